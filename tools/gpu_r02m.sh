#!/bin/bash
T=r02m
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -n 3 gpurun_out/${T}_bench.err; cut -c1-600 gpurun_out/${T}_bench.json
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2>> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_ref.json
