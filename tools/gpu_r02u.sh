#!/bin/bash
T=r02u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_gram_fused.py tests/test_gpu_gram_i8.py tests/test_gpu_scale.py -x -q 2>&1 | tail -8 > gpurun_out/${T}_t1.log; cat gpurun_out/${T}_t1.log
timeout -s KILL 300 python tools/bench_configs.py cfg3 > gpurun_out/${T}_cfg3.jsonl 2> gpurun_out/${T}_cfg3.err; cat gpurun_out/${T}_cfg3.jsonl; tail -n 5 gpurun_out/${T}_cfg3.err
