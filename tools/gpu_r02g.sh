#!/bin/bash
T=r02g
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_gram_fused.py -x -q 2>&1 | tail -25 > gpurun_out/${T}_t1.log; cat gpurun_out/${T}_t1.log
timeout -s KILL 900 python -m pytest tests/test_gpu_scale.py -x -q -k "shard or cfg3_full or cfg5_ten or elementwise" 2>&1 | tail -25 > gpurun_out/${T}_t2.log; cat gpurun_out/${T}_t2.log
timeout -s KILL 300 python tools/bench_configs.py cfg3 > gpurun_out/${T}_cfg3.jsonl 2> gpurun_out/${T}_cfg3.err; cat gpurun_out/${T}_cfg3.jsonl; tail -n 5 gpurun_out/${T}_cfg3.err
