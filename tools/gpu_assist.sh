#!/bin/bash
T=r02assist
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_gram_fused.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -6
for i in 1 2; do timeout -s KILL 300 python tools/bench_configs.py cfg3 2>&1 | tail -1 | cut -c80-330; done
timeout -s KILL 150 python tools/stress_cfg3.py 8000 > gpurun_out/${T}_stress.log 2>&1; echo "stress exit $?"; tail -2 gpurun_out/${T}_stress.log
