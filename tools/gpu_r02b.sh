#!/bin/bash
# r02_b: first run of the fused small-instance kernel: parity tests (bounded by timeout: a protocol bug would hang), then cfg3 timing
T=r02b
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_gram_fused.py -x -q -k "cfg3_instance_shape" 2>&1 | tail -25 > gpurun_out/${T}_t1.log
cat gpurun_out/${T}_t1.log
if grep -q "passed" gpurun_out/${T}_t1.log && ! grep -q "failed" gpurun_out/${T}_t1.log; then
  timeout -s KILL 600 python -m pytest tests/test_gpu_gram_fused.py -q 2>&1 | tail -40 > gpurun_out/${T}_t2.log
  cat gpurun_out/${T}_t2.log
  timeout -s KILL 300 python tools/bench_configs.py cfg3 > gpurun_out/${T}_cfg3.jsonl 2> gpurun_out/${T}_cfg3.err
  cat gpurun_out/${T}_cfg3.jsonl; tail -n 5 gpurun_out/${T}_cfg3.err
fi
