#!/bin/bash
# quick check of the int8 Gram probe (correctness on small shapes, timing at cfg2 / 6144 / 6 planes) + the i8 GPU tests
mkdir -p gpurun_out
P=tools/probe/i8_probe
{
timeout 30 $P 520 1000 2 5 1 1 | grep -E "PROBE|^run"; echo "exit $?"
timeout 30 $P 300 500 1 6 0 1 | grep -E "PROBE|^run"; echo "exit $?"
timeout 120 $P 4096 4096 1 5 0 6 | grep -E "PROBE|^run"; echo "exit $?"
timeout 120 $P 4096 4096 1 6 0 4 | grep -E "PROBE|^run"; echo "exit $?"
timeout 120 $P 6144 4096 1 5 0 4 | grep -E "PROBE|^run"; echo "exit $?"
timeout 120 $P 256 256 1024 5 0 3 | grep -E "PROBE|^run"; echo "exit $?"
PMB_I8_DBG=8 timeout 60 $P 4096 4096 1 5 0 4 | grep -E "^run"
timeout 600 python -m pytest tests/test_gpu_gram_i8.py -x -q 2>&1 | tail -5
} > gpurun_out/i8_quick.txt 2>&1
cat gpurun_out/i8_quick.txt
