#!/bin/bash
# quick check of the int8 Gram probe (correctness on small shapes, timing at cfg2 / other full-column sizes) + the i8 GPU tests
mkdir -p gpurun_out
P=tools/probe/i8_probe
{
timeout 30 $P 520 1000 2 5 1 1 | grep -E "PROBE|^run"; echo "exit $?"
timeout 60 $P 600 2048 2 0 1 2 | grep -E "PROBE|^run"; echo "exit $?"
timeout 60 $P 300 3072 1 6 2 2 | grep -E "PROBE|^run"; echo "exit $?"
timeout 120 $P 4096 4096 1 5 0 6 | grep -E "PROBE|^run"; echo "exit $?"
echo "--- PMB_I8_SLICE_GENERIC=1"
PMB_I8_SLICE_GENERIC=1 timeout 120 $P 4096 4096 1 5 0 6 | grep -E "PROBE|^run"; echo "exit $?"
timeout 120 $P 4096 4096 1 5 1 3 | grep -E "PROBE|^run"; echo "exit $?"
timeout 600 python -m pytest tests/test_gpu_gram_i8.py tests/test_gpu_scale.py -x -q 2>&1 | tail -5
} > gpurun_out/i8_quick.txt 2>&1
cat gpurun_out/i8_quick.txt
