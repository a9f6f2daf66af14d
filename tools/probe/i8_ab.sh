#!/bin/bash
# A/B of experimental builds of the int8 Gram probe (tools/probe/i8_probe_*): correctness on small shapes, timing at cfg2
mkdir -p gpurun_out
{
for P in tools/probe/i8_probe_*; do
  echo "=== $P"
  timeout 60 $P 520 1000 2 0 1 1 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 60 $P 300 500 1 6 0 1 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 4096 4096 1 0 0 8 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 4096 4096 1 6 0 4 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 256 256 1024 5 0 3 | grep -E "PROBE|^run"; echo "exit $?"
done
} > gpurun_out/i8_ab.txt 2>&1
cat gpurun_out/i8_ab.txt
