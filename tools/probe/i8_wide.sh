#!/bin/bash
# 128-byte-row split-ring variant of the int8 Gram kernel against the 64-byte one: correctness, timing, TMA-only / MMA-only diagnostics
mkdir -p gpurun_out
P=tools/probe/i8_probe
{
for kc in 128 64; do
  export PMB_I8_KC=$kc
  echo "=== PMB_I8_KC=$kc"
  timeout 30 $P 128 64 1 5 0 1 | grep -E "PROBE|^run|mismatch"; echo "exit $?"
  timeout 30 $P 520 1000 2 5 1 1 | grep -E "PROBE|^run|mismatch" | head -5; echo "exit $?"
  timeout 30 $P 300 500 1 6 0 1 | grep -E "PROBE|^run|mismatch" | head -5; echo "exit $?"
  timeout 60 $P 520 1000 3 0 2 1 | grep -E "PROBE|^run|mismatch" | head -5; echo "exit $?"
  timeout 120 $P 6144 4096 1 5 0 4 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 4096 4096 1 5 0 6 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 4096 4096 1 6 0 3 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 2048 2048 1 5 0 3 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 256 256 1024 5 0 3 | grep -E "PROBE|^run"; echo "exit $?"
  for d in 1 2; do
    echo "--- PMB_I8_DBG=$d"
    PMB_I8_DBG=$d timeout 60 $P 4096 4096 1 5 0 3 | grep -E "^run"
  done
done
} > gpurun_out/i8_wide.txt 2>&1
cat gpurun_out/i8_wide.txt
