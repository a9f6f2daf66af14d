#!/bin/bash
# CTA-pair (cta_group::2) and 128-byte split-ring variants of the int8 Gram kernel against the default: correctness on small shapes under a short timeout
# (a hang costs 30 s), then timing; last, the timing diagnostics (PMB_I8_DBG, wrong results by construction) of every variant
mkdir -p gpurun_out
P=tools/probe/i8_probe
{
for cfg in "0 64" "1 64" "0 128"; do
  set -- $cfg
  export PMB_I8_PAIR=$1 PMB_I8_KC=$2
  echo "=== PMB_I8_PAIR=$1 PMB_I8_KC=$2"
  timeout 30 $P 128 64 1 5 0 1 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 30 $P 520 1000 2 5 1 1 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 30 $P 300 500 1 6 0 1 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 6144 4096 1 5 0 4 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 4096 4096 1 5 0 5 | grep -E "PROBE|^run"; echo "exit $?"
  timeout 120 $P 4096 4096 1 6 0 3 | grep -E "PROBE|^run"; echo "exit $?"
  for d in 1 2; do
    echo "--- PMB_I8_DBG=$d"
    PMB_I8_DBG=$d timeout 60 $P 6144 4096 1 5 0 3 | grep -E "^run"
  done
done
} > gpurun_out/i8_pair.txt 2>&1
cat gpurun_out/i8_pair.txt
