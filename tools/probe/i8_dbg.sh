#!/bin/bash
# timing diagnostics of the one-CTA int8 Gram kernel: which resource bounds the main loop (results are wrong by construction)
mkdir -p gpurun_out
P=tools/probe/i8_probe
{
for d in 0 1 2 3 4 5; do
  echo "=== PMB_I8_DBG=$d (1 no TMA, 2 no MMA, 4 only s=0 row)"
  PMB_I8_DBG=$d timeout 60 $P 4096 4096 1 5 0 4 | grep -E "^run"; echo "exit $?"
done
} > gpurun_out/i8_dbg.txt 2>&1
cat gpurun_out/i8_dbg.txt
