#!/bin/bash
# timing diagnostics of the one-CTA int8 Gram kernel: which resource bounds the main loop (results are wrong by construction)
# PMB_I8_DBG bits: 1 no TMA loads, 2 no MMAs, 4 only the s=0 row, 8 no epilogue work, 16 s=0 row three times (N=240 instructions only)
mkdir -p gpurun_out
P=tools/probe/i8_probe
{
for d in 0 1 2 3 8 9 11 16 17 25; do
  echo "=== PMB_I8_DBG=$d"
  PMB_I8_DBG=$d timeout 60 $P 4096 4096 1 5 0 4 | grep -E "^run [123]"
done
} > gpurun_out/i8_dbg2.txt 2>&1
cat gpurun_out/i8_dbg2.txt
