#!/bin/bash
# sustained (power-capped) timing of the 5-plane Gram variants: 4000 back-to-back slice + Gram pairs (~1.6 s), last runs reported
mkdir -p gpurun_out
P=tools/probe/i8_probe
{
for kc in 64 128 64 128; do
  echo "=== PMB_I8_KC=$kc (5 planes, n = k = 4096)"
  PMB_I8_KC=$kc timeout 120 $P 4096 4096 1 5 0 4000 | grep -E "^run (5|50|500|1000|2000|3000|399[0-9]):"
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
done
} > gpurun_out/i8_sustained.txt 2>&1
cat gpurun_out/i8_sustained.txt
