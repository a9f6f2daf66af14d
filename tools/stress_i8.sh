#!/bin/bash
# whole GPU suite with every SYRK-shaped step forced through the int8 route (PMB_GRAM_I8_ALWAYS=1), once per tcgen05 pipeline variant
mkdir -p gpurun_out
{
for v in "" "PMB_I8_PAIR=1" "PMB_I8_KC=128"; do
  echo "=== PMB_GRAM_I8_ALWAYS=1 $v"
  env PMB_GRAM_I8_ALWAYS=1 $v timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
done
} > gpurun_out/r01o_stress_i8.txt 2>&1
cat gpurun_out/r01o_stress_i8.txt
