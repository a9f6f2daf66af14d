#!/bin/bash
# ncu duration of the affine-assembly launch at cfg2 for a few ring configurations
for cfg in "0 0" "6 2" "4 3" "3 4" "2 5" "3 3" "5 2" "2 4"; do
  set -- $cfg
  PMB_AFF_STAGES=$1 PMB_AFF_CTAS_PER_SM=$2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:affine_assemble_tma -s 2 -c 3 --csv \
     python tools/kbench.py 4096 4096 1024 1 3 2>/dev/null | grep -E "affine_assemble_tma" | awk -F, -v c="$cfg" '{printf "stages/ctas_per_sm=%s %s\n", c, $(NF)}'
done
