#!/bin/bash
# r02_a: baseline of cfg3 (8192 x n=256, m=64) on the round-1 kernels + ncu capture of the 64-tile fp64 Gram there
T=r02a
mkdir -p gpurun_out
nproc > gpurun_out/${T}_host.txt; lscpu | head -20 >> gpurun_out/${T}_host.txt; nvidia-smi topo -m >> gpurun_out/${T}_host.txt 2>&1
python tools/bench_configs.py cfg3 > gpurun_out/${T}_cfg3.jsonl 2> gpurun_out/${T}_cfg3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tma -s 3 -c 1 -f -o gpurun_out/${T}_prof_gram_tma_cfg3 python tools/bench_configs.py cfg3 > /dev/null 2> gpurun_out/${T}_ncu.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 6 --csv --log-file gpurun_out/${T}_launches_cfg3.csv python tools/bench_configs.py cfg3 > /dev/null 2>&1
cat gpurun_out/${T}_cfg3.jsonl; tail -3 gpurun_out/${T}_cfg3.err gpurun_out/${T}_ncu.err
