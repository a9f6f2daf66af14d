#!/bin/bash
# r02_z: ncu evidence of the round-2 kernels + sanitizer passes
T=r02z
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file gpurun_out/${T}_launches_cfg3.csv python tools/bench_configs.py cfg3 > /dev/null 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gram_i8_fused -s 3 -c 1 -f -o gpurun_out/${T}_prof_fused_cfg3 python tools/bench_configs.py cfg3 > /dev/null 2> gpurun_out/${T}_ncu1.err
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gram_i8_kernel -s 6 -c 1 -f -o gpurun_out/${T}_prof_i8_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2> gpurun_out/${T}_ncu2.err
timeout -s KILL 900 ncu --set full --clock-control none -k regex:gram_i8_slice -s 6 -c 1 -f -o gpurun_out/${T}_prof_slice_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>> gpurun_out/${T}_ncu2.err
timeout -s KILL 900 ncu --set full --clock-control none -k regex:prologue -s 6 -c 1 -f -o gpurun_out/${T}_prof_prologue_cfg3 python tools/bench_configs.py cfg3 > /dev/null 2>> gpurun_out/${T}_ncu1.err
(timeout -s KILL 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_gram_fused.py -x -q -k "ragged or certificate or permuted" 2>&1 | tail -6) > gpurun_out/${T}_memcheck_fused.txt
(timeout -s KILL 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_gram_fused.py -x -q -k "shape3 or shape4 or certificate" 2>&1 | tail -8) > gpurun_out/${T}_racecheck_fused.txt
(timeout -s KILL 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_gram_i8.py -x -q -k "forced_planes and shape0" 2>&1 | tail -8) > gpurun_out/${T}_racecheck_i8.txt
cat gpurun_out/${T}_memcheck_fused.txt gpurun_out/${T}_racecheck_fused.txt gpurun_out/${T}_racecheck_i8.txt; ls -la gpurun_out | grep ${T}
