#!/bin/bash
# one GPU call: full GPU suite, bench (both arms), ncu launch list + full capture of the int8 Gram kernels.  Outputs under gpurun_out/$1_*
T=${1:-r}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2>> gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 18 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_i8 -s 6 -c 3 -f -o gpurun_out/${T}_prof_i8 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cat gpurun_out/${T}_tests.log; cut -c1-400 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
