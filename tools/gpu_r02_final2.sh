#!/bin/bash
# round-2 final evidence run (generator fusion + prologue split on): GPU suite, smoke, bench (both arms), all configs, stress, ncu
T=r02l
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -s KILL 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; cut -c1-200 gpurun_out/${T}_ref.json
timeout -s KILL 900 python tools/bench_configs.py > gpurun_out/${T}_all_configs.jsonl 2> gpurun_out/${T}_all_configs.err; cut -c1-200 gpurun_out/${T}_all_configs.jsonl
timeout -s KILL 200 python tools/stress_cfg3.py 10000 > gpurun_out/${T}_stress.log 2>&1; echo "stress exit $?"; tail -2 gpurun_out/${T}_stress.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 10 --csv --log-file gpurun_out/${T}_launches_cfg3.csv python tools/bench_configs.py cfg3 > /dev/null 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gram_i8_fused_gen -s 3 -c 1 -f -o gpurun_out/${T}_prof_fused_gen_cfg3 python tools/bench_configs.py cfg3 > /dev/null 2> gpurun_out/${T}_ncu1.err
timeout -s KILL 900 ncu --set full --clock-control none -k regex:gram_i8_slice -s 6 -c 1 -f -o gpurun_out/${T}_prof_slice_gen_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2> gpurun_out/${T}_ncu2.err
(timeout -s KILL 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_gram_fused.py tests/test_gpu_scale.py -m gpu -x -q -k "ragged or certificate or permuted or generator_fused or prologue_split or shard or back_to_back" 2>&1 | grep -E "=========|passed|failed" | tail -4) > gpurun_out/${T}_memcheck.txt; cat gpurun_out/${T}_memcheck.txt
ls -la gpurun_out | grep ${T}
