#!/bin/bash
# round-2 evidence run on one B200: GPU suite, smoke, bench (both arms), all BASELINE configs, stress, sanitizer passes
T=r02j
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -s KILL 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json
timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; cut -c1-300 gpurun_out/${T}_ref.json
timeout -s KILL 900 python tools/bench_configs.py > gpurun_out/${T}_all_configs.jsonl 2> gpurun_out/${T}_all_configs.err; cut -c1-260 gpurun_out/${T}_all_configs.jsonl
timeout -s KILL 120 python tools/stress_cfg3.py 2500 > gpurun_out/${T}_stress.log 2>&1; echo "stress exit $?"; tail -2 gpurun_out/${T}_stress.log
{
echo "== memcheck: fused small-instance kernel through the planner (tests/test_gpu_gram_fused.py, all 14 tests)"
timeout -s KILL 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_gram_fused.py -x -q 2>&1 | grep -E "=========|passed|failed" | tail -4
echo "== memcheck: parity / export / edge / result-path / golden suites (small models)"
timeout -s KILL 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_export.py tests/test_gpu_edge.py tests/test_gpu_result_path.py tests/test_golden.py -m gpu -x -q 2>&1 | grep -E "=========|passed|failed" | tail -4
echo "== racecheck, shipped build: fused kernel (ragged shape 100 x 36 and the certificate test)"
timeout -s KILL 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_gram_fused.py -x -q -k "shape4 or certificate" 2>&1 | grep -E "=========|passed|failed" | grep -v "Host Frame\|Saved host" | tail -12
cp parametron.jl_b200/lib/libpmb200.so /tmp/lib_orig.so; cp alt_libs/libpmb200_strict.so parametron.jl_b200/lib/libpmb200.so
echo "== racecheck, -DFZ_STRICT_SYNC build (release arrivals instead of relaxed): same tests"
timeout -s KILL 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_gram_fused.py -x -q -k "shape4 or certificate" 2>&1 | grep -E "=========|passed|failed" | grep -v "Host Frame\|Saved host" | tail -12
echo "== racecheck, -DFZ_STRICT_SYNC build: k = 256 shapes only (no scratch restore)"
timeout -s KILL 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_gram_fused.py -x -q -k "shape0 or shape2 or shape5" 2>&1 | grep -E "=========|passed|failed" | grep -v "Host Frame\|Saved host" | tail -12
cp /tmp/lib_orig.so parametron.jl_b200/lib/libpmb200.so
echo "== racecheck: two-kernel int8 route (n = 2048 forced planes) and fp64 route kernels"
timeout -s KILL 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_gram_i8.py -x -q -k "bit_identical_to_the_numpy" 2>&1 | grep -E "=========|passed|failed" | grep -v "Host Frame\|Saved host" | tail -8
} > gpurun_out/${T}_sanitizer.txt 2>&1
tail -40 gpurun_out/${T}_sanitizer.txt
