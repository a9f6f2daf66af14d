#!/bin/bash
T=r02w
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_gram_i8.py tests/test_gpu_gram_fused.py -x -q 2>&1 | tail -6 > gpurun_out/${T}_t1.log; cat gpurun_out/${T}_t1.log
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -n 3 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02w_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step']); print(json.dumps(d['kernel_ms_per_update'])); print('roofline frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
print('sustained', d['sustained']['value'], 'cfg3', d['cfg3_shard']['value'], d['cfg3_shard']['ms_per_sweep'])
PY
timeout -s KILL 200 python tools/sweep_gram.py 2048:2048 4096:4096 6144:6144 8192:8192 2>/dev/null | cut -c1-400
