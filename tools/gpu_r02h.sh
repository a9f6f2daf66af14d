#!/bin/bash
T=r02h
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_gram_fused.py tests/test_gpu_gram_i8.py -x -q 2>&1 | tail -8 > gpurun_out/${T}_t1.log; cat gpurun_out/${T}_t1.log
timeout -s KILL 300 python tools/bench_configs.py cfg3 > gpurun_out/${T}_cfg3.jsonl 2> gpurun_out/${T}_cfg3.err; cat gpurun_out/${T}_cfg3.jsonl; tail -n 5 gpurun_out/${T}_cfg3.err
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -n 5 gpurun_out/${T}_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
for k in ('value','ms_per_step','kernel_ms_per_update','sustained','fp64_route','cfg3_shard'):
    print(k, json.dumps(d.get(k))[:900])
print('e2e', json.dumps(d.get('e2e'))[:300])
print('cpu keys', list((d.get('cpu_baseline') or {}).keys()))
print(json.dumps((d.get('cpu_baseline') or {}).get('termlist_full_runs_cubic_fit'))[:600])
print(json.dumps((d.get('cpu_baseline') or {}).get('dense_blas_every_config'))[:900])
PY
