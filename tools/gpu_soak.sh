#!/bin/bash
# soak of the shipped build: long back-to-back update streams on every route + the new tests
T=r02soak
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "prologue_split or back_to_back" 2>&1 | tail -3
timeout -s KILL 400 python tools/stress_cfg3.py 30000 > gpurun_out/${T}_cfg3.log 2>&1; echo "cfg3 30000 updates: exit $?"; tail -2 gpurun_out/${T}_cfg3.log
timeout -s KILL 300 python - <<'PY' 2>&1 | tail -4
import sys, time
sys.path.insert(0, ".")
import numpy as np
import tools.bench_configs as B
m = B.ls_model(4096, 4096, 1024)            # cfg2, two-kernel int8 route, device generators
t0 = time.perf_counter()
for i in range(0, 10000, 1000):
    for _ in range(1000):
        m.update()
    m.synchronize()
print("cfg2 device-resident: 10000 updates", round(time.perf_counter() - t0, 2), "s, levels", m.gram_level_counts(), flush=True)
m.close()
PY
timeout -s KILL 600 python bench.py --no-cpu-baseline --no-extra --steps 1000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench 1000 steps', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('steps'), d['clocks'])"
