#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "generator_fused or cfg2_full or device_generator or idempotent" 2>&1 | tail -15
timeout -s KILL 300 python bench.py --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:round(v*1e3,1) for k,v in d['kernel_ms_per_update'].items() if v}, d['clocks']['sm_mhz'])"
