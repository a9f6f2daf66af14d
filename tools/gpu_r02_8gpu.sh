#!/bin/bash
# 8-GPU run: concurrent PCIe probe (host ceiling of the e2e path) + bench under torchrun at N = 8 (cfg2 replicas, cfg3 instance-shard over 8 ranks)
T=r02q
N=${1:-8}
mkdir -p gpurun_out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29507 tools/pcie_probe_multi.py > gpurun_out/${T}_pcie${N}.json 2> gpurun_out/${T}_pcie${N}.err; cut -c1-600 gpurun_out/${T}_pcie${N}.json; tail -n 2 gpurun_out/${T}_pcie${N}.err
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/${T}_bench${N}.json 2> gpurun_out/${T}_bench${N}.err; tail -n 3 gpurun_out/${T}_bench${N}.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d=json.loads(open(f'gpurun_out/r02q_bench{n}.json').read().strip().splitlines()[-1])
print('value', d['value'], 'n_gpus', d['n_gpus'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'))
c=d.get('cfg3_shard') or {}
print('cfg3', c.get('value'), c.get('ms_per_sweep'), c.get('instances_on_rank0'))
PY
