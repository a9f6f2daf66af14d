#!/bin/bash
# 2-GPU run: bench under torchrun (cfg3 instance-shard across 2 ranks, e2e with host pinning) + the concurrent PCIe probe
T=r02y
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29507 tools/pcie_probe_multi.py > gpurun_out/${T}_pcie2.json 2> gpurun_out/${T}_pcie2.err; cat gpurun_out/${T}_pcie2.json; tail -n 2 gpurun_out/${T}_pcie2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err; tail -n 3 gpurun_out/${T}_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02y_bench2.json'))
print('value', d['value'], 'n_gpus', d['n_gpus'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('cfg3', d['cfg3_shard']['value'], d['cfg3_shard']['ms_per_sweep'], d['cfg3_shard']['instances_on_rank0'])
print('kernel ms', json.dumps(d['kernel_ms_per_update']), 'sum', sum(d['kernel_ms_per_update'].values()), 'ms/step', d['ms_per_step'])
PY
