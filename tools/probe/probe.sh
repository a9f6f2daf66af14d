#!/bin/bash
# First GPU probe: DMMA shapes, cuBLAS fp64 peaks, ncu metric names, host facts.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/probe_nvsmi.txt 2>&1
lscpu > gpurun_out/probe_lscpu.txt 2>&1; nproc >> gpurun_out/probe_lscpu.txt; free -g >> gpurun_out/probe_lscpu.txt
./tools/probe/dmma_probe > gpurun_out/probe_dmma.txt 2>&1
python tools/probe/torch_fp64.py > gpurun_out/probe_torch_fp64.txt 2>&1
ncu --query-metrics 2>/dev/null | grep -i -E "fp64|dmma|pipe_tensor|dram__bytes|dram__throughput" > gpurun_out/probe_ncu_metrics.txt
which julia go node javac > gpurun_out/probe_toolchains.txt 2>&1
cat gpurun_out/probe_dmma.txt gpurun_out/probe_torch_fp64.txt
