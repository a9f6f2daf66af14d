"""Host ceiling of the e2e arm across GPUs: every rank copies cfg2's per-update traffic (168 MB H2D + 100 MB D2H, pinned,
both directions at once) at the same time.  Launched with torchrun, one rank per GPU; prints the per-rank and aggregate GB/s, the CPU
affinity / NUMA node of every GPU, and the same with each rank pinned to its own slice of the GPU's affine cores.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def pin_to_gpu_cores(local_rank, world):
    """disjoint slice of the cores NVML lists as affine to this rank's GPU; returns the core list (None when NVML is unavailable)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cores = [c for c in cores if c in os.sched_getaffinity(0)] or sorted(os.sched_getaffinity(0))
        # ranks whose GPUs share the same core set split it evenly
        per = max(1, len(cores) // world)
        mine = cores[(local_rank * per) % len(cores):][:per] or cores
        os.sched_setaffinity(0, mine)
        return mine
    except Exception as exc:   # noqa: BLE001
        return None


def measure(tag, local_rank, world):
    n_up, n_dn = 168_000_000 // 8, 100_000_000 // 8
    hu = torch.empty(n_up, dtype=torch.float64).pin_memory(); du = torch.empty(n_up, dtype=torch.float64, device="cuda")
    hd = torch.empty(n_dn, dtype=torch.float64).pin_memory(); dd = torch.empty(n_dn, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for _ in range(6):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            du.copy_(hu, non_blocking=True)
        with torch.cuda.stream(s2):
            hd.copy_(dd, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    t = torch.tensor([best], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = float(t.item())
    return {"tag": tag, "ms_this_rank": best * 1e3, "ms_slowest_rank": worst * 1e3, "aggregate_gbs": world * 0.268 / worst,
            "per_rank_gbs": 0.268 / worst}


if __name__ == "__main__":
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    before = sorted(os.sched_getaffinity(0))
    a = measure("unpinned", local, world)
    mine = pin_to_gpu_cores(local, world)
    b = measure("pinned to a slice of the GPU's affine cores", local, world)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "host_cores": os.cpu_count(), "affinity_before": [before[0], before[-1]], "rank0_cores_after_pinning": mine,
                          "results": [a, b]}), flush=True)
    if world > 1:
        dist.destroy_process_group()
