"""All five BASELINE.json configs through the C ABI on one GPU (or one instance-shard per rank under torchrun for cfg3).
Prints one JSON object per config.  bench.py remains the driver contract (cfg2); this is the per-config evidence.

  python tools/bench_configs.py [cfg1 cfg2 cfg3 cfg4 cfg5]
  python -m torch.distributed.run --nproc-per-node N tools/bench_configs.py cfg3        (instance-shard sweep)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parametron_b200 as pm  # noqa: E402

RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
HBM_PEAK = 6455.6   # MEASURED_PEAKS.json (hbm_gbs)
I8_PEAK = 2 * 1677.0  # 2 x bf16_tflops (burst) of MEASURED_PEAKS.json: int8 dense rate of tcgen05 = 2 x bf16
DMMA_PEAK = 37.1    # profiles/r01_probe_dmma.txt


def ls_model(n, k, m, batch=1, fixed_matrices=False, gram_mode=None, instance_offset=0):
    model = pm.Model(device=LOCAL, batch=batch, gram_mode=gram_mode, instance_offset=instance_offset)
    x = model.variables(n)
    if fixed_matrices:   # MPC style: A, C resident and constant between solves; b, d re-parameterised every solve
        A = pm.Parameter(model, shape=(k, n), device=True)
        Cm = pm.Parameter(model, shape=(m, n), device=True)
    else:
        A = pm.Parameter(model, shape=(k, n), uniform=(1234, 1.0))
        Cm = pm.Parameter(model, shape=(m, n), uniform=(1236, 1.0))
    b = pm.Parameter(model, shape=(k,), uniform=(1235, 1.0))
    d = pm.Parameter(model, shape=(m,), uniform=(1237, 2.0))
    r = A * x - b
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, Cm * x, "==", d)
    model.initialize()
    if fixed_matrices:
        rng = np.random.default_rng(0)
        A.upload(rng.random((k, n)) if batch == 1 else rng.random((batch, k, n)))
        Cm.upload(rng.random((m, n)) if batch == 1 else rng.random((batch, m, n)))
    return model


def timed_updates(model, steps, warmup=5):
    for _ in range(warmup):
        model.update()
    model.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        model.update()
    model.synchronize()
    return (time.perf_counter() - t0) / steps


def kernel_breakdown(model, steps):
    model.enable_kernel_timing(True)
    for _ in range(steps):
        model.update()
    kt = model.kernel_times()
    model.enable_kernel_timing(False)
    return {k: v[0] / steps * 1e3 for k, v in kt.items() if v[1]}  # us per update


def affine_record(nbytes, us, rows):
    """GB/s of the affine copy kernel -- or a note when the block aliases the Parameter's buffer (zero-copy: only the offsets are assembled)"""
    if nbytes <= 16 * rows:
        return {"affine": "zero-copy block (the view aliases the Parameter's buffer; offsets only)", "affine_offsets_us": us}
    return {"affine_gbs": nbytes / us / 1e3, "affine_frac_of_hbm_peak": nbytes / us / 1e3 / HBM_PEAK}


def report(name, **kw):
    if RANK == 0:
        print(json.dumps({"config": name, **kw}), flush=True)


def allreduce_max(x):
    if WORLD == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{LOCAL}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cfg1():
    m = ls_model(8, 8, 2)
    sec = timed_updates(m, 2000)
    report("cfg1 README Example 1: n=8, m=2", us_per_update=sec * 1e6, updates_per_s=1 / sec, launches=m.stats()["kernel_launches_per_update"],
           note="latency-bound plumbing case; the reference README reports 51.9 us per solve! including the OSQP solve")


def cfg2():
    for mode, label in ((1, "certified int8 digit planes on tcgen05 (default)"), (0, "fp64 DMMA only (pmb_set_gram_mode 0)")):
        m = ls_model(4096, 4096, 1024, gram_mode=mode)
        st = m.stats()
        sec = timed_updates(m, 30)
        kb = kernel_breakdown(m, 10)
        lv = m.gram_level_counts()
        extra = {}
        syrk = 4096 * 4096 * 4097
        if lv[0] + lv[1]:
            planes = 5 if lv[0] else 6
            nprod = planes * (planes + 1) // 2
            extra = dict(gram_planes=planes, gram_int8_tops=nprod * syrk / kb["gram_i8_tcgen05"] / 1e6,
                         gram_frac_of_int8_peak=nprod * syrk / kb["gram_i8_tcgen05"] / 1e6 / I8_PEAK,
                         gram_fp64_equivalent_tflops=syrk / (kb["gram_i8_tcgen05"] + kb["gram_i8_slice"]) / 1e6)
        else:
            extra = dict(gram_tflops=st["gram_algorithmic_flops"] / kb["gram_dmma"] / 1e6,
                         gram_frac_of_dmma_peak=st["gram_algorithmic_flops"] / kb["gram_dmma"] / 1e6 / DMMA_PEAK)
        report(f"cfg2 n=4096, m=1024 dense -- {label}", ms_per_update=sec * 1e3, updates_per_s=1 / sec, kernel_us=kb, gram_levels=lv,
               **affine_record(st["affine_algorithmic_bytes"], kb["affine_assemble"], 1024), **extra)
        m.close()


def cfg3(total=8192):
    first, batch = pm.shard.instance_shard(total, RANK, WORLD)
    m = ls_model(256, 256, 64, batch=batch, instance_offset=first)
    st = m.stats()
    sec = allreduce_max(timed_updates(m, 10))
    kb = kernel_breakdown(m, 5)
    lv = m.gram_level_counts()
    extra = {}
    if kb.get("gram_i8_tcgen05"):   # fused small-instance kernel: HBM-bound by design (reads A once, writes Q once)
        gbytes = (8 * 256 * 256 + 8 * 256 * 256) * batch
        extra = dict(gram_fused_gbs_per_gpu=gbytes / kb["gram_i8_tcgen05"] / 1e3, gram_fused_frac_of_hbm_peak=gbytes / kb["gram_i8_tcgen05"] / 1e3 / HBM_PEAK,
                     gram_fp64_equivalent_tflops=st["gram_algorithmic_flops"] * batch / kb["gram_i8_tcgen05"] / 1e6)
    if kb.get("gram_dmma", 0) > 50:
        extra.update(gram_tflops_per_gpu=st["gram_algorithmic_flops"] * batch / kb["gram_dmma"] / 1e6,
                     gram_frac_of_dmma_peak=st["gram_algorithmic_flops"] * batch / kb["gram_dmma"] / 1e6 / DMMA_PEAK)
    report(f"cfg3 batch={total} x (n=256, m=64), {WORLD} GPU(s), {batch} instances per GPU", ms_per_update=sec * 1e3,
           instance_updates_per_s=total / sec, kernel_us_rank0=kb, gram_levels=lv,
           **affine_record(st["affine_algorithmic_bytes"] * batch, kb["affine_assemble"], 64 * batch), **extra)


def cfg4():
    rows, cols = 4096, 16384
    for b in (16, 32):
        rng = np.random.default_rng(b)
        nbr, nbc = rows // b, cols // b
        nblocks = int(round(0.01 * nbr * nbc))
        flat = rng.choice(nbr * nbc, size=nblocks, replace=False)
        br, bc = (flat // nbc).astype(np.int64), (flat % nbc).astype(np.int64)
        model = pm.Model(device=LOCAL)
        x = model.variables(cols)
        Cp = pm.Parameter(model, uniform=(7, 1.0), bsr=(rows, cols, b, br, bc))
        d = pm.Parameter(model, shape=(rows,), uniform=(8, 2.0))
        pm.constraint(model, Cp * x, "==", d)
        model.initialize()
        st = model.stats()
        sec = timed_updates(model, 500)
        kb = kernel_breakdown(model, 50)
        report(f"cfg4 n=16384, m=4096, 1% block-sparse C (b={b}), affine-only", us_per_update=sec * 1e6, updates_per_s=1 / sec,
               stored_values=nblocks * b * b, algorithmic_bytes=st["affine_algorithmic_bytes"], kernel_us=kb,
               affine_gbs=st["affine_algorithmic_bytes"] / kb["affine_assemble"] / 1e3,
               note="launch-latency-bound (10.8 MB per update): reported, not graded against the 70 % HBM target (SURVEY 8d)")
        model.close()


def cfg5(solves=10000):
    m = ls_model(512, 512, 256, fixed_matrices=True)
    sec = timed_updates(m, solves, warmup=20)
    kb = kernel_breakdown(m, 200)
    report("cfg5 MPC-style n=512, m=256: 10k sequential re-parameterisations, device-resident (Q,q,C,d)", us_per_update=sec * 1e6,
           updates_per_s=1 / sec, solves=solves, launches=m.stats()["kernel_launches_per_update"], cuda_graph=bool(m.stats()["uses_cuda_graph"]),
           kernel_us=kb, note="A, C resident and fixed; b, d regenerated on device every solve; zero host<->device traffic per update")


if __name__ == "__main__":
    if WORLD > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(LOCAL)
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    for w in which:
        if WORLD > 1 and w != "cfg3":
            continue
        globals()[w]()
    if WORLD > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
