"""Objective-only sweep over problem sizes: ms per update of `min ||Ax - b||^2` (A k x n regenerated on the device every update) on the
certified int8 route (default) and on the fp64 route (gram_mode=0), with the level the certificate picked and the per-kernel split.
usage: python tools/sweep_gram.py [n:k ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parametron_b200 as pm  # noqa: E402


def run(n, k, mode, steps=20):
    model = pm.Model(gram_mode=mode)
    x = model.variables(n)
    A = pm.Parameter(model, shape=(k, n), uniform=(1, 1.0))
    b = pm.Parameter(model, shape=(k,), uniform=(2, 1.0))
    r = A * x - b
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    model.initialize()
    for _ in range(3):
        model.update()
    model.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        model.update()
    model.synchronize()
    sec = (time.perf_counter() - t0) / steps
    model.enable_kernel_timing(True)
    for _ in range(5):
        model.update()
    kt = {k_: v[0] / 5 * 1e3 for k_, v in model.kernel_times().items() if v[1]}
    lv = model.gram_level_counts()
    model.close()
    return sec, kt, lv


if __name__ == "__main__":
    sizes = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(1024, 1024), (2048, 2048), (3000, 3000), (4096, 4096), (4096, 16384), (6144, 6144), (8192, 8192)]
    for n, k in sizes:
        row = {"n": n, "k": k}
        for mode, name in ((1, "int8"), (0, "fp64")):
            sec, kt, lv = run(n, k, mode)
            row[name] = {"ms_per_update": sec * 1e3, "levels_5_6_fp64": lv, "kernel_us": kt}
        row["speedup"] = row["fp64"]["ms_per_update"] / row["int8"]["ms_per_update"]
        print(json.dumps(row), flush=True)
