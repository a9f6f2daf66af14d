"""Per-kernel-class timing of one model shape through the C ABI (CUDA events on the context stream, eager replay).
usage: python tools/kbench.py [n k m batch steps]   (env PMB_NO_TMA=1 / PMB_NO_GRAPH=1 select code paths)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parametron_b200 as pm  # noqa: E402


def main():
    a = [int(x) for x in sys.argv[1:]]
    n, k, m, batch, steps = (a + [4096, 4096, 1024, 1, 20][len(a):])[:5]
    model = pm.Model(batch=batch)
    x = model.variables(n)
    A = pm.Parameter(model, shape=(k, n), uniform=(1, 1.0))
    b = pm.Parameter(model, shape=(k,), uniform=(2, 1.0))
    Cm = pm.Parameter(model, shape=(m, n), uniform=(3, 1.0))
    d = pm.Parameter(model, shape=(m,), uniform=(4, 2.0))
    r = A * x - b
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, Cm * x, "==", d)
    model.initialize()
    st = model.stats()
    for _ in range(3):
        model.update()
    model.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        model.update()
    model.synchronize()
    wall = (time.perf_counter() - t0) / steps
    model.enable_kernel_timing(True)
    for _ in range(steps):
        model.update()
    kt = model.kernel_times()
    print(f"n={n} k={k} m={m} batch={batch}: graph replay {wall * 1e3:.4f} ms/update  ({batch / wall:.1f} instance-updates/s)  "
          f"tma={'off' if os.environ.get('PMB_NO_TMA') == '1' else 'on'}")
    for name, (ms, cnt) in kt.items():
        if cnt:
            per = ms / steps
            extra = ""
            if name == "affine_assemble":
                extra = f"  {st['affine_algorithmic_bytes'] * batch / (per * 1e-3) / 1e9:8.1f} GB/s algorithmic"
            if name == "gram_dmma":
                extra = f"  {st['gram_algorithmic_flops'] * batch / (per * 1e-3) / 1e12:8.2f} TFLOP/s (SYRK count)"
            print(f"   {name:18s} {per * 1e3:10.2f} us/update ({cnt // steps} launches){extra}")


if __name__ == "__main__":
    main()
