"""PCIe facts of the box the e2e arm runs on: pinned H2D / D2H bandwidth alone and concurrently (full duplex?), chunked vs whole copies."""
import time
import torch

def t(fn, reps=5):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3

n_up, n_dn = 168_000_000 // 8, 100_000_000 // 8
hu = torch.empty(n_up, dtype=torch.float64).pin_memory(); du = torch.empty(n_up, dtype=torch.float64, device="cuda")
hd = torch.empty(n_dn, dtype=torch.float64).pin_memory(); dd = torch.empty(n_dn, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def up():
    with torch.cuda.stream(s1): du.copy_(hu, non_blocking=True)
def dn():
    with torch.cuda.stream(s2): hd.copy_(dd, non_blocking=True)
def both():
    up(); dn()
def up_chunks(c=9):
    step = n_up // c
    with torch.cuda.stream(s1):
        for i in range(c): du[i * step:(i + 1) * step].copy_(hu[i * step:(i + 1) * step], non_blocking=True)
a, b, c, d = t(up), t(dn), t(both), t(up_chunks)
print(f"H2D 168 MB: {a:.3f} ms ({168 / a:.1f} GB/s)   D2H 100 MB: {b:.3f} ms ({100 / b:.1f} GB/s)   both concurrently: {c:.3f} ms   H2D in 9 chunks: {d:.3f} ms")
