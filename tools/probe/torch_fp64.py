# Probe: cuBLAS fp64 GEMM/SYRK-shaped throughput (roofline denominator for the Gram kernel) + host info.
import torch, time, os, json
print("nproc", os.cpu_count())
p = torch.cuda.get_device_properties(0); print(p)
def bench(f, n=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
res = {}
for N in (4096, 8192):
    A = torch.rand(N, N, dtype=torch.float64, device="cuda"); B = torch.rand(N, N, dtype=torch.float64, device="cuda")
    ms = bench(lambda: torch.matmul(A, B)); res[f"dgemm_nn_{N}"] = (ms, 2 * N**3 / ms / 1e9)
    ms = bench(lambda: torch.matmul(A.t(), B)); res[f"dgemm_tn_{N}"] = (ms, 2 * N**3 / ms / 1e9)
    ms = bench(lambda: torch.matmul(A.t(), A)); res[f"dgemm_ata_{N}"] = (ms, 2 * N**3 / ms / 1e9)
    C = torch.empty(N, N, dtype=torch.float64, device="cuda")
    ms = bench(lambda: C.copy_(A)); res[f"copy_{N}"] = (ms, 2 * 8 * N * N / ms / 1e6)
# sustained dgemm for 3 s
N = 8192; A = torch.rand(N, N, dtype=torch.float64, device="cuda"); B = torch.rand(N, N, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.time(); it = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    torch.matmul(A, B); it += 1
    if it % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
res["dgemm_sustained_8192"] = (ms / it, 2 * N**3 * it / ms / 1e9)
for k, v in res.items(): print(f"{k:28s} {v[0]:10.3f} ms  {v[1]:10.1f} GFLOP/s|GB/s")
json.dump(res, open("gpurun_out/probe_fp64.json", "w"), indent=1)
