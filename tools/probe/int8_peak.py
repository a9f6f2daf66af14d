"""Measured dense int8 tensor-core rate of this B200 through the library path (torch._int_mm -> cuBLASLt), to put a MEASURED number next to the
2 x bf16 denominator of bench.py's roofline.  usage: python tools/probe/int8_peak.py"""
import json
import torch

torch.cuda.init()
res = {}
for n in (4096, 8192, 16384):
    a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
    try:
        for _ in range(3):
            c = torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); c = torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[f"int8_{n}"] = {"ms": best, "tops": 2.0 * n ** 3 / (best * 1e-3) / 1e12}
    except Exception as exc:
        res[f"int8_{n}"] = {"error": repr(exc)[:200]}
    x = torch.randn(n, n, dtype=torch.bfloat16, device="cuda"); y = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        z = x @ y
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); z = x @ y; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"bf16_{n}"] = {"ms": best, "tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12}
    del a, b, x, y
print(json.dumps(res, indent=1))
