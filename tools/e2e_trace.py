"""Timeline of one pipelined e2e update (PMB_TRACE=1 makes pmb_update execute the graph's multi-stream schedule eagerly with timing
events and print when each chunk landed, was multiplied and left).  usage: PMB_TRACE=1 python tools/e2e_trace.py"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parametron_b200 as pm  # noqa: E402

n, k, m = 4096, 4096, 1024
rng = np.random.default_rng(0)
host = {"A": rng.random((k, n)), "b": rng.random(k), "C": rng.random((m, n)), "d": 2 * rng.random(m)}
model = pm.Model()
x = model.variables(n)
P = {key: pm.Parameter(model, val=v) for key, v in host.items()}
r = P["A"] * x - P["b"]
pm.objective(model, pm.Minimize, pm.dot(r, r))
pm.constraint(model, P["C"] * x, "==", P["d"])
model.enable_qp_pack()
model.initialize()
sink = model.qp_host_sink()
for i in range(3):
    sys.stderr.write(f"--- update {i}\n")
    model._ck(model._lib.pmb_update(model._ctx))
    model.synchronize()
print("levels", model.gram_level_counts())
