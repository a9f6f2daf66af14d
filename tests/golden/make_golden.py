"""Generates tests/golden/*.npz: seeded inputs and what the LITERAL oracle (oracle/model.py, the term-list restatement of the
reference's update! + moi_interop translation) pushes for them, densified -- (Q, q, c0) of the objective and (C_i, d_i) of every
constraint block in push order, plus the raw term order (row, optimizer variable) of every block.

The reference is Julia and cannot run in this image or on the GPU box, so these are NOT outputs of the reference
itself; the oracle that produced them is pinned on the reference's RNG-free known answers (tests/test_oracle_functions.py,
tests/test_oracle_model.py).  Their job: freeze oracle + device together -- a later change that moves both the oracle and the CUDA path
in the same wrong direction no longer passes.

usage (repo root):  python -m tests.golden.make_golden
"""
import os

import numpy as np

from tests.golden import models as M
from tests.twin import OracleSide, oracle_outputs

HERE = os.path.dirname(os.path.abspath(__file__))


def build(name):
    vals = M.inputs(name)
    holders = {k: {"v": v} for k, v in vals.items()}
    o = OracleSide()
    M.declare(name, o, holders)
    o.model.initialize()
    obj, blocks, terms = oracle_outputs(o.model, M.nvars(name))
    out = {f"in_{k}": np.asarray(v) for k, v in vals.items()}
    if obj is not None:
        if obj[0] is not None:
            out["Q"] = obj[0]
        out["q"], out["c0"] = obj[1], np.float64(obj[2])
    out["nblocks"] = np.int64(len(blocks))
    for i, (Cb, db) in enumerate(blocks):
        out[f"C{i}"], out[f"d{i}"] = Cb, db
        out[f"term_rows{i}"] = np.array([r for r, _ in terms[i]], dtype=np.int64)
        out[f"term_vars{i}"] = np.array([getattr(v, "value", v) for _, v in terms[i]], dtype=np.int64)
    return out


def main():
    for name in M.NAMES:
        np.savez(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)


if __name__ == "__main__":
    main()
