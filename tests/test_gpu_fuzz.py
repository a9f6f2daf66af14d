"""Randomised expression-DAG parity (needs a B200: -m gpu): seeded random models assembled from the rewrite rules of
src/lazyexpression.jl:198-302 are built on both sides (literal oracle / device) and compared after several
re-parameterisations -- 1e-10 relative on every coefficient array, == on the term index where the device can reproduce it."""
import numpy as np
import pytest

from tests.twin import DeviceSide, OracleSide, assert_close, oracle_outputs

pytestmark = pytest.mark.gpu


class Recipe:
    """A seeded random model description, replayable on either side."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.seed = seed
        r = self.rng
        self.n = int(r.integers(2, 7))
        self.holders = {}
        self.ops = []          # list of (kind, spec) for constraints
        self.obj = None
        nvec = self.n
        k = int(r.integers(1, 6))
        self.k = k
        self._h("A", (k, nvec)); self._h("B", (k, nvec)); self._h("b", (k,)); self._h("c", (k,)); self._h("p", (nvec,))
        self._h("w", ()); self._h("S", (nvec, nvec)); self._h("H", (nvec, nvec)); self._h("v0", (nvec,))
        self.obj = str(r.choice(["ls", "ls_weighted", "bilinear", "sum_two", "cross", "linear", "none"]))
        ncon = int(r.integers(0, 4))
        for _ in range(ncon):
            self.ops.append((str(r.choice(["matvec", "scaled_sum", "box", "vcat", "mpc", "scalar", "permvars"])),
                             str(r.choice(["==", "<=", ">="]))))
        self.perm = [int(i) + 1 for i in r.permutation(nvec)] if r.random() < 0.5 else None
        self.sub = sorted(int(i) for i in r.choice(nvec, size=max(1, nvec // 2), replace=False))

    def _h(self, name, shape):
        self.holders[name] = {"v": self.rng.random(shape) if shape else float(self.rng.random()) + 0.25}

    def mutate(self):
        for h in self.holders.values():
            h["v"] = self.rng.random(np.shape(h["v"])) - 0.3 if np.ndim(h["v"]) else float(self.rng.random()) + 0.25

    def build(self, s):
        x = s.variables(self.n)
        P = {k_: s.parameter(h) for k_, h in self.holders.items()}
        r = s.sub(s.mul(P["A"], x), P["b"])
        if self.obj == "ls":
            s.objective("min", s.dot(r, r))
        elif self.obj == "ls_weighted":
            s.objective("min", s.mul(P["w"], s.dot(r, r)))
        elif self.obj == "bilinear":
            s.objective("min", s.add(s.mul(s.adjoint(x), P["S"], x), s.dot(P["p"], x), P["w"]))
        elif self.obj == "sum_two":
            s.objective("min", s.add(s.dot(r, r), s.dot(x, x)))
        elif self.obj == "cross":
            h = s.add(s.mul(P["B"], x), P["c"])
            s.objective("max", s.dot(r, h))
        elif self.obj == "linear":
            s.objective("min", s.dot(P["p"], x))
        exact = True
        for kind, rel in self.ops:
            if kind == "matvec":
                s.constraint(s.mul(P["A"], x), rel, P["b"])
            elif kind == "scaled_sum":
                s.constraint(s.add(s.mul(P["w"], s.mul(P["A"], x)), s.mul(P["B"], x)), rel, P["c"])
            elif kind == "box":
                s.constraint(x, rel, P["p"])
            elif kind == "vcat":
                s.constraint(s.vcat(s.mul(P["A"], x), s.mul(P["B"], x)), rel, np.concatenate([np.zeros(self.k), np.ones(self.k)]))
            elif kind == "mpc":
                s.constraint(s.sub(s.mul(P["H"], s.sub(x, P["v0"])), s.mul(P["w"], s.sub(x, P["p"]))), rel, np.zeros(self.n))
            elif kind == "scalar":
                s.constraint(s.dot(P["p"], x), rel, P["w"])
            elif kind == "permvars":
                xs = [x[i] for i in self.sub]
                s.constraint(s.mul(P["w"], xs), rel, np.ones(len(xs)))
        return self.n, exact


@pytest.mark.parametrize("seed", list(range(40)))
def test_random_models(seed):
    rec = Recipe(seed)
    o, d = OracleSide(), DeviceSide()
    nv, _ = rec.build(o)
    rec.build(d)
    o.model.initialize(idxmap=({i + 1: rec.perm[i] for i in range(nv)} if rec.perm else None))
    d.model.initialize(idxmap=rec.perm)
    for step in range(3):
        rec.mutate()
        obj, blocks, terms = oracle_outputs(o.model, nv)
        d.model.update()
        Q, q, c0 = d.model.objective_arrays()
        if obj is not None:
            if obj[0] is not None:
                assert_close(Q, obj[0], 1e-10, f"seed {seed} Q")
            assert_close(q, obj[1], 1e-10, f"seed {seed} q")
            assert_close([c0], [obj[2]], 1e-10, f"seed {seed} c0")
        pos = [p for p in range(d.model.num_blocks()) if not d.model.block_view(p).is_constant]
        assert len(pos) == len(blocks)
        for p, (Cb, db), tl in zip(pos, blocks, terms):
            Cd, dd = d.model.block_arrays(p)
            assert_close(Cd, Cb, 1e-10, f"seed {seed} C[{p}]")
            assert_close(dd, db, 1e-10, f"seed {seed} d[{p}]")
            rows, vars_, _ = d.model.block_terms(p)
            assert [(int(r_), int(v)) for r_, v in zip(rows, vars_)] == tl, f"seed {seed} term index of block {p}"
    d.model.close()
