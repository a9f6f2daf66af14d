"""Builds the SAME model twice -- in the oracle (literal restatement of the reference) and on the device through the
C ABI -- from one description, so every GPU parity test reads like the reference's own model tests."""
import numpy as np

from oracle import dense as D
from oracle import model as OM


class OracleSide:
    """namespace of model-building verbs backed by oracle/model.py"""

    def __init__(self):
        self.model = OM.Model()
        self.holders = []

    def variables(self, n):
        return [self.model.variable() for _ in range(n)]

    def parameter(self, holder):
        """holder: dict with key 'v' -> current numpy value (or python float for scalars)"""
        if np.ndim(holder["v"]) == 0:
            return OM.Parameter(lambda: float(holder["v"]), self.model)
        val = np.array(holder["v"], dtype=float)

        def f(buf):
            buf[...] = holder["v"]
        return OM.Parameter(f, self.model, val=val)

    def mul(self, *a):
        return OM.expression(OM.g_mul, *a)

    def add(self, *a):
        return OM.expression(OM.g_add, *a)

    def sub(self, a, b):
        return OM.expression(OM.g_sub, a, b)

    def dot(self, a, b):
        return OM.expression(OM.g_dot, a, b)

    def adjoint(self, a):
        return OM.expression(OM.g_adjoint, a)

    def vcat(self, *a):
        return OM.expression(OM.g_vcat, *a)

    def vect(self, a):
        return OM.expression(OM.g_vect, a)

    def objective(self, sense, e):
        self.model.setobjective(OM.Minimize if sense == "min" else OM.Maximize, e)

    def constraint(self, lhs, rel, rhs):
        return self.model.constraint(lhs, rel, rhs)


class DeviceSide:
    """the same verbs backed by parametron_b200 (C ABI -> CUDA kernels)"""

    def __init__(self, batch=1, device=0):
        import parametron_b200 as pm
        self.pm = pm
        self.model = pm.Model(batch=batch, device=device)

    def variables(self, n):
        return self.model.variables(n)

    def parameter(self, holder):
        if np.ndim(holder["v"]) == 0:
            return self.pm.Parameter(self.model, f=lambda: np.float64(holder["v"]))
        val = np.array(holder["v"], dtype=float)

        def f(buf):
            buf[...] = holder["v"]
        return self.pm.Parameter(self.model, f=f, val=val)

    def mul(self, *a):
        if len(a) == 3:
            return self.pm.bilinear(*a, model=self.model)
        return self.model.expr(a[0]) * a[1]

    def add(self, *a):
        r = self.model.expr(a[0]) + a[1]
        for x in a[2:]:  # n-ary + folds pairwise, lazyexpression.jl:234-236
            r = r + x
        return r

    def sub(self, a, b):
        return self.model.expr(a) - b

    def dot(self, a, b):
        return self.pm.dot(a, b, model=self.model)

    def adjoint(self, a):
        return self.pm.adjoint(a, model=self.model)

    def vcat(self, *a):
        return self.pm.vcat(*a, model=self.model)

    def vect(self, a):
        return self.pm.vect(a, model=self.model)

    def objective(self, sense, e):
        self.model.setobjective(self.pm.Minimize if sense == "min" else self.pm.Maximize, e)

    def constraint(self, lhs, rel, rhs):
        return self.pm.constraint(self.model, lhs, rel, rhs)


def oracle_outputs(om, nvars):
    """Run one oracle update! and densify what it pushed: (objective (Q,q,c0) or None, [(C,d), ...] in push order,
    [terms...] raw VectorAffine/ScalarAffine term lists)."""
    om.optimizer.log.clear()
    om.update()
    obj, blocks, terms = None, [], []
    for entry in om.optimizer.log:
        if entry[0] == "objective":
            f = entry[1]
            if isinstance(f, OM.ScalarQuadraticFunctionMOI):
                obj = D.moi_quadratic_to_dense(f, nvars)
            else:
                q, c0 = D.moi_scalar_affine_to_dense(f, nvars)
                obj = (None, q, c0)
        else:
            f = entry[2]
            if isinstance(f, OM.VectorAffineFunctionMOI):
                blocks.append(D.moi_vector_affine_to_dense(f, nvars))
                terms.append([(r, v) for r, _, v in f.terms])
            else:
                q, c0 = D.moi_scalar_affine_to_dense(f, nvars)
                blocks.append((q[None, :], np.array([c0])))
                terms.append([(1, v) for _, v in f.terms])
    return obj, blocks, terms


def assert_close(a, b, rtol=1e-10, what=""):
    """the north_star tolerance: 1e-10 relative (to the largest magnitude of the array) for fp64 coefficient arrays"""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = max(np.abs(b).max(initial=0.0), 1e-300)
    err = np.abs(a - b).max(initial=0.0)
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"
