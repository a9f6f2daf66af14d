"""The golden model descriptions: each entry builds one model through the verbs shared by tests/twin.py (OracleSide / DeviceSide) from a
dict of named input arrays.  Shapes restate the reference's own model tests: README Example 1 (README.md:31-56, test/model.jl:83-125),
the box-constrained model (test/model.jl:127-171), the MPC-shape expression of issue 32 (test/lazyexpression.jl:330-347) and the
`x' Q x + q' x` objective of issue 26 (test/lazyexpression.jl:316-328)."""
import numpy as np


def inputs(name):
    rng = np.random.default_rng({"readme": 11, "box": 12, "mpc": 13, "bilinear": 14}[name])
    if name == "readme":
        n, m = 8, 2
        return {"A": rng.random((n, n)), "b": rng.random(n), "C": rng.random((m, n)), "d": 2 * rng.random(m)}
    if name == "box":
        n = 6
        return {"p": rng.standard_normal(n), "lo": -rng.random(n), "hi": rng.random(n)}
    if name == "mpc":
        nv, nu = 5, 3
        return {"H": rng.standard_normal((nv, nv)), "v0": rng.standard_normal(nv), "B": rng.standard_normal((nv, nu)),
                "c": rng.standard_normal(nv), "dt": np.float64(0.05)}
    if name == "bilinear":
        n = 7
        return {"Q": rng.standard_normal((n, n)), "q": rng.standard_normal(n)}
    raise KeyError(name)


def nvars(name):
    return {"readme": 8, "box": 6, "mpc": 8, "bilinear": 7}[name]


def declare(name, s, holders):
    """s: OracleSide or DeviceSide; holders: {input name: {'v': array}}"""
    P = {k: s.parameter(h) for k, h in holders.items()}
    if name == "readme":      # min ||A x - b||^2  s.t.  C x == d
        x = s.variables(8)
        r = s.sub(s.mul(P["A"], x), P["b"])
        s.objective("min", s.dot(r, r))
        s.constraint(s.mul(P["C"], x), "==", P["d"])
    elif name == "box":       # min (x - p).(x - p)  s.t.  x >= lo, x <= hi
        x = s.variables(6)
        e = s.sub(x, P["p"])
        s.objective("min", s.dot(e, e))
        s.constraint(x, ">=", P["lo"])
        s.constraint(x, "<=", P["hi"])
    elif name == "mpc":       # min (H (v - v0)).(H (v - v0))  s.t.  v == dt * (B u - c): scale!, matvecmul on affine vectors, vcat-free
        v = s.variables(5)
        u = s.variables(3)
        e = s.mul(P["H"], s.sub(v, P["v0"]))
        s.objective("min", s.dot(e, e))
        s.constraint(v, "==", s.mul(P["dt"], s.sub(s.mul(P["B"], u), P["c"])))
    elif name == "bilinear":  # min x' Q x + q . x   (issue 26)
        x = s.variables(7)
        s.objective("min", s.add(s.mul(s.adjoint(x), P["Q"], x), s.dot(P["q"], x)))
    else:
        raise KeyError(name)


NAMES = ("readme", "box", "mpc", "bilinear")
