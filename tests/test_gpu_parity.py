"""GPU parity: the CUDA path behind the C ABI against the literal oracle on the same inputs (needs a B200: -m gpu).

Tolerance (north_star): 1e-10 relative for fp64 coefficient arrays; bit-exact (==) for variable-index / sparsity bookkeeping.
Each case restates a model of the reference's tests (file:line in the docstring)."""
import numpy as np
import pytest

from tests.twin import DeviceSide, OracleSide, assert_close, oracle_outputs

pytestmark = pytest.mark.gpu


def run_twin(build, holders, mutate, nupdates=3, idxmap=None, check_terms=True, rtol=1e-10):
    """build(side) declares the model on one side and returns nvars; `mutate(step)` changes the holders in place."""
    o, d = OracleSide(), DeviceSide()
    nv = build(o)
    assert build(d) == nv
    o.model.initialize(idxmap=({i + 1: idxmap[i] for i in range(nv)} if idxmap is not None else None))
    d.model.initialize(idxmap=idxmap)
    for step in range(nupdates):
        mutate(step)
        obj, blocks, terms = oracle_outputs(o.model, nv)
        d.model.update()
        Q, q, c0 = d.model.objective_arrays()
        if obj is not None:
            if obj[0] is not None:
                assert_close(Q, obj[0], rtol, "Q")
            assert_close(q, obj[1], rtol, "q")
            assert_close([c0], [obj[2]], rtol, "c0")
        # the oracle pushes only non-constant blocks, in update order
        pos = [p for p in range(d.model.num_blocks()) if not d.model.block_view(p).is_constant]
        assert len(pos) == len(blocks)
        for p, (Cb, db), tl in zip(pos, blocks, terms):
            Cd, dd = d.model.block_arrays(p)
            assert_close(Cd, Cb, rtol, f"C[{p}]")
            assert_close(dd, db, rtol, f"d[{p}]")
            if check_terms:
                rows, vars_, src = d.model.block_terms(p)
                assert [(int(r), int(v)) for r, v in zip(rows, vars_)] == tl  # bit-exact bookkeeping
                assert np.array_equal(src, (rows - 1) + (vars_ - 1) * Cb.shape[0])
    return o, d


def test_readme_example_1():
    """README.md:31-56 / test/model.jl:83-125: n=8, m=2, min ||Ax-b||^2 s.t. Cx == d, 4 re-parameterisations."""
    n, m = 8, 2
    rng = np.random.default_rng(1234)
    A, b, Cm, dv = ({"v": rng.random(s)} for s in ((n, n), (n,), (m, n), (m,)))

    def build(s):
        x = s.variables(n)
        pA, pb, pC, pd = (s.parameter(h) for h in (A, b, Cm, dv))
        residual = s.sub(s.mul(pA, x), pb)
        s.objective("min", s.dot(residual, residual))
        s.constraint(s.mul(pC, x), "==", pd)
        return n

    def mutate(step):
        A["v"] = rng.random((n, n)); b["v"] = rng.random(n); Cm["v"] = rng.random((m, n)); dv["v"] = 2 * rng.random(m)

    o, d = run_twin(build, None, mutate, nupdates=4)
    # dense meaning stated in the reference's own comment test/model.jl:84
    Q, q, c0 = d.model.objective_arrays()
    assert_close(Q, 2 * A["v"].T @ A["v"], 1e-10, "Q=2A'A")
    assert_close(q, -2 * A["v"].T @ b["v"], 1e-10, "q=-2A'b")
    st = d.model.stats()
    assert st["uses_cuda_graph"] == 1 and st["kernel_launches_per_update"] > 0


def test_readme_example_1_with_index_map():
    """mapindices! model.jl:100-107: the optimizer permutes variable indices; columns must follow (bit-exact)."""
    n, m = 6, 3
    rng = np.random.default_rng(7)
    A, b, Cm, dv = ({"v": rng.random(s)} for s in ((n, n), (n,), (m, n), (m,)))
    perm = [int(i) + 1 for i in rng.permutation(n)]

    def build(s):
        x = s.variables(n)
        pA, pb, pC, pd = (s.parameter(h) for h in (A, b, Cm, dv))
        r = s.sub(s.mul(pA, x), pb)
        s.objective("min", s.dot(r, r))
        s.constraint(s.mul(pC, x), "==", pd)
        return n

    run_twin(build, None, lambda step: A.__setitem__("v", rng.random((n, n))), idxmap=perm)


def test_unconstrained_bilinear_objective():
    """test/model.jl:26-81: min x'Qx + r.x + s  via the 3-argument `*` rule (lazyexpression.jl:219-226)."""
    n = 5
    rng = np.random.default_rng(3)
    Qh, rh, sh = {"v": np.eye(n)}, {"v": rng.random(n)}, {"v": 1.5}

    def build(s):
        x = s.variables(n)
        Q, r, sc = s.parameter(Qh), s.parameter(rh), s.parameter(sh)
        xT = s.adjoint(x)
        s.objective("min", s.add(s.mul(xT, Q, x), s.dot(r, x), sc))
        return n

    def mutate(step):
        M = rng.random((n, n))
        Qh["v"] = M if step == 1 else M @ M.T  # step 1: NON-symmetric Q exercises the k-index pairing of functions.jl:849-856
        rh["v"] = rng.random(n); sh["v"] = float(rng.random())

    run_twin(build, None, mutate)


def test_box_constraints():
    """test/model.jl:127-171: min x.x - p.x  s.t.  x >= l, x <= u with Parameter bounds."""
    n = 7
    rng = np.random.default_rng(5)
    ph, lh, uh = {"v": rng.random(n)}, {"v": -rng.random(n)}, {"v": rng.random(n)}

    def build(s):
        x = s.variables(n)
        p, l, u = s.parameter(ph), s.parameter(lh), s.parameter(uh)
        s.objective("min", s.sub(s.dot(x, x), s.dot(p, x)))
        s.constraint(x, ">=", l)
        s.constraint(x, "<=", u)
        return n

    def mutate(step):
        ph["v"] = rng.random(n); lh["v"] = -rng.random(n); uh["v"] = rng.random(n)

    run_twin(build, None, mutate)


def test_adjoint_shapes():
    """README Example 2 / test/model.jl:350-351: resid = X'g - p ; objective resid'resid (Adjoint{vector} * vector)."""
    k, n = 4, 6
    rng = np.random.default_rng(11)
    Xh, ph = {"v": rng.random((n, k))}, {"v": rng.random(k)}

    def build(s):
        g = s.variables(n)
        X, p = s.parameter(Xh), s.parameter(ph)
        resid = s.sub(s.mul(s.adjoint(X), g), p)
        s.objective("min", s.mul(s.adjoint(resid), resid))
        return n

    def mutate(step):
        Xh["v"] = rng.random((n, k)); ph["v"] = rng.random(k)

    run_twin(build, None, mutate)


def test_issue_32_mpc_shape():
    """test/lazyexpression.jl:330-347: H * (v - v0) - dt * (u - c) == 0 -- matvec over Vector{AffineFunction}, scale!, vecsubtract!."""
    n = 4
    rng = np.random.default_rng(13)
    Hh, v0h, ch, dth = {"v": rng.random((n, n))}, {"v": rng.random(n)}, {"v": rng.random(n)}, {"v": 0.01}

    def build(s):
        v = s.variables(n)
        u = s.variables(n)
        H, v0, c, dt = (s.parameter(h) for h in (Hh, v0h, ch, dth))
        lhs = s.sub(s.mul(H, s.sub(v, v0)), s.mul(dt, s.sub(u, c)))
        s.constraint(lhs, "==", np.zeros(n))
        s.objective("min", s.dot(u, u))
        return 2 * n

    def mutate(step):
        Hh["v"] = rng.random((n, n)); v0h["v"] = rng.random(n); ch["v"] = rng.random(n); dth["v"] = float(rng.random())

    run_twin(build, None, mutate)


def test_vcat_blocks():
    """test/lazyexpression.jl:188-245: vcat(A*x, B*y) as one constraint block; y shares no variable with x."""
    rng = np.random.default_rng(17)
    Ah, Bh, dh = {"v": rng.random((3, 4))}, {"v": rng.random((2, 3))}, {"v": rng.random(5)}

    def build(s):
        x = s.variables(4)
        y = s.variables(3)
        A, B, d = s.parameter(Ah), s.parameter(Bh), s.parameter(dh)
        s.constraint(s.vcat(s.mul(A, x), s.mul(B, y)), "<=", d)
        return 7

    def mutate(step):
        Ah["v"] = rng.random((3, 4)); Bh["v"] = rng.random((2, 3)); dh["v"] = rng.random(5)

    run_twin(build, None, mutate)


def test_scaled_and_summed_affine_rows():
    """scale! (functions.jl:873-925) + vecadd! (:751-764) with two matrices hitting the SAME variables (duplicates summed
    only in the dense meaning, never in the term list)."""
    n, m = 5, 4
    rng = np.random.default_rng(19)
    Ah, Bh, wh, dh = {"v": rng.random((m, n))}, {"v": rng.random((m, n))}, {"v": 0.5}, {"v": rng.random(m)}

    def build(s):
        x = s.variables(n)
        A, B, w, d = (s.parameter(h) for h in (Ah, Bh, wh, dh))
        e = s.add(s.mul(w, s.mul(A, x)), s.mul(B, x))
        s.constraint(e, ">=", d)
        r = s.sub(s.mul(A, x), d)
        s.objective("min", s.mul(w, s.dot(r, r)))  # mul!(::QuadraticFunction, ...) lazyexpression.jl:260-274
        return n

    def mutate(step):
        Ah["v"] = rng.random((m, n)); Bh["v"] = rng.random((m, n)); wh["v"] = float(rng.random()) + 0.1; dh["v"] = rng.random(m)

    run_twin(build, None, mutate)


def test_dot_of_different_affine_vectors():
    """vecdot!(::QuadraticFunction, f, h) with f != h (functions.jl:702-709): Q = M + M', M = F'H."""
    n, k = 5, 6
    rng = np.random.default_rng(23)
    Ah, Bh, bh, ch = {"v": rng.random((k, n))}, {"v": rng.random((k, n))}, {"v": rng.random(k)}, {"v": rng.random(k)}

    def build(s):
        x = s.variables(n)
        A, B, b, c = (s.parameter(h) for h in (Ah, Bh, bh, ch))
        f = s.sub(s.mul(A, x), b)
        h = s.add(s.mul(B, x), c)
        s.objective("min", s.dot(f, h))
        return n

    def mutate(step):
        Ah["v"] = rng.random((k, n)); Bh["v"] = rng.random((k, n)); bh["v"] = rng.random(k); ch["v"] = rng.random(k)

    run_twin(build, None, mutate)


def test_constant_blocks_are_skipped():
    """test/model.jl:281-300 + moi_interop.jl:123,132,153,169: constraints without Parameters are never re-pushed."""
    import parametron_b200 as pm
    m = pm.Model()
    x, y, z = m.variables(3)
    e = m.expr(x) + 2.0 * m.expr(y) + 3.0 * m.expr(z)
    pm.constraint(m, e, ">=", 4.0)
    pm.constraint(m, m.expr(x) + y, ">=", 1.0)
    p = pm.Parameter(m, f=lambda: np.float64(2.0))
    pm.constraint(m, p * m.expr(z), "<=", 1.0)
    pm.objective(m, pm.Minimize, pm.dot([x, y, z], [x, y, z], model=m))
    m.initialize()
    m.update()
    assert [m.block_view(i).is_constant for i in range(3)] == [1, 1, 0]
    C0, d0 = m.block_arrays(0)
    assert np.array_equal(C0, [[1.0, 2.0, 3.0]]) and np.array_equal(d0, [-4.0])
    C2, d2 = m.block_arrays(2)
    assert np.array_equal(C2, [[0.0, 0.0, 2.0]]) and np.array_equal(d2, [-1.0])
    Q, q, c0 = m.objective_arrays()
    assert np.array_equal(Q, 2 * np.eye(3)) and not q.any() and c0 == 0.0
    assert m.objective_view().is_constant == 1


def test_bucket_order():
    """moi_interop.jl:180-193,236-247: scalar GreaterThan/LessThan/EqualTo, then vector Nonnegatives/Nonpositives/Zeros,
    insertion order within a bucket -- independent of declaration order."""
    import parametron_b200 as pm
    m = pm.Model()
    x = m.variables(3)
    p = pm.Parameter(m, f=lambda: np.ones(3))
    s = pm.Parameter(m, f=lambda: np.float64(1.0))
    pm.constraint(m, x, "==", p)                     # vector Zeros        bucket 5
    pm.constraint(m, x, ">=", p)                     # vector Nonnegatives bucket 3
    pm.constraint(m, s * m.expr(x[0]), "==", 1.0)    # scalar EqualTo      bucket 2
    pm.constraint(m, x, "<=", p)                     # vector Nonpositives bucket 4
    pm.constraint(m, s * m.expr(x[1]), ">=", 1.0)    # scalar GreaterThan  bucket 0
    pm.constraint(m, 2.0 * p * 1.0 + x, ">=", 0.0 * np.ones(3))  # vector Nonnegatives bucket 3 (second)
    m.initialize()
    assert [m.block_view(i).bucket for i in range(m.num_blocks())] == [0, 2, 3, 3, 4, 5]
    m.update()
    C3, d3 = m.block_arrays(3)
    assert np.array_equal(d3, 2.0 * np.ones(3)) and np.array_equal(C3, np.eye(3))


def test_errors_mirror_the_reference():
    """DimensionMismatch (functions.jl:669,755,780-781,845) and "already initialized" (model.jl:50,61,69)."""
    import parametron_b200 as pm
    m = pm.Model()
    x = m.variables(3)
    A = pm.Parameter(m, f=lambda: np.ones((2, 4)))
    b = pm.Parameter(m, f=lambda: np.ones(5))
    with pytest.raises(pm.DimensionMismatch):
        A * x
    with pytest.raises(pm.DimensionMismatch):
        m.expr(x) - b
    with pytest.raises(pm.DimensionMismatch):
        pm.dot(x, b, model=m)
    with pytest.raises(pm.DimensionMismatch):
        pm.bilinear(pm.adjoint(x, model=m), A, x)
    m.initialize()
    with pytest.raises(pm.AlreadyInitialized):
        m.variables(1)
    with pytest.raises(pm.AlreadyInitialized):
        m.setobjective(pm.Minimize, pm.Expr(m, 0))
    with pytest.raises(pm.AlreadyInitialized):
        m.add_zero_constraint(pm.Expr(m, 0))
    with pytest.raises(pm.AlreadyInitialized):
        m.initialize()


def test_parameters_evaluated_once_per_update():
    """test/parameter.jl:8-40 + model.jl:133: each Parameter's function runs exactly once per update!, even when the
    Parameter feeds several expressions (objective twice + a constraint)."""
    import parametron_b200 as pm
    m = pm.Model()
    x = m.variables(4)
    calls = [0]

    def fA(buf):
        calls[0] += 1
        buf[...] = calls[0]
    A = pm.Parameter(m, f=fA, val=np.zeros((4, 4)))
    r = A * x
    pm.objective(m, pm.Minimize, pm.dot(r, r))
    pm.constraint(m, A * x, "==", np.zeros(4))
    m.initialize()
    for k in range(1, 4):
        m.update()
        assert calls[0] == k
        Cb, _ = m.block_arrays(0)
        assert np.array_equal(Cb, np.full((4, 4), float(k)))
        Q, _, _ = m.objective_arrays()
        assert np.array_equal(Q, np.full((4, 4), 2.0 * 4 * k * k))
