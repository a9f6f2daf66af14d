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
    import os
    assert st["uses_cuda_graph"] == (0 if os.environ.get("PMB_NO_GRAPH") == "1" else 1) and st["kernel_launches_per_update"] > 0


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


def test_nested_matrix_products():
    """A * (B * x) - c : matvecmul! over Vector{AffineFunction} whose rows are matrix rows (functions.jl:800-822) -> dense A*B."""
    rng = np.random.default_rng(29)
    Ah, Bh, ch = {"v": rng.random((3, 4))}, {"v": rng.random((4, 5))}, {"v": rng.random(3)}

    def build(s):
        x = s.variables(5)
        A, B, c = s.parameter(Ah), s.parameter(Bh), s.parameter(ch)
        s.constraint(s.mul(A, s.sub(s.mul(B, x), np.ones(4))), "<=", c)
        return 5

    def mutate(step):
        Ah["v"] = rng.random((3, 4)); Bh["v"] = rng.random((4, 5)); ch["v"] = rng.random(3)

    run_twin(build, None, mutate, check_terms=False)  # the reference emits 4 terms per variable here; the dense block sums them


def test_numbers_dot_affine_vector_objective():
    """dot(p, A*x + q) with numeric p (functions.jl:665-674): affine objective, Maximize sense; plus a Variable-sum constraint."""
    n, k = 6, 4
    rng = np.random.default_rng(31)
    Ah, ph, qh, wh = {"v": rng.random((k, n))}, {"v": rng.random(k)}, {"v": rng.random(k)}, {"v": 2.0}

    def build(s):
        x = s.variables(n)
        A, p, q, w = (s.parameter(h) for h in (Ah, ph, qh, wh))
        e = s.dot(p, s.add(s.mul(A, x), q))
        s.objective("max", s.mul(w, e))
        s.constraint(s.dot(p, s.mul(A, x)), "<=", 1.0)      # scalar affine constraint (bucket ScalarAffine/LessThan)
        return n

    def mutate(step):
        Ah["v"] = rng.random((k, n)); ph["v"] = rng.random(k); qh["v"] = rng.random(k); wh["v"] = float(rng.random()) + 0.5

    o, d = run_twin(build, None, mutate, check_terms=False)
    assert d.model.objective_view().sense == 1 and d.model.objective_view().kind == 5  # Maximize, AffineFunction


def test_weighted_sum_of_quadratics_with_two_dynamic_scalars():
    """w1 * (x.x) + w2 * w1 * ||Ax - b||^2 : mul!(::QuadraticFunction, scalar) twice -> a product of two device scalars."""
    n, k = 5, 3
    rng = np.random.default_rng(37)
    Ah, bh, w1h, w2h = {"v": rng.random((k, n))}, {"v": rng.random(k)}, {"v": 0.7}, {"v": 1.3}

    def build(s):
        x = s.variables(n)
        A, b, w1, w2 = (s.parameter(h) for h in (Ah, bh, w1h, w2h))
        r = s.sub(s.mul(A, x), b)
        s.objective("min", s.add(s.mul(w1, s.dot(x, x)), s.mul(w2, s.mul(w1, s.dot(r, r)))))
        return n

    def mutate(step):
        Ah["v"] = rng.random((k, n)); bh["v"] = rng.random(k); w1h["v"] = float(rng.random()) + 0.1; w2h["v"] = float(rng.random()) + 0.1

    run_twin(build, None, mutate)


def test_repeated_variables_are_summed_in_the_dense_block():
    """x appears twice in the variable vector: the reference keeps two terms per row, MOI sums duplicates; the Gram operand
    must be materialised (non-unique scatter map) and the dense C column holds the sum."""
    import parametron_b200 as pm
    from oracle import dense as D
    rng = np.random.default_rng(41)
    A = rng.random((4, 3))
    b = rng.random(4)
    m = pm.Model()
    x = m.variables(2)
    xx = [x[0], x[1], x[0]]
    pA, pb = pm.Parameter(m, val=A), pm.Parameter(m, val=b)
    r = pA * xx - pb
    pm.objective(m, pm.Minimize, pm.dot(r, r))
    pm.constraint(m, pA * xx, ">=", pb)
    m.initialize()
    m.update()
    Aeff = np.stack([A[:, 0] + A[:, 2], A[:, 1]], axis=1)
    Qo, qo, c0o = D.least_squares_objective(Aeff, b, [1, 2], 2)
    Q, q, c0 = m.objective_arrays()
    assert_close(Q, Qo, 1e-10, "Q"); assert_close(q, qo, 1e-10, "q"); assert_close([c0], [c0o], 1e-10, "c0")
    Cb, db = m.block_arrays(0)
    assert_close(Cb, Aeff, 1e-15, "C"); assert np.array_equal(db, -b)
    rows, vars_, _ = m.block_terms(0)
    assert list(vars_) == [1, 2, 1] * 4 and list(rows) == [r_ for r_ in range(1, 5) for _ in range(3)]  # reference term list keeps both


def test_constant_quadratic_terms_and_linear_objective():
    """test/model.jl:281-321 shapes: constant QuadraticFunction objective terms (COO path) + x'Px with a Parameter scale."""
    import ctypes as C
    import parametron_b200 as pm
    m = pm.Model()
    x = m.variables(3)
    # 2 x1^2 + x1 x2 + 0.5 x3^2 + 3 x2 + 1  as a constant QuadraticFunction
    qc = np.array([2.0, 1.0, 0.5]); rv = np.array([1, 1, 3], dtype=np.int64); cv = np.array([1, 2, 3], dtype=np.int64)
    lc = np.array([3.0]); lv = np.array([2], dtype=np.int64)
    out = C.c_int32()
    m._ck(m._lib.pmb_const_quadratic(m._ctx, 3, C.c_void_p(qc.ctypes.data), C.c_void_p(rv.ctypes.data), C.c_void_p(cv.ctypes.data),
                                     1, C.c_void_p(lc.ctypes.data), C.c_void_p(lv.ctypes.data), 1.0, out))
    base = pm.Expr(m, out.value)
    w = pm.Parameter(m, f=lambda: np.float64(2.0))
    pm.objective(m, pm.Minimize, w * base + pm.dot(x, np.array([1.0, 1.0, 1.0]), model=m))
    m.initialize()
    m.update()
    Q, q, c0 = m.objective_arrays()
    # MOI Hessian: diagonal terms doubled (moi_interop.jl:58), off-diagonal term counted in both triangles
    assert np.array_equal(Q, 2.0 * np.array([[4.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]))
    assert np.array_equal(q, [1.0, 7.0, 1.0]) and c0 == 2.0


def test_c_abi_host_callback_and_external_device_pointer():
    """The two Parameter plumbing modes only the raw C ABI exposes: a C host callback run exactly once per pmb_update
    (parameter.jl:93-102 through @cfunction) and a caller-owned device buffer (manual mode, parameter.jl:88)."""
    import ctypes as C
    import parametron_b200 as pm
    from parametron_b200 import _capi as K
    n, k = 4, 3
    m = pm.Model()
    lib, ctx = m._lib, m._ctx
    x = m.variables(n)
    # host Parameter A with a C callback that fills it with (call count)
    A = C.c_int32()
    m._ck(lib.pmb_param_create(ctx, 2, k, n, K.PARAM_HOST, A))
    hostA = m._pinned((k * n,))
    calls = []

    @K.HOST_CALLBACK
    def cb(val, rows, cols, batch, user):
        calls.append((rows, cols, batch))
        for i in range(rows * cols):
            val[i] = float(len(calls))
    m._ck(lib.pmb_param_set_host_ptr(ctx, A.value, C.c_void_p(hostA.ctypes.data), C.cast(cb, C.c_void_p), None))
    # device Parameter d living in a buffer this test owns
    dbuf = C.c_void_p(m._dev_alloc(k * 8))     # a device buffer this test owns (pmb_device_alloc: on the context's device)
    d = C.c_int32()
    m._ck(lib.pmb_param_create(ctx, 1, k, 0, K.PARAM_DEVICE, d))
    m._ck(lib.pmb_param_set_device_ptr(ctx, d.value, dbuf))
    pm.constraint(m, pm.Expr(m, A.value) * x, "==", pm.Expr(m, d.value))
    m.initialize()
    p = C.c_void_p()
    m._ck(lib.pmb_param_device_ptr(ctx, d.value, p))
    assert p.value == dbuf.value                       # the library reads the caller's buffer in place
    for step in (1, 2, 3):
        dv = np.arange(k, dtype=np.float64) * step
        m._ck(lib.pmb_memcpy_h2d(ctx, dbuf, C.c_void_p(dv.ctypes.data), k * 8))
        m.update()
        Cb, db = m.block_arrays(0)
        assert len(calls) == step and calls[-1] == (k, n, 1)
        assert np.array_equal(Cb, np.full((k, n), float(step))) and np.array_equal(db, -dv)
    m.close()                                          # frees the tracked scratch as well


def test_quadratic_constraint_blocks():
    """test/model.jl:364-394 (the Gurobi QCQP case, skipped on the reference's CI): a QuadraticFunction-valued constraint goes
    to the ScalarQuadraticFunction buckets (moi_interop.jl:186-188), after every affine bucket; same (Q, q, c0) translation
    (diagonal x2, moi_interop.jl:45-62) as the objective."""
    from oracle import dense as D
    from oracle import model as OM
    n, k = 5, 4
    rng = np.random.default_rng(43)
    Ah, bh, rh = {"v": rng.random((k, n))}, {"v": rng.random(k)}, {"v": 3.0}
    sides = []
    for side in (OracleSide(), DeviceSide()):
        x = side.variables(n)
        A, b, r = (side.parameter(h) for h in (Ah, bh, rh))
        res = side.sub(side.mul(A, x), b)
        side.constraint(side.dot(res, res), "<=", r)            # ||Ax - b||^2 <= r       (ScalarQuadratic / LessThan: bucket 7)
        side.constraint(side.mul(A, x), ">=", b)                # vector Nonnegatives      (bucket 3)
        side.objective("min", side.dot(x, x))
        side.model.initialize()
        sides.append(side)
    o, d = sides
    assert [d.model.block_view(i).bucket for i in range(2)] == [3, 7]
    for step in range(3):
        Ah["v"] = rng.random((k, n)); bh["v"] = rng.random(k); rh["v"] = float(rng.random()) + 1.0
        o.model.optimizer.log.clear()
        o.model.update()
        d.model.update()
        fq = [e[2] for e in o.model.optimizer.log if e[0] == "constraint" and isinstance(e[2], OM.ScalarQuadraticFunctionMOI)][0]
        Qo, qo, c0o = D.moi_quadratic_to_dense(fq, n)
        Q, q, c0 = d.model.block_quadratic(1)
        assert_close(Q, Qo, 1e-10, "Qc"); assert_close(q, qo, 1e-10, "qc"); assert_close([c0], [c0o], 1e-10, "c0c")
        assert d.model.block_view(1).set_kind == 1 and d.model.block_view(1).is_scalar == 1


def test_string_macros_read_like_the_reference():
    """README.md:31-56 written with the string macros (the @expression / @objective / @constraint analogues) gives the same
    arrays as the operator-overload API; test/model.jl:66's objective uses the 3-argument `*` rule."""
    import parametron_b200 as pm
    from parametron_b200.macros import constraint, expression, objective
    from oracle import dense as D
    n, m_ = 8, 2
    rng = np.random.default_rng(1234)
    Av, bv, Cv, dv = rng.random((n, n)), rng.random(n), rng.random((m_, n)), rng.random(m_)
    model = pm.Model()
    x = model.variables(n)
    A, b, C_, d = (pm.Parameter(model, val=v) for v in (Av, bv, Cv, dv))
    residual = expression(model, "A * x - b", A=A, x=x, b=b)
    objective(model, pm.Minimize, "residual ⋅ residual", residual=residual)
    constraint(model, "C * x == d", C=C_, x=x, d=d)
    constraint(model, "[C * x; x] >= [d; lo]", C=C_, x=x, d=d, lo=np.zeros(n))
    model.initialize()
    model.update()
    Q, q, c0 = model.objective_arrays()
    Qo, qo, c0o = D.least_squares_objective(Av, bv, np.arange(1, n + 1), n)
    assert_close(Q, Qo, 1e-10, "Q"); assert_close(q, qo, 1e-10, "q"); assert_close([c0], [c0o], 1e-10, "c0")
    C1, d1 = model.block_arrays(1)          # Zeros bucket comes after Nonnegatives
    assert np.array_equal(C1, Cv) and np.array_equal(d1, -dv)
    C0, d0 = model.block_arrays(0)
    assert np.array_equal(C0, np.vstack([Cv, np.eye(n)])) and np.array_equal(d0, np.concatenate([-dv, np.zeros(n)]))
    # test/model.jl:66
    m2 = pm.Model()
    y = m2.variables(3)
    Qp, r, s = pm.Parameter(m2, val=np.eye(3)), pm.Parameter(m2, val=np.array([1.0, 2.0, 3.0])), pm.Parameter(m2, f=lambda: np.float64(0.5))
    objective(m2, pm.Minimize, "transpose(x) * Q * x + r ⋅ x + s", x=y, Q=Qp, r=r, s=s)
    m2.initialize()
    m2.update()
    Q2, q2, c2 = m2.objective_arrays()
    assert np.array_equal(Q2, 2 * np.eye(3)) and np.array_equal(q2, [1.0, 2.0, 3.0]) and c2 == 0.5


def test_two_models_interleaved():
    """Two live contexts on one device (each with its own stream, graph and buffers) updated alternately do not interfere."""
    import parametron_b200 as pm
    rng = np.random.default_rng(47)
    models = []
    for n in (20, 33):
        Av, bv = rng.random((n + 3, n)), rng.random(n + 3)
        m = pm.Model()
        x = m.variables(n)
        A, b = pm.Parameter(m, val=Av), pm.Parameter(m, val=bv)
        r = A * x - b
        pm.objective(m, pm.Minimize, pm.dot(r, r))
        pm.constraint(m, A * x, "<=", b)
        m.initialize()
        models.append((m, A, b))
    for step in range(3):
        for m, A, b in models:
            A.val = rng.random(A.val.shape); b.val = rng.random(b.val.shape)
            m.update()
        for m, A, b in models:
            Q, q, c0 = m.objective_arrays()
            assert_close(Q, 2 * A.val.T @ A.val, 1e-10, "Q"); assert_close(q, -2 * A.val.T @ b.val, 1e-10, "q")
            Cb, db = m.block_arrays(0)
            assert np.array_equal(Cb, A.val) and np.array_equal(db, -b.val)
    for m, _, _ in models:
        m.close()


def test_getfield_rule_on_a_struct_valued_parameter():
    """test/lazyexpression.jl:355-382 (`ex1 = @expression p.x`, `ex2 = @expression p.x + 1`) restated on the device: a struct-valued
    Parameter is one flat record; `p.field` lowers to GetField (lazyexpression.jl:300-302, functions.jl:998-999).  The fields feed an
    objective and a constraint, are re-evaluated ONCE per update! with the record, and match the literal oracle."""
    import parametron_b200 as pm
    n = 4
    rng = np.random.default_rng(21)

    class MyWrapper:
        def __init__(self):
            self.x = float(rng.random())            # scalar field
            self.w = rng.random(n)                  # vector field
            self.M = rng.random((3, n))             # matrix field (column-major on the device)

    m = pm.Model()
    x = m.variables(n)
    p = pm.RecordParameter(m, MyWrapper, {"x": (), "w": (n,), "M": (3, n)})
    ex1 = p.x                                        # @expression p.x
    ex2 = p.x + 1.0                                  # @expression p.x + 1
    pm.objective(m, pm.Minimize, ex2 * pm.dot(p.w, x) + ex1)          # (p.x + 1) * (w . x) + p.x
    pm.constraint(m, p.M * x, "==", np.arange(3.0))
    m.initialize()
    for _ in range(3):
        evals = p.evaluations
        m.update()
        assert p.evaluations == evals + 1           # parameter.jl:93-104: once per update!
        rec = p.record
        Q, q, c0 = m.objective_arrays()
        assert_close(q, (rec.x + 1.0) * rec.w, 1e-14, "q"); assert_close([c0], [rec.x], 1e-14, "c0")
        Cb, db = m.block_arrays(0)
        assert np.array_equal(Cb, rec.M) and np.array_equal(db, -np.arange(3.0))
    # the string-macro form lowers `p.x` through the same rule
    m2 = pm.Model()
    x2 = m2.variables(n)
    p2 = pm.RecordParameter(m2, MyWrapper, {"x": (), "w": (n,), "M": (3, n)})
    pm.macros.objective(m2, pm.Minimize, "(p.x + 1) * (p.w ⋅ x) + p.x", p=p2, x=x2)
    m2.initialize(); m2.update()
    _, q2, c2 = m2.objective_arrays()
    assert_close(q2, (p2.record.x + 1.0) * p2.record.w, 1e-14, "q macro"); assert_close([c2], [p2.record.x], 1e-14, "c0 macro")
    with pytest.raises(pm.DimensionMismatch):
        m3 = pm.Model()
        p3 = pm.RecordParameter(m3, MyWrapper, {"x": (), "w": (n,), "M": (3, n)})
        out = __import__("ctypes").c_int32()
        m3._ck(m3._lib.pmb_op_getfield(m3._ctx, p3.node, 10, 1, 100, 0, out))   # field outside the record


def test_convert_rule():
    """test/lazyexpression.jl:265-277: `convert(Vector, A * x)` == A * x as a Vector{AffineFunction} (copyto!, lazyexpression.jl:280-282)."""
    import parametron_b200 as pm
    rng = np.random.default_rng(22)
    Av = rng.integers(1, 4, size=(3, 3)).astype(float)
    m = pm.Model()
    x = m.variables(3)
    A = pm.Parameter(m, f=lambda: Av)
    e = pm.convert(A * x)
    pm.constraint(m, e, "==", np.zeros(3))
    pm.macros.constraint(m, "convert(Vector, x) >= l", x=x, l=np.ones(3))   # a Vector{Variable} converts too
    m.initialize(); m.update()
    Cb, db = m.block_arrays(0)
    views = [m.block_view(i) for i in range(m.num_blocks())]
    arrs = [m.block_arrays(i) for i in range(m.num_blocks())]
    assert any(np.array_equal(C_, Av) and np.array_equal(d_, np.zeros(3)) for C_, d_ in arrs)
    assert any(np.array_equal(C_, np.eye(3)) and np.array_equal(d_, -np.ones(3)) for C_, d_ in arrs)
    assert len(views) == 2


def test_variable_arithmetic_builds_affine_rows():
    """`x + 2y + 3z >= 4`, `x + y >= 1` written with operators on Variables (test/model.jl:288-290): each Variable lowers to an AffineFunction
    with one unit term; unused Parameters are never evaluated (the reference evaluates Parameters lazily, parameter.jl:93-102)."""
    import parametron_b200 as pm
    model = pm.Model()
    x, y, z = model.variables(3)
    calls = {"used": 0, "unused": 0}

    def f_used():
        calls["used"] += 1
        return 4.0

    def f_unused():
        calls["unused"] += 1
        return np.ones(3)
    rhs = pm.Parameter(model, f=f_used)
    pm.Parameter(model, f=f_unused)        # declared, reached by nothing
    pm.constraint(model, x + 2 * y + 3 * z, ">=", rhs)
    pm.constraint(model, x + y, ">=", 1.0)
    model.initialize()
    base_used, base_unused = calls["used"], calls["unused"]
    for _ in range(3):
        model.update()
    assert calls["used"] - base_used == 3 and calls["unused"] == base_unused
    blocks = sorted((model.block_arrays(i) for i in range(2)), key=lambda cd: cd[1][0])
    assert np.array_equal(blocks[0][0], [[1.0, 2.0, 3.0]]) and np.array_equal(blocks[0][1], [-4.0])
    assert np.array_equal(blocks[1][0], [[1.0, 1.0, 0.0]]) and np.array_equal(blocks[1][1], [-1.0])
    with pytest.raises(pm.Unsupported):
        model.add_integer_constraint(x)
    with pytest.raises(pm.Unsupported):
        x ** 2
