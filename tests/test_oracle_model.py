"""Pins oracle/model.py (Parameter, rewrite table, MOI translation, update! order) on the
reference's tests: test/parameter.jl, test/lazyexpression.jl, test/model.jl.  CPU only."""
import numpy as np
import pytest

from oracle import functions as F
from oracle import model as M
from oracle import dense as D
from oracle.functions import AffineFunction, QuadraticFunction, Variable
from oracle.model import (Model, Parameter, expression, g_add, g_adjoint, g_convert, g_dot, g_mul, g_sub,
                          g_transpose, g_vcat, g_vect, g_getproperty)


def test_parameter_caching():  # test/parameter.jl:8-40
    model = Model()
    val = [True]
    p1 = Parameter(lambda: val[0], model)
    assert p1()
    val[0] = False
    assert p1()
    p1.setdirty()
    assert not p1()
    A = np.zeros((3, 4))
    aval = [1.0]

    def f(m):
        m[:] = aval[0]
    p2 = Parameter(f, model, val=A)
    assert p2() is A and np.all(p2() == 1)
    aval[0] = 2.0
    assert p2() is A and np.all(p2() == 1.0)
    p2.setdirty()
    assert np.all(p2() == 2.0)


def test_basics_and_nested():  # test/lazyexpression.jl:12-49
    model = Model()
    cval = [0]
    c = Parameter(lambda: cval[0], model)
    e = expression(g_add, 3, expression(g_mul, 4.0, c))
    assert e() == 3 + 4.0 * 0
    cval[0] = 4
    c.setdirty()
    assert e() == 3 + 4.0 * 4
    e2 = expression(g_mul, 4, e)
    assert e2() == 4 * e()


def test_matvecmul_rule():  # test/lazyexpression.jl:84-108
    m = Model()
    rng = np.random.default_rng(1)
    A = Parameter(lambda a: a.__setitem__(slice(None), rng.random(a.shape)), m, val=np.zeros((3, 4)))
    x = [Variable(i) for i in range(1, 5)]
    e = expression(g_mul, A, x)
    assert e.wrapped
    assert e() == F.matvec(A(), x)
    m.setdirty()
    r1 = e()
    assert r1 == F.matvec(A(), x)
    assert e() is r1  # result aliases the preallocated dest


def test_scale_rules():  # test/lazyexpression.jl:150-185
    model = Model()
    x = [model.variable() for _ in range(3)]
    dt = Parameter(lambda: 2.0, model)
    e = expression(g_mul, dt, np.array([1, 2, 3, 4, 5]))
    assert list(e()) == [2.0, 4.0, 6.0, 8.0, 10.0]
    e = expression(g_mul, dt, x)
    assert e() == [2.0 * v for v in x]
    e = expression(g_mul, x, dt)
    assert e() == [v * 2.0 for v in x]
    Ax = F.matvec(np.ones((3, 3)), x)
    e = expression(g_mul, dt, Ax)
    assert e() == [2.0 * f for f in Ax]
    e = expression(g_mul, Ax, dt)
    assert e() == [f * 2.0 for f in Ax]


def test_vcat_rule():  # test/lazyexpression.jl:188-245
    m = Model()
    rng = np.random.default_rng(42)
    fill = lambda a: a.__setitem__(slice(None), rng.random(a.shape))
    A = Parameter(fill, m, val=np.zeros((3, 4)))
    B = Parameter(fill, m, val=np.zeros((3, 3)))
    x = [Variable(i) for i in range(1, 5)]
    y = [Variable(i) for i in range(5, 8)]
    f1 = expression(g_mul, A, x)
    f2 = expression(g_mul, B, y)
    v1 = expression(g_vcat, f1, f2)
    assert v1() == f1() + f2()
    m.setdirty()
    assert v1() == F.matvec(A(), x) + F.matvec(B(), y)
    v4 = expression(g_vcat, f1)
    assert v4() == f1()


def test_vect_convert_vecdot_adjoint_rules():  # test/lazyexpression.jl:247-314
    m = Model()
    x = [Variable(1), Variable(2)]
    p = Parameter(lambda: np.array([1.0, 2.0]), m)
    e = expression(g_vect, expression(g_dot, p, x))
    assert e() == [F.dot([1.0, 2.0], x)]
    x3 = [Variable(i) for i in (1, 2, 3)]
    p3 = Parameter(lambda: np.array([1.0, 2.0, 3.0]), m)
    ex1 = expression(g_dot, x3, p3)
    ex2 = expression(g_dot, p3, x3)
    assert ex1() == ex2() == F.dot([1.0, 2.0, 3.0], x3)
    assert isinstance(ex1(), AffineFunction)
    ex3 = expression(g_dot, expression(g_add, x3, p3), p3)
    ex4 = expression(g_dot, p3, expression(g_add, x3, p3))
    xp = [xi + pi for xi, pi in zip(x3, [1.0, 2.0, 3.0])]
    assert ex3() == ex4() == F.dot([1.0, 2.0, 3.0], xp)
    assert isinstance(ex3(), AffineFunction)
    pm = Parameter(lambda: np.array([[1.0, 2.0], [3.0, 4.0]]), m)
    ex = expression(g_mul, expression(g_adjoint, pm), x)
    assert ex() == F.matvec(np.array([[1.0, 3.0], [2.0, 4.0]]), x)
    A = Parameter(lambda: np.ones((3, 3), dtype=np.int64), m)
    ec = expression(g_convert, list, expression(g_mul, A, x3))
    assert ec() == F.matvec(np.ones((3, 3), dtype=np.int64), x3)


def test_mul_rule():  # test/lazyexpression.jl:137-147
    m = Model()
    weight = Parameter(lambda: 3, m)
    x = [Variable(i) for i in (1, 2, 3)]
    e = expression(g_mul, weight, expression(g_dot, x, x))
    vals = dict(zip(x, [1, 2, 3]))
    assert e()(vals) == 3 * (1 + 4 + 9)


def test_issue_26():  # test/lazyexpression.jl:316-328
    model = Model()
    x = [Variable(1), Variable(2)]
    vals = dict(zip(x, [0.3, 0.9]))

    def fq(q):
        q[0] = 1
        q[1] = 2
    q = Parameter(fq, model, val=np.zeros(2))
    I2 = np.eye(2)
    quad = lambda: expression(g_mul, expression(g_transpose, x), I2, x)
    lin = lambda: expression(g_mul, expression(g_adjoint, q), x)
    e1 = expression(g_add, quad(), lin())
    e2 = expression(g_add, lin(), quad())
    assert abs(e1()(vals) - e2()(vals)) <= 1e-14
    assert abs(e1()(vals) - (0.3**2 + 0.9**2 + 0.3 + 1.8)) <= 1e-14


def test_issue_32_mpc_shape():  # test/lazyexpression.jl:330-347
    model = Model()
    v = [model.variable() for _ in range(2)]
    u = [model.variable() for _ in range(2)]
    v0 = np.zeros(2)
    dt = 0.01
    rng = np.random.default_rng(3)

    def fH(H):
        H[0, 0] = rng.random()
        H[1, 1] = rng.random()
    H = Parameter(fH, model, val=np.zeros((2, 2)))
    c = Parameter(lambda c: c.__setitem__(slice(None), rng.random(2)), model, val=np.zeros(2))
    e = expression(g_sub, expression(g_mul, H, expression(g_sub, v, v0)),
                   expression(g_mul, dt, expression(g_sub, u, c)))
    got = e()
    Hv, cv = H(), c()
    vv0 = [vi - 0.0 for vi in v]
    lhs = F.matvec(Hv, vv0)
    rhs = [dt * (ui - ci) for ui, ci in zip(u, cv)]
    assert got == [a - b for a, b in zip(lhs, rhs)]


def test_adjoint_times_vars():  # test/lazyexpression.jl:410-418
    model = Model()
    rng = np.random.default_rng(5)
    Adata = rng.standard_normal((5, 15))
    A = Parameter.manual(model, Adata)
    x = [model.variable() for _ in range(5)]
    resid = expression(g_mul, expression(g_adjoint, A), x)
    assert resid() == F.matvec(Adata.T.copy(), x)


def test_getfield_rule():  # test/lazyexpression.jl:364-382
    model = Model()

    class W:
        def __init__(self, x):
            self.x = x
    rng = np.random.default_rng(0)
    p = Parameter(lambda: W(rng.random()), model)
    e = expression(g_getproperty, p, "x")
    assert e() == p().x
    p.setdirty()
    assert e() == p().x
    e2 = expression(g_add, e, 1)
    p.setdirty()
    assert e2() == p().x + 1


# --------------------------------------------------------------------------------------------
# model.jl tests: the external solver is replaced by closed-form / KKT checks on the data the
# reference pushes through MOI.set.
# --------------------------------------------------------------------------------------------
def kkt_equality_qp(Q, q, C, d):
    """min 1/2 x'Qx + q'x  s.t.  Cx + d = 0."""
    n, m = Q.shape[0], C.shape[0]
    K = np.block([[Q, C.T], [C, np.zeros((m, m))]])
    sol = np.linalg.lstsq(K, np.concatenate([-q, -d]), rcond=None)[0]
    return sol[:n]


def readme_model(n, m, rng):
    model = Model()
    x = [model.variable() for _ in range(n)]
    fill = lambda a: a.__setitem__(slice(None), rng.random(a.shape))
    A = Parameter(fill, model, val=np.zeros((n, n)))
    b = Parameter(fill, model, val=np.zeros(n))
    C = Parameter(fill, model, val=np.zeros((m, n)))
    d = Parameter(fill, model, val=np.zeros(m))
    residual = expression(g_sub, expression(g_mul, A, x), b)
    model.setobjective(M.Minimize, expression(g_dot, residual, residual))
    model.constraint(expression(g_mul, C, x), "==", d)
    return model, x, (A, b, C, d)


def test_model_equality_constrained():  # test/model.jl:83-125 (the check, not Julia's RNG stream)
    n, m = 8, 2
    rng = np.random.default_rng(1234)
    model, x, (A, b, C, d) = readme_model(n, m, rng)
    for it in range(5):
        model.solve()
        log = model.optimizer.log
        assert [e[0] for e in log[-2:]] == ["objective", "constraint"]  # order of MOI.set calls
        Q, q, c0 = D.moi_quadratic_to_dense(log[-2][1], n)
        Cb, db = D.moi_vector_affine_to_dense(log[-1][2], n)
        assert all(p.evaluations == it + 2 for p in model.params)  # once per update (+1 at build time)
        Av, bv, Cv, dv = A(), b(), C(), d()
        # dense identities, test/model.jl:84
        np.testing.assert_allclose(Q, 2 * Av.T @ Av, rtol=1e-12)
        np.testing.assert_allclose(q, -2 * Av.T @ bv, rtol=1e-12)
        np.testing.assert_allclose(c0, bv @ bv, rtol=1e-12)
        np.testing.assert_array_equal(Cb, Cv)
        np.testing.assert_array_equal(db, -dv)
        xs = kkt_equality_qp(Q, q, Cb, db)
        Cp = np.linalg.pinv(Cv)
        P = np.eye(n) - Cp @ Cv
        expected = P @ (np.linalg.pinv(Av @ P) @ (bv - Av @ Cp @ dv)) + Cp @ dv
        np.testing.assert_allclose(xs, expected, rtol=1e-4)


def test_moi_term_order_and_counts():
    n, m = 3, 2
    rng = np.random.default_rng(7)
    model, x, (A, b, C, d) = readme_model(n, m, rng)
    model.initialize(idxmap={1: 3, 2: 1, 3: 2})  # a non-trivial optimizer variable permutation
    model.update()
    fq = model.optimizer.log[-2][1]
    assert len(fq.quadratic_terms) == n * n * n  # k * n^2, never combined (functions.jl:548-559)
    assert len(fq.affine_terms) == 2 * n * n
    Av = A()
    vm = [None, 3, 1, 2]
    # first residual row: i outer, j inner, diagonal doubled, varmap applied (moi_interop.jl:53-60)
    k = 0
    for i in range(n):
        for j in range(n):
            coeff, vi, vj = fq.quadratic_terms[k]
            expect = Av[0, i] * Av[0, j] * (2 if i == j else 1)
            assert (coeff, vi, vj) == (expect, vm[i + 1], vm[j + 1])
            k += 1
    fv = model.optimizer.log[-1][2]
    Cv = C()
    assert [(r, v) for r, _, v in fv.terms] == [(r, vm[c]) for r in (1, 2) for c in (1, 2, 3)]  # row-major flatten
    assert [c for _, c, _ in fv.terms] == [Cv[r, c] for r in range(m) for c in range(n)]
    Q, q, c0 = D.moi_quadratic_to_dense(fq, n)
    Qd, qd, c0d = D.least_squares_objective(Av, b(), [1, 2, 3], n, varmap=[3, 1, 2])
    np.testing.assert_allclose(Q, Qd, rtol=1e-13)
    np.testing.assert_allclose(q, qd, rtol=1e-13)
    Cb, db = D.moi_vector_affine_to_dense(fv, n)
    Cd, dd = D.affine_block(Cv, d(), [1, 2, 3], n, varmap=[3, 1, 2])
    np.testing.assert_array_equal(Cb, Cd)
    np.testing.assert_array_equal(db, dd)


def test_model_unconstrained():  # test/model.jl:26-81
    model = Model()
    n = 2
    x = [model.variable() for _ in range(n)]
    assert [v.index for v in x] == [1, 2]
    rng = np.random.default_rng(1)
    sval = [1.0]

    def fQ(Q):
        Q[0, 0] = rng.random()
        Q[1, 1] = rng.random()
    Q = Parameter(fQ, model, val=np.eye(n))
    r = Parameter(lambda r: r.__setitem__(slice(None), rng.random(n)), model, val=np.zeros(n))
    s = Parameter(lambda: sval[0], model)
    e = expression(g_add, expression(g_mul, expression(g_transpose, x), Q, x), expression(g_dot, r, x), s)
    model.setobjective(M.Minimize, e)
    model.initialize()
    for sv in (1.0, 2.0):
        sval[0] = sv
        model.solve()
        H, q, c0 = D.moi_quadratic_to_dense(model.optimizer.log[-1][1], n)
        xs = np.linalg.solve(H, -q)
        np.testing.assert_allclose(xs, -np.linalg.solve(2 * Q(), r()), atol=1e-8)
        obj = 0.5 * xs @ H @ xs + q @ xs + c0
        np.testing.assert_allclose(obj, xs @ Q() @ xs + r() @ xs + s(), atol=1e-8)


def test_model_box_constrained():  # test/model.jl:127-171 (KKT check of the known answer x = p/2)
    n = 10
    model = Model()
    x = [model.variable() for _ in range(n)]
    rng = np.random.default_rng(1234)
    l = Parameter(lambda l: l.__setitem__(slice(None), -rng.random(n)), model, val=np.zeros(n))
    u = Parameter(lambda u: u.__setitem__(slice(None), rng.random(n)), model, val=np.zeros(n))
    choose = [True]

    def fp(p):
        choose[0] = bool(rng.integers(2))
        p[:] = 2 * (l() if choose[0] else u())
    p = Parameter(fp, model, val=np.zeros(n))
    residual = expression(g_sub, x, p)
    model.setobjective(M.Minimize, expression(g_sub, expression(g_dot, residual, residual), expression(g_dot, p, p)))
    model.constraint(x, ">=", l)
    model.constraint(x, "<=", u)
    for _ in range(5):
        model.solve()
        log = model.optimizer.log[-3:]
        assert [e[0] for e in log] == ["objective", "constraint", "constraint"]
        H, q, c0 = D.moi_quadratic_to_dense(log[0][1], n)
        Cl, dl = D.moi_vector_affine_to_dense(log[1][2], n)  # x - l in Nonnegatives
        Cu, du = D.moi_vector_affine_to_dense(log[2][2], n)  # x - u in Nonpositives
        xs = p() / 2
        np.testing.assert_allclose(H, 2 * np.eye(n))
        np.testing.assert_allclose(c0, 0.0, atol=1e-12)
        grad = H @ xs + q  # = -p: pushes towards p, blocked by the active bound
        assert np.all(Cl @ xs + dl >= -1e-12) and np.all(Cu @ xs + du <= 1e-12)
        active = (Cl @ xs + dl) if choose[0] else (Cu @ xs + du)
        np.testing.assert_allclose(active, 0, atol=1e-12)
        assert np.all(grad >= 0) if choose[0] else np.all(grad <= 0)


def test_constant_constraints_skipped():  # test/model.jl:281-300, moi_interop.jl:123,132,153,169
    model = Model()
    x, y, z = (model.variable() for _ in range(3))
    model.setobjective(M.Minimize, x**2 + x * y + y**2 + y * z + z**2)
    model.constraint(x + 2 * y + 3 * z, ">=", 4)
    model.constraint(x + y, ">=", 1)
    model.solve()
    assert model.optimizer.log == []  # everything constant: nothing is pushed by update!
    H, q, c0 = D.moi_quadratic_to_dense(model.objective.f, 3)
    np.testing.assert_array_equal(H, [[2, 1, 0], [1, 2, 1], [0, 1, 2]])
    xs = np.array([4 / 7, 3 / 7, 6 / 7])
    assert abs(0.5 * xs @ H @ xs + q @ xs + c0 - 13 / 7) < 1e-12
    c1 = model.constraints[0][0].f
    assert sorted(c1.terms) == [(1, 1), (2, 2), (3, 3)] and c1.constant == -4
    # KKT: both constraints active at the known answer with non-negative multipliers
    Cm = np.array([[1.0, 2, 3], [1, 1, 0]])
    lam = np.linalg.lstsq(Cm.T, H @ xs + q, rcond=None)[0]
    np.testing.assert_allclose(Cm.T @ lam, H @ xs + q, atol=1e-12)
    assert np.all(lam >= 0)


def test_errors_after_initialize():  # model.jl:50,61,69
    model = Model()
    x = model.variable()
    model.initialize()
    with pytest.raises(RuntimeError):
        model.variable()
    with pytest.raises(RuntimeError):
        model.setobjective(M.Minimize, x**2)
    with pytest.raises(RuntimeError):
        model.constraint(x, ">=", 0)
