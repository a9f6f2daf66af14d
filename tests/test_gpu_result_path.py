"""Result path (needs a B200: -m gpu): `value.(model, x)` gathers the solver's primal solution through
`model_var_to_optimizer` and `objectivevalue(model)` evaluates the objective that `update!` handed over (model.jl:166-173).  The
"solver" of these tests is numpy / scipy on the host acting on the device-assembled QP -- the external solve is out of scope."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def newton_solver(model):
    """an `optimizer` callable: unconstrained minimiser of 1/2 x'Qx + q'x + c0 in OPTIMIZER variable order"""
    Q, q, _ = model.objective_arrays()
    model.set_primal(np.linalg.solve(Q, -q))


def test_unconstrained_quadratic_value_and_objectivevalue():
    """test/model.jl:26-35 restated: min x'Qx + r.x + s;  x* = -(Q + Q')^-1 r;  objectivevalue == x'Qx + r.x + s (atol 1e-8)."""
    import parametron_b200 as pm
    n = 6
    rng = np.random.default_rng(1)
    perm = rng.permutation(n) + 1
    model = pm.Model(optimizer=newton_solver)
    x = model.variables(n)
    state = {}

    def newQ():
        M = rng.random((n, n))
        state["Q"] = M @ M.T + n * np.eye(n)
        return state["Q"]

    def newr():
        state["r"] = rng.random(n)
        return state["r"]

    Qp = pm.Parameter(model, f=newQ)
    rp = pm.Parameter(model, f=newr)
    sp = pm.Parameter(model, f=lambda: np.float64(1.5))
    pm.objective(model, pm.Minimize, pm.bilinear(pm.adjoint(x, model=model), Qp, x) + pm.dot(rp, x) + sp)
    model.initialize(idxmap=perm)
    for _ in range(3):
        model.solve()
        Q, r = state["Q"], state["r"]
        expected = -np.linalg.solve(Q + Q.T, r)
        xval = model.value(x)
        np.testing.assert_allclose(xval, expected, atol=1e-10)
        assert abs(model.value(x[2]) - expected[2]) <= 1e-10
        assert abs(model.objectivevalue() - (xval @ Q @ xval + r @ xval + 1.5)) <= 1e-8


def test_contquadratic_known_answer_through_an_external_qp_solver():
    """test/model.jl:281-300 (MOI contquadratic 1): min x^2 + xy + y^2 + yz + z^2  s.t. x + 2y + 3z >= 4, x + y >= 1
    -> objective 13/7 at (4/7, 3/7, 6/7) (atol 1e-6).  The QP the device assembled is solved on the host with scipy (SLSQP)."""
    import parametron_b200 as pm
    from scipy.optimize import minimize

    def scipy_qp(model):
        Q, q, c0 = model.objective_arrays()
        cons = []
        for pos in range(model.num_blocks()):
            Cb, db = model.block_arrays(pos)
            cons.append({"type": "ineq", "fun": lambda v, Cb=Cb, db=db: Cb @ v + db, "jac": lambda v, Cb=Cb: Cb})
        res = minimize(lambda v: 0.5 * v @ Q @ v + q @ v + c0, np.zeros(Q.shape[0]), jac=lambda v: Q @ v + q, constraints=cons, method="SLSQP",
                       options={"ftol": 1e-14, "maxiter": 200})
        model.set_primal(res.x)

    m = pm.Model(optimizer=scipy_qp)
    x, y, z = m.variables(3)
    # x^2 + xy + y^2 + yz + z^2 as a constant QuadraticFunction (functions.jl:326-329)
    qc = np.array([1.0, 1.0, 1.0, 1.0, 1.0]); rv = np.array([1, 1, 2, 2, 3], dtype=np.int64); cv = np.array([1, 2, 2, 3, 3], dtype=np.int64)
    out = C.c_int32()
    m._ck(m._lib.pmb_const_quadratic(m._ctx, 5, C.c_void_p(qc.ctypes.data), C.c_void_p(rv.ctypes.data), C.c_void_p(cv.ctypes.data),
                                     0, None, None, 0.0, out))
    pm.objective(m, pm.Minimize, pm.Expr(m, out.value))
    one = pm.Parameter(m, f=lambda: np.float64(1.0))   # keeps the blocks non-constant so that update! pushes them
    pm.constraint(m, one * pm.dot(np.array([1.0, 2.0, 3.0]), [x, y, z], model=m), ">=", 4.0)
    pm.constraint(m, one * pm.dot(np.array([1.0, 1.0]), [x, y], model=m), ">=", 1.0)
    m.solve()
    assert abs(m.objectivevalue() - 13 / 7) <= 1e-6
    np.testing.assert_allclose(m.value([x, y, z]), [4 / 7, 3 / 7, 6 / 7], atol=1e-6)


def test_value_errors_and_batched_gather():
    import parametron_b200 as pm
    n, B = 5, 3
    model = pm.Model(batch=B)
    x = model.variables(n)
    c = pm.Parameter(model, shape=(n,), device=True)
    pm.objective(model, pm.Minimize, pm.dot(c, x))
    perm = np.array([3, 1, 5, 2, 4])
    model.initialize(idxmap=perm)
    with pytest.raises(pm.PmbError):
        model.value(x)                      # no solution handed back yet
    cv = np.arange(B * n, dtype=float).reshape(B, n)
    c.upload(cv)
    model.update()
    sol = np.arange(B * n, dtype=float).reshape(B, n) * 0.5 + 1.0     # optimizer order
    model.set_primal(sol)
    got = model.value(x)
    assert np.array_equal(got, sol[:, perm - 1])
    # affine objective: q'x + c0 with q scattered through the index map
    np.testing.assert_allclose(model.objectivevalue(), np.einsum("bi,bi->b", cv, sol[:, perm - 1]), rtol=1e-14)
