"""Degenerate and ragged shapes through the C ABI (-m gpu), each against the literal oracle (tests/twin.py): the smallest models the
reference accepts (model.jl:18 default zero objective, one variable, one row), odd sizes that meet none of the 16-byte / tile alignments,
and a batch whose instances differ."""
import numpy as np
import pytest

from tests.test_gpu_parity import run_twin
from tests.twin import DeviceSide, assert_close

pytestmark = pytest.mark.gpu


def test_variables_only_model_updates_to_zero_objective():
    """no @objective, no @constraint: update! has nothing to push (model.jl:18: objective defaults to zero(AffineFunction))"""
    d = DeviceSide()
    d.variables(3)
    d.model.initialize()
    for _ in range(2):
        d.model.update()
    Q, q, c0 = d.model.objective_arrays()
    assert not np.any(q) and c0 == 0.0 and (Q is None or not np.any(Q))
    assert d.model.num_blocks() == 0


def test_one_variable_one_row():
    a, b, c, dd = ({"v": np.array([[2.5]])}, {"v": np.array([-1.5])}, {"v": np.array([[4.0]])}, {"v": np.array([3.0])})

    def build(s):
        x = s.variables(1)
        pa, pb, pc, pd = (s.parameter(h) for h in (a, b, c, dd))
        r = s.sub(s.mul(pa, x), pb)
        s.objective("min", s.dot(r, r))
        s.constraint(s.mul(pc, x), "==", pd)
        return 1

    def mutate(step):
        a["v"] = np.array([[2.5 + step]]); b["v"] = np.array([-1.5 * (step + 1)])

    o, d = run_twin(build, None, mutate)
    Q, q, c0 = d.model.objective_arrays()
    assert Q.shape == (1, 1) and Q[0, 0] == 2.0 * a["v"][0, 0] ** 2     # diagonal x2 of moi_interop.jl:58


@pytest.mark.parametrize("shape", [(3, 1, 1), (7, 13, 5), (33, 17, 9), (129, 97, 65), (1, 19, 3)])
def test_ragged_least_squares_shapes(shape):
    """(k, n, m) with k != n, odd everything: no TMA-aligned leading dimension, partial tiles on every edge"""
    k, n, m = shape
    rng = np.random.default_rng(k * 1000 + n)
    A, b, Cm, dv = ({"v": rng.standard_normal(s)} for s in ((k, n), (k,), (m, n), (m,)))

    def build(s):
        x = s.variables(n)
        pA, pb, pC, pd = (s.parameter(h) for h in (A, b, Cm, dv))
        r = s.sub(s.mul(pA, x), pb)
        s.objective("min", s.dot(r, r))
        s.constraint(s.mul(pC, x), "<=", pd)
        return n

    def mutate(step):
        A["v"] = rng.standard_normal((k, n)); dv["v"] = rng.standard_normal(m)

    run_twin(build, None, mutate, nupdates=2)


def test_batch_of_different_instances_matches_per_instance_oracle():
    """instance-shard semantics: instance i of a batch equals a batch-1 model fed with instance i's Parameters"""
    import parametron_b200 as pm
    from oracle import dense as D
    B, k, n, m = 5, 21, 11, 4
    rng = np.random.default_rng(99)
    A, b, Cm, dv = rng.standard_normal((B, k, n)), rng.standard_normal((B, k)), rng.standard_normal((B, m, n)), rng.standard_normal((B, m))
    model = pm.Model(batch=B)
    x = model.variables(n)
    pA, pb, pC, pd = (pm.Parameter(model, shape=s, device=True) for s in ((k, n), (k,), (m, n), (m,)))
    r = pA * x - pb
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, pC * x, "==", pd)
    model.initialize()
    for p, v in zip((pA, pb, pC, pd), (A, b, Cm, dv)):
        p.upload(v)
    model.update()
    Q, q, c0 = model.objective_arrays()
    Cb, db = model.block_arrays(0)
    for i in range(B):
        Qo, qo, c0o = D.least_squares_objective(A[i], b[i], np.arange(1, n + 1), n)
        assert_close(Q[i], Qo, 1e-10, f"Q[{i}]"); assert_close(q[i], qo, 1e-10, f"q[{i}]"); assert_close([c0[i]], [c0o], 1e-10, f"c0[{i}]")
        assert np.array_equal(Cb[i], Cm[i]) and np.array_equal(db[i], -dv[i])
