"""Pins oracle/termlist.c (the C term-list port timed as the CPU baseline) against the Python literal oracle, term by term:
same coefficients (bit-exact: identical operation order), same variable indices, same term order."""
import numpy as np
import pytest

from oracle import model as OM
from oracle.termlist import TermListModel
from tests.twin import OracleSide


@pytest.mark.parametrize("n,k,m,perm", [(8, 8, 2, False), (5, 7, 3, True), (1, 1, 1, False), (6, 4, 0, False)])
def test_termlist_matches_python_oracle(n, k, m, perm):
    rng = np.random.default_rng(n * 100 + k)
    hs = [{"v": rng.random(s)} for s in ((k, n), (k,), (m, n), (m,))]
    o = OracleSide()
    x = o.variables(n)
    A, b, Cm, d = (o.parameter(h) for h in hs)
    r = o.sub(o.mul(A, x), b)
    o.objective("min", o.dot(r, r))
    if m:
        o.constraint(o.mul(Cm, x), "==", d)
    varmap = [int(i) + 1 for i in rng.permutation(n)] if perm else None
    o.model.initialize(idxmap=({i + 1: varmap[i] for i in range(n)} if perm else None))
    t = TermListModel(n, k, m, varmap)
    for _ in range(2):
        for h in hs:
            h["v"] = rng.random(h["v"].shape)
        o.model.optimizer.log.clear()
        o.model.update()
        t.update(hs[0]["v"], hs[1]["v"], hs[2]["v"], hs[3]["v"])
        fq = o.model.optimizer.log[0][1]
        c, v1, v2 = t.quadratic_terms()
        assert len(fq.quadratic_terms) == k * n * n == len(c)
        assert [(float(a), int(i), int(j)) for a, i, j in zip(c, v1, v2)] == [(float(a), i, j) for a, i, j in fq.quadratic_terms]
        ca, va, const = t.affine_terms()
        assert [(float(a), int(i)) for a, i in zip(ca, va)] == [(float(a), i) for a, i in fq.affine_terms]
        assert const == fq.constant
        if m:
            fv = o.model.optimizer.log[1][2]
            rr, cc, vv, kk = t.vector_terms()
            assert [(int(a), float(b_), int(c_)) for a, b_, c_ in zip(rr, cc, vv)] == [(a, float(b_), c_) for a, b_, c_ in fv.terms]
            assert list(kk) == [float(z) for z in fv.constants]
    t.close()


def test_row_limit_is_a_prefix():
    """bench sampling: row_limit = R processes exactly the first R residual rows of vecdot!/the MOI copy."""
    n, k = 6, 5
    rng = np.random.default_rng(0)
    A, b = rng.random((k, n)), rng.random(k)
    full, part = TermListModel(n, k, 0), TermListModel(n, k, 0)
    full.update(A, b, np.zeros((0, n)), np.zeros(0))
    part.update(A, b, np.zeros((0, n)), np.zeros(0), row_limit=2)
    cf, pf = full.quadratic_terms(), part.quadratic_terms()
    assert len(pf[0]) == 2 * n * n
    for a, b_ in zip(cf, pf):
        assert np.array_equal(a[:2 * n * n], b_)
