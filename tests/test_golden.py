"""Golden fixtures (tests/golden/*.npz, generator tests/golden/make_golden.py): seeded inputs and the densified output of the literal
oracle for four model shapes of the reference's own tests.  CPU: the oracle still reproduces them exactly and they carry the dense
meaning of SURVEY.md section 8a (Q = 2 A'A, q = -2 A'b, c0 = b'b, block (C, -d)).  GPU (-m gpu): the CUDA path reproduces them through
the C ABI within the north_star tolerance (1e-10 relative on fp64 coefficient arrays, == on the term bookkeeping)."""
import os

import numpy as np
import pytest

from tests.golden import make_golden as G
from tests.golden import models as M
from tests.twin import DeviceSide, assert_close

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-10


def load(name):
    return dict(np.load(os.path.join(HERE, name + ".npz")))


@pytest.mark.parametrize("name", M.NAMES)
def test_oracle_reproduces_golden(name):
    want, got = load(name), G.build(name)
    assert sorted(want) == sorted(got)
    for k in want:
        assert np.array_equal(np.asarray(want[k]), np.asarray(got[k])), k


def test_readme_golden_has_the_dense_meaning():
    g = load("readme")
    A, b, Cm, d = g["in_A"], g["in_b"], g["in_C"], g["in_d"]
    assert_close(g["Q"], 2.0 * A.T @ A, 1e-13, "Q")        # moi_interop.jl:45-62 on functions.jl:548-576 (reference comment test/model.jl:84)
    assert_close(g["q"], -2.0 * A.T @ b, 1e-13, "q")
    assert_close([g["c0"]], [b @ b], 1e-13, "c0")
    assert np.array_equal(g["C0"], Cm) and np.array_equal(g["d0"], -d)
    # row-outer / term-inner order of moi_interop.jl:64-81
    assert list(g["term_rows0"]) == [r for r in (1, 2) for _ in range(8)] and list(g["term_vars0"]) == list(range(1, 9)) * 2


@pytest.mark.gpu
@pytest.mark.parametrize("name", M.NAMES)
def test_device_reproduces_golden(name):
    g = load(name)
    holders = {k[3:]: {"v": g[k]} for k in g if k.startswith("in_")}
    d = DeviceSide()
    M.declare(name, d, holders)
    d.model.initialize()
    for _ in range(2):   # a second update! of unchanged Parameters gives the same arrays
        d.model.update()
        Q, q, c0 = d.model.objective_arrays()
        if "Q" in g:
            assert_close(Q, g["Q"], RTOL, "Q")
        assert_close(q, g["q"], RTOL, "q")
        assert_close([c0], [g["c0"]], RTOL, "c0")
        pos = [p for p in range(d.model.num_blocks()) if not d.model.block_view(p).is_constant]
        assert len(pos) == int(g["nblocks"])
        for i, p in enumerate(pos):
            Cd, dd = d.model.block_arrays(p)
            assert_close(Cd, g[f"C{i}"], RTOL, f"C{i}")
            assert_close(dd, g[f"d{i}"], RTOL, f"d{i}")
            rows, vars_, _ = d.model.block_terms(p)
            assert np.array_equal(rows, g[f"term_rows{i}"]) and np.array_equal(vars_, g[f"term_vars{i}"])   # bit-exact bookkeeping
