"""CPU checks of the certified int8-slice Gram's arithmetic (oracle/int8_slices.py restates csrc/kernels_gram_i8.cu in numpy):
digit planes are an exact fixed-point representation, the truncated plane products meet the certified error bound -- also on inputs
built to make every dropped digit product add up with the same sign -- and the level the certificate picks is the expected one."""
import numpy as np
import pytest

from oracle import int8_slices as S


def exact_gram(X):
    Xl = X.astype(np.longdouble)
    return np.asarray(Xl.T @ Xl, dtype=np.longdouble)


def cases():
    rng = np.random.default_rng(42)
    k, n = 257, 37
    yield "uniform", rng.random((k, n))
    yield "normal", rng.standard_normal((k, n))
    yield "wide range", rng.standard_normal((k, n)) * 10.0 ** (6.0 * rng.random((k, n)) - 3.0)
    yield "column scales", rng.standard_normal((k, n)) * 10.0 ** rng.integers(-8, 8, size=n)[None, :]
    yield "zero column", np.where(np.arange(n)[None, :] == 5, 0.0, rng.random((k, n)))
    # adversarial: every entry has lower digits -128 (so all dropped digit products are positive and add up), top digit 100;
    # one entry per column equal to 0.99 pins the scale at exactly 2^47
    Xi = 100 * 256 ** 5 + sum(-128 * 256 ** (5 - s) for s in range(1, 6))
    adv = np.full((k, n), Xi / 2.0 ** 47)
    adv[0, :] = 0.99
    yield "aligned digits", adv


@pytest.mark.parametrize("name,X", list(cases()), ids=[c[0] for c in cases()])
def test_planes_are_exact_fixed_point_and_bound_dominates_error(name, X):
    digits, sigma, stats = S.slice_columns(X)
    assert digits.min() >= -128 and digits.max() <= 127
    # representation: |x - sigma * sum_s d_s 2^(-7-8s)| <= 1.1 * 2^(-8 ns) * sigma  for ns = 5, 6
    for ns in (5, 6):
        err = np.abs(S.reconstruct(digits, sigma, ns) - X) / sigma[None, :]
        assert err.max() <= 1.1 * 2.0 ** (-8 * ns), (name, ns, err.max())
    G = exact_gram(X)
    for ns in (5, 6):
        Q, st = S.gram(X, ns, alpha=1.0)
        actual = float(np.abs(Q.astype(np.longdouble) - G).max())
        bound = S.error_bound(st, ns)
        assert actual <= bound, (name, ns, actual, bound)
        # the certificate is not vacuous: within a factor 2000 of the truth even on benign data (it is a worst-case bound)
        assert bound <= 2000 * max(actual, 1e-300 + 2.0 ** -52 * float(np.abs(G).max())), (name, ns, actual, bound)


def test_adversarial_bound_is_tight_within_an_order_of_magnitude():
    X = dict(cases())["aligned digits"]
    Q, st = S.gram(X, 5, alpha=1.0)
    actual = float(np.abs(Q.astype(np.longdouble) - exact_gram(X)).max())
    assert 0.05 * S.error_bound(st, 5) <= actual <= S.error_bound(st, 5)


def test_levels():
    c = dict(cases())
    assert S.level(S.slice_columns(c["uniform"])[2]) == 5
    assert S.level(S.slice_columns(c["normal"])[2]) in (5, 6)
    assert S.level(S.slice_columns(c["wide range"])[2]) in (6, 64)
    bad = c["uniform"].copy(); bad[3, 4] = np.inf
    assert S.level(S.slice_columns(bad)[2]) == 64
    tiny = c["uniform"].copy(); tiny[:, 2] *= 1e-200
    assert S.level(S.slice_columns(tiny)[2]) == 64      # outside [1e-150, 1e150]: fp64 kernel
    assert S.level(S.slice_columns(np.zeros((8, 3)))[2]) == 5
    # certified means: the int8 result is within tol of max |Q| -- the north_star tolerance is relative to the largest magnitude
    for name in ("uniform", "normal", "column scales"):
        st = S.slice_columns(c[name])[2]
        lv = S.level(st)
        if lv in (5, 6):
            Q, _ = S.gram(c[name], lv, alpha=1.0)
            G = exact_gram(c[name])
            assert float(np.abs(Q.astype(np.longdouble) - G).max()) <= 9e-11 * float(np.abs(G).max())
