"""Pins oracle/functions.py on the RNG-free known answers of the reference's own tests
(/root/reference/test/functions.jl, test/util.jl).  CPU only."""
import numpy as np
import pytest

from oracle import functions as F
from oracle.functions import (AffineFunction, LinearTerm, QuadraticFunction, QuadraticTerm, Variable,
                              canonicalize, prune_zero)


def test_canonicalize_prune_zero():  # test/functions.jl:9-28
    x, y = Variable(1), Variable(2)
    assert canonicalize(3 * x * y) == canonicalize(3 * y * x) == QuadraticTerm(3, x, y)
    assert canonicalize(y + x - 2 * y + 3) == x - y + 3
    assert canonicalize(x * y + y * x + y + y + x - y + 4) == 2 * x * y + x + y + 4
    assert canonicalize(0 * y + x + 1) == x + 0 * y + 1
    assert prune_zero(canonicalize(0 * y + x + 1)) == x + 1
    assert canonicalize(0 * y**2 + x**2 + 0 * x + y + 1) == x**2 + 0 * y**2 + 0 * x + y + 1
    assert prune_zero(canonicalize(0 * y**2 + x**2 + 0 * x + y + 1)) == x**2 + y + 1


def test_docstring_examples():  # functions.jl:16-17, 285-289, 400-404, 317-323
    x, y = Variable(1), Variable(2)
    assert repr(3 * (x + y) + 4) == "3 * x1 + 3 * x2 + 4"
    f = y + x - 2 * y + 3
    assert repr(f) == "1 * x2 + 1 * x1 + -2 * x2 + 3"
    assert repr(canonicalize(f)) == "1 * x1 + -1 * x2 + 3"
    g = x * y + y * x + y + y + x - y
    assert repr(g) == "1 * x1 * x2 + 1 * x2 * x1 + 1 * x2 + 1 * x2 + 1 * x1 + -1 * x2 + 0"
    assert repr(canonicalize(g)) == "2 * x1 * x2 + 1 * x1 + 1 * x2 + 0"
    q = 2 * x**2 + 3 * x * y - 2 * y + 4
    assert repr(q) == "2 * x1 * x1 + 3 * x1 * x2 + -2 * x2 + 4"
    assert q({x: 4, y: -3}) == 6
    assert (2 * x + 3 * y + 4)({x: 4, y: -3}) == 3
    assert repr(canonicalize(3 * y * x)) == "3 * x1 * x2"


def test_linear_term():  # test/functions.jl:30-44
    x = Variable(1)
    assert LinearTerm(4.5, x) == 4.5 * x == x * 4.5
    assert +(2 * x) == 2 * x
    assert -(2 * x) == -2 * x
    assert 2 * (3 * x) == x * 6
    assert (2 * x) * 3 == 6 * x
    assert repr(LinearTerm(4.5, x)) == "4.5 * x1"


def test_quadratic_term():  # test/functions.jl:46-60
    x, y = Variable(1), Variable(2)
    assert x * y == QuadraticTerm(1, x, y)
    assert (2 * x) * y == QuadraticTerm(2, x, y)
    assert x * (2 * y) == QuadraticTerm(2, x, y)
    assert +(2 * x * y) == 2 * x * y
    assert -(2 * x * y) == -2 * x * y
    assert 2 * (3 * x * y) == x * 6 * y
    assert (y * 2 * x) * 3 == 6 * y * x
    assert repr(QuadraticTerm(1, x, y)) == "1 * x1 * x2"


def test_affine_function():  # test/functions.jl:62-98
    x, y = Variable(1), Variable(2)
    vals1 = {x: 1.0, y: 2.0}
    f1 = 2 * x + 3 * y + 5
    assert f1(vals1) == 2.0 + 3.0 * 2.0 + 5
    assert repr(f1) == "2 * x1 + 3 * x2 + 5"
    vals2 = {x: 2, y: -1}
    f2 = 2.5 * x + 4 * y + 1
    assert f2(vals2) == 2.5 * 2 - 4 + 1
    f3 = AffineFunction()
    assert f3(vals2) == 0.0
    f4 = f1 + f2
    assert f4(vals1) == f1(vals1) + f2(vals1)
    assert f4(vals2) == f1(vals2) + f2(vals2)
    f5 = f1 + 4.0
    assert f5(vals1) == f1(vals1) + 4.0
    f6 = 1 + f5
    assert f6(vals1) == 1 + f5(vals1)
    f7 = f6 - f4
    assert f7(vals1) == f6(vals1) - f4(vals1)
    f8 = f7 - LinearTerm(4, y)
    assert f8(vals1) == f7(vals1) - 4 * vals1[y]
    assert AffineFunction.of(x) == x + 0


def test_quadratic_function():  # test/functions.jl:100-146
    x = [Variable(1), Variable(2)]
    vals = dict(zip(x, [1.0, 2.0]))
    xv = np.array([1.0, 2.0])
    y = QuadraticFunction()
    assert F.dot(x, x)(vals) == xv @ xv
    assert F.vecdot_(y, x, x)(vals) == xv @ xv
    a = [4.0, 5.0]
    ax = [ai * xi for ai, xi in zip(a, x)]
    assert F.dot(x, ax)(vals) == xv @ (a * xv)
    assert F.vecdot_(y, ax, x)(vals) == (a * xv) @ xv
    assert F.vecdot_(y, ax, ax)(vals) == (a * xv) @ (a * xv)
    b = [6.0, 7.0]
    axb = [ai * xi + bi for ai, xi, bi in zip(a, x, b)]
    bxa = [bi * xi + ai for ai, xi, bi in zip(a, x, b)]
    assert F.dot(x, axb)(vals) == xv @ (a * xv + b)
    assert F.vecdot_(y, axb, x)(vals) == xv @ (a * xv + b)
    assert F.vecdot_(y, bxa, axb)(vals) == (b * xv + a) @ (a * xv + b)
    f1 = x[0]**2 + 2 * x[0] * x[1] + 3 * x[1] + 4
    f2 = x[1]**2 - 2 * x[0] + 3 * x[1] - 1
    assert (f1 - f2)(vals) == f1(vals) - f2(vals)
    f3 = (3 * x[0]**2) - x[1]
    assert f3(vals) == 3 * xv[0]**2 - xv[1]
    f4 = (3 * x[0] * x[1]) + (0.5 * x[1])
    assert f4(vals) == 3 * xv[0] * xv[1] + 0.5 * xv[1]
    f5 = (0.5 * x[1]) + (3 * x[0] * x[1])
    assert f5(vals) == 3 * xv[0] * xv[1] + 0.5 * xv[1]
    assert QuadraticFunction.of(x[0]) == QuadraticFunction.of(x[0] + 0)


def test_matvecmul_known_answer():  # test/functions.jl:148-163
    A = np.ones((3, 4), dtype=np.int64)
    x = [Variable(i) for i in range(1, 5)]
    y = [None] * 3
    F.matvecmul_(y, A, x)
    sumx = AffineFunction([LinearTerm(1, v) for v in x], 0)
    assert y == [sumx] * 3
    rng = np.random.default_rng(0)
    Bx = F.matvec(rng.random((4, 4)), x)
    y2 = [None] * 3
    F.matvecmul_(y2, A, Bx)
    s = AffineFunction()
    for f in Bx:
        F.muladd_(s, 1, f)
    assert y2 == [s] * 3


def test_mul_known_answers():  # test/functions.jl:165-199
    x = [Variable(i) for i in (1, 2, 3)]
    aff = F.dot([1, 2, 3], x) + 4
    quad = x[0]**2 + 2 * x[0] * x[2] + 3 * x[1] + 4
    dest = AffineFunction()
    for _ in range(2):
        F.mul_(dest, aff, 2)
        assert dest == 2 * x[0] + 4 * x[1] + 6 * x[2] + 8
        F.mul_(dest, 3, aff)
        assert dest == 3 * x[0] + 6 * x[1] + 9 * x[2] + 12
    dest = QuadraticFunction()
    for _ in range(2):
        F.mul_(dest, quad, 2)
        assert dest == 2 * x[0]**2 + 4 * x[0] * x[2] + 6 * x[1] + 8
        F.mul_(dest, 2, quad)
        assert dest == 2 * x[0]**2 + 4 * x[0] * x[2] + 6 * x[1] + 8
        expect = F.dot([1, 2, 3], [xi * x[0] for xi in x]) + 4 * x[0]
        F.mul_(dest, aff, x[0])
        assert dest == expect
        F.mul_(dest, x[0], aff)
        assert dest == expect
        F.mul_(dest, aff, aff)
        assert dest == (F.dot([1, 2, 3], x) + 4)**2


def test_dot_with_matrix_variables():  # test/functions.jl:201-206 (column-major eachindex)
    x = [Variable(i) for i in (1, 2, 3, 4)]  # x[i,j] with LinearIndices column-major
    w = np.array([[0.1, 0.2], [0.3, 0.4]])
    wl = [v.item() for v in w.ravel(order="F")]
    expect = 0.1 * x[0] + 0.3 * x[1] + 0.2 * x[2] + 0.4 * x[3] + 0.0
    assert F.dot(x, wl) == F.dot(wl, x) == expect


def test_matrix_operations():  # test/functions.jl:208-233
    x = [Variable(1), Variable(2)]
    A1 = np.array([[1.0, 2.0], [3.0, 4.0]])
    fs = F.matvec(A1, x)
    vals = dict(zip(x, [2.0, 5.0]))
    xv = np.array([2.0, 5.0])
    assert [f(vals) for f in fs] == list(A1 @ xv)
    a = [1, 2]
    assert F.dot(a, x) == F.dot(x, a) == a[0] * x[0] + a[1] * x[1]
    gs = [f + ai for f, ai in zip(fs, a)]
    gv = np.array([g(vals) for g in gs])
    assert F.dot(gs, gs)(vals) == gv @ gv
    assert F.dot(gs, x)(vals) == gv @ xv
    assert F.dot(x, gs)(vals) == xv @ gv
    h = F.dot(x, x)
    assert h(vals) == xv @ xv
    assert repr(h) == "1 * x1 * x1 + 1 * x2 * x2 + 0"  # term order pin


def test_vcat():  # test/functions.jl:235-253
    x, y = Variable(1), Variable(2)
    f1, f2, f3 = 3 * x + 10, 0.1 * y - 0.5, x + y
    v1, v2, v3 = [f1, f2], [f2, f3], [f3, f2, f1]
    z = lambda n: [AffineFunction() for _ in range(n)]
    assert F.vcat_(z(2), v1) == v1
    assert F.vcat_(z(4), v1, v2) == v1 + v2
    assert F.vcat_(z(7), v1, v2, v3) == v1 + v2 + v3
    for n, args in ((1, (v1,)), (3, (v1,)), (3, (v1, v2)), (5, (v1, v2)), (6, (v1, v2, v3)), (8, (v1, v2, v3))):
        with pytest.raises(F.DimensionMismatch):
            F.vcat_(z(n), *args)


def test_scalar_times_vector_and_zero():  # test/functions.jl:255-276
    x, y = Variable(1), Variable(2)
    assert [1 * v for v in (x, y)] == [v * 1 for v in (x, y)] == [LinearTerm(1, x), LinearTerm(1, y)]
    assert [x * v for v in (x, y)] == [x * x, x * y]
    assert canonicalize(F.dot([x, y], [x, y])) == x**2 + y**2


def test_sort_and_combine():  # test/util.jl:11-15
    comb = lambda a, b: (a[0], a[1] + b[1])
    first = lambda p: p[0]
    assert F.sort_and_combine([(2, 1.0), (2, 2.0)], combine=comb, by=first) == [(2, 3.0)]
    assert F.sort_and_combine([(3, 1.0), (2, 4.0), (3, 2.0), (1, 2.0)], combine=comb, by=first) == [(1, 2.0), (2, 4.0), (3, 3.0)]
    n = 100
    assert len(F.sort_and_combine([(i, 0.5) for i in range(n)], combine=comb, by=first)) == n
    assert len(F.sort_and_combine([(3, 0.5)] * n, combine=comb, by=first)) == 1


def test_bilinearmul_column_major_quirk():  # functions.jl:840-858
    x = [Variable(1), Variable(2)]
    Q = np.array([[1.0, 2.0], [3.0, 4.0]])
    d = F.bilinearmul_(QuadraticFunction(), Q, x, x)
    # k runs (row, col) = (1,1),(1,2),(2,1),(2,2) while Q[k] runs column-major: 1, 3, 2, 4
    assert [t.coeff for t in d.quadratic] == [1.0, 3.0, 2.0, 4.0]
    assert [(t.rowvar.index, t.colvar.index) for t in d.quadratic] == [(1, 1), (1, 2), (2, 1), (2, 2)]
