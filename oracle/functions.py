"""ORACLE (test infrastructure only -- never imported by the product path).

Literal CPU restatement of Parametron.jl's symbolic value types and in-place
kernels, `/root/reference/src/functions.jl` and `src/util.jl`.  Term lists are
kept in exactly the order the reference produces them (duplicates are never
combined on the hot path), so results can be compared term by term.

Parity status: PINNED by the RNG-free known answers of the reference's own
tests (`test/functions.jl`, `test/util.jl`), restated in
`tests/test_oracle_functions.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.
"""
from __future__ import annotations

import numbers
from typing import Dict, Iterable, List, Sequence

Number = numbers.Number


def _isnum(x) -> bool:
    return isinstance(x, numbers.Number) and not isinstance(x, bool) or isinstance(x, bool)


# ---------------------------------------------------------------------------
# Variable / LinearTerm / QuadraticTerm            functions.jl:96-162
# ---------------------------------------------------------------------------
class Variable:
    """functions.jl:96-98"""
    __slots__ = ("index",)

    def __init__(self, index: int):
        self.index = int(index)

    def __eq__(self, o):
        return isinstance(o, Variable) and o.index == self.index

    def __hash__(self):
        return hash(("Variable", self.index))

    def __repr__(self):
        return f"x{self.index}"

    # functions.jl:101,122
    def __pos__(self):
        return self

    def __neg__(self):
        return LinearTerm(-1, self)

    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return sub(self, o)

    def __rsub__(self, o):
        return sub(o, self)

    def __mul__(self, o):
        return mul(self, o)

    def __rmul__(self, o):
        return mul(o, self)

    def __pow__(self, p):
        return power(self, p)


class LinearTerm:
    """functions.jl:110-113"""
    __slots__ = ("coeff", "var")

    def __init__(self, coeff, var: Variable):
        self.coeff = coeff
        self.var = var

    def __eq__(self, o):
        return isinstance(o, LinearTerm) and o.coeff == self.coeff and o.var == self.var

    def __hash__(self):
        return hash((self.coeff, self.var))

    def __repr__(self):  # functions.jl:117
        return f"{_show(self.coeff)} * x{self.var.index}"

    def __pos__(self):
        return self

    def __neg__(self):  # functions.jl:158
        return LinearTerm(-self.coeff, self.var)

    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return sub(self, o)

    def __rsub__(self, o):
        return sub(o, self)

    def __mul__(self, o):
        return mul(self, o)

    def __rmul__(self, o):
        return mul(o, self)

    def __pow__(self, p):
        return power(self, p)


class QuadraticTerm:
    """functions.jl:136-140"""
    __slots__ = ("coeff", "rowvar", "colvar")

    def __init__(self, coeff, rowvar: Variable, colvar: Variable):
        self.coeff = coeff
        self.rowvar = rowvar
        self.colvar = colvar

    def __eq__(self, o):
        return (isinstance(o, QuadraticTerm) and o.coeff == self.coeff
                and o.rowvar == self.rowvar and o.colvar == self.colvar)

    def __hash__(self):
        return hash((self.coeff, self.rowvar, self.colvar))

    def __repr__(self):  # functions.jl:143
        return f"{_show(self.coeff)} * x{self.rowvar.index} * x{self.colvar.index}"

    def __pos__(self):
        return self

    def __neg__(self):
        return QuadraticTerm(-self.coeff, self.rowvar, self.colvar)

    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return sub(self, o)

    def __rsub__(self, o):
        return sub(o, self)

    def __mul__(self, o):
        return mul(self, o)

    def __rmul__(self, o):
        return mul(o, self)


def _show(c):
    """Julia prints Int coefficients without a decimal point and Float64 with one."""
    if isinstance(c, bool):
        return str(c).lower()
    if isinstance(c, int):
        return str(c)
    f = float(c)
    return repr(f)


def canonicalize_term(term: QuadraticTerm) -> QuadraticTerm:
    """functions.jl:182-184"""
    lo, hi = sorted((term.rowvar.index, term.colvar.index))
    return QuadraticTerm(term.coeff, Variable(lo), Variable(hi))


def combine_linear(t1: LinearTerm, t2: LinearTerm) -> LinearTerm:
    """functions.jl:126-129"""
    if t1.var != t2.var:
        raise ValueError("ArgumentError")
    return LinearTerm(t1.coeff + t2.coeff, t1.var)


def combine_quadratic(t1: QuadraticTerm, t2: QuadraticTerm) -> QuadraticTerm:
    """functions.jl:186-191"""
    t1 = canonicalize_term(t1)
    t2 = canonicalize_term(t2)
    if not (t1.rowvar == t2.rowvar and t1.colvar == t2.colvar):
        raise ValueError("ArgumentError")
    return QuadraticTerm(t1.coeff + t2.coeff, t1.rowvar, t1.colvar)


def sort_and_combine(v: list, *, combine, by=lambda x: x) -> list:
    """util.jl:9-26.  In place.  The reference sorts with (unstable) QuickSort; the
    order in which equal-key terms are summed is therefore unspecified there.  A
    stable sort is one admissible order."""
    if not v:
        return v
    v.sort(key=by)
    j = 0
    for i in range(1, len(v)):
        x = v[i]
        if by(v[j]) < by(x):
            j += 1
            v[j] = x
        else:
            v[j] = combine(v[j], x)
    del v[j + 1:]
    return v


# ---------------------------------------------------------------------------
# AffineFunction                                    functions.jl:218-299
# ---------------------------------------------------------------------------
class AffineFunction:
    __slots__ = ("linear", "constant")

    def __init__(self, linear: Iterable[LinearTerm] = (), constant=0):
        self.linear: List[LinearTerm] = list(linear)  # functions.jl:223-224 copies
        self.constant = constant

    @staticmethod
    def of(x) -> "AffineFunction":
        """Converting constructors functions.jl:228-235."""
        if isinstance(x, AffineFunction):
            return AffineFunction(x.linear, x.constant)
        if isinstance(x, LinearTerm):
            return AffineFunction([x], 0)
        if isinstance(x, Variable):
            return AffineFunction([LinearTerm(1, x)], 0)
        if _isnum(x):
            return AffineFunction([], x)
        raise TypeError(type(x))

    def __eq__(self, o):  # functions.jl:239
        return isinstance(o, AffineFunction) and self.linear == o.linear and self.constant == o.constant

    def __hash__(self):
        return hash((tuple(self.linear), self.constant))

    def __repr__(self):  # functions.jl:252-257
        return "".join(f"{t} + " for t in self.linear) + _show(self.constant)

    def __call__(self, vals: Dict[Variable, Number]):  # functions.jl:259-267
        ret = self.constant
        for term in self.linear:
            ret = ret + term.coeff * vals[term.var]
        return ret

    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return sub(self, o)

    def __rsub__(self, o):
        return sub(o, self)

    def __mul__(self, o):
        return mul(self, o)

    def __rmul__(self, o):
        return mul(o, self)

    def __pow__(self, p):
        return power(self, p)


def zero_(f):
    """zero! functions.jl:244,355"""
    if isinstance(f, AffineFunction):
        f.linear.clear()
        f.constant = 0
    else:
        f.quadratic.clear()
        zero_(f.affine)
    return f


# ---------------------------------------------------------------------------
# QuadraticFunction                                 functions.jl:326-415
# ---------------------------------------------------------------------------
class QuadraticFunction:
    __slots__ = ("quadratic", "affine")

    def __init__(self, quadratic: Iterable[QuadraticTerm] = (), affine: AffineFunction | None = None):
        self.quadratic: List[QuadraticTerm] = list(quadratic)
        self.affine = affine if affine is not None else AffineFunction()

    @staticmethod
    def of(x) -> "QuadraticFunction":
        """Converting constructors functions.jl:334-346."""
        if isinstance(x, QuadraticFunction):
            return QuadraticFunction(x.quadratic, AffineFunction.of(x.affine))
        if isinstance(x, QuadraticTerm):
            return QuadraticFunction([x], AffineFunction())
        return QuadraticFunction([], AffineFunction.of(x))

    def __eq__(self, o):  # functions.jl:350
        return isinstance(o, QuadraticFunction) and self.quadratic == o.quadratic and self.affine == o.affine

    def __hash__(self):
        return hash((tuple(self.quadratic), self.affine))

    def __repr__(self):  # functions.jl:366-371
        return "".join(f"{t} + " for t in self.quadratic) + repr(self.affine)

    def __call__(self, vals):  # functions.jl:373-379
        ret = self.affine(vals)
        for term in self.quadratic:
            ret = ret + term.coeff * vals[term.rowvar] * vals[term.colvar]
        return ret

    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return sub(self, o)

    def __rsub__(self, o):
        return sub(o, self)

    def __mul__(self, o):
        return mul(self, o)

    def __rmul__(self, o):
        return mul(o, self)


SCALAR_FUNCTIONS = (Variable, LinearTerm, QuadraticTerm, AffineFunction, QuadraticFunction)


def canonicalize_(f):
    """canonicalize! functions.jl:269-272, 381-386"""
    if isinstance(f, AffineFunction):
        sort_and_combine(f.linear, by=lambda t: t.var.index, combine=combine_linear)
        return f
    canonicalize_(f.affine)

    def key(term):
        t = canonicalize_term(term)
        return (t.rowvar.index, t.colvar.index)

    sort_and_combine(f.quadratic, by=key, combine=combine_quadratic)
    return f


def canonicalize(f):
    """functions.jl:182, 292, 407"""
    if isinstance(f, QuadraticTerm):
        return canonicalize_term(f)
    if isinstance(f, AffineFunction):
        return canonicalize_(AffineFunction.of(f))
    return canonicalize_(QuadraticFunction.of(f))


def prune_zero_(f, atol=0):
    """prune_zero! functions.jl:294-297, 409-413 (the quadratic method prunes its affine part with atol=0)."""
    if isinstance(f, AffineFunction):
        f.linear[:] = [t for t in f.linear if abs(t.coeff) > atol]
        return f
    prune_zero_(f.affine)
    f.quadratic[:] = [t for t in f.quadratic if abs(t.coeff) > atol]
    return f


def prune_zero(f, atol=0):
    if isinstance(f, AffineFunction):
        return prune_zero_(AffineFunction.of(f), atol)
    return prune_zero_(QuadraticFunction.of(f), atol)


# ---------------------------------------------------------------------------
# copyto! / add! / subtract! / muladd! / mul!       functions.jl:419-578
# ---------------------------------------------------------------------------
def copyto_(f, x):
    """functions.jl:419-439"""
    if isinstance(f, AffineFunction):
        if _isnum(x):
            zero_(f)
            f.constant = x
        elif isinstance(x, LinearTerm):
            zero_(f)
            f.linear.append(x)
        elif isinstance(x, Variable):
            zero_(f)
            f.linear.append(LinearTerm(1, x))
        elif isinstance(x, AffineFunction):
            f.linear[:] = x.linear
            f.constant = x.constant
        else:
            raise TypeError(type(x))
        return f
    if isinstance(x, QuadraticTerm):
        zero_(f)
        f.quadratic.append(x)
    elif isinstance(x, QuadraticFunction):
        f.quadratic[:] = x.quadratic
        copyto_(f.affine, x.affine)
    else:
        f.quadratic.clear()
        copyto_(f.affine, x)
    return f


def add_(f, x, y=None):
    """add! functions.jl:452-461 (2-arg: f += x; 3-arg: dest = x + y)."""
    if y is not None:
        copyto_(f, x)
        add_(f, y)
        return f
    if isinstance(f, AffineFunction):
        if _isnum(x):
            f.constant = f.constant + x
        elif isinstance(x, Variable):
            f.linear.append(LinearTerm(1, x))
        elif isinstance(x, LinearTerm):
            f.linear.append(x)
        elif isinstance(x, AffineFunction):
            f.linear.extend(x.linear)
            f.constant = f.constant + x.constant
        else:
            raise TypeError(type(x))
        return f
    if isinstance(x, QuadraticTerm):
        f.quadratic.append(x)
    elif isinstance(x, QuadraticFunction):
        f.quadratic.extend(x.quadratic)
        add_(f.affine, x.affine)
    else:
        add_(f.affine, x)
    return f


def subtract_(f, x, y=None):
    """subtract! functions.jl:474-502"""
    if y is not None:
        copyto_(f, x)
        subtract_(f, y)
        return f
    if isinstance(f, AffineFunction):
        if _isnum(x):
            f.constant = f.constant - x
        elif isinstance(x, Variable):
            f.linear.append(LinearTerm(-1, x))
        elif isinstance(x, LinearTerm):
            f.linear.append(-x)
        elif isinstance(x, AffineFunction):
            f.linear.extend(-t for t in x.linear)
            f.constant = f.constant - x.constant
        else:
            raise TypeError(type(x))
        return f
    if isinstance(x, QuadraticTerm):
        f.quadratic.append(-x)
    elif isinstance(x, QuadraticFunction):
        f.quadratic.extend(-t for t in x.quadratic)
        subtract_(f.affine, x.affine)
    else:
        subtract_(f.affine, x)
    return f


def muladd_(dest, x, y):
    """muladd! functions.jl:515-576: dest += x * y."""
    if isinstance(dest, AffineFunction):
        if _isnum(x) and isinstance(y, AffineFunction):  # :524
            x, y = y, x
        if isinstance(x, AffineFunction) and _isnum(y):  # :515-523
            dest.linear.extend(t * y for t in x.linear)
            dest.constant = dest.constant + x.constant * y
            return dest
        raise TypeError((type(x), type(y)))
    # QuadraticFunction destination
    if _isnum(x) and isinstance(y, QuadraticFunction):  # :535
        x, y = y, x
    if isinstance(x, QuadraticFunction) and _isnum(y):  # :526-534
        dest.quadratic.extend(t * y for t in x.quadratic)
        muladd_(dest.affine, x.affine, y)
        return dest
    if isinstance(x, (Variable, LinearTerm)) and isinstance(y, AffineFunction):  # :546
        x, y = y, x
    if isinstance(x, AffineFunction) and isinstance(y, (Variable, LinearTerm)):  # :537-545
        dest.quadratic.extend(t * y for t in x.linear)
        add_(dest.affine, x.constant * y)
        return dest
    if isinstance(x, AffineFunction) and isinstance(y, AffineFunction):  # :548-576
        for xi in x.linear:
            for yj in y.linear:
                dest.quadratic.append(xi * yj)
        xconst, yconst = x.constant, y.constant
        aff = dest.affine
        aff.linear.extend(t * yconst for t in x.linear)
        aff.linear.extend(t * xconst for t in y.linear)
        aff.constant = aff.constant + xconst * yconst
        return dest
    raise TypeError((type(x), type(y)))


def mul_(dest, x, y):
    """mul! functions.jl:578"""
    zero_(dest)
    return muladd_(dest, x, y)


# ---------------------------------------------------------------------------
# Operator overloads (generic, allocating path)     functions.jl:581-652
# ---------------------------------------------------------------------------
def _lt(x):
    return LinearTerm(1, x) if isinstance(x, Variable) else x


def _addsub(x, y, neg: bool):
    fun_ = subtract_ if neg else add_
    sgn = (lambda t: -t) if neg else (lambda t: t)
    if _isnum(x) and _isnum(y):
        return x - y if neg else x + y
    xq = isinstance(x, (QuadraticTerm, QuadraticFunction))
    yq = isinstance(y, (QuadraticTerm, QuadraticFunction))
    if xq or yq:
        if isinstance(x, QuadraticTerm) and isinstance(y, QuadraticTerm):  # :599
            return QuadraticFunction([x, sgn(y)], AffineFunction())
        return fun_(QuadraticFunction.of(x), y)  # :601-614
    x, y = _lt(x) if not _isnum(x) else x, _lt(y) if not _isnum(y) else y
    if isinstance(x, LinearTerm) and isinstance(y, LinearTerm):  # :583
        return AffineFunction([x, sgn(y)], 0)
    if isinstance(x, LinearTerm) and _isnum(y):  # :586
        return AffineFunction([x], sgn(y))
    if _isnum(x) and isinstance(y, LinearTerm):  # :587
        return AffineFunction([sgn(y)], x)
    return fun_(AffineFunction.of(x), y)  # :592-597


def add(x, y):
    return _addsub(x, y, False)


def sub(x, y):
    return _addsub(x, y, True)


def mul(x, y):
    """`*` functions.jl:120-121,146-149,159-160,618-626"""
    if _isnum(x) and _isnum(y):
        return x * y
    if _isnum(x):
        if isinstance(y, Variable):
            return LinearTerm(x, y)
        if isinstance(y, LinearTerm):
            return LinearTerm(x * y.coeff, y.var)
        if isinstance(y, QuadraticTerm):
            return QuadraticTerm(x * y.coeff, y.rowvar, y.colvar)
        if isinstance(y, AffineFunction):
            return muladd_(AffineFunction(), x, y)
        if isinstance(y, QuadraticFunction):
            return muladd_(QuadraticFunction(), x, y)
    if _isnum(y):
        if isinstance(x, Variable):
            return LinearTerm(y, x)
        if isinstance(x, (LinearTerm, QuadraticTerm)):
            return mul(y, x)  # :160  term * c = c * term
        if isinstance(x, AffineFunction):
            return muladd_(AffineFunction(), x, y)
        if isinstance(x, QuadraticFunction):
            return muladd_(QuadraticFunction(), x, y)
    if isinstance(x, Variable) and isinstance(y, Variable):
        return QuadraticTerm(1, x, y)
    if isinstance(x, Variable) and isinstance(y, LinearTerm):
        return QuadraticTerm(y.coeff, x, y.var)
    if isinstance(x, LinearTerm) and isinstance(y, Variable):
        return QuadraticTerm(x.coeff, x.var, y)
    if isinstance(x, LinearTerm) and isinstance(y, LinearTerm):
        return QuadraticTerm(x.coeff * y.coeff, x.var, y.var)
    if isinstance(x, AffineFunction) or isinstance(y, AffineFunction):
        if isinstance(x, (Variable, LinearTerm, AffineFunction)) and isinstance(y, (Variable, LinearTerm, AffineFunction)):
            return muladd_(QuadraticFunction(), x, y)
    raise TypeError(f"unsupported product {type(x).__name__} * {type(y).__name__}")


def power(x, p: int):
    """`^` functions.jl:628-630 (power_by_squaring); only p in {1, 2} yields a representable function."""
    if p == 1:
        return x
    if p == 2:
        return mul(x, x)
    raise ValueError("only powers 1 and 2 are representable")


# ---------------------------------------------------------------------------
# Array kernels                                     functions.jl:656-994
# ---------------------------------------------------------------------------
class DimensionMismatch(Exception):
    pass


def _is_simple(v):
    return _isnum(v) or isinstance(v, (Variable, LinearTerm))


def vecdot_(dest, x: Sequence, y: Sequence):
    """vecdot! / _vecdot! functions.jl:665-731"""
    zero_(dest)
    if len(x) != len(y):
        raise DimensionMismatch()
    if isinstance(dest, AffineFunction):
        if all(_is_simple(v) for v in x) and all(_is_simple(v) for v in y):  # :676-687
            dest.linear.extend(mul(xi, yi) for xi, yi in zip(x, y))
        else:  # :665-674
            for xi, yi in zip(x, y):
                muladd_(dest, xi, yi)
        return dest
    simpleq = lambda v: _isnum(v) or isinstance(v, (Variable, LinearTerm, QuadraticTerm))
    if all(simpleq(v) for v in x) and all(simpleq(v) for v in y):  # :689-700
        dest.quadratic.extend(mul(xi, yi) for xi, yi in zip(x, y))
    else:  # :702-709
        for xi, yi in zip(x, y):
            muladd_(dest, xi, yi)
    return dest


def dot(x: Sequence, y: Sequence):
    """LinearAlgebra.dot overloads functions.jl:931-956"""
    xnum = all(_isnum(v) for v in x)
    ynum = all(_isnum(v) for v in y)
    if xnum and ynum:
        return sum(a * b for a, b in zip(x, y))
    if xnum or ynum:
        other = y if xnum else x
        if all(isinstance(v, QuadraticTerm) for v in other):
            return vecdot_(QuadraticFunction(), x, y)
        return vecdot_(AffineFunction(), x, y)
    return vecdot_(QuadraticFunction(), x, y)


def _vec_addsub_(dest: list, x: Sequence, y: Sequence, scalarfun_):
    """vecadd!/vecsubtract! functions.jl:751-764"""
    n = len(x)
    if n != len(y):
        raise DimensionMismatch()
    if len(dest) != n:
        del dest[n:]
        while len(dest) < n:
            dest.append(AffineFunction())
    for i in range(n):
        zero_(dest[i])
        scalarfun_(dest[i], x[i], y[i])
    return dest


def vecadd_(dest, x, y):
    return _vec_addsub_(dest, x, y, add_)


def vecsubtract_(dest, x, y):
    return _vec_addsub_(dest, x, y, subtract_)


def matvecmul_(y: list, A, x: Sequence):
    """matvecmul! functions.jl:775-829.  `A` is any object with .shape and column-major
    linear indexing via A[row, col] (numpy arrays are fine)."""
    if isinstance(y, (AffineFunction, QuadraticFunction)):  # :824-829  (A is an adjoint/transposed vector)
        return vecdot_(y, list(A), x)
    rows, cols = A.shape
    if len(y) != rows or len(x) != cols:
        raise DimensionMismatch()
    for row in range(rows):
        if y[row] is None:
            y[row] = AffineFunction()
        else:
            zero_(y[row])
    if all(isinstance(v, Variable) for v in x):  # :775-798
        for row in range(rows):
            y[row].linear.extend([None] * cols)
        for col in range(cols):
            xc = x[col]
            for row in range(rows):
                y[row].linear[col] = LinearTerm(A[row, col].item() if hasattr(A[row, col], "item") else A[row, col], xc)
    else:  # :800-822
        for col in range(cols):
            for row in range(rows):
                a = A[row, col]
                muladd_(y[row], a.item() if hasattr(a, "item") else a, x[col])
    return y


def matvec(A, x: Sequence) -> list:
    """`A * x`, functions.jl:860-862"""
    return matvecmul_([None] * A.shape[0], A, x)


def bilinearmul_(dest: QuadraticFunction, Q, x: Sequence[Variable], y: Sequence[Variable]):
    """bilinearmul! functions.jl:840-858.  NOTE the reference pairs the k-th (row-outer, col-inner) variable
    pair with Q's k-th COLUMN-MAJOR entry, i.e. with Q[col, row]; reproduced literally."""
    if Q.shape != (len(x), len(y)):
        raise DimensionMismatch()
    zero_(dest)
    nrow = Q.shape[0]
    k = 0
    for row in range(len(x)):
        for col in range(len(y)):
            q = Q[k % nrow, k // nrow]  # column-major linear index k
            dest.quadratic.append(QuadraticTerm(q.item() if hasattr(q, "item") else q, x[row], y[col]))
            k += 1
    return dest


def scale_(dest: list, x, y):
    """scale! functions.jl:873-925"""
    if _isnum(y) and not _isnum(x):
        vec, num, num_first = x, y, False
    else:
        vec, num, num_first = y, x, True
    if len(dest) != len(vec):
        raise DimensionMismatch()
    for i, v in enumerate(vec):
        if isinstance(v, AffineFunction):  # :895-915
            if dest[i] is None:
                dest[i] = AffineFunction()
            mul_(dest[i], num, v) if num_first else mul_(dest[i], v, num)
        else:  # :873-893, 917-925
            dest[i] = mul(num, v) if num_first else mul(v, num)
    return dest


def vcat_(y: list, *args: Sequence[AffineFunction]):
    """vcat! functions.jl:969-994"""
    for f in y:
        zero_(f)
    i = 0
    for source in args:
        if i + len(source) > len(y):
            raise DimensionMismatch()
        for s in source:
            copyto_(y[i], s)
            i += 1
    if i != len(y):
        raise DimensionMismatch()
    return y


def deepcopy_value(v):
    """`deepcopy` of an evaluated expression value (used for `dest` preallocation, lazyexpression.jl:202 etc.)."""
    if isinstance(v, AffineFunction):
        return AffineFunction.of(v)
    if isinstance(v, QuadraticFunction):
        return QuadraticFunction.of(v)
    if isinstance(v, list):
        return [deepcopy_value(e) for e in v]
    if hasattr(v, "copy"):
        return v.copy()
    return v
