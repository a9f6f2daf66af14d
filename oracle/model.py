"""ORACLE (test infrastructure only -- never imported by the product path).

Literal CPU restatement of the control flow of Parametron.jl's hot path:
`Parameter` (src/parameter.jl), `LazyExpression` + the rewrite table
(src/lazyexpression.jl), native->MOI translation and the `Objective` /
`Constraint` / `Constraints` update loop (src/moi_interop.jl), and
`Model` / `initialize!` / `update!` / `solve!` (src/model.jl).

MathOptInterface ~0.8.0 (Project.toml:10,14) is NOT under /root/reference; it
supplies only container structs and the `set` sink.  The containers are
restated here as plain records; the sink is a `RecordingOptimizer` that logs
every `MOI.set` call in order.  The convention that makes the diagonal x2 of
moi_interop.jl:58 correct -- an MOI ScalarQuadraticFunction means
1/2 x'Qx + a'x + c with Q symmetric and duplicate terms summed -- is MOI's
published definition; it is restated in `oracle/dense.py::moi_quadratic_to_dense`.

Parity status: PINNED for the rewrite table by the RNG-free cases of
`test/lazyexpression.jl` and `test/parameter.jl`, and for the MOI boundary
through the closed-form solution checks of `test/model.jl` (restated in
tests/test_oracle_model.py); the raw contents of `moi_f` after `update!` are
not inspected by any reference test.
"""
from __future__ import annotations

import numpy as np

from . import functions as F
from .functions import (AffineFunction, QuadraticFunction, Variable, LinearTerm, QuadraticTerm,
                        deepcopy_value)

Minimize, Maximize = "Minimize", "Maximize"  # Parametron.jl:50


# ---------------------------------------------------------------------------
# Parameter                                          src/parameter.jl
# ---------------------------------------------------------------------------
class Parameter:
    """parameter.jl:36-104.  `inplace=True`: f(val) mutates val; else val = f()."""

    def __init__(self, f, model, val=None, inplace=None):
        if inplace is None:
            inplace = val is not None
        self.dirty = True
        self.f = f
        self.val = val
        self.inplace = inplace
        self.evaluations = 0
        model.addparameter(self)

    @staticmethod
    def manual(model, val):
        """Parameter(model; val=...) = Parameter(identity, val, model), parameter.jl:88"""
        return Parameter(lambda v: v, model, val=val, inplace=True)

    def __call__(self):  # parameter.jl:93-99
        if self.dirty:
            self.update()
            self.dirty = False
        return self.val

    def update(self):  # parameter.jl:101-102
        self.evaluations += 1
        if self.inplace:
            self.f(self.val)
        else:
            self.val = self.f()

    def setdirty(self):  # parameter.jl:104
        self.dirty = True


# ---------------------------------------------------------------------------
# LazyExpression + rewrite table                     src/lazyexpression.jl
# ---------------------------------------------------------------------------
class LazyExpression:
    """lazyexpression.jl:37-61"""

    def __init__(self, f, *args, wrapped=False):
        self.f = f
        self.args = args
        self.wrapped = wrapped

    def __call__(self):
        return self.f(*[evalarg(a) for a in self.args])


def evalarg(x):  # lazyexpression.jl:50-52
    if isinstance(x, (Parameter, LazyExpression)):
        return x()
    return x


def wrap(expr: LazyExpression) -> LazyExpression:
    """lazyexpression.jl:84-87 (type erasure; value semantics unchanged)."""
    if expr.wrapped:
        return expr
    expr()  # `T = typeof(expr())` evaluates once
    return LazyExpression(expr, wrapped=True)


# ---- generic (allocating) functions the macro can meet -------------------------------------
def _isvec(v):
    return isinstance(v, (list, tuple)) or (isinstance(v, np.ndarray) and v.ndim == 1)


def _ismat(v):
    return isinstance(v, np.ndarray) and v.ndim == 2


def _isnumvec(v):
    return isinstance(v, np.ndarray) and v.ndim == 1


def _isvarvec(v):
    return isinstance(v, (list, tuple)) and len(v) > 0 and all(isinstance(e, Variable) for e in v)


def _isfunvec(v):
    return isinstance(v, (list, tuple)) and len(v) > 0 and all(isinstance(e, F.SCALAR_FUNCTIONS) for e in v)


class AdjointVec:
    """transpose(v)/adjoint(v) of a vector (numbers, Variables or functions): Julia's Adjoint/Transpose
    wrapper, which is an AbstractMatrix for dispatch purposes (lazyexpression.jl:200, functions.jl:824-829)."""

    def __init__(self, parent):
        self.parent = [e.item() if hasattr(e, "item") else e for e in parent]


TransposedVarVec = AdjointVec


def g_mul(*args):
    """Julia `*` for the operand kinds the reference's tests use."""
    if len(args) == 3:  # transpose(x) * Q * y, lazyexpression.jl:219-226 / generic left fold
        x, Q, y = args
        if isinstance(x, AdjointVec) and _isvarvec(x.parent) and _ismat(Q) and _isvarvec(y):
            return F.bilinearmul_(QuadraticFunction(), Q, x.parent, list(y))
        return g_mul(g_mul(x, Q), y)
    x, y = args
    if _ismat(x) and _isnumvec(y):
        return x @ y
    if _ismat(x) and _isvec(y):
        return F.matvec(x, list(y))  # functions.jl:860-862 (order of the in-place kernel)
    if isinstance(x, AdjointVec):  # u' * v = dot(u, v)
        return F.dot(x.parent, _aslist(y))
    if F._isnum(x) and _isvec(y) and not _isnumvec(y):
        return F.scale_([None] * len(y), x, list(y))
    if _isvec(x) and not _isnumvec(x) and F._isnum(y):
        return F.scale_([None] * len(x), list(x), y)
    if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
        return x * y
    return F.mul(x, y)


def _elementwise(op, x, y):
    if _isvec(x) and _isvec(y) and not (_isnumvec(x) and _isnumvec(y)):
        if len(x) != len(y):
            raise F.DimensionMismatch()
        return [op(a.item() if hasattr(a, "item") else a, b.item() if hasattr(b, "item") else b) for a, b in zip(x, y)]
    return op(x, y)


def g_add(*args):
    acc = args[0]
    for a in args[1:]:
        acc = _elementwise(F.add, acc, a) if not (isinstance(acc, np.ndarray) and isinstance(a, np.ndarray)) else acc + a
    return acc


def g_sub(x, y):
    if isinstance(x, np.ndarray) and isinstance(y, np.ndarray):
        return x - y
    return _elementwise(F.sub, x, y)


def g_dot(x, y):
    if isinstance(x, np.ndarray) and isinstance(y, np.ndarray):
        return float(np.dot(x.ravel(order="F"), y.ravel(order="F")))
    fx = list(x.ravel(order="F")) if isinstance(x, np.ndarray) else list(x)
    fy = list(y.ravel(order="F")) if isinstance(y, np.ndarray) else list(y)
    return F.dot([v.item() if hasattr(v, "item") else v for v in fx], [v.item() if hasattr(v, "item") else v for v in fy])


def g_adjoint(x):
    if _ismat(x):
        return x.T.copy()
    if _isvec(x):
        return AdjointVec(x)
    return x  # scalars and scalar functions, functions.jl:635,645


g_transpose = g_adjoint


def g_vcat(*args):
    out = []
    for a in args:
        out.extend(list(a) if _isvec(a) else [a])
    return out


def g_vect(*args):
    return list(args)


def g_convert(T, x):
    if T is list:
        return list(x)
    if T is AffineFunction:
        return x if isinstance(x, AffineFunction) else AffineFunction.of(x)
    if T is QuadraticFunction:
        return x if isinstance(x, QuadraticFunction) else QuadraticFunction.of(x)
    if T == "Vector{AffineFunction}":
        return [e if isinstance(e, AffineFunction) else AffineFunction.of(e) for e in x]
    raise TypeError(T)


def g_getproperty(obj, name):
    return getattr(obj, name)


# ---- the optimizer -------------------------------------------------------------------------
def expression(f, *args):
    """One step of the `@expression` macro: `optimize_toplevel(LazyExpression(f, args...))`,
    lazyexpression.jl:161-163,184-193."""
    return optimize_toplevel(LazyExpression(f, *args))


def optimize_toplevel(expr: LazyExpression):
    if expr.wrapped:
        raise ValueError("Cannot optimize wrapped expressions")
    if any(isinstance(a, (Parameter, LazyExpression)) for a in expr.args):
        vals = [evalarg(a) for a in expr.args]  # `argtypes`, :187
        return wrap(optimize(expr, vals))
    return expr()  # constant folding, :191


def _isaffvec(v):
    return isinstance(v, (list, tuple)) and all(isinstance(e, AffineFunction) for e in v)




def optimize(expr: LazyExpression, vals):
    f, args = expr.f, expr.args
    n = len(args)
    if f is g_mul and n == 2 and _ismat(vals[0]) and (_isvarvec(vals[1]) or _isaffvec(vals[1])):  # :200-204
        dest = deepcopy_value(expr())
        return LazyExpression(F.matvecmul_, dest, args[0], args[1])
    if f is g_mul and n == 2 and isinstance(vals[0], AdjointVec) and (_isvarvec(vals[1]) or _isaffvec(vals[1])):
        # same rule :200-204 (an Adjoint vector is an AbstractMatrix) -> matvecmul! method functions.jl:824-829
        dest = deepcopy_value(expr())
        return LazyExpression(lambda dest, A, x: F.vecdot_(dest, A.parent, _aslist(x)), dest, args[0], args[1])
    if f in (g_adjoint,) and n == 1 and _ismat(vals[0]):  # :206-217
        dest = np.empty_like(expr())

        def adjoint_into(dest, A):
            if A.shape != dest.shape[::-1]:
                raise F.DimensionMismatch()
            for j in range(A.shape[1]):
                for i in range(A.shape[0]):
                    dest[j, i] = A[i, j]
            return dest

        return LazyExpression(adjoint_into, dest, args[0])
    if f is g_mul and n == 3 and isinstance(vals[0], AdjointVec) and _isvarvec(vals[0].parent) and _ismat(vals[1]) and _isvarvec(vals[2]):  # :219-226
        x, Q, y = args
        return LazyExpression(lambda dest, Q, x, y: F.bilinearmul_(dest, Q, x.parent, y), QuadraticFunction(), Q, x, y)
    if f is g_dot and n == 2 and _isvec(vals[0]) and _isvec(vals[1]):  # :228-232
        ret = expr()
        if isinstance(ret, (AffineFunction, QuadraticFunction)):
            dest = deepcopy_value(ret)
            return LazyExpression(lambda dest, x, y: F.vecdot_(dest, _aslist(x), _aslist(y)), dest, args[0], args[1])
        return expr  # numeric dot: generic vecdot! fallback functions.jl:711-714
    if f is g_add and n >= 3:  # :234-236
        inner = optimize_toplevel(LazyExpression(g_add, args[0], args[1]))
        return _unwrap_for_nesting(optimize_toplevel(LazyExpression(g_add, inner, *args[2:])))
    if f in (g_add, g_sub) and n == 2:  # :238-258
        ret = expr()
        scalar_, vec_ = (F.add_, F.vecadd_) if f is g_add else (F.subtract_, F.vecsubtract_)
        if isinstance(ret, (AffineFunction, QuadraticFunction)):
            return LazyExpression(scalar_, type(ret)(), args[0], args[1])
        if _isaffvec(ret) and len(ret) > 0:
            return LazyExpression(lambda dest, x, y: vec_(dest, _aslist(x), _aslist(y)), deepcopy_value(ret), args[0], args[1])
        return expr
    if f is g_mul and n == 2:
        a, b = vals
        scal = (Variable, LinearTerm)
        if (isinstance(a, AffineFunction) and (F._isnum(b) or isinstance(b, scal))) or \
           (isinstance(b, AffineFunction) and (F._isnum(a) or isinstance(a, scal))) or \
           (isinstance(a, QuadraticFunction) and (F._isnum(b) or isinstance(b, Variable))) or \
           (isinstance(b, QuadraticFunction) and (F._isnum(a) or isinstance(a, Variable))):  # :260-274
            return LazyExpression(F.mul_, deepcopy_value(expr()), args[0], args[1])
        numvarvec = lambda v: _isvec(v)  # Vector{<:Union{Number, Variable, AffineFunction}}
        if (F._isnum(a) and numvarvec(b)) or (numvarvec(a) and F._isnum(b)):  # :284-290
            ret = expr()
            if isinstance(ret, np.ndarray):
                def numscale(dest, x, y):
                    np.multiply(x, y, out=dest)
                    return dest
                return LazyExpression(numscale, ret.copy(), args[0], args[1])
            return LazyExpression(lambda dest, x, y: F.scale_(dest, _aslist(x) if _isvec(x) else x, _aslist(y) if _isvec(y) else y),
                                  deepcopy_value(ret), args[0], args[1])
    if f is g_vcat and all(_isaffvec(v) for v in vals):  # :276-278
        return LazyExpression(F.vcat_, deepcopy_value(expr()), *args)
    if f is g_convert and n == 2 and _isvec(vals[1]):  # :280-282
        dest = deepcopy_value(expr())

        def copyto(dest, src):
            dest[:] = src  # copyto!(dest, src): element references
            return dest

        return LazyExpression(copyto, dest, args[1])
    if f is g_vect and n == 1 and (F._isnum(vals[0]) or isinstance(vals[0], (Variable, LinearTerm, AffineFunction))):  # :292-298
        dest = deepcopy_value(expr())

        def vect_into(dest, x):
            dest[0] = x
            return dest

        return LazyExpression(vect_into, dest, args[0])
    if f is g_getproperty and n == 2 and isinstance(args[1], str):  # :300-302
        name = args[1]
        return LazyExpression(lambda obj: getattr(obj, name), args[0])
    return expr  # :198


def _unwrap_for_nesting(e):
    # optimize() must return an un-wrapped expression for wrap(); a wrapped one is returned as is by wrap().
    return e


def _aslist(v):
    if isinstance(v, np.ndarray):
        return [e.item() for e in v]
    return list(v)


# ---------------------------------------------------------------------------
# MOI containers + native->MOI translation           src/moi_interop.jl:35-92
# ---------------------------------------------------------------------------
class ScalarAffineFunctionMOI:
    def __init__(self):
        self.terms = []  # (coeff, variable_index)
        self.constant = 0.0


class ScalarQuadraticFunctionMOI:
    def __init__(self):
        self.affine_terms = []  # (coeff, variable_index)
        self.quadratic_terms = []  # (coeff, variable_index_1, variable_index_2)
        self.constant = 0.0


class VectorAffineFunctionMOI:
    def __init__(self):
        self.terms = []  # (output_index, coeff, variable_index)
        self.constants = []


class IdentityVarMap:  # moi_interop.jl:32-33
    def __getitem__(self, i):
        return i


def moi_update(moi_f, f, varmap=IdentityVarMap()):
    if isinstance(moi_f, ScalarAffineFunctionMOI):  # :35-43
        moi_f.constant = f.constant
        moi_f.terms = [(t.coeff, varmap[t.var.index]) for t in f.linear]
    elif isinstance(moi_f, ScalarQuadraticFunctionMOI):  # :45-62
        moi_f.constant = f.affine.constant
        moi_f.affine_terms = [(t.coeff, varmap[t.var.index]) for t in f.affine.linear]
        moi_f.quadratic_terms = [
            (2 * t.coeff if t.rowvar == t.colvar else t.coeff, varmap[t.rowvar.index], varmap[t.colvar.index])
            for t in f.quadratic]
    elif isinstance(moi_f, VectorAffineFunctionMOI):  # :64-81
        terms = []
        consts = []
        for row, fr in enumerate(f, start=1):
            for t in fr.linear:
                terms.append((row, t.coeff, varmap[t.var.index]))
            consts.append(fr.constant)
        moi_f.terms = terms
        moi_f.constants = consts
    return moi_f


def canonical_function_type(val):  # moi_interop.jl:96-101
    if isinstance(val, (Variable, LinearTerm, AffineFunction)):
        return AffineFunction
    if isinstance(val, (QuadraticTerm, QuadraticFunction)):
        return QuadraticFunction
    if isinstance(val, (list, tuple)) and all(isinstance(e, (AffineFunction, Variable, LinearTerm)) for e in val):
        return "Vector{AffineFunction}"
    if F._isnum(val):
        return AffineFunction
    raise TypeError(f"no canonical function type for {type(val)}")


def make_moi_equivalent(E):  # moi_interop.jl:87-92
    if E is AffineFunction:
        return ScalarAffineFunctionMOI()
    if E is QuadraticFunction:
        return ScalarQuadraticFunctionMOI()
    return VectorAffineFunctionMOI()


class _FunctionHolder:
    """Common part of Objective (moi_interop.jl:113-129) and Constraint (:141-159)."""

    def __init__(self, expr):
        val = evalarg(expr)
        self.E = canonical_function_type(val)
        converted = expression(g_convert, self.E, expr)
        self.isconstant = not isinstance(converted, (LazyExpression, Parameter))
        self.expr = converted if isinstance(converted, LazyExpression) else LazyExpression(lambda v: v, converted, wrapped=True)
        self.f = make_moi_equivalent(self.E)
        moi_update(self.f, self.expr())


class Objective(_FunctionHolder):
    def update(self, optimizer, varmap):  # moi_interop.jl:131-137
        if not self.isconstant:
            moi_update(self.f, self.expr(), varmap)
            optimizer.set_objective_function(self.f)


class Constraint(_FunctionHolder):
    def __init__(self, expr, set_):
        super().__init__(expr)
        self.set = set_
        self.modelindex = None
        self.optimizerindex = None

    def update(self, optimizer, varmap):  # moi_interop.jl:168-175
        if not self.isconstant:
            moi_update(self.f, self.expr(), varmap)
            optimizer.set_constraint_function(self.optimizerindex, self.f)


# bucket order, moi_interop.jl:180-193 (quadratic-vector and single-variable buckets carry no updatable data)
BUCKETS = [("ScalarAffine", "GreaterThan"), ("ScalarAffine", "LessThan"), ("ScalarAffine", "EqualTo"),
           ("VectorAffine", "Nonnegatives"), ("VectorAffine", "Nonpositives"), ("VectorAffine", "Zeros"),
           ("ScalarQuadratic", "GreaterThan"), ("ScalarQuadratic", "LessThan"), ("ScalarQuadratic", "EqualTo")]


def _bucket_of(c: Constraint):
    kind = {ScalarAffineFunctionMOI: "ScalarAffine", ScalarQuadraticFunctionMOI: "ScalarQuadratic",
            VectorAffineFunctionMOI: "VectorAffine"}[type(c.f)]
    return BUCKETS.index((kind, c.set))


class RecordingOptimizer:
    """Stands in for the external MOI optimizer: logs the `MOI.set` calls of one update! in order."""

    def __init__(self):
        self.log = []

    def set_objective_function(self, f):
        self.log.append(("objective", f))

    def set_constraint_function(self, ci, f):
        self.log.append(("constraint", ci, f))


# ---------------------------------------------------------------------------
# Model                                               src/model.jl
# ---------------------------------------------------------------------------
class Model:
    def __init__(self, optimizer=None):  # model.jl:10-21
        self.params = []
        self.optimizer = optimizer if optimizer is not None else RecordingOptimizer()
        self.initialized = False
        self.nvars = 0
        self.sense = Minimize
        self.objective = Objective(AffineFunction())
        self.constraints = [[] for _ in BUCKETS]
        self.model_var_to_optimizer = []
        self._nconstraints = 0

    def addparameter(self, p):  # model.jl:42
        self.params.append(p)
        return p

    def setdirty(self):  # model.jl:40
        for p in self.params:
            p.setdirty()

    def variable(self) -> Variable:  # model.jl:49-53
        if self.initialized:
            raise RuntimeError("Model has already been initialized.")
        self.nvars += 1
        return Variable(self.nvars)

    def setobjective(self, sense, expr):  # model.jl:60-66
        if self.initialized:
            raise RuntimeError("Model was already initialized. setobjective! can only be called before initialization.")
        self.objective = Objective(expr)
        self.sense = sense

    def add_constraint(self, c: Constraint):  # model.jl:68-73
        if self.initialized:
            raise RuntimeError("Model was already initialized. add_constraint can only be called before initialization.")
        self._nconstraints += 1
        c.modelindex = self._nconstraints
        self.constraints[_bucket_of(c)].append(c)
        return c

    def _add(self, expr, scalar_set, vector_set):  # model.jl:75-98
        val = evalarg(expr)
        if isinstance(val, (list, tuple)):
            return self.add_constraint(Constraint(expr, vector_set))
        return self.add_constraint(Constraint(expr, scalar_set))

    def add_nonnegative_constraint(self, expr):
        return self._add(expr, "GreaterThan", "Nonnegatives")

    def add_nonpositive_constraint(self, expr):
        return self._add(expr, "LessThan", "Nonpositives")

    def add_zero_constraint(self, expr):
        return self._add(expr, "EqualTo", "Zeros")

    def constraint(self, lhs, relation: str, rhs):
        """`@constraint(model, lhs relation rhs)`, model.jl:224-249."""
        e = expression(g_sub, lhs, rhs)
        if relation == ">=":
            return self.add_nonnegative_constraint(e)
        if relation == "<=":
            return self.add_nonpositive_constraint(e)
        if relation == "==":
            return self.add_zero_constraint(e)
        raise ValueError("Relation not recognized")

    def initialize(self, idxmap=None):  # model.jl:100-122
        """`idxmap` plays MOI.copy_to's index map: model variable index -> optimizer variable index."""
        for bucket in self.constraints:
            for c in bucket:
                c.optimizerindex = c.modelindex
        self.model_var_to_optimizer = [None] + [(idxmap[i] if idxmap is not None else i) for i in range(1, self.nvars + 1)]
        self.initialized = True

    def update(self):  # model.jl:132-143
        self.setdirty()
        self.objective.update(self.optimizer, self.model_var_to_optimizer)
        for bucket in self.constraints:  # moi_interop.jl:236-247
            for c in bucket:
                c.update(self.optimizer, self.model_var_to_optimizer)

    def solve(self):  # model.jl:151-159 (the external MOI.optimize! is out of scope)
        if not self.initialized:
            self.initialize()
        self.update()
