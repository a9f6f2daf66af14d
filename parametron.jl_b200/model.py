"""Host-side mirror of Parametron.jl's user API for the `update!` hot path, above the C ABI (libpmb200.so).

Julia is not installed here (nor on the GPU box), so this Python module plays the role of the Julia shim of
INTEGRATION.md: same names, argument meaning and error behaviour as the reference --

    Model(optimizer)                      src/model.jl:1-30        -> Model(device=0, batch=1)
    Variable(model)                       src/model.jl:49-53       -> Variable(model) / model.variables(n)
    Parameter(f, val, model) etc.         src/parameter.jl:48-88   -> Parameter(...)
    @expression  A * x - b                src/lazyexpression.jl:140-302 -> operator overloads on `Expr` handles
    @objective(model, Minimize, e)        src/model.jl:267-271     -> objective(model, Minimize, e)
    @constraint(model, lhs == rhs)        src/model.jl:224-249     -> constraint(model, lhs, "==", rhs)
    initialize! / update! / solve!        src/model.jl:117-159     -> initialize / update / solve

All arithmetic runs in the CUDA kernels behind `pmb_update`; this module only builds the DAG and moves bytes.
There is no CPU path: constructing a Model without the compiled library or without a B200 raises.
"""
import ctypes as C

import numpy as np

from . import _capi as K

Minimize, Maximize = K.MINIMIZE, K.MAXIMIZE  # Parametron.jl:50


class Variable:
    """src/functions.jl:96-98 -- an index into the model's variables (1-based).  Arithmetic on a Variable that knows its model builds
    expression nodes (`x + 2 * y >= 4`, test/model.jl:289): Variable -> AffineFunction with one unit term."""
    __slots__ = ("index", "model")

    def __init__(self, model_or_index, model=None):
        if isinstance(model_or_index, Model):
            self.model = model_or_index
            self.index = model_or_index._add_variables(1)
        else:
            self.index = int(model_or_index)
            self.model = model

    def _expr(self):
        if self.model is None:
            raise TypeError("this Variable is not attached to a Model (use Variable(model) or model.variables(n))")
        return self.model.expr(self)

    def __add__(self, o):
        return self._expr() + o

    def __radd__(self, o):
        return o + self._expr()

    def __sub__(self, o):
        return self._expr() - o

    def __rsub__(self, o):
        return o - self._expr()

    def __mul__(self, o):
        return self._expr() * o

    def __rmul__(self, o):
        return o * self._expr()

    def __neg__(self):
        return -self._expr()

    def __pow__(self, o):
        raise K.Unsupported(K.EUNSUPPORTED, "`^` has no device lowering: write x * x through a bilinear form (reference: lazyexpression.jl has no pow rule either)")

    def __repr__(self):
        return f"x{self.index}"

    def __eq__(self, o):
        return isinstance(o, Variable) and o.index == self.index

    def __hash__(self):
        return hash(self.index)


class Expr:
    """Handle of a DAG node (what `@expression` returns, lazyexpression.jl:65 `WrappedExpression`)."""
    __array_priority__ = 1000
    __array_ufunc__ = None  # numpy defers to our __r*__ operators

    def __init__(self, model, node):
        self.model = model
        self.node = int(node)

    # -- introspection
    def info(self):
        kind, rows, cols, const = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int32()
        self.model._ck(self.model._lib.pmb_node_info(self.model._ctx, self.node, kind, rows, cols, const))
        return kind.value, rows.value, cols.value, bool(const.value)

    @property
    def kind(self):
        return self.info()[0]

    # -- operators (each is one rewrite rule of lazyexpression.jl:198-302)
    def _bin(self, fn, a, b):
        out = C.c_int32()
        m = self.model
        m._ck(fn(m._ctx, m._as_node(a), m._as_node(b), out))
        return Expr(m, out.value)

    def __mul__(self, o):
        return self._bin(self.model._lib.pmb_op_mul, self, o)

    def __rmul__(self, o):
        return self._bin(self.model._lib.pmb_op_mul, o, self)

    __matmul__ = __mul__
    __rmatmul__ = __rmul__

    def __add__(self, o):
        return self._bin(self.model._lib.pmb_op_add, self, o)

    def __radd__(self, o):
        return self._bin(self.model._lib.pmb_op_add, o, self)

    def __sub__(self, o):
        return self._bin(self.model._lib.pmb_op_sub, self, o)

    def __rsub__(self, o):
        return self._bin(self.model._lib.pmb_op_sub, o, self)

    def __neg__(self):
        return (-1.0) * self

    @property
    def T(self):
        return adjoint(self)


class Parameter(Expr):
    """src/parameter.jl:36-104.

    Parameter(model, val=buf)            manual mode (:88): the caller mutates `buf`/the device buffer between solves
    Parameter(model, f=f, val=buf)       in-place (:57): `f(buf)` is called exactly once per update!
    Parameter(model, f=f)                out-of-place (:48,73): `val = f()` once per update! (shape taken from the first call)
    Parameter(model, shape=..., uniform=(seed, scale))   built-in device generator scale*U[0,1)  (README.md:36-43 `rand!`)
    Parameter(model, shape=..., device=True)             device-resident manual mode; write through `.upload()` / `.device_ptr`
    Matrices are column-major (Julia layout); with batch > 1 every value has a leading instance axis.
    """

    def __init__(self, model, f=None, val=None, shape=None, uniform=None, device=False, device_ptr=None, bsr=None):
        """bsr=(rows, cols, b, block_rows, block_cols): block-sparse matrix; the value is an (nblocks, b, b) array of blocks."""
        self.bsr = None
        if bsr is not None:
            rows_, cols_, b_, br, bc = bsr
            br, bc = np.ascontiguousarray(br, dtype=np.int64), np.ascontiguousarray(bc, dtype=np.int64)
            self.bsr = (int(rows_), int(cols_), int(b_), br, bc)
            shape = (len(br), int(b_), int(b_))
        self.f = f
        self.inplace = val is not None
        self.evaluations = 0
        self.host = None
        if uniform is not None:
            mode = K.PARAM_UNIFORM
        elif device or device_ptr is not None:
            mode = K.PARAM_DEVICE
        else:
            mode = K.PARAM_HOST
            if val is None:
                if f is None:
                    raise ValueError("Parameter needs f, val, uniform= or device=")
                val = f()  # out-of-place constructor evaluates once to learn the type (parameter.jl:48-50)
                self.evaluations += 1
            val = np.asarray(val, dtype=np.float64)
            shape = val.shape[1:] if model.batch > 1 else val.shape
        shape = tuple(int(s) for s in (shape if shape is not None else ()))
        if len(shape) > 2 and self.bsr is None:
            raise ValueError("Parameters are scalars, vectors or matrices")
        self.shape = shape
        rows = shape[0] if len(shape) >= 1 else 0
        cols = shape[1] if len(shape) == 2 else 0
        out = C.c_int32()
        if self.bsr is not None:
            r_, c_, b_, br, bc = self.bsr
            model._ck(model._lib.pmb_param_create_bsr(model._ctx, r_, c_, b_, len(br), C.c_void_p(br.ctypes.data), C.c_void_p(bc.ctypes.data),
                                                      mode, out))
        else:
            model._ck(model._lib.pmb_param_create(model._ctx, len(shape), rows, cols, mode, out))
        super().__init__(model, out.value)
        self.mode = mode
        self.numel = int(np.prod(shape)) if shape else 1
        if mode == K.PARAM_UNIFORM:
            seed, scale = uniform
            model._ck(model._lib.pmb_param_set_uniform(model._ctx, self.node, int(seed), float(scale)))
        elif mode == K.PARAM_DEVICE:
            if device_ptr is not None:
                model._ck(model._lib.pmb_param_set_device_ptr(model._ctx, self.node, C.c_void_p(int(device_ptr))))
        else:
            # pinned staging buffer in the library's layout: [instance][column-major value]
            self.host = model._pinned((model.batch, self.numel))
            self._store(val)
            model._ck(model._lib.pmb_param_set_host_ptr(model._ctx, self.node, C.c_void_p(self.host.ctypes.data), None, None))
            self.val = val
        model.params.append(self)

    def _to_device_layout(self, val):
        """[instance][column-major value] (BSR: [instance][block][column-major block])"""
        val = np.asarray(val, dtype=np.float64)
        b = self.model.batch
        full = ((b,) if b > 1 else ()) + self.shape
        if val.shape != full:
            raise K.DimensionMismatch(K.EDIM, f"Parameter value has shape {val.shape}, expected {full}")
        v = val.reshape((b,) + self.shape)
        if len(self.shape) == 2:
            v = np.transpose(v, (0, 2, 1))
        elif len(self.shape) == 3:
            v = np.transpose(v, (0, 1, 3, 2))
        return v.reshape(b, self.numel)

    def _store(self, val):
        self.host[...] = self._to_device_layout(val)

    def _evaluate(self):
        """parameter.jl:93-102 -- called once per update! by Model.update (the dirty flag is the update itself)."""
        if self.mode != K.PARAM_HOST:
            return
        if self.f is not None:
            if self.inplace:
                self.f(self.val)
            else:
                self.val = self.f()
            self.evaluations += 1
        self._store(self.val)

    def _needed(self):
        """reachable from the objective or a constraint?  (pmb_param_device_ptr is null for dead Parameters: the reference never evaluates them either)"""
        if getattr(self, "_needed_cache", None) is None:
            self._needed_cache = (not self.model.initialized) or bool(self.device_ptr)
        return self._needed_cache

    # device-resident manual mode
    def upload(self, val):
        m = self.model
        b = m.batch
        stage = m._pinned((b, self.numel))
        stage[...] = self._to_device_layout(val)
        m._ck(m._lib.pmb_param_upload(m._ctx, self.node, C.c_void_p(stage.ctypes.data)))
        m.synchronize()

    @property
    def device_ptr(self):
        p = C.c_void_p()
        self.model._ck(self.model._lib.pmb_param_device_ptr(self.model._ctx, self.node, p))
        return p.value

    def download(self):
        m = self.model
        flat = m._d2h(self.device_ptr, m.batch * self.numel)
        if len(self.shape) == 3:
            out = np.transpose(flat.reshape((m.batch, self.shape[0], self.shape[2], self.shape[1])), (0, 1, 3, 2))
        else:
            out = flat.reshape((m.batch,) + self.shape[::-1])
            if len(self.shape) == 2:
                out = np.transpose(out, (0, 2, 1))
        return out if m.batch > 1 else out[0]


class RecordParameter(Parameter):
    """A Parameter whose value is a struct (`Parameter(model) do MyWrapper(rand()) end`, test/lazyexpression.jl:355-382).  On the
    device the fields are packed into one flat vector in declaration order (matrices column-major); `p.field` is the GetField rewrite
    (lazyexpression.jl:300-302, functions.jl:998-999).  `fields` = ordered {name: shape}; `f()` returns an object (or dict) carrying
    those fields and is called exactly once per update! like any other Parameter."""

    def __init__(self, model, f, fields):
        self._fields, off = {}, 0
        for name, shp in fields.items():
            shp = tuple(int(v) for v in shp)
            self._fields[name] = (off, shp)
            off += int(np.prod(shp)) if shp else 1
        self._record_f, self.record = f, None

        def flat():
            rec = self._record_f()
            self.record = rec
            parts = []
            for name, (_, shp) in self._fields.items():
                a = np.asarray(rec[name] if isinstance(rec, dict) else getattr(rec, name), dtype=np.float64)
                if a.shape != shp:
                    raise K.DimensionMismatch(K.EDIM, f"field {name} has shape {a.shape}, declared {shp}")
                parts.append(a.reshape(-1, order="F"))
            return np.concatenate(parts) if parts else np.zeros(0)
        super().__init__(model, f=flat)

    def field(self, name):
        off, shp = self._fields[name]
        rows = shp[0] if len(shp) >= 1 else 0
        cols = shp[1] if len(shp) == 2 else 0
        out = C.c_int32()
        m = self.model
        m._ck(m._lib.pmb_op_getfield(m._ctx, self.node, off, len(shp), rows, cols, out))
        return Expr(m, out.value)

    def __getattr__(self, name):
        if name.startswith("_") or "_fields" not in self.__dict__ or name not in self.__dict__["_fields"]:
            raise AttributeError(name)
        return self.field(name)


def _flatten_vars(x):
    if isinstance(x, np.ndarray) and x.dtype == object:
        x = list(x.ravel())
    if isinstance(x, (list, tuple)) and len(x) > 0 and all(isinstance(v, Variable) for v in x):
        return [v.index for v in x]
    return None


class Model:
    """src/model.jl:1-30.  `optimizer` is an optional callable invoked by solve() with the model after update() -- the
    external solve itself is out of scope and untouched (north_star); results stay on the device."""

    def __init__(self, optimizer=None, device=0, batch=1, stream=None, gram_mode=None, gram_tol=0.0, instance_offset=0):
        self._lib = K.load()
        self._ctx = C.c_void_p()
        code = self._lib.pmb_ctx_create(int(device), C.c_void_p(stream) if stream else None, int(batch), C.byref(self._ctx))
        if code != K.OK:
            raise K.PmbError(code, f"pmb_ctx_create failed with {code}: a B200 (sm_100) and the CUDA runtime are required; "
                                   "there is no CPU fallback")
        if gram_mode is not None:  # 0: fp64 DMMA only, 1: certified int8 digit planes on tcgen05 (default), 5 / 6: forced planes
            self._ck(self._lib.pmb_set_gram_mode(self._ctx, int(gram_mode), float(gram_tol)))
        # instance-shard: this model owns global instances [instance_offset, instance_offset + batch) of a larger batch
        self.instance_offset = int(instance_offset)
        if self.instance_offset:
            self._ck(self._lib.pmb_set_instance_offset(self._ctx, self.instance_offset))
        self.optimizer = optimizer
        self.device = int(device)
        self.batch = int(batch)
        self.params = []
        self.initialized = False
        self.nvars = 0
        self._pins = []
        self._varvec_cache = {}
        self._has_host_params = None

    # -- plumbing
    def _ck(self, code):
        K.check(self._lib, self._ctx, code)

    def gram_level_counts(self):
        """instances of the last update whose Hessian contraction ran with (5 int8 planes, 6 int8 planes, fp64)"""
        out = (C.c_int64 * 3)()
        self._ck(self._lib.pmb_gram_level_counts(self._ctx, out))
        return tuple(int(v) for v in out)

    def close(self):
        if self._ctx:
            self._lib.pmb_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _pinned(self, shape):
        n = int(np.prod(shape))
        p = C.c_void_p()
        self._ck(self._lib.pmb_host_alloc(self._ctx, max(n, 1) * 8, C.byref(p)))
        buf = (C.c_double * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
        self._pins.append(arr)
        return arr

    def _d2h(self, dptr, count, dtype=np.float64):
        out = np.empty(count, dtype=dtype)
        if count:
            self._ck(self._lib.pmb_memcpy_d2h(self._ctx, C.c_void_p(out.ctypes.data), C.c_void_p(dptr), out.nbytes))
            self.synchronize()
        return out

    def _dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._ck(self._lib.pmb_device_alloc(self._ctx, int(nbytes), C.byref(p)))
        return p.value

    def _dev_free(self, ptr):
        if ptr:
            self._ck(self._lib.pmb_device_free(self._ctx, C.c_void_p(ptr)))

    def synchronize(self):
        self._ck(self._lib.pmb_synchronize(self._ctx))

    def _add_variables(self, n):
        first = C.c_int64()
        self._ck(self._lib.pmb_add_variables(self._ctx, n, first))
        self.nvars += n
        return first.value

    def variables(self, n):
        """[Variable(model) for _ = 1 : n]"""
        first = self._add_variables(n)
        return [Variable(first + i, model=self) for i in range(n)]

    def _as_node(self, x):
        """numbers / arrays -> constants (constant folding, lazyexpression.jl:189-192); Variable lists -> Vector{Variable}."""
        if isinstance(x, Expr):
            if x.model is not self:
                raise ValueError("expression belongs to another model")
            return x.node
        out = C.c_int32()
        if isinstance(x, Variable):
            c = np.ones(1)
            v = np.array([x.index], dtype=np.int64)
            self._ck(self._lib.pmb_const_affine(self._ctx, 1, C.c_void_p(c.ctypes.data), C.c_void_p(v.ctypes.data), 0.0, out))
            return out.value
        idx = _flatten_vars(x)
        if idx is not None:
            key = tuple(idx)
            if key not in self._varvec_cache:
                v = np.asarray(idx, dtype=np.int64)
                self._ck(self._lib.pmb_varvec(self._ctx, len(idx), C.c_void_p(v.ctypes.data), out))
                self._varvec_cache[key] = out.value
            return self._varvec_cache[key]
        a = np.asarray(x, dtype=np.float64)
        if a.ndim == 0:
            rows = cols = 0
        elif a.ndim == 1:
            rows, cols = a.shape[0], 0
        elif a.ndim == 2:
            rows, cols = a.shape
        else:
            raise ValueError("constants are scalars, vectors or matrices")
        flat = np.ascontiguousarray(a.T).ravel() if a.ndim == 2 else np.ascontiguousarray(a).ravel()
        self._ck(self._lib.pmb_const_numeric(self._ctx, a.ndim, rows, cols, C.c_void_p(flat.ctypes.data), out))
        return out.value

    def expr(self, x):
        return x if isinstance(x, Expr) else Expr(self, self._as_node(x))

    # -- model.jl:60-98
    def setobjective(self, sense, e):
        self._ck(self._lib.pmb_set_objective(self._ctx, int(sense), self._as_node(e)))

    def _add(self, e, set_kind):
        out = C.c_int32()
        self._ck(self._lib.pmb_add_constraint(self._ctx, self._as_node(e), set_kind, out))
        return out.value

    def add_nonnegative_constraint(self, e):
        return self._add(e, K.SET_NONNEGATIVE)

    def add_nonpositive_constraint(self, e):
        return self._add(e, K.SET_NONPOSITIVE)

    def add_zero_constraint(self, e):
        return self._add(e, K.SET_ZERO)

    def add_integer_constraint(self, x):
        """model.jl:97: SingleVariable-in-Integer carries no updatable data and belongs to the external solver, which is out of scope"""
        raise K.Unsupported(K.EUNSUPPORTED, "integer constraints are not lowered to the device (no coefficient data; declare them on the solver)")

    def add_binary_constraint(self, x):
        """model.jl:98"""
        raise K.Unsupported(K.EUNSUPPORTED, "binary constraints are not lowered to the device (no coefficient data; declare them on the solver)")

    # -- model.jl:117-159
    def initialize(self, idxmap=None):
        """initialize!: `idxmap[i-1]` = optimizer index of model variable i (MOI.copy_to's index map); None = identity."""
        vm = None
        if idxmap is not None:
            vm = np.asarray(idxmap, dtype=np.int64)
        self._ck(self._lib.pmb_initialize(self._ctx, C.c_void_p(vm.ctypes.data) if vm is not None else None))
        self.initialized = True

    def update(self):
        """update!: every Parameter evaluated exactly once, then ONE device pass re-assembles (Q, q, c0) and all (C_i, d_i)."""
        if self._has_host_params is None:   # only Parameters the objective / a constraint reaches are evaluated (the reference evaluates lazily)
            self._has_host_params = any(p.mode == K.PARAM_HOST for p in self.params)
        if self._has_host_params and self.initialized:
            self.synchronize()               # the previous update's H2D copies have consumed the pinned staging buffers rewritten below
        for p in self.params:
            if p.mode == K.PARAM_HOST and not p._needed():
                continue
            p._evaluate()
        self._ck(self._lib.pmb_update(self._ctx))

    def solve(self):
        if not self.initialized:
            self.initialize()
        self.update()
        if self.optimizer is not None:
            self.optimizer(self)

    # -- result path (model.jl:166-173): the optimizer hands its primal solution back in OPTIMIZER variable order
    def set_primal(self, x, device_ptr=None):
        """Store the solver's VariablePrimal vector ([batch,] N in optimizer index order); device_ptr: it already lives on the device."""
        if device_ptr is not None:
            self._ck(self._lib.pmb_set_primal(self._ctx, C.c_void_p(int(device_ptr)), 1))
            return
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(self.batch, self.nvars))
        self._ck(self._lib.pmb_set_primal(self._ctx, C.c_void_p(x.ctypes.data), 0))

    def value(self, x):
        """value(m, x) / value.(m, xs): VariablePrimal[model_var_to_optimizer[x.index]] (model.jl:166)"""
        scalar = isinstance(x, Variable)
        idx = np.array([x.index] if scalar else _flatten_vars(x), dtype=np.int64)
        out = np.empty((self.batch, len(idx)))
        self._ck(self._lib.pmb_value(self._ctx, len(idx), C.c_void_p(idx.ctypes.data), C.c_void_p(out.ctypes.data), 0))
        out = out[0] if self.batch == 1 else out
        return out[..., 0] if scalar else out

    def objectivevalue(self):
        """objectivevalue(m) (model.jl:173): the objective of the last update evaluated at the stored solution, on the device"""
        out = (C.c_double * self.batch)()
        self._ck(self._lib.pmb_objective_value(self._ctx, out))
        return float(out[0]) if self.batch == 1 else np.array(out[:])

    # -- device-resident handoff (replaces MOI.set, moi_interop.jl:134,171)
    def objective_view(self):
        v = K.ObjectiveView()
        self._ck(self._lib.pmb_get_objective(self._ctx, C.byref(v)))
        return v

    def num_blocks(self):
        n = C.c_int32()
        self._ck(self._lib.pmb_num_blocks(self._ctx, n))
        return n.value

    def block_view(self, pos):
        v = K.BlockView()
        self._ck(self._lib.pmb_get_block(self._ctx, pos, C.byref(v)))
        return v

    def objective_arrays(self):
        """(Q, q, c0) copied to the host: Q is the symmetric MOI Hessian (1/2 x'Qx + q'x + c0); leading batch axis if batch > 1."""
        v = self.objective_view()
        N, b = v.nvars, self.batch
        Q = None
        if v.Q:
            Q = np.transpose(self._d2h(v.Q, b * N * N).reshape(b, N, N), (0, 2, 1))
        q = self._d2h(v.q, b * N).reshape(b, N)
        c0 = self._d2h(v.c0, b)
        if b == 1:
            return (Q[0] if Q is not None else None), q[0], float(c0[0])
        return Q, q, c0

    def block_arrays(self, pos):
        """(C, d) of the constraint block at update position `pos` (C x + d in the set), copied to the host."""
        v = self.block_view(pos)
        m, N, b = v.rows, v.nvars, self.batch
        d = self._d2h(v.d, b * m).reshape(b, m)
        if v.is_sparse:  # densify on the host for inspection: duplicates (a variable used twice) are summed
            vals, rows, cols, _ = self.block_sparse(pos)
            Cm = np.zeros((b, m, N))
            for i in range(b):
                np.add.at(Cm[i], (rows, cols), vals.reshape(b, -1)[i])
        else:
            Cm = np.transpose(self._d2h(v.C, b * m * N).reshape(b, N, m), (0, 2, 1))
        return (Cm[0], d[0]) if b == 1 else (Cm, d)

    def block_quadratic(self, pos):
        """(Q, q, c0) of a scalar quadratic constraint block (1/2 x'Qx + q'x + c0 in the set)."""
        v = self.block_view(pos)
        if not v.Q:
            raise ValueError("block is not a quadratic constraint")
        N, b = v.nvars, self.batch
        Q = np.transpose(self._d2h(v.Q, b * N * N).reshape(b, N, N), (0, 2, 1))
        q = self._d2h(v.C, b * N).reshape(b, N)
        c0 = self._d2h(v.d, b)
        return (Q[0], q[0], float(c0[0])) if b == 1 else (Q, q, c0)

    def block_sparse(self, pos):
        """(vals, rows, cols, d) of a sparse block: coefficient e is vals[e] at 0-based (rows[e], cols[e]); the index is static."""
        v = self.block_view(pos)
        if not v.is_sparse:
            raise ValueError("block is dense")
        b = self.batch
        vals = self._d2h(v.vals, b * v.nnz).reshape(b, v.nnz)
        rows = self._d2h(v.sp_rows, v.nnz, np.int32)
        cols = self._d2h(v.sp_cols, v.nnz, np.int32)
        d = self._d2h(v.d, b * v.rows).reshape(b, v.rows)
        return (vals[0], rows, cols, d[0]) if b == 1 else (vals, rows, cols, d)

    def block_terms(self, pos):
        """(rows, vars, src) of the reference's VectorAffineTerm list for this block (moi_interop.jl:64-81), int64, 1-based."""
        v = self.block_view(pos)
        n = v.nterms
        rows, vars_, src = (np.empty(n, dtype=np.int64) for _ in range(3))
        self._ck(self._lib.pmb_block_term_index(self._ctx, pos, n, C.c_void_p(rows.ctypes.data), C.c_void_p(vars_.ctypes.data),
                                                C.c_void_p(src.ctypes.data)))
        return rows, vars_, src

    # -- "next" rows of the scope table: MOI-compatible term list, solver-side QP ingest
    TERM_DTYPE = np.dtype([("output_index", np.int64), ("coefficient", np.float64), ("variable", np.int64)])

    def export_block_terms(self, pos):
        """The reference's MOI.VectorAffineTerm list of block `pos` after update! (moi_interop.jl:64-81), produced on the device:
        (records[batch?, nterms] of TERM_DTYPE, constants[batch?, rows])."""
        v = self.block_view(pos)
        b, nt, m = self.batch, v.nterms, v.rows
        dterms, dconst = self._dev_alloc(max(b * nt * 24, 8)), None   # two scratch device buffers owned by the caller, as a Julia/MOI consumer would
        try:
            dconst = self._dev_alloc(max(b * m * 8, 8))
            self._ck(self._lib.pmb_block_export_terms(self._ctx, pos, C.c_void_p(dterms), C.c_void_p(dconst)))
            recs = self._d2h(dterms, b * nt, self.TERM_DTYPE).reshape(b, nt)
            consts = self._d2h(dconst, b * m).reshape(b, m)
        finally:
            self._dev_free(dterms)
            self._dev_free(dconst)
        return (recs[0], consts[0]) if b == 1 else (recs, consts)

    QUAD_TERM_DTYPE = np.dtype([("coefficient", np.float64), ("variable_1", np.int64), ("variable_2", np.int64)])
    AFF_TERM_DTYPE = np.dtype([("coefficient", np.float64), ("variable", np.int64)])

    def export_objective_terms(self):
        """The objective as MOI.ScalarQuadraticFunction records (moi_interop.jl:45-62, combined form): (quadratic_terms[batch?, N(N+1)/2] or None
        for an affine objective, affine_terms[batch?, N], constant[batch?]) -- produced on the device from the last update."""
        v = self.objective_view()
        b, N = self.batch, v.nvars
        npairs = N * (N + 1) // 2
        dq = self._dev_alloc(max(b * npairs * 24, 8)) if v.Q else None
        da = dc = None
        try:
            da = self._dev_alloc(max(b * N * 16, 8))
            dc = self._dev_alloc(max(b * 8, 8))
            self._ck(self._lib.pmb_objective_export_terms(self._ctx, C.c_void_p(dq) if dq else None, C.c_void_p(da), C.c_void_p(dc)))
            quad = self._d2h(dq, b * npairs, self.QUAD_TERM_DTYPE).reshape(b, npairs) if dq else None
            aff = self._d2h(da, b * N, self.AFF_TERM_DTYPE).reshape(b, N)
            const = self._d2h(dc, b)
        finally:
            for p_ in (dq, da, dc):
                self._dev_free(p_)
        if b == 1:
            return (quad[0] if quad is not None else None), aff[0], float(const[0])
        return quad, aff, const

    def enable_qp_pack(self):
        self._ck(self._lib.pmb_enable_qp_pack(self._ctx))

    def qp_view(self):
        v = K.QpView()
        self._ck(self._lib.pmb_get_qp(self._ctx, C.byref(v)))
        return v

    def qp_host_sink(self):
        """Register pinned host buffers that every update() fills with the packed QP as part of the update graph (call after
        initialize()).  Returns the dict of numpy views; valid after synchronize()."""
        v, b = self.qp_view(), self.batch
        bufs = {"P_x": self._pinned((b, v.nnzP)), "q": self._pinned((b, v.n)), "c0": self._pinned((b,)),
                "A_x": self._pinned((b, v.nnzA)), "l": self._pinned((b, v.m)), "u": self._pinned((b, v.m))}
        ptr = lambda a: C.c_void_p(a.ctypes.data) if a.size else None
        self._ck(self._lib.pmb_qp_set_host_sink(self._ctx, ptr(bufs["P_x"]), ptr(bufs["q"]), ptr(bufs["c0"]), ptr(bufs["A_x"]),
                                                ptr(bufs["l"]), ptr(bufs["u"])))
        return bufs

    def qp_arrays(self):
        """Host copies of the OSQP-style QP (P upper-triangular CSC, q, c0, A CSC, l, u) -- values refreshed by every update()."""
        v, b = self.qp_view(), self.batch
        out = {"n": v.n, "m": v.m,
               "P_p": self._d2h(v.P_p, v.n + 1, np.int64), "P_i": self._d2h(v.P_i, v.nnzP, np.int64),
               "A_p": self._d2h(v.A_p, v.n + 1, np.int64), "A_i": self._d2h(v.A_i, v.nnzA, np.int64),
               "P_x": self._d2h(v.P_x, b * v.nnzP).reshape(b, v.nnzP), "q": self._d2h(v.q, b * v.n).reshape(b, v.n),
               "c0": self._d2h(v.c0, b), "A_x": self._d2h(v.A_x, b * v.nnzA).reshape(b, v.nnzA),
               "l": self._d2h(v.l, b * v.m).reshape(b, v.m), "u": self._d2h(v.u, b * v.m).reshape(b, v.m)}
        if b == 1:
            for k_ in ("P_x", "q", "c0", "A_x", "l", "u"):
                out[k_] = out[k_][0]
        return out

    def stats(self):
        s = K.Stats()
        self._ck(self._lib.pmb_get_stats(self._ctx, C.byref(s)))
        return {f: getattr(s, f) for f, _ in K.Stats._fields_}

    def update_count(self):
        n = C.c_int64()
        self._ck(self._lib.pmb_update_count(self._ctx, n))
        return n.value

    def enable_kernel_timing(self, on=True):
        self._ck(self._lib.pmb_enable_kernel_timing(self._ctx, 1 if on else 0))

    def kernel_times(self):
        out = {}
        for cls in range(self._lib.pmb_kernel_class_count()):
            ms, n = C.c_double(), C.c_int64()
            self._ck(self._lib.pmb_kernel_time_ms(self._ctx, cls, ms, n))
            out[self._lib.pmb_kernel_class_name(cls).decode()] = (ms.value, n.value)
        return out


# ---- free functions mirroring the names used inside @expression ------------------------------------
def _model_of(*args):
    for a in args:
        if isinstance(a, Expr):
            return a.model
    raise ValueError("at least one argument must be an expression of a model (constants fold on the host)")


def adjoint(x, model=None):
    m = model or _model_of(x)
    out = C.c_int32()
    m._ck(m._lib.pmb_op_adjoint(m._ctx, m._as_node(x), out))
    return Expr(m, out.value)


transpose = adjoint


def dot(x, y, model=None):
    m = model or _model_of(x, y)
    out = C.c_int32()
    m._ck(m._lib.pmb_op_dot(m._ctx, m._as_node(x), m._as_node(y), out))
    return Expr(m, out.value)


def bilinear(xT, Q, y, model=None):
    """`transpose(x) * Q * y` as ONE 3-argument `*` call (lazyexpression.jl:219-226)."""
    m = model or _model_of(xT, Q, y)
    out = C.c_int32()
    m._ck(m._lib.pmb_op_bilinear(m._ctx, m._as_node(xT), m._as_node(Q), m._as_node(y), out))
    return Expr(m, out.value)


def vcat(*parts, model=None):
    m = model or _model_of(*parts)
    nodes = (C.c_int32 * len(parts))(*[m._as_node(p) for p in parts])
    out = C.c_int32()
    m._ck(m._lib.pmb_op_vcat(m._ctx, len(parts), nodes, out))
    return Expr(m, out.value)


def convert(x, model=None):
    """`convert(Vector, x)` (lazyexpression.jl:280-282): any vector value as a Vector{AffineFunction}"""
    m = model or _model_of(x)
    out = C.c_int32()
    m._ck(m._lib.pmb_op_convert(m._ctx, m._as_node(x), out))
    return Expr(m, out.value)


def vect(x, model=None):
    m = model or _model_of(x)
    out = C.c_int32()
    m._ck(m._lib.pmb_op_vect(m._ctx, m._as_node(x), out))
    return Expr(m, out.value)


def objective(model, sense, e):
    """@objective(model, sense, expr), model.jl:267-271"""
    model.setobjective(sense, e)


def constraint(model, lhs, relation, rhs):
    """@constraint(model, lhs relation rhs), model.jl:224-249: the constraint function is `lhs - rhs`."""
    e = model.expr(lhs) - rhs
    if relation == ">=":
        return model.add_nonnegative_constraint(e)
    if relation == "<=":
        return model.add_nonpositive_constraint(e)
    if relation == "==":
        return model.add_zero_constraint(e)
    raise ValueError("Relation not recognized")  # model.jl:243
