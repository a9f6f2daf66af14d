"""ctypes binding of include/parametron_b200.h (the C ABI of libpmb200.so).

This is the binding the tests and bench.py exercise; the Julia `ccall` shim a Parametron.jl maintainer would add
binds exactly the same symbols.  There is NO CPU fallback: if the library is missing or no B200 is
visible every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpmb200.so")

OK, EINVAL, EDIM, EINITIALIZED, ENOTINIT, ECUDA, EUNSUPPORTED, ENOMEM = 0, -1, -2, -3, -4, -5, -6, -7
KIND_NUM_SCALAR, KIND_NUM_VECTOR, KIND_NUM_MATRIX, KIND_VARVEC, KIND_AFFVEC, KIND_AFF, KIND_QUAD, KIND_ADJVEC = range(8)
MINIMIZE, MAXIMIZE = 0, 1
SET_NONNEGATIVE, SET_NONPOSITIVE, SET_ZERO = 0, 1, 2
PARAM_DEVICE, PARAM_HOST, PARAM_UNIFORM = 0, 1, 2

c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)


class ObjectiveView(C.Structure):
    _fields_ = [("kind", C.c_int32), ("sense", C.c_int32), ("is_constant", C.c_int32), ("nvars", C.c_int64),
                ("Q", C.c_void_p), ("q", C.c_void_p), ("c0", C.c_void_p)]


class BlockView(C.Structure):
    _fields_ = [("set_kind", C.c_int32), ("is_scalar", C.c_int32), ("is_constant", C.c_int32), ("bucket", C.c_int32),
                ("rows", C.c_int64), ("nvars", C.c_int64), ("C", C.c_void_p), ("d", C.c_void_p), ("nterms", C.c_int64),
                ("is_sparse", C.c_int32), ("nnz", C.c_int64), ("vals", C.c_void_p), ("sp_rows", C.c_void_p), ("sp_cols", C.c_void_p),
                ("Q", C.c_void_p)]


class QpView(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("nnzP", C.c_int64), ("nnzA", C.c_int64), ("P_x", C.c_void_p), ("q", C.c_void_p),
                ("c0", C.c_void_p), ("A_x", C.c_void_p), ("l", C.c_void_p), ("u", C.c_void_p), ("P_p", C.c_void_p), ("P_i", C.c_void_p),
                ("A_p", C.c_void_p), ("A_i", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches_per_update", C.c_int64), ("memcpy_nodes_per_update", C.c_int64),
                ("device_bytes_allocated", C.c_int64), ("uses_cuda_graph", C.c_int64),
                ("affine_algorithmic_bytes", C.c_int64), ("gram_algorithmic_flops", C.c_int64)]


HOST_CALLBACK = C.CFUNCTYPE(None, c_f64p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p)

# name -> (restype, argtypes); every symbol include/parametron_b200.h declares
SIGNATURES = {
    "pmb_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "pmb_ctx_destroy": (C.c_int, [C.c_void_p]),
    "pmb_set_instance_offset": (C.c_int, [C.c_void_p, C.c_int64]),
    "pmb_get_instance_offset": (C.c_int, [C.c_void_p, c_i64p]),
    "pmb_last_error": (C.c_char_p, [C.c_void_p]),
    "pmb_version": (C.c_char_p, []),
    "pmb_add_variables": (C.c_int, [C.c_void_p, C.c_int64, c_i64p]),
    "pmb_num_variables": (C.c_int, [C.c_void_p, c_i64p]),
    "pmb_param_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, c_i32p]),
    "pmb_param_create_bsr": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, c_i32p]),
    "pmb_param_set_device_ptr": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pmb_param_set_host_ptr": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pmb_param_set_uniform": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_double]),
    "pmb_param_upload": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pmb_param_device_ptr": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "pmb_const_numeric": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, c_i32p]),
    "pmb_varvec": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, c_i32p]),
    "pmb_const_affine": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, c_i32p]),
    "pmb_const_quadratic": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_double, c_i32p]),
    "pmb_op_matvec": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_i32p]),
    "pmb_op_adjoint": (C.c_int, [C.c_void_p, C.c_int32, c_i32p]),
    "pmb_op_bilinear": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_i32p]),
    "pmb_op_dot": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_i32p]),
    "pmb_op_add": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_i32p]),
    "pmb_op_sub": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_i32p]),
    "pmb_op_mul": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_i32p]),
    "pmb_op_vcat": (C.c_int, [C.c_void_p, C.c_int32, c_i32p, c_i32p]),
    "pmb_op_vect": (C.c_int, [C.c_void_p, C.c_int32, c_i32p]),
    "pmb_op_convert": (C.c_int, [C.c_void_p, C.c_int32, c_i32p]),
    "pmb_op_getfield": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int, C.c_int64, C.c_int64, c_i32p]),
    "pmb_node_info": (C.c_int, [C.c_void_p, C.c_int32, c_i32p, c_i64p, c_i64p, c_i32p]),
    "pmb_set_objective": (C.c_int, [C.c_void_p, C.c_int, C.c_int32]),
    "pmb_add_constraint": (C.c_int, [C.c_void_p, C.c_int32, C.c_int, c_i32p]),
    "pmb_initialize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pmb_update": (C.c_int, [C.c_void_p]),
    "pmb_synchronize": (C.c_int, [C.c_void_p]),
    "pmb_update_count": (C.c_int, [C.c_void_p, c_i64p]),
    "pmb_get_objective": (C.c_int, [C.c_void_p, C.POINTER(ObjectiveView)]),
    "pmb_num_blocks": (C.c_int, [C.c_void_p, c_i32p]),
    "pmb_get_block": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(BlockView)]),
    "pmb_block_term_index": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pmb_block_export_terms": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "pmb_objective_export_terms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pmb_enable_qp_pack": (C.c_int, [C.c_void_p]),
    "pmb_get_qp": (C.c_int, [C.c_void_p, C.POINTER(QpView)]),
    "pmb_qp_set_host_sink": (C.c_int, [C.c_void_p] * 7),
    "pmb_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "pmb_set_primal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "pmb_value": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]),
    "pmb_objective_value": (C.c_int, [C.c_void_p, c_f64p]),
    "pmb_set_gram_mode": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "pmb_gram_level_counts": (C.c_int, [C.c_void_p, c_i64p]),
    "pmb_enable_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "pmb_kernel_class_count": (C.c_int, []),
    "pmb_kernel_class_name": (C.c_char_p, [C.c_int]),
    "pmb_kernel_time_ms": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_i64p]),
    "pmb_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "pmb_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pmb_device_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "pmb_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pmb_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pmb_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
}

_lib = None


def load():
    """dlopen libpmb200.so and bind every declared symbol.  Raises (never falls back) when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not found: build it with `python parametron.jl_b200/build.py` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PmbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class DimensionMismatch(PmbError):  # functions.jl:669,755,780-781,845,877
    pass


class AlreadyInitialized(PmbError):  # model.jl:50,61,69
    pass


class Unsupported(PmbError):
    pass


def check(lib, ctx, code):
    if code == OK:
        return
    msg = lib.pmb_last_error(ctx)
    msg = msg.decode() if msg else f"pmb error {code}"
    cls = {EDIM: DimensionMismatch, EINITIALIZED: AlreadyInitialized, EUNSUPPORTED: Unsupported}.get(code, PmbError)
    raise cls(code, msg)
