"""ORACLE (test infrastructure only).  ctypes wrapper of oracle/termlist.c -- the literal single-threaded C restatement of
the reference's term-list `update!` for the README-Example-1 model shape (see the header of termlist.c for file:line).
Used by tests/ (validated against the Python oracle) and by bench.py's cpu_baseline / --impl reference leg."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libtermlist.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        lib = C.CDLL(_LIB)
        lib.tl_create.restype = C.c_void_p
        lib.tl_create.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
        lib.tl_destroy.argtypes = [C.c_void_p]
        lib.tl_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        for f in ("tl_num_quadratic_terms", "tl_num_affine_terms", "tl_num_vector_terms"):
            getattr(lib, f).restype = C.c_size_t
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.tl_objective_constant.restype = C.c_double
        lib.tl_objective_constant.argtypes = [C.c_void_p]
        lib.tl_get_quadratic_terms.argtypes = [C.c_void_p] * 4
        lib.tl_get_affine_terms.argtypes = [C.c_void_p] * 3
        lib.tl_get_vector_terms.argtypes = [C.c_void_p] * 5
        _lib = lib
    return _lib


class TermListModel:
    """minimize ||A x - b||^2 s.t. C x == d with A: k x n, C: m x n (column-major float64 arrays, order='F')."""

    def __init__(self, n, k, m, varmap=None):
        self.lib = load()
        self.n, self.k, self.m = n, k, m
        vm = None if varmap is None else np.ascontiguousarray(varmap, dtype=np.int64)
        self.h = C.c_void_p(self.lib.tl_create(n, k, m, vm.ctypes.data if vm is not None else None))

    def close(self):
        if self.h:
            self.lib.tl_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def update(self, A, b, Cm, d, row_limit=0, parts=3):
        A, Cm = np.asfortranarray(A, dtype=np.float64), np.asfortranarray(Cm, dtype=np.float64)
        b, d = np.ascontiguousarray(b, dtype=np.float64), np.ascontiguousarray(d, dtype=np.float64)
        assert A.shape == (self.k, self.n) and Cm.shape == (self.m, self.n) and b.shape == (self.k,) and d.shape == (self.m,)
        self.lib.tl_update(self.h, A.ctypes.data, b.ctypes.data, Cm.ctypes.data, d.ctypes.data, row_limit, parts)

    def quadratic_terms(self):
        n = self.lib.tl_num_quadratic_terms(self.h)
        c, v1, v2 = np.empty(n), np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        self.lib.tl_get_quadratic_terms(self.h, c.ctypes.data, v1.ctypes.data, v2.ctypes.data)
        return c, v1, v2

    def affine_terms(self):
        n = self.lib.tl_num_affine_terms(self.h)
        c, v = np.empty(n), np.empty(n, dtype=np.int64)
        self.lib.tl_get_affine_terms(self.h, c.ctypes.data, v.ctypes.data)
        return c, v, self.lib.tl_objective_constant(self.h)

    def vector_terms(self):
        n = self.lib.tl_num_vector_terms(self.h)
        r, c, v, k = np.empty(n, dtype=np.int64), np.empty(n), np.empty(n, dtype=np.int64), np.empty(self.m)
        self.lib.tl_get_vector_terms(self.h, r.ctypes.data, c.ctypes.data, v.ctypes.data, k.ctypes.data)
        return r, c, v, k
