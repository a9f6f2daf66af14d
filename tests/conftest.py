import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_available():
    try:
        import ctypes
        import parametron_b200 as pm
        lib = pm._capi.load()
        ctx = ctypes.c_void_p()
        if lib.pmb_ctx_create(0, None, 1, ctypes.byref(ctx)) != 0:
            return False
        lib.pmb_ctx_destroy(ctx)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a B200: skip (not fail) them on hosts where the library cannot create a context"""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _gpu_available():
        return
    skip = pytest.mark.skip(reason="no sm_100 device / libpmb200.so context (these tests run on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)
