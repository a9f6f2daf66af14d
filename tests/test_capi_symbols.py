"""CPU-side checks of the C ABI: the shared library loads, exports every symbol include/parametron_b200.h declares, the
ctypes binding covers exactly that set, and (no GPU here) context creation fails LOUDLY instead of falling back."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry
import parametron_b200 as pm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "parametron_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmb_[a-z0-9_]+)\s*\(", src)) - {"pmb_host_callback"})


def test_library_exports_every_declared_symbol():
    entry.build()
    lib = ctypes.CDLL(pm.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(pm._capi.SIGNATURES) == names  # the binding covers exactly the header


def test_version_and_no_cpu_fallback():
    lib = pm._capi.load()
    assert b"sm_100a" in lib.pmb_version()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for CPU-only hosts")
    with pytest.raises(pm.PmbError):
        pm.Model()
    ctx = ctypes.c_void_p()
    assert lib.pmb_ctx_create(0, None, 1, ctypes.byref(ctx)) < 0 and not ctx.value


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "parametron.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", text, flags=re.M), f
            if f.endswith((".cu", ".h")):
                assert "oracle/" not in open(os.path.join(dirpath, f)).read(), f
