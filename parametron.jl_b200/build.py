"""Builds libpmb200.so (the C-ABI shared library of include/parametron_b200.h) in-tree with nvcc for sm_100a.

Usage: python parametron.jl_b200/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libpmb200.so")
SOURCES = ["kernels_affine.cu", "kernels_affine_tma.cu", "kernels_gram.cu", "kernels_gram_tma.cu", "kernels_gram_i8.cu", "kernels_gram_i8_fused.cu", "kernels_misc.cu", "kernels_export.cu", "pmb_export.cu", "pmb_build.cu", "pmb_plan.cu"]
HEADERS = ["pmb_kernels.h", "pmb_internal.h", "pmb_host.h", "pmb_i8_ptx.h", "pmb_philox.h", "kernels_gram_i8_fused_body.inc", os.path.join("..", "..", "include", "parametron_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
         "-Xptxas", "-v", "--use_fast_math=false"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]
# extra nvcc flags for diagnostic builds, e.g. PMB_NVCC_EXTRA=-DFZ_WATCHDOG (kernels_gram_i8_fused.cu: watchdog warp + milestone record); forces a rebuild
EXTRA = os.environ.get("PMB_NVCC_EXTRA", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    force = force or bool(EXTRA)
    os.makedirs(OUT_DIR, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + EXTRA + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(o + ".ptxas.log", "w") as f:
                f.write(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on " + src)
            if verbose:
                sys.stderr.write(r.stderr)
    if force or _stale(LIB, objs):
        tmp = LIB + ".tmp"   # link beside the target and rename: a snapshot of the tree never sees a half-written library
        cmd = [NVCC, "-shared", "-o", tmp] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
