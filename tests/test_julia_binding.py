"""The Julia side (julia/ParametronB200) cannot run here (no Julia in the image or on the GPU box), so its `ccall` layer is checked
statically: every `ccall((:pmb_*, lib), Ret, (ArgTypes...), ...)` found in the .jl sources is compared -- symbol name, return type, arity and
the machine type of every argument -- with the ctypes binding that the GPU tests exercise (parametron.jl_b200/_capi.py: SIGNATURES, itself
checked against include/parametron_b200.h by test_capi_symbols.py).  Also: the package layout a Julia user needs is present, README
Example 1 (reference README.md:31-56) appears in the package test with only the optimizer line changed, and the reference-side timing
script of SURVEY 8(d)(1) exists."""
import ctypes as C
import glob
import os
import re

import parametron_b200 as pm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = sorted(glob.glob(os.path.join(ROOT, "julia", "**", "*.jl"), recursive=True))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _category_julia(t):
    t = t.strip()
    if t in ("Cint", "Int32"):
        return "i32"
    if t == "Int64":
        return "i64"
    if t in ("UInt64", "Csize_t"):     # LP64: size_t is the 64-bit unsigned integer (ctypes aliases c_size_t and c_uint64 to c_ulong)
        return "u64"
    if t in ("Cdouble", "Float64"):
        return "f64"
    if t == "Cstring":
        return "cstr"
    if t.startswith("Ptr{") or t.startswith("Ref{") or t == "Any":
        return "ptr"
    raise AssertionError(f"unmapped Julia ccall type {t!r}")


def _category_ctypes(t):
    if t is None:
        return "void"
    if t in (C.c_int, C.c_int32):
        return "i32"
    if t is C.c_int64:
        return "i64"
    if t in (C.c_uint64, C.c_size_t):
        return "u64"
    if t is C.c_double:
        return "f64"
    if t is C.c_char_p:
        return "cstr"
    if t is C.c_void_p or hasattr(t, "_type_") and hasattr(t, "contents") or issubclass(t, C._Pointer):
        return "ptr"
    raise AssertionError(f"unmapped ctypes type {t!r}")


def _ccalls():
    """[(file, symbol, ret, [argtypes])] for literal symbols; @eval-generated ones are expanded from their (julia_name, c_symbol) tables"""
    found = []
    for path in JL:
        text = open(path).read()
        for mt in re.finditer(r"ccall\(\s*\(\s*(:\w+|\$\(QuoteNode\(c\)\))\s*,\s*lib\s*\)\s*,", text):
            # parse "Ret, (types...)" after the match with bracket counting
            i = mt.end()
            depth, j = 0, i
            while True:
                ch = text[j]
                if ch in "({[":
                    depth += 1
                elif ch in ")}]":
                    if depth == 0:
                        break
                    depth -= 1
                j += 1
            parts = _split_top(text[i:j])
            ret, argt = parts[0], parts[1]
            assert argt.startswith("(") and argt.endswith(")"), (path, argt)
            args = _split_top(argt[1:-1])
            sym = mt.group(1)
            if sym.startswith(":"):
                found.append((path, sym[1:], ret, args))
            else:   # for (jl, c) in ((:matvec, :pmb_op_matvec), ...) just above
                head = text[:mt.start()]
                loop = head[head.rfind("for (jl, c) in"):]
                for c_sym in re.findall(r"\(:\w+,\s*:(pmb_\w+)\)", loop.split("\n")[0]):
                    found.append((path, c_sym, ret, args))
    return found


def test_julia_sources_are_present():
    names = {os.path.relpath(p, ROOT) for p in JL}
    for need in ("julia/ParametronB200/src/ParametronB200.jl", "julia/ParametronB200/src/libpmb.jl", "julia/ParametronB200/src/lowering.jl",
                 "julia/ParametronB200/src/model_hooks.jl", "julia/ParametronB200/test/readme_example.jl"):
        assert need in names, need
    assert os.path.exists(os.path.join(ROOT, "julia", "ParametronB200", "Project.toml"))
    assert os.path.exists(os.path.join(ROOT, "baseline", "julia", "bench_update.jl"))


def test_every_ccall_matches_the_tested_binding():
    calls = _ccalls()
    assert len(calls) >= 35
    seen = set()
    for path, sym, ret, args in calls:
        assert sym in pm._capi.SIGNATURES, f"{os.path.basename(path)}: {sym} is not a symbol of include/parametron_b200.h"
        res, argtypes = pm._capi.SIGNATURES[sym]
        assert _category_julia(ret) == _category_ctypes(res), (sym, ret, res)
        assert len(args) == len(argtypes), f"{sym}: Julia passes {len(args)} arguments, the C ABI takes {len(argtypes)}"
        for k, (jt, ct) in enumerate(zip(args, argtypes)):
            assert _category_julia(jt) == _category_ctypes(ct), f"{sym} argument {k}: Julia {jt} vs C {ct}"
        seen.add(sym)
    # the calls the update! path needs are all bound
    for need in ("pmb_ctx_create", "pmb_ctx_destroy", "pmb_add_variables", "pmb_param_create", "pmb_param_set_host_ptr", "pmb_const_numeric", "pmb_varvec",
                 "pmb_const_affine", "pmb_const_quadratic", "pmb_op_matvec", "pmb_op_dot", "pmb_op_add", "pmb_op_sub", "pmb_op_mul", "pmb_op_adjoint",
                 "pmb_op_bilinear", "pmb_op_vcat", "pmb_op_vect", "pmb_op_convert", "pmb_set_objective", "pmb_add_constraint", "pmb_initialize", "pmb_update",
                 "pmb_synchronize", "pmb_get_objective", "pmb_get_block", "pmb_block_export_terms", "pmb_objective_export_terms", "pmb_device_alloc",
                 "pmb_memcpy_d2h", "pmb_set_primal", "pmb_value", "pmb_objective_value", "pmb_last_error", "pmb_set_instance_offset"):
        assert need in seen, f"{need} has no Julia binding"


def test_julia_structs_mirror_the_c_views():
    """ObjectiveView / BlockView in libpmb.jl must list the fields of pmb_objective_view / pmb_block_view in order (same as the ctypes Structures)"""
    text = open(os.path.join(ROOT, "julia", "ParametronB200", "src", "libpmb.jl")).read()
    for jl_name, ct in (("ObjectiveView", pm._capi.ObjectiveView), ("BlockView", pm._capi.BlockView)):
        body = re.search(r"struct " + jl_name + r"\n(.*?)\nend", text, flags=re.S).group(1)
        fields = re.findall(r"(\w+)::", body)
        assert fields == [f for f, _ in ct._fields_], (jl_name, fields)


def test_readme_example_is_unchanged_but_for_the_optimizer():
    """the statements of README Example 1 that define the model (reference README.md:44-56) appear verbatim in the package test"""
    text = open(os.path.join(ROOT, "julia", "ParametronB200", "test", "readme_example.jl")).read()
    for stmt in ("model = Model(optimizer)", "x = [Variable(model) for _ = 1 : n]", "residual = @expression A * x - b",
                 "@objective(model, Minimize, residual ⋅ residual)", "@constraint(model, C * x == d)", "solve!(model)"):
        assert stmt in text, stmt
    assert "B200Optimizer(Optimizer())" in text


def test_lowering_covers_every_rewrite_rule():
    """one case per in-place kernel the reference's `optimize` table can produce (lazyexpression.jl:198-302)"""
    text = open(os.path.join(ROOT, "julia", "ParametronB200", "src", "lowering.jl")).read()
    for fn in ("matvecmul!", "vecdot!", "add!", "vecadd!", "subtract!", "vecsubtract!", "mul!", "scale!", "bilinearmul!", "vcat!", "copyto!", "GetField",
               "adjoint", "Base.vect", "convert", "identity"):
        assert fn in text, fn
