"""`@expression` / `@constraint` / `@objective` as string "macros": the host-side analogue of src/lazyexpression.jl:140-180 and
src/model.jl:224-271.  The expression text is parsed (Python `ast`; Julia's `'` adjoint and `⋅` dot are pre-translated), every
call / operator node is lowered bottom-up to the matching `pmb_op_*` opcode -- one opcode per rewrite rule, exactly what the
reference's postwalk + `optimize_toplevel` does -- and names are looked up in the caller-supplied environment (Parameters,
Variables, vectors of Variables, numbers, numpy arrays).  Unhandled syntax raises ArgumentError like the macro does
(lazyexpression.jl:165-176).

    r = expression(model, "A * x - b", A=A, x=x, b=b)
    objective(model, Minimize, "r ⋅ r", r=r)                       # model.jl:267-271
    constraint(model, "C * x == d", C=C, x=x, d=d)                 # model.jl:224-249
    expression(model, "transpose(x) * Q * x + r ⋅ x + s", ...)     # 3-argument `*` rule, lazyexpression.jl:219-226
"""
import ast
import re

import numpy as np

from . import model as M


class ArgumentError(ValueError):
    pass


_DOT = "__pmb_dot__"


def _translate(text):
    """Julia surface syntax -> parseable Python: `a ⋅ b` -> dot infix marker, postfix `'` -> .T, `[a; b]` -> vcat(a, b)."""
    t = text.replace("⋅", f" @{_DOT}@ ")
    t = re.sub(r"(\w|\)|\])'", r"\1.T", t)
    t = re.sub(r"\[([^\[\];]+(?:;[^\[\];]+)+)\]", lambda m_: "vcat(" + ", ".join(p.strip() for p in m_.group(1).split(";")) + ")", t)
    return t


def _top_level_split(t, seps):
    """split `t` at top-level (outside (), []) occurrences of any separator in `seps`; returns (pieces, separators)"""
    pieces, ops, depth, cur, i = [], [], 0, "", 0
    while i < len(t):
        ch = t[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        hit = None
        if depth == 0:
            for sp in seps:
                if t.startswith(sp, i):
                    hit = sp
                    break
        if hit is not None and (hit not in "+-" or (cur.strip() and not cur.rstrip().endswith(("*", "/", "@", "(", ",")))):
            pieces.append(cur); ops.append(hit); cur = ""; i += len(hit)
        else:
            cur += ch; i += 1
    pieces.append(cur)
    return pieces, ops


def _split_dots(t):
    """`⋅` has the precedence of `*` in Julia (left-associative): rewrite `a * b ⋅ c + d` as `dot(a * b, c) + d`, recursing
    into parenthesised groups and call arguments."""
    marker = f"@{_DOT}@"
    if marker not in t:
        return t
    # 1. recurse into top-level parenthesised groups
    out, i = "", 0
    while i < len(t):
        if t[i] == "(":
            depth, k = 1, i + 1
            while k < len(t) and depth:
                depth += t[k] == "("
                depth -= t[k] == ")"
                k += 1
            inner = t[i + 1:k - 1]
            args, _ = _top_level_split(inner, [","])
            out += "(" + ", ".join(_split_dots(a) for a in args) + ")"
            i = k
        else:
            out += t[i]; i += 1
    # 2. additive terms, then multiplicative factors folded left
    terms, addops = _top_level_split(out, ["+", "-"])
    res = ""
    for n_, term in enumerate(terms):
        factors, mulops = _top_level_split(term, [marker, "*"])
        acc = factors[0].strip()
        for op, f in zip(mulops, factors[1:]):
            acc = f"dot({acc}, {f.strip()})" if op == marker else f"{acc} * {f.strip()}"
        res += acc if n_ == 0 else f" {addops[n_ - 1]} {acc}"
    return res


class _Lower(ast.NodeVisitor):
    def __init__(self, model, env):
        self.m, self.env = model, env

    def lower(self, text):
        src = _split_dots(_translate(text)).strip()
        try:
            tree = ast.parse(src, mode="eval")
        except SyntaxError as e:
            raise ArgumentError(f"Unhandled expression: {text!r} ({e.msg})")
        return self.visit(tree.body)

    def _val(self, x):
        return x

    def generic_visit(self, node):
        raise ArgumentError(f"Unhandled expression head: {type(node).__name__}")  # lazyexpression.jl:165-176

    def visit_Name(self, node):
        if node.id not in self.env:
            raise ArgumentError(f"unknown name {node.id!r} in expression")
        return self.env[node.id]

    def visit_Constant(self, node):
        if not isinstance(node.value, (int, float)):
            raise ArgumentError(f"Unhandled literal {node.value!r}")
        return float(node.value)

    def visit_UnaryOp(self, node):
        v = self.visit(node.operand)
        if isinstance(node.op, ast.USub):
            return -v if not isinstance(v, M.Expr) else (-1.0) * v
        if isinstance(node.op, ast.UAdd):
            return v
        raise ArgumentError(f"Unhandled expression head: {type(node.op).__name__}")

    def visit_Attribute(self, node):
        v = self.visit(node.value)
        if isinstance(v, M.RecordParameter) and node.attr in v._fields:   # p.field: the GetField rule, lazyexpression.jl:300-302
            return v.field(node.attr)
        if node.attr == "T":  # Julia's postfix '
            return M.adjoint(v, model=self.m)
        raise ArgumentError("getproperty: only fields of a struct-valued Parameter (RecordParameter) are lowered (lazyexpression.jl:300-302)")

    def _is_sym(self, v):
        return isinstance(v, (M.Expr, M.Variable)) or M._flatten_vars(v) is not None

    def visit_BinOp(self, node):
        # 3-argument `*`:  transpose(x) * Q * y parses as ((x' * Q) * y): recognise the bilinear rule first
        if isinstance(node.op, ast.Mult) and isinstance(node.left, ast.BinOp) and isinstance(node.left.op, ast.Mult):
            a, Q, y = self.visit(node.left.left), self.visit(node.left.right), self.visit(node.right)
            if isinstance(a, M.Expr) and a.kind == M.K.KIND_ADJVEC and isinstance(Q, M.Expr) and Q.kind == M.K.KIND_NUM_MATRIX and \
                    M._flatten_vars(y) is not None:
                try:
                    return M.bilinear(a, Q, y, model=self.m)
                except M.K.Unsupported:
                    pass
            left = self._binop(ast.Mult(), a, Q)
            return self._binop(node.op, left, y)
        return self._binop(node.op, self.visit(node.left), self.visit(node.right))

    def _binop(self, op, a, b):
        if not (self._is_sym(a) or self._is_sym(b)):  # constant folding on the host (lazyexpression.jl:189-192)
            a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
            if isinstance(op, ast.Add):
                return a + b
            if isinstance(op, ast.Sub):
                return a - b
            if isinstance(op, (ast.Mult, ast.MatMult)):
                return a @ b if a.ndim and b.ndim else a * b
            raise ArgumentError(f"Unhandled operator {type(op).__name__}")
        ea = a if isinstance(a, M.Expr) else self.m.expr(a)
        if isinstance(op, ast.Add):
            return ea + b
        if isinstance(op, ast.Sub):
            return ea - b
        if isinstance(op, (ast.Mult, ast.MatMult)):
            return ea * b
        raise ArgumentError(f"Unhandled operator {type(op).__name__}")

    def visit_Call(self, node):
        if not isinstance(node.func, ast.Name):
            raise ArgumentError("Unhandled expression head: call of a non-name")
        f = node.func.id
        if f == "convert" and len(node.args) == 2:   # convert(Vector, x): the type argument is only a tag (lazyexpression.jl:280-282)
            return M.convert(self.visit(node.args[1]), model=self.m)
        args = [self.visit(a) for a in node.args]
        if f == "dot" and len(args) == 2:
            return M.dot(args[0], args[1], model=self.m)
        if f in ("adjoint", "transpose") and len(args) == 1:
            return M.adjoint(args[0], model=self.m)
        if f == "vcat":
            return M.vcat(*args, model=self.m)
        if f == "vect" and len(args) == 1:
            return M.vect(args[0], model=self.m)
        raise ArgumentError(f"function {f!r} has no device lowering (only the rewrite table of lazyexpression.jl:198-302 is lowered)")

    def visit_List(self, node):  # [x] : Base.vect
        if len(node.elts) == 1:
            return M.vect(self.visit(node.elts[0]), model=self.m)
        raise ArgumentError("Unhandled expression head: vect with several elements (use [a; b] for vcat)")


def expression(model, text, **env):
    """@expression <text>  (src/lazyexpression.jl:140-180)"""
    v = _Lower(model, env).lower(text)
    return v if isinstance(v, M.Expr) else model.expr(v)


def objective(model, sense, text, **env):
    """@objective(model, sense, <text>)  (src/model.jl:267-271)"""
    model.setobjective(sense, expression(model, text, **env))


def constraint(model, text, **env):
    """@constraint(model, lhs <rel> rhs) with rel in ==, <=, >=  (src/model.jl:224-249)"""
    m_ = re.search(r"(==|<=|>=)", text)
    if not m_ or len(re.findall(r"(==|<=|>=)", text)) != 1:
        raise ArgumentError("Relation not recognized")  # model.jl:243,246
    lhs, rhs = text[:m_.start()], text[m_.end():]
    low = _Lower(model, env)
    return M.constraint(model, low.lower(lhs), m_.group(1), low.lower(rhs))
