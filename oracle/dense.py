"""ORACLE (test infrastructure only -- never imported by the product path).

Tier-2 oracle: dense numpy restatement of what the reference's term lists MEAN,
used (a) to compare device arrays with the literal term-list oracle after
densification and (b) as the comparator at sizes where the literal term list
cannot exist (n=4096: 1.65 TB, SURVEY.md section 7 "Hard parts").

Conventions restated here
* MathOptInterface ~0.8 `ScalarQuadraticFunction` = 1/2 x'Qx + a'x + c, Q symmetric, duplicate
  terms summed, an off-diagonal term (i, j, c) contributes c to BOTH Q[i,j] and Q[j,i]
  (published MOI definition; the dependency is not vendored under /root/reference).  With the
  diagonal doubling of src/moi_interop.jl:58 this gives  Q = M + M'  where M[i,j] is the sum of
  native coefficients of ordered pairs (rowvar=i, colvar=j).
* `VectorAffineFunction` rows: src/moi_interop.jl:64-81.

Parity status: PINNED transitively -- equal to the literal oracle (oracle/functions.py, itself
pinned on the reference's known answers) on seeded inputs in tests/test_oracle_dense.py, and
against the comment `test/model.jl:84` (||Ax-b||^2 = x'A'Ax - (2A'b)'x + b'b).
"""
from __future__ import annotations

import numpy as np


def moi_quadratic_to_dense(moi_f, nvars: int):
    """(Q, q, c0) with Q the symmetric MOI Hessian; summation in term order."""
    Q = np.zeros((nvars, nvars))
    q = np.zeros(nvars)
    for coeff, i, j in moi_f.quadratic_terms:
        if i == j:
            Q[i - 1, i - 1] += coeff
        else:
            Q[i - 1, j - 1] += coeff
            Q[j - 1, i - 1] += coeff
    for coeff, i in moi_f.affine_terms:
        q[i - 1] += coeff
    return Q, q, float(moi_f.constant)


def moi_scalar_affine_to_dense(moi_f, nvars: int):
    q = np.zeros(nvars)
    for coeff, i in moi_f.terms:
        q[i - 1] += coeff
    return q, float(moi_f.constant)


def moi_vector_affine_to_dense(moi_f, nvars: int):
    """(C, d) with C[row, var] the summed coefficients and d the constants."""
    m = len(moi_f.constants)
    C = np.zeros((m, nvars))
    for row, coeff, var in moi_f.terms:
        C[row - 1, var - 1] += coeff
    return C, np.asarray(moi_f.constants, dtype=float)


def native_quadratic_to_ordered_pairs(f, nvars: int):
    """M[i,j] = sum of coefficients of native QuadraticTerms with (rowvar=i, colvar=j) (no symmetrisation)."""
    M = np.zeros((nvars, nvars))
    for t in f.quadratic:
        M[t.rowvar.index - 1, t.colvar.index - 1] += t.coeff
    return M


# ---- dense restatement of the expression shapes ------------------------------------------
def least_squares_objective(A, b, xidx, nvars, varmap=None):
    """`residual . residual` with residual = A*x - b  ->  (Q = 2A'A, q = -2A'b, c0 = b'b)   [test/model.jl:84]."""
    cols = _cols(xidx, varmap)
    Q = np.zeros((nvars, nvars))
    q = np.zeros(nvars)
    G = 2.0 * (A.T @ A)
    np.add.at(Q, (cols[:, None], cols[None, :]), G)
    np.add.at(q, cols, -2.0 * (A.T @ b))
    return Q, q, float(b @ b)


def affine_block(C, d, xidx, nvars, varmap=None, sign=-1.0):
    """`C*x - d`  ->  block (C scattered by variable column, -d)."""
    cols = _cols(xidx, varmap)
    out = np.zeros((C.shape[0], nvars))
    np.add.at(out, (slice(None), cols), C)
    return out, sign * np.asarray(d, dtype=float)


def _cols(xidx, varmap):
    xidx = np.asarray(xidx, dtype=np.int64)
    if varmap is not None:
        varmap = np.asarray(varmap, dtype=np.int64)
        xidx = varmap[xidx - 1]
    return xidx - 1
