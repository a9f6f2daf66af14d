/* ORACLE (test infrastructure only -- never linked into or called by the product path; only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this).
 *
 * Literal single-threaded C restatement of the reference's term-list hot path for the README-Example-1 model shape
 *     minimize ||A x - b||^2   subject to  C x == d                      (README.md:31-56, test/model.jl:83-125)
 * i.e. exactly the work `update!(model)` (src/model.jl:132-143) performs per call:
 *   objective.expr():  vecdot!(destQ, residual(), residual())            src/functions.jl:730, :702-709
 *       residual() = vecsubtract!(destR, matvecmul!(destAx, A, x), b)    :751-764 -> :502 -> :422-427,:474 ; :775-798
 *       (evaluated TWICE: LazyExpressions are not memoised, only Parameters are -- src/lazyexpression.jl:50-61)
 *       muladd!(destQ, r_i, r_i) per row: n^2 QuadraticTerms + 2n LinearTerms, never combined      :548-576
 *   update!(::MOI.ScalarQuadraticFunction, f, varmap)  copy + diagonal x2 + index remap            src/moi_interop.jl:45-62
 *   constraint.expr(): vecsubtract!(.., matvecmul!(.., C, x), d); update!(::MOI.VectorAffineFunction) row-major flatten :64-81
 * Structs mirror the Julia isbits layouts: LinearTerm{Float64} 16 B (:110-113), QuadraticTerm{Float64} 24 B (:136-140),
 * MOI.VectorAffineTerm 24 B, MOI.ScalarQuadraticTerm 24 B.
 *
 * `row_limit` (bench sampling): only the first row_limit residual rows enter vecdot!/the MOI copy; per-row work is identical
 * (n^2 + 2n terms), so objective time scales exactly linearly in the row count.  row_limit = k is the full reference work.
 *
 * Parity status: pinned against oracle/functions.py + oracle/model.py (themselves pinned on the reference's known answers)
 * in tests/test_oracle_termlist.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double coeff; int64_t var; } LinearTerm;
typedef struct { double coeff; int64_t rowvar, colvar; } QuadraticTerm;
typedef struct { LinearTerm* linear; size_t n, cap; double constant; } AffineFunction;
typedef struct { QuadraticTerm* quadratic; size_t n, cap; AffineFunction affine; } QuadraticFunction;
typedef struct { int64_t output_index; double coeff; int64_t var; } MoiVectorAffineTerm;
typedef struct { double coeff; int64_t var1, var2; } MoiScalarQuadraticTerm;
typedef struct { double coeff; int64_t var; } MoiScalarAffineTerm;

typedef struct {
    int64_t n, k, m;
    int64_t* x;              /* Variable indices 1..n */
    int64_t* varmap;         /* model_var_to_optimizer, 1-based values, indexed var-1 */
    AffineFunction *Ax, *res, *Cx, *con;     /* preallocated dests (lazyexpression.jl:202 deepcopy(expr())) */
    QuadraticFunction obj;
    /* MOI function objects (moi_interop.jl:87-92) */
    MoiScalarQuadraticTerm* moi_quad; size_t moi_nquad, moi_quad_cap;
    MoiScalarAffineTerm* moi_aff; size_t moi_naff, moi_aff_cap;
    double moi_const;
    MoiVectorAffineTerm* moi_vterms; size_t moi_nv, moi_v_cap;
    double* moi_vconst;
} tl_model;

static void lin_resize(AffineFunction* f, size_t n) {      /* resize!(f.linear, n) */
    if (n > f->cap) { f->linear = (LinearTerm*)realloc(f->linear, n * sizeof(LinearTerm)); f->cap = n; }
    f->n = n;
}
static void quad_resize(QuadraticFunction* f, size_t n) {
    if (n > f->cap) { f->quadratic = (QuadraticTerm*)realloc(f->quadratic, n * sizeof(QuadraticTerm)); f->cap = n; }
    f->n = n;
}

/* matvecmul!(y, A, x::Vector{Variable})  functions.jl:775-798: y[row].linear[col] = LinearTerm(A[i], x[col]), i column-major */
static void matvecmul(AffineFunction* y, const double* A, int64_t rows, int64_t cols, const int64_t* x) {
    for (int64_t r = 0; r < rows; ++r) { lin_resize(&y[r], (size_t)cols); y[r].constant = 0.0; }
    size_t i = 0;
    for (int64_t col = 0; col < cols; ++col) {
        int64_t xcol = x[col];
        for (int64_t row = 0; row < rows; ++row) {
            y[row].linear[col].coeff = A[i++];
            y[row].linear[col].var = xcol;
        }
    }
}

/* vecsubtract!(dest, x, b) functions.jl:751-764 -> subtract!(dest[i], x[i], b[i]) :502 = copyto! (:422-427) then constant -= b */
static void vecsubtract(AffineFunction* dest, const AffineFunction* x, const double* b, int64_t rows) {
    for (int64_t r = 0; r < rows; ++r) {
        lin_resize(&dest[r], x[r].n);
        memcpy(dest[r].linear, x[r].linear, x[r].n * sizeof(LinearTerm));
        dest[r].constant = x[r].constant;
        dest[r].constant -= b[r];
    }
}

/* muladd!(dest::QuadraticFunction, x::AffineFunction, y::AffineFunction)  functions.jl:548-576 */
static void muladd_quad(QuadraticFunction* dest, const AffineFunction* x, const AffineFunction* y) {
    size_t quadoffset = dest->n;
    quad_resize(dest, quadoffset + x->n * y->n);
    QuadraticTerm* q = dest->quadratic + quadoffset;
    size_t k = 0;
    for (size_t i = 0; i < x->n; ++i)
        for (size_t j = 0; j < y->n; ++j) {
            q[k].coeff = x->linear[i].coeff * y->linear[j].coeff;
            q[k].rowvar = x->linear[i].var;
            q[k].colvar = y->linear[j].var;
            ++k;
        }
    AffineFunction* da = &dest->affine;
    double xconst = x->constant, yconst = y->constant;
    size_t linoffset = da->n;
    lin_resize(da, linoffset + x->n + y->n);
    k = linoffset;
    for (size_t i = 0; i < x->n; ++i) { da->linear[k].coeff = x->linear[i].coeff * yconst; da->linear[k].var = x->linear[i].var; ++k; }
    for (size_t i = 0; i < y->n; ++i) { da->linear[k].coeff = y->linear[i].coeff * xconst; da->linear[k].var = y->linear[i].var; ++k; }
    da->constant += xconst * yconst;
}

/* _vecdot!(dest::QuadraticFunction, x, y)  functions.jl:702-709 */
static void vecdot_quad(QuadraticFunction* dest, const AffineFunction* x, const AffineFunction* y, int64_t rows) {
    dest->n = 0; dest->affine.n = 0; dest->affine.constant = 0.0;   /* zero! :355 */
    for (int64_t i = 0; i < rows; ++i) muladd_quad(dest, &x[i], &y[i]);
}

/* update!(moi_f::MOI.ScalarQuadraticFunction, f, varmap)  moi_interop.jl:45-62 */
static void moi_update_quadratic(tl_model* m, const QuadraticFunction* f) {
    m->moi_const = f->affine.constant;
    if (f->affine.n > m->moi_aff_cap) { m->moi_aff = (MoiScalarAffineTerm*)realloc(m->moi_aff, f->affine.n * sizeof *m->moi_aff); m->moi_aff_cap = f->affine.n; }
    m->moi_naff = f->affine.n;
    for (size_t i = 0; i < f->affine.n; ++i) {
        m->moi_aff[i].coeff = f->affine.linear[i].coeff;
        m->moi_aff[i].var = m->varmap[f->affine.linear[i].var - 1];
    }
    if (f->n > m->moi_quad_cap) { m->moi_quad = (MoiScalarQuadraticTerm*)realloc(m->moi_quad, f->n * sizeof *m->moi_quad); m->moi_quad_cap = f->n; }
    m->moi_nquad = f->n;
    for (size_t i = 0; i < f->n; ++i) {
        const QuadraticTerm* t = &f->quadratic[i];
        double coeff = t->rowvar == t->colvar ? 2 * t->coeff : t->coeff;    /* :58 */
        m->moi_quad[i].coeff = coeff;
        m->moi_quad[i].var1 = m->varmap[t->rowvar - 1];
        m->moi_quad[i].var2 = m->varmap[t->colvar - 1];
    }
}

/* update!(moi_f::MOI.VectorAffineFunction, fs, varmap)  moi_interop.jl:64-81 */
static void moi_update_vector_affine(tl_model* m, const AffineFunction* fs, int64_t rows) {
    size_t num = 0;
    for (int64_t r = 0; r < rows; ++r) num += fs[r].n;
    if (num > m->moi_v_cap) { m->moi_vterms = (MoiVectorAffineTerm*)realloc(m->moi_vterms, num * sizeof *m->moi_vterms); m->moi_v_cap = num; }
    m->moi_nv = num;
    size_t i = 0;
    for (int64_t row = 0; row < rows; ++row) {
        const AffineFunction* f = &fs[row];
        for (size_t t = 0; t < f->n; ++t) {
            m->moi_vterms[i].output_index = row + 1;
            m->moi_vterms[i].coeff = f->linear[t].coeff;
            m->moi_vterms[i].var = m->varmap[f->linear[t].var - 1];
            ++i;
        }
        m->moi_vconst[row] = f->constant;
    }
}

tl_model* tl_create(int64_t n, int64_t k, int64_t mrows, const int64_t* varmap) {
    tl_model* m = (tl_model*)calloc(1, sizeof(tl_model));
    m->n = n; m->k = k; m->m = mrows;
    m->x = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    m->varmap = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) { m->x[i] = i + 1; m->varmap[i] = varmap ? varmap[i] : i + 1; }
    m->Ax = (AffineFunction*)calloc((size_t)k, sizeof(AffineFunction));
    m->res = (AffineFunction*)calloc((size_t)k, sizeof(AffineFunction));
    m->Cx = (AffineFunction*)calloc((size_t)mrows, sizeof(AffineFunction));
    m->con = (AffineFunction*)calloc((size_t)mrows, sizeof(AffineFunction));
    m->moi_vconst = (double*)calloc((size_t)(mrows > 0 ? mrows : 1), sizeof(double));
    return m;
}

void tl_destroy(tl_model* m) {
    if (!m) return;
    for (int64_t r = 0; r < m->k; ++r) { free(m->Ax[r].linear); free(m->res[r].linear); }
    for (int64_t r = 0; r < m->m; ++r) { free(m->Cx[r].linear); free(m->con[r].linear); }
    free(m->Ax); free(m->res); free(m->Cx); free(m->con);
    free(m->obj.quadratic); free(m->obj.affine.linear);
    free(m->moi_quad); free(m->moi_aff); free(m->moi_vterms); free(m->moi_vconst);
    free(m->x); free(m->varmap); free(m);
}

/* One update!(model): objective first, then the constraint (model.jl:134-142).  A: k x n, C: m x n, column-major.
 * parts: bit 0 = objective, bit 1 = constraint. */
void tl_update(tl_model* m, const double* A, const double* b, const double* C, const double* d, int64_t row_limit, int parts) {
    if (row_limit <= 0 || row_limit > m->k) row_limit = m->k;
    if (parts & 1) {
        /* residual . residual: the residual expression is evaluated once per operand (no memoisation) */
        for (int pass = 0; pass < 2; ++pass) {
            matvecmul(m->Ax, A, m->k, m->n, m->x);
            vecsubtract(m->res, m->Ax, b, m->k);
        }
        vecdot_quad(&m->obj, m->res, m->res, row_limit);
        moi_update_quadratic(m, &m->obj);
    }
    if (parts & 2) {
        matvecmul(m->Cx, C, m->m, m->n, m->x);
        vecsubtract(m->con, m->Cx, d, m->m);
        moi_update_vector_affine(m, m->con, m->m);
    }
}

/* accessors for the tests */
size_t tl_num_quadratic_terms(const tl_model* m) { return m->moi_nquad; }
size_t tl_num_affine_terms(const tl_model* m) { return m->moi_naff; }
size_t tl_num_vector_terms(const tl_model* m) { return m->moi_nv; }
double tl_objective_constant(const tl_model* m) { return m->moi_const; }
void tl_get_quadratic_terms(const tl_model* m, double* coeff, int64_t* v1, int64_t* v2) {
    for (size_t i = 0; i < m->moi_nquad; ++i) { coeff[i] = m->moi_quad[i].coeff; v1[i] = m->moi_quad[i].var1; v2[i] = m->moi_quad[i].var2; }
}
void tl_get_affine_terms(const tl_model* m, double* coeff, int64_t* v) {
    for (size_t i = 0; i < m->moi_naff; ++i) { coeff[i] = m->moi_aff[i].coeff; v[i] = m->moi_aff[i].var; }
}
void tl_get_vector_terms(const tl_model* m, int64_t* row, double* coeff, int64_t* v, double* constants) {
    for (size_t i = 0; i < m->moi_nv; ++i) { row[i] = m->moi_vterms[i].output_index; coeff[i] = m->moi_vterms[i].coeff; v[i] = m->moi_vterms[i].var; }
    for (int64_t r = 0; r < m->m; ++r) constants[r] = m->moi_vconst[r];
}
