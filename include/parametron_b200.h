/*
 * parametron_b200.h -- C ABI of the B200-native replacement for Parametron.jl's
 * parameter-update hot path (`update!` + coefficient assembly).
 *
 * The reference (pure Julia) has no FFI of its own; the seams this library sits
 * behind are listed per entry point as `ref:` citations into /root/reference/src.
 * A Julia `ccall` shim that binds exactly these symbols is shown in INTEGRATION.md;
 * the Python ctypes binding in parametron.jl_b200/_capi.py is the one exercised by tests.
 *
 * Conventions
 *   - every call returns PMB_OK (0) or a negative PMB_E* status; pmb_last_error(ctx) holds the
 *     message (the Julia shim rethrows: `error(...)`, DimensionMismatch, ArgumentError).
 *   - matrices are COLUMN-MAJOR fp64 (Julia `Matrix{Float64}`); variable indices are 1-based
 *     int64 (ref: functions.jl:96-98, model.jl:49-53).
 *   - a context owns one CUDA stream's worth of work and is NOT re-entrant, like the reference's
 *     preallocated `dest` buffers (ref: lazyexpression.jl:202, test/lazyexpression.jl:73).
 *   - `batch` > 1 stacks `batch` independent instances of the same model: every Parameter buffer
 *     gets a leading instance stride, every output too (instance-shard data parallelism).
 *   - build-phase calls after pmb_initialize() fail with PMB_EINITIALIZED
 *     (ref: model.jl:50,61,69 "Model has already been initialized").
 *   - pmb_update() performs no allocation and, for device-resident Parameters, no host<->device
 *     copy (ref: `@allocated solve!(model) == 0`, test/model.jl:117-122).
 */
#ifndef PARAMETRON_B200_H
#define PARAMETRON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmb_ctx pmb_ctx;
typedef int32_t pmb_node; /* handle of a DAG node (>= 0) */

enum {
    PMB_OK = 0,
    PMB_EINVAL = -1,       /* ArgumentError */
    PMB_EDIM = -2,         /* DimensionMismatch (ref: functions.jl:669,755,780-781,845,877) */
    PMB_EINITIALIZED = -3, /* build call after initialize (ref: model.jl:50,61,69) */
    PMB_ENOTINIT = -4,
    PMB_ECUDA = -5,
    PMB_EUNSUPPORTED = -6, /* expression shape with no device lowering ("Unhandled ...", model.jl:139) */
    PMB_ENOMEM = -7
};

/* value kinds of DAG nodes == Julia types of evaluated expressions */
enum {
    PMB_KIND_NUM_SCALAR = 0, /* T                          */
    PMB_KIND_NUM_VECTOR = 1, /* Vector{T}                  */
    PMB_KIND_NUM_MATRIX = 2, /* Matrix{T}                  */
    PMB_KIND_VARVEC = 3,     /* Vector{Variable}           */
    PMB_KIND_AFFVEC = 4,     /* Vector{AffineFunction{T}}  */
    PMB_KIND_AFF = 5,        /* AffineFunction{T}          */
    PMB_KIND_QUAD = 6,       /* QuadraticFunction{T}       */
    PMB_KIND_ADJVEC = 7      /* Adjoint/Transpose of a vector (numeric or Variables) */
};

enum { PMB_MINIMIZE = 0, PMB_MAXIMIZE = 1 }; /* ref: Parametron.jl:50 `Sense` */

/* constraint sets; scalar kinds for AFF/QUAD nodes, vector kinds for AFFVEC nodes
 * (ref: model.jl:75-95, bucket order moi_interop.jl:180-193) */
enum {
    PMB_SET_NONNEGATIVE = 0, /* expr >= 0 : GreaterThan(0) / Nonnegatives(n) */
    PMB_SET_NONPOSITIVE = 1, /* expr <= 0 : LessThan(0)    / Nonpositives(n) */
    PMB_SET_ZERO = 2         /* expr == 0 : EqualTo(0)     / Zeros(n)        */
};

/* Parameter modes (ref: parameter.jl:48,57,88) */
enum {
    PMB_PARAM_DEVICE = 0,  /* manual mode: caller-owned DEVICE buffer, written by the caller before update
                              (ref: Parameter(model; val=buf), parameter.jl:88) */
    PMB_PARAM_HOST = 1,    /* caller-owned HOST buffer (+ optional host callback run once per update, then one
                              H2D copy) (ref: in-place Parameter(f, val, model), parameter.jl:57,101) */
    PMB_PARAM_UNIFORM = 2  /* built-in device generator: scale * U[0,1) from Philox4x32-10 keyed by
                              (seed, GLOBAL instance id = pmb_set_instance_offset + local instance, update counter) (ref: Parameter(rand!, zeros(n,n), model), README.md:36-43) */
};

typedef void (*pmb_host_callback)(double* val, int64_t rows, int64_t cols, int64_t batch, void* user);

/* ---- context ------------------------------------------------------------------------------- */
/* ref: Model(optimizer) model.jl:10-30.  `stream` is a cudaStream_t (NULL = the library creates its own). */
int pmb_ctx_create(int device, void* stream, int64_t batch, pmb_ctx** out);
int pmb_ctx_destroy(pmb_ctx* ctx);
/* Instance-shard data parallelism: this context owns the contiguous shard [first, first + batch) of a larger
 * batch of independent instances of the model (ref: the same update! loop, model.jl:132-143, run once per instance).  `first` is the
 * GLOBAL id of local instance 0: the built-in generators (PMB_PARAM_UNIFORM) key Philox by the global id, so the shard reproduces
 * bit for bit the matching slice of the unsharded batch.  Build phase only; default 0. */
int pmb_set_instance_offset(pmb_ctx* ctx, int64_t first);
int pmb_get_instance_offset(const pmb_ctx* ctx, int64_t* first);
const char* pmb_last_error(const pmb_ctx* ctx);
const char* pmb_version(void);

/* ---- variables / parameters / constants (build phase) ------------------------------------------ */
/* ref: Variable(m::Model) model.jl:49-53 -- returns the 1-based index of the first new variable. */
int pmb_add_variables(pmb_ctx* ctx, int64_t n, int64_t* first_index);
int pmb_num_variables(const pmb_ctx* ctx, int64_t* n);

/* ref: Parameter constructors parameter.jl:48-88.  fp64 value of `ndims` dimensions: 0 = scalar (rows, cols ignored),
 * 1 = vector of length rows (may be 0), 2 = rows x cols column-major matrix. */
int pmb_param_create(pmb_ctx* ctx, int ndims, int64_t rows, int64_t cols, int mode, pmb_node* out);
/* Block-sparse Parameter matrix (BSR): rows x cols with `nblocks` stored b x b blocks at 0-based block coordinates
 * (block_rows[i], block_cols[i]); the value buffer is nblocks * b * b doubles, each block column-major.  The reference has no
 * sparse path (matvecmul! visits every A[i], functions.jl:790-796); this is its dense matrix with the structural zeros left
 * out.  Supported use: `A * x` with x::Vector{Variable} as the only coefficient source of a constraint block, which is then
 * emitted in sparse form (pmb_block_view.is_sparse). */
int pmb_param_create_bsr(pmb_ctx* ctx, int64_t rows, int64_t cols, int64_t b, int64_t nblocks, const int64_t* block_rows,
                         const int64_t* block_cols, int mode, pmb_node* out);
int pmb_param_set_device_ptr(pmb_ctx* ctx, pmb_node p, const double* dptr);           /* PMB_PARAM_DEVICE */
int pmb_param_set_host_ptr(pmb_ctx* ctx, pmb_node p, double* hptr,
                           pmb_host_callback f, void* user);                          /* PMB_PARAM_HOST */
int pmb_param_set_uniform(pmb_ctx* ctx, pmb_node p, uint64_t seed, double scale);     /* PMB_PARAM_UNIFORM */
/* stream-ordered host -> device copy of a whole Parameter value (batch * rows * max(cols,1) doubles); usable after
 * pmb_initialize for PMB_PARAM_DEVICE parameters whose buffer the library owns (no pmb_param_set_device_ptr call). */
int pmb_param_upload(pmb_ctx* ctx, pmb_node p, const double* host_values);
/* device address of the Parameter's current value (batch-major: instance stride = rows*max(cols,1)); valid in stream order on the context
 * stream.  A PMB_PARAM_UNIFORM matrix whose only reader is the int8 slice pass of the objective is generated INSIDE that pass and not written
 * to memory by pmb_update (its values are a pure function of seed, global instance id and update counter); this call materialises the values of
 * the LAST update into the buffer before it returns the address -- call it again after later updates instead of keeping the pointer
 * (PMB_GEN_MATERIALIZE=1 in the environment makes every update write the buffer, as for all other Parameters). */
int pmb_param_device_ptr(pmb_ctx* ctx, pmb_node p, double** dptr);

/* constants folded at build time (ref: optimize_toplevel constant folding, lazyexpression.jl:189-192) */
int pmb_const_numeric(pmb_ctx* ctx, int ndims, int64_t rows, int64_t cols, const double* host_values, pmb_node* out);
/* Vector{Variable} from 1-based indices */
int pmb_varvec(pmb_ctx* ctx, int64_t n, const int64_t* var_indices, pmb_node* out);
/* constant symbolic values (terms given in reference order): scalar affine / quadratic, affine vector */
int pmb_const_affine(pmb_ctx* ctx, int64_t nterms, const double* coeffs, const int64_t* vars, double constant, pmb_node* out);
int pmb_const_quadratic(pmb_ctx* ctx, int64_t nquad, const double* qcoeffs, const int64_t* rowvars, const int64_t* colvars,
                        int64_t nlin, const double* lcoeffs, const int64_t* lvars, double constant, pmb_node* out);

/* ---- expression DAG: one builder per rewrite rule of lazyexpression.jl:198-302 ---------------- */
int pmb_op_matvec(pmb_ctx* ctx, pmb_node A, pmb_node x, pmb_node* out);   /* A*x   -> matvecmul!  :200-204, functions.jl:775-829 */
int pmb_op_adjoint(pmb_ctx* ctx, pmb_node A, pmb_node* out);              /* A' / transpose :206-217 */
int pmb_op_bilinear(pmb_ctx* ctx, pmb_node xT, pmb_node Q, pmb_node y, pmb_node* out); /* x'Qy -> bilinearmul! :219-226, functions.jl:840-858 */
int pmb_op_dot(pmb_ctx* ctx, pmb_node x, pmb_node y, pmb_node* out);      /* x . y -> vecdot!     :228-232, functions.jl:665-731 */
int pmb_op_add(pmb_ctx* ctx, pmb_node x, pmb_node y, pmb_node* out);      /* x + y -> add!/vecadd!           :238-247 */
int pmb_op_sub(pmb_ctx* ctx, pmb_node x, pmb_node y, pmb_node* out);      /* x - y -> subtract!/vecsubtract! :249-258 */
int pmb_op_mul(pmb_ctx* ctx, pmb_node x, pmb_node y, pmb_node* out);      /* scalar*fn -> mul! :260-274; number*vector -> scale! :284-290 */
int pmb_op_vcat(pmb_ctx* ctx, int32_t n, const pmb_node* parts, pmb_node* out); /* vcat! :276-278, functions.jl:969-994 */
int pmb_op_vect(pmb_ctx* ctx, pmb_node x, pmb_node* out);                 /* [x]   :292-298 */
int pmb_op_convert(pmb_ctx* ctx, pmb_node x, pmb_node* out);              /* convert(Vector, x) -> copyto! :280-282 */
/* p.field -> GetField :300-302, functions.jl:998-999.  A struct-valued Parameter is a flat record (fields packed in declaration
 * order, matrices column-major); the field is the ndims/rows/cols sub-array starting at element `offset` of the record. */
int pmb_op_getfield(pmb_ctx* ctx, pmb_node p, int64_t offset, int ndims, int64_t rows, int64_t cols, pmb_node* out);
int pmb_node_info(pmb_ctx* ctx, pmb_node n, int32_t* kind, int64_t* rows, int64_t* cols, int32_t* is_constant);

/* ---- model ------------------------------------------------------------------------------------ */
/* ref: setobjective! model.jl:60-66 / @objective :267-271.  node kind: AFF, QUAD (or NUM_SCALAR). */
int pmb_set_objective(pmb_ctx* ctx, int sense, pmb_node expr);
/* ref: add_*_constraint! model.jl:75-98 / @constraint :224-249 (expr = lhs - rhs).  Returns the block id. */
int pmb_add_constraint(pmb_ctx* ctx, pmb_node expr, int set_kind, int32_t* block_id);
/* ref: initialize! model.jl:117-122 + mapindices! :100-107.  varmap[i-1] = optimizer index of model variable i
 * (NULL = identity); plans launches, allocates all device buffers, captures the CUDA graph. */
int pmb_initialize(pmb_ctx* ctx, const int64_t* varmap);
/* ref: update!(m::Model) model.jl:132-143: mark all Parameters dirty, evaluate each once, re-assemble the objective
 * (Q, q, c0) then every non-constant constraint block in bucket order (moi_interop.jl:180-193, 236-247).
 * Asynchronous on the context's stream. */
int pmb_update(pmb_ctx* ctx);
int pmb_synchronize(pmb_ctx* ctx);
int pmb_update_count(const pmb_ctx* ctx, int64_t* n);

/* ---- device-resident handoff (replaces MOI.set, moi_interop.jl:134,171) ------------------------- */
typedef struct {
    int32_t kind;        /* PMB_KIND_AFF or PMB_KIND_QUAD */
    int32_t sense;
    int32_t is_constant; /* skipped by update (ref: isconstant, moi_interop.jl:123,132) */
    int64_t nvars;       /* N: Q is N x N column-major (ld = N), symmetric = MOI Hessian of 1/2 x'Qx + q'x + c0 */
    double* Q;           /* device, batch stride N*N; NULL for affine objectives */
    double* q;           /* device, batch stride N */
    double* c0;          /* device, batch stride 1 */
} pmb_objective_view;

typedef struct {
    int32_t set_kind;    /* PMB_SET_* */
    int32_t is_scalar;   /* 1: scalar (Aff/Quad in GreaterThan/LessThan/EqualTo), 0: vector set */
    int32_t is_constant;
    int32_t bucket;      /* index into the reference's fixed bucket order (moi_interop.jl:180-193) */
    int64_t rows;        /* m_i */
    int64_t nvars;       /* N */
    double* C;           /* device, m_i x N column-major (ld = m_i), batch stride m_i*N; columns are OPTIMIZER indices.  READ-ONLY for the caller:
                          * when the block is `M * x` over all variables in optimizer order with unit coefficient and M is a Parameter the library
                          * fills only inside pmb_update (PMB_PARAM_UNIFORM / PMB_PARAM_HOST), this IS the Parameter's buffer (no copy is made;
                          * PMB_NO_BLOCK_ALIAS=1 in the environment restores the copy) */
    double* d;           /* device, m_i, batch stride m_i   (the block is  C x + d  in the set) */
    int64_t nterms;      /* number of reference LinearTerms per instance (static sparsity bookkeeping) */
    /* sparse form (block built from one BSR matrix): C is NULL; coefficient e of an instance is vals[e] at
     * (sp_rows[e], sp_cols[e]) (0-based row, 0-based OPTIMIZER column), e in value-buffer order; the index tables are static */
    int32_t is_sparse;
    int64_t nnz;
    double* vals;            /* device, batch stride nnz */
    const int32_t* sp_rows;  /* device, [nnz] */
    const int32_t* sp_cols;  /* device, [nnz] */
    /* scalar quadratic constraint (bucket 6-8: 1/2 x'Qx + C x + d in GreaterThan/LessThan/EqualTo(0)): Q is the symmetric MOI
     * Hessian (N x N column-major, batch stride N*N), C the 1 x N linear part, d the constant; NULL for affine blocks */
    double* Q;
} pmb_block_view;

int pmb_get_objective(pmb_ctx* ctx, pmb_objective_view* out);
int pmb_num_blocks(const pmb_ctx* ctx, int32_t* n);
/* blocks are numbered in UPDATE order: bucket order, insertion order within a bucket */
int pmb_get_block(pmb_ctx* ctx, int32_t update_pos, pmb_block_view* out);
/* Static index bookkeeping, bit-exact with the reference's MOI.VectorAffineTerm list after `update!`
 * (row-outer / term-inner, varmap applied; moi_interop.jl:64-81).  `src` gives, for term t, the offset
 * (column-major, within one instance) of the dense C entry its coefficient is summed into. */
int pmb_block_term_index(pmb_ctx* ctx, int32_t update_pos, int64_t capacity, int64_t* rows, int64_t* vars, int64_t* src);

/* ---- MOI-compatible term-list export -------------------------------------------------------
 * One record per reference term of a vector (or scalar) affine constraint block, in EXACTLY the order and layout of
 * update!(::MOI.VectorAffineFunction, ...) (moi_interop.jl:64-81): row-outer / term-inner, 1-based output index, variable =
 * optimizer index (varmap applied), duplicates NOT combined (a variable used twice yields two records, as in the reference). */
typedef struct {
    int64_t output_index;  /* MOI.VectorAffineTerm.output_index (1-based row) */
    double coefficient;    /* MOI.ScalarAffineTerm.coefficient */
    int64_t variable;      /* MOI.ScalarAffineTerm.variable_index.value (1-based) */
} pmb_vector_affine_term;
/* Writes batch * nterms records to the DEVICE buffer `terms_dev` and (if non-NULL) batch * rows constants to `constants_dev`,
 * asynchronously on the context stream, from the Parameter values of the most recent pmb_update.  Rows with different term layouts
 * (a vcat of blocks of different shapes) go through a per-term gather table instead of the per-row template.  PMB_EUNSUPPORTED for
 * sparse blocks (the reference would emit the structural zeros). */
int pmb_block_export_terms(pmb_ctx* ctx, int32_t update_pos, pmb_vector_affine_term* terms_dev, double* constants_dev);

/* The objective as the MOI function update!(::MOI.ScalarQuadraticFunction, f, varmap) fills (moi_interop.jl:45-62), in COMBINED form: the
 * reference appends one ScalarQuadraticTerm per product (k n^2 of them for ||Ax - b||^2, duplicates left to the solver, coefficient x 2 when
 * both variables coincide); the device holds the sum, so it writes one term per unordered pair of OPTIMIZER variables i <= j (row-outer:
 * (1,1), (1,2), .., (1,N), (2,2), ..), coefficient = entry (i, j) of the symmetric MOI Hessian of 1/2 x'Qx -- the value MOI's
 * duplicate-summing gives the reference's list -- plus one ScalarAffineTerm per variable and the constant.  quad_dev: batch * N (N + 1) / 2
 * records (NULL: skip; error for an affine objective), aff_dev: batch * N records (NULL: skip), constant_dev: batch doubles (NULL: skip). */
typedef struct {
    double coefficient;    /* MOI.ScalarQuadraticTerm.coefficient */
    int64_t variable_1;    /* .variable_index_1.value (1-based) */
    int64_t variable_2;    /* .variable_index_2.value */
} pmb_scalar_quadratic_term;
typedef struct {
    double coefficient;    /* MOI.ScalarAffineTerm.coefficient */
    int64_t variable;      /* .variable_index.value (1-based) */
} pmb_scalar_affine_term;
int pmb_objective_export_terms(pmb_ctx* ctx, pmb_scalar_quadratic_term* quad_dev, pmb_scalar_affine_term* aff_dev, double* constant_dev);

/* ---- solver-side ingest: the QP in the form an OSQP-style solver consumes ---------------------------
 *     minimize 1/2 x'Px + q'x + c0   subject to   l <= A x <= u
 * P_x: values of the UPPER triangle of the MOI Hessian in CSC order (column j holds rows 0..j); A: all constraint blocks stacked
 * in update order, CSC; bounds from the set kinds (block  Cx + d in set):  Zeros l = u = -d;  Nonnegatives l = -d, u = +inf;
 * Nonpositives l = -inf, u = -d.  The patterns are static (written once by pmb_initialize); every pmb_update refreshes the
 * VALUES only (value-only updates into a fixed pattern, replacing MOI.set at moi_interop.jl:134,171). */
typedef struct {
    int64_t n, m;            /* variables, total constraint rows */
    int64_t nnzP, nnzA;
    double* P_x;             /* device, batch stride nnzP (NULL for affine objectives) */
    double* q;               /* device, batch stride n  (same buffer as pmb_objective_view.q) */
    double* c0;              /* device, batch stride 1 */
    double* A_x;             /* device, batch stride nnzA */
    double* l;               /* device, batch stride m */
    double* u;               /* device, batch stride m */
    const int64_t* P_p;      /* device [n + 1] */
    const int64_t* P_i;      /* device [nnzP] */
    const int64_t* A_p;      /* device [n + 1] */
    const int64_t* A_i;      /* device [nnzA] */
} pmb_qp_view;
int pmb_enable_qp_pack(pmb_ctx* ctx);             /* build phase: adds the pack launches to every update */
int pmb_get_qp(pmb_ctx* ctx, pmb_qp_view* out);   /* after pmb_initialize */
/* After pmb_enable_qp_pack (before or after pmb_initialize -- the sizes come from pmb_get_qp): pinned HOST buffers (pmb_host_alloc; sizes batch * nnzP / n / 1 / nnzA / m / m doubles, any
 * may be NULL) that every pmb_update fills as part of the update graph: the constraint pieces (A_x, l, u) leave the device as soon
 * as the constraint chain is done, concurrently with the objective chain; (P_x, q, c0) follow it.  Device-resident results (Q, q, c0, the
 * packed P_x on the device) are valid in stream order after every update.  The HOST sink is valid after pmb_synchronize (or any call that
 * returns results to the host: pmb_value, pmb_objective_value, pmb_gram_level_counts): when the objective's operand is a pinned HOST
 * Parameter on the int8 route, the update is a pipelined schedule (chunked upload || per-chunk slice + Gram || streamed P_x download) in which
 * P_x leaves range by range BEFORE the error certificate's verdict is final; if the verdict ends at fp64, the in-stream fp64 launch recomputes
 * Q and P_x on the device and those calls re-download P_x into the sink before they return (a bare cudaStreamSynchronize would leave the
 * host copy of P_x -- and only that -- from the refused int8 attempt). */
int pmb_qp_set_host_sink(pmb_ctx* ctx, double* P_x, double* q, double* c0, double* A_x, double* l, double* u);

/* ---- host staging helpers (pinned memory + stream-ordered copies; the e2e path of bench.py and the ctypes tests) -- */
int pmb_host_alloc(pmb_ctx* ctx, size_t bytes, void** out);   /* cudaHostAlloc; freed by pmb_host_free or pmb_ctx_destroy */
int pmb_host_free(pmb_ctx* ctx, void* p);
int pmb_memcpy_d2h(pmb_ctx* ctx, void* host, const void* dev, size_t bytes);  /* async on the context stream */
int pmb_memcpy_h2d(pmb_ctx* ctx, void* dev, const void* host, size_t bytes);  /* async on the context stream */
/* scratch DEVICE memory on the context's device (the export calls write into caller-provided device buffers): cudaSetDevice(ctx device) +
 * cudaMalloc, tracked by the context (freed by pmb_device_free or pmb_ctx_destroy) */
int pmb_device_alloc(pmb_ctx* ctx, size_t bytes, void** out);
int pmb_device_free(pmb_ctx* ctx, void* p);

/* ---- result path: value.(model, x) / objectivevalue(model), model.jl:166-173 ----------------------------
 * The external solver hands back its primal solution in OPTIMIZER variable order (MOI.VariablePrimal per optimizer index).
 * pmb_set_primal stores it on the device (x: [batch][N]; on_device != 0: x is a device pointer, copied device-to-device on the context
 * stream -- a device-resident solver never leaves the GPU).  pmb_value gathers model Variables through model_var_to_optimizer
 * (value(m, x) = VariablePrimal[m.model_var_to_optimizer[x.index]]): out is [batch][nv], host (out_on_device == 0, synchronous) or
 * device memory.  pmb_objective_value evaluates the objective the last update handed over, 1/2 x'Qx + q'x + c0 (q'x + c0 for an
 * affine objective), at the stored solution on the device: out_host[batch]. */
int pmb_set_primal(pmb_ctx* ctx, const double* x, int on_device);
int pmb_value(pmb_ctx* ctx, int64_t nv, const int64_t* var_indices, double* out, int out_on_device);
int pmb_objective_value(pmb_ctx* ctx, double* out_host);

/* ---- introspection for tests / bench --------------------------------------------------------- */
typedef struct {
    int64_t kernel_launches_per_update; /* launches of THIS library's kernels in one update */
    int64_t memcpy_nodes_per_update;    /* H2D staging copies (PMB_PARAM_HOST only) */
    int64_t device_bytes_allocated;
    int64_t uses_cuda_graph;
    int64_t affine_algorithmic_bytes;   /* 16 B per stored coefficient + 16 B per row, all blocks, one instance */
    int64_t gram_algorithmic_flops;     /* SYRK count k*N*(N+1) + 2kN + 2k per Gram term, one instance */
} pmb_stats;
int pmb_get_stats(pmb_ctx* ctx, pmb_stats* out);
/* Precision mode of the objective-Hessian contraction (build phase).  The reference accumulates every product in Float64
 * (functions.jl:548-576); the north_star tolerance is 1e-10 relative.  mode 1 (default): contractions large enough to amortise it run
 * as exact integer products of int8 digit planes on the tcgen05 tensor cores whenever a worst-case error bound computed from the
 * operand (tol, default 9e-11, relative to the largest |Q| entry) certifies the result -- per instance, 5 or 6 planes; everything
 * else runs on the fp64 DMMA pipe.  mode 0: fp64 only.  modes 5 / 6: force that many planes (experiments, no certificate). */
int pmb_set_gram_mode(pmb_ctx* ctx, int mode, double tol);
/* instances of the most recent update whose contraction ran with 5 planes / 6 planes / in fp64: counts[3] */
int pmb_gram_level_counts(pmb_ctx* ctx, int64_t* counts);
/* time (ms, CUDA events on the context stream) of the most recent launch of each kernel class; class names via pmb_kernel_class_name */
int pmb_enable_kernel_timing(pmb_ctx* ctx, int enable);
int pmb_kernel_class_count(void);
const char* pmb_kernel_class_name(int cls);
int pmb_kernel_time_ms(pmb_ctx* ctx, int cls, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* PARAMETRON_B200_H */
