// Device-side argument structs and host launchers of the hand-written sm_100a kernels.
// Host code (pmb_api.cu) only sees these launchers; every launcher enqueues on `stream` and never allocates.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmb {

// alpha = c * (ptr ? ptr[instance * bstride] : 1)     -- dynamic scalar multipliers (scale!/mul!, functions.jl:578,873-925)
struct AlphaDev {
    double c;
    const double* ptr;
    int64_t bstride;
};

// ---------------- K1: batched affine-coefficient assembly ------------------------------------
// One coefficient source of a block:  G[row_off + i, outcol] += alpha * src[i + srccol * ld]   (matvecmul!, functions.jl:775-798)
struct AffMatTermDev {
    const double* src;
    int64_t src_bstride;  // elements between instances (0: shared constant)
    int64_t ld;
    int64_t row_off, rows;
    AlphaDev alpha;
};
// identity-like source: G[row, outcol] += alpha  for the single row stored in the column entry (Vector{Variable} operands)
struct AffUnitTermDev {
    AlphaDev alpha;
};
struct AffColEntry {  // one contribution to one output column
    int32_t term;     // >= 0: index into mats[];  < 0: ~term indexes units[]
    int32_t src;      // mats: source column;  units: the (block) row that receives alpha
};
// offset source: d[row_off + i] += alpha * v[src_off + i]        (subtract!/add! of numbers, functions.jl:452,474)
struct AffOffTermDev {
    const double* src;
    int64_t src_bstride;
    int64_t src_off, row_off, rows;
    AlphaDev alpha;
};
struct AffSeg {        // one contiguous piece: dst[dst_off ...] = alpha * src[src_off ...], `bytes` long (multiple of 16)
    int64_t src_off;   // elements from the source matrix base (within one instance)
    int64_t dst_off;   // elements from the block's C base (within one instance)
    int32_t bytes;
    int32_t stage_off; // byte offset inside the chunk's shared-memory stage
};
constexpr int kAffStageBytes = 16384;
struct AffBlockDev {
    double* C;            // rows x N column-major, ld = rows
    double* d;            // rows
    int64_t C_bstride, d_bstride;
    int64_t rows;
    int64_t d_rows;       // length of d (== rows except for sparse blocks, whose "rows" is the number of stored values)
    int32_t ntouched;     // number of output columns with at least one contribution (the others stay 0 forever)
    const int32_t* col_out;   // [ntouched] output column (0-based optimizer variable index)
    const int32_t* col_ptr;   // [ntouched + 1] CSR pointer into entries
    const AffColEntry* entries;
    const AffMatTermDev* mats;
    const AffUnitTermDev* units;
    int32_t noffs;
    const AffOffTermDev* offs;
    const double* d_const;    // [rows] folded constant part of the offsets (may be null)
    int32_t simple;           // every touched column has exactly one matrix entry from mats[0] covering all rows
    const int32_t* col_src;   // [ntouched] source column (simple path)
    int32_t vec2;             // 16-byte path legal (even rows/ld/strides, aligned bases)
    // TMA bulk-copy plan of the simple path (static, built once): the block is a list of contiguous segments
    // (merged runs of adjacent columns), grouped into chunks of at most one shared-memory stage.
    int32_t tma;              // plan present (requires vec2)
    int32_t nchunks;
    const AffSeg* segs;
    const int32_t* chunk_first;   // [nchunks + 1] first segment of each chunk
    // regular plan (one segment per chunk, offsets in arithmetic progression -- e.g. the identity variable map): the kernel
    // computes the descriptors instead of reading the tables
    int32_t uniform;
    int64_t u_src0, u_dst0, u_src_stride, u_dst_stride;
    int32_t u_bytes, u_last_bytes;
};
void launch_affine_assemble(const AffBlockDev& blk, int64_t batch, int sm_count, cudaStream_t stream);
void launch_affine_assemble_tma(const AffBlockDev& blk, int64_t batch, int sm_count, cudaStream_t stream);  // blk.tma != 0
void configure_affine_tma_kernels();

// ---------------- K2: fp64 Gram (SYRK-shaped) on the DMMA tensor pipe ---------------------------
// H[map[i], map[j]] (=|+=) scale * sum_r X[r, i] * X[r, j]   for i, j in [0, n), both triangles written
// (vecdot! + muladd!, functions.jl:548-576,702-709, followed by the diag x2 / symmetric MOI Hessian of moi_interop.jl:45-62)
struct GramArgs {
    const double* X;       // k x n column-major
    int64_t ldx, x_bstride;
    int64_t k, n;
    double* H;             // N x N column-major
    int64_t ldh, h_bstride;
    const int32_t* map;    // [n] 0-based output index of source column i; null = identity
    AlphaDev scale;
    int32_t accumulate;    // 0: overwrite, 1: +=
    // Arrival gating (upload/compute overlap): X is being copied host->device in chunks of `gate_cols` columns while the kernel
    // runs; gate[c] becomes non-zero (written by a stream-ordered copy AFTER chunk c) when columns [c*gate_cols, (c+1)*gate_cols)
    // have landed.  Tiles are enumerated column-major, so early tiles only need early chunks.  null = no gating.
    const int32_t* gate;
    int32_t gate_cols;
    // Fused solver-side ingest: besides H, the epilogue writes the UPPER triangle in packed CSC order (column j holds rows 0..j,
    // P_x[j(j+1)/2 + i]) to a device buffer (packed_host: same, to mapped pinned host memory -- kept for experiments, measured
    // slower than a D2H copy).  Only legal when this launch is the only writer of H (accumulate == 0).  null = off.
    double* packed_dev;
    double* packed_host;
    int64_t packed_bstride;
    // Fused linear part (TMA kernel only): while a DIAGONAL tile's panel sits in shared memory the consumers also accumulate
    //   q[map[col]] += qscale * sum_k X[k, col] * gvec[k]      and (tile (0,0))   c0 += cscale * sum_k gvec[k]^2
    // which removes the second pass over X (the gemv + dot launch).  gvec == null: off.  Requires k even, gvec 16-byte aligned.
    const double* gvec; int64_t g_bstride;
    double* qout; int64_t q_bstride;
    double* c0out;
    AlphaDev qscale, cscale;
    // Certified int8-slice path (kernels_gram_i8.cu) ran first: skip the instances it certified.  null = compute everything.
    const unsigned long long* cert_stats;
    double cert_tol;
    int32_t cert_force;    // experiments: the int8 path was forced for every instance
    const unsigned long long* cert_none;   // fused small-instance route: device counter of the instances it left uncertified (0: the fp64 launch exits at once)
};
void launch_gram_syrk(const GramArgs& a, int64_t batch, int sm_count, cudaStream_t stream);
// general TN product  M (n x p, ld) = X' * Y  (scratch for dot(f, h) with f != h and for A * Vector{AffineFunction})
struct GemmTNArgs {
    const double* X; int64_t ldx, x_bstride;
    const double* Y; int64_t ldy, y_bstride;
    int64_t k, n, p;
    double* M; int64_t ldm, m_bstride;
};
void launch_gemm_tn(const GemmTNArgs& a, int64_t batch, int sm_count, cudaStream_t stream);
// TMA + warp-specialised variants (kernels_gram_tma.cu).  Return false when the operands do not meet the TMA alignment rules
// (the caller then uses the cp.async kernels above).
struct GramSplitPlan {          // k-split of the last partial wave of tiles (computed once by the planner)
    int32_t nslices;            // 1: no split
    int64_t split_tiles;        // number of tiles that are split (the tail of the tile enumeration)
    int64_t workspace_doubles;  // split_tiles * nslices * bm * bm
    int32_t bm;                 // tile edge: 128, or 64 for problems smaller than one wave of 128-tiles
};
GramSplitPlan plan_gram_split(bool syrk, int64_t k, int64_t n, int64_t p, int64_t batch, int sm_count);
struct GramScratch {
    GramSplitPlan plan;
    int32_t* counters;   // [split_tiles], zero between launches
    double* workspace;   // [workspace_doubles]
    double* qworkspace;  // [split_tiles * nslices * (bm + 1)] partial (q, c0) sums of the fused linear part
    int32_t qworkspace_ok;  // the planner provided what the fused linear part needs
};
bool launch_gram_syrk_tma(const GramArgs& a, const GramScratch& sc, int64_t batch, int sm_count, cudaStream_t stream);
bool launch_gemm_tn_tma(const GemmTNArgs& a, const GramScratch& sc, int64_t batch, int sm_count, cudaStream_t stream);
void configure_gram_tma_kernels();
bool gram_fused_linear_supported(int64_t k, int64_t n, int64_t batch, int sm_count);
bool gram_tma_applicable(const double* X, int64_t ld, int64_t bstride, int64_t k, int64_t n, int64_t batch);

// ---------------- K2i: certified int8-slice Gram on tcgen05 (kernels_gram_i8.cu) ----------------------------------------------
// Every column j of X is scaled by 1 / sigma_j (sigma_j = largest magnitude / 0.99) and rounded to a 47-bit fixed-point integer cut
// into kGramI8Planes balanced base-256 digits (int8 planes).  The Gram kernel multiplies the leading NS = 5 or 6 planes.
// Per-instance statistics written by the slice kernel (unsigned long long[kGramI8StatsStride]):
//   [0] bits of max_j sigma_j (0: every column is zero; +inf: non-finite or out-of-range data)   [1] bits of max_j ||x_j||_2^2
//   [2] bits of max_j ||x_j / sigma_j||_1                                                          [3 + s] max_j sum_r d_s[r, j]^2
constexpr int kGramI8Planes = 6;
constexpr int kGramI8StatsStride = 16;   // [14] columns sliced so far (arrival counter), [15] the certificate's verdict: 5 / 6 planes or kGramI8Fp64
constexpr int kGramI8Fp64 = 64;          // verdict: not certifiable, the fp64 kernel owns the instance (0 = no verdict yet)
#ifdef __CUDACC__
// Number of digit planes (5 or 6) whose worst-case error -- representation error of the planes used plus every dropped digit
// product, bounded with the measured column norms -- is below tol * max |H| (max |H| of a Gram matrix = its largest diagonal entry);
// kGramI8Fp64: not certifiable, the fp64 DMMA kernel computes this instance.  Evaluated ONCE per instance and update by the slice
// kernel (the group that slices the instance's last column) and stored in stats[15].
__device__ __forceinline__ int gram_i8_compute_level(const unsigned long long* st, double tol) {
    const unsigned long long s0 = __ldcg(st);
    if (s0 == 0ull) return 5;                       // all-zero operand: every digit is zero
    const double sig = __longlong_as_double((long long)s0);
    if (!(sig < 1e200)) return kGramI8Fp64;
    const double ss = __longlong_as_double((long long)__ldcg(st + 1)), l1 = __longlong_as_double((long long)__ldcg(st + 2));
    double R[kGramI8Planes];
#pragma unroll
    for (int s = 0; s < kGramI8Planes; ++s) R[s] = sqrt((double)__ldcg(st + 3 + s));
#pragma unroll
    for (int ns = 5; ns <= kGramI8Planes; ++ns) {
        double by = 2.0 * 1.1 * exp2(-8.0 * ns) * l1;  // sum_r (|y_ri| + |y_rj|) * |representation error|
#pragma unroll
        for (int s = 1; s < ns; ++s)
#pragma unroll
            for (int t = 1; t < ns; ++t)
                if (s + t >= ns) by += exp2(-14.0 - 8.0 * (s + t)) * R[s] * R[t];   // dropped digit products (Cauchy-Schwarz)
        if (1.01 * sig * sig * by <= tol * ss) return ns;
    }
    return kGramI8Fp64;
}
__device__ __forceinline__ int gram_i8_level(const unsigned long long* stats, int64_t inst) {
    return (int)stats[inst * kGramI8StatsStride + 15];
}
#endif
struct GramI8Plan {
    int32_t ok;            // path applicable
    int64_t kp;            // k rounded up to 128
    int64_t slice_bytes;   // batch * kGramI8Planes * n * kp
};
GramI8Plan plan_gram_i8(int64_t k, int64_t n, int64_t batch);
struct GramI8Scratch {
    GramI8Plan plan;
    int8_t* slices;              // [batch][kGramI8Planes][n][kp]
    double* sigma;               // [batch][n]
    unsigned long long* stats;   // [batch][kGramI8StatsStride]
    double tol;                  // certification tolerance (relative to max |H|)
    int32_t force_ns;            // experiments: 5 / 6 = run that many planes whatever the certificate says, 0 = adaptive
    int32_t fused;               // the fused small-instance kernel (kernels_gram_i8_fused.cu) runs instead of slice + Gram launches: no planes in HBM
    unsigned long long* diag;    // host-mapped 8-word record the fused kernel fills when one of its waits times out (may be null)
    // generator-fused slice pass: the Gram operand is a UNIFORM Parameter only this step reads; its Philox fill moves from the prologue launch
    // into the slice kernel (gen_counter = the context's update counter, already ticked by the prologue; null: off)
    const unsigned long long* gen_counter; uint64_t gen_seed; double gen_scale; int64_t gen_inst_offset;
    int32_t gen_writeback;       // 1: the pass also writes the generated values to the Parameter's buffer; 0: the buffer is filled on demand (pmb_param_device_ptr)
};
// text for a non-empty diagnostic record of the fused kernel ("" when the record is empty)
const char* gram_i8_fused_diag_text(const unsigned long long* diag, char* buf, size_t len);
// memset(stats) + slice kernel (gram_i8_slice) and the two tcgen05 Gram launches, 5 and 6 planes, of which at most one does work per
// instance (gram_i8_mma).  false: not applicable (the caller runs the fp64 kernel alone).
// [col0, col1): column range of this launch (upload pipelining: one slice + one Gram launch per arrived chunk); col1 < 0 = all columns.
// The Gram launch of a range computes the tile columns whose last column lies in the range.
bool gram_i8_slice_can_generate(const GramArgs& g, const GramI8Plan& plan);
bool gram_i8_fused_can_generate(const GramArgs& g);   // same for the fused small-instance kernel   // GramI8Scratch::gen_counter may be set for this operand
bool launch_gram_i8_slice(const GramArgs& g, const GramI8Scratch& sc, int64_t batch, int sm_count, cudaStream_t stream, int64_t col0 = 0,
                          int64_t col1 = -1);
bool launch_gram_i8_mma(const GramArgs& g, const GramI8Scratch& sc, int64_t batch, int sm_count, cudaStream_t stream, int64_t col0 = 0,
                        int64_t col1 = -1);
void launch_gram_i8_flag(const GramI8Scratch& sc, int64_t batch, int32_t* flag_dev, cudaStream_t stream);
// counts[0..2] += number of instances whose certificate picked 5 planes / 6 planes / the fp64 kernel (introspection for tests and bench)
void launch_gram_i8_count_levels(const GramI8Scratch& sc, int64_t batch, unsigned long long* counts_dev, cudaStream_t stream);
// predicted seconds of the two ways to compute one SYRK step; the planner picks the int8 path when it is the faster one
bool gram_i8_pays_off(int64_t k, int64_t n, int64_t batch, int sm_count);
void configure_gram_i8_kernels();
// K3: fused small-instance route (n <= 256, k <= 256, even k, 16-byte aligned operands): slicing, the 15 digit-plane products and the
// epilogue of one instance per CTA pair in ONE launch; statistics / verdict land in sc.stats (stride kGramI8StatsStride, followed by the
// count of uncertified instances).  The fp64 launch that follows computes H of the uncertified instances only.
bool gram_i8_fused_applicable(const GramArgs& g, int64_t batch);
bool gram_i8_fused_pays_off(int64_t k, int64_t n, int64_t batch, int sm_count);
bool launch_gram_i8_fused(const GramArgs& g, const GramI8Scratch& sc, int64_t batch, int sm_count, cudaStream_t stream);
void configure_gram_i8_fused_kernel();

// ---------------- small kernels ---------------------------------------------------------------
// fused prologue of one update: all built-in Parameter generators (kind 0: Philox uniform fill, per instance) + zero fills
// (kind 1: numel doubles in total) + the update-counter tick (published by the last block to finish)
struct FillJob {
    double* out;
    int64_t numel;
    uint64_t seed;
    double scale;
    int32_t kind;
    int32_t nblocks;   // blocks of 256 threads assigned to this job
};
// inst_offset: global id of local instance 0 (Philox key).  tick = 1: the update's first prologue launch (update id = counter + 1, published by
// the last block); tick = 0: a later launch of the same update (update id = counter, nothing published)
void launch_prologue(const FillJob* jobs_dev, int njobs, int total_blocks, int64_t batch, int64_t inst_offset, unsigned long long* counter,
                     unsigned int* ticket, int tick, cudaStream_t stream);
// out[i] = cst[i] + sum_t alpha_t * src_t[off_t + i]
struct LinCombTermDev { const double* src; int64_t src_bstride; int64_t src_off; AlphaDev alpha; };
void launch_lincomb(double* out, int64_t out_bstride, int64_t n, const LinCombTermDev* terms, int nterms,
                    const double* cst, int64_t batch, cudaStream_t stream);
// out (cols x rows) = in (rows x cols)'       (adjoint buffer, lazyexpression.jl:206-217)
void launch_transpose(double* out, int64_t out_bstride, const double* in, int64_t in_bstride, int64_t rows, int64_t cols,
                      int64_t batch, cudaStream_t stream);
// y = A v (trans = 0, len rows) or y = A' v (trans = 1, len cols); optional scatter: y[map[i]] (=|+=) scale * value
struct GemvArgs {
    const double* A; int64_t lda, a_bstride; int64_t rows, cols; int32_t trans;
    const double* v; int64_t v_bstride; int64_t v_off;
    double* y; int64_t y_bstride;
    const int32_t* map; AlphaDev scale; int32_t accumulate;
};
void launch_gemv(const GemvArgs& a, int64_t batch, cudaStream_t stream);
// out[0] (=|+=) scale * sum_i x[xoff+i] * y[yoff+i]
struct DotArgs {
    const double* x; int64_t x_bstride, x_off;
    const double* y; int64_t y_bstride, y_off;
    int64_t n; double* out; int64_t out_bstride; AlphaDev scale; int32_t accumulate;
};
void launch_dot(const DotArgs& a, int64_t batch, cudaStream_t stream);
void launch_gemv_t_dot(const GemvArgs& a, const DotArgs& d, int64_t batch, cudaStream_t stream);  // y = A'v and out (+)= s * x.y in one launch
// scalar chain: out = c * prod_i s_i
void launch_scalar_prod(double* out, int64_t out_bstride, double c, const double* s0, int64_t s0_bstride,
                        const double* s1, int64_t s1_bstride, int64_t batch, cudaStream_t stream);
// q[map[j]] (=|+=) alpha * v[off + j]   (dot(numbers, variables), functions.jl:676-687)
struct ScatterVecArgs {
    const double* v; int64_t v_bstride, v_off; int64_t n; const int32_t* map;
    double* q; int64_t q_bstride; AlphaDev alpha; int32_t accumulate;
};
void launch_scatter_vec(const ScatterVecArgs& a, int64_t batch, cudaStream_t stream);
// bilinearmul! (functions.jl:840-858) + symmetric MOI Hessian:  H[xm[r], ym[c]] += a*Q[c,r];  H[ym[c], xm[r]] += a*Q[c,r]
struct BilinearArgs {
    const double* Q; int64_t ldq, q_bstride; int64_t nx, ny;
    const int32_t* xmap; const int32_t* ymap;
    double* H; int64_t ldh, h_bstride; AlphaDev alpha;
};
void launch_bilinear_scatter(const BilinearArgs& a, int64_t batch, cudaStream_t stream);
// constant quadratic terms: H[i,j] += a*c ; H[j,i] += a*c  (i == j: += 2ac, moi_interop.jl:58)
struct CooQuadArgs {
    const int32_t* i; const int32_t* j; const double* c; int64_t nnz;
    double* H; int64_t ldh, h_bstride; AlphaDev alpha;
};
void launch_coo_quad(const CooQuadArgs& a, int64_t batch, cudaStream_t stream);
// general scatter of a scratch product into the symmetric Hessian: H[xm[i], ym[j]] += a*M[i,j]; H[ym[j], xm[i]] += a*M[i,j]
struct ScatterSymArgs {
    const double* M; int64_t ldm, m_bstride; int64_t n, p;
    const int32_t* xmap; const int32_t* ymap;
    double* H; int64_t ldh, h_bstride; AlphaDev alpha;
};
void launch_scatter_sym(const ScatterSymArgs& a, int64_t batch, cudaStream_t stream);
void launch_fill(double* p, int64_t n, double v, cudaStream_t stream);
// out[b * nv + i] = x[b * N + idx[i]]     (value.(model, x): gather through model_var_to_optimizer, model.jl:166)
void launch_gather(double* out, const double* x, const int64_t* idx, int64_t nv, int64_t N, int64_t batch, cudaStream_t stream);

// ---------------- term-list export + QP pack (kernels_export.cu) ---------------------------------------------------------
struct TermEntryDev {      // one term of the per-row template
    int32_t kind;          // 0: matrix term (coefficient = alpha * src[row + srccol * ld]),  1: unit term (coefficient = alpha)
    int32_t term;          // index into mats / units
    int64_t srccol;        // matrix: source column;  unit: offset into unit_vars of this unit's per-row variable table
    int64_t var;           // matrix: 1-based optimizer variable;  unit: unused
};
struct ExportArgs {
    const TermEntryDev* tmpl; int32_t T;          // terms per row
    const AffMatTermDev* mats; const AffUnitTermDev* units;
    const int64_t* unit_vars;                     // per-row 1-based optimizer variable of unit terms
    int64_t rows;
    void* out;                                    // batch * rows * T records of 24 bytes
};
void launch_export_terms(const ExportArgs& a, int64_t batch, cudaStream_t stream);
struct TermGatherDev {     // one term of a block whose rows have different layouts (static, reference order)
    int32_t kind;          // 0: matrix term, 1: unit term
    int32_t term;          // index into mats / units
    int64_t src_off;       // matrix: row + srccol * ld inside one instance of the source
    int64_t row;           // 1-based output row
    int64_t var;           // 1-based optimizer variable
};
void launch_export_terms_gather(const TermGatherDev* tab, int64_t nterms, const AffMatTermDev* mats, const AffUnitTermDev* units, void* out, int64_t batch,
                                cudaStream_t stream);
// objective as MOI.ScalarQuadraticFunction records (combined upper triangle, row-outer) + ScalarAffineTerm records
void launch_export_objective(const double* Q, const double* q, int64_t N, void* quad_out, void* aff_out, int64_t batch, cudaStream_t stream);
// P_x[j(j+1)/2 + i] = Q[i + j*N]  (i <= j)
void launch_pack_upper(const double* Q, int64_t N, int64_t q_bstride, double* Px, int64_t px_bstride, int64_t batch, cudaStream_t stream);
struct PackBlockDev {      // one constraint block inside the stacked CSC matrix
    const double* C;       // dense: rows x N column-major;  sparse: value buffer
    const double* d;
    int64_t c_bstride, d_bstride;
    int64_t rows, row_off; // rows of the block, first row in the stacked matrix
    int32_t set_kind, sparse;
    int64_t nnz;                 // sparse: stored values
    const int64_t* colbase;      // dense: [N] position of this block's first entry in column j of A_x
    const int64_t* pos;          // sparse: [nnz] position of each stored value in A_x
};
void launch_pack_constraints(const PackBlockDev* blocks_dev, int nblocks, int64_t N, double* Ax, int64_t ax_bstride, double* l, double* u,
                             int64_t m_total, int64_t batch, int sm_count, cudaStream_t stream);

}  // namespace pmb
