// K2i -- certified int8-slice Gram on the 5th-generation tensor cores (tcgen05 kind::i8, accumulators in TMEM).
//
// Same reference semantics as kernels_gram_tma.cu (vecdot! + muladd!, functions.jl:548-576,702-709, then the symmetric MOI
// Hessian of moi_interop.jl:45-62:  H = scale * X'X, both triangles), but the fp64 contraction is carried out EXACTLY in
// integer arithmetic, because tcgen05 has no f64 kind and the DMMA pipe (37 TFLOP/s) is ~100x slower than the int8 pipe:
//
//   1. slice kernel: every column j of X is scaled by 1 / sigma_j (sigma_j = largest magnitude / 0.99), rounded to a 47-bit signed
//      fixed-point integer and cut into 6 balanced base-256 digits d_0..d_5 in [-128, 127] (int8 planes, K-major like X).  The same
//      pass records, per instance, max sigma_j, max ||x_j||^2, max ||x_j / sigma_j||_1 and the largest 2-norm of every digit plane.
//   2. Gram kernel <NS planes, BN, CTAS, KC>: for an output tile (128 columns i) x (BN columns j) all digit-plane products with
//      s + t <= NS-1 are issued as tcgen05.mma kind::i8 (M = 128, K = 32) from TMA-loaded, swizzled shared-memory panels; planes with
//      equal s + t share an int32 accumulator in TMEM (NS accumulators x BN columns <= 512 TMEM columns), and one instruction covers
//      up to 256 consecutive accumulator columns (several planes / levels).  Integer accumulation is exact, so the result does not
//      depend on tiling, k order, pipeline variant or the number of SMs.  Three pipelines share the kernel body (I8Cfg):
//        KC = 64,  CTAS = 1: 3-stage ring of whole k steps (default for 5 planes)
//        KC = 128, CTAS = 1: 128-byte rows, split ring: 2 column-panel stages + 6 row-panel slots (default for 6 planes)
//        KC = 64,  CTAS = 2: CTA pair, tcgen05 cta_group::2 with M = 256 (measured slower; kept as an experiment, PMB_I8_PAIR=1)
//   3. epilogue: phase 1 pulls the NS int32 accumulators of the thread's 48 entries out of TMEM (Horner chain in fp64, one rounding at
//      2^-53) and hands TMEM back, so the next tile's MMAs start; phase 2 scales by sigma_i sigma_j 2^(-8 NS - 6) and the caller's
//      alpha, scatters through the variable map and writes both triangles (and, when asked, the packed upper triangle P_x).
//
// What is dropped are the digit products with s + t >= NS and the representation error of the planes used.  gram_i8_level()
// (pmb_kernels.h, shared with the fp64 kernel) bounds both with the recorded norms and picks, per instance, the smallest NS in {5, 6}
// whose worst-case error is below tol * max |H| (max |H| of a Gram matrix is its largest diagonal entry) -- the north_star tolerance
// is relative to the largest magnitude of the array.  The <5, 96> and <6, 80> launches each process the instances of their level;
// instances that cannot be certified (wide dynamic range inside the matrix, non-finite values) are left to the fp64 DMMA kernel.
#include <cuda.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <math_constants.h>
#include <cstdlib>

#include "pmb_kernels.h"
#include "pmb_i8_ptx.h"
#include "pmb_philox.h"

namespace pmb {

constexpr int kI8BM = 128;          // tile rows = tcgen05 M (one TMEM lane per row)
constexpr int kI8KC = 64;           // default k (bytes) per pipeline stage: one 64-byte swizzle span, two K=32 MMAs per digit-plane pair
                                    // (KC = 32: one 32-byte swizzle span, one MMA; twice the stages for the same bytes in flight)
constexpr int kI8EpiWarps = 8;       // two warps per TMEM lane quarter, alternating 16-column chunks
constexpr int kI8Threads = 64 + 32 * kI8EpiWarps;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int kI8TmemCols = 512;

// CTAS = 1: one CTA per (128 x BN) tile, tcgen05 cta_group::1.  CTAS = 2: a CTA pair (cluster of 2, one TPC) per (256 x BN) tile,
// cta_group::2 with M = 256: CTA r holds rows 128 r .. 128 r + 127 of the tile (own row panel, own TMEM accumulators) and supplies HALF of
// every column-panel operand -- for an instruction over two consecutive planes t, t+1 (N = 2 BN) CTA 0 holds plane t and CTA 1 plane t+1 of
// the tile's BN columns ("pair slot" t/2), for an instruction over a single plane (N = BN; the planes t = NS-1-s with NS-s odd, all even)
// CTA r holds columns BN/2 r .. of it ("half slot" t/2).  Per k step a CTA then writes 27.5-30.5 KB and the tensor pipe reads 58.5 KB of
// shared memory instead of 35 + 81 KB: the one-CTA kernel is bound by exactly that traffic (161 B/clk wanted, 128 B/clk available).
template <int NS, int BN, int CTAS = 1, int KC = kI8KC>
struct I8Cfg {
    static constexpr int kRows = kI8BM * CTAS;     // tile rows
    static constexpr int kKC = KC;
    static constexpr int kITile = kI8BM * KC;   // bytes of one digit plane of the (CTA's) row panel
    static constexpr int kJTile = BN * KC;
    static constexpr int kNPair = NS / 2, kNHalf = (NS + 1) / 2;                    // CTAS = 2: pair slots and half slots per stage
    static constexpr int kHalfBytes = BN / 2 * KC;                               // bytes a half slot receives
    static constexpr int kHalfTile = (kHalfBytes + 16 * KC - 1) / (16 * KC) * (16 * KC);   // its footprint (slots stay aligned to two swizzle periods)
    // KC = 128 ("wide"): rows of 128 bytes under the 128-byte swizzle.  The TMA engine moves about one box ROW per clock whatever its
    // length (measured: 64-byte and 32-byte rows take the same time, one box per plane or one per panel as well), so 128-byte rows
    // double the fill rate.  A whole k step (NS (16 + BN / 8) KB = 140 KB) no longer double-buffers, so the ring is split: kStages = 2
    // column-panel stages (all NS planes, consecutive: the grouped MMAs span plane boundaries) and kNA row-panel slots of one plane;
    // the MMA issuer walks the rows s = 0 .. NS-1 of a k step and releases each row-panel slot as soon as its row is done.
    static constexpr bool kWide = KC == 128;
    static constexpr int kBStage = NS * kJTile;
    static constexpr int kStageBytes = kWide ? kBStage : CTAS == 1 ? NS * (kITile + kJTile) : NS * kITile + kNPair * kJTile + kNHalf * kHalfTile;
    static constexpr int kTxBytes = CTAS == 1 ? kStageBytes : NS * kITile + kNPair * kJTile + kNHalf * kHalfBytes;   // per CTA and stage
    static constexpr int kStages = kWide ? 2 : (227 * 1024 - 2048) / kStageBytes;
    static constexpr int kNA = kWide ? (227 * 1024 - 2048 - kStages * kStageBytes) / kITile : 1;
    static constexpr int kSmemBytes = kStages * kStageBytes + (kWide ? kNA * kITile : 0) + 1024;
    static constexpr int kGrp = 256 / BN;          // column-panel planes covered by one MMA (N = kGrp * BN <= 256)
    static constexpr int kAccBufs = (2 * NS * BN <= kI8TmemCols) ? 2 : 1;   // narrow tiles: two accumulator sets, the epilogue of a
                                                                            // tile overlaps the MMAs of the next (short-k problems)
    static_assert(CTAS == 1 || CTAS == 2, "one CTA or a CTA pair");
    static_assert(KC == 128 || KC == 64 || KC == 32, "one swizzle span per stage");
    static_assert(!kWide || (CTAS == 1 && kNA >= 2 && kNA <= 8), "wide stages: one-CTA kernel");
    static_assert(NS * BN <= kI8TmemCols, "accumulators do not fit TMEM");
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "tcgen05 kind::i8, M = 128: N % 16 == 0");
    static_assert(CTAS == 1 || (2 * BN <= 256 && BN % 16 == 0), "cta_group::2, M = 256: N % 16 == 0, N <= 256");
    static_assert(kStages >= 2, "pipeline needs two stages");
};

struct GramI8KernelArgs {
    CUtensorMap tmi, tmj;   // digit planes (kp, n, 6, batch) uint8; boxes (64, 128, 1, 1) and (64, BN, 1, 1), 64-byte swizzle
    CUtensorMap tmh;        // CTA-pair kernel: box (64, BN / 2, 1, 1)
    CUtensorMap tmja;       // wide stages: all NS planes of the column panel in ONE box (128, BN, NS, 1)
    int64_t k, kp, n;
    const double* sigma;    // [batch][n]
    const unsigned long long* stats;
    double tol;
    int32_t force_ns;
    double* out; int64_t ldo, o_bstride;
    const int32_t* map;
    AlphaDev scale;
    int32_t accumulate;
    double* packed_dev; int64_t packed_bstride;
    int32_t tj0, tj1;       // tile columns [tj0, tj1) of every instance
    int64_t tiles_per_inst, total_tiles;
    int32_t dbg;            // timing diagnostics only (PMB_I8_DBG): 1 = no TMA loads, 2 = no MMAs, 4 = only the s = 0 row of MMAs,
                            // 8 = no epilogue work, 16 = the s = 0 row three times (same MAC count, N = 240 instructions only)
};

namespace i8 {
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kI8EpiWarps) : "memory"); }
}  // namespace i8

// ---- slice kernel --------------------------------------------------------------------------------------------------------------
struct GramI8SliceArgs {
    const double* X; int64_t ldx, x_bstride;
    int64_t k, kp, n, batch;
    int8_t* slices;        // [batch][6][n][kp]
    double* sigma;         // [batch][n]
    unsigned long long* stats;
    double tol;
    int64_t col0, ncols;   // this launch slices columns [col0, col0 + ncols) of every instance (upload pipelining: one launch per chunk)
    // fused linear part of the least-squares objective (the pass reads X anyway):  q[map[j]] += qscale * sum_r X[r, j] g[r],
    // c0 += cscale * sum_r g[r]^2 (by the group that owns column 0).  gvec == null: off
    const double* gvec; int64_t g_bstride;
    double* qout; int64_t q_bstride;
    double* c0out;
    const int32_t* map;
    AlphaDev qscale, cscale;
    // GEN: X is a UNIFORM Parameter nobody else reads -- its Philox fill runs HERE instead of in the prologue launch (same bits: pmb_philox.h); the
    // values are still written to X (the Parameter stays readable), but never read back
    uint64_t gen_seed; double gen_sc53; const unsigned long long* gen_counter; int64_t gen_inst_offset; int32_t gen_writeback;
};

// max of two doubles as one compare + select (fmax() expands to 8 instructions for its NaN rules: DSETP.MAX, FSEL, SEL, a predicated
// LOP3 and moves).  A NaN operand loses every comparison, which is fine here: non-finite columns are detected through the 1-norm
__device__ __forceinline__ double dmax(double a, double b) { return b > a ? b : a; }

__device__ __forceinline__ void load4(const double* x, int64_t r0, int64_t k, bool vec, double (&v)[4]) {
    if (vec && r0 + 3 < k) {
        const double2 p = __ldg((const double2*)(x + r0)), q = __ldg((const double2*)(x + r0 + 2));
        v[0] = p.x; v[1] = p.y; v[2] = q.x; v[3] = q.y;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = r0 + e < k ? __ldg(x + r0 + e) : 0.0;
    }
}

// A group of T threads (8 / 16 lanes of a warp for short columns, one warp, or the whole CTA for long columns) owns one column of X at
// a time.  Columns of up to 16 T rows
// stay in registers between the statistics pass (largest magnitude, 2-norm, 1-norm, X'g) and the digit pass: X is read ONCE.
// RING (T = 256, columns of at most 4096 rows, 16-byte aligned): the columns travel global -> shared memory as 1-D bulk copies
// (cp.async.bulk on an mbarrier, SASS UBLKCP) three columns ahead of the arithmetic, so the HBM latency is off the dependent chain
// load -> reduce -> digits -> store of a column.
#ifndef PMB_SLICE_RING
#define PMB_SLICE_RING 3
#endif
#ifndef PMB_SLICE_CTAS
#define PMB_SLICE_CTAS 2
#endif
constexpr int kI8SliceRing = PMB_SLICE_RING;
// FULL (ring only): k == kp is a multiple of 1024 and the offset vector, if any, is 16-byte aligned -- every row of every cached iteration
// that runs exists, so the bounds tests, masked loads and their reconvergence points (about a fifth of the instructions of this issue-bound
// kernel) are compiled out
template <int T, bool RING, bool FULL = false, bool GEN = false>
__global__ void __launch_bounds__(256, PMB_SLICE_CTAS) gram_i8_slice_kernel(const GramI8SliceArgs a) {
    constexpr int NP = kGramI8Planes, G = 256 / T, CIT = 4;   // CIT cached iterations of 4 rows per thread
    static_assert(!RING || T == 256, "the ring feeds one column per CTA");
    static_assert(!FULL || RING || GEN, "the full-column fast path belongs to the ring and generator kernels");
    static_assert(!GEN || (T == 256 && FULL && !RING), "the generator-fused pass is the full-column, one-column-per-CTA variant without the ring");
    extern __shared__ __align__(128) unsigned char slice_smem[];
    __shared__ __align__(8) uint64_t ring_full[kI8SliceRing];
    __shared__ double red[8][5];
    __shared__ unsigned int redu[8][NP];
    const int64_t kpe = (a.k + 1) / 2 * 2;   // rows copied per column (ld is even on this path, so the pad row exists)
    if (RING) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < kI8SliceRing; ++s) i8::mb_init(&ring_full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            for (int s = 0; s < kI8SliceRing; ++s) {
                const int64_t c = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
                if (c < a.batch * a.ncols) {
                    const int64_t inst = c / a.ncols, j = a.col0 + (c - inst * a.ncols);
                    i8::mb_expect_tx(&ring_full[s], (uint32_t)(kpe * 8));
                    i8::bulk_load_1d(i8::s_u32(slice_smem) + (uint32_t)(s * kpe * 8), a.X + inst * a.x_bstride + j * a.ldx, (uint32_t)(kpe * 8), &ring_full[s]);
                }
            }
        }
        __syncthreads();
    }
    uint32_t ring_it = 0;
    const uint32_t gen_upd = GEN ? (uint32_t)*a.gen_counter : 0u;   // the prologue launch of this update has already published the update id
    // per-group running statistics (thread tig == 0 only)
    int64_t acc_inst = 0;
    unsigned long long acc_n = 0;
    bool acc_bad = false;
    double acc_sig = 0.0, acc_ss = 0.0, acc_l1 = 0.0;
    uint32_t acc_dsq[NP];
#pragma unroll
    for (int s = 0; s < NP; ++s) acc_dsq[s] = 0u;
    auto flush = [&]() {
        if (!acc_n) return;
        unsigned long long* st = a.stats + acc_inst * kGramI8StatsStride;
        if (acc_bad) atomicMax(st, (unsigned long long)__double_as_longlong(CUDART_INF));
        else if (acc_sig > 0.0) {
            atomicMax(st, (unsigned long long)__double_as_longlong(acc_sig));
            atomicMax(st + 1, (unsigned long long)__double_as_longlong(acc_ss));
            atomicMax(st + 2, (unsigned long long)__double_as_longlong(acc_l1));
#pragma unroll
            for (int s = 0; s < NP; ++s) atomicMax(st + 3 + s, (unsigned long long)acc_dsq[s]);
        }
        // the group that completes the instance's column count evaluates the certificate once for everybody
        __threadfence();
        if (atomicAdd(st + 14, acc_n) + acc_n == (unsigned long long)(a.col0 + a.ncols)) {
            __threadfence();
            // verdict for everything sliced so far; sticky across the launches of one update (5 planes < 6 planes < fp64): a tile computed
            // under an earlier, smaller max |H| stays certified because max |H| only grows as columns arrive
            const unsigned long long lv = (unsigned long long)gram_i8_compute_level(st, a.tol), prev = __ldcg(st + 15);
            st[15] = lv > prev ? lv : prev;
        }
        acc_n = 0; acc_bad = false; acc_sig = 0.0; acc_ss = 0.0; acc_l1 = 0.0;
#pragma unroll
        for (int s = 0; s < NP; ++s) acc_dsq[s] = 0u;
    };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tig = threadIdx.x % T, grp = threadIdx.x / T;
    const bool vec = (a.ldx % 2 == 0) && (((uintptr_t)a.X & 15) == 0) && (a.x_bstride % 2 == 0);
    const bool gvecal = a.gvec && (((uintptr_t)a.gvec & 15) == 0) && (a.g_bstride % 2 == 0);
    const bool cached = FULL || a.kp <= (int64_t)4 * T * CIT;
    const bool k4 = FULL || (a.k & 3) == 0;
    const int nit = (int)(a.k / (4 * T));   // FULL: k is a multiple of 4 T = 1024 rows, whole cached iterations
    const int64_t plane_stride = a.n * a.kp;
    const int64_t total = a.batch * a.ncols;
    for (int64_t c0 = (int64_t)blockIdx.x * G; c0 < total; c0 += (int64_t)gridDim.x * G) {   // uniform per CTA (block-wide barriers below)
        const int64_t c = c0 + grp;
        const bool live = c < total;
        int64_t inst = 0, j = 0;
        if (live) {   // 64-bit division is ~70 instructions: the planner guarantees batch * n < 2^31
            const uint32_t q32 = (uint32_t)c / (uint32_t)a.ncols;
            inst = q32; j = a.col0 + (int64_t)((uint32_t)c - q32 * (uint32_t)a.ncols);
        }
        const double* x = a.X + inst * a.x_bstride + j * a.ldx;
        const double* g = a.gvec ? a.gvec + inst * a.g_bstride : nullptr;
        const int rs = (int)(ring_it % kI8SliceRing);
        const double* xs = (const double*)slice_smem + (int64_t)rs * kpe;
        if (RING) i8::mb_wait(&ring_full[rs], (ring_it / kI8SliceRing) & 1);
        double vc[CIT][4];
        double amax = 0.0, ss = 0.0, l1 = 0.0, dg = 0.0, dg2 = 0.0, gg = 0.0;
        bool bad = false;
        auto stat4 = [&](const double (&v)[4], int64_t r0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double av = fabs(v[e]);
                amax = dmax(amax, av);
                l1 += av;   // Inf / NaN entries surface here (fmax would hide a NaN): tested once per column below
                ss = fma(v[e], v[e], ss);
            }
            if (g) {
                double gv[4];
                if (FULL) {
                    const double2 p = __ldg((const double2*)(g + r0)), q = __ldg((const double2*)(g + r0 + 2));
                    gv[0] = p.x; gv[1] = p.y; gv[2] = q.x; gv[3] = q.y;
                } else
                load4(g, r0, a.k, gvecal, gv);
                dg = fma(v[0], gv[0], dg); dg2 = fma(v[1], gv[1], dg2); dg = fma(v[2], gv[2], dg); dg2 = fma(v[3], gv[3], dg2);
                if (j == 0) {   // g.g is needed once per instance: the group that owns column 0 computes it
#pragma unroll
                    for (int e = 0; e < 4; ++e) gg = fma(gv[e], gv[e], gg);
                }
            }
        };
        if (live) {
            if (cached) {
#pragma unroll
                for (int it = 0; it < CIT; ++it) {
                    const int64_t r0 = ((int64_t)it * T + tig) * 4;
                    if (FULL ? it < nit : r0 < a.k) {
                        if (GEN) {
                            // rows r0 .. r0 + 3 of column j: elements j k + r0 .. of the instance's buffer = pairs p, p + 1 (k is even)
                            const uint64_t p0 = ((uint64_t)j * (uint64_t)a.k + (uint64_t)r0) >> 1;
                            philox_uniform_pairs2(p0, p0 + 1, (uint32_t)(inst + a.gen_inst_offset), gen_upd, a.gen_seed, a.gen_sc53, vc[it]);
                            if (a.gen_writeback) {
                                double2* o = reinterpret_cast<double2*>(const_cast<double*>(x) + r0);
                                o[0] = make_double2(vc[it][0], vc[it][1]); o[1] = make_double2(vc[it][2], vc[it][3]);
                            }
                        } else if (RING) {
                            const double2 p = *(const double2*)(xs + r0), q = *(const double2*)(xs + r0 + 2);   // r0 + 3 < kpe + 2 <= slot + pad
                            vc[it][0] = p.x; vc[it][2] = q.x;
                            if (k4) { vc[it][1] = p.y; vc[it][3] = q.y; }
                            else { vc[it][1] = r0 + 1 < a.k ? p.y : 0.0; vc[it][2] = r0 + 2 < a.k ? q.x : 0.0; vc[it][3] = r0 + 3 < a.k ? q.y : 0.0; }
                        } else {
                            load4(x, r0, a.k, vec, vc[it]);
                        }
                        stat4(vc[it], r0);
                    }
                    else { vc[it][0] = vc[it][1] = vc[it][2] = vc[it][3] = 0.0; }
                }
            } else {
                for (int64_t r0 = (int64_t)tig * 4; r0 < a.k; r0 += 4 * T) {
                    double v[4];
                    load4(x, r0, a.k, vec, v);
                    stat4(v, r0);
                }
            }
        }
        dg += dg2;
        constexpr int W = T < 32 ? T : 32;   // lanes that share a column inside one warp (short columns: several columns per warp)
#pragma unroll
        for (int o = W / 2; o; o >>= 1) {
            amax = dmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
            l1 += __shfl_xor_sync(0xffffffffu, l1, o);
            dg += __shfl_xor_sync(0xffffffffu, dg, o);
            gg += __shfl_xor_sync(0xffffffffu, gg, o);
        }
        bad = !(l1 <= DBL_MAX);
        if (T == 256) {
            if (lane == 0) { red[warp][0] = bad ? CUDART_INF : amax; red[warp][1] = ss; red[warp][2] = l1; red[warp][3] = dg; red[warp][4] = gg; }
            __syncthreads();
            if (RING && threadIdx.x == 0) {   // every thread holds its rows in registers: the slot can take the column 3 ahead
                const int64_t cn = c0 + (int64_t)kI8SliceRing * gridDim.x;
                if (cn < total) {
                    const uint32_t q2 = (uint32_t)cn / (uint32_t)a.ncols;
                    const int64_t in2 = q2, j2 = a.col0 + (int64_t)((uint32_t)cn - q2 * (uint32_t)a.ncols);
                    i8::mb_expect_tx(&ring_full[rs], (uint32_t)(kpe * 8));
                    i8::bulk_load_1d(i8::s_u32(slice_smem) + (uint32_t)(rs * kpe * 8), a.X + in2 * a.x_bstride + j2 * a.ldx, (uint32_t)(kpe * 8), &ring_full[rs]);
                }
            }
            ++ring_it;
            amax = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) amax = dmax(amax, red[w][0]);
            if (tig == 0) {   // the column's sums are only consumed by the thread that keeps the running statistics
                ss = 0.0; l1 = 0.0; dg = 0.0; gg = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { ss += red[w][1]; l1 += red[w][2]; dg += red[w][3]; gg += red[w][4]; }
            }
            bad = !(amax <= DBL_MAX);
        }
        if (amax > 0.0 && !(amax >= 1e-150 && amax <= 1e150)) bad = true;
        const bool zero = !(amax > 0.0) || bad;
        const double sigma = zero ? 1.0 : amax * (1.0 / 0.99);
        const double mult = zero ? 0.0 : (0.99 / amax) * 140737488355328.0;   // 2^47: x -> fixed point with 47 fractional bits of x / sigma
        int8_t* dst = a.slices + ((inst * NP) * a.n + j) * a.kp;
        asm volatile("" : "+l"(dst));   // keep the column's base in registers: the compiler otherwise rebuilds the 64-bit product per 4-row group
        uint32_t dsq[NP];
#pragma unroll
        for (int s = 0; s < NP; ++s) dsq[s] = 0u;
        // Balanced base-256 digits of the 48-bit two's-complement integer X without a carry chain: adding 0x80 to each of the five low
        // bytes turns byte s of the sum into (digit + 128), XOR 0x80 maps that to the digit's int8 bit pattern; the top byte is the
        // (signed) leading digit.  Byte b of the result belongs to plane 5 - b.  Four rows are transposed into one 32-bit word per plane
        // with byte permutes; sum d^2 per plane is one dp4a.
        auto digits4 = [&](const double (&v)[4], int64_t r0) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                // round-to-nearest-even integer of v * mult (|.| < 2^47) without the slow 64-bit F2I: adding 1.5 * 2^52 leaves the
                // two's-complement integer in the low mantissa bits; only the low 48 bits are used below
                const unsigned long long X = (unsigned long long)__double_as_longlong(fma(v[t], mult, 6755399441055744.0));
                const unsigned long long Y = (X + 0x0000008080808080ull) ^ 0x0000008080808080ull;
                lo[t] = (uint32_t)Y; hi[t] = (uint32_t)(Y >> 32);
            }
            int8_t* dp = dst + r0;
#pragma unroll
            for (int p = 0; p < NP; ++p, dp += plane_stride) {
                const int bidx = (NP - 1 - p) & 3;
                const uint32_t* src = (NP - 1 - p) < 4 ? lo : hi;
                const uint32_t sel = (uint32_t)bidx | ((uint32_t)(4 + bidx) << 4);
                const uint32_t w = __byte_perm(__byte_perm(src[0], src[1], sel), __byte_perm(src[2], src[3], sel), 0x5410);
                dsq[p] = (uint32_t)__dp4a((int)w, (int)w, (int)dsq[p]);
                *(uint32_t*)dp = w;
            }
        };
        if (live) {
            if (cached) {
#pragma unroll
                for (int it = 0; it < CIT; ++it) {
                    const int64_t r0 = ((int64_t)it * T + tig) * 4;
                    if (FULL ? it < nit : r0 < a.kp) digits4(vc[it], r0);
                }
            } else {
                for (int64_t r0 = (int64_t)tig * 4; r0 < a.kp; r0 += 4 * T) {
                    double v[4];
                    load4(x, r0, a.k, vec, v);
                    digits4(v, r0);
                }
            }
        }
        if (W == 32) {
#pragma unroll
            for (int s = 0; s < NP; ++s) dsq[s] = __reduce_add_sync(0xffffffffu, dsq[s]);
        } else {
#pragma unroll
            for (int s = 0; s < NP; ++s)
#pragma unroll
                for (int o = W / 2; o; o >>= 1) dsq[s] += __shfl_xor_sync(0xffffffffu, dsq[s], o);
        }
        if (T == 256) {
            if (lane == 0) {
#pragma unroll
                for (int s = 0; s < NP; ++s) redu[warp][s] = dsq[s];
            }
            __syncthreads();
            if (tig == 0) {
#pragma unroll
                for (int s = 0; s < NP; ++s) {
                    dsq[s] = 0u;
#pragma unroll
                    for (int w = 0; w < 8; ++w) dsq[s] += redu[w][s];
                }
            }
            __syncthreads();   // red / redu are rewritten by the next column
        }
        if (live && tig == 0) {
            a.sigma[inst * a.n + j] = sigma;
            if (g) {
                const double qs = a.qscale.ptr ? a.qscale.c * a.qscale.ptr[inst * a.qscale.bstride] : a.qscale.c;
                atomicAdd(a.qout + inst * a.q_bstride + (a.map ? a.map[j] : j), qs * dg);   // one writer per entry: a fire-and-forget reduction
                if (j == 0) {
                    const double cs = a.cscale.ptr ? a.cscale.c * a.cscale.ptr[inst * a.cscale.bstride] : a.cscale.c;
                    atomicAdd(a.c0out + inst, cs * gg);
                }
            }
            // running maxima of this group's columns; the global atomics happen once per (group, instance), not once per column
            if (acc_n && acc_inst != inst) flush();
            acc_inst = inst; ++acc_n;
            if (bad) acc_bad = true;
            else if (amax > 0.0) {
                acc_sig = fmax(acc_sig, sigma); acc_ss = fmax(acc_ss, ss); acc_l1 = fmax(acc_l1, l1 / sigma);
#pragma unroll
                for (int s = 0; s < NP; ++s) acc_dsq[s] = acc_dsq[s] > dsq[s] ? acc_dsq[s] : dsq[s];
            }
        }
    }
    if (tig == 0) flush();
}

// ---- Gram kernel ---------------------------------------------------------------------------------------------------------------
struct I8Unit { int64_t inst; int ti, tj; };

// upper tiles, column-major: tile column tj (BN wide) holds the row tiles ti = 0 .. (last column of tj) / 128
template <int BN, int ROWS = kI8BM>
__device__ __forceinline__ I8Unit i8_decode(const GramI8KernelArgs& a, int64_t u) {
    I8Unit w;
    w.inst = u / a.tiles_per_inst;
    int64_t tl = u - w.inst * a.tiles_per_inst;
    int tj = a.tj0;
    for (;; ++tj) {
        const int64_t lastcol = ((int64_t)(tj + 1) * BN < a.n ? (int64_t)(tj + 1) * BN : a.n) - 1;
        const int64_t cnt = lastcol / ROWS + 1;
        if (tl < cnt) break;
        tl -= cnt;
    }
    w.ti = (int)tl; w.tj = tj;
    return w;
}

// instruction descriptor: D = s32 (bits 4-5 = 2), A and B signed int8 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at 17, M >> 4 at 24
__device__ __forceinline__ constexpr uint32_t i8_idesc(int n, int m = kI8BM) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// this launch (NS planes) owns the instance?
template <int NS>
__device__ __forceinline__ bool i8_runs(const GramI8KernelArgs& a, int64_t inst) {
    return a.force_ns ? a.force_ns == NS : gram_i8_level(a.stats, inst) == NS;
}

template <int NS, int BN, int CTAS, int KC>
__global__ void __launch_bounds__(kI8Threads, 1) gram_i8_kernel(const __grid_constant__ GramI8KernelArgs a) {
    using C = I8Cfg<NS, BN, CTAS, KC>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[C::kStages];
    __shared__ __align__(8) uint64_t empty_bar[C::kStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ __align__(8) uint64_t full_a[8], empty_a[8];   // wide stages: row-panel slots
    __shared__ uint32_t tmem_slot;
    __shared__ double sfj[BN];       // sigma_j 2^(-4 NS - 3) of the tile's columns
    __shared__ int32_t mapj[BN];     // output index of the tile's columns
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ring = (i8::s_u32(smem_raw) + 1023u) & ~1023u;
    const int nkc = (int)(a.kp / KC);
    if (a.total_tiles == a.tiles_per_inst && !i8_runs<NS>(a, 0)) return;   // one instance and it is not ours: leave before touching TMEM
    // CTA pair: rank in the cluster; the leader (rank 0) owns the full / tmem_empty barriers and issues the MMAs for both CTAs
    const uint32_t rank = CTAS == 2 ? i8::cluster_ctarank() : 0u;
    const int64_t unit0 = CTAS == 2 ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
    const int64_t ustride = CTAS == 2 ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < C::kStages; ++s) { i8::mb_init(&full_bar[s], CTAS); i8::mb_init(&empty_bar[s], 1); }
#pragma unroll
        for (int s = 0; s < C::kNA; ++s) { i8::mb_init(&full_a[s], 1); i8::mb_init(&empty_a[s], 1); }
#pragma unroll
        for (int b = 0; b < 2; ++b) { i8::mb_init(&tmem_full_bar[b], 1); i8::mb_init(&tmem_empty_bar[b], kI8EpiWarps * CTAS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CTAS == 2) i8::tmem_alloc_pair(&tmem_slot, kI8TmemCols); else i8::tmem_alloc(&tmem_slot, kI8TmemCols);
    }
    i8::tc_fence_before();
    if constexpr (CTAS == 2) i8::cluster_sync_all(); else __syncthreads();   // the peer's barriers are initialised before anything arrives on them
    i8::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            [[maybe_unused]] uint32_t ita = 0;
            for (int64_t u = unit0; u < a.total_tiles; u += ustride) {
                const I8Unit w = i8_decode<BN, C::kRows>(a, u);
                if (!i8_runs<NS>(a, w.inst)) continue;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % C::kStages;
                    i8::mb_wait(&empty_bar[s], ((it / C::kStages) & 1) ^ 1);
                    const uint32_t stage = ring + (uint32_t)s * C::kStageBytes;
                    if constexpr (C::kWide) {
                        // column panel: all NS planes in one box (they are consecutive in shared memory in box order k, column, plane)
                        if (a.dbg & 1) i8::mb_arrive(&full_bar[s]);
                        else {
                            i8::mb_expect_tx(&full_bar[s], C::kBStage);
                            i8::tma_load_4d(stage, &a.tmja, kc * KC, w.tj * BN, 0, (int)w.inst, &full_bar[s]);
                        }
                        // row panel: one slot per plane, in the order the MMA issuer walks them
#pragma unroll 1
                        for (int sl = 0; sl < NS; ++sl, ++ita) {
                            const int sa = ita % C::kNA;
                            i8::mb_wait(&empty_a[sa], ((ita / C::kNA) & 1) ^ 1);
                            if (a.dbg & 1) { i8::mb_arrive(&full_a[sa]); continue; }
                            i8::mb_expect_tx(&full_a[sa], C::kITile);
                            i8::tma_load_4d(ring + C::kStages * C::kStageBytes + sa * C::kITile, &a.tmi, kc * KC, w.ti * kI8BM, sl, (int)w.inst, &full_a[sa]);
                        }
                    } else if constexpr (CTAS == 2) {
                        // both CTAs count their bytes on the LEADER's full barrier (2 arrivals + the bytes of both)
                        const uint32_t fb = i8::mapa_u32(i8::s_u32(&full_bar[s]), 0);
                        if (a.dbg & 1) { i8::mb_arrive_cluster(fb); continue; }   // timing diagnostics
                        i8::mb_expect_tx_cluster(fb, C::kTxBytes);
                        const int row0 = w.ti * C::kRows + (int)rank * kI8BM;
#pragma unroll
                        for (int sl = 0; sl < NS; ++sl) i8::tma_load_4d_pair(stage + sl * C::kITile, &a.tmi, kc * KC, row0, sl, (int)w.inst, fb);
#pragma unroll
                        for (int ps = 0; ps < C::kNPair; ++ps)   // pair slot ps: plane 2 ps + rank, all BN columns
                            i8::tma_load_4d_pair(stage + NS * C::kITile + ps * C::kJTile, &a.tmj, kc * KC, w.tj * BN, 2 * ps + (int)rank, (int)w.inst, fb);
#pragma unroll
                        for (int hs = 0; hs < C::kNHalf; ++hs)   // half slot hs: plane 2 hs, columns BN/2 rank ..
                            i8::tma_load_4d_pair(stage + NS * C::kITile + C::kNPair * C::kJTile + hs * C::kHalfTile, &a.tmh, kc * KC,
                                                 w.tj * BN + (int)rank * (BN / 2), 2 * hs, (int)w.inst, fb);
                    } else {
                        if (a.dbg & 1) { i8::mb_arrive(&full_bar[s]); continue; }   // timing diagnostics
                        i8::mb_expect_tx(&full_bar[s], C::kStageBytes);
#pragma unroll
                        for (int sl = 0; sl < NS; ++sl)
                            i8::tma_load_4d(stage + sl * C::kITile, &a.tmi, kc * KC, w.ti * kI8BM, sl, (int)w.inst, &full_bar[s]);
#pragma unroll
                        for (int sl = 0; sl < NS; ++sl)
                            i8::tma_load_4d(stage + NS * C::kITile + sl * C::kJTile, &a.tmj, kc * KC, w.tj * BN, sl, (int)w.inst, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: all digit-plane pairs with s + t <= NS - 1, level s + t -> TMEM columns [(s+t) BN, +BN).  The WHOLE warp runs the
        // loop with warp-uniform values (descriptors in uniform registers), one elected lane issues (pmb_i8_ptx.h) =====
        if (rank == 0) {
            uint32_t it = 0, tile_it = 0;
            [[maybe_unused]] uint32_t ita = 0;
            for (int64_t u = unit0; u < a.total_tiles; u += ustride) {
                const I8Unit w = i8_decode<BN, C::kRows>(a, u);
                if (!i8_runs<NS>(a, w.inst)) continue;
                const uint32_t ab = tile_it % C::kAccBufs, ause = tile_it / C::kAccBufs;   // accumulator set and how often it was used
                i8::mb_wait(&tmem_empty_bar[ab], (ause & 1) ^ 1);   // the epilogue has drained this accumulator set
                i8::tc_fence_after();
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % C::kStages;
                    i8::mb_wait(&full_bar[s], (it / C::kStages) & 1);
                    i8::tc_fence_after();
                    const uint32_t stage = ring + (uint32_t)s * C::kStageBytes;
                    if constexpr (C::kWide) {
#pragma unroll
                        for (int si = 0; si < NS; ++si, ++ita) {
                            const int sa = ita % C::kNA;
                            i8::mb_wait(&full_a[sa], (ita / C::kNA) & 1);
                            i8::tc_fence_after();
                            const uint32_t apanel = ring + C::kStages * C::kStageBytes + sa * C::kITile;
                            const int W = (NS - si) * BN, nm = (W + 255) / 256, per = ((W + nm - 1) / nm + 15) / 16 * 16;
                            if (!((a.dbg & 2) || ((a.dbg & 4) && si > 0))) {
#pragma unroll
                                for (int kk = 0; kk < KC / 32; ++kk) {
                                    const uint64_t ad = i8::smem_desc_sw64<KC>(apanel + kk * 32);
#pragma unroll
                                    for (int c = 0; c < W; c += per) {
                                        const int wn = (W - c) < per ? (W - c) : per;
                                        const uint64_t bd = i8::smem_desc_sw64<KC>(stage + c * KC + kk * 32);
                                        i8::mma_i8_elect(tmem_base + (uint32_t)(ab * NS * BN + si * BN + c), ad, bd, i8_idesc(wn), (kc > 0 || kk > 0 || si > 0) ? 1u : 0u);
                                    }
                                }
                            }
                            i8::mma_commit_elect(&empty_a[sa]);   // this row-panel slot may be refilled
                        }
                        i8::mma_commit_elect(&empty_bar[s]);
                    } else {
#pragma unroll
                    for (int kk = 0; kk < KC / 32; ++kk) {
                        if constexpr (CTAS == 2) {
                            // CTA pair: plane s of the 256-row panel against planes (t, t+1) of the column panel (N = 2 BN, one plane from
                            // each CTA) for t = 0, 2, ..; a left-over single plane t = NS-1-s (N = BN, half of its columns from each CTA)
                            // always lands on the last level
#pragma unroll
                            for (int si = 0; si < NS; ++si) {
                                if ((a.dbg & 2) || ((a.dbg & 4) && si > 0)) continue;
                                const uint64_t ad = i8::smem_desc_sw64<KC>(stage + si * C::kITile + kk * 32);
                                const uint32_t first = (kc > 0 || kk > 0 || si > 0) ? 1u : 0u;
#pragma unroll
                                for (int t0 = 0; t0 + 1 < NS - si; t0 += 2) {
                                    const uint64_t bd = i8::smem_desc_sw64<KC>(stage + NS * C::kITile + (t0 / 2) * C::kJTile + kk * 32);
                                    i8::mma_i8_pair_elect(tmem_base + (uint32_t)(ab * NS * BN + (si + t0) * BN), ad, bd, i8_idesc(2 * BN, 256), first);
                                }
                                if ((NS - si) & 1) {
                                    const uint64_t bd = i8::smem_desc_sw64<KC>(stage + NS * C::kITile + C::kNPair * C::kJTile + ((NS - 1 - si) / 2) * C::kHalfTile + kk * 32);
                                    i8::mma_i8_pair_elect(tmem_base + (uint32_t)(ab * NS * BN + (NS - 1) * BN), ad, bd, i8_idesc(BN, 256), first);
                                }
                            }
                        } else {
                        // plane s of the row panel against GRP consecutive planes t .. t+g-1 of the column panel in ONE instruction
                        // (N = g BN): the planes are contiguous in shared memory and the levels s+t .. s+t+g-1 are adjacent TMEM column
                        // blocks, so the row panel is read from shared memory 9 times per k step instead of 15 (21 for 6 planes).
                        // Every level is first written by the s = 0 instructions (accumulate off at the first k step).
                        if (a.dbg & 16) {
                            for (int rep3 = 0; rep3 < 3; ++rep3) {
                                const uint64_t ad = i8::smem_desc_sw64<KC>(stage + rep3 * C::kITile + kk * 32);
#pragma unroll
                                for (int c = 0; c < NS * BN; c += NS * BN / 2) {
                                    const uint64_t bd = i8::smem_desc_sw64<KC>(stage + NS * C::kITile + c * KC + kk * 32);
                                    i8::mma_i8_elect(tmem_base + (uint32_t)(ab * NS * BN + c), ad, bd, i8_idesc(NS * BN / 2), (kc > 0 || kk > 0 || rep3 > 0) ? 1u : 0u);
                                }
                            }
                        } else
#pragma unroll
                        for (int si = 0; si < NS; ++si) {
                            if ((a.dbg & 2) || ((a.dbg & 4) && si > 0)) continue;
                            const uint64_t ad = i8::smem_desc_sw64<KC>(stage + si * C::kITile + kk * 32);
                            // the (NS - si) BN columns of this row are cut into the fewest instructions of N <= 256 (N % 16 == 0; a cut need
                            // not fall on a plane boundary: 5 planes of 96 = 240 + 240), so the row panel is read 8 times per k step for
                            // 5 planes (9 for 6)
                            const int W = (NS - si) * BN, nm = (W + 255) / 256;
                            const int per = ((W + nm - 1) / nm + 15) / 16 * 16;
#pragma unroll
                            for (int c = 0; c < W; c += per) {
                                const int wn = (W - c) < per ? (W - c) : per;
                                const uint64_t bd = i8::smem_desc_sw64<KC>(stage + NS * C::kITile + c * KC + kk * 32);
                                i8::mma_i8_elect(tmem_base + (uint32_t)(ab * NS * BN + si * BN + c), ad, bd, i8_idesc(wn), (kc > 0 || kk > 0 || si > 0) ? 1u : 0u);
                            }
                        }
                        }
                    }
                    if constexpr (CTAS == 2) i8::mma_commit_pair_elect(&empty_bar[s]); else i8::mma_commit_elect(&empty_bar[s]);   // the stage may be refilled once these MMAs have read it
                    }
                }
                if constexpr (CTAS == 2) i8::mma_commit_pair_elect(&tmem_full_bar[ab]); else i8::mma_commit_elect(&tmem_full_bar[ab]);
                ++tile_it;
            }
        }
    } else {
        // ===== epilogue: 8 warps, warp w owns TMEM lanes 32 (w % 4) .. +31 = tile rows and every other 16-column chunk =====
        const int q = warp & 3;
        const int et = tid - 64;   // 0 .. 32 * kI8EpiWarps - 1
        const int chalf = (warp - 2) >> 2;   // which of the quarter's two warps: it takes the 16-column chunks of its parity
        uint32_t tile_it = 0;
        for (int64_t u = unit0; u < a.total_tiles; u += ustride) {
            const I8Unit w = i8_decode<BN, C::kRows>(a, u);
            if (!i8_runs<NS>(a, w.inst)) continue;
            const int64_t i0 = (int64_t)w.ti * C::kRows + (int64_t)rank * kI8BM, j0 = (int64_t)w.tj * BN;
            const double* sg = a.sigma + w.inst * a.n;
            if (et < BN) {
                const int64_t gj = j0 + et;
                sfj[et] = gj < a.n ? sg[gj] * exp2(-4.0 * NS - 3.0) : 0.0;
                mapj[et] = gj < a.n ? (a.map ? a.map[gj] : (int32_t)gj) : 0;
            }
            const int m = q * 32 + lane;
            const int64_t gi = i0 + m;
            const bool row_ok = gi < a.n;
            const double sc = a.scale.ptr ? a.scale.c * a.scale.ptr[w.inst * a.scale.bstride] : a.scale.c;
            const double fi = row_ok ? sc * (sg[gi] * exp2(-4.0 * NS - 3.0)) : 0.0;
            const int64_t om = row_ok ? (a.map ? a.map[gi] : gi) : 0;
            double* out = a.out + w.inst * a.o_bstride;
            double* pk = a.packed_dev ? a.packed_dev + w.inst * a.packed_bstride : nullptr;
            i8::epi_bar();   // sfj / mapj visible
            const uint32_t ab = tile_it % C::kAccBufs, ause = tile_it / C::kAccBufs;
            i8::mb_wait(&tmem_full_bar[ab], ause & 1);
            i8::tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + ab * (uint32_t)(NS * BN);
            // identity map, overwrite, aligned output: a 16-column chunk that lies strictly above the diagonal for this thread's row and
            // inside the matrix is stored without predicates -- the direct triangle coalesces across the lanes (consecutive rows), the
            // mirrored one is 16 consecutive doubles per thread (4 x 256-bit stores).  Decided per (row, chunk), so tiles that straddle
            // the diagonal or the ragged edge (all tiles of a 256 x 256 instance) still take it for most of their chunks.
            const bool vec_ok = !a.map && !a.accumulate && (a.ldo & 3) == 0 && (a.o_bstride & 3) == 0 && (((uintptr_t)a.out) & 31) == 0;
            // phase 1: TMEM -> registers (Horner chain) for all of this thread's chunks, then hand the accumulators back: the MMAs of the
            // CTA's next tile run while phase 2 scales and stores (cfg2: 0.351 -> 0.337 ms).  An exact 64-bit integer combination of the
            // levels (three IMAD.WIDE + one conversion per entry instead of the fp64 chain) was measured and is no faster: what the tile
            // boundary costs (~5 us, tools/probe/i8_dbg.sh bit 8) is the latency chain commit -> tmem_full -> tcgen05.ld -> tmem_empty,
            // not fp64 throughput
            constexpr int NCH = (BN / 16 + 1) / 2;
            double v[NCH][16];
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = 16 * chalf + 32 * ci;
                if (c0 < BN && !(a.dbg & 8)) {
                    int32_t acc[NS][16];
#pragma unroll
                    for (int l = 0; l < NS; ++l) i8::tmem_ld16(trow + (uint32_t)(l * BN + c0), acc[l]);
                    i8::tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        double t = (double)acc[0][jj];
#pragma unroll
                        for (int l = 1; l < NS; ++l) t = fma(t, 256.0, (double)acc[l][jj]);
                        v[ci][jj] = t;
                    }
                }
            }
            i8::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTAS == 2) i8::mb_arrive_cluster(i8::mapa_u32(i8::s_u32(&tmem_empty_bar[ab]), 0)); else i8::mb_arrive(&tmem_empty_bar[ab]);
            }
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = 16 * chalf + 32 * ci;
                if (c0 >= BN || (a.dbg & 8)) continue;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    v[ci][jj] = v[ci][jj] * fi * sfj[c0 + jj];
                }
                if (vec_ok && row_ok && gi < j0 + c0 && j0 + c0 + 16 <= a.n) {
                    double* p1 = out + gi + (j0 + c0) * a.ldo;
#pragma unroll
                    // streaming stores (evict-first): Q is written once and not re-read here, the digit planes (100 MB) should keep the L2
                    for (int jj = 0; jj < 16; ++jj) __stcs(p1 + (int64_t)jj * a.ldo, v[ci][jj]);
                    double* p2 = out + (j0 + c0) + gi * a.ldo;
#pragma unroll
                    for (int jj = 0; jj < 16; jj += 4)
                        asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p2 + jj), "d"(v[ci][jj]), "d"(v[ci][jj + 1]), "d"(v[ci][jj + 2]), "d"(v[ci][jj + 3]) : "memory");
                    if (pk) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            const int64_t gj = j0 + c0 + jj;
                            __stcs(pk + gj * (gj + 1) / 2 + gi, v[ci][jj]);
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int64_t gj = j0 + c0 + jj;
                    if (row_ok && gj < a.n && gi <= gj) {   // every unordered pair of columns exactly once
                        const int64_t on = mapj[c0 + jj];
                        double* p1 = out + om + on * a.ldo;
                        if (a.accumulate) *p1 += v[ci][jj]; else *p1 = v[ci][jj];
                        if (gi != gj) {
                            double* p2 = out + on + om * a.ldo;
                            if (a.accumulate) *p2 += v[ci][jj]; else *p2 = v[ci][jj];
                        }
                        if (pk) {
                            const int64_t lo = om < on ? om : on, hi = om < on ? on : om;
                            pk[hi * (hi + 1) / 2 + lo] = v[ci][jj];
                        }
                    }
                }
            }
            i8::epi_bar();   // sfj / mapj are rewritten for the next tile
            ++tile_it;
        }
    }
    i8::tc_fence_before();
    if constexpr (CTAS == 2) i8::cluster_sync_all(); else __syncthreads();   // nobody leaves (or frees TMEM) while the peer can still signal into this CTA
    if (warp == 1) {
        i8::tc_fence_after();
        if constexpr (CTAS == 2) i8::tmem_dealloc_pair(tmem_base, kI8TmemCols); else i8::tmem_dealloc(tmem_base, kI8TmemCols);
    }
}

// ---- host ----------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnI8)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnI8 encode_fn_i8() {
    static EncodeTiledFnI8 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFnI8)p;
    }
    return fn;
}

static bool make_plane_map(CUtensorMap* tm, const int8_t* base, int64_t kp, int64_t n, int64_t batch, int box_rows, int kc = kI8KC, int box_planes = 1) {
    EncodeTiledFnI8 enc = encode_fn_i8();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)kp, (cuuint64_t)n, (cuuint64_t)kGramI8Planes, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)kp, (cuuint64_t)kp * n, (cuuint64_t)kp * n * kGramI8Planes};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)box_rows, (cuuint32_t)box_planes, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<int8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     kc == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : kc == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

GramI8Plan plan_gram_i8(int64_t k, int64_t n, int64_t batch) {
    GramI8Plan p{};
    if (!encode_fn_i8() || k <= 0 || n <= 0) return p;
    if (k * kGramI8Planes * 16384 >= (int64_t)1 << 31) return p;   // int32 accumulators: (s + t + 1) * k * 128 * 128 must stay below 2^31
    if (n >= ((int64_t)1 << 31) / 2 || batch >= ((int64_t)1 << 31) || batch * n >= ((int64_t)1 << 31)) return p;   // 32-bit column ids in the slice kernel
    p.ok = 1;
    p.kp = (k + 127) / 128 * 128;
    p.slice_bytes = batch * kGramI8Planes * n * p.kp;
    return p;
}

void configure_gram_i8_kernels() {
    cudaFuncSetAttribute(gram_i8_kernel<5, 96, 1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<5, 96, 1, 64>::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_kernel<6, 80, 1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<6, 80, 1, 64>::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_kernel<5, 48, 1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<5, 48, 1, 64>::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_kernel<5, 96, 2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<5, 96, 2, 64>::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_kernel<6, 80, 2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<6, 80, 2, 64>::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_kernel<5, 96, 1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<5, 96, 1, 128>::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_kernel<6, 80, 1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<6, 80, 1, 128>::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_slice_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SliceRing * 4096 * 8 + 64);
    cudaFuncSetAttribute(gram_i8_slice_kernel<256, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SliceRing * 4096 * 8 + 64);
}

// the generator-fused slice pass handles whole, densely stored columns of 1024 .. 4096 rows (the full-column fast path of the slice kernel)
bool gram_i8_slice_can_generate(const GramArgs& g, const GramI8Plan& plan) {
    const bool aligned = g.ldx % 2 == 0 && (((uintptr_t)g.X) & 15) == 0 && g.x_bstride % 2 == 0;
    const bool galigned = !g.gvec || ((((uintptr_t)g.gvec) & 15) == 0 && g.g_bstride % 2 == 0);
    return plan.ok && plan.kp == g.k && g.k % 1024 == 0 && g.k >= 1024 && g.k <= 4096 && g.ldx == g.k && aligned && galigned && !g.gate && !g.packed_host;
}

bool launch_gram_i8_slice(const GramArgs& g, const GramI8Scratch& sc, int64_t batch, int sm_count, cudaStream_t stream, int64_t col0, int64_t col1) {
    if (!sc.plan.ok || !sc.slices || g.packed_host || g.gate) return false;
    if (col1 < 0) col1 = g.n;
    if (col0 == 0) cudaMemsetAsync(sc.stats, 0, sizeof(unsigned long long) * kGramI8StatsStride * (size_t)batch, stream);   // first (or only) chunk
    GramI8SliceArgs s{};
    s.col0 = col0; s.ncols = col1 - col0;
    s.X = g.X; s.ldx = g.ldx; s.x_bstride = g.x_bstride; s.k = g.k; s.kp = sc.plan.kp; s.n = g.n; s.batch = batch;
    s.slices = sc.slices; s.sigma = sc.sigma; s.stats = sc.stats; s.tol = sc.tol;
    if (g.gvec) {
        s.gvec = g.gvec; s.g_bstride = g.g_bstride; s.qout = g.qout; s.q_bstride = g.q_bstride; s.c0out = g.c0out; s.map = g.map;
        s.qscale = g.qscale; s.cscale = g.cscale;
    }
    const int64_t cols = batch * (col1 - col0);
    const bool aligned = g.ldx % 2 == 0 && (((uintptr_t)g.X) & 15) == 0 && g.x_bstride % 2 == 0;
    if (sc.gen_counter) {   // generator-fused pass: the planner removed the operand's fill from the prologue, so this path MUST run
        if (!gram_i8_slice_can_generate(g, sc.plan) || col0 != 0 || col1 != g.n) return false;
        s.gen_seed = sc.gen_seed; s.gen_sc53 = sc.gen_scale * (1.0 / 9007199254740992.0); s.gen_counter = sc.gen_counter; s.gen_inst_offset = sc.gen_inst_offset; s.gen_writeback = sc.gen_writeback;
        const int sgrid = (int)std::min<int64_t>(cols, (int64_t)sm_count * PMB_SLICE_CTAS);
        gram_i8_slice_kernel<256, false, true, true><<<sgrid, 256, 0, stream>>>(s);
        return true;
    }
    if (sc.plan.kp > 1024 && sc.plan.kp <= 4096 && aligned && !getenv("PMB_I8_NO_RING")) {   // bulk-copy ring, 2 CTAs per SM
        const int sgrid = (int)std::min<int64_t>(cols, (int64_t)sm_count * PMB_SLICE_CTAS);
        const size_t smem = (size_t)kI8SliceRing * (size_t)((g.k + 1) / 2 * 2) * 8 + 64;
        const bool galigned = !g.gvec || ((((uintptr_t)g.gvec) & 15) == 0 && g.g_bstride % 2 == 0);
        if (sc.plan.kp == g.k && g.k % 1024 == 0 && galigned && !getenv("PMB_I8_SLICE_GENERIC")) gram_i8_slice_kernel<256, true, true><<<sgrid, 256, smem, stream>>>(s);
        else gram_i8_slice_kernel<256, true><<<sgrid, 256, smem, stream>>>(s);
    } else if (sc.plan.kp > 1024) {   // long columns: the whole CTA owns one column (up to 4096 rows cached in registers)
        const int sgrid = (int)std::min<int64_t>(cols, (int64_t)sm_count * 4);
        gram_i8_slice_kernel<256, false><<<sgrid, 256, 0, stream>>>(s);
    } else if (sc.plan.kp <= 128) {   // short columns: 8 / 16 lanes per column, 32 / 16 columns per CTA pass (16 rows per lane, cached)
        const int sgrid = (int)std::min<int64_t>((cols + 31) / 32, (int64_t)sm_count * 4);
        gram_i8_slice_kernel<8, false><<<sgrid, 256, 0, stream>>>(s);
    } else if (sc.plan.kp <= 256) {
        const int sgrid = (int)std::min<int64_t>((cols + 15) / 16, (int64_t)sm_count * 4);
        gram_i8_slice_kernel<16, false><<<sgrid, 256, 0, stream>>>(s);
    } else {
        const int sgrid = (int)std::min<int64_t>((cols + 7) / 8, (int64_t)sm_count * 4);
        gram_i8_slice_kernel<32, false><<<sgrid, 256, 0, stream>>>(s);
    }
    return true;
}

template <int NS, int BN, int CTAS = 1, int KC = kI8KC>
static bool launch_i8_impl(const GramArgs& g, const GramI8Scratch& sc, int64_t batch, int sm_count, cudaStream_t stream, int64_t col0, int64_t col1) {
    using C = I8Cfg<NS, BN, CTAS, KC>;
    GramI8KernelArgs a{};
    if (!make_plane_map(&a.tmi, sc.slices, sc.plan.kp, g.n, batch, kI8BM, KC)) return false;
    if (!make_plane_map(&a.tmj, sc.slices, sc.plan.kp, g.n, batch, BN, KC)) return false;
    if constexpr (CTAS == 2) { if (!make_plane_map(&a.tmh, sc.slices, sc.plan.kp, g.n, batch, BN / 2, KC)) return false; }
    if constexpr (C::kWide) { if (!make_plane_map(&a.tmja, sc.slices, sc.plan.kp, g.n, batch, BN, KC, NS)) return false; }
    a.k = g.k; a.kp = sc.plan.kp; a.n = g.n; a.sigma = sc.sigma; a.stats = sc.stats; a.tol = sc.tol; a.force_ns = sc.force_ns;
    a.out = g.H; a.ldo = g.ldh; a.o_bstride = g.h_bstride; a.map = g.map; a.scale = g.scale; a.accumulate = g.accumulate;
    a.packed_dev = g.accumulate ? nullptr : g.packed_dev; a.packed_bstride = g.packed_bstride;
    // tile columns whose LAST column lies in [col0, col1): every tile column belongs to exactly one column range
    const int ntj = (int)((g.n + BN - 1) / BN);
    auto lastcol = [&](int tj) { return std::min<int64_t>((int64_t)(tj + 1) * BN, g.n) - 1; };
    a.tj0 = 0;
    while (a.tj0 < ntj && lastcol(a.tj0) < col0) ++a.tj0;
    a.tj1 = a.tj0;
    while (a.tj1 < ntj && lastcol(a.tj1) < col1) ++a.tj1;
    int64_t tpi = 0;
    for (int tj = a.tj0; tj < a.tj1; ++tj) tpi += lastcol(tj) / C::kRows + 1;
    a.tiles_per_inst = tpi;
    a.total_tiles = tpi * batch;
    if (const char* e = getenv("PMB_I8_DBG")) a.dbg = atoi(e);
    if (a.total_tiles == 0) return true;
    if constexpr (CTAS == 2) {   // one cluster of two CTAs (a TPC) per tile
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * std::min<int64_t>(a.total_tiles, sm_count / 2)));
        cfg.blockDim = dim3(kI8Threads);
        cfg.dynamicSmemBytes = C::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, gram_i8_kernel<NS, BN, 2, KC>, a) == cudaSuccess;
    } else {
        const int grid = (int)std::min<int64_t>(a.total_tiles, sm_count);
        gram_i8_kernel<NS, BN, 1, KC><<<grid, kI8Threads, C::kSmemBytes, stream>>>(a);
        return true;
    }
}

__global__ void gram_i8_count_levels_kernel(const unsigned long long* stats, double tol, int force_ns, int64_t batch, unsigned long long* counts) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < batch; i += (int64_t)gridDim.x * blockDim.x) {
        const int lv = force_ns ? force_ns : gram_i8_level(stats, i);
        atomicAdd(counts + (lv == 5 ? 0 : lv == 6 ? 1 : 2), 1ull);   // anything else (kGramI8Fp64): the fp64 kernel
    }
}

void launch_gram_i8_count_levels(const GramI8Scratch& sc, int64_t batch, unsigned long long* counts_dev, cudaStream_t stream) {
    const int grid = (int)std::min<int64_t>((batch + 255) / 256, 1024);
    gram_i8_count_levels_kernel<<<grid, 256, 0, stream>>>(sc.stats, sc.tol, sc.force_ns, batch, counts_dev);
}

// Cost model (seconds) of one SYRK step, fitted to measurements on B200 (profiles/r01_n_*): the int8 path pays a fixed price
// (slice pass over X, three launches, ~12 us of pipeline fill + epilogue per tile) that only large contractions amortise.
bool gram_i8_pays_off(int64_t k, int64_t n, int64_t batch, int sm_count) {
    if (const char* e = getenv("PMB_GRAM_I8_ALWAYS")) { if (atoi(e)) return true; }
    int64_t tpi = 0;
    for (int64_t tj = 0; tj * 96 < n; ++tj) tpi += (std::min<int64_t>((tj + 1) * 96, n) - 1) / kI8BM + 1;
    const double tiles = (double)tpi * (double)batch;
    const double waves = std::ceil(tiles / sm_count);
    const double kp = (double)((k + 127) / 128 * 128);
    const double t_i8 = waves * (kp / 32.0 * 15.0 * 48.0 / 1.45e9 * 1.15 + 12e-6) + (double)batch * n * (16.0 * k + 6.0 * kp) / 4.0e12 + 15e-6;
    const double t_f64 = (double)batch * k * n * (n + 1) / 30e12 + 6e-6;
    return t_i8 < t_f64;
}

bool launch_gram_i8_mma(const GramArgs& g, const GramI8Scratch& sc, int64_t batch, int sm_count, cudaStream_t stream, int64_t col0, int64_t col1) {
    if (!sc.plan.ok || !sc.slices || g.packed_host || g.gate) return false;
    if (col1 < 0) col1 = g.n;
    bool ok = true;
    // experiment (PMB_I8_NARROW=1): 48-wide tiles keep TWO accumulator sets in TMEM, so the epilogue of a tile overlaps the MMAs of the
    // next, and multiply plane s against up to five column-panel planes per instruction.  Measured slower than the 96-wide tiles even
    // for short contractions (1024 x (256 x 256): 0.55 vs 0.40 ms; n = 4096, k = 512: 0.109 vs 0.083 ms): twice the tiles, twice the
    // row-panel traffic.  Kept for the accumulator double-buffering it exercises.
    const bool narrow = sc.plan.kp <= 512 && getenv("PMB_I8_NARROW") != nullptr;
    // Defaults, measured on B200 at n = k = 4096 (tools/probe/i8_wide.sh, i8_pair.sh; profiles/r01_o_gram_i8_variants.txt):
    //   5 planes: 64-byte stages, 3-stage ring (0.336 ms; 128-byte split ring 0.344 ms; CTA pair 0.49-0.51 ms)
    //   6 planes: 128-byte split ring (0.469 ms; 64-byte stages fit only 2 stages of 78 KB: 0.498 ms; CTA pair 0.62-0.65 ms)
    // PMB_I8_KC = 64 | 128 and PMB_I8_PAIR = 1 override for experiments.
    static const int kc_env = getenv("PMB_I8_KC") ? atoi(getenv("PMB_I8_KC")) : 0;
    static const bool pair = getenv("PMB_I8_PAIR") != nullptr && atoi(getenv("PMB_I8_PAIR")) != 0;
    if (pair) {
        if (sc.force_ns != 6) ok = ok && launch_i8_impl<5, 96, 2, 64>(g, sc, batch, sm_count, stream, col0, col1);
        if (sc.force_ns != 5) ok = ok && launch_i8_impl<6, 80, 2, 64>(g, sc, batch, sm_count, stream, col0, col1);
        return ok;
    }
    if (!narrow) {
        if (sc.force_ns != 6) ok = ok && (kc_env == 128 ? launch_i8_impl<5, 96, 1, 128>(g, sc, batch, sm_count, stream, col0, col1)
                                                        : launch_i8_impl<5, 96, 1, 64>(g, sc, batch, sm_count, stream, col0, col1));
        if (sc.force_ns != 5) ok = ok && (kc_env == 64 ? launch_i8_impl<6, 80, 1, 64>(g, sc, batch, sm_count, stream, col0, col1)
                                                       : launch_i8_impl<6, 80, 1, 128>(g, sc, batch, sm_count, stream, col0, col1));
        return ok;
    }
    if (sc.force_ns != 6) ok = ok && launch_i8_impl<5, 48, 1, 64>(g, sc, batch, sm_count, stream, col0, col1);
    if (sc.force_ns != 5) ok = ok && launch_i8_impl<6, 80, 1, 128>(g, sc, batch, sm_count, stream, col0, col1);
    return ok;
}

// flag[0] = 1 when any instance ended up on the fp64 path (the pipelined e2e graph has no fp64 launch: the host repairs)
__global__ void gram_i8_flag_kernel(const unsigned long long* stats, int64_t batch, int32_t* flag) {
    int bad = 0;
    for (int64_t i = threadIdx.x; i < batch; i += blockDim.x) bad |= gram_i8_level(stats, i) != 5 && gram_i8_level(stats, i) != 6;
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) flag[0] = bad;
}
void launch_gram_i8_flag(const GramI8Scratch& sc, int64_t batch, int32_t* flag_dev, cudaStream_t stream) {
    gram_i8_flag_kernel<<<1, 256, 0, stream>>>(sc.stats, batch, flag_dev);
}

}  // namespace pmb
