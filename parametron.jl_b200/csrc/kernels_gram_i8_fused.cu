// K3 -- fused small-instance Gram on tcgen05: ONE launch per batch of independent instances with n <= 256 variables and k <= 256
// residual rows (BASELINE cfg3: 8192 x (n = 256, k = 256)).  Same reference semantics as kernels_gram_i8.cu / kernels_gram_tma.cu
// (vecdot! + muladd!, functions.jl:548-576,702-709, then the symmetric MOI Hessian of moi_interop.jl:45-62:  H = scale * X'X, plus the
// linear part q += qscale * X'g, c0 += cscale * g.g of the least-squares objective) and the SAME arithmetic as the two-kernel int8
// route (47-bit fixed point per column, balanced base-256 digits, exact int32 accumulation of the digit-plane products with
// s + t <= 4, one fp64 rounding) -- so results are bit-identical to it (and to the numpy restatement the tests check it against).
//
// Why a separate kernel: at this size the two-kernel route moves every digit plane through HBM (write 6, read 5 bytes per element
// on top of the 8 B read of X and the 8 B write of H) and the fp64 DMMA route is compute-bound at ~4 ms per 8192 instances.  Here a
// CTA PAIR (cluster of 2 = one TPC, tcgen05 cta_group::2) owns one instance at a time and keeps everything on chip:
//
//   warp 0   producer   X columns of the CTA's half (128 columns) travel global -> shared as 1-D bulk copies (UBLKCP), 8 columns per
//                       stage, 3 stages ahead; the offset vector g rides along with the first stage of an instance
//   warps 2-9 workers   (a) one column per warp and stage: largest magnitude / 2-norm / 1-norm / X'g by warp shuffles, fixed-point
//                       digits, five int8 planes written straight into the 128-byte-swizzled K-major operand layout tcgen05 reads
//                       (160 KB per CTA: 5 planes x 2 k-blocks x 128 columns x 128 B); the per-instance statistics of the error
//                       certificate are exchanged between the two CTAs through distributed shared memory
//                       (b) epilogue: tcgen05.ld, exact combination of the five int32 levels, sigma_i sigma_j scaling, stores
//   warp 1   MMA issuer (leader CTA only) tcgen05.mma.cta_group::2.kind::i8, M = 256 (CTA r supplies rows 128 r ..), N = 128 (each CTA
//                       supplies 64 of its columns), K = 32: the 15 digit-plane products of an output half go out as three units
//                       {level 0, 1}, {2, 3}, {4} of at most 256 TMEM columns, double-buffered against the epilogue
//
// HBM traffic per instance is the algorithmic minimum (read X once, write H once); the full square is computed (no mirrored writes).
// The certificate is evaluated per instance inside the kernel (5 planes or "not certified"); uncertified instances are left to the
// fp64 kernel that follows in the update graph (GramArgs::cert_stats), exactly as on the two-kernel route.
#include <cuda.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <math_constants.h>

#include "pmb_i8_ptx.h"
#include "pmb_kernels.h"

namespace pmb {

namespace fz {
constexpr int kNS = 5;                       // digit planes multiplied (the certificate accepts 5 or hands the instance to fp64)
constexpr int kCols = 128;                   // X columns per CTA = rows of its operand panels = TMEM lanes
constexpr int kKBlock = 128;                 // bytes of k per swizzle span
constexpr int kPanel = kCols * kKBlock;      // 16 KB: one plane, one k-block
constexpr int kPlaneBytes = 2 * kPanel;      // k <= 256
constexpr int kPlanesBytes = kNS * kPlaneBytes;          // 160 KB
constexpr int kStageCols = 8;                // one column per warp of the worker group that owns the stage
constexpr int kStages = 3;
constexpr int kStageBytes = kStageCols * 256 * 8;        // 16 KB
constexpr int kEpiWarps = 8;                 // two warps per TMEM lane quarter
constexpr int kSlicers = 8;                  // one column of a stage each
constexpr int kThreads = 32 * (kEpiWarps + kSlicers + 4);   // + producer warp + MMA warp + two idle warps (setmaxnreg works on groups of four)
constexpr int kSmemBytes = 1024 + kPlanesBytes + kStages * kStageBytes + 2 * 2048 /* g */;
constexpr int kTmemCols = 512;

__device__ __forceinline__ void mb_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAITC:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONEC;\n"
        "bra LAB_WAITC;\n"
        "LAB_DONEC:\n"
        "}\n" ::"r"(i8::s_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void mb_arrive_cluster_relaxed(uint32_t bar_cluster_addr) {   // no data is published with this arrival: no fence
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {   // 8 consecutive TMEM columns of this thread's lane
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void slicer_bar() { asm volatile("bar.sync 2, %0;" ::"n"(32 * kSlicers) : "memory"); }
__device__ __forceinline__ double2 lds_v2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ double dmax(double a, double b) { return b > a ? b : a; }
}  // namespace fz

struct GramI8FusedArgs {
    const double* X; int64_t ldx, x_bstride;
    int64_t k, n, batch;
    const double* gvec; int64_t g_bstride;
    double* qout; int64_t q_bstride;
    double* c0out;
    AlphaDev qscale, cscale;
    double* out; int64_t ldo, o_bstride;
    const int32_t* map;
    AlphaDev scale;
    int32_t accumulate;
    unsigned long long* stats;     // [batch][kGramI8StatsStride] (+ [batch * stride]: number of uncertified instances of this launch)
    double tol;
    int32_t force;                 // experiments (pmb_set_gram_mode 5): no certificate
    int32_t npairs;
    unsigned long long* tl;        // timeline of CTA 0 (PMB_FUSED_DBG & 32): [role][256] clock64 stamps, role 0 worker 0, 1 worker 8, 2 MMA issuer, 3 producer
    int32_t dbg;                   // timing diagnostics only (PMB_FUSED_DBG; results wrong by construction): 1 = no MMAs, 2 = no digit / plane work,
                                   // 4 = no epilogue arithmetic or stores, 8 = no global stores, 16 = only one of the 15 products per half
};

// the certificate of gram_i8_compute_level for 5 planes, on values (the merged statistics of the two CTAs)
__device__ __forceinline__ int fused_level(unsigned long long s0, unsigned long long s1, unsigned long long s2, const unsigned long long* dsq, double tol) {
    if (s0 == 0ull) return 5;
    const double sig = __longlong_as_double((long long)s0);
    if (!(sig < 1e200)) return kGramI8Fp64;
    const double ss = __longlong_as_double((long long)s1), l1 = __longlong_as_double((long long)s2);
    double R[fz::kNS];
#pragma unroll
    for (int s = 0; s < fz::kNS; ++s) R[s] = sqrt((double)dsq[s]);
    double by = 2.0 * 1.1 * exp2(-8.0 * fz::kNS) * l1;
#pragma unroll
    for (int s = 1; s < fz::kNS; ++s)
#pragma unroll
        for (int t = 1; t < fz::kNS; ++t)
            if (s + t >= fz::kNS) by += exp2(-14.0 - 8.0 * (s + t)) * R[s] * R[t];
    return (1.01 * sig * sig * by <= tol * ss) ? 5 : kGramI8Fp64;
}

// generic store of four consecutive entries of one output row (ragged edge, index map, accumulation): kept OUT of line -- the unrolled
// epilogue must stay small enough for the instruction cache (five warps per scheduler run three different code regions)
__device__ __noinline__ void fused_store_generic(double* outp, int64_t ldo, const int32_t* map, int accumulate, int64_t om, int gc, int nn, double v0,
                                                 double v1, double v2, double v3) {
    const double v[4] = {v0, v1, v2, v3};
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        const int gj = gc + j;
        if (gj < nn) {
            const int64_t on = map ? (int64_t)map[gj] : (int64_t)gj;
            double* p1 = outp + on + om * ldo;   // the MIRRORED entry, like the vector path: every (row, column) is written exactly once
            if (accumulate) *p1 += v[j]; else *p1 = v[j];
        }
    }
}

// exact int64 -> fp64 with ONE rounding (to nearest even), without the slow I2F.F64 of the XU pipe: both halves become doubles
// exactly through the 2^52 mantissa trick, one fma combines them.  |t| < 2^62.
__device__ __forceinline__ double i64_to_f64(long long t) {
    const int hi = (int)(t >> 32);
    const unsigned lo = (unsigned)t;
    const double dh = __hiloint2double(0x43300000, hi ^ (int)0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
    const double dl = __hiloint2double(0x43300000, (int)lo) - 4503599627370496.0;                 // 2^52
    return fma(dh, 4294967296.0, dl);
}

#ifndef FZ_TIMELINE
#define TL_STAMP(role, tag) do { } while (0)
#else
#define TL_STAMP(role, tag)                                                                                         \
    do {                                                                                                            \
        if (tl_on && tl_n < 127) { a.tl[(role) * 256 + 2 * tl_n] = (unsigned long long)(tag); a.tl[(role) * 256 + 2 * tl_n + 1] = (unsigned long long)clock64(); ++tl_n; } \
    } while (0)
#endif

// Warp roles (20 warps = 5 per scheduler: 2 epilogue + 2 slicers + 1 of the last group; launched with 96 registers each, setmaxnreg moves
// registers to the epilogue warps -- per scheduler 2 x 144 + 2 x 72 + 48 = 480 = 5 x 96):
//   warps 0-7   epilogue: TMEM -> exact integer combination of the five levels -> one rounding -> scale -> store (64 columns per thread)   144 registers
//   warps 8-15  slicers: one column per warp and stage -> statistics, digits, operand panels, q and c0                                 72 registers
//   warp 16     lane 0 = producer: bulk copies (UBLKCP) of the CTA's columns, 8 per stage, 3 stages ahead                                48 registers
//   warp 17     MMA issuer of the leader CTA (whole warp, warp-uniform descriptors, one elected lane issues); owns the TMEM allocation; warps 18, 19 idle
// The slicers start instance i+1 as soon as the products of instance i have read the panels, while the epilogue warps are still converting
// and storing instance i: slicing and epilogue overlap, the tensor pipe idles only while the panels are rewritten.
// Level order per output half: 4, then {2, 3}, then {0, 1}.  The epilogue carries w = ((a2 256 + a3) 256 + a4) between the units as a 42-bit
// integer -- 32 low bits per column plus three 10-bit high parts per register (86 registers for 64 columns) -- and finishes each entry as
// t = (a0 256 + a1) 2^24 + w the moment levels 0 and 1 arrive: exact (|t| < 2^55), one rounding to fp64.
__global__ void __launch_bounds__(fz::kThreads, 1) gram_i8_fused_kernel(const GramI8FusedArgs a) {
    using namespace fz;
    using fz::mb_arrive_cluster_relaxed;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t stage_full[kStages], stage_empty[kStages];
    __shared__ __align__(8) uint64_t conv_done, planes_free;
    __shared__ __align__(8) uint64_t g_full[2], g_empty[2];       // the offset vector of instance parity p has landed in / been read from gbuf[p]
    __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long cstat[8];          // this CTA's running maxima: sigma, ss, l1 / sigma, dsq[0..4]
    __shared__ __align__(8) unsigned long long xstat[2][2][8];    // [instance parity][CTA rank][..]: both CTAs' statistics, in both CTAs
    __shared__ __align__(8) double sig[2][2 * kCols];             // [instance parity][global column]: sigma_j 2^-23 of all 256 columns
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = i8::cluster_ctarank();
    const uint32_t base = (i8::s_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t planes = base, staging = base + kPlanesBytes, gbuf = staging + kStages * kStageBytes;
    const int64_t n = a.n, k = a.k;
    const int nloc = (int)max((int64_t)0, min((int64_t)kCols, n - (int64_t)kCols * rank));   // live columns of this CTA
    const int nst = (nloc + kStageCols - 1) / kStageCols;
    const int nks = (int)((k + 31) / 32);            // K = 32 steps
    const int nkb = (int)((k + kKBlock - 1) / kKBlock);   // k-blocks in use (1 or 2)
    const int nh = n > 64 ? 2 : 1;                   // output halves with live columns
    const uint32_t colbytes = (uint32_t)(k * 8);
    const int64_t pair0 = blockIdx.x >> 1;
    [[maybe_unused]] const bool tl_on = (a.dbg & 32) && a.tl && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8 || warp == 16 || warp == 17);
    [[maybe_unused]] int tl_n = 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { i8::mb_init(&stage_full[s], 1); i8::mb_init(&stage_empty[s], kSlicers); }
        i8::mb_init(&conv_done, 2);
        i8::mb_init(&planes_free, 1);
        i8::mb_init(&g_full[0], 1); i8::mb_init(&g_full[1], 1); i8::mb_init(&g_empty[0], 1); i8::mb_init(&g_empty[1], 1);
#pragma unroll
        for (int b = 0; b < 2; ++b) { i8::mb_init(&tmem_full[b], 1); i8::mb_init(&tmem_empty[b], 2 * kEpiWarps); }
#pragma unroll
        for (int s = 0; s < 8; ++s) cstat[s] = 0ull;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // staging columns are 256 rows apart whatever k is: rows k .. 255 (and the tail of the offset vector) are zeroed ONCE, the bulk copies
    // never touch them, so the slicers read whole columns without bounds tests (zero rows add nothing to any statistic or digit)
    if (k < 256) {
        for (uint32_t o = (uint32_t)tid * 8u; o < (uint32_t)(kStages * kStageBytes + 2 * 2048); o += (uint32_t)kThreads * 8u)
            asm volatile("st.shared.u64 [%0], %1;" ::"r"(staging + o), "l"(0ull) : "memory");
        fence_proxy_async_smem();
    }
    if (warp == kEpiWarps + kSlicers + 1) i8::tmem_alloc_pair(&tmem_slot, kTmemCols);
    i8::tc_fence_before();
    i8::cluster_sync_all();
    i8::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp >= kEpiWarps + kSlicers) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        if (warp == kEpiWarps + kSlicers) {
            if (lane == 0) {
                // ===== producer: bulk copies of this CTA's columns, 8 per stage =====
                uint32_t slot = 0, ph = 0, ii = 0;
                const bool contiguous = k == 256 && a.ldx == 256;
                for (int64_t inst = pair0; inst < a.batch; inst += a.npairs, ++ii) {
                    const double* xi = a.X + inst * a.x_bstride + (int64_t)kCols * rank * a.ldx;
                    if (a.gvec && nloc > 0) {   // gbuf[ii & 1] was last read while instance ii - 2 was sliced
                        i8::mb_wait(&g_empty[ii & 1u], ((ii >> 1) & 1u) ^ 1u);
                        i8::mb_expect_tx(&g_full[ii & 1u], colbytes);
                        i8::bulk_load_1d(gbuf + (ii & 1u) * 2048u, a.gvec + inst * a.g_bstride, colbytes, &g_full[ii & 1u]);
                    }
                    for (int st = 0; st < nst; ++st) {
                        i8::mb_wait(&stage_empty[slot], ph ^ 1u);
                        if (st == 0 || st == nst - 1) TL_STAMP(3, 1000 * ii + st);
                        const int c0 = st * kStageCols, nc = min(kStageCols, nloc - c0);
                        i8::mb_expect_tx(&stage_full[slot], (uint32_t)nc * colbytes);
                        const uint32_t dst = staging + slot * kStageBytes;
                        if (contiguous) {
                            i8::bulk_load_1d(dst, xi + (int64_t)c0 * 256, (uint32_t)nc * 2048u, &stage_full[slot]);
                        } else {
                            for (int c = 0; c < nc; ++c) i8::bulk_load_1d(dst + (uint32_t)c * 2048u, xi + (int64_t)(c0 + c) * a.ldx, colbytes, &stage_full[slot]);
                        }
                        if (++slot == kStages) { slot = 0; ph ^= 1u; }
                    }
                }
            }
        } else if (warp == kEpiWarps + kSlicers + 1 && rank == 0) {
            // ===== MMA issuer (leader CTA, whole warp): per instance 2 halves x 3 units of digit-plane products into alternating 256-column
            // TMEM buffers.  Descriptors: the high word is constant (128-byte rows, 1024 bytes between 8-row groups, SWIZZLE_128B), the low
            // word is 1 << 16 | address >> 4 -- one add per instruction =====
            // TMEM regions: A = columns [0, 256) takes unit 0 (level 4, 128 columns) and then unit 1 (levels 2, 3) -- both drained by a few
            // tcgen05.ld; B = columns [256, 512) takes unit 2 (levels 0, 1), whose drain is interleaved with the conversions and stores: while it
            // lasts the products of the NEXT half's units 0 and 1 already run in region A
            uint32_t ii = 0, ucA = 0, ucB = 0;
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t free_remote = i8::mapa_u32(i8::s_u32(&planes_free), 1);
            const uint64_t dproto = i8::smem_desc_sw64<128>(0);
            const uint64_t dhi = dproto & 0xFFFFFFFF00000000ull;
            const uint32_t dlo0 = (uint32_t)dproto | ((planes & 0x3FFFFu) >> 4);
            for (int64_t inst = pair0; inst < a.batch; inst += a.npairs, ++ii) {
                const uint32_t par = ii & 1u;
                i8::mb_wait(&conv_done, par);
                TL_STAMP(2, 1000 * ii + 1);
                int lv = 5;
                if (!a.force) {
                    unsigned long long m[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s) m[s] = xstat[par][0][s] > xstat[par][1][s] ? xstat[par][0][s] : xstat[par][1][s];
                    lv = fused_level(m[0], m[1], m[2], m + 3, a.tol);
                }
                if (lv != 5) {   // not certified: the fp64 kernel owns the instance; the operand panels are free at once
                    if (lane == 0) { i8::mb_arrive(&planes_free); i8::mb_arrive_cluster(free_remote); }
                    __syncwarp();
                    continue;
                }
                i8::tc_fence_after();
                for (int h = 0; h < nh; ++h) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const uint32_t buf = u == 2 ? 1u : 0u;
                        const uint32_t use = u == 2 ? ucB++ : ucA++;
                        i8::mb_wait(&tmem_empty[buf], (use & 1u) ^ 1u);
                        i8::tc_fence_after();
                        TL_STAMP(2, 1000 * ii + 10 + 3 * h + u);
                        uint32_t dlo = dlo0;
                        asm volatile("" : "+r"(dlo));   // opaque: descriptors are formed next to their instruction, not hoisted (register budget: 48)
                        // unit 0: level 4;  unit 1: levels 2, 3;  unit 2: levels 0, 1.  Level l = all (s, t) with s + t = l.
#pragma unroll
                        for (int lsub = 0; lsub < 2; ++lsub) {
                            const int l = u == 0 ? 4 : (u == 1 ? 2 + lsub : lsub);
                            if (u == 0 && lsub == 1) break;
                            const uint32_t dcol = tmem_base + buf * 256u + (uint32_t)lsub * 128u;
#pragma unroll
                            for (int s = 0; s <= l; ++s) {
                                const int t = l - s;
#ifdef FZ_DEBUG
                                if ((a.dbg & 1) || ((a.dbg & 16) && l > 0)) continue;
#endif
                                const uint32_t alo = dlo + (uint32_t)(s * kPlaneBytes >> 4), blo = dlo + (uint32_t)(t * kPlaneBytes >> 4) + (uint32_t)h * (64u * kKBlock >> 4);
#pragma unroll
                                for (int ks = 0; ks < 8; ++ks) {
                                    if (ks < nks) {
                                        const uint32_t off = (uint32_t)((ks >> 2) * kPanel + (ks & 3) * 32) >> 4;
                                        i8::mma_i8_pair_elect(dcol, dhi | (uint64_t)(alo + off), dhi | (uint64_t)(blo + off), idesc, (s > 0 || ks > 0) ? 1u : 0u);
                                    }
                                }
                            }
                        }
                        i8::mma_commit_pair_elect(&tmem_full[buf]);
                    }
                }
                i8::mma_commit_pair_elect(&planes_free);   // every product of this instance has read its operands
                TL_STAMP(2, 1000 * ii + 99);
            }
        }
    } else if (warp >= kEpiWarps) {
        // ===== slicers =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        const int wig = warp - kEpiWarps;         // this warp's column inside a stage
        const uint32_t peer = rank ^ 1u;
        const uint32_t conv_remote = i8::mapa_u32(i8::s_u32(&conv_done), peer);
        // staging reads: rows 4 lane .. 4 lane + 3 of each k-block as two 16-byte loads with the halves SWAPPED for lanes 4-7 of every 8, so a
        // quarter warp touches all 32 banks; the swap is undone for free by the selectors of the byte permutes that assemble the plane words
        const bool swp = (lane & 4) != 0;
        const uint32_t lane_off = (uint32_t)lane * 32u + (swp ? 16u : 0u);
        const uint32_t selA = swp ? 0x1054u : 0x5410u, selB = swp ? 0x3276u : 0x7632u;
        const uint32_t word_off = (((uint32_t)(lane >> 2)) << 4) | ((uint32_t)(lane & 3) * 4u);   // this lane's 4 bytes inside a 128-byte row, before the swizzle
        const bool restore = k < 96;              // the reduction scratch (first 768 bytes of the warp's own column slot) reaches into the zeroed tail rows
        uint32_t slot = 0, ph = 0, ii = 0;
        for (int64_t inst = pair0; inst < a.batch; inst += a.npairs, ++ii) {
            const uint32_t par = ii & 1u;
            TL_STAMP(1, 1000 * ii + 0);
            if (ii > 0) i8::mb_wait(&planes_free, (ii - 1u) & 1u);   // the previous instance's products have read the panels
            TL_STAMP(1, 1000 * ii + 1);
            double w_sig = 0.0, w_ss = 0.0, w_l1 = 0.0;   // running maxima: sigma (uniform), ||x||^2 (lane 0), ||x / sigma||_1 (lane 8)
            uint32_t w_dsq[kNS];
            bool w_bad = false;
#pragma unroll
            for (int s = 0; s < kNS; ++s) w_dsq[s] = 0u;
            const double qs = a.gvec ? (a.qscale.ptr ? a.qscale.c * a.qscale.ptr[inst * a.qscale.bstride] : a.qscale.c) : 0.0;
            if (a.gvec && nloc > 0) i8::mb_wait(&g_full[par], (ii >> 1) & 1u);   // the offset vector of this instance (same, swapped, element order as the columns)
            const uint32_t gb = gbuf + par * 2048u + lane_off;
            for (int st = 0; st < nst; ++st) {
                i8::mb_wait(&stage_full[slot], ph);
                const uint32_t myslot = slot;
                if (++slot == kStages) { slot = 0; ph ^= 1u; }
                const uint32_t colbase = staging + myslot * kStageBytes + (uint32_t)wig * 2048u;
                const uint32_t cb = colbase + lane_off;
                const int lc = st * kStageCols + wig;   // local column of this warp
                const bool live = lc < nloc;            // warp-uniform
                double v[8];
                {
                    const double2 p0 = lds_v2(cb), p1 = lds_v2(cb ^ 16u), p2 = lds_v2(cb + 1024u), p3 = lds_v2((cb + 1024u) ^ 16u);
                    v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
                }
                __syncwarp();
#ifdef FZ_DEBUG
                if (a.dbg & 2) { if (lane == 0) i8::mb_arrive(&stage_empty[myslot]); continue; }
#endif
                if (!live) { if (lane == 0) i8::mb_arrive(&stage_empty[myslot]); continue; }
                double am = 0.0, ss = 0.0, l1 = 0.0, dg = 0.0;
                {
                    double gv[8];
                    if (a.gvec) {
                        const double2 g0 = lds_v2(gb), g1 = lds_v2(gb ^ 16u), g2 = lds_v2(gb + 1024u), g3 = lds_v2((gb + 1024u) ^ 16u);
                        gv[0] = g0.x; gv[1] = g0.y; gv[2] = g1.x; gv[3] = g1.y; gv[4] = g2.x; gv[5] = g2.y; gv[6] = g3.x; gv[7] = g3.y;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) gv[e] = 0.0;
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const double av = fabs(v[e]);
                        am = dmax(am, av);
                        l1 += av;
                        ss = fma(v[e], v[e], ss);
                        dg = fma(v[e], gv[e], dg);
                    }
                    if (a.gvec && rank == 0 && lc == 0) {   // c0 += cscale g.g, once per instance
                        double gg = 0.0;
#pragma unroll
                        for (int e = 0; e < 8; ++e) gg = fma(gv[e], gv[e], gg);
#pragma unroll
                        for (int o = 16; o; o >>= 1) gg += __shfl_xor_sync(0xffffffffu, gg, o);
                        if (lane == 0) {
                            const double cs = a.cscale.ptr ? a.cscale.c * a.cscale.ptr[inst * a.cscale.bstride] : a.cscale.c;
                            atomicAdd(a.c0out + inst, cs * gg);
                        }
                    }
                }
                // per-lane partial sums -> scratch (the head of this warp's own column slot: the column sits in registers); lanes 0-23 add four
                // partials each, three butterfly steps finish the three sums: ||x||^2 in lane 0, ||x||_1 in lane 8, x.g in lane 16
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(colbase + (uint32_t)lane * 8u), "d"(ss) : "memory");
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(colbase + 256u + (uint32_t)lane * 8u), "d"(l1) : "memory");
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(colbase + 512u + (uint32_t)lane * 8u), "d"(dg) : "memory");
                // largest magnitude of the column in every lane: two integer warp reductions on the bit pattern (monotonic for values >= 0)
                const uint32_t hmax = __reduce_max_sync(0xffffffffu, (uint32_t)__double2hiint(am));
                const uint32_t lmax = __reduce_max_sync(0xffffffffu, (uint32_t)__double2hiint(am) == hmax ? (uint32_t)__double2loint(am) : 0u);
                const double amax = __hiloint2double((int)hmax, (int)lmax);
                bool bad = __any_sync(0xffffffffu, !(l1 <= DBL_MAX));   // an Inf / NaN entry poisons this lane's partial 1-norm
                if (amax > 0.0 && !(amax >= 1e-150 && amax <= 1e150)) bad = true;
                const bool zero = !(amax > 0.0) || bad;
                const double inv = zero ? 0.0 : 0.99 / amax;
                const double sigma = zero ? 1.0 : amax * (1.0 / 0.99);
                const double mult = inv * 140737488355328.0;   // 2^47
                __syncwarp();
                double tot = 0.0;
                if (lane < 24) {
                    const double2 r0 = lds_v2(colbase + (uint32_t)lane * 32u), r1 = lds_v2(colbase + (uint32_t)lane * 32u + 16u);
                    tot = (r0.x + r0.y) + (r1.x + r1.y);
                }
                __syncwarp();
                if (restore) {
                    asm volatile("st.shared.u64 [%0], %1;" ::"r"(colbase + (uint32_t)lane * 8u), "l"(0ull) : "memory");
                    asm volatile("st.shared.u64 [%0], %1;" ::"r"(colbase + 256u + (uint32_t)lane * 8u), "l"(0ull) : "memory");
                    asm volatile("st.shared.u64 [%0], %1;" ::"r"(colbase + 512u + (uint32_t)lane * 8u), "l"(0ull) : "memory");
                    fence_proxy_async_smem();
                    __syncwarp();
                }
                if (lane == 0) i8::mb_arrive(&stage_empty[myslot]);   // column and scratch are done with the slot: it can take the stage 3 ahead
                tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                // digits: byte b of Y = (X + 0x8080808080) ^ 0x8080808080 is plane 5 - b; the XOR is applied to the assembled plane words
                uint32_t dsq[kNS];
#pragma unroll
                for (int s2 = 0; s2 < kNS; ++s2) dsq[s2] = 0u;
                const uint32_t rowaddr = planes + (uint32_t)lc * kKBlock + ((word_off & ~0x70u) | ((word_off ^ ((uint32_t)(lc & 7) << 4)) & 0x70u));
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    if (hb >= nkb) break;
                    uint32_t lo[4], hi[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const unsigned long long Y = (unsigned long long)__double_as_longlong(fma(v[4 * hb + t], mult, 6755399441055744.0)) + 0x0000008080808080ull;
                        lo[t] = (uint32_t)Y; hi[t] = (uint32_t)(Y >> 32);
                    }
                    const uint32_t u0 = __byte_perm(lo[0], lo[1], 0x7362), u1 = __byte_perm(lo[2], lo[3], 0x7362);     // bytes 2, 3 of two rows
                    const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x0051), t1 = __byte_perm(lo[2], lo[3], 0x0051);     // byte 1 of two rows
                    const uint32_t h0 = __byte_perm(hi[0], hi[1], 0x5140), h1 = __byte_perm(hi[2], hi[3], 0x5140);     // bytes 4, 5 of two rows
                    uint32_t w[kNS];
                    w[0] = __byte_perm(h0, h1, selB);                      // byte 5: the signed leading digit, no XOR
                    w[1] = __byte_perm(h0, h1, selA) ^ 0x80808080u;        // byte 4
                    w[2] = __byte_perm(u0, u1, selB) ^ 0x80808080u;        // byte 3
                    w[3] = __byte_perm(u0, u1, selA) ^ 0x80808080u;        // byte 2
                    w[4] = __byte_perm(t0, t1, selA) ^ 0x80808080u;        // byte 1
#pragma unroll
                    for (int p = 0; p < kNS; ++p) {
                        dsq[p] = (uint32_t)__dp4a((int)w[p], (int)w[p], (int)dsq[p]);
                        sts_u32(rowaddr + (uint32_t)hb * kPanel + (uint32_t)p * kPlaneBytes, w[p]);
                    }
                }
#pragma unroll
                for (int s2 = 0; s2 < kNS; ++s2) dsq[s2] = __reduce_add_sync(0xffffffffu, dsq[s2]);
                const int64_t gj = (int64_t)kCols * rank + lc;
                if (lane == 0) {
                    const double s23 = sigma * 0x1p-23;
                    const uint32_t so = i8::s_u32(&sig[par][gj]);
                    sig[par][gj] = s23;
                    st_cluster_f64(i8::mapa_u32(so, peer), s23);
                    if (!zero) w_ss = dmax(w_ss, tot);
                } else if (lane == 8) {
                    if (!zero) w_l1 = dmax(w_l1, tot * inv);
                } else if (lane == 16 && a.gvec) {
                    atomicAdd(a.qout + inst * a.q_bstride + (a.map ? a.map[gj] : gj), qs * tot);
                }
                if (bad) w_bad = true;
                else if (!zero) {
                    w_sig = dmax(w_sig, sigma);
#pragma unroll
                    for (int s2 = 0; s2 < kNS; ++s2) w_dsq[s2] = w_dsq[s2] > dsq[s2] ? w_dsq[s2] : dsq[s2];
                }
            }
            if (lane == 0) {
                if (w_bad) atomicMax(&cstat[0], (unsigned long long)__double_as_longlong(CUDART_INF));
                else if (w_sig > 0.0) {
                    atomicMax(&cstat[0], (unsigned long long)__double_as_longlong(w_sig));
                    atomicMax(&cstat[1], (unsigned long long)__double_as_longlong(w_ss));
#pragma unroll
                    for (int s2 = 0; s2 < kNS; ++s2) atomicMax(&cstat[3 + s2], (unsigned long long)w_dsq[s2]);
                }
            } else if (lane == 8 && !w_bad && w_l1 > 0.0) {
                atomicMax(&cstat[2], (unsigned long long)__double_as_longlong(w_l1));
            }
            TL_STAMP(1, 1000 * ii + 2);
            fence_proxy_async_smem();   // the panel writes (generic proxy) become visible to the tensor core (async proxy)
            slicer_bar();
            if (wig == 0 && lane == 0) {   // publish this CTA's statistics in both CTAs, then count the CTA in on both conv_done barriers
                if (a.gvec && nloc > 0) i8::mb_arrive(&g_empty[par]);   // every slicer of this CTA is done with the offset vector
#pragma unroll
                for (int s2 = 0; s2 < 8; ++s2) {
                    const unsigned long long vv = cstat[s2];
                    xstat[par][rank][s2] = vv;
                    st_cluster_u64(i8::mapa_u32(i8::s_u32(&xstat[par][rank][s2]), peer), vv);
                    cstat[s2] = 0ull;
                }
                i8::mb_arrive_cluster(i8::mapa_u32(i8::s_u32(&conv_done), rank));
                i8::mb_arrive_cluster(conv_remote);
                TL_STAMP(1, 1000 * ii + 3);
                if (rank == 0) {   // the record the fp64 companion launch and pmb_gram_level_counts read
                    i8::mb_wait(&conv_done, par);
                    unsigned long long m[8];
#pragma unroll
                    for (int s2 = 0; s2 < 8; ++s2) m[s2] = xstat[par][0][s2] > xstat[par][1][s2] ? xstat[par][0][s2] : xstat[par][1][s2];
                    const int lv = a.force ? 5 : fused_level(m[0], m[1], m[2], m + 3, a.tol);
                    unsigned long long* sp = a.stats + inst * kGramI8StatsStride;
#pragma unroll
                    for (int s2 = 0; s2 < 8; ++s2) sp[s2] = m[s2];
                    sp[8] = 0ull; sp[14] = (unsigned long long)n; sp[15] = (unsigned long long)lv;
                    if (lv != 5) atomicAdd(a.stats + a.batch * kGramI8StatsStride, 1ull);
                }
            }
        }
    } else {
        // ===== epilogue =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        const int q = warp & 3;                   // TMEM lane quarter this warp may read
        const int chalf = warp >> 2;              // which 64 of a unit's 128 output columns (= which CTA supplied them)
        const uint32_t empty_remote0 = i8::mapa_u32(i8::s_u32(&tmem_empty[0]), 0), empty_remote1 = i8::mapa_u32(i8::s_u32(&tmem_empty[1]), 0);
        const bool vec_ok = !a.map && !a.accumulate && (a.ldo & 3) == 0 && (a.o_bstride & 3) == 0 && (((uintptr_t)a.out) & 31) == 0;
        const int nn = (int)n;
        const int row0 = kCols * (int)rank + q * 32;   // first row of this warp
        const int gi = row0 + lane;
        const bool row_ok = gi < nn;
        uint32_t ii = 0, ucA = 0, ucB = 0;
        for (int64_t inst = pair0; inst < a.batch; inst += a.npairs, ++ii) {
            const uint32_t par = ii & 1u;
            TL_STAMP(0, 1000 * ii + 0);
            i8::mb_wait(&conv_done, par);
            TL_STAMP(0, 1000 * ii + 4);
            int lv = 5;
            if (!a.force) {
                unsigned long long m[8];
#pragma unroll
                for (int s2 = 0; s2 < 8; ++s2) m[s2] = xstat[par][0][s2] > xstat[par][1][s2] ? xstat[par][0][s2] : xstat[par][1][s2];
                lv = fused_level(m[0], m[1], m[2], m + 3, a.tol);
            }
            if (lv != 5) continue;
            const double scl = a.scale.ptr ? a.scale.c * a.scale.ptr[inst * a.scale.bstride] : a.scale.c;
            const double sgi = row_ok ? sig[par][gi] : 0.0;   // sigma_i 2^-23
            const double fgi = scl * sgi;
            double* outp = a.out + inst * a.o_bstride;
            for (int h = 0; h < nh; ++h) {
                const int gj0 = kCols * chalf + 64 * h;   // global column of this warp's first output column
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * chalf);
                uint32_t wlo[64];     // low 32 bits of w = (a2 256 + a3) 256 + a4   (first: a4 itself)
                uint32_t whi[22];     // bits 32-41 of w, three 10-bit fields per register
#pragma unroll
                for (int r = 0; r < 22; ++r) whi[r] = 0u;
                // unit 0: level 4 (region A)
                {
                    const uint32_t buf = 0u;
                    i8::mb_wait(&tmem_full[buf], ucA & 1u);
                    TL_STAMP(0, 1000 * ii + 10 * (h + 1) + 1);
                    i8::tc_fence_after();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        int32_t a4[8];
                        tmem_ld8(trow + buf * 256u + 8u * c, a4);
                        i8::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) wlo[8 * c + j] = (uint32_t)a4[j];
                    }
                    i8::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mb_arrive_cluster_relaxed(empty_remote0);
                    ++ucA;
                }
                // unit 1: levels 2, 3 -> w = (a2 256 + a3) 256 + a4, |w| < 2^40 (region A)
                {
                    const uint32_t buf = 0u;
                    i8::mb_wait(&tmem_full[buf], ucA & 1u);
                    TL_STAMP(0, 1000 * ii + 10 * (h + 1) + 2);
                    i8::tc_fence_after();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        int32_t a2[8], a3[8];
                        tmem_ld8(trow + buf * 256u + 8u * c, a2);
                        tmem_ld8(trow + buf * 256u + 128u + 8u * c, a3);
                        i8::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int e = 8 * c + j;
                            const long long w = (((long long)a2[j] << 8) + (long long)a3[j]) * 256 + (long long)(int32_t)wlo[e];
                            wlo[e] = (uint32_t)w;
                            whi[e / 3] |= ((uint32_t)(w >> 32) & 0x3FFu) << (10 * (e % 3));
                        }
                    }
                    i8::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mb_arrive_cluster_relaxed(empty_remote0);
                    ++ucA;
                }
                // unit 2 (region B): levels 0, 1 -> t = (a0 256 + a1) 2^24 + w, the one rounding to fp64, scale, store.  Every unordered pair of columns is
                // scaled as (t * (scale sigma_lo 2^-23)) * (sigma_hi 2^-23), lo <= hi the column indices: what the two-kernel route computes once
                // and mirrors -- both triangles are bit-identical to it.  A chunk of 8 columns lies right of, left of, or across the diagonal for
                // the whole warp (rows row0 .. row0 + 31)
                {
                    const uint32_t buf = 1u;
                    i8::mb_wait(&tmem_full[buf], ucB & 1u);
                    TL_STAMP(0, 1000 * ii + 10 * (h + 1) + 3);
                    i8::tc_fence_after();
                    const bool fast = vec_ok && row0 + 32 <= nn && gj0 + 64 <= nn;
                    const int64_t om = (!fast && row_ok) ? (a.map ? (int64_t)a.map[gi] : (int64_t)gi) : 0;
                    double* pcol = outp + gj0 + (int64_t)gi * a.ldo;
                    const double* sg = &sig[par][gj0];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        int32_t a0[4], a1[4];
                        tmem_ld4(trow + buf * 256u + 4u * c, a0);
                        tmem_ld4(trow + buf * 256u + 128u + 4u * c, a1);
                        i8::tmem_ld_wait();
                        if (c == 15) {   // the last read of this unit: hand the buffer back before the last conversions and stores
                            i8::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mb_arrive_cluster_relaxed(empty_remote1);
                        }
#ifdef FZ_DEBUG
                        if (a.dbg & 4) continue;
#endif
                        const int gc = gj0 + 4 * c;
                        double val[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int e = 4 * c + j;
                            // t = (a0 256 + a1) 2^24 + w as a 64-bit integer in two words: low = wlo + (p << 24), high = hi10 + (p >> 8) + carry
                            const int32_t p01 = a0[j] * 256 + a1[j];
                            const int32_t hi10 = ((int32_t)(whi[e / 3] << (22 - 10 * (e % 3)))) >> 22;   // sign-extended 10-bit field
                            const uint32_t plo = (uint32_t)p01 << 24;
                            const uint32_t tlo = wlo[e] + plo;
                            const int32_t thi = hi10 + (p01 >> 8) + (int32_t)(tlo < plo);
                            const double dh = __hiloint2double(0x43300000, thi ^ (int)0x80000000) - 4503601774854144.0;   // exact: 2^52 + 2^31 offset
                            const double dl = __hiloint2double(0x43300000, (int)tlo) - 4503599627370496.0;                 // exact: 2^52 offset
                            const double t = fma(dh, 4294967296.0, dl);                                                     // the one rounding
                            const double sgj = sg[e];
                            const bool up = gi <= gc + j;
                            const double fa = up ? fgi : scl * sgj, fb = up ? sgj : sgi;
                            val[j] = (t * fa) * fb;
                        }
#ifdef FZ_DEBUG
                        if (a.dbg & 8) { if (val[0] == 1.2345e-300 && val[3] == 5.0) outp[0] = val[1]; continue; }
#endif
                        if (fast) {
                            // H is symmetric and the values are bit-symmetric: this thread's entries of ROW gi go out as consecutive entries of
                            // COLUMN gi (256-bit streaming stores)
                            asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(pcol + 4 * c), "d"(val[0]), "d"(val[1]), "d"(val[2]), "d"(val[3]) : "memory");
                        } else if (row_ok) {
                            fused_store_generic(outp, a.ldo, a.map, a.accumulate, om, gc, nn, val[0], val[1], val[2], val[3]);
                        }
                    }
                    TL_STAMP(0, 1000 * ii + 10 * (h + 1) + 4);
                    ++ucB;
                }
            }
        }
    }
    i8::tc_fence_before();
    i8::cluster_sync_all();
    if (warp == fz::kEpiWarps + fz::kSlicers + 1) {
        i8::tc_fence_after();
        i8::tmem_dealloc_pair(tmem_base, fz::kTmemCols);
    }
}

// ---- host ----------------------------------------------------------------------------------------------------------------------
static int g_fused_max_pairs = -1;

void configure_gram_i8_fused_kernel() {
    cudaFuncSetAttribute(gram_i8_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fz::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
}

static int fused_max_pairs(int sm_count) {
    if (g_fused_max_pairs >= 0) return g_fused_max_pairs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(sm_count / 2 * 2));
    cfg.blockDim = dim3(fz::kThreads);
    cfg.dynamicSmemBytes = fz::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, gram_i8_fused_kernel, &cfg) != cudaSuccess || nclusters <= 0) {
        (void)cudaGetLastError();
        nclusters = sm_count / 2;
    }
    g_fused_max_pairs = std::min(nclusters, sm_count / 2);
    return g_fused_max_pairs;
}

bool gram_i8_fused_applicable(const GramArgs& g, int64_t batch) {
    if (g.n < 1 || g.n > 2 * fz::kCols || g.k < 2 || g.k > 256 || (g.k & 1)) return false;
    if (g.gate || g.packed_dev || g.packed_host) return false;
    if ((((uintptr_t)g.X) & 15) || (g.ldx & 1) || (g.x_bstride & 1) || g.ldx < g.k) return false;
    if (g.gvec && ((((uintptr_t)g.gvec) & 15) || ((g.g_bstride & 1) && batch > 1))) return false;
    if (batch >= ((int64_t)1 << 31)) return false;
    return true;
}

// predicted seconds, fitted to the measured timeline of a CTA pair (profiles/r02_e_fused_v6_timeline.txt, n = k = 256: 37 us per instance):
// ~6 us of hand-shakes and TMEM drains that nothing hides, the slicing of the CTA's 128 columns at ~14 GB/s per SM, and the span of the 15
// products x nh halves of (M 256, N 128, K 32) instructions, stretched 1.5x by the epilogue it waits for.  The fp64 kernel reaches ~19 TFLOP/s
// of SYRK-counted flops on small instances (profiles/r02_a_gram_tma_cfg3_ncu_full.txt)
bool gram_i8_fused_pays_off(int64_t k, int64_t n, int64_t batch, int sm_count) {
    if (const char* e = getenv("PMB_GRAM_I8_FUSED")) return atoi(e) != 0;
    const double per = 6e-6 + (double)std::min<int64_t>(n, 128) * (double)k * 8.0 / 14e9 + 1.5 * (double)((k + 31) / 32) * 15.0 * (n > 64 ? 2.0 : 1.0) * 64.0 / 1.9e9;
    const double t_fused = std::ceil((double)batch / (double)(sm_count / 2)) * per + 8e-6;
    const double t_f64 = (double)batch * k * n * (n + 1) / 19e12 + 6e-6;
    return t_fused < t_f64;
}

bool launch_gram_i8_fused(const GramArgs& g, const GramI8Scratch& sc, int64_t batch, int sm_count, cudaStream_t stream) {
    if (!sc.stats || !gram_i8_fused_applicable(g, batch)) return false;
    GramI8FusedArgs a{};
    a.X = g.X; a.ldx = g.ldx; a.x_bstride = g.x_bstride; a.k = g.k; a.n = g.n; a.batch = batch;
    a.gvec = g.gvec; a.g_bstride = g.g_bstride; a.qout = g.qout; a.q_bstride = g.q_bstride; a.c0out = g.c0out; a.qscale = g.qscale; a.cscale = g.cscale;
    a.out = g.H; a.ldo = g.ldh; a.o_bstride = g.h_bstride; a.map = g.map; a.scale = g.scale; a.accumulate = g.accumulate;
    a.stats = sc.stats; a.tol = sc.tol; a.force = sc.force_ns != 0;
    a.npairs = (int32_t)std::min<int64_t>(batch, fused_max_pairs(sm_count));
    if (const char* e = getenv("PMB_FUSED_DBG")) a.dbg = atoi(e);
    a.tl = sc.stats + batch * kGramI8StatsStride + 2;   // the planner / probe reserves 1024 more entries when the timeline is wanted
    if (a.npairs < 1) return false;
    cudaMemsetAsync(sc.stats + batch * kGramI8StatsStride, 0, sizeof(unsigned long long), stream);   // uncertified-instance counter
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * a.npairs));
    cfg.blockDim = dim3(fz::kThreads);
    cfg.dynamicSmemBytes = fz::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gram_i8_fused_kernel, a) == cudaSuccess;
}

}  // namespace pmb
