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
//   warps 0-7   epilogue   tcgen05.ld, exact integer combination of the five int32 levels (carried between the units in packed form), ONE
//                          rounding to fp64, sigma_i sigma_j scaling, 256-bit streaming stores                              (144 registers)
//   warps 8-15  slicers    one column per warp and stage: largest magnitude / 2-norm / 1-norm / X'g by warp reductions, fixed-point
//                          digits, five int8 planes written straight into the 128-byte-swizzled K-major operand layout tcgen05 reads
//                          (160 KB per CTA: 5 planes x 2 k-blocks x 128 columns x 128 B); the per-instance statistics of the error
//                          certificate are exchanged between the two CTAs through distributed shared memory              (72 registers)
//   warp 16     producer   X columns of the CTA's half (128 columns) travel global -> shared as 1-D bulk copies (UBLKCP), 8 columns per
//                          stage, 3 stages ahead; the offset vector g rides along
//   warp 17     MMA issuer (leader CTA only, whole warp, elect.sync) tcgen05.mma.cta_group::2.kind::i8, M = 256 (CTA r supplies rows
//                          128 r ..), N = 128 (each CTA supplies 64 of its columns), K = 32: the 15 digit-plane products of an output
//                          half go out as three units {level 4}, {2, 3}, {0, 1} into two 256-column TMEM regions
//   warps 18-19 idle (setmaxnreg works on groups of four); with -DFZ_WATCHDOG warp 18 watches the others' milestones
//
// Two kernels are compiled from one body (kernels_gram_i8_fused_body.inc):
//   gram_i8_fused_kernel      the operand X is read from memory (HBM traffic = the algorithmic minimum: X once, H once)
//   gram_i8_fused_gen_kernel  X is a UNIFORM Parameter nothing else reads: the slicers generate their columns with Philox (pmb_philox.h, the
//                             same bits as the prologue's fill), nothing is loaded, and -- since nothing has to arrive -- they slice the
//                             first four stages of the NEXT instance while the products of the current one still read the operand panels,
//                             parking digits and sigma in the 48 KB the loading variant uses for its ring; the epilogue warps, idle until the
//                             next instance is sliced, slice one column of its last six stages themselves
// The full square is computed (no mirrored writes).  The certificate is evaluated per instance inside the kernel (5 planes or "not
// certified"); uncertified instances are left to the fp64 kernel that follows in the update graph (GramArgs::cert_stats), exactly as on
// the two-kernel route.  Measured variants that did not make it are kept as text under tools/probe/attic/ (DESIGN.md section 4).
#include <cuda.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <math_constants.h>

#include "pmb_i8_ptx.h"
#include "pmb_kernels.h"
#include "pmb_philox.h"

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
// "TMEM region read" arrival: no DATA is published with it (the tcgen05.ld results are in registers, tcgen05.wait::ld has returned and
// tcgen05.fence::before_thread_sync orders them) -- relaxed keeps the MEMBAR of a release.cluster arrival out of the drain.  compute-sanitizer's
// racecheck does not model the relaxed form (it reports the barrier word itself as a hazard): build with -DFZ_STRICT_SYNC for that tool
__device__ __forceinline__ void mb_arrive_cluster_relaxed(uint32_t bar_cluster_addr) {
#ifdef FZ_STRICT_SYNC
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
#else
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
#endif
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
    // generator-fused variant (gram_i8_fused_gen_kernel): X is not read; the slicers generate its columns (same bits as the prologue's fill)
    const unsigned long long* gen_counter; uint64_t gen_seed; double gen_sc53; int64_t gen_inst_offset;
    unsigned long long* diag;      // host-mapped record of the first wait that timed out (32 words; null: plain trap): see fused_wait
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

// Watchdog (build with -DFZ_WATCHDOG; tools/stress_cfg3.py and the probe use it to turn a hang into a message).  The mbarrier waits of the
// working warps stay the plain two-instruction poll loops: any bound inside them -- a counter, a clock read, a sleep, a suspend-time hint --
// cost 15-19 % of the kernel, and even the milestone stores below cost 10 % (the kernel's ~9100 instructions sit at an instruction-cache
// cliff: profiles/r02_h_fused_wait_variants.txt), so the shipped build has neither.  With FZ_WATCHDOG one of the two otherwise idle warps
// watches the CTA: every working warp leaves a milestone (instance counter, phase) in shared memory at a handful of points per instance; the
// watchdog lane sleeps 20 us at a time and, when NO milestone of the CTA has moved for ~2 s (a whole launch takes milliseconds), copies the
// milestones to the host-mapped diagnostic record and traps -- pmb_synchronize appends the record to its error message.
#ifdef FZ_WATCHDOG
__shared__ uint32_t fused_progress[24];   // per warp: (instance counter << 4) | phase, written by lane 0
__shared__ uint32_t fused_done;           // roles that have finished their instance loop (producer, MMA issuer / its idle twin, slicers, epilogue)
#define FZ_MARK(ii_, phase_) do { if ((threadIdx.x & 31) == 0) *reinterpret_cast<volatile uint32_t*>(&fused_progress[threadIdx.x >> 5]) = ((uint32_t)(ii_) << 4) | (uint32_t)(phase_); } while (0)
#define FZ_ROLE_DONE() do { if ((threadIdx.x & 31) == 0) atomicAdd(&fused_done, 1u); } while (0)
__device__ __noinline__ void fused_watchdog(unsigned long long* diag, uint32_t roles) {
    uint32_t last = 0;
    long long t_last = clock64();
    while (*reinterpret_cast<volatile uint32_t*>(&fused_done) < roles) {
        __nanosleep(20000);
        uint32_t sum = 0;
        for (int w = 0; w < 20; ++w) sum = sum * 31u + *reinterpret_cast<volatile uint32_t*>(&fused_progress[w]);
        if (sum != last) { last = sum; t_last = clock64(); continue; }
        if (clock64() - t_last > 4000000000ll) {
            if (diag && atomicCAS(diag, 0ull, 1ull | ((unsigned long long)blockIdx.x << 32)) == 0ull) {
                for (int w = 0; w < 20; ++w) diag[1 + w] = *reinterpret_cast<volatile uint32_t*>(&fused_progress[w]);
                __threadfence_system();
            }
            __trap();
        }
    }
}
#else
#define FZ_MARK(ii_, phase_) do { } while (0)
#define FZ_ROLE_DONE() do { } while (0)
#endif
__device__ __forceinline__ void fused_wait(uint64_t* bar, uint32_t parity) { i8::mb_wait(bar, parity); }

#define FZ_STR2(x) #x
#define FZ_STR(x) FZ_STR2(x)
#define FZ_GEN 0
#define FZ_KERNEL_NAME gram_i8_fused_kernel
#include "kernels_gram_i8_fused_body.inc"
#undef FZ_GEN
#undef FZ_KERNEL_NAME
#define FZ_GEN 1
#define FZ_KERNEL_NAME gram_i8_fused_gen_kernel
#include "kernels_gram_i8_fused_body.inc"
#undef FZ_GEN
#undef FZ_KERNEL_NAME

// ---- host ----------------------------------------------------------------------------------------------------------------------
static int g_fused_max_pairs = -1;

void configure_gram_i8_fused_kernel() {
    cudaFuncSetAttribute(gram_i8_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fz::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
    cudaFuncSetAttribute(gram_i8_fused_gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fz::kSmemBytes);
    cudaFuncSetAttribute(gram_i8_fused_gen_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
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

const char* gram_i8_fused_diag_text(const unsigned long long* diag, char* buf, size_t len) {
    if (!diag || !diag[0] || !buf || len == 0) return "";
    int o = snprintf(buf, len, "fused int8 Gram kernel: no progress for ~2 s in CTA %u; (instance counter.phase) per warp -- epilogue 0-7, slicers 8-15, producer 16, MMA issuer 17:",
                     (unsigned)(diag[0] >> 32));
    for (int w = 0; w < 18 && o > 0 && (size_t)o < len; ++w) o += snprintf(buf + o, len - (size_t)o, " %u.%u", (unsigned)(diag[1 + w] >> 4), (unsigned)(diag[1 + w] & 15));
    return buf;
}

// the generator-fused variant covers the densely stored full-size instance (k = ldx = 256 rows: cfg3)
bool gram_i8_fused_can_generate(const GramArgs& g) { return g.k == 256 && g.ldx == 256 && g.x_bstride == 256 * g.n && !g.gate; }

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
    a.diag = sc.diag;
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
    if (sc.gen_counter) {   // the planner removed the operand's fill from the prologue: this variant MUST run
        if (!gram_i8_fused_can_generate(g)) return false;
        a.gen_counter = sc.gen_counter; a.gen_seed = sc.gen_seed; a.gen_sc53 = sc.gen_scale * (1.0 / 9007199254740992.0); a.gen_inst_offset = sc.gen_inst_offset;
        return cudaLaunchKernelEx(&cfg, gram_i8_fused_gen_kernel, a) == cudaSuccess;
    }
    return cudaLaunchKernelEx(&cfg, gram_i8_fused_kernel, a) == cudaSuccess;
}

}  // namespace pmb
