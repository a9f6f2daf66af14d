// K2 (TMA, warp-specialised) -- fp64 Gram / SYRK-shaped contraction on the DMMA tensor pipe.
//
// Same reference semantics as kernels_gram.cu (vecdot! + muladd!, functions.jl:548-576,702-709, then the symmetric MOI
// Hessian of moi_interop.jl:45-62:  H = scale * X'X written to both triangles), restructured for Blackwell:
//   * one producer warp: a single elected thread issues 3-D TMA tile loads (cp.async.bulk.tensor, SASS UTMALDG) of the
//     K-major operand panels straight from the Parameter buffer into a 6-stage shared-memory ring, completion on mbarriers.
//     Box = 16 k x 128 columns (one full 128-byte line per column) with the hardware 128-byte swizzle; together with a fixed
//     permutation of the 8 rows of every MMA fragment (pi = 0,2,4,6,1,3,5,7) every fragment load is bank-conflict-free.
//     Ragged edges (k, n not multiples of the tile) are zero-filled by the TMA unit;
//   * eight consumer warps (64x32 each, 4x4 mma.m16n8k4.f64 = SASS DMMA) wait on the "full" mbarrier of a stage and release
//     it through the "empty" mbarrier: no block-wide barrier and no address arithmetic in the MMA instruction stream, so the
//     two warps of a scheduler drift out of phase and keep the DMMA pipe fed;
//   * persistent CTAs walk "units" = (instance, upper-triangular 128x128 tile, k-slice).  Tiles of the last partial wave are
//     split along k so every SM finishes together: each slice parks its partial tile in a workspace, and the slice that arrives
//     last (one atomic counter per tile) sums the partials in slice order -- nobody waits, and the summation order is fixed, so
//     results are bit-identical run to run.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "pmb_kernels.h"

namespace pmb {

constexpr int kTBM = 128;  // large-problem tile; small problems (less than one wave of 128-tiles) use 64 x 64 tiles
constexpr int kTKT = 16;
constexpr int kTStages = 6;
constexpr int kTConsumerWarps = 8;
constexpr int kTThreads = (kTConsumerWarps + 4) * 32;  // 2 consumer warpgroups + 1 producer warpgroup (setmaxnreg is per warpgroup)
constexpr int kTSmemBytes = kTStages * 2 * kTBM * kTKT * 8 + 1024 + kTStages * 128;  // 192 KB ring + 1 KB alignment slack + g slots

struct GramTmaKernelArgs {
    CUtensorMap tmx, tmy;
    int64_t k, n, p;
    double* out; int64_t ldo, o_bstride;
    const int32_t* map;
    AlphaDev scale;
    int32_t accumulate;
    int32_t tiles_m, tiles_n, kiters;
    int64_t tiles_per_inst, total_tiles;
    int64_t full_tiles;   // tiles [0, full_tiles) are one unit each; the rest are cut into nslices k-slices
    int32_t nslices;
    const int32_t* gate;  // see GramArgs::gate
    int32_t gate_cols;
    double* packed_dev;   // see GramArgs::packed_*
    double* packed_host;
    int64_t packed_bstride;
    const double* gvec; int64_t g_bstride;   // see GramArgs::gvec (fused q / c0)
    double* qout; int64_t q_bstride;
    double* c0out;
    AlphaDev qscale, cscale;
    double* qworkspace;
    const unsigned long long* cert_stats;  // see GramArgs::cert_stats: instances certified for the int8 path are skipped here
    double cert_tol;
    int32_t cert_force;
    const unsigned long long* cert_none;   // number of instances the int8 launch left uncertified (null: unknown); 0 -> nothing to do here
    int32_t* counters;    // [total_tiles - full_tiles] arrival counters, zero between launches
    double* workspace;    // [(total_tiles - full_tiles) * nslices][64][256] partial accumulators
    int64_t total_units;
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(s_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(s_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gmem_src),
                 "r"(bytes), "r"(s_u32(bar))
                 : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma_16x8x4(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a0), "d"(a1), "d"(b0));
}
// the certified int8-slice kernels (kernels_gram_i8.cu) own this instance: nothing to do here
__device__ __forceinline__ bool skip_certified(const GramTmaKernelArgs& a, int64_t inst) {
    return a.cert_stats && (a.cert_force || gram_i8_level(a.cert_stats, inst) == 5 || gram_i8_level(a.cert_stats, inst) == 6);
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kTConsumerWarps * 32) : "memory"); }

struct Unit {
    int64_t inst;
    int ti, tj;
    int kt0, kt1;    // k-tile range
    int slice, nslices;
    int64_t split_id;  // index of the split tile: counters[split_id], workspace[(split_id * nslices + slice)]
};

template <bool SYRK>
__device__ __forceinline__ Unit decode_unit(const GramTmaKernelArgs& a, int64_t u) {
    Unit w;
    int64_t tile;
    if (u < a.full_tiles) {
        tile = u; w.slice = 0; w.nslices = 1; w.split_id = 0;
    } else {
        int64_t v = u - a.full_tiles;
        w.split_id = v / a.nslices;
        w.slice = (int)(v - w.split_id * a.nslices);
        w.nslices = a.nslices;
        tile = a.full_tiles + w.split_id;
    }
    w.inst = tile / a.tiles_per_inst;
    int64_t tl = tile - w.inst * a.tiles_per_inst;
    if (SYRK) {  // upper-triangular tile (ti <= tj), column-major enumeration: tile column tj holds tiles ti = 0..tj
        int tj = 0;
        while (tl >= tj + 1) { tl -= tj + 1; ++tj; }
        w.ti = (int)tl; w.tj = tj;
    } else {
        w.ti = (int)(tl / a.tiles_n);
        w.tj = (int)(tl - (int64_t)w.ti * a.tiles_n);
    }
    w.kt0 = (int)((int64_t)a.kiters * w.slice / w.nslices);
    w.kt1 = (int)((int64_t)a.kiters * (w.slice + 1) / w.nslices);
    return w;
}

template <bool SYRK, bool ADD, int BM, bool PACKED>
__device__ __forceinline__ void store_tile(const GramTmaKernelArgs& a, double (&acc)[BM / 32][BM / 32][4], double* out, double sc, int64_t i0,
                                           int64_t j0, bool diag, int wm, int wn, int pg, int t, int64_t inst) {
    double* pk_dev = (PACKED && a.packed_dev) ? a.packed_dev + inst * a.packed_bstride : nullptr;
    double* pk_host = (PACKED && a.packed_host) ? a.packed_host + inst * a.packed_bstride : nullptr;
#pragma unroll
    for (int mi = 0; mi < BM / 32; ++mi) {
#pragma unroll
        for (int ni = 0; ni < BM / 32; ++ni) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int cj = 2 * t + (r & 1);  // accumulator column index inside the 8-wide fragment -> panel column pi(cj)
                const int64_t m = i0 + wm * (BM / 2) + mi * 16 + pg + (r >> 1) * 8;
                const int64_t nn = j0 + wn * (BM / 4) + ni * 8 + (((2 * cj) & 7) + (cj >> 2));
                if (m < a.n && nn < a.p) {
                    const double v = sc * acc[mi][ni][r];
                    if (SYRK) {
                        const int64_t om = a.map ? a.map[m] : m, on = a.map ? a.map[nn] : nn;
                        double* p1 = out + om + on * a.ldo;
                        *p1 = ADD ? *p1 + v : v;
                        if (PACKED && (!diag || m <= nn)) {  // each unordered pair exactly once
                            const int64_t lo = om < on ? om : on, hi = om < on ? on : om;
                            const int64_t idx = hi * (hi + 1) / 2 + lo;
                            if (pk_dev) pk_dev[idx] = v;
                            if (pk_host) __stcs(pk_host + idx, v);
                        }
                        if (!diag) {
                            double* p2 = out + on + om * a.ldo;
                            *p2 = ADD ? *p2 + v : v;
                        }
                    } else {
                        double* p1 = out + m + nn * a.ldo;
                        *p1 = ADD ? *p1 + v : v;
                    }
                }
            }
        }
    }
}

template <bool SYRK, int BM, bool FUSEQ>
__global__ void __launch_bounds__(kTThreads, 1) gram_tma_kernel(const __grid_constant__ GramTmaKernelArgs a) {
    constexpr int MI = BM / 32, NI = BM / 32;      // warp tile (BM/2) x (BM/4) = MI x NI fragments of 16 x 8
    constexpr int kPanelBytes = BM * kTKT * 8;     // one operand panel of one stage: BM columns x 128 bytes
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kTStages];
    __shared__ __align__(8) uint64_t empty_bar[kTStages];
    __shared__ int arrival_flag;
    __shared__ double qred[kTConsumerWarps * 32 + 8];   // fused q / c0 partial sums of one unit
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ring = (s_u32(smem_raw) + 1023u) & ~1023u;  // the 128-byte swizzle pattern repeats every 1024 bytes
    const uint32_t gring = ring + (uint32_t)kTStages * 2 * kTBM * kTKT * 8;  // 128-byte slots: 16 entries of gvec per stage
    constexpr bool fuse_q = SYRK && FUSEQ;  // compiled out of the large-problem instantiations: it costs MMA-loop scheduling freedom

    if (SYRK && a.cert_stats && (a.cert_force || (a.cert_none && __ldcg(a.cert_none) == 0ull))) return;   // every instance was certified: leave at once
    if (SYRK && a.total_tiles == a.tiles_per_inst && skip_certified(a, 0)) return;   // one instance, owned by the int8 path: nothing to do
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kTStages; ++s) { mb_init(&full_bar[s], 1); mb_init(&empty_bar[s], kTConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= kTConsumerWarps) {
        // ===== producer warpgroup: gives its registers back; one elected thread drives the TMA ring =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == kTConsumerWarps && lane == 0) {
            uint32_t it = 0;
            for (int64_t u = blockIdx.x; u < a.total_units; u += gridDim.x) {
                const Unit w = decode_unit<SYRK>(a, u);
                if (SYRK && skip_certified(a, w.inst)) continue;
                const bool diag = SYRK && w.ti == w.tj;
                if (SYRK && a.gate) {
                    // wait until the column chunk holding this tile's right-most panel has been uploaded (ti <= tj: the left
                    // panel is in the same or an earlier chunk); the flag is written by the copy engine after the data
                    const int32_t* flag = a.gate + ((int64_t)(w.tj + 1) * BM - 1 < a.n ? ((int64_t)(w.tj + 1) * BM - 1) : (a.n - 1)) / a.gate_cols;
                    // The flag is written by a copy node of a sibling branch of the same graph: CUDA does not promise that it runs concurrently
                    // with this kernel, so the wait is BOUNDED (~4 s); a stalled upload aborts the kernel loudly instead of hanging the update
                    int v;
                    unsigned int spins = 0;
                    do {
                        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if (!v) { __nanosleep(400); if (++spins > 10000000u) __trap(); }
                    } while (!v);
                    asm volatile("fence.proxy.async;" ::: "memory");  // order the generic-proxy flag read before the async-proxy TMA reads
                }
                for (int kt = w.kt0; kt < w.kt1; ++kt, ++it) {
                    const int s = it % kTStages;
                    mb_wait(&empty_bar[s], ((it / kTStages) & 1) ^ 1);
                    const uint32_t stage = ring + (uint32_t)s * 2 * kPanelBytes;
                    const int64_t krem = a.k - (int64_t)kt * kTKT;
                    const uint32_t gbytes = (diag && fuse_q) ? (uint32_t)((krem < kTKT ? krem : kTKT) * 8) : 0u;
                    mb_expect_tx(&full_bar[s], (diag ? kPanelBytes : 2 * kPanelBytes) + gbytes);
                    if (gbytes) bulk_load_1d(gring + (uint32_t)s * 128, a.gvec + w.inst * a.g_bstride + (int64_t)kt * kTKT, gbytes, &full_bar[s]);
                    tma_load_3d(stage, &a.tmx, kt * kTKT, w.ti * BM, (int)w.inst, &full_bar[s]);
                    if (!diag) tma_load_3d(stage + kPanelBytes, &a.tmy, kt * kTKT, w.tj * BM, (int)w.inst, &full_bar[s]);
                }
            }
        }
        return;
    }

    // ===== consumers: 8 warps (2 warpgroups), warp tile 64 x 32 =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    // MMA fragment row g lives in panel column (8-group base) + pi(g): with the 128-byte swizzle (16-byte chunk index XOR
    // (column & 7)) the 16 lanes of each half-warp then hit 16 distinct 8-byte bank pairs.
    const int pg = ((2 * g) & 7) + (g >> 2);
    uint32_t koff[kTKT / 4];  // byte offset of this thread's k element inside a swizzled 128-byte column line, per k-block
#pragma unroll
    for (int kb = 0; kb < kTKT / 4; ++kb) koff[kb] = (uint32_t)((((kb * 2 + (t >> 1)) ^ pg) << 4) + (t & 1) * 8);
    const uint32_t a_row = (uint32_t)((wm * (BM / 2) + pg) * 128);
    const uint32_t b_row = (uint32_t)((wn * (BM / 4) + pg) * 128);
    uint32_t it = 0;
    for (int64_t u = blockIdx.x; u < a.total_units; u += gridDim.x) {
        const Unit w = decode_unit<SYRK>(a, u);
        if (SYRK && skip_certified(a, w.inst)) continue;
        const bool diag = SYRK && w.ti == w.tj;
        double acc[MI][NI][4];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[mi][ni][r] = 0.0;

        // fused linear part: thread -> (column qcol of the diagonal tile, k-group qkg)
        constexpr int KG = (kTConsumerWarps * 32) / BM, KPT = kTKT / KG;
        const int qcol = tid % BM, qkg = tid / BM;
        const bool do_q = fuse_q && diag;
        double qacc = 0.0, cacc = 0.0;

        for (int kt = w.kt0; kt < w.kt1; ++kt, ++it) {
            const int s = it % kTStages;
            mb_wait(&full_bar[s], (it / kTStages) & 1);
            if (do_q) {
                const uint32_t sX = ring + (uint32_t)s * 2 * kPanelBytes + (uint32_t)qcol * 128;
                const uint32_t sG = gring + (uint32_t)s * 128;
#pragma unroll
                for (int j = 0; j < KPT; ++j) {
                    const int kk = qkg * KPT + j;
                    if ((int64_t)kt * kTKT + kk < a.k) {
                        const double gv = lds_f64(sG + kk * 8);
                        qacc += lds_f64(sX + ((((kk >> 1) ^ (qcol & 7))) << 4) + (kk & 1) * 8) * gv;
                        if (qcol == 0) cacc += gv * gv;
                    }
                }
            }
            const uint32_t sA = ring + (uint32_t)s * 2 * kPanelBytes + a_row;
            const uint32_t sB = ring + (uint32_t)s * 2 * kPanelBytes + (diag ? 0u : (uint32_t)kPanelBytes) + b_row;
#pragma unroll
            for (int kb = 0; kb < kTKT / 4; ++kb) {
                double af[MI][2], bf[NI];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    af[mi][0] = lds_f64(sA + mi * 16 * 128 + koff[kb]);
                    af[mi][1] = lds_f64(sA + (mi * 16 + 8) * 128 + koff[kb]);
                }
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) bf[ni] = lds_f64(sB + ni * 8 * 128 + koff[kb]);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) dmma_16x8x4(acc[mi][ni], af[mi][0], af[mi][1], bf[ni]);
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&empty_bar[s]);  // this warp is done reading the stage
        }

        // ---- fused linear part: reduce the k-groups of this unit ----
        double qv = 0.0, cv = 0.0;
        if (do_q) {
            qred[tid] = qacc;
            if (qcol == 0) qred[kTConsumerWarps * 32 + qkg] = cacc;
            consumer_bar();
            if (tid < BM) {
#pragma unroll
                for (int gq = 0; gq < KG; ++gq) qv += qred[gq * BM + tid];
            }
            if (tid == 0) {
#pragma unroll
                for (int gq = 0; gq < KG; ++gq) cv += qred[kTConsumerWarps * 32 + gq];
            }
            consumer_bar();  // qred is reused by the next unit
            if (w.nslices > 1) {
                double* qw = a.qworkspace + (w.split_id * w.nslices + w.slice) * (int64_t)(BM + 1);
                if (tid < BM) __stcg(qw + tid, qv);
                if (tid == 0) __stcg(qw + BM, cv);
            }
        }

        // ---- k-slices: park the partial, the last arriver sums all partials of the tile in slice order ----
        if (w.nslices > 1) {
            double* ws = a.workspace + ((w.split_id * w.nslices + w.slice) * (MI * NI * 4)) * (int64_t)(kTConsumerWarps * 32) + tid;
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                    for (int r = 0; r < 4; ++r) __stcg(ws + ((mi * NI + ni) * 4 + r) * (kTConsumerWarps * 32), acc[mi][ni][r]);
            __threadfence();
            consumer_bar();
            if (tid == 0) {
                const int old = atomicAdd(a.counters + w.split_id, 1);
                const int last = old == w.nslices - 1;
                if (last) a.counters[w.split_id] = 0;  // ready for the next launch (graph replay)
                arrival_flag = last;
            }
            consumer_bar();
            if (!arrival_flag) continue;  // uniform across the consumer warps
            __threadfence();
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[mi][ni][r] = 0.0;
            if (do_q) {
                qv = 0.0; cv = 0.0;
                const double* qw = a.qworkspace + (w.split_id * w.nslices) * (int64_t)(BM + 1);
                for (int sl = 0; sl < w.nslices; ++sl, qw += BM + 1) {
                    if (tid < BM) qv += __ldcg(qw + tid);
                    if (tid == 0) cv += __ldcg(qw + BM);
                }
            }
            const double* wr = a.workspace + (w.split_id * w.nslices * (MI * NI * 4)) * (int64_t)(kTConsumerWarps * 32) + tid;
            for (int sl = 0; sl < w.nslices; ++sl, wr += (MI * NI * 4) * (kTConsumerWarps * 32)) {
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[mi][ni][r] += __ldcg(wr + ((mi * NI + ni) * 4 + r) * (kTConsumerWarps * 32));
            }
        }
        // ---- fused linear part: one writer per column of q (this diagonal tile), tile (0,0) owns c0 ----
        if (do_q) {
            const int64_t col = (int64_t)w.ti * BM + tid;
            if (tid < BM && col < a.n) {
                const double qs = a.qscale.ptr ? a.qscale.c * a.qscale.ptr[w.inst * a.qscale.bstride] : a.qscale.c;
                double* dst = a.qout + w.inst * a.q_bstride + (a.map ? a.map[col] : col);
                *dst += qs * qv;
            }
            if (tid == 0 && w.ti == 0) {
                const double cs = a.cscale.ptr ? a.cscale.c * a.cscale.ptr[w.inst * a.cscale.bstride] : a.cscale.c;
                a.c0out[w.inst] += cs * cv;
            }
        }
        // ---- epilogue: scale, (scatter by variable map), both triangles ----
        const double sc = a.scale.ptr ? a.scale.c * a.scale.ptr[w.inst * a.scale.bstride] : a.scale.c;
        double* out = a.out + w.inst * a.o_bstride;
        const int64_t i0 = (int64_t)w.ti * BM, j0 = (int64_t)w.tj * BM;
        if (a.accumulate) store_tile<SYRK, true, BM, false>(a, acc, out, sc, i0, j0, diag, wm, wn, pg, t, w.inst);
        else if (SYRK && (a.packed_dev || a.packed_host)) store_tile<SYRK, false, BM, true>(a, acc, out, sc, i0, j0, diag, wm, wn, pg, t, w.inst);
        else store_tile<SYRK, false, BM, false>(a, acc, out, sc, i0, j0, diag, wm, wn, pg, t, w.inst);
    }
}

void configure_gram_tma_kernels() {
    cudaFuncSetAttribute(gram_tma_kernel<true, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmemBytes);
    cudaFuncSetAttribute(gram_tma_kernel<false, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmemBytes);
    cudaFuncSetAttribute(gram_tma_kernel<true, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmemBytes);
    cudaFuncSetAttribute(gram_tma_kernel<true, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmemBytes);
    cudaFuncSetAttribute(gram_tma_kernel<false, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmemBytes);
}

// ---- host: tensor maps and the unit decomposition ------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// X: k x n column-major (ld), `batch` instances `bstride` elements apart.  Tensor (k, n, batch), box (16, 128, 1), 128-byte swizzle.
static bool make_panel_map(CUtensorMap* tm, const double* X, int64_t ld, int64_t bstride, int64_t k, int64_t n, int64_t batch, int bm) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    if (ld % 2 != 0 || ((uintptr_t)X & 15) != 0 || (batch > 1 && bstride % 2 != 0) || k == 0 || n == 0) return false;
    cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)n, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)(batch > 1 && bstride > 0 ? bstride : ld * n) * 8};
    cuuint32_t box[3] = {(cuuint32_t)kTKT, (cuuint32_t)bm, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// unit decomposition: whole waves of tiles run unsplit; the tiles of the last partial wave are cut into k-slices so that
// (tail tiles x slices) fills whole waves again.  Fewer tiles than SMs: everything is split.
GramSplitPlan plan_gram_split(bool syrk, int64_t k, int64_t n, int64_t p, int64_t batch, int sm_count) {
    GramSplitPlan pl{1, 0, 0, kTBM};
    auto tiles = [&](int bm) {
        const int64_t tm = (n + bm - 1) / bm, tn = (p + bm - 1) / bm;
        return (syrk ? tn * (tn + 1) / 2 : tm * tn) * batch;
    };
    // less than one wave of 128 x 128 tiles: the problem is latency-bound, spread it over more SMs with 64 x 64 tiles.
    // Small instances (n <= 256: at most 3 upper 128-tiles, two of them diagonal and half redundant): 64-tiles compute 17 % fewer
    // redundant entries -- measured 6.50 ms vs 7.05 ms at cfg3.
    if (tiles(kTBM) < sm_count || std::max(n, p) <= 256) pl.bm = 64;
    if (const char* e = getenv("PMB_GRAM_BM")) { if (atoi(e) == 64) pl.bm = 64; if (atoi(e) == 128) pl.bm = kTBM; }  // experiments
    const int64_t total = tiles(pl.bm);
    const int64_t kiters = (k + kTKT - 1) / kTKT;
    const int64_t G = sm_count;
    const int64_t tail = total % G;
    if (tail == 0) return pl;
    const int64_t max_slices = std::min<int64_t>(8, kiters / 8);  // >= 8 k-tiles (128 k) per slice, workspace <= 8 tiles per SM
    double best = 1e30;
    for (int64_t s = 1; s <= max_slices; ++s) {
        const double waves = (double)((tail * s + G - 1) / G) / (double)s;  // tail time in tile-times
        const double cost = waves + 0.012 * (double)s;                      // parking + summing the partials is not free
        if (cost < best - 1e-12) { best = cost; pl.nslices = (int32_t)s; }
    }
    if (pl.nslices > 1) {
        pl.split_tiles = tail;
        pl.workspace_doubles = tail * pl.nslices * (int64_t)pl.bm * pl.bm;
    }
    return pl;
}

// Returns false when the TMA path does not apply (alignment / shared constant operand with batch > 1); the caller falls back
// to the cp.async kernel of kernels_gram.cu.
static bool launch_tma_impl(bool syrk, const double* X, int64_t ldx, int64_t xbs, const double* Y, int64_t ldy, int64_t ybs, int64_t k,
                            int64_t n, int64_t p, double* out, int64_t ldo, int64_t obs, const int32_t* map, AlphaDev scale, int accumulate,
                            const GramScratch& sc, int64_t batch, int sm_count, cudaStream_t stream, const int32_t* gate = nullptr,
                            int32_t gate_cols = 0, double* packed_dev = nullptr, double* packed_host = nullptr, int64_t packed_bstride = 0,
                            const GramArgs* fused = nullptr) {
    if (batch > 1 && (xbs == 0 || ybs == 0)) return false;
    GramTmaKernelArgs a{};
    const int bm = sc.plan.bm == 64 ? 64 : kTBM;
    if (!make_panel_map(&a.tmx, X, ldx, xbs, k, n, batch, bm)) return false;
    if (!make_panel_map(&a.tmy, Y, ldy, ybs, k, p, batch, bm)) return false;
    a.k = k; a.n = n; a.p = p; a.out = out; a.ldo = ldo; a.o_bstride = obs; a.map = map; a.scale = scale; a.accumulate = accumulate;
    a.tiles_m = (int)((n + bm - 1) / bm);
    a.tiles_n = (int)((p + bm - 1) / bm);
    a.kiters = (int)((k + kTKT - 1) / kTKT);
    a.tiles_per_inst = syrk ? (int64_t)a.tiles_n * (a.tiles_n + 1) / 2 : (int64_t)a.tiles_m * a.tiles_n;
    a.total_tiles = a.tiles_per_inst * batch;
    if (a.total_tiles == 0) return true;
    const bool split = sc.plan.nslices > 1 && sc.counters && sc.workspace;
    a.nslices = split ? sc.plan.nslices : 1;
    a.full_tiles = split ? a.total_tiles - sc.plan.split_tiles : a.total_tiles;
    a.counters = sc.counters;
    a.workspace = sc.workspace;
    a.gate = gate; a.gate_cols = gate_cols;
    a.packed_dev = packed_dev; a.packed_host = packed_host; a.packed_bstride = packed_bstride;
    if (fused && fused->gvec && sc.qworkspace_ok) {
        a.gvec = fused->gvec; a.g_bstride = fused->g_bstride; a.qout = fused->qout; a.q_bstride = fused->q_bstride; a.c0out = fused->c0out;
        a.qscale = fused->qscale; a.cscale = fused->cscale; a.qworkspace = sc.qworkspace;
    }
    if (fused) { a.cert_stats = fused->cert_stats; a.cert_tol = fused->cert_tol; a.cert_force = fused->cert_force; a.cert_none = fused->cert_none; }
    a.total_units = a.full_tiles + (a.total_tiles - a.full_tiles) * a.nslices;
    const int grid = (int)(a.total_units < sm_count ? a.total_units : sm_count);
    if (bm == 64) {
        if (syrk && a.gvec) gram_tma_kernel<true, 64, true><<<grid, kTThreads, kTSmemBytes, stream>>>(a);
        else if (syrk) gram_tma_kernel<true, 64, false><<<grid, kTThreads, kTSmemBytes, stream>>>(a);
        else gram_tma_kernel<false, 64, false><<<grid, kTThreads, kTSmemBytes, stream>>>(a);
    } else {
        if (syrk) gram_tma_kernel<true, 128, false><<<grid, kTThreads, kTSmemBytes, stream>>>(a);
        else gram_tma_kernel<false, 128, false><<<grid, kTThreads, kTSmemBytes, stream>>>(a);
    }
    return true;
}

bool gram_tma_applicable(const double* X, int64_t ld, int64_t bstride, int64_t k, int64_t n, int64_t batch) {
    return encode_fn() != nullptr && ld % 2 == 0 && ((uintptr_t)X & 15) == 0 && !(batch > 1 && (bstride % 2 != 0 || bstride == 0)) && k > 0 && n > 0;
}

// the fused linear part exists only in the 64-tile instantiation: latency-bound problems (less than one wave of 128-tiles) and small
// instances in large batches (n <= 256; cfg3: 9.42 -> 9.09 ms per sweep, the Gram gets 14 % slower but the gemv + dot launch that
// re-read every A goes away).  For large single problems the fusion costs more MMA-loop throughput than the 34 us gemv it removes.
bool gram_fused_linear_supported(int64_t k, int64_t n, int64_t batch, int sm_count) {
    const GramSplitPlan pl = plan_gram_split(true, k, n, n, batch, sm_count);
    return pl.bm == 64;
}

bool launch_gram_syrk_tma(const GramArgs& g, const GramScratch& sc, int64_t batch, int sm_count, cudaStream_t stream) {
    return launch_tma_impl(true, g.X, g.ldx, g.x_bstride, g.X, g.ldx, g.x_bstride, g.k, g.n, g.n, g.H, g.ldh, g.h_bstride, g.map, g.scale,
                           g.accumulate, sc, batch, sm_count, stream, g.gate, g.gate_cols, g.accumulate ? nullptr : g.packed_dev,
                           g.accumulate ? nullptr : g.packed_host, g.packed_bstride, &g);
}

bool launch_gemm_tn_tma(const GemmTNArgs& g, const GramScratch& sc, int64_t batch, int sm_count, cudaStream_t stream) {
    return launch_tma_impl(false, g.X, g.ldx, g.x_bstride, g.Y, g.ldy, g.y_bstride, g.k, g.n, g.p, g.M, g.ldm, g.m_bstride, nullptr,
                           AlphaDev{1.0, nullptr, 0}, 0, sc, batch, sm_count, stream);
}

}  // namespace pmb
