// K2 -- fp64 Gram / SYRK-shaped contraction on the DMMA tensor pipe (mma.sync f64; tcgen05 has no f64 kind).
//
// Device restatement of   vecdot!(dest::QuadraticFunction, r, r)  with r = A*x - b
//   _vecdot!  functions.jl:702-709  ->  one muladd!(dest, r_i, r_i) per residual row
//   muladd!   functions.jl:548-576  ->  N^2 QuadraticTerms  a_ri * a_rj * x_i * x_j  per row, never summed
//   update!   moi_interop.jl:45-62  ->  diagonal coefficients doubled; MOI sums duplicates into the symmetric Hessian
// whose dense meaning is  H = 2 * A'A  (test/model.jl:84).  Only the upper-triangular tiles are computed; each
// tile is written to both triangles.  A is read in place from the Parameter buffer (column-major, k contiguous),
// so both MMA operands are K-major and no residual / term list is ever materialised.
//
// Tiling: CTA tile 128x128, 8 warps of 64x32 (4x4 mma.m16n8k4 each, SASS DMMA.8x8x4), K step 16 through a
// 4-stage cp.async ring.  Shared layout [k/4][column][4 doubles]: an MMA fragment load touches 8 columns x 32 B
// = all 32 banks exactly once per half-warp (conflict-free without padding or swizzle).
#include "pmb_kernels.h"

namespace pmb {

constexpr int kBM = 128;           // tile edge (both operands are column panels of X / Y)
constexpr int kKT = 16;            // k per stage
constexpr int kStages = 4;
constexpr int kGramThreads = 256;
constexpr int kPanelDoubles = (kKT / 4) * kBM * 4;                       // one operand, one stage: 2048 doubles = 16 KB
constexpr int kGramSmemBytes = kStages * 2 * kPanelDoubles * 8;          // 128 KB

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_m16n8k4(double (&c)[4], double a0, double a1, double b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a0), "d"(a1), "d"(b0));
}

// Load one 128-column x 16-k panel of a column-major matrix into the [k/4][col][4] layout (zero-filled out of range).
template <bool ALIGN16>
__device__ __forceinline__ void load_panel(double* sm, const double* __restrict__ X, int64_t ld, int64_t k, int64_t n,
                                           int64_t col0, int64_t k0, int tid) {
    if (ALIGN16) {
#pragma unroll
        for (int i = 0; i < (kBM * kKT / 2) / kGramThreads; ++i) {  // 1024 chunks of 2 doubles
            int c = tid + i * kGramThreads;
            int col = c >> 3, kc = c & 7;
            int64_t gk = k0 + kc * 2, gc = col0 + col;
            int bytes = 0;
            if (gc < n) bytes = gk + 1 < k ? 16 : (gk < k ? 8 : 0);
            const double* src = bytes ? X + gk + gc * ld : X;
            cp_async16(sm + (((kc >> 1) * kBM + col) * 4 + (kc & 1) * 2), src, bytes);
        }
    } else {
#pragma unroll
        for (int i = 0; i < (kBM * kKT) / kGramThreads; ++i) {  // 2048 single doubles
            int c = tid + i * kGramThreads;
            int col = c >> 4, kk = c & 15;
            int64_t gk = k0 + kk, gc = col0 + col;
            int bytes = (gc < n && gk < k) ? 8 : 0;
            const double* src = bytes ? X + gk + gc * ld : X;
            cp_async8(sm + (((kk >> 2) * kBM + col) * 4 + (kk & 3)), src, bytes);
        }
    }
}

struct GramKernelArgs {
    const double* X; int64_t ldx, x_bstride;
    const double* Y; int64_t ldy, y_bstride;   // SYRK: Y == X
    int64_t k, n, p;                            // X: k x n, Y: k x p
    double* out; int64_t ldo, o_bstride;
    const int32_t* map;
    AlphaDev scale;
    int32_t accumulate;
    int32_t tiles_m, tiles_n;
    int64_t tiles_per_inst, total_work;
};

template <bool SYRK, bool ALIGN16>
__global__ void __launch_bounds__(kGramThreads, 1) gram_dmma_kernel(GramKernelArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int64_t kiters = (a.k + kKT - 1) / kKT;

    for (int64_t work = blockIdx.x; work < a.total_work; work += gridDim.x) {
        const int64_t inst = work / a.tiles_per_inst;
        int64_t tl = work - inst * a.tiles_per_inst;
        int ti, tj;
        if (SYRK) {  // upper-triangular tile (ti <= tj), row-major enumeration
            ti = 0;
            int rowlen = a.tiles_n;
            while (tl >= rowlen) { tl -= rowlen; --rowlen; ++ti; }
            tj = ti + (int)tl;
        } else {
            ti = (int)(tl / a.tiles_n);
            tj = (int)(tl - (int64_t)ti * a.tiles_n);
        }
        const bool diag = SYRK && ti == tj;
        const double* X = a.X + inst * a.x_bstride;
        const double* Y = a.Y + inst * a.y_bstride;
        const int64_t i0 = (int64_t)ti * kBM, j0 = (int64_t)tj * kBM;

        double acc[4][4][4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[mi][ni][r] = 0.0;

        auto issue = [&](int64_t kt) {
            if (kt < kiters) {
                double* sA = smem + (kt % kStages) * 2 * kPanelDoubles;
                load_panel<ALIGN16>(sA, X, a.ldx, a.k, a.n, i0, kt * kKT, tid);
                if (!diag) load_panel<ALIGN16>(sA + kPanelDoubles, Y, a.ldy, a.k, a.p, j0, kt * kKT, tid);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int s = 0; s < kStages - 1; ++s) issue(s);

        for (int64_t kt = 0; kt < kiters; ++kt) {
            cp_async_wait<kStages - 2>();
            __syncthreads();
            issue(kt + kStages - 1);
            const double* sA = smem + (kt % kStages) * 2 * kPanelDoubles;
            const double* sB = diag ? sA : sA + kPanelDoubles;
#pragma unroll
            for (int kb = 0; kb < kKT / 4; ++kb) {
                double af[4][2], bf[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const double* p = sA + ((kb * kBM + wm * 64 + mi * 16 + g) * 4 + t);
                    af[mi][0] = p[0];
                    af[mi][1] = p[8 * 4];
                }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) bf[ni] = sB[(kb * kBM + wn * 32 + ni * 8 + g) * 4 + t];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma_m16n8k4(acc[mi][ni], af[mi][0], af[mi][1], bf[ni]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // all warps done with the ring before the next tile's prologue overwrites it

        // ---- epilogue: scale, (scatter by variable map), write both triangles ----
        const double sc = a.scale.ptr ? a.scale.c * a.scale.ptr[inst * a.scale.bstride] : a.scale.c;
        double* out = a.out + inst * a.o_bstride;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int64_t m = i0 + wm * 64 + mi * 16 + g + (r >> 1) * 8;
                    int64_t nn = j0 + wn * 32 + ni * 8 + 2 * t + (r & 1);
                    if (m < a.n && nn < a.p) {
                        double v = sc * acc[mi][ni][r];
                        if (SYRK) {
                            int64_t om = a.map ? a.map[m] : m, on = a.map ? a.map[nn] : nn;
                            double* p1 = out + om + on * a.ldo;
                            if (a.accumulate) *p1 += v; else *p1 = v;
                            if (!diag) {
                                double* p2 = out + on + om * a.ldo;
                                if (a.accumulate) *p2 += v; else *p2 = v;
                            }
                        } else {
                            out[m + nn * a.ldo] = v;
                        }
                    }
                }
            }
        }
    }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <bool SYRK>
static void launch_impl(GramKernelArgs& ka, int64_t batch, int sm_count, cudaStream_t stream) {
    ka.tiles_m = (int)((ka.n + kBM - 1) / kBM);
    ka.tiles_n = (int)((ka.p + kBM - 1) / kBM);
    ka.tiles_per_inst = SYRK ? (int64_t)ka.tiles_n * (ka.tiles_n + 1) / 2 : (int64_t)ka.tiles_m * ka.tiles_n;
    ka.total_work = ka.tiles_per_inst * batch;
    if (ka.total_work == 0 || ka.k == 0) return;
    bool al = aligned16(ka.X) && aligned16(ka.Y) && (ka.ldx % 2 == 0) && (ka.ldy % 2 == 0) &&
              (ka.x_bstride % 2 == 0) && (ka.y_bstride % 2 == 0);
    int grid = (int)(ka.total_work < sm_count ? ka.total_work : sm_count);
    if (al) gram_dmma_kernel<SYRK, true><<<grid, kGramThreads, kGramSmemBytes, stream>>>(ka);
    else gram_dmma_kernel<SYRK, false><<<grid, kGramThreads, kGramSmemBytes, stream>>>(ka);
}

// called once per context (outside any stream capture)
void configure_gram_kernels() {
    cudaFuncSetAttribute(gram_dmma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSmemBytes);
    cudaFuncSetAttribute(gram_dmma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSmemBytes);
    cudaFuncSetAttribute(gram_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSmemBytes);
    cudaFuncSetAttribute(gram_dmma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSmemBytes);
}

void launch_gram_syrk(const GramArgs& a, int64_t batch, int sm_count, cudaStream_t stream) {
    GramKernelArgs ka{};
    ka.X = a.X; ka.ldx = a.ldx; ka.x_bstride = a.x_bstride;
    ka.Y = a.X; ka.ldy = a.ldx; ka.y_bstride = a.x_bstride;
    ka.k = a.k; ka.n = a.n; ka.p = a.n;
    ka.out = a.H; ka.ldo = a.ldh; ka.o_bstride = a.h_bstride;
    ka.map = a.map; ka.scale = a.scale; ka.accumulate = a.accumulate;
    launch_impl<true>(ka, batch, sm_count, stream);
}

void launch_gemm_tn(const GemmTNArgs& a, int64_t batch, int sm_count, cudaStream_t stream) {
    GramKernelArgs ka{};
    ka.X = a.X; ka.ldx = a.ldx; ka.x_bstride = a.x_bstride;
    ka.Y = a.Y; ka.ldy = a.ldy; ka.y_bstride = a.y_bstride;
    ka.k = a.k; ka.n = a.n; ka.p = a.p;
    ka.out = a.M; ka.ldo = a.ldm; ka.o_bstride = a.m_bstride;
    ka.map = nullptr; ka.scale = AlphaDev{1.0, nullptr, 0}; ka.accumulate = 0;
    launch_impl<false>(ka, batch, sm_count, stream);
}

}  // namespace pmb
