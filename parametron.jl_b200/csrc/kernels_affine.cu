// K1 -- batched affine-coefficient assembly (HBM-bound).
//
// Device restatement of the reference chain
//   matvecmul!(y, C, x)            functions.jl:775-798   (one LinearTerm per C[i])
//   vecsubtract!/vecadd!           functions.jl:751-764   (copy the term list, fold numbers into the constant)
//   scale!/mul!                    functions.jl:873-925, 578
//   vcat!                          functions.jl:969-994
//   update!(::VectorAffineFunction) moi_interop.jl:64-81  (flatten + varmap)
// on the dense representation: the block  (C_i, d_i)  with  C_i[row, optimizer_var]  column-major.
// The sparsity structure is static (built once by the planner), so one launch per block only moves
// coefficient VALUES: 8 B read + 8 B written per stored coefficient, 16 B per row for the offsets.
#include "pmb_kernels.h"

namespace pmb {

__device__ __forceinline__ double alpha_of(const AlphaDev& a, int64_t inst) {
    return a.ptr ? a.c * a.ptr[inst * a.bstride] : a.c;
}

__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream2(double* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

constexpr int kAffThreads = 256;
constexpr int kAffUnroll = 4;  // independent 16-byte loads in flight per thread

// offsets: d[i] = d_const[i] + sum_t alpha_t * v_t[...]; handled by the CTAs after the coefficient CTAs
__device__ void affine_offsets(const AffBlockDev& blk, int64_t inst, int tid, int nthreads) {
    if (!blk.d) return;
    for (int64_t i = tid; i < blk.d_rows; i += nthreads) {
        double acc = blk.d_const ? blk.d_const[i] : 0.0;
        for (int t = 0; t < blk.noffs; ++t) {
            const AffOffTermDev& o = blk.offs[t];
            int64_t r = i - o.row_off;
            if (r >= 0 && r < o.rows) acc += alpha_of(o.alpha, inst) * o.src[inst * o.src_bstride + o.src_off + r];
        }
        blk.d[inst * blk.d_bstride + i] = acc;
    }
}

// Fast path: every touched column is alpha * (one column of one source matrix).  grid.y strides instances.
template <int VEC>
__global__ void __launch_bounds__(kAffThreads) affine_assemble_simple(AffBlockDev blk, int64_t batch, int main_blocks) {
    const int64_t rowsv = blk.rows / VEC;
    const int64_t items = (int64_t)blk.ntouched * rowsv;
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        if ((int)blockIdx.x >= main_blocks) {
            affine_offsets(blk, inst, (blockIdx.x - main_blocks) * kAffThreads + threadIdx.x,
                           (gridDim.x - main_blocks) * kAffThreads);
            continue;
        }
        const AffMatTermDev m = blk.mats[0];
        const double a = alpha_of(m.alpha, inst);
        const double* src = m.src + inst * m.src_bstride;
        double* out = blk.C + inst * blk.C_bstride;
        for (int64_t base = (int64_t)blockIdx.x * (kAffThreads * kAffUnroll); base < items;
             base += (int64_t)main_blocks * (kAffThreads * kAffUnroll)) {
            double2 v[kAffUnroll];
            int64_t dst[kAffUnroll];
#pragma unroll
            for (int u = 0; u < kAffUnroll; ++u) {
                int64_t it = base + u * kAffThreads + threadIdx.x;
                dst[u] = -1;
                if (it < items) {
                    int tc = (int)(it / rowsv);
                    int64_t r = (it - (int64_t)tc * rowsv) * VEC;
                    const double* p = src + r + (int64_t)blk.col_src[tc] * m.ld;
                    dst[u] = r + (int64_t)blk.col_out[tc] * blk.rows;
                    if (VEC == 2) v[u] = ld_stream2(p);
                    else v[u].x = __ldg(p);
                }
            }
#pragma unroll
            for (int u = 0; u < kAffUnroll; ++u) {
                if (dst[u] >= 0) {
                    if (VEC == 2) st_stream2(out + dst[u], make_double2(a * v[u].x, a * v[u].y));
                    else out[dst[u]] = a * v[u].x;
                }
            }
        }
    }
}

// General path: sums every contribution of a column (several matrices, unit terms, partial row ranges).
__global__ void __launch_bounds__(kAffThreads) affine_assemble_general(AffBlockDev blk, int64_t batch, int main_blocks) {
    const int64_t items = (int64_t)blk.ntouched * blk.rows;
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        if ((int)blockIdx.x >= main_blocks) {
            affine_offsets(blk, inst, (blockIdx.x - main_blocks) * kAffThreads + threadIdx.x,
                           (gridDim.x - main_blocks) * kAffThreads);
            continue;
        }
        double* out = blk.C + inst * blk.C_bstride;
        for (int64_t it = (int64_t)blockIdx.x * kAffThreads + threadIdx.x; it < items; it += (int64_t)main_blocks * kAffThreads) {
            int tc = (int)(it / blk.rows);
            int64_t row = it - (int64_t)tc * blk.rows;
            double acc = 0.0;
            for (int e = blk.col_ptr[tc]; e < blk.col_ptr[tc + 1]; ++e) {
                AffColEntry ce = blk.entries[e];
                if (ce.term >= 0) {
                    const AffMatTermDev& m = blk.mats[ce.term];
                    int64_t r = row - m.row_off;
                    if (r >= 0 && r < m.rows) acc += alpha_of(m.alpha, inst) * m.src[inst * m.src_bstride + r + (int64_t)ce.src * m.ld];
                } else if (ce.src == row) {
                    acc += alpha_of(blk.units[~ce.term].alpha, inst);
                }
            }
            out[row + (int64_t)blk.col_out[tc] * blk.rows] = acc;
        }
    }
}

void launch_affine_assemble(const AffBlockDev& blk, int64_t batch, int sm_count, cudaStream_t stream) {
    if (blk.rows == 0) return;
    if (blk.tma) { launch_affine_assemble_tma(blk, batch, sm_count, stream); return; }
    const int off_blocks = 1;
    int64_t gy = batch < 65535 ? batch : 65535;
    if (blk.simple) {
        int vec = blk.vec2 ? 2 : 1;
        int64_t items = (int64_t)blk.ntouched * (blk.rows / vec);
        int64_t per_block = (int64_t)kAffThreads * kAffUnroll;
        int64_t want = (items + per_block - 1) / per_block;
        // several CTAs per SM worth of 16-byte loads in flight; the grid-stride loop covers the rest
        int64_t cap = (int64_t)sm_count * 8;
        int64_t per_inst_cap = batch >= cap ? 1 : (cap + batch - 1) / batch;
        int main_blocks = (int)(want < per_inst_cap ? want : per_inst_cap);
        if (main_blocks < 1) main_blocks = 1;
        dim3 grid(main_blocks + off_blocks, (unsigned)gy);
        if (vec == 2) affine_assemble_simple<2><<<grid, kAffThreads, 0, stream>>>(blk, batch, main_blocks);
        else affine_assemble_simple<1><<<grid, kAffThreads, 0, stream>>>(blk, batch, main_blocks);
    } else {
        int64_t items = (int64_t)blk.ntouched * blk.rows;
        int64_t want = (items + kAffThreads - 1) / kAffThreads;
        int64_t cap = (int64_t)sm_count * 8;
        int64_t per_inst_cap = batch >= cap ? 1 : (cap + batch - 1) / batch;
        int main_blocks = (int)(want < per_inst_cap ? want : per_inst_cap);
        if (main_blocks < 1) main_blocks = 1;
        dim3 grid(main_blocks + off_blocks, (unsigned)gy);
        affine_assemble_general<<<grid, kAffThreads, 0, stream>>>(blk, batch, main_blocks);
    }
}

}  // namespace pmb
