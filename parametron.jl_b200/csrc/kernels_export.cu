// Packing kernels either side of the hot path: all HBM-bound, static patterns, values only.
//   export_terms_kernel  -- the reference's MOI.VectorAffineTerm list after update! (moi_interop.jl:64-81), 24-byte records,
//                           row-outer / term-inner: a transposing gather from the column-major Parameter buffers
//   pack_upper_kernel    -- upper-triangular CSC values of the MOI Hessian (what an OSQP-style solver ingests as P)
//   pack_constraints     -- all constraint blocks stacked into one CSC value array + row bounds from the set kinds
#include <math_constants.h>

#include <algorithm>

#include "pmb_kernels.h"

namespace pmb {

__device__ __forceinline__ double alpha_e(const AlphaDev& a, int64_t inst) { return a.ptr ? a.c * a.ptr[inst * a.bstride] : a.c; }

struct TermRec { int64_t row; double coeff; int64_t var; };

// tile: 32 rows x 32 terms.  Reads run along rows (contiguous in the column-major sources), writes along terms (contiguous records).
__global__ void __launch_bounds__(256) export_terms_kernel(ExportArgs a, int64_t batch) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t r0 = (int64_t)blockIdx.x * 32, t0 = (int64_t)blockIdx.y * 32;
    for (int64_t inst = blockIdx.z; inst < batch; inst += gridDim.z) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int tl = ty + q * 8;
            const int64_t t = t0 + tl, r = r0 + tx;
            double v = 0.0;
            if (t < a.T && r < a.rows) {
                const TermEntryDev e = a.tmpl[t];
                if (e.kind == 0) {
                    const AffMatTermDev m = a.mats[e.term];
                    v = alpha_e(m.alpha, inst) * m.src[inst * m.src_bstride + r + e.srccol * m.ld];
                } else {
                    v = alpha_e(a.units[e.term].alpha, inst);
                }
            }
            tile[tl][tx] = v;
        }
        __syncthreads();
        TermRec* out = reinterpret_cast<TermRec*>(a.out) + inst * a.rows * a.T;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int rl = ty + q * 8;
            const int64_t r = r0 + rl, t = t0 + tx;
            if (t < a.T && r < a.rows) {
                const TermEntryDev e = a.tmpl[t];
                TermRec rec;
                rec.row = r + 1;
                rec.coeff = tile[tx][rl];
                rec.var = e.kind == 0 ? e.var : a.unit_vars[e.srccol + r];
                out[r * a.T + t] = rec;
            }
        }
        __syncthreads();
    }
}
// generic term list (rows with different term layouts, e.g. a vcat of blocks of different shapes): one static gather entry per term
__global__ void __launch_bounds__(256) export_terms_gather_kernel(const TermGatherDev* tab, int64_t nterms, const AffMatTermDev* mats, const AffUnitTermDev* units,
                                                                  void* outv, int64_t batch) {
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        TermRec* out = reinterpret_cast<TermRec*>(outv) + inst * nterms;
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nterms; t += (int64_t)gridDim.x * blockDim.x) {
            const TermGatherDev e = tab[t];
            TermRec rec;
            rec.row = e.row;
            rec.var = e.var;
            if (e.kind == 0) {
                const AffMatTermDev m = mats[e.term];
                rec.coeff = alpha_e(m.alpha, inst) * m.src[inst * m.src_bstride + e.src_off];
            } else {
                rec.coeff = alpha_e(units[e.term].alpha, inst);
            }
            out[t] = rec;
        }
    }
}
void launch_export_terms_gather(const TermGatherDev* tab, int64_t nterms, const AffMatTermDev* mats, const AffUnitTermDev* units, void* out, int64_t batch,
                                cudaStream_t stream) {
    if (nterms == 0) return;
    dim3 grid((unsigned)std::min<int64_t>((nterms + 255) / 256, 4096), (unsigned)(batch < 65535 ? batch : 65535));
    export_terms_gather_kernel<<<grid, 256, 0, stream>>>(tab, nterms, mats, units, out, batch);
}

// MOI.ScalarQuadraticFunction of the objective (moi_interop.jl:45-62) in combined form: one ScalarQuadraticTerm per unordered pair of
// variables (i <= j, i outer), coefficient = the entry of the symmetric MOI Hessian (1/2 x'Qx: diagonal terms carry Q_ii, an off-diagonal
// term carries Q_ij = Q_ji once), one ScalarAffineTerm per variable, the constant.  Records: {coefficient, variable_1, variable_2} 24 bytes,
// {coefficient, variable} 16 bytes, 1-based optimizer indices.
struct QuadRec { double coeff; int64_t v1; int64_t v2; };
struct AffRec { double coeff; int64_t v; };
__global__ void __launch_bounds__(256) export_objective_kernel(const double* Q, const double* q, int64_t N, int64_t q_bstride, int64_t qq_bstride, void* quadv,
                                                               void* affv, int64_t batch) {
    const int64_t npairs = N * (N + 1) / 2;
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        if (Q && quadv) {
            QuadRec* out = reinterpret_cast<QuadRec*>(quadv) + inst * npairs;
            const double* Qi = Q + inst * q_bstride;
            for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npairs; t += (int64_t)gridDim.x * blockDim.x) {
                // row-outer enumeration of the upper triangle: row i holds N - i pairs, starting at i N - i (i - 1) / 2
                int64_t i = (int64_t)((2.0 * N + 1.0 - sqrt((2.0 * N + 1.0) * (2.0 * N + 1.0) - 8.0 * (double)t)) * 0.5);
                while (i > 0 && i * N - i * (i - 1) / 2 > t) --i;
                while ((i + 1) * N - (i + 1) * i / 2 <= t) ++i;
                const int64_t j = i + (t - (i * N - i * (i - 1) / 2));
                QuadRec r;
                r.coeff = Qi[i + j * N]; r.v1 = i + 1; r.v2 = j + 1;
                out[t] = r;
            }
        }
        if (affv) {
            AffRec* out = reinterpret_cast<AffRec*>(affv) + inst * N;
            for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (int64_t)gridDim.x * blockDim.x) {
                AffRec r;
                r.coeff = q[inst * qq_bstride + t]; r.v = t + 1;
                out[t] = r;
            }
        }
    }
}
void launch_export_objective(const double* Q, const double* q, int64_t N, void* quad_out, void* aff_out, int64_t batch, cudaStream_t stream) {
    if (N == 0) return;
    const int64_t work = std::max<int64_t>(N * (N + 1) / 2, N);
    dim3 grid((unsigned)std::min<int64_t>((work + 255) / 256, 8192), (unsigned)(batch < 65535 ? batch : 65535));
    export_objective_kernel<<<grid, 256, 0, stream>>>(Q, q, N, N * N, N, quad_out, aff_out, batch);
}

void launch_export_terms(const ExportArgs& a, int64_t batch, cudaStream_t stream) {
    if (a.rows == 0 || a.T == 0) return;
    dim3 grid((unsigned)((a.rows + 31) / 32), (unsigned)((a.T + 31) / 32), (unsigned)(batch < 65535 ? batch : 65535));
    export_terms_kernel<<<grid, 256, 0, stream>>>(a, batch);
}

// one CTA per (column, instance): column j of Q (rows 0..j, contiguous) -> P_x[j(j+1)/2 ...] (contiguous)
__global__ void __launch_bounds__(256) pack_upper_kernel(const double* Q, int64_t N, int64_t qbs, double* Px, int64_t pbs, int64_t batch) {
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y)
        for (int64_t j = blockIdx.x; j < N; j += gridDim.x) {
            const double* src = Q + inst * qbs + j * N;
            double* dst = Px + inst * pbs + j * (j + 1) / 2;
            for (int64_t i = threadIdx.x; i <= j; i += 256) dst[i] = src[i];
        }
}
void launch_pack_upper(const double* Q, int64_t N, int64_t q_bstride, double* Px, int64_t px_bstride, int64_t batch, cudaStream_t stream) {
    if (N == 0) return;
    int64_t by = batch < 65535 ? batch : 65535;
    int64_t bx = N;
    if (bx * by > 65535 * 16) bx = (65535 * 16) / by > 0 ? (65535 * 16) / by : 1;
    pack_upper_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(Q, N, q_bstride, Px, px_bstride, batch);
}

__global__ void __launch_bounds__(256) pack_constraints_kernel(const PackBlockDev* blocks, int nblocks, int64_t N, double* Ax, int64_t abs_,
                                                               double* l, double* u, int64_t m_total, int64_t batch) {
    const int64_t gtid = (int64_t)blockIdx.x * 256 + threadIdx.x, gsz = (int64_t)gridDim.x * 256;
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        double* ax = Ax + inst * abs_;
        for (int b = 0; b < nblocks; ++b) {
            const PackBlockDev B = blocks[b];
            const double* C = B.C + inst * B.c_bstride;
            if (B.sparse) {
                for (int64_t e = gtid; e < B.nnz; e += gsz) ax[B.pos[e]] = C[e];
            } else {
                const int64_t total = B.rows * N;
                for (int64_t e = gtid; e < total; e += gsz) {
                    const int64_t j = e / B.rows, r = e - j * B.rows;
                    ax[B.colbase[j] + r] = C[e];
                }
            }
            const double* d = B.d + inst * B.d_bstride;
            for (int64_t r = gtid; r < B.rows; r += gsz) {
                const double nd = -d[r];
                // block: C x + d in set   (model.jl:75-95)
                l[inst * m_total + B.row_off + r] = B.set_kind == 1 ? -CUDART_INF : nd;
                u[inst * m_total + B.row_off + r] = B.set_kind == 0 ? CUDART_INF : nd;
            }
        }
    }
}
void launch_pack_constraints(const PackBlockDev* blocks_dev, int nblocks, int64_t N, double* Ax, int64_t ax_bstride, double* l, double* u,
                             int64_t m_total, int64_t batch, int sm_count, cudaStream_t stream) {
    if (nblocks == 0) return;
    int64_t by = batch < 65535 ? batch : 65535;
    int64_t bx = (int64_t)sm_count * 8;
    if (by >= 64) bx = sm_count;
    pack_constraints_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(blocks_dev, nblocks, N, Ax, ax_bstride, l, u, m_total, batch);
}

}  // namespace pmb
