// Small kernels around the two hot kernels: Parameter generators, numeric prologue ops, reductions for the affine
// offsets / linear objective part (warp-shuffle), scatter of the rarely used quadratic shapes.
#include "pmb_kernels.h"
#include "pmb_philox.h"

namespace pmb {

__device__ __forceinline__ double alpha_val(const AlphaDev& a, int64_t inst) {
    return a.ptr ? a.c * a.ptr[inst * a.bstride] : a.c;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum for 256 threads; result valid in thread 0
__device__ __forceinline__ double block_sum_256(double v) {
    __shared__ double part[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < 8 ? part[threadIdx.x] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// ---------------- fused prologue: every built-in Parameter generator + the zeroing of (q, c0) + the update tick -------------
// One launch replaces  tick + one Philox launch per Parameter + a memset node  (launch latency dominates the small configs).
// All blocks read the update counter, use counter + 1 as this update's id, and the block that finishes LAST (atomic ticket)
// publishes the incremented counter -- every other block has read it by then.
__global__ void __launch_bounds__(256) prologue_kernel(const FillJob* jobs, int njobs, int total_blocks, int64_t batch, int64_t inst_offset,
                                                       unsigned long long* counter, unsigned int* ticket, int tick) {
    const uint32_t upd = (uint32_t)(*counter + (tick ? 1ULL : 0ULL));   // tick = 0: the update's first prologue launch has already published the id
    // which job does this block belong to?
    int j = 0;
    int b = blockIdx.x;
    while (j + 1 < njobs && b >= jobs[j].nblocks) { b -= jobs[j].nblocks; ++j; }
    const FillJob job = jobs[j];
    if (job.kind == 0) {
        const int64_t pairs = (job.numel + 1) / 2;
        const bool vec_ok = (job.numel % 2 == 0) && (((uintptr_t)job.out & 15) == 0);
        const uint32_t k0 = (uint32_t)job.seed, k1 = (uint32_t)(job.seed >> 32);
        const double sc = job.scale * (1.0 / 9007199254740992.0);   // scale 2^-53: a power of two times scale, the product below rounds once like (x 2^-53) scale
        const int64_t stride = (int64_t)job.nblocks * 512;   // a block covers 512 consecutive pairs per pass: two per thread, 256 apart
        for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
            double* o = job.out + inst * job.numel;
            const uint32_t ci = (uint32_t)(inst + inst_offset);
            // two independent Philox chains per iteration: the ten dependent rounds of one hide behind the other
            for (int64_t p = (int64_t)b * 512 + threadIdx.x; p < pairs; p += stride) {
                const int64_t p2 = p + 256;
                uint32_t c[4] = {(uint32_t)p, (uint32_t)((uint64_t)p >> 32), ci, upd};
                uint32_t e[4] = {(uint32_t)p2, (uint32_t)((uint64_t)p2 >> 32), ci, upd};
                philox4x32_10x2(c, e, k0, k1);
                const double d0 = u53_to_f64(c[1], c[0]) * sc, d1 = u53_to_f64(c[3], c[2]) * sc;
                const double d2 = u53_to_f64(e[1], e[0]) * sc, d3 = u53_to_f64(e[3], e[2]) * sc;
                if (vec_ok) {
                    *reinterpret_cast<double2*>(o + 2 * p) = make_double2(d0, d1);
                    if (p2 < pairs) *reinterpret_cast<double2*>(o + 2 * p2) = make_double2(d2, d3);
                } else {
                    o[2 * p] = d0;
                    if (2 * p + 1 < job.numel) o[2 * p + 1] = d1;
                    if (p2 < pairs) {
                        o[2 * p2] = d2;
                        if (2 * p2 + 1 < job.numel) o[2 * p2 + 1] = d3;
                    }
                }
            }
        }
    } else if (blockIdx.y == 0) {  // zero job: numel covers all instances
        for (int64_t i = (int64_t)b * 256 + threadIdx.x; i < job.numel; i += (int64_t)job.nblocks * 256) job.out[i] = 0.0;
    }
    if (!tick) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(ticket, 1u);
        if (done == (unsigned int)(total_blocks * gridDim.y) - 1u) {
            *counter = (unsigned long long)upd;
            *ticket = 0u;
        }
    }
}
void launch_prologue(const FillJob* jobs_dev, int njobs, int total_blocks, int64_t batch, int64_t inst_offset, unsigned long long* counter,
                     unsigned int* ticket, int tick, cudaStream_t stream) {
    if (njobs == 0 || total_blocks == 0) return;
    // at most 64 instance rows: the per-job block counts were sized for min(batch, 64) rows, and the end-of-kernel ticket is one
    // atomic per block (hundreds of thousands of blocks would serialise on it)
    int64_t by = batch < 64 ? batch : 64;
    prologue_kernel<<<dim3((unsigned)total_blocks, (unsigned)by), 256, 0, stream>>>(jobs_dev, njobs, total_blocks, batch, inst_offset, counter, ticket, tick);
}

// ---------------- numeric prologue ---------------------------------------------------------------
__global__ void lincomb_kernel(double* out, int64_t out_bstride, int64_t n, const LinCombTermDev* terms, int nterms, const double* cst,
                               int64_t batch) {
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            double acc = cst ? cst[i] : 0.0;
            for (int t = 0; t < nterms; ++t)
                acc += alpha_val(terms[t].alpha, inst) * terms[t].src[inst * terms[t].src_bstride + terms[t].src_off + i];
            out[inst * out_bstride + i] = acc;
        }
}
void launch_lincomb(double* out, int64_t out_bstride, int64_t n, const LinCombTermDev* terms, int nterms, const double* cst,
                    int64_t batch, cudaStream_t stream) {
    if (n == 0) return;
    int64_t bx = (n + 255) / 256; if (bx > 2048) bx = 2048;
    int64_t by = batch < 65535 ? batch : 65535;
    lincomb_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(out, out_bstride, n, terms, nterms, cst, batch);
}

__global__ void transpose_kernel(double* out, int64_t out_bstride, const double* in, int64_t in_bstride, int64_t rows, int64_t cols) {
    __shared__ double tile[32][33];
    const double* src = in + (int64_t)blockIdx.z * in_bstride;
    double* dst = out + (int64_t)blockIdx.z * out_bstride;
    int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int64_t r = r0 + threadIdx.x, c = c0 + j;
        if (r < rows && c < cols) tile[j][threadIdx.x] = src[r + c * rows];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int64_t c = c0 + threadIdx.x, r = r0 + j;  // out is cols x rows: out[c + r*cols] = in[r + c*rows]
        if (r < rows && c < cols) dst[c + r * cols] = tile[threadIdx.x][j];
    }
}
void launch_transpose(double* out, int64_t out_bstride, const double* in, int64_t in_bstride, int64_t rows, int64_t cols, int64_t batch,
                      cudaStream_t stream) {
    if (rows == 0 || cols == 0) return;
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32), (unsigned)nb);
        transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(out + b0 * out_bstride, out_bstride, in + b0 * in_bstride, in_bstride, rows, cols);
    }
}

// y = A'v : one warp per column, rows contiguous -> coalesced, warp-shuffle reduction
__device__ void dot_body(const DotArgs& a, int64_t inst);

// y = A'v (one warp per column) with an optional fused  out (+)= scale * x.y  done by the last block of the launch
__global__ void __launch_bounds__(256) gemv_t_kernel(GemvArgs a, int64_t batch, DotArgs dt, int has_dot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (has_dot && blockIdx.x == gridDim.x - 1) {
        for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) dot_body(dt, inst);
        return;
    }
    const int gx = gridDim.x - (has_dot ? 1 : 0);
    const bool vec2 = (a.rows % 2 == 0) && (a.lda % 2 == 0) && (a.a_bstride % 2 == 0) && (a.v_bstride % 2 == 0) && (a.v_off % 2 == 0) &&
                      (((uintptr_t)a.A | (uintptr_t)a.v) & 15) == 0;
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        const double* A = a.A + inst * a.a_bstride;
        const double* v = a.v + inst * a.v_bstride + a.v_off;
        for (int64_t col = (int64_t)blockIdx.x * 8 + warp; col < a.cols; col += (int64_t)gx * 8) {
            const double* ac = A + col * a.lda;
            double acc = 0.0;
            if (vec2) {
                // 16-byte loads, four independent loads in flight per lane
                const double2* ac2 = reinterpret_cast<const double2*>(ac);
                const double2* v2 = reinterpret_cast<const double2*>(v);
                const int64_t n2 = a.rows >> 1;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int64_t r = lane;
                for (; r + 96 < n2; r += 128) {
                    const double2 x0 = ac2[r], x1 = ac2[r + 32], x2 = ac2[r + 64], x3 = ac2[r + 96];
                    const double2 y0 = v2[r], y1 = v2[r + 32], y2 = v2[r + 64], y3 = v2[r + 96];
                    s0 += x0.x * y0.x + x0.y * y0.y;
                    s1 += x1.x * y1.x + x1.y * y1.y;
                    s2 += x2.x * y2.x + x2.y * y2.y;
                    s3 += x3.x * y3.x + x3.y * y3.y;
                }
                for (; r < n2; r += 32) {
                    const double2 x0 = ac2[r], y0 = v2[r];
                    s0 += x0.x * y0.x + x0.y * y0.y;
                }
                acc = (s0 + s1) + (s2 + s3);
            } else {
                for (int64_t r = lane; r < a.rows; r += 32) acc += ac[r] * v[r];
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                double val = alpha_val(a.scale, inst) * acc;
                double* dst = a.y + inst * a.y_bstride + (a.map ? a.map[col] : col);
                if (a.accumulate) *dst += val; else *dst = val;
            }
        }
    }
}
// y = A v : 32 rows per CTA, 8 column slices reduced through shared memory
__global__ void __launch_bounds__(256) gemv_n_kernel(GemvArgs a, int64_t batch) {
    __shared__ double red[8][33];
    const int rl = threadIdx.x & 31, cs = threadIdx.x >> 5;
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        const double* A = a.A + inst * a.a_bstride;
        const double* v = a.v + inst * a.v_bstride + a.v_off;
        for (int64_t r0 = (int64_t)blockIdx.x * 32; r0 < a.rows; r0 += (int64_t)gridDim.x * 32) {
            int64_t r = r0 + rl;
            double acc = 0.0;
            if (r < a.rows)
                for (int64_t c = cs; c < a.cols; c += 8) acc += A[r + c * a.lda] * v[c];
            red[cs][rl] = acc;
            __syncthreads();
            if (cs == 0 && r < a.rows) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) s += red[i][rl];
                double val = alpha_val(a.scale, inst) * s;
                double* dst = a.y + inst * a.y_bstride + (a.map ? a.map[r] : r);
                if (a.accumulate) *dst += val; else *dst = val;
            }
            __syncthreads();
        }
    }
}
void launch_gemv_t_dot(const GemvArgs& a, const DotArgs& d, int64_t batch, cudaStream_t stream) {
    int64_t by = batch < 65535 ? batch : 65535;
    int64_t bx = (a.cols + 7) / 8; if (bx > 4096) bx = 4096;
    if (bx < 1) bx = 1;
    gemv_t_kernel<<<dim3((unsigned)(bx + 1), (unsigned)by), 256, 0, stream>>>(a, batch, d, 1);
}
void launch_gemv(const GemvArgs& a, int64_t batch, cudaStream_t stream) {
    int64_t outlen = a.trans ? a.cols : a.rows;
    if (outlen == 0) return;
    int64_t by = batch < 65535 ? batch : 65535;
    if (a.trans) {
        int64_t bx = (a.cols + 7) / 8; if (bx > 4096) bx = 4096;
        gemv_t_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(a, batch, DotArgs{}, 0);
    } else {
        int64_t bx = (a.rows + 31) / 32; if (bx > 4096) bx = 4096;
        gemv_n_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(a, batch);
    }
}

__device__ void dot_body(const DotArgs& a, int64_t inst) {
    const double* x = a.x + inst * a.x_bstride + a.x_off;
    const double* y = a.y + inst * a.y_bstride + a.y_off;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < a.n; i += 256) acc += x[i] * y[i];
    acc = block_sum_256(acc);
    if (threadIdx.x == 0) {
        double val = alpha_val(a.scale, inst) * acc;
        double* dst = a.out + inst * a.out_bstride;
        if (a.accumulate) *dst += val; else *dst = val;
    }
}
__global__ void __launch_bounds__(256) dot_kernel(DotArgs a, int64_t batch) {
    for (int64_t inst = blockIdx.x; inst < batch; inst += gridDim.x) dot_body(a, inst);
}
void launch_dot(const DotArgs& a, int64_t batch, cudaStream_t stream) {
    int64_t bx = batch < 65535 ? batch : 65535;
    dot_kernel<<<(unsigned)bx, 256, 0, stream>>>(a, batch);
}

__global__ void scalar_prod_kernel(double* out, int64_t out_bstride, double c, const double* s0, int64_t s0_bstride, const double* s1,
                                   int64_t s1_bstride, int64_t batch) {
    for (int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; inst < batch; inst += (int64_t)gridDim.x * blockDim.x) {
        double v = c;
        if (s0) v *= s0[inst * s0_bstride];
        if (s1) v *= s1[inst * s1_bstride];
        out[inst * out_bstride] = v;
    }
}
void launch_scalar_prod(double* out, int64_t out_bstride, double c, const double* s0, int64_t s0_bstride, const double* s1,
                        int64_t s1_bstride, int64_t batch, cudaStream_t stream) {
    int64_t bx = (batch + 255) / 256; if (bx > 1024) bx = 1024;
    scalar_prod_kernel<<<(unsigned)bx, 256, 0, stream>>>(out, out_bstride, c, s0, s0_bstride, s1, s1_bstride, batch);
}

__global__ void scatter_vec_kernel(ScatterVecArgs a, int64_t batch) {
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        const double al = alpha_val(a.alpha, inst);
        for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += (int64_t)gridDim.x * blockDim.x) {
            double val = al * a.v[inst * a.v_bstride + a.v_off + j];
            double* dst = a.q + inst * a.q_bstride + (a.map ? a.map[j] : j);
            if (a.accumulate) atomicAdd(dst, val); else *dst = val;
        }
    }
}
void launch_scatter_vec(const ScatterVecArgs& a, int64_t batch, cudaStream_t stream) {
    if (a.n == 0) return;
    int64_t bx = (a.n + 255) / 256; if (bx > 2048) bx = 2048;
    int64_t by = batch < 65535 ? batch : 65535;
    scatter_vec_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(a, batch);
}

__global__ void bilinear_scatter_kernel(BilinearArgs a, int64_t batch) {
    const int64_t total = a.nx * a.ny;
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        const double al = alpha_val(a.alpha, inst);
        const double* Q = a.Q + inst * a.q_bstride;
        double* H = a.H + inst * a.h_bstride;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            // e is Q's column-major linear index == the reference's running k: (row, col) = (e / ny, e % ny)   functions.jl:849-856
            int64_t qi = e % a.nx, qj = e / a.nx;
            int64_t row = e / a.ny, col = e % a.ny;
            double v = al * Q[qi + qj * a.ldq];
            int64_t hi = a.xmap[row], hj = a.ymap[col];
            atomicAdd(H + hi + hj * a.ldh, v);
            atomicAdd(H + hj + hi * a.ldh, v);
        }
    }
}
void launch_bilinear_scatter(const BilinearArgs& a, int64_t batch, cudaStream_t stream) {
    int64_t total = a.nx * a.ny;
    if (total == 0) return;
    int64_t bx = (total + 255) / 256; if (bx > 4096) bx = 4096;
    int64_t by = batch < 65535 ? batch : 65535;
    bilinear_scatter_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(a, batch);
}

__global__ void coo_quad_kernel(CooQuadArgs a, int64_t batch) {
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        const double al = alpha_val(a.alpha, inst);
        double* H = a.H + inst * a.h_bstride;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.nnz; e += (int64_t)gridDim.x * blockDim.x) {
            double v = al * a.c[e];
            atomicAdd(H + a.i[e] + (int64_t)a.j[e] * a.ldh, v);
            atomicAdd(H + a.j[e] + (int64_t)a.i[e] * a.ldh, v);
        }
    }
}
void launch_coo_quad(const CooQuadArgs& a, int64_t batch, cudaStream_t stream) {
    if (a.nnz == 0) return;
    int64_t bx = (a.nnz + 255) / 256; if (bx > 2048) bx = 2048;
    int64_t by = batch < 65535 ? batch : 65535;
    coo_quad_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(a, batch);
}

__global__ void scatter_sym_kernel(ScatterSymArgs a, int64_t batch) {
    const int64_t total = a.n * a.p;
    for (int64_t inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        const double al = alpha_val(a.alpha, inst);
        const double* M = a.M + inst * a.m_bstride;
        double* H = a.H + inst * a.h_bstride;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
            int64_t i = e % a.n, j = e / a.n;
            double v = al * M[i + j * a.ldm];
            int64_t hi = a.xmap ? a.xmap[i] : i, hj = a.ymap ? a.ymap[j] : j;
            atomicAdd(H + hi + hj * a.ldh, v);
            atomicAdd(H + hj + hi * a.ldh, v);
        }
    }
}
void launch_scatter_sym(const ScatterSymArgs& a, int64_t batch, cudaStream_t stream) {
    int64_t total = a.n * a.p;
    if (total == 0) return;
    int64_t bx = (total + 255) / 256; if (bx > 4096) bx = 4096;
    int64_t by = batch < 65535 ? batch : 65535;
    scatter_sym_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, stream>>>(a, batch);
}

__global__ void fill_kernel(double* p, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
void launch_fill(double* p, int64_t n, double v, cudaStream_t stream) {
    if (n == 0) return;
    int64_t bx = (n + 255) / 256; if (bx > 4096) bx = 4096;
    fill_kernel<<<(unsigned)bx, 256, 0, stream>>>(p, n, v);
}

__global__ void gather_kernel(double* out, const double* x, const int64_t* idx, int64_t nv, int64_t N, int64_t batch) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nv * batch; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / nv, i = t - b * nv;
        out[t] = x[b * N + idx[i]];
    }
}
void launch_gather(double* out, const double* x, const int64_t* idx, int64_t nv, int64_t N, int64_t batch, cudaStream_t stream) {
    if (nv * batch == 0) return;
    const int grid = (int)std::min<int64_t>((nv * batch + 255) / 256, 4096);
    gather_kernel<<<grid, 256, 0, stream>>>(out, x, idx, nv, N, batch);
}

}  // namespace pmb
