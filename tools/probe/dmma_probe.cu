// Probe: fp64 mma.sync (DMMA) fragment layouts + issue throughput on sm_100a.
// Not product code: informs the tile shape of the Gram kernel.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---- layout check: C(16x8) = A(16xK) * B(Kx8), A row-major [16][K], B stored [K][8]
template <int SHAPE>
__global__ void layout_kernel(const double* A, const double* B, double* C, int K) {
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    if (SHAPE == 0) {  // m8n8k4 : M=8
        double c[2] = {0, 0};
        mma884(c, A[g * K + t], B[t * 8 + g]);
        C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1];
    } else if (SHAPE == 1) {
        double c[4] = {0, 0, 0, 0}; double a[2] = {A[g * K + t], A[(g + 8) * K + t]};
        mma1684(c, a, B[t * 8 + g]);
        C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1]; C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
    } else if (SHAPE == 2) {
        double c[4] = {0, 0, 0, 0};
        double a[4] = {A[g * K + t], A[(g + 8) * K + t], A[g * K + t + 4], A[(g + 8) * K + t + 4]};
        double b[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
        mma1688(c, a, b);
        C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1]; C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
    } else {
        double c[4] = {0, 0, 0, 0}; double a[8]; double b[4];
        for (int i = 0; i < 8; ++i) a[i] = A[(g + 8 * (i & 1)) * K + t + 4 * (i >> 1)];
        for (int i = 0; i < 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];
        mma16816(c, a, b);
        C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1]; C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
}

template <int SHAPE, int NACC>
__global__ void __launch_bounds__(256) tput_kernel(double* out, int iters, double seed) {
    double c[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i + 1);
    for (int i = 0; i < 4; ++i) b[i] = seed * (threadIdx.x * 3 + i + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (SHAPE == 0) { double cc[2] = {c[i][0], c[i][1]}; mma884(cc, a[0], b[0]); c[i][0] = cc[0]; c[i][1] = cc[1]; }
            else if (SHAPE == 1) { double aa[2] = {a[0], a[1]}; mma1684(c[i], aa, b[0]); }
            else if (SHAPE == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; mma1688(c[i], aa, bb); }
            else if (SHAPE == 3) { mma16816(c[i], a, b); }
            else { // DFMA: 4 fma per acc
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = fma(a[j], b[j], c[i][j]);
            }
        }
    }
    double s = 0; for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}

template <int SHAPE, int NACC>
void run_tput(const char* name, double fma_per_warp_op, int blocks_per_sm, int threads) {
    int dev; CK(cudaGetDevice(&dev)); cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    double* out; CK(cudaMalloc(&out, 8));
    int iters = 4096; int grid = p.multiProcessorCount * blocks_per_sm;
    tput_kernel<SHAPE, NACC><<<grid, threads>>>(out, 64, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        tput_kernel<SHAPE, NACC><<<grid, threads>>>(out, iters, 1e-9);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warps = (double)grid * threads / 32.0;
    double fmas = (SHAPE == 4) ? (double)grid * threads * iters * NACC * 4.0 : warps * iters * NACC * fma_per_warp_op;
    printf("%-10s nacc=%d blocks/sm=%d thr=%d : %.3f ms  %.2f TFLOP/s (2*FMA)\n", name, NACC, blocks_per_sm, threads, best, 2.0 * fmas / (best * 1e-3) / 1e12);
    cudaFree(out);
}

template <int SHAPE> void check_layout(const char* name, int M, int K) {
    double hA[16 * 16], hB[16 * 8], hC[16 * 8], ref[16 * 8];
    for (int i = 0; i < 16 * 16; ++i) hA[i] = (double)((i * 7 + 3) % 23) - 11.0;
    for (int i = 0; i < 16 * 8; ++i) hB[i] = (double)((i * 5 + 1) % 19) - 9.0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < 8; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += hA[m * K + k] * hB[k * 8 + n]; ref[m * 8 + n] = s; }
    double *dA, *dB, *dC; CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dC, sizeof(hC)));
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, sizeof(hC)));
    layout_kernel<SHAPE><<<1, 32>>>(dA, dB, dC, K); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost));
    double err = 0; for (int i = 0; i < M * 8; ++i) err = fmax(err, fabs(hC[i] - ref[i]));
    printf("layout %-10s max|err| = %g  %s\n", name, err, err == 0 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz smem/SM=%zu regs/SM=%d L2=%d\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate, p.sharedMemPerMultiprocessor, p.regsPerMultiprocessor, p.l2CacheSize);
    check_layout<0>("m8n8k4", 8, 4); check_layout<1>("m16n8k4", 16, 4); check_layout<2>("m16n8k8", 16, 8); check_layout<3>("m16n8k16", 16, 16);
    for (int bps = 1; bps <= 2; ++bps) {
        run_tput<0, 8>("m8n8k4", 8 * 8 * 4, bps, 256);
        run_tput<1, 8>("m16n8k4", 16 * 8 * 4, bps, 256);
        run_tput<2, 8>("m16n8k8", 16 * 8 * 8, bps, 256);
        run_tput<3, 8>("m16n8k16", 16 * 8 * 16, bps, 256);
        run_tput<4, 8>("dfma", 0, bps, 256);
    }
    run_tput<1, 16>("m16n8k4", 16 * 8 * 4, 1, 256);
    run_tput<3, 16>("m16n8k16", 16 * 8 * 16, 1, 256);
    run_tput<1, 16>("m16n8k4", 16 * 8 * 4, 1, 128);
    run_tput<3, 4>("m16n8k16", 16 * 8 * 16, 1, 128);
    run_tput<3, 2>("m16n8k16", 16 * 8 * 16, 1, 128);
    run_tput<4, 16>("dfma", 0, 2, 512);
    return 0;
}
