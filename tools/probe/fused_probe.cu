// Standalone harness of the fused small-instance int8 Gram kernel (kernels_gram_i8_fused.cu): launches it directly (no planner, no graph),
// prints the launch resources / errors, checks H, q, c0 against a plain fp64 loop on the host and times the launch with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/fused_probe tools/probe/fused_probe.cu parametron.jl_b200/csrc/kernels_gram_i8_fused.cu
//   tools/probe/fused_probe [batch n k reps]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../parametron.jl_b200/csrc/pmb_kernels.h"

namespace pmb { struct GramI8FusedArgs; __global__ void gram_i8_fused_kernel(const GramI8FusedArgs a); }
#define CKP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int64_t batch = argc > 1 ? atoll(argv[1]) : 4, n = argc > 2 ? atoll(argv[2]) : 256, k = argc > 3 ? atoll(argv[3]) : 256;
    const int reps = argc > 4 ? atoi(argv[4]) : 3;
    cudaDeviceProp prop; CKP(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, smem optin %zu, regs/block %d\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin, prop.regsPerBlock);
    pmb::configure_gram_i8_fused_kernel();
    cudaFuncAttributes fa; CKP(cudaFuncGetAttributes(&fa, pmb::gram_i8_fused_kernel));
    printf("kernel: regs %d, static smem %zu, max dyn smem %d, maxThreadsPerBlock %d, local %zu\n", fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes,
           fa.maxThreadsPerBlock, fa.localSizeBytes);
    std::vector<double> X((size_t)(batch * k * n)), g((size_t)(batch * k));
    srand(1);
    for (auto& v : X) v = rand() / (double)RAND_MAX;
    for (auto& v : g) v = -(rand() / (double)RAND_MAX);
    double *dX, *dg, *dH, *dq, *dc0; unsigned long long* dstats;
    CKP(cudaMalloc(&dX, X.size() * 8)); CKP(cudaMalloc(&dg, g.size() * 8)); CKP(cudaMalloc(&dH, (size_t)(batch * n * n) * 8));
    CKP(cudaMalloc(&dq, (size_t)(batch * n) * 8)); CKP(cudaMalloc(&dc0, (size_t)batch * 8)); CKP(cudaMalloc(&dstats, (size_t)(batch * 16 + 2 + 1024) * 8)); CKP(cudaMemset(dstats, 0, (size_t)(batch * 16 + 2 + 1024) * 8));
    CKP(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice)); CKP(cudaMemcpy(dg, g.data(), g.size() * 8, cudaMemcpyHostToDevice));
    CKP(cudaMemset(dH, 0xff, (size_t)(batch * n * n) * 8));
    pmb::GramArgs ga{};
    ga.X = dX; ga.ldx = k; ga.x_bstride = k * n; ga.k = k; ga.n = n; ga.H = dH; ga.ldh = n; ga.h_bstride = n * n; ga.map = nullptr;
    ga.scale = pmb::AlphaDev{2.0, nullptr, 0}; ga.accumulate = 0;
    ga.gvec = dg; ga.g_bstride = k; ga.qout = dq; ga.q_bstride = n; ga.c0out = dc0; ga.qscale = pmb::AlphaDev{2.0, nullptr, 0}; ga.cscale = pmb::AlphaDev{1.0, nullptr, 0};
    pmb::GramI8Scratch sc{}; sc.stats = dstats; sc.tol = 9e-11; sc.force_ns = 0; sc.fused = 1;
    unsigned long long* diag = nullptr; CKP(cudaHostAlloc(&diag, 256, cudaHostAllocMapped)); for (int i = 0; i < 32; ++i) diag[i] = 0; sc.diag = diag;
    cudaStream_t st; CKP(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < reps; ++r) {
        CKP(cudaMemsetAsync(dq, 0, (size_t)(batch * n) * 8, st)); CKP(cudaMemsetAsync(dc0, 0, (size_t)batch * 8, st));
        cudaEventRecord(e0, st);
        const bool ok = pmb::launch_gram_i8_fused(ga, sc, batch, prop.multiProcessorCount, st);
        cudaEventRecord(e1, st);
        printf("launch %d: ok=%d, last error: %s\n", r, (int)ok, cudaGetErrorString(cudaGetLastError()));
        cudaError_t se = cudaStreamSynchronize(st);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        printf("  sync: %s, %.3f ms (%.1f us / instance-slot, %.0f GB/s of 16 B per element)\n", cudaGetErrorString(se), ms,
               1e3 * ms / std::ceil(batch / 74.0), 16.0 * batch * n * k / (ms * 1e6));
        if (se != cudaSuccess) { char buf[1024]; printf("%s\n", pmb::gram_i8_fused_diag_text(diag, buf, sizeof buf)); return 1; }
    }
    std::vector<double> H((size_t)(batch * n * n)), q((size_t)(batch * n)), c0((size_t)batch);
    std::vector<unsigned long long> stats((size_t)(batch * 16 + 2 + 1024));
    CKP(cudaMemcpy(H.data(), dH, H.size() * 8, cudaMemcpyDeviceToHost)); CKP(cudaMemcpy(q.data(), dq, q.size() * 8, cudaMemcpyDeviceToHost));
    CKP(cudaMemcpy(c0.data(), dc0, c0.size() * 8, cudaMemcpyDeviceToHost)); CKP(cudaMemcpy(stats.data(), dstats, stats.size() * 8, cudaMemcpyDeviceToHost));
    printf("uncertified: %llu; verdict[0] = %llu\n", stats[(size_t)batch * 16], stats[15]);
    if (getenv("PMB_FUSED_DBG") && (atoi(getenv("PMB_FUSED_DBG")) & 32)) {
        const unsigned long long* tl = stats.data() + batch * 16 + 2;
        unsigned long long t0 = ~0ull;
        for (int r = 0; r < 4; ++r) for (int e = 0; e < 127; ++e) if (tl[r * 256 + 2 * e + 1] && tl[r * 256 + 2 * e + 1] < t0) t0 = tl[r * 256 + 2 * e + 1];
        const char* names[4] = {"worker0", "worker8", "mma", "producer"};
        for (int r = 0; r < 4; ++r) {
            printf("timeline %s:", names[r]);
            for (int e = 0; e < 127; ++e) { if (!tl[r * 256 + 2 * e + 1]) break; printf(" [%llu]%llu", tl[r * 256 + 2 * e], tl[r * 256 + 2 * e + 1] - t0); }
            printf("\n");
        }
    }
    double worst = 0, worstq = 0, worstc = 0; int64_t bad_inst = -1;
    const int64_t check[] = {0, 1, batch / 2, batch - 1};
    for (int64_t inst : check) {
        if (inst < 0 || inst >= batch) continue;
        const double* x = X.data() + inst * k * n; const double* gg = g.data() + inst * k;
        double hmax = 0;
        for (int64_t j = 0; j < n; ++j) { double s = 0; for (int64_t r = 0; r < k; ++r) s += x[r + j * k] * x[r + j * k]; hmax = std::fmax(hmax, 2 * s); }
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j) {
                long double s = 0; for (int64_t r = 0; r < k; ++r) s += (long double)x[r + i * k] * x[r + j * k];
                const double err = std::fabs((double)(2 * s) - H[(size_t)(inst * n * n + i + j * n)]) / hmax;
                if (!(err <= worst)) { worst = err; bad_inst = inst; }
            }
        double cc = 0; for (int64_t r = 0; r < k; ++r) cc += gg[r] * gg[r];
        worstc = std::fmax(worstc, std::fabs(cc - c0[(size_t)inst]) / cc);
        for (int64_t j = 0; j < n; ++j) { double s = 0; for (int64_t r = 0; r < k; ++r) s += x[r + j * k] * gg[r]; worstq = std::fmax(worstq, std::fabs(2 * s - q[(size_t)(inst * n + j)]) / std::fabs(2 * s)); }
    }
    printf("max |H - ref| / max|H| = %.3e (instance %lld), q rel %.3e, c0 rel %.3e  -> %s\n", worst, (long long)bad_inst, worstq, worstc,
           (worst <= 1e-10 && worstq <= 1e-12 && worstc <= 1e-12) ? "PASS" : "FAIL");
    return 0;
}
