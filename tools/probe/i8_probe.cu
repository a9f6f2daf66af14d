// Standalone check + timing of the certified int8-slice Gram (kernels_gram_i8.cu) against a host fp64/long-double Gram.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I parametron.jl_b200/csrc tools/probe/i8_probe.cu \
//             parametron.jl_b200/csrc/kernels_gram_i8.cu -o tools/probe/i8_probe -lcuda
// run:   tools/probe/i8_probe [n k batch ns dist reps]     dist: 0 uniform[0,1), 1 normal-ish mixed sign, 2 wide dynamic range
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "pmb_kernels.h"

using namespace pmb;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

int main(int argc, char** argv) {
    int64_t n = argc > 1 ? atoll(argv[1]) : 200, k = argc > 2 ? atoll(argv[2]) : 300, batch = argc > 3 ? atoll(argv[3]) : 1;
    int ns = argc > 4 ? atoi(argv[4]) : 5, dist = argc > 5 ? atoi(argv[5]) : 0, reps = argc > 6 ? atoi(argv[6]) : 3;
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sm = prop.multiProcessorCount;
    printf("device %s sm=%d  n=%lld k=%lld batch=%lld ns=%d dist=%d\n", prop.name, sm, (long long)n, (long long)k, (long long)batch, ns, dist);
    configure_gram_i8_kernels();
    const int64_t ld = (k + 1) / 2 * 2;
    std::vector<double> X((size_t)(batch * ld * n));
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> G(0.0, 1.0);
    for (auto& v : X) v = dist == 0 ? U(rng) : dist == 1 ? G(rng) : G(rng) * std::pow(10.0, 6.0 * U(rng) - 3.0);
    double *dX, *dH;
    CK(cudaMalloc(&dX, X.size() * 8));
    CK(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dH, (size_t)(batch * n * n) * 8));
    CK(cudaMemset(dH, 0xFF, (size_t)(batch * n * n) * 8));
    GramI8Scratch sc{};
    sc.plan = plan_gram_i8(k, n, batch);
    if (!sc.plan.ok) { printf("plan: not applicable\n"); return 1; }
    sc.tol = 9e-11;
    sc.force_ns = ns;   // 0: adaptive
    CK(cudaMalloc(&sc.slices, (size_t)sc.plan.slice_bytes));
    CK(cudaMalloc(&sc.sigma, (size_t)(batch * n) * 8));
    CK(cudaMalloc(&sc.stats, (size_t)batch * 8 * kGramI8StatsStride));
    GramArgs g{};
    g.X = dX; g.ldx = ld; g.x_bstride = ld * n; g.k = k; g.n = n; g.H = dH; g.ldh = n; g.h_bstride = n * n;
    g.scale = AlphaDev{2.0, nullptr, 0};
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&e2));
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0, st));
        if (!launch_gram_i8_slice(g, sc, batch, sm, st)) { printf("launch refused\n"); return 1; }
        CK(cudaEventRecord(e1, st));
        if (!launch_gram_i8_mma(g, sc, batch, sm, st)) { printf("launch refused\n"); return 1; }
        CK(cudaEventRecord(e2, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        float ms1, ms2;
        CK(cudaEventElapsedTime(&ms1, e0, e1));
        CK(cudaEventElapsedTime(&ms2, e1, e2));
        const double ops = (double)batch * k * n * (n + 1);
        printf("run %d: slice %.3f ms + gram %.3f ms = %.1f TFLOP/s SYRK-equivalent fp64\n", r, ms1, ms2, ops / (ms1 + ms2) * 1e-9);
    }
    std::vector<unsigned long long> stats((size_t)batch * kGramI8StatsStride);
    CK(cudaMemcpy(stats.data(), sc.stats, stats.size() * 8, cudaMemcpyDeviceToHost));
    {
        double sig, ss, l1;
        memcpy(&sig, &stats[0], 8); memcpy(&ss, &stats[1], 8); memcpy(&l1, &stats[2], 8);
        printf("inst 0: sigma_max=%.4e ssmax=%.6e l1max=%.4e  R=", sig, ss, l1);
        double R[kGramI8Planes];
        for (int s2 = 0; s2 < kGramI8Planes; ++s2) { R[s2] = std::sqrt((double)stats[3 + s2]); printf("%.3e ", R[s2]); }
        printf("\n");
        for (int nsx = 5; nsx <= 6; ++nsx) {
            double by = 2.0 * 1.1 * std::exp2(-8.0 * nsx) * l1;
            for (int s2 = 1; s2 < nsx; ++s2)
                for (int t2 = 1; t2 < nsx; ++t2)
                    if (s2 + t2 >= nsx) by += std::exp2(-14.0 - 8.0 * (s2 + t2)) * R[s2] * R[t2];
            printf("  ns=%d: bound/ssmax = %.3e (tol %.1e)\n", nsx, 1.01 * sig * sig * by / ss, sc.tol);
        }
    }
    std::vector<double> H((size_t)(batch * n * n));
    CK(cudaMemcpy(H.data(), dH, H.size() * 8, cudaMemcpyDeviceToHost));
    // reference: full for small problems, sampled entries for large ones
    const bool full = (double)n * n * k * batch < 3e9;
    double maxref = 0, maxerr = 0, maxasym = 0;
    int printed = 0;
    auto check = [&](int64_t b, int64_t i, int64_t j) {
        long double s = 0;
        const double* xi = &X[(size_t)((b * n + i) * ld)];
        const double* xj = &X[(size_t)((b * n + j) * ld)];
        for (int64_t r = 0; r < k; ++r) s += (long double)xi[r] * xj[r];
        const double ref = (double)(2 * s), got = H[(size_t)(b * n * n + i + j * n)];
        maxref = std::fmax(maxref, std::fabs(ref));
        const double err = std::fabs(got - ref);
        if (!(err <= 1e-9 * std::fabs(ref) + 1e-300) && printed < 12) {
            printf("  mismatch inst %lld (%lld,%lld): got %.17g ref %.17g\n", (long long)b, (long long)i, (long long)j, got, ref);
            ++printed;
        }
        if (!(err <= maxerr)) maxerr = err;
        const double asym = std::fabs(got - H[(size_t)(b * n * n + j + i * n)]);
        if (!(asym <= maxasym)) maxasym = asym;
    };
    if (full) {
        for (int64_t b = 0; b < batch; ++b)
            for (int64_t j = 0; j < n; ++j)
                for (int64_t i = 0; i < n; ++i) check(b, i, j);
    } else {
        std::uniform_int_distribution<int64_t> I(0, n - 1), B(0, batch - 1);
        for (int t = 0; t < 20000; ++t) check(B(rng), I(rng), I(rng));
        for (int64_t i = 0; i < n; i += 7) check(0, i, i);
        for (int64_t i = 0; i < n; ++i) { check(batch - 1, i, n - 1); check(0, 0, i); check(0, i, (i * 97 + 13) % n); }
    }
    printf("%s check: max|ref|=%.6e  max abs err=%.3e  rel(to max)=%.3e  max asym=%.3e\n", full ? "full" : "sampled", maxref, maxerr, maxerr / maxref,
           maxasym);
    printf(maxerr <= 1e-10 * maxref ? "PROBE PASS\n" : "PROBE FAIL\n");
    return 0;
}
