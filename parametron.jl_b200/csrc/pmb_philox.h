// Philox4x32-10 (Random123; the tests pin it on Random123's known-answer vectors) and the exact 53-bit -> fp64 conversion: shared by the
// fused prologue (kernels_misc.cu) and the generator-fused slice pass (kernels_gram_i8.cu) -- both must produce the same bits for the same
// (pair index, global instance id, update id, seed).
#pragma once
#include <stdint.h>

namespace pmb {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// two chains, rounds interleaved
__device__ __forceinline__ void philox4x32_10x2(uint32_t (&c)[4], uint32_t (&e)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t gi0 = __umulhi(0xD2511F53u, e[0]), go0 = 0xD2511F53u * e[0];
        const uint32_t gi1 = __umulhi(0xCD9E8D57u, e[2]), go1 = 0xCD9E8D57u * e[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        const uint32_t m0 = gi1 ^ e[1] ^ k0, m2 = gi0 ^ e[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        e[0] = m0; e[1] = go1; e[2] = m2; e[3] = go0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// (double)((hi:lo) >> 11), exactly, without the 64-bit integer conversion of the XU pipe: both words of the 53-bit integer become doubles
// through the 2^52 mantissa trick, one fma joins them (the sum has at most 53 significant bits: no rounding)
__device__ __forceinline__ double u53_to_f64(uint32_t hi, uint32_t lo) {
    const uint32_t xh = hi >> 11, xl = (lo >> 11) | (hi << 21);
    const double dh = __hiloint2double(0x43300000, (int)xh) - 4503599627370496.0;
    const double dl = __hiloint2double(0x43300000, (int)xl) - 4503599627370496.0;
    return fma(dh, 4294967296.0, dl);
}


// Element pair p of a UNIFORM Parameter (elements 2p, 2p + 1 of the instance's buffer) at update `upd`:  counter = (p, global instance, upd),
// key = seed;  value = ((hi:lo) >> 11) * 2^-53 * scale
__device__ __forceinline__ void philox_uniform_pairs2(uint64_t p0, uint64_t p1, uint32_t inst_global, uint32_t upd, uint64_t seed, double sc53, double (&v)[4]) {
    uint32_t c[4] = {(uint32_t)p0, (uint32_t)(p0 >> 32), inst_global, upd};
    uint32_t e[4] = {(uint32_t)p1, (uint32_t)(p1 >> 32), inst_global, upd};
    philox4x32_10x2(c, e, (uint32_t)seed, (uint32_t)(seed >> 32));
    v[0] = u53_to_f64(c[1], c[0]) * sc53; v[1] = u53_to_f64(c[3], c[2]) * sc53;
    v[2] = u53_to_f64(e[1], e[0]) * sc53; v[3] = u53_to_f64(e[3], e[2]) * sc53;
}

}  // namespace pmb
