"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the synthetic-input generator of SURVEY.md section 8(d): Philox4x32-10 (Salmon et al., SC'11; the
published round constants) producing U[0,1) fp64 with 53 random bits, the stand-in for the reference's
`Parameter(rand!, zeros(n, n), model)` (README.md:36-43).  The device generator (csrc/kernels_misc.cu
philox_uniform_kernel) must produce the same bits:
    element pair p = e // 2 of instance `inst` at update `upd`:  counter = (p_lo, p_hi, inst, upd), key = (seed_lo, seed_hi)
    out[2p]   = ((c1 << 32 | c0) >> 11) * 2^-53 * scale
    out[2p+1] = ((c3 << 32 | c2) >> 11) * 2^-53 * scale
Parity status: pinned on the Random123 known-answer vectors (tests/test_oracle_philox.py).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint32 arrays c0..c3; k0, k1 python ints.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def uniform(numel, seed, scale=1.0, inst=0, upd=1):
    """The `numel` doubles the device writes for one instance (flat, column-major order for matrices)."""
    pairs = (numel + 1) // 2
    p = np.arange(pairs, dtype=np.uint64)
    c0, c1, c2, c3 = philox4x32_10(p & MASK, p >> np.uint64(32), np.full(pairs, inst, dtype=np.uint64),
                                   np.full(pairs, upd, dtype=np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lo = (c1.astype(np.uint64) << np.uint64(32)) | c0.astype(np.uint64)
    hi = (c3.astype(np.uint64) << np.uint64(32)) | c2.astype(np.uint64)
    out = np.empty(2 * pairs)
    out[0::2] = (lo >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * scale
    out[1::2] = (hi >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * scale
    return out[:numel]


def uniform_matrix(rows, cols, seed, scale=1.0, inst=0, upd=1):
    """rows x cols matrix whose column-major flattening is `uniform(rows*cols, ...)`."""
    return uniform(rows * cols, seed, scale, inst, upd).reshape(cols, rows).T.copy()
