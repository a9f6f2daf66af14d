"""TEST INFRASTRUCTURE ONLY (never imported by the package): numpy restatement of the certified int8-slice Gram of
parametron.jl_b200/csrc/kernels_gram_i8.cu, used to check on the CPU that (1) the digit planes represent the scaled operand exactly
as claimed, (2) the truncated digit-plane products reproduce `2 * X'X` to the accuracy the certificate promises, and (3) the
worst-case bound evaluated by `gram_i8_compute_level` (pmb_kernels.h) really dominates the actual error -- on benign and on
adversarial inputs.  The reference semantics being approximated is the fp64 contraction `vecdot!` + `muladd!`
(/root/reference/src/functions.jl:548-576,702-709) followed by the symmetric MOI Hessian of moi_interop.jl:45-62.

Everything below follows the device code line by line (same scale, same rounding, same Horner chain); integers are Python/NumPy
int64 so the plane products are exact, as they are in the int32 TMEM accumulators for k < 2^17 / 6."""
import numpy as np

PLANES = 6           # kGramI8Planes
MAGIC = 6755399441055744.0   # 1.5 * 2^52: the device rounds v * mult with one fma against this constant


def slice_columns(X):
    """gram_i8_slice_kernel: per column sigma = amax / 0.99, fixed point with 47 fractional bits, 6 balanced base-256 digits.
    Returns (digits [PLANES, k, n] int64 in [-128, 127], sigma [n], stats dict) -- stats as accumulated on the device."""
    X = np.asarray(X, dtype=np.float64)
    k, n = X.shape
    amax = np.abs(X).max(axis=0) if k else np.zeros(n)
    l1 = np.abs(X).sum(axis=0)
    ss = (X * X).sum(axis=0)
    bad = ~(l1 <= np.finfo(np.float64).max) | ((amax > 0) & ~((amax >= 1e-150) & (amax <= 1e150)))
    zero = ~(amax > 0) | bad
    safe = np.where(zero, 1.0, amax)
    sigma = np.where(zero, 1.0, amax * (1.0 / 0.99))
    mult = np.where(zero, 0.0, np.ldexp(0.99 / safe, 8 * PLANES - 1))
    # round-to-nearest-even of the exact product (the device uses ONE fma against MAGIC); long double keeps the product exact enough
    with np.errstate(invalid="ignore", over="ignore"):
        prod = np.where(np.isfinite(X), X, 0.0).astype(np.longdouble) * mult.astype(np.longdouble)
        Xi = np.rint(prod).astype(np.int64)
    Y = (Xi + 0x0000008080808080) ^ 0x0000008080808080        # carry-free balanced digits: byte b of Y is plane 5 - b
    digits = np.empty((PLANES, k, n), dtype=np.int64)
    for p in range(PLANES):
        b = (Y >> (8 * (PLANES - 1 - p))) & 0xFF
        digits[p] = np.where(b >= 128, b - 256, b)
    live = ~zero
    stats = {
        "bad": bool(bad.any()),
        "sigma_max": float(sigma[live].max()) if live.any() else 0.0,
        "ss_max": float(ss[live].max()) if live.any() else 0.0,
        "l1_max": float((l1[live] / sigma[live]).max()) if live.any() else 0.0,
        "dsq_max": [int((digits[p] ** 2).sum(axis=0)[live].max()) if live.any() else 0 for p in range(PLANES)],
    }
    return digits, sigma, stats


def reconstruct(digits, sigma, ns=PLANES):
    """value represented by the leading `ns` planes: sigma * sum_s d_s 2^(-7 - 8 s)"""
    y = sum(digits[s].astype(np.float64) * 2.0 ** (-7 - 8 * s) for s in range(ns))
    return y * sigma[None, :]


def level(stats, tol=9e-11):
    """gram_i8_compute_level (pmb_kernels.h): 5, 6 or 64 (= fp64)"""
    if stats["bad"]:
        return 64
    if stats["sigma_max"] == 0.0:
        return 5
    R = [np.sqrt(float(v)) for v in stats["dsq_max"]]
    for ns in (5, 6):
        by = bound_scaled(stats, ns, R)
        if 1.01 * stats["sigma_max"] ** 2 * by <= tol * stats["ss_max"]:
            return ns
    return 64


def bound_scaled(stats, ns, R=None):
    R = R or [np.sqrt(float(v)) for v in stats["dsq_max"]]
    by = 2.0 * 1.1 * 2.0 ** (-8.0 * ns) * stats["l1_max"]
    for s in range(1, ns):
        for t in range(1, ns):
            if s + t >= ns:
                by += 2.0 ** (-14.0 - 8.0 * (s + t)) * R[s] * R[t]
    return by


def error_bound(stats, ns):
    """worst-case |error| of any entry of X'X (before the caller's alpha), as certified on the device"""
    return 1.01 * stats["sigma_max"] ** 2 * bound_scaled(stats, ns)


def gram(X, ns, alpha=2.0):
    """gram_i8_kernel<ns, BN>: exact integer products of the digit planes with s + t <= ns - 1, one accumulator per level s + t,
    Horner chain in fp64, scale by sigma_i sigma_j 2^(-8 ns - 6) and alpha.  Returns the full symmetric matrix."""
    digits, sigma, stats = slice_columns(X)
    n = X.shape[1]
    acc = [np.zeros((n, n), dtype=np.int64) for _ in range(ns)]
    for s in range(ns):
        for t in range(ns - s):
            acc[s + t] += digits[s].T @ digits[t]
    assert all(np.abs(a).max(initial=0) < 2 ** 31 for a in acc), "int32 accumulator overflow"
    T = acc[0].astype(np.float64)
    for lv in range(1, ns):
        T = T * 256.0 + acc[lv].astype(np.float64)
    f = sigma * 2.0 ** (-4.0 * ns - 3.0)
    return (T * (alpha * f)[:, None]) * f[None, :], stats
