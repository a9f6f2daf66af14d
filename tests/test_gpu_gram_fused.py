"""GPU parity of the fused small-instance int8 Gram (csrc/kernels_gram_i8_fused.cu: one CTA pair per instance, tcgen05 cta_group::2,
digit planes never leave the chip) -- needs a B200: -m gpu.

Reference semantics: vecdot! + muladd! (functions.jl:548-576,702-709) then the symmetric MOI Hessian of moi_interop.jl:45-62, one
`update!` (model.jl:132-143) per instance of the batch.  Checked: (1) 1e-10 relative (normwise, of max |Q|; written per assert) against
the dense oracle on BASELINE configs[2]'s instance shape (n=256, k=256, m=64) with more instances than CTA pairs, (2) bit-identity with
the two-kernel int8 route and with the numpy restatement oracle/int8_slices.py (integer accumulation is exact, the scaling order is the
same), (3) ragged shapes (n < 128: the second CTA of the pair owns no column; n <= 64: one output half; k = 2), (4) the per-instance
certificate: instances the 5-plane bound refuses are computed by the fp64 kernel that follows, (5) permuted index maps and accumulation
next to a second quadratic term (the generic store path), (6) determinism."""
import os

import numpy as np
import pytest

from oracle import dense as D
from oracle import int8_slices as S
from tests.twin import assert_close

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(bool(os.environ.get("PMB_NO_TMA")) or os.environ.get("PMB_GRAM_MODE") is not None or bool(os.environ.get("PMB_NO_FUSED_I8")),
                                 reason="int8 route disabled by PMB_NO_TMA / PMB_GRAM_MODE / PMB_NO_FUSED_I8")]
RTOL = 1e-10


def ls_model(pm, n, k, m, batch, gram_mode=1, idxmap=None):
    model = pm.Model(batch=batch, gram_mode=gram_mode)
    x = model.variables(n)
    A, b, Cm, d = (pm.Parameter(model, shape=s, device=True) for s in ((k, n), (k,), (m, n), (m,)))
    r = A * x - b
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, Cm * x, "==", d)
    model.initialize(idxmap=idxmap)
    return model, (A, b, Cm, d)


def run(pm, vals, n, k, m, batch, gram_mode=1, idxmap=None):
    model, params = ls_model(pm, n, k, m, batch, gram_mode, idxmap)
    for p, v in zip(params, vals):
        p.upload(v)
    model.update()
    out = model.objective_arrays()
    lv = model.gram_level_counts()
    kinds = model.kernel_classes_per_update() if hasattr(model, "kernel_classes_per_update") else None
    model.close()
    return out, lv, kinds


def rand_vals(rng, n, k, m, batch, dist="uniform"):
    f = rng.random if dist == "uniform" else rng.standard_normal
    shp = lambda s: (batch,) + s if batch > 1 else s
    return [f(shp((k, n))), f(shp((k,))), f(shp((m, n))), f(shp((m,)))]


def test_cfg3_instance_shape_more_instances_than_pairs(monkeypatch):
    """n = 256, k = 256, m = 64 (BASELINE configs[2]), 300 instances: every CTA pair runs at least four instances back to back (operand
    panels, sigma and statistics buffers are reused with both parities)."""
    import parametron_b200 as pm
    monkeypatch.setenv("PMB_GRAM_I8_FUSED", "1")
    n, k, m, B = 256, 256, 64, 300
    rng = np.random.default_rng(1)
    vals = rand_vals(rng, n, k, m, B)
    (Q, q, c0), lv, _ = run(pm, vals, n, k, m, B)
    assert lv == (B, 0, 0)                      # uniform [0, 1) data certifies with 5 planes
    for inst in list(range(0, B, 37)) + [B - 1]:
        Qo, qo, c0o = D.least_squares_objective(vals[0][inst], vals[1][inst], np.arange(1, n + 1), n)
        assert_close(Q[inst], Qo, RTOL, f"Q[{inst}]"); assert_close(q[inst], qo, RTOL, f"q[{inst}]"); assert_close([c0[inst]], [c0o], RTOL, "c0")
        assert np.array_equal(Q[inst], Q[inst].T)
    # bit-identical to the numpy restatement of the int8 route (exact integer accumulation, same scaling order) on the upper triangle
    for inst in (0, 149, B - 1):
        Go, _ = S.gram(vals[0][inst], 5, alpha=2.0)
        iu = np.triu_indices(n)
        assert np.array_equal(Q[inst][iu], Go[iu])
    # and to the two-kernel int8 route
    monkeypatch.setenv("PMB_GRAM_I8_FUSED", "0")
    monkeypatch.setenv("PMB_GRAM_I8_ALWAYS", "1")
    (Q2, q2, c02), lv2, _ = run(pm, vals, n, k, m, B, gram_mode=5)
    assert lv2 == (B, 0, 0)
    assert np.array_equal(Q, Q2)
    assert_close(q, q2, 1e-13, "q"); assert_close(c0, c02, 1e-13, "c0")


@pytest.mark.parametrize("shape", [(256, 256), (200, 130), (129, 256), (128, 64), (100, 36), (64, 256), (40, 2), (1, 2), (255, 254), (256, 34)])
def test_ragged_shapes_forced_and_certified(monkeypatch, shape):
    import parametron_b200 as pm
    monkeypatch.setenv("PMB_GRAM_I8_FUSED", "1")
    n, k = shape
    m, B = 3, 5
    rng = np.random.default_rng(n * 1000 + k)
    vals = rand_vals(rng, n, k, m, B, "normal")
    for mode in (5, 1):
        (Q, q, c0), lv, _ = run(pm, vals, n, k, m, B, gram_mode=mode)
        assert sum(lv) == B and (mode == 1 or lv == (B, 0, 0))
        for inst in range(B):
            Qo, qo, c0o = D.least_squares_objective(vals[0][inst], vals[1][inst], np.arange(1, n + 1), n)
            assert_close(Q[inst], Qo, RTOL, f"Q[{inst}] mode {mode}"); assert_close(q[inst], qo, RTOL, "q"); assert_close([c0[inst]], [c0o], RTOL, "c0")
            assert np.array_equal(Q[inst], Q[inst].T)


def test_certificate_hands_uncertifiable_instances_to_fp64(monkeypatch):
    """six instances: uniform -> 5 planes; six decades of dynamic range -> refused by the 5-plane bound -> fp64 kernel; a column outside the
    fixed-point range -> fp64; an Inf entry -> fp64 (propagates as in the fp64 kernel); all-zero -> 5 planes (exact zero); uniform again."""
    import parametron_b200 as pm
    monkeypatch.setenv("PMB_GRAM_I8_FUSED", "1")
    n, k, m, B = 256, 256, 2, 6
    rng = np.random.default_rng(7)
    A = np.empty((B, k, n))
    A[0] = rng.random((k, n))
    A[1] = rng.standard_normal((k, n)) * 10.0 ** (6.0 * rng.random((k, n)) - 3.0)
    A[2] = rng.standard_normal((k, n)); A[2][:, 3] *= 1e200; A[2][:, 5] *= 1e-300
    A[3] = rng.standard_normal((k, n)); A[3][17, 11] = np.inf
    A[4] = 0.0
    A[5] = rng.random((k, n))
    vals = [A, rng.standard_normal((B, k)), rng.standard_normal((B, m, n)), rng.standard_normal((B, m))]
    (Q, q, c0), lv, _ = run(pm, vals, n, k, m, B)
    s1 = S.level(S.slice_columns(A[1])[2])
    assert lv == ((3, 0, 3) if s1 != 5 else (4, 0, 2)), (lv, s1)   # the numpy restatement of the certificate predicts instance 1's verdict
    for inst in (0, 1, 4, 5):
        Qo, qo, c0o = D.least_squares_objective(A[inst], vals[1][inst], np.arange(1, n + 1), n)
        assert_close(Q[inst], Qo, RTOL, f"Q[{inst}]"); assert_close(q[inst], qo, RTOL, f"q[{inst}]")
    assert not Q[4].any()
    with np.errstate(over="ignore", invalid="ignore"):
        Qo2 = 2.0 * A[2].T @ A[2]
    fin = np.isfinite(Qo2)
    assert np.array_equal(np.isfinite(Q[2]), fin)
    keep = np.arange(n) != 3
    assert_close(np.where(fin, Q[2], 0.0)[np.ix_(keep, keep)], np.where(fin, Qo2, 0.0)[np.ix_(keep, keep)], RTOL, "Q[2] finite part")
    Q3 = Q[3]
    assert not np.isfinite(Q3[11, :]).any() and not np.isfinite(Q3[:, 11]).any()
    keep = np.arange(n) != 11
    A3 = A[3].copy(); A3[17, 11] = 0.0
    assert_close(Q3[np.ix_(keep, keep)], (2.0 * A3.T @ A3)[np.ix_(keep, keep)], RTOL, "Q[3] finite part")


def test_permuted_map_and_accumulation_with_second_quadratic_term(monkeypatch):
    """objective = ||Ax - b||^2 + x'Px with a permuted index map: the fused kernel accumulates through the variable map (generic stores)."""
    import parametron_b200 as pm
    monkeypatch.setenv("PMB_GRAM_I8_FUSED", "1")
    n, k, B = 192, 128, 4
    rng = np.random.default_rng(11)
    Av, bv, Pm = rng.random((B, k, n)), rng.random((B, k)), rng.standard_normal((B, n, n))
    perm = rng.permutation(n) + 1
    outs = {}
    for mode in (1, 0):
        model = pm.Model(batch=B, gram_mode=mode)
        x = model.variables(n)
        A = pm.Parameter(model, shape=(k, n), device=True)
        b = pm.Parameter(model, shape=(k,), device=True)
        P = pm.Parameter(model, shape=(n, n), device=True)
        r = A * x - b
        pm.objective(model, pm.Minimize, pm.dot(r, r) + pm.bilinear(pm.adjoint(x, model=model), P, x))
        model.initialize(idxmap=perm)
        A.upload(Av); b.upload(bv); P.upload(Pm)
        model.update()
        outs[mode] = model.objective_arrays()
        lv = model.gram_level_counts()
        assert lv == ((B, 0, 0) if mode == 1 else (0, 0, B))
        model.update()   # a second update gives the same values: nothing stale accumulates (the bilinear scatter adds with atomics, so not bitwise)
        assert_close(model.objective_arrays()[0], outs[mode][0], 1e-14, "second update")
        model.close()
    pi = perm - 1
    for inst in range(B):
        Qo = np.zeros((n, n))
        Qo[np.ix_(pi, pi)] = 2.0 * Av[inst].T @ Av[inst] + Pm[inst] + Pm[inst].T
        qo = np.zeros(n); qo[pi] = -2.0 * Av[inst].T @ bv[inst]
        for mode in (1, 0):
            assert_close(outs[mode][0][inst], Qo, RTOL, f"Q[{inst}] mode {mode}"); assert_close(outs[mode][1][inst], qo, RTOL, "q")
            assert_close([outs[mode][2][inst]], [bv[inst] @ bv[inst]], RTOL, "c0")


def test_cost_model_keeps_single_small_instances_on_fp64():
    """batch 1 of README Example 1's shape stays on the fused-linear fp64 kernel (latency-bound: 3 launches per update)"""
    import parametron_b200 as pm
    rng = np.random.default_rng(0)
    vals = rand_vals(rng, 8, 8, 2, 1)
    (Q, q, c0), lv, _ = run(pm, vals, 8, 8, 2, 1)
    assert lv == (0, 0, 1)
    Qo, qo, c0o = D.least_squares_objective(vals[0], vals[1], np.arange(1, 9), 8)
    assert_close(Q, Qo, RTOL, "Q")


@pytest.mark.parametrize("n", [256, 200, 128, 100, 24])
def test_generator_fused_variant_on_ragged_instance_widths(monkeypatch, n):
    """gram_i8_fused_gen_kernel (the operand is a UNIFORM Parameter nothing else reads: the slicer warps generate their columns, the head of the
    next instance is parked in shared memory while the current products run) on instance widths where the second CTA of the pair owns a
    partial panel (200), nothing (128, 100) and where less than one parked stage exists (24); 7 instances = several per CTA pair only for the
    first pairs.  Q bit-identical to the loading kernel fed by the prologue's fill (PMB_NO_GEN_FUSE=1), 1e-10 against the dense oracle, and the
    Parameter reads back as the oracle's matrix (materialised on demand)."""
    import parametron_b200 as pm
    from oracle import philox
    monkeypatch.setenv("PMB_GRAM_I8_FUSED", "1")
    k, m, B, first = 256, 3, 7, 11
    outs = {}
    for tag in ("gen", "load"):
        monkeypatch.delenv("PMB_NO_GEN_FUSE", raising=False)
        if tag == "load":
            monkeypatch.setenv("PMB_NO_GEN_FUSE", "1")
        model = pm.Model(batch=B, instance_offset=first)
        x = model.variables(n)
        A = pm.Parameter(model, shape=(k, n), uniform=(77, 1.0))
        b = pm.Parameter(model, shape=(k,), uniform=(78, 1.0))
        Cm = pm.Parameter(model, shape=(m, n), uniform=(79, 1.0))
        d = pm.Parameter(model, shape=(m,), uniform=(80, 2.0))
        r = A * x - b
        pm.objective(model, pm.Minimize, pm.dot(r, r))
        pm.constraint(model, Cm * x, "==", d)
        model.initialize()
        for _ in range(3):
            model.update()
        assert model.gram_level_counts() == (B, 0, 0)
        outs[tag] = (model.objective_arrays(), A.download())
        model.close()
    assert np.array_equal(outs["gen"][0][0], outs["load"][0][0])          # Q: exact integer accumulation, same sigma -> bit-identical
    assert_close(outs["gen"][0][1], outs["load"][0][1], 1e-14, "q")     # q, c0: the generated column sits in registers in natural row order, the
    assert_close(outs["gen"][0][2], outs["load"][0][2], 1e-14, "c0")    # loaded one with swapped halves in half of the lanes: fp64 sums in another order
    assert np.array_equal(outs["gen"][1], outs["load"][1])
    Q, q, c0 = outs["gen"][0]
    for inst in range(B):
        Ao = philox.uniform_matrix(k, n, 77, 1.0, first + inst, 3)
        bo = philox.uniform(k, 78, 1.0, first + inst, 3)
        assert np.array_equal(outs["gen"][1][inst], Ao)
        Qo, qo, c0o = D.least_squares_objective(Ao, bo, np.arange(1, n + 1), n)
        assert_close(Q[inst], Qo, RTOL, f"Q[{inst}]"); assert_close(q[inst], qo, RTOL, "q"); assert_close([c0[inst]], [c0o], RTOL, "c0")
