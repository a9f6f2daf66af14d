"""GPU parity at the BASELINE.json sizes and through size-independent properties (needs a B200: -m gpu).

The literal term list cannot exist at n=4096 (1.65 TB), so parity is transitive:
device == dense oracle (numpy/BLAS) at full size, and dense oracle == literal oracle at small size (tests/test_oracle_*)
plus device == literal oracle at small size (tests/test_gpu_parity.py).  Tolerance 1e-10 relative, written per assert."""
import os

import numpy as np
import pytest

from oracle import dense as D
from oracle import philox
from oracle.termlist import TermListModel
from tests.twin import assert_close

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def ls_model(pm, n, k, m, batch=1, seeds=(1234, 1235, 1236, 1237), mode="uniform", host=None):
    model = pm.Model(batch=batch)
    x = model.variables(n)
    if mode == "uniform":
        A = pm.Parameter(model, shape=(k, n), uniform=(seeds[0], 1.0))
        b = pm.Parameter(model, shape=(k,), uniform=(seeds[1], 1.0))
        Cm = pm.Parameter(model, shape=(m, n), uniform=(seeds[2], 1.0))
        d = pm.Parameter(model, shape=(m,), uniform=(seeds[3], 2.0))
    elif mode == "device":
        A, b, Cm, d = (pm.Parameter(model, shape=s, device=True) for s in ((k, n), (k,), (m, n), (m,)))
    else:
        A, b, Cm, d = (pm.Parameter(model, val=host[key]) for key in "AbCd")
    r = A * x - b
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, Cm * x, "==", d)
    return model, (A, b, Cm, d)


def oracle_inputs(n, k, m, inst, upd, seeds=(1234, 1235, 1236, 1237)):
    return (philox.uniform_matrix(k, n, seeds[0], 1.0, inst, upd), philox.uniform(k, seeds[1], 1.0, inst, upd),
            philox.uniform_matrix(m, n, seeds[2], 1.0, inst, upd), philox.uniform(m, seeds[3], 2.0, inst, upd))


def test_device_generator_matches_philox_oracle_bit_exact():
    import parametron_b200 as pm
    model, (A, b, Cm, d) = ls_model(pm, 33, 17, 5, batch=3)
    model.initialize()
    for upd in (1, 2, 3):
        model.update()
        Ad, bd, Cd, dd = A.download(), b.download(), Cm.download(), d.download()
        for inst in range(3):
            Ao, bo, Co, do = oracle_inputs(33, 17, 5, inst, upd)
            assert np.array_equal(Ad[inst], Ao) and np.array_equal(bd[inst], bo)
            assert np.array_equal(Cd[inst], Co) and np.array_equal(dd[inst], do)
    assert model.update_count() == 3


def test_cfg2_full_size_against_dense_oracle():
    """BASELINE.json configs[1]: n=4096, m=1024, dense random A,b,C,d; fp64."""
    import parametron_b200 as pm
    n, k, m = 4096, 4096, 1024
    model, _ = ls_model(pm, n, k, m)
    model.initialize()
    for upd in (1, 2):
        model.update()
    Q, q, c0 = model.objective_arrays()
    Cb, db = model.block_arrays(0)
    A, b, Cm, d = oracle_inputs(n, k, m, 0, 2)
    Qo, qo, c0o = D.least_squares_objective(A, b, np.arange(1, n + 1), n)
    assert_close(Q, Qo, RTOL, "Q"); assert_close(q, qo, RTOL, "q"); assert_close([c0], [c0o], RTOL, "c0")
    assert np.array_equal(Cb, Cm) and np.array_equal(db, -d)  # pure copies / negations: bit-exact
    # size-independent properties
    assert np.array_equal(Q, Q.T)                      # both triangles come from the same accumulators
    assert (np.diag(Q) > 0).all()
    xs = np.random.default_rng(0).standard_normal(n)   # 1/2 x'Qx + q'x + c0 == ||Ax-b||^2   (test/model.jl:84)
    lhs = 0.5 * xs @ Q @ xs + q @ xs + c0
    rhs = float(np.sum((A @ xs - b) ** 2))
    assert abs(lhs - rhs) <= 1e-9 * abs(rhs)
    st = model.stats()
    # the block C x == d aliases the (library-filled) Parameter: only the offsets are assembled; with the copy it is 16 m n + 16 m
    assert st["affine_algorithmic_bytes"] == (16 * m * n + 16 * m if os.environ.get("PMB_NO_BLOCK_ALIAS") else 16 * m)
    assert st["gram_algorithmic_flops"] == k * n * (n + 1) + 2 * k * n + 2 * k
    assert model.block_view(0).nterms == m * n


def test_idempotent_and_deterministic_with_manual_device_parameters():
    """manual mode (parameter.jl:88): unchanged Parameters -> bit-identical outputs on every update (no atomics on this path)."""
    import parametron_b200 as pm
    n, k, m = 512, 384, 128
    rng = np.random.default_rng(2)
    model, (A, b, Cm, d) = ls_model(pm, n, k, m, mode="device")
    model.initialize()
    vals = [rng.standard_normal(s) for s in ((k, n), (k,), (m, n), (m,))]
    for p, v in zip((A, b, Cm, d), vals):
        p.upload(v)
    model.update()
    first = model.objective_arrays() + model.block_arrays(0)
    model.update()
    second = model.objective_arrays() + model.block_arrays(0)
    for a, c in zip(first, second):
        assert np.array_equal(a, c)
    Qo, qo, c0o = D.least_squares_objective(vals[0], vals[1], np.arange(1, n + 1), n)
    assert_close(first[0], Qo, RTOL, "Q"); assert_close(first[1], qo, RTOL, "q"); assert_close([first[2]], [c0o], RTOL, "c0")
    # linearity in b: doubling b doubles q, quadruples c0, leaves Q unchanged
    b.upload(2 * vals[1])
    model.update()
    Q2, q2, c2 = model.objective_arrays()
    assert np.array_equal(Q2, first[0]) and np.array_equal(q2, 2 * first[1]) and c2 == 4 * first[2]


def test_permuted_index_map_medium():
    import parametron_b200 as pm
    n, k, m = 300, 257, 65  # odd sizes: the unaligned (8-byte) load paths
    rng = np.random.default_rng(3)
    host = {"A": rng.standard_normal((k, n)), "b": rng.standard_normal(k), "C": rng.standard_normal((m, n)), "d": rng.standard_normal(m)}
    perm = rng.permutation(n) + 1
    model, _ = ls_model(pm, n, k, m, mode="host", host=host)
    model.initialize(idxmap=perm)
    model.update()
    Q, q, c0 = model.objective_arrays()
    Cb, db = model.block_arrays(0)
    Qo, qo, c0o = D.least_squares_objective(host["A"], host["b"], np.arange(1, n + 1), n, varmap=perm)
    Co, do = D.affine_block(host["C"], host["d"], np.arange(1, n + 1), n, varmap=perm)
    assert_close(Q, Qo, RTOL, "Q"); assert_close(q, qo, RTOL, "q"); assert_close([c0], [c0o], RTOL, "c0")
    assert np.array_equal(Cb, Co) and np.array_equal(db, do)
    rows, vars_, src = model.block_terms(0)
    assert np.array_equal(rows, np.repeat(np.arange(1, m + 1), n))           # row-outer
    assert np.array_equal(vars_, np.tile(perm, m))                          # term-inner, varmap applied (moi_interop.jl:64-81)


def test_against_c_termlist_port_small():
    """device == the literal C port (the thing timed as the CPU baseline) after densification."""
    import parametron_b200 as pm
    n, k, m = 24, 20, 6
    rng = np.random.default_rng(4)
    host = {"A": rng.random((k, n)), "b": rng.random(k), "C": rng.random((m, n)), "d": rng.random(m)}
    model, _ = ls_model(pm, n, k, m, mode="host", host=host)
    model.initialize()
    model.update()
    t = TermListModel(n, k, m)
    t.update(host["A"], host["b"], host["C"], host["d"])
    c, v1, v2 = t.quadratic_terms()
    Qo = np.zeros((n, n))
    np.add.at(Qo, (v1 - 1, v2 - 1), c)
    off = v1 != v2
    np.add.at(Qo, (v2[off] - 1, v1[off] - 1), c[off])   # MOI: off-diagonal terms count for both triangles
    ca, va, const = t.affine_terms()
    qo = np.zeros(n); np.add.at(qo, va - 1, ca)
    Q, q, c0 = model.objective_arrays()
    assert_close(Q, Qo, RTOL, "Q"); assert_close(q, qo, RTOL, "q"); assert_close([c0], [const], RTOL, "c0")
    rr, cc, vv, kk = t.vector_terms()
    rows, vars_, src = model.block_terms(0)
    assert np.array_equal(rows, rr) and np.array_equal(vars_, vv)
    Cb, db = model.block_arrays(0)
    assert np.array_equal(Cb.T.ravel()[src], cc) and np.array_equal(db, kk)


@pytest.mark.parametrize("batch,n,k,m", [(5, 64, 48, 16), (64, 256, 256, 64)])
def test_batched_instances(batch, n, k, m):
    """BASELINE.json configs[2] shape (n=256, m=64), reduced batch: every instance independent, batch-major layout."""
    import parametron_b200 as pm
    model, _ = ls_model(pm, n, k, m, batch=batch)
    model.initialize()
    model.update()
    Q, q, c0 = model.objective_arrays()
    Cb, db = model.block_arrays(0)
    assert Q.shape == (batch, n, n) and Cb.shape == (batch, m, n)
    for inst in range(batch):
        A, b, Cm, d = oracle_inputs(n, k, m, inst, 1)
        Qo, qo, c0o = D.least_squares_objective(A, b, np.arange(1, n + 1), n)
        assert_close(Q[inst], Qo, RTOL, f"Q[{inst}]"); assert_close(q[inst], qo, RTOL, f"q[{inst}]")
        assert_close([c0[inst]], [c0o], RTOL, f"c0[{inst}]")
        assert np.array_equal(Cb[inst], Cm) and np.array_equal(db[inst], -d)


def test_empty_and_degenerate_shapes():
    """ragged / empty inputs: zero-row constraint block, one-variable model, k < n (rank-deficient Gram)."""
    import parametron_b200 as pm
    rng = np.random.default_rng(6)
    for n, k, m in [(1, 1, 1), (7, 3, 0), (130, 5, 2), (5, 130, 1)]:
        host = {"A": rng.random((k, n)), "b": rng.random(k), "C": rng.random((m, n)), "d": rng.random(m)}
        model, _ = ls_model(pm, n, k, m, mode="host", host=host)
        model.initialize()
        model.update()
        Q, q, c0 = model.objective_arrays()
        Qo, qo, c0o = D.least_squares_objective(host["A"], host["b"], np.arange(1, n + 1), n)
        assert_close(Q, Qo, RTOL, "Q"); assert_close(q, qo, RTOL, "q"); assert_close([c0], [c0o], RTOL, "c0")
        Cb, db = model.block_arrays(0)
        assert Cb.shape == (m, n) and np.array_equal(Cb, host["C"]) and np.array_equal(db, -host["d"])
        model.close()


def _random_bsr(rng, rows, cols, b, density):
    nbr, nbc = rows // b, cols // b
    mask = rng.random((nbr, nbc)) < density
    mask[0, 0] = True
    br, bc = np.nonzero(mask)
    return br.astype(np.int64), bc.astype(np.int64)


@pytest.mark.parametrize("rows,cols,b,density,perm", [(64, 128, 16, 0.2, False), (96, 160, 32, 0.3, True), (4096, 16384, 32, 0.01, False)])
def test_block_sparse_constraint(rows, cols, b, density, perm):
    """BASELINE.json configs[3]: 1 % block-sparse C, affine-only assembly.  The reference has no sparse path (matvecmul! visits
    every A[i], functions.jl:790-796), so parity is against the dense meaning: the sparse block densified == (C_dense, -d)."""
    import parametron_b200 as pm
    rng = np.random.default_rng(rows + b)
    br, bc = _random_bsr(rng, rows, cols, b, density)
    blocks = rng.standard_normal((len(br), b, b))
    dv = rng.standard_normal(rows)
    model = pm.Model()
    x = model.variables(cols)
    Cp = pm.Parameter(model, val=blocks, bsr=(rows, cols, b, br, bc))
    dp = pm.Parameter(model, val=dv)
    pm.constraint(model, Cp * x, "==", dp)
    vm = (rng.permutation(cols) + 1) if perm else None
    model.initialize(idxmap=vm)
    model.update()
    v = model.block_view(0)
    assert v.is_sparse == 1 and v.nnz == len(br) * b * b and not v.C
    vals, rws, cls_, d = model.block_sparse(0)
    dense = np.zeros((rows, cols))
    for k_, (i, j) in enumerate(zip(br, bc)):
        dense[i * b:(i + 1) * b, j * b:(j + 1) * b] = blocks[k_]
    Co, do = D.affine_block(dense, dv, np.arange(1, cols + 1), cols, varmap=vm)
    got = np.zeros((rows, cols))
    got[rws, cls_] = vals
    assert np.array_equal(got, Co) and np.array_equal(d, do)      # pure copies: bit-exact
    assert model.stats()["affine_algorithmic_bytes"] == 16 * v.nnz + 16 * rows
    # values-only update over the static index
    blocks2 = rng.standard_normal(blocks.shape)
    Cp.val = blocks2
    model.update()
    vals2, rws2, cls2, _ = model.block_sparse(0)
    assert np.array_equal(rws2, rws) and np.array_equal(cls2, cls_)
    assert np.array_equal(vals2, np.transpose(blocks2, (0, 2, 1)).ravel())
    model.close()


def test_block_sparse_small_against_literal_oracle():
    """small enough for the literal term list: the oracle runs the DENSE matrix (zeros included), densified results must agree."""
    import parametron_b200 as pm
    from tests.twin import OracleSide, oracle_outputs
    rows, cols, b = 8, 12, 4
    rng = np.random.default_rng(8)
    br, bc = _random_bsr(rng, rows, cols, b, 0.5)
    blocks = rng.random((len(br), b, b))
    dv = rng.random(rows)
    dense = np.zeros((rows, cols))
    for k_, (i, j) in enumerate(zip(br, bc)):
        dense[i * b:(i + 1) * b, j * b:(j + 1) * b] = blocks[k_]
    o = OracleSide()
    xo = o.variables(cols)
    o.constraint(o.mul(o.parameter({"v": dense}), xo), ">=", o.parameter({"v": dv}))
    o.model.initialize()
    _, oblocks, _ = oracle_outputs(o.model, cols)
    model = pm.Model()
    x = model.variables(cols)
    pm.constraint(model, pm.Parameter(model, val=blocks, bsr=(rows, cols, b, br, bc)) * x, ">=", pm.Parameter(model, val=dv))
    model.initialize()
    model.update()
    Cd, dd = model.block_arrays(0)
    assert np.array_equal(Cd, oblocks[0][0]) and np.array_equal(dd, oblocks[0][1])
    assert model.block_view(0).nterms == rows * cols  # what the reference emits: every entry, zeros included


@pytest.mark.parametrize("env", ["PMB_NO_TMA", "PMB_NO_GRAPH"])
def test_fallback_paths_agree(env, monkeypatch):
    """The cp.async Gram / LDG-STG affine fallbacks (PMB_NO_TMA=1) and eager launches (PMB_NO_GRAPH=1) produce the same arrays
    as the default TMA + CUDA-graph path (the switches are read when the context is created)."""
    import parametron_b200 as pm
    n, k, m = 384, 320, 96
    rng = np.random.default_rng(12)
    host = {"A": rng.standard_normal((k, n)), "b": rng.standard_normal(k), "C": rng.standard_normal((m, n)), "d": rng.standard_normal(m)}
    outs = []
    for on in (False, True):
        if on:
            monkeypatch.setenv(env, "1")
        model, _ = ls_model(pm, n, k, m, mode="host", host=host)
        model.initialize()
        model.update()
        assert model.stats()["uses_cuda_graph"] == (0 if (on and env == "PMB_NO_GRAPH") else 1)
        outs.append(model.objective_arrays() + model.block_arrays(0))
        model.close()
    for a, b_ in zip(*outs):
        assert_close(a, b_, 1e-12, env)


# ---- BASELINE configs[2] (cfg3) at full batch, instance-shard partition, cfg5 drift, elementwise error report (round 2) --------------

def _ls_model_shard(pm, n, k, m, batch, first, gram_mode=None, seeds=(1234, 1235, 1236, 1237)):
    model = pm.Model(batch=batch, instance_offset=first, gram_mode=gram_mode)
    x = model.variables(n)
    A = pm.Parameter(model, shape=(k, n), uniform=(seeds[0], 1.0))
    b = pm.Parameter(model, shape=(k,), uniform=(seeds[1], 1.0))
    Cm = pm.Parameter(model, shape=(m, n), uniform=(seeds[2], 1.0))
    d = pm.Parameter(model, shape=(m,), uniform=(seeds[3], 2.0))
    r = A * x - b
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, Cm * x, "==", d)
    model.initialize()
    return model, (A, b, Cm, d)


@pytest.mark.parametrize("gram_mode", [1, 0])
def test_instance_shard_equals_slice_of_unsharded_batch(gram_mode, monkeypatch):
    """rank r of G owns the contiguous global instances [first, first + count) (pm.shard.instance_shard); the built-in
    generators are keyed by the GLOBAL instance id (pmb_set_instance_offset), so a shard's Parameters, (C, d) and -- on the int8 route,
    whose integer accumulation does not depend on tiling -- (Q, q, c0) are bit-identical to the matching slice of the unsharded batch.
    On the fp64 route the k-split of the last wave of tiles depends on the batch, so Q agrees to rounding (1e-13), not bitwise.
    Ref: the same update! loop per instance, model.jl:132-143; BASELINE.json configs[2]."""
    import parametron_b200 as pm
    if gram_mode == 1:
        monkeypatch.setenv("PMB_GRAM_I8_FUSED", "1")
    n, k, m, total = 256, 256, 64, 50
    full, fp = _ls_model_shard(pm, n, k, m, total, 0, gram_mode)
    for _ in range(2):
        full.update()
    ref = full.objective_arrays() + full.block_arrays(0) + tuple(p.download() for p in fp)
    assert full.gram_level_counts() == ((total, 0, 0) if gram_mode == 1 else (0, 0, total))
    full.close()
    for world in (2, 3, 8):
        spans = [pm.shard.instance_shard(total, r, world) for r in range(world)]
        assert sum(c for _, c in spans) == total
        for first, count in spans:
            if count == 0:
                continue
            sh, sp = _ls_model_shard(pm, n, k, m, count, first, gram_mode)
            for _ in range(2):
                sh.update()
            got = sh.objective_arrays() + sh.block_arrays(0) + tuple(p.download() for p in sp)
            sh.close()
            for name, g, r_ in zip(("Q", "q", "c0", "C", "d", "A", "b", "Cm", "dm"), got, ref):
                g = np.asarray(g).reshape((count,) + np.asarray(r_).shape[1:])
                want = np.asarray(r_)[first:first + count]
                if gram_mode == 0 and name in ("Q", "q", "c0"):
                    assert_close(g, want, 1e-13, f"{name} shard [{first}, {first + count}) of {world}")
                else:
                    assert np.array_equal(g, want), f"{name} shard [{first}, {first + count}) of {world}"


def test_cfg3_full_batch_sampled_against_dense_oracle():
    """BASELINE.json configs[2] at its full size: 8192 independent instances of n=256, m=64 on one GPU; 72 sampled instances (first,
    last, around every 128th) against the dense oracle at 1e-10 relative (normwise, of max |Q| per instance), every instance through
    size-independent properties computed on the device side of the copy (finite, symmetric, positive diagonal)."""
    import parametron_b200 as pm
    n, k, m, B = 256, 256, 64, 8192
    model, _ = _ls_model_shard(pm, n, k, m, B, 0)
    for _ in range(2):
        model.update()
    lv = model.gram_level_counts()
    assert sum(lv) == B and lv[2] == 0          # uniform data certifies on the int8 route (fused kernel) for every instance
    ov, bv = model.objective_view(), model.block_view(0)
    sample = sorted(set([0, 1, B - 1, B - 2] + list(range(73, B, 128)) + [4095, 4096, 4097, 147, 148, 149]))
    for inst in sample:
        Q = model._d2h(ov.Q + inst * n * n * 8, n * n).reshape(n, n).T
        q = model._d2h(ov.q + inst * n * 8, n)
        c0 = model._d2h(ov.c0 + inst * 8, 1)
        Cb = model._d2h(bv.C + inst * m * n * 8, m * n).reshape(n, m).T
        db = model._d2h(bv.d + inst * m * 8, m)
        A, b, Cm, d = oracle_inputs(n, k, m, inst, 2)
        Qo, qo, c0o = D.least_squares_objective(A, b, np.arange(1, n + 1), n)
        assert_close(Q, Qo, RTOL, f"Q[{inst}]"); assert_close(q, qo, RTOL, f"q[{inst}]"); assert_close(c0, [c0o], RTOL, f"c0[{inst}]")
        assert np.array_equal(Q, Q.T)
        assert np.array_equal(Cb, Cm) and np.array_equal(db, -d)
    # every instance: trace(Q) = 2 ||A||_F^2 > 0 and c0 = b.b > 0, read as two small arrays (checksum of checksums)
    c0_all = model._d2h(ov.c0, B)
    q_all = model._d2h(ov.q, B * n).reshape(B, n)
    assert np.isfinite(c0_all).all() and (c0_all > 0).all() and np.isfinite(q_all).all() and (q_all < 0).all()
    model.close()


def test_cfg5_ten_thousand_sequential_updates_no_drift():
    """BASELINE.json configs[4] (MPC style): n=512, m=256, A and C resident and fixed, b and d re-parameterised on the device every
    solve; after 10 000 sequential updates (graph replays) the outputs equal the oracle evaluated at update 10 000 -- nothing
    accumulates across updates (the reference re-materialises everything per update!, model.jl:132-143)."""
    import parametron_b200 as pm
    n, k, m, solves = 512, 512, 256, 10000
    rng = np.random.default_rng(5)
    Av, Cv = rng.random((k, n)), rng.random((m, n))
    model = pm.Model()
    x = model.variables(n)
    A = pm.Parameter(model, shape=(k, n), device=True)
    Cm = pm.Parameter(model, shape=(m, n), device=True)
    b = pm.Parameter(model, shape=(k,), uniform=(1235, 1.0))
    d = pm.Parameter(model, shape=(m,), uniform=(1237, 2.0))
    r = A * x - b
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, Cm * x, "==", d)
    model.initialize()
    A.upload(Av); Cm.upload(Cv)
    model.update()
    Q1 = model.objective_arrays()[0]
    for _ in range(solves - 1):
        model.update()
    assert model.update_count() == solves
    Q, q, c0 = model.objective_arrays()
    Cb, db = model.block_arrays(0)
    bo, do = philox.uniform(k, 1235, 1.0, 0, solves), philox.uniform(m, 1237, 2.0, 0, solves)
    assert np.array_equal(b.download(), bo) and np.array_equal(d.download(), do)
    Qo, qo, c0o = D.least_squares_objective(Av, bo, np.arange(1, n + 1), n)
    assert_close(Q, Qo, RTOL, "Q after 10k updates"); assert_close(q, qo, RTOL, "q after 10k updates"); assert_close([c0], [c0o], RTOL, "c0")
    assert np.array_equal(Q, Q1)                        # A did not change: the Hessian is bit-identical to the first update's
    assert np.array_equal(Cb, Cv) and np.array_equal(db, -do)
    model.close()


@pytest.mark.parametrize("dist", ["uniform", "normal"])
def test_cfg2_elementwise_error_report_both_routes(dist):
    """The 1e-10 of north_star is asserted NORMWISE (max |Q - Qref| <= 1e-10 max |Qref|) on both routes -- that is the contract the int8
    certificate bounds.  The ELEMENTWISE relative error |Q_ij - ref_ij| / |ref_ij| is measured here on 40 000 sampled entries against a
    long-double reference and pinned: for U[0,1) data (cfg2's synthetic input: no cancellation, every entry ~ k/2) it meets 1e-10
    entry by entry on both routes; for standard-normal data the off-diagonal entries are small by cancellation (|Q_ij| ~ sqrt(k) against
    a diagonal ~ k) and the fixed-point route loses there the digits its 40-bit operands do not carry, the fp64 route does not."""
    import json
    import parametron_b200 as pm
    n = k = 4096
    rng = np.random.default_rng(17)
    Av = rng.random((k, n)) if dist == "uniform" else rng.standard_normal((k, n))
    bv = rng.random(k)
    idx = rng.integers(0, n, size=(40000, 2))
    idx[:2000, 1] = idx[:2000, 0]                        # some diagonal entries
    AL = Av.astype(np.longdouble)
    ref = np.empty(len(idx), dtype=np.longdouble)
    for s in range(0, len(idx), 2000):
        sl = slice(s, s + 2000)
        ref[sl] = 2 * np.einsum("ij,ij->j", AL[:, idx[sl, 0]], AL[:, idx[sl, 1]])
    report = {"dist": dist, "n": n, "k": k, "entries_sampled": len(idx)}
    for mode, name in ((1, "int8"), (0, "fp64")):
        model = pm.Model(gram_mode=mode)
        x = model.variables(n)
        A = pm.Parameter(model, shape=(k, n), device=True)
        b = pm.Parameter(model, shape=(k,), device=True)
        r = A * x - b
        pm.objective(model, pm.Minimize, pm.dot(r, r))
        model.initialize()
        A.upload(Av); b.upload(bv)
        model.update()
        lv = model.gram_level_counts()
        Q = model.objective_arrays()[0]
        model.close()
        got = Q[idx[:, 0], idx[:, 1]].astype(np.longdouble)
        err = np.abs(got - ref)
        scale = float(np.abs(ref).max())
        normwise = float(err.max()) / scale
        elem = (err / np.abs(ref)).astype(np.float64)
        report[name] = {"levels": lv, "normwise_max": normwise, "elementwise_max": float(elem.max()), "elementwise_p999": float(np.quantile(elem, 0.999)),
                        "elementwise_median": float(np.median(elem)), "smallest_abs_ref_over_max": float(np.abs(ref).min()) / scale}
        assert normwise <= RTOL, (name, normwise)            # the contract (normwise 1e-10), both routes
        if dist == "uniform":
            assert elem.max() <= RTOL, (name, float(elem.max()))   # no cancellation: 1e-10 holds entry by entry as well
        elif name == "fp64":      # measured (profiles/r02_elementwise_error_cfg2_normal.json): max 1.0e-10, 99.9 % 4.5e-13, median 1.5e-15
            assert elem.max() <= 1e-8 and np.quantile(elem, 0.999) <= 1e-11, (float(elem.max()), float(np.quantile(elem, 0.999)))
        else:                     # measured: max 7.7e-7 (an entry 3.6e-7 of max |Q|), 99.9 % 4.7e-9, median 8.6e-12
            assert elem.max() <= 1e-5 and np.quantile(elem, 0.999) <= 1e-7 and np.median(elem) <= 1e-10, (float(elem.max()), float(np.quantile(elem, 0.999)))
    if dist == "uniform":
        assert report["int8"]["levels"][2] == 0
    print("ELEMENTWISE_REPORT " + json.dumps(report))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/elementwise_error_cfg2_{dist}.json", "w") as f:
        json.dump(report, f, indent=1)


@pytest.mark.timeout(180)
def test_cfg3_back_to_back_updates_without_synchronisation():
    """Regression for a class of bug a single update cannot show: 600 updates of the cfg3 graph (generators -> fused int8 Gram || affine
    assembly) queued back to back with no host synchronisation in between -- a protocol error between the warps of the persistent fused kernel
    that needs a particular interleaving with the concurrent affine branch shows up as a hang (bounded here by the test's time-out) or a launch
    failure (a withdrawn rewrite of the kernel failed this within ~300 updates; tools/stress_cfg3.py is the long version).  After the last
    update the outputs equal the oracle evaluated at that update counter: nothing stale, nothing skipped.  BASELINE.json configs[2] shape,
    2048 instances."""
    import parametron_b200 as pm
    n, k, m, B, U = 256, 256, 64, 2048, 600
    model, _ = _ls_model_shard(pm, n, k, m, B, 0)
    for _ in range(U):
        model.update()
    model.synchronize()
    assert model.gram_level_counts() == (B, 0, 0)
    ov = model.objective_view()
    for inst in (0, 777, B - 1):
        Q = model._d2h(ov.Q + inst * n * n * 8, n * n).reshape(n, n).T
        q = model._d2h(ov.q + inst * n * 8, n)
        A, b, _, _ = oracle_inputs(n, k, m, inst, U)
        Qo, qo, _ = D.least_squares_objective(A, b, np.arange(1, n + 1), n)
        assert_close(Q, Qo, RTOL, f"Q[{inst}] after {U} updates"); assert_close(q, qo, RTOL, f"q[{inst}]")
    model.close()


def test_prologue_split_is_bit_identical_to_the_single_launch(monkeypatch):
    """The generators of constraint-only Parameters (here C: 8 MB, d) run as a second prologue launch on the constraint branch (update id taken
    from the already published counter instead of counter + 1); PMB_NO_PROLOGUE_SPLIT=1 keeps the single launch.  Every Parameter and every
    output must be equal bit for bit, over several updates.  Ref: each Parameter is evaluated exactly once per update!, parameter.jl:93-104."""
    import parametron_b200 as pm
    n, k, m = 1024, 256, 1024
    outs = []
    for single in (True, False):
        monkeypatch.delenv("PMB_NO_PROLOGUE_SPLIT", raising=False)
        if single:
            monkeypatch.setenv("PMB_NO_PROLOGUE_SPLIT", "1")
        model, _ = _ls_model_shard(pm, n, k, m, 1, 0)
        for _ in range(3):
            model.update()
        launches = model.stats()["kernel_launches_per_update"]
        outs.append((model.objective_arrays() + model.block_arrays(0), launches))
        model.close()
    assert outs[1][1] == outs[0][1] + 1          # the second prologue launch exists
    for a, b_ in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(np.asarray(a), np.asarray(b_))
    A, b, Cm, d = oracle_inputs(n, k, m, 0, 3)
    assert np.array_equal(outs[1][0][3], Cm) and np.array_equal(outs[1][0][4], -d)


def test_generator_fused_slice_pass_is_bit_identical_and_lazily_materialised(monkeypatch):
    """A UNIFORM matrix whose only reader is the int8 slice pass of the objective is generated INSIDE that pass (no prologue fill, no
    read-back) and its buffer is only written on demand by pmb_param_device_ptr.  The objective must equal the unfused schedule
    (PMB_NO_GEN_FUSE=1) and the write-every-update variant (PMB_GEN_MATERIALIZE=1) bit for bit over several updates, batch 1 and batch 2
    (per-instance Philox keys), and A.download() must return the oracle's matrix of the LAST update -- also when it is called only after
    later updates.  Ref: each Parameter is evaluated once per update! (parameter.jl:93-104); values are what rand! would have produced."""
    import parametron_b200 as pm
    n, k, m = 1536, 2048, 8
    for batch in (1, 2):
        runs = {}
        for tag, env in (("fused", None), ("unfused", "PMB_NO_GEN_FUSE"), ("writeback", "PMB_GEN_MATERIALIZE")):
            monkeypatch.delenv("PMB_NO_GEN_FUSE", raising=False); monkeypatch.delenv("PMB_GEN_MATERIALIZE", raising=False)
            if env:
                monkeypatch.setenv(env, "1")
            model, (A, b, Cm, d) = _ls_model_shard(pm, n, k, m, batch, 5, gram_mode=5)
            cls = None
            for _ in range(3):
                model.update()
            kt = model.stats()
            out = model.objective_arrays()
            Ad = A.download()                      # materialises update 3 on the fused route
            runs[tag] = (out, Ad, kt["kernel_launches_per_update"])
            if tag == "fused":
                model.update(); model.update()     # two more updates without touching the Parameter ...
                A5 = A.download()                  # ... then its values are those of update 5
                for inst in range(batch):
                    Ao = oracle_inputs(n, k, m, 5 + inst, 5)[0]
                    assert np.array_equal(A5[inst] if batch > 1 else A5, Ao)
            model.close()
        for inst in range(batch):
            Ao = oracle_inputs(n, k, m, 5 + inst, 3)[0]
            for tag in runs:
                Ad = runs[tag][1]
                assert np.array_equal(Ad[inst] if batch > 1 else Ad, Ao), tag
        for tag in ("unfused", "writeback"):
            for x, y in zip(runs["fused"][0], runs[tag][0]):
                assert np.array_equal(np.asarray(x), np.asarray(y)), tag
        assert runs["fused"][2] == runs["unfused"][2]


def test_identity_constraint_block_aliases_the_parameter_buffer(monkeypatch):
    """`C x == d` over all variables in optimizer order with unit coefficient, C a Parameter the library fills only inside update (UNIFORM
    generator, HOST upload): the coefficient block is the Parameter's own buffer -- no copy launch, the view points at it -- and only the
    offsets -d are assembled.  A permuted variable map, a scaled block and PMB_NO_BLOCK_ALIAS=1 keep the copy; values are identical in all
    cases.  Ref: the block is what moi_interop.jl:64-81 would emit for C * x - d (coefficients C[row, col], constants -d)."""
    import parametron_b200 as pm
    n, k, m, B = 96, 64, 40, 3

    def build(perm=None, scale=None, host=False):
        model = pm.Model(batch=B, instance_offset=2)
        x = model.variables(n)
        A = pm.Parameter(model, shape=(k, n), uniform=(1, 1.0)); b = pm.Parameter(model, shape=(k,), uniform=(2, 1.0))
        if host:
            rng = np.random.default_rng(5)
            Cv = rng.random((B, m, n)); dv = rng.random((B, m))
            Cm = pm.Parameter(model, val=Cv); d = pm.Parameter(model, val=dv)
        else:
            Cm = pm.Parameter(model, shape=(m, n), uniform=(3, 1.0)); d = pm.Parameter(model, shape=(m,), uniform=(4, 2.0))
        r = A * x - b
        pm.objective(model, pm.Minimize, pm.dot(r, r))
        pm.constraint(model, (scale * (Cm * x)) if scale else (Cm * x), "==", d)
        model.initialize(idxmap=perm)
        for _ in range(2):
            model.update()
        v = model.block_view(0)
        out = (int(v.C), int(Cm.device_ptr), model.block_arrays(0), Cm.download() if not host else Cv)
        model.close()
        return out

    ptr, pptr, (Cb, db), Cp = build()
    assert ptr == pptr and np.array_equal(Cb, Cp)                 # the view IS the Parameter's buffer
    for inst in range(B):
        _, _, Co, do = oracle_inputs(n, k, m, 2 + inst, 2, seeds=(1, 2, 3, 4))
        assert np.array_equal(Cb[inst], Co) and np.array_equal(db[inst], -do)
    monkeypatch.setenv("PMB_NO_BLOCK_ALIAS", "1")
    ptr2, pptr2, (Cb2, db2), _ = build()
    assert ptr2 != pptr2 and np.array_equal(Cb2, Cb) and np.array_equal(db2, db)
    monkeypatch.delenv("PMB_NO_BLOCK_ALIAS")
    perm = np.random.default_rng(0).permutation(n) + 1
    ptr3, pptr3, (Cb3, _), _ = build(perm=perm)
    assert ptr3 != pptr3 and np.array_equal(Cb3[:, :, perm - 1], Cb)
    ptr4, pptr4, (Cb4, _), _ = build(scale=2.0)
    assert ptr4 != pptr4 and np.array_equal(Cb4, 2.0 * Cb)
    ptr5, pptr5, (Cb5, _), Cv = build(host=True)
    assert ptr5 == pptr5 and np.array_equal(Cb5, Cv)
