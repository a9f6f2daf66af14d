"""GPU parity at the BASELINE.json sizes and through size-independent properties (needs a B200: -m gpu).

The literal term list cannot exist at n=4096 (1.65 TB), so parity is transitive:
device == dense oracle (numpy/BLAS) at full size, and dense oracle == literal oracle at small size (tests/test_oracle_*)
plus device == literal oracle at small size (tests/test_gpu_parity.py).  Tolerance 1e-10 relative, written per assert."""
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
    assert st["affine_algorithmic_bytes"] == 16 * m * n + 16 * m
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
