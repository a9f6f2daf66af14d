"""GPU parity of the two "next" rows: the MOI-compatible term-list export (bit-exact order, indices AND
coefficients against the literal oracle's VectorAffineFunction) and the OSQP-style QP pack (patterns == and values 1e-10)."""
import numpy as np
import pytest

from oracle import dense as D
from tests.twin import DeviceSide, OracleSide, assert_close, oracle_outputs

pytestmark = pytest.mark.gpu


def _moi_vector_terms(om):
    return [e[2] for e in om.optimizer.log if e[0] == "constraint"]


@pytest.mark.parametrize("shape", ["dense", "mpc", "box"])
def test_term_list_export_matches_reference_order_and_values(shape):
    rng = np.random.default_rng(hash(shape) % 1000)
    n, m = 6, 4
    hs = {"C": {"v": rng.random((m, n))}, "d": {"v": rng.random(m)}, "H": {"v": rng.random((n, n))}, "v0": {"v": rng.random(n)},
          "dt": {"v": 0.25}, "c": {"v": rng.random(n)}, "lb": {"v": rng.random(n)}}
    perm = [int(i) + 1 for i in rng.permutation(2 * n)]

    def build(s):
        x = s.variables(n)
        u = s.variables(n)
        P = {k_: s.parameter(h) for k_, h in hs.items()}
        if shape == "dense":
            s.constraint(s.mul(P["C"], x), "==", P["d"])
        elif shape == "mpc":  # test/lazyexpression.jl:330-347
            s.constraint(s.sub(s.mul(P["H"], s.sub(x, P["v0"])), s.mul(P["dt"], s.sub(u, P["c"]))), "==", np.zeros(n))
        else:                 # test/model.jl:162-163
            s.constraint(x, ">=", P["lb"])
        return 2 * n

    o, d = OracleSide(), DeviceSide()
    build(o); build(d)
    o.model.initialize(idxmap={i + 1: perm[i] for i in range(2 * n)})
    d.model.initialize(idxmap=perm)
    for step in range(2):
        for h in hs.values():
            h["v"] = rng.random(np.shape(h["v"])) if np.ndim(h["v"]) else float(rng.random())
        oracle_outputs(o.model, 2 * n)
        d.model.update()
        f = _moi_vector_terms(o.model)[0]
        recs, consts = d.model.export_block_terms(0)
        assert [(int(r["output_index"]), float(r["coefficient"]), int(r["variable"])) for r in recs] == \
               [(int(a), float(b), int(c)) for a, b, c in f.terms]          # order, indices and coefficients: bit-exact
        assert_close(consts, [float(z) for z in f.constants], 1e-10, "constants")  # H*v0 is a reduction: 1e-10, not bit-exact


def test_term_list_export_full_size_properties():
    """cfg2's constraint block (m=1024, n=4096: 4.19 M records, 100 MB): row-outer / term-inner, coefficients == C row-major."""
    import parametron_b200 as pm
    n, m = 4096, 1024
    rng = np.random.default_rng(0)
    Cv, dv = rng.random((m, n)), rng.random(m)
    model = pm.Model()
    x = model.variables(n)
    pm.constraint(model, pm.Parameter(model, val=Cv) * x, "==", pm.Parameter(model, val=dv))
    model.initialize()
    model.update()
    recs, consts = model.export_block_terms(0)
    assert recs.shape == (m * n,)
    assert np.array_equal(recs["output_index"], np.repeat(np.arange(1, m + 1), n))
    assert np.array_equal(recs["variable"], np.tile(np.arange(1, n + 1), m))
    assert np.array_equal(recs["coefficient"], Cv.ravel()) and np.array_equal(consts, -dv)


def test_qp_pack_matches_dense_oracle():
    """README Example 1 shape + box constraints: P upper CSC, stacked A CSC (update order: Nonnegatives, Nonpositives, Zeros),
    bounds from the set kinds."""
    import parametron_b200 as pm
    n, k, m = 12, 9, 5
    rng = np.random.default_rng(3)
    A, b, Cm, dv, lo, hi = rng.random((k, n)), rng.random(k), rng.random((m, n)), rng.random(m), -rng.random(n), rng.random(n)
    model = pm.Model()
    x = model.variables(n)
    r = pm.Parameter(model, val=A) * x - pm.Parameter(model, val=b)
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, pm.Parameter(model, val=Cm) * x, "==", pm.Parameter(model, val=dv))   # Zeros        (bucket 5)
    pm.constraint(model, x, "<=", pm.Parameter(model, val=hi))                                # Nonpositives (bucket 4)
    pm.constraint(model, x, ">=", pm.Parameter(model, val=lo))                                # Nonnegatives (bucket 3)
    model.enable_qp_pack()
    perm = rng.permutation(n) + 1
    model.initialize(idxmap=perm)
    model.update()
    qp = model.qp_arrays()
    Qo, qo, c0o = D.least_squares_objective(A, b, np.arange(1, n + 1), n, varmap=perm)
    # P: upper triangle, CSC
    assert np.array_equal(qp["P_p"], np.array([j * (j + 1) // 2 for j in range(n + 1)]))
    assert np.array_equal(qp["P_i"], np.concatenate([np.arange(j + 1) for j in range(n)]))
    assert_close(qp["P_x"], np.concatenate([Qo[:j + 1, j] for j in range(n)]), 1e-10, "P_x")
    assert_close(qp["q"], qo, 1e-10, "q"); assert_close([qp["c0"]], [c0o], 1e-10, "c0")
    # A: stacked in update order  [x - lo >= 0 ; x - hi <= 0 ; Cx - d == 0]
    I = np.zeros((n, n)); I[np.arange(n), perm - 1] = 1.0
    Co, _ = D.affine_block(Cm, dv, np.arange(1, n + 1), n, varmap=perm)
    Astack = np.vstack([I, I, Co])
    M = 2 * n + m
    assert qp["m"] == M and np.array_equal(qp["A_p"], np.arange(n + 1) * M) and np.array_equal(qp["A_i"], np.tile(np.arange(M), n))
    assert np.array_equal(qp["A_x"], Astack.T.ravel())
    assert np.array_equal(qp["l"], np.concatenate([lo, np.full(n, -np.inf), dv]))
    assert np.array_equal(qp["u"], np.concatenate([np.full(n, np.inf), hi, dv]))
    # KKT sanity with the packed data: the equality-constrained least-squares solution satisfies A_eq x = d
    Pfull = np.zeros((n, n))
    for j in range(n):
        Pfull[:j + 1, j] = qp["P_x"][qp["P_p"][j]:qp["P_p"][j + 1]]
    Pfull = Pfull + Pfull.T - np.diag(np.diag(Pfull))
    assert_close(Pfull, Qo, 1e-10, "P symmetric completion")


def test_qp_pack_with_sparse_block():
    import parametron_b200 as pm
    rows, cols, b = 8, 16, 4
    rng = np.random.default_rng(5)
    br, bc = np.array([0, 1, 1], dtype=np.int64), np.array([2, 0, 3], dtype=np.int64)
    blocks = rng.random((3, b, b))
    dv, lo = rng.random(rows), rng.random(cols)
    model = pm.Model()
    x = model.variables(cols)
    pm.constraint(model, pm.Parameter(model, val=blocks, bsr=(rows, cols, b, br, bc)) * x, "==", pm.Parameter(model, val=dv))
    pm.constraint(model, x, ">=", pm.Parameter(model, val=lo))
    model.enable_qp_pack()
    model.initialize()
    model.update()
    qp = model.qp_arrays()
    dense = np.zeros((rows, cols))
    for k_, (i, j) in enumerate(zip(br, bc)):
        dense[i * b:(i + 1) * b, j * b:(j + 1) * b] = blocks[k_]
    Astack = np.vstack([np.eye(cols), dense])       # update order: Nonnegatives (bucket 3) before Zeros (bucket 5)
    got = np.zeros_like(Astack)
    for j in range(cols):
        sl = slice(qp["A_p"][j], qp["A_p"][j + 1])
        assert np.all(np.diff(qp["A_i"][sl]) > 0)   # rows ascending inside a column
        got[qp["A_i"][sl], j] = qp["A_x"][sl]
    assert np.array_equal(got, Astack)
    assert qp["A_p"][-1] == cols * cols + 3 * b * b and qp["P_x"].size == 0  # the Variable block is dense (identity incl. zeros)
    assert np.array_equal(qp["l"], np.concatenate([lo, dv])) and np.array_equal(qp["u"], np.concatenate([np.full(cols, np.inf), dv]))


def test_qp_host_sink_filled_by_the_update_graph():
    """pmb_qp_set_host_sink: the packed QP lands in pinned host buffers as part of pmb_update (two-branch graph with in-graph
    H2D of the pinned HOST Parameters); values equal the explicit device reads, across several re-parameterisations."""
    import parametron_b200 as pm
    n, k, m = 40, 33, 12
    rng = np.random.default_rng(9)
    vals = {"A": rng.random((k, n)), "b": rng.random(k), "C": rng.random((m, n)), "d": rng.random(m)}
    model = pm.Model()
    x = model.variables(n)
    P = {key: pm.Parameter(model, val=v) for key, v in vals.items()}
    r = P["A"] * x - P["b"]
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, P["C"] * x, "<=", P["d"])
    model.enable_qp_pack()
    model.initialize()
    sink = model.qp_host_sink()
    for step in range(3):
        for key in vals:
            P[key].val = rng.random(vals[key].shape)
        model.update()
        model.synchronize()
        qp = model.qp_arrays()
        for key in ("P_x", "q", "A_x", "l", "u"):
            assert np.array_equal(sink[key][0], qp[key]), key
        assert sink["c0"][0] == qp["c0"]
        A, b = P["A"].val, P["b"].val
        Qo = 2 * A.T @ A
        assert_close(sink["P_x"][0], np.concatenate([Qo[:j + 1, j] for j in range(n)]), 1e-10, "P_x")
        assert np.array_equal(sink["A_x"][0], P["C"].val.T.ravel())
        assert np.all(np.isneginf(sink["l"][0])) and np.array_equal(sink["u"][0], P["d"].val)


@pytest.mark.parametrize("n,k", [(1024, 192), (1152, 130)])
def test_upload_compute_overlap_gated_gram(n, k):
    """n >= 1024 (8 tile columns) with a pinned HOST matrix used only by the objective: the matrix is uploaded in column chunks
    while the flag-gated Gram kernel consumes them (pmb_plan.cu build_update_graph).  Results must not depend on the overlap."""
    import parametron_b200 as pm
    m = 16
    rng = np.random.default_rng(n + k)
    vals = {"A": rng.standard_normal((k, n)), "b": rng.standard_normal(k), "C": rng.standard_normal((m, n)), "d": rng.standard_normal(m)}
    model = pm.Model()
    x = model.variables(n)
    P = {key: pm.Parameter(model, val=v) for key, v in vals.items()}
    r = P["A"] * x - P["b"]
    pm.objective(model, pm.Minimize, pm.dot(r, r))
    pm.constraint(model, P["C"] * x, "==", P["d"])
    model.enable_qp_pack()
    model.initialize()
    sink = model.qp_host_sink()
    for step in range(4):
        for key in vals:
            P[key].val = rng.standard_normal(vals[key].shape)
        model.update()
        model.synchronize()
        A, b = P["A"].val, P["b"].val
        Q, q, c0 = model.objective_arrays()
        assert_close(Q, 2 * A.T @ A, 1e-10, "Q"); assert_close(q, -2 * A.T @ b, 1e-10, "q"); assert_close([c0], [b @ b], 1e-10, "c0")
        assert np.array_equal(Q, Q.T)
        iu = np.triu_indices(n)
        assert np.array_equal(sink["P_x"][0], Q.T[iu[1], iu[0]][np.lexsort((iu[0], iu[1]))]) or \
            np.array_equal(sink["P_x"][0], np.concatenate([Q[:j + 1, j] for j in range(n)]))
        assert np.array_equal(sink["A_x"][0], P["C"].val.T.ravel())


def test_term_list_export_of_rows_with_different_layouts():
    """A vcat of blocks of different shapes (test/lazyexpression.jl:188-245): rows 0..m-1 carry n matrix terms, the remaining rows one unit
    term each -- no common per-row template, the export goes through the per-term gather table.  Order, indices and coefficients bit-exact
    against the literal oracle's VectorAffineFunction (moi_interop.jl:64-81)."""
    rng = np.random.default_rng(21)
    n, m = 5, 3
    hs = {"C": {"v": rng.random((m, n))}, "d": {"v": rng.random(m)}, "lb": {"v": rng.random(n)}, "s": {"v": 0.5}}
    perm = [int(i) + 1 for i in rng.permutation(n)]

    def build(s):
        x = s.variables(n)
        P = {k_: s.parameter(h) for k_, h in hs.items()}
        top = s.sub(s.mul(P["C"], x), P["d"])                 # m rows, n terms each
        bottom = s.sub(s.mul(P["s"], x), P["lb"])             # n rows, one term each
        s.constraint(s.vcat(top, bottom), ">=", np.zeros(m + n))

    o, d = OracleSide(), DeviceSide()
    build(o); build(d)
    o.model.initialize(idxmap={i + 1: perm[i] for i in range(n)})
    d.model.initialize(idxmap=perm)
    for step in range(2):
        for h in hs.values():
            h["v"] = rng.random(np.shape(h["v"])) if np.ndim(h["v"]) else float(rng.random())
        oracle_outputs(o.model, n)
        d.model.update()
        f = _moi_vector_terms(o.model)[0]
        recs, consts = d.model.export_block_terms(0)
        assert len(recs) == m * n + n
        assert [(int(r["output_index"]), float(r["coefficient"]), int(r["variable"])) for r in recs] == \
               [(int(a), float(b), int(c)) for a, b, c in f.terms]
        assert_close(consts, [float(z) for z in f.constants], 1e-10, "constants")


@pytest.mark.parametrize("n,k", [(8, 8), (37, 20), (256, 96)])
def test_objective_term_export_matches_the_reference_function(n, k):
    """update!(::MOI.ScalarQuadraticFunction, f, varmap) (moi_interop.jl:45-62) appends one ScalarQuadraticTerm per product, coefficient x 2 when
    both variables coincide, duplicates left to the solver; MOI reads them as 1/2 x'Qx with duplicates summed.  The device exports the SUM:
    one term per unordered optimizer-variable pair (row-outer), the affine terms, the constant.  Checked: structure (indices) exactly, values
    at 1e-10 against the literal oracle's term list summed the way MOI does, and the quadratic form itself at a random point."""
    rng = np.random.default_rng(n + k)
    hs = {"A": {"v": rng.random((k, n))}, "b": {"v": rng.random(k)}}
    perm = [int(i) + 1 for i in rng.permutation(n)]

    def build(s):
        x = s.variables(n)
        A, b = s.parameter(hs["A"]), s.parameter(hs["b"])
        r = s.sub(s.mul(A, x), b)
        s.objective("min", s.dot(r, r))

    small = n <= 64
    d = DeviceSide()
    build(d)
    d.model.initialize(idxmap=perm)
    if small:
        o = OracleSide()
        build(o)
        o.model.initialize(idxmap={i + 1: perm[i] for i in range(n)})
    for step in range(2):
        hs["A"]["v"] = rng.random((k, n)); hs["b"]["v"] = rng.random(k)
        d.model.update()
        quad, aff, const = d.model.export_objective_terms()
        iu = np.triu_indices(n)
        assert np.array_equal(quad["variable_1"], iu[0] + 1) and np.array_equal(quad["variable_2"], iu[1] + 1)   # row-outer upper triangle
        assert np.array_equal(aff["variable"], np.arange(1, n + 1))
        Q, q, c0 = d.model.objective_arrays()
        assert np.array_equal(quad["coefficient"], Q[iu]) and np.array_equal(aff["coefficient"], q) and const == c0
        if small:
            obj, _, _ = oracle_outputs(o.model, n)       # the literal oracle's MOI function, densified with MOI's conventions
            assert_close(Q, obj[0], 1e-10, "Q"); assert_close(q, obj[1], 1e-10, "q"); assert_close([c0], [obj[2]], 1e-10, "c0")
        xs = rng.standard_normal(n)                      # the exported records ARE the function: evaluate them the way MOI does
        val = const + float(aff["coefficient"] @ xs[aff["variable"] - 1])
        v1, v2, cc = quad["variable_1"] - 1, quad["variable_2"] - 1, quad["coefficient"]
        val += float(np.sum(np.where(v1 == v2, 0.5, 1.0) * cc * xs[v1] * xs[v2]))
        pinv = np.argsort(np.asarray(perm) - 1)          # optimizer variable -> model variable
        xm = xs[np.asarray(perm) - 1]                    # model-order point
        ref = float(np.sum((hs["A"]["v"] @ xm - hs["b"]["v"]) ** 2))
        assert abs(val - ref) <= 1e-10 * abs(ref)


def test_objective_term_export_batched_and_affine_objective():
    import parametron_b200 as pm
    n, B = 12, 3
    rng = np.random.default_rng(2)
    model = pm.Model(batch=B)
    x = model.variables(n)
    c = pm.Parameter(model, shape=(n,), device=True)
    pm.objective(model, pm.Minimize, pm.dot(c, x))
    model.initialize()
    cv = rng.random((B, n))
    c.upload(cv)
    model.update()
    quad, aff, const = model.export_objective_terms()
    assert quad is None and aff.shape == (B, n)
    assert np.array_equal(aff["coefficient"], cv) and np.array_equal(aff["variable"], np.tile(np.arange(1, n + 1), (B, 1)))
    assert not const.any()
