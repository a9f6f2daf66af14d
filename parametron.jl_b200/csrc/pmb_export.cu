// Host side of the two "next" rows of SURVEY.md section 8(f) that sit right after the hot path:
//   pmb_block_export_terms  -- the reference's MOI.VectorAffineTerm list (moi_interop.jl:64-81) for unmodified MOI consumers
//   pmb_enable_qp_pack / pmb_get_qp -- value-only updates of an OSQP-style (P upper CSC, q, A CSC, l, u), replacing the
//                                      MOI.set sink (moi_interop.jl:134,171)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "pmb_host.h"

namespace pmb {

// ---- QP pack plan (called from pmb_initialize when enabled) -------------------------------------------------------------------
void plan_qp_pack(pmb_ctx* c) {
    const int64_t N = c->nvars, nb = c->batch;
    for (auto& con : c->cons)
        if (con.is_quad) fail(PMB_EUNSUPPORTED, "the OSQP-style QP form has no quadratic constraints");
    // P: upper triangle of the dense Hessian, column j holds rows 0..j
    if (c->obj_kind == PMB_KIND_QUAD && c->Q) {
        c->qp_nnzP = N * (N + 1) / 2;
        std::vector<int64_t> Pp((size_t)N + 1), Pi((size_t)c->qp_nnzP);
        for (int64_t j = 0; j <= N; ++j) Pp[(size_t)j] = j * (j + 1) / 2;
        for (int64_t j = 0; j < N; ++j)
            for (int64_t i = 0; i <= j; ++i) Pi[(size_t)(j * (j + 1) / 2 + i)] = i;
        c->qp_Pp = upload(c, Pp);
        c->qp_Pi = upload(c, Pi);
        c->qp_Px = dalloc<double>(c, nb * c->qp_nnzP, true);
        // If the Hessian is produced by ONE overwriting TMA Gram launch, its epilogue writes the packed upper triangle itself: no pack
        // launch.  (Writing the host sink from the epilogue as well -- posted PCIe writes of 64-byte segments -- was measured and
        // rejected: e2e 5.71 -> 6.93 ms at cfg2; the sink stays a D2H copy node.)
        Step* gram = nullptr;
        int ngram_like = 0;
        for (auto& st : c->steps) {
            if (st.type == ST_GRAM && st.gram.H == c->Q) { gram = &st; ++ngram_like; }
            if ((st.type == ST_BILINEAR && st.bl.H == c->Q) || (st.type == ST_COO && st.coo.H == c->Q) || (st.type == ST_SCATTER_SYM && st.ss.H == c->Q))
                ++ngram_like;
        }
        if (!c->obj_const && ngram_like == 1 && gram && gram->gram.accumulate == 0 && c->use_tma && !getenv("PMB_NO_FUSED_PACK") &&
            gram_tma_applicable(gram->gram.X, gram->gram.ldx, gram->gram.x_bstride, gram->gram.k, gram->gram.n, nb)) {
            gram->gram.packed_dev = c->qp_Px;
            gram->gram.packed_bstride = c->qp_nnzP;
            for (auto& st : c->steps)   // the int8-slice companion of the same contraction writes P_x from its own epilogue
                if (st.type == ST_GRAM_I8_MMA && st.gram.H == c->Q) { st.gram.packed_dev = c->qp_Px; st.gram.packed_bstride = c->qp_nnzP; }
            c->qp_P_fused = true;
        } else {
            Step s;
            s.type = ST_PACK_P; s.cls = KC_PACK; s.batch = nb;
            s.src = c->Q; s.n = N; s.src_bstride = N * N; s.out = c->qp_Px; s.out_bstride = c->qp_nnzP;
            (c->obj_const ? c->init_steps : c->steps).push_back(s);
        }
    } else {
        c->qp_Pp = upload(c, std::vector<int64_t>((size_t)N + 1, 0));
    }
    // A: blocks stacked in update order; per column, the entries of each block in block order (rows ascending inside a block)
    std::vector<int64_t> row_off(c->cons.size(), 0);
    int64_t M = 0;
    for (int ci : c->update_order) { row_off[(size_t)ci] = M; M += c->cons[(size_t)ci].rows; }
    c->qp_m = M;
    // count per column
    std::vector<int64_t> colcount((size_t)N, 0);
    std::vector<std::vector<int32_t>> sp_rows(c->cons.size()), sp_cols(c->cons.size());
    for (int ci : c->update_order) {
        Constraint& con = c->cons[(size_t)ci];
        if (!con.sparse) {
            for (int64_t j = 0; j < N; ++j) colcount[(size_t)j] += con.rows;
        } else {
            sp_rows[(size_t)ci].resize((size_t)con.nnz);
            sp_cols[(size_t)ci].resize((size_t)con.nnz);
            CK(cudaMemcpy(sp_rows[(size_t)ci].data(), con.sp_rows, (size_t)con.nnz * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(sp_cols[(size_t)ci].data(), con.sp_cols, (size_t)con.nnz * 4, cudaMemcpyDeviceToHost));
            for (int64_t e = 0; e < con.nnz; ++e) colcount[(size_t)sp_cols[(size_t)ci][(size_t)e]] += 1;
        }
    }
    std::vector<int64_t> Ap((size_t)N + 1, 0);
    for (int64_t j = 0; j < N; ++j) Ap[(size_t)j + 1] = Ap[(size_t)j] + colcount[(size_t)j];
    c->qp_nnzA = Ap[(size_t)N];
    std::vector<int64_t> Ai((size_t)c->qp_nnzA), cursor(Ap.begin(), Ap.end() - 1);
    std::vector<PackBlockDev> blocks;
    for (int ci : c->update_order) {
        Constraint& con = c->cons[(size_t)ci];
        PackBlockDev B{};
        B.d = con.d; B.d_bstride = con.rows; B.rows = con.rows; B.row_off = row_off[(size_t)ci]; B.set_kind = con.set_kind;
        if (!con.sparse) {
            std::vector<int64_t> colbase((size_t)N);
            for (int64_t j = 0; j < N; ++j) {
                colbase[(size_t)j] = cursor[(size_t)j];
                for (int64_t r = 0; r < con.rows; ++r) Ai[(size_t)(cursor[(size_t)j] + r)] = row_off[(size_t)ci] + r;
                cursor[(size_t)j] += con.rows;
            }
            B.C = con.C; B.c_bstride = con.rows * N; B.sparse = 0; B.colbase = upload(c, colbase);
        } else {
            // stored values sorted by (column, row) inside the block; ties keep value-buffer order
            std::vector<int64_t> order((size_t)con.nnz);
            for (int64_t e = 0; e < con.nnz; ++e) order[(size_t)e] = e;
            const auto& rr = sp_rows[(size_t)ci];
            const auto& cc = sp_cols[(size_t)ci];
            std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
                return cc[(size_t)a] != cc[(size_t)b] ? cc[(size_t)a] < cc[(size_t)b] : rr[(size_t)a] < rr[(size_t)b];
            });
            std::vector<int64_t> pos((size_t)con.nnz);
            for (int64_t q = 0; q < con.nnz; ++q) {
                const int64_t e = order[(size_t)q];
                const int64_t j = cc[(size_t)e];
                pos[(size_t)e] = cursor[(size_t)j];
                Ai[(size_t)cursor[(size_t)j]] = row_off[(size_t)ci] + rr[(size_t)e];
                cursor[(size_t)j] += 1;
            }
            B.C = con.vals; B.c_bstride = con.nnz; B.sparse = 1; B.nnz = con.nnz; B.pos = upload(c, pos);
        }
        blocks.push_back(B);
    }
    c->qp_Ap = upload(c, Ap);
    c->qp_Ai = upload(c, Ai);
    c->qp_Ax = dalloc<double>(c, nb * c->qp_nnzA, true);
    c->qp_l = dalloc<double>(c, nb * M, true);
    c->qp_u = dalloc<double>(c, nb * M, true);
    c->qp_blocks_dev = upload(c, blocks);
    if (!blocks.empty()) {
        bool all_const = true;
        for (auto& con : c->cons) all_const = all_const && con.is_const;
        Step s;
        s.type = ST_PACK_A; s.cls = KC_PACK; s.batch = nb;
        (all_const ? c->init_steps : c->steps).push_back(s);
    }
}

}  // namespace pmb

using namespace pmb;

extern "C" {

int pmb_enable_qp_pack(pmb_ctx* ctx) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "pmb_enable_qp_pack");
    ctx->qp_pack = true;
    API_END(ctx)
}

int pmb_qp_set_host_sink(pmb_ctx* ctx, double* P_x, double* q, double* c0, double* A_x, double* l, double* u) {
    API_BEGIN(ctx)
    if (!ctx->qp_pack) fail(PMB_EINVAL, "pmb_enable_qp_pack was not called");
    ctx->sink_Px = P_x; ctx->sink_q = q; ctx->sink_c0 = c0; ctx->sink_Ax = A_x; ctx->sink_l = l; ctx->sink_u = u;
    if (ctx->initialized) {  // the sizes are only known after pmb_initialize: re-capture the update graph with the new memcpy nodes
        CK(cudaStreamSynchronize(ctx->stream));
        build_update_graph(ctx);
    }
    API_END(ctx)
}

int pmb_get_qp(pmb_ctx* ctx, pmb_qp_view* out) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_get_qp before pmb_initialize");
    if (!ctx->qp_pack) fail(PMB_EINVAL, "pmb_enable_qp_pack was not called");
    if (!out) fail(PMB_EINVAL, "null out");
    out->n = ctx->nvars; out->m = ctx->qp_m; out->nnzP = ctx->qp_nnzP; out->nnzA = ctx->qp_nnzA;
    out->P_x = ctx->qp_Px; out->q = ctx->q; out->c0 = ctx->c0; out->A_x = ctx->qp_Ax; out->l = ctx->qp_l; out->u = ctx->qp_u;
    out->P_p = ctx->qp_Pp; out->P_i = ctx->qp_Pi; out->A_p = ctx->qp_Ap; out->A_i = ctx->qp_Ai;
    API_END(ctx)
}

int pmb_block_export_terms(pmb_ctx* ctx, int32_t update_pos, pmb_vector_affine_term* terms_dev, double* constants_dev) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_block_export_terms before pmb_initialize");
    if (update_pos < 0 || update_pos >= (int32_t)ctx->update_order.size()) fail(PMB_EINVAL, "block position %d out of range", update_pos);
    if (!terms_dev) fail(PMB_EINVAL, "null output buffer");
    Constraint& con = ctx->cons[(size_t)ctx->update_order[(size_t)update_pos]];
    const AffForm& f = con.form;
    if (con.sparse) fail(PMB_EUNSUPPORTED, "sparse blocks have no reference term list (the reference would emit the structural zeros)");
    if (con.is_quad) fail(PMB_EUNSUPPORTED, "quadratic constraints have no affine term list");
    if (!f.order_exact) fail(PMB_EUNSUPPORTED, "the reference's term order of this block is not reproducible from the device lowering");
    if (!con.export_ready) {
        // per-row template when every segment covers all rows (every row has the same layout); otherwise one gather entry per term
        bool uniform_rows = true;
        for (auto& seg : f.order) {
            if (seg.is_unit) { const UnitTerm& u = f.units[(size_t)seg.idx[0]]; uniform_rows = uniform_rows && u.row_off == 0 && u.rows == f.rows; }
            else { const MatTerm& m0 = f.mats[(size_t)seg.idx[0]]; uniform_rows = uniform_rows && m0.row_off == 0 && m0.rows == f.rows; }
        }
        std::vector<AffMatTermDev> mats;
        for (auto& m : f.mats) {
            const Node& sn = ctx->nodes[(size_t)m.mat];
            mats.push_back(AffMatTermDev{sn.dptr, sn.bstride, sn.kind == PMB_KIND_NUM_MATRIX ? sn.rows : 1, m.row_off, m.rows, alpha_dev(ctx, m.a)});
        }
        std::vector<AffUnitTermDev> units;
        for (auto& u : f.units) units.push_back(AffUnitTermDev{alpha_dev(ctx, u.a)});
        con.ex.mats = upload(ctx, mats); con.ex.units = upload(ctx, units);
        con.ex.rows = f.rows;
        if (uniform_rows) {
            std::vector<TermEntryDev> tmpl;
            std::vector<int64_t> unit_vars;
            for (auto& seg : f.order) {
                if (seg.is_unit) {
                    const UnitTerm& u = f.units[(size_t)seg.idx[0]];
                    TermEntryDev e{1, seg.idx[0], (int64_t)unit_vars.size(), 0};
                    for (int64_t r = 0; r < u.rows; ++r) unit_vars.push_back(ctx->varmap[(size_t)(u.vars[(size_t)r] - 1)]);
                    tmpl.push_back(e);
                } else {
                    const MatTerm& m0 = f.mats[(size_t)seg.idx[0]];
                    for (int64_t j = 0; j < m0.cols; ++j)
                        for (int mi : seg.idx) {
                            const MatTerm& m = f.mats[(size_t)mi];
                            tmpl.push_back(TermEntryDev{0, mi, m.src_col_off + j, ctx->varmap[(size_t)(m.vars[(size_t)j] - 1)]});
                        }
                }
            }
            if ((int64_t)tmpl.size() * f.rows != con.nterms) fail(PMB_EINVAL, "internal: template covers %lld terms, expected %lld",
                                                                  (long long)(tmpl.size() * f.rows), (long long)con.nterms);
            con.ex.tmpl = upload(ctx, tmpl); con.ex.T = (int32_t)tmpl.size();
            con.ex.unit_vars = upload(ctx, unit_vars);
        } else {
            // update!(::VectorAffineFunction): row-outer, term-inner, varmap applied (moi_interop.jl:64-81) -- the enumeration of pmb_block_term_index
            std::vector<TermGatherDev> tab;
            tab.reserve((size_t)con.nterms);
            for (int64_t r = 0; r < f.rows; ++r)
                for (auto& seg : f.order) {
                    if (seg.is_unit) {
                        const UnitTerm& u = f.units[(size_t)seg.idx[0]];
                        const int64_t i = r - u.row_off;
                        if (i < 0 || i >= u.rows) continue;
                        tab.push_back(TermGatherDev{1, seg.idx[0], 0, r + 1, ctx->varmap[(size_t)(u.vars[(size_t)i] - 1)]});
                    } else {
                        const MatTerm& m0 = f.mats[(size_t)seg.idx[0]];
                        const int64_t i = r - m0.row_off;
                        if (i < 0 || i >= m0.rows) continue;
                        for (int64_t j = 0; j < m0.cols; ++j)
                            for (int mi : seg.idx) {
                                const MatTerm& m = f.mats[(size_t)mi];
                                const Node& sn = ctx->nodes[(size_t)m.mat];
                                const int64_t ld = sn.kind == PMB_KIND_NUM_MATRIX ? sn.rows : 1;
                                tab.push_back(TermGatherDev{0, mi, i + (m.src_col_off + j) * ld, r + 1, ctx->varmap[(size_t)(m.vars[(size_t)j] - 1)]});
                            }
                    }
                }
            if ((int64_t)tab.size() != con.nterms) fail(PMB_EINVAL, "internal: enumerated %lld terms, expected %lld", (long long)tab.size(), (long long)con.nterms);
            con.ex_gather = upload(ctx, tab);
        }
        con.export_ready = true;
    }
    if (con.ex_gather) {
        launch_export_terms_gather(con.ex_gather, con.nterms, con.ex.mats, con.ex.units, terms_dev, ctx->batch, ctx->stream);
        CK(cudaGetLastError());
        if (constants_dev && con.rows > 0)
            CK(cudaMemcpyAsync(constants_dev, con.d, (size_t)(ctx->batch * con.rows) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        return PMB_OK;
    }
    ExportArgs a = con.ex;
    a.out = terms_dev;
    launch_export_terms(a, ctx->batch, ctx->stream);
    CK(cudaGetLastError());
    if (constants_dev && con.rows > 0)
        CK(cudaMemcpyAsync(constants_dev, con.d, (size_t)(ctx->batch * con.rows) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    API_END(ctx)
}

int pmb_objective_export_terms(pmb_ctx* ctx, pmb_scalar_quadratic_term* quad_dev, pmb_scalar_affine_term* aff_dev, double* constant_dev) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_objective_export_terms before pmb_initialize");
    const int64_t N = ctx->nvars;
    if (quad_dev && !(ctx->obj_kind == PMB_KIND_QUAD && ctx->Q)) fail(PMB_EINVAL, "the objective is affine: it has no quadratic terms");
    if (N > 0 && (quad_dev || aff_dev)) {
        launch_export_objective(quad_dev ? ctx->Q : nullptr, ctx->q, N, quad_dev, aff_dev, ctx->batch, ctx->stream);
        CK(cudaGetLastError());
    }
    if (constant_dev && ctx->c0) CK(cudaMemcpyAsync(constant_dev, ctx->c0, (size_t)ctx->batch * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    API_END(ctx)
}

}  // extern "C"
