// Launch planner and run loop of the C ABI: the device restatement of
//   initialize!(m)  model.jl:117-122 (+ mapindices! :100-107)   -> pmb_initialize: static index tables, buffers, CUDA graph
//   update!(m)      model.jl:132-143                             -> pmb_update: every Parameter once, objective, then the
//                                                                   constraint blocks in bucket order (moi_interop.jl:180-193,236-247)
//   MOI.set(...)    moi_interop.jl:134,171                       -> pmb_get_objective / pmb_get_block device views
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "pmb_host.h"

namespace pmb {

void configure_gram_kernels();  // kernels_gram.cu
void plan_qp_pack(pmb_ctx* c);   // pmb_export.cu

// ---- reachability ----------------------------------------------------------------------------
static void mark(pmb_ctx* c, int id) {
    if (id < 0) return;
    Node& n = c->nodes[id];
    if (n.needed) return;
    n.needed = true;
    for (int i : n.in) mark(c, i);
    for (auto& t : n.lin) { mark(c, t.src); mark(c, t.a.node); }
}
static void mark_aff(pmb_ctx* c, const AffForm& f) {
    for (auto& t : f.mats) { mark(c, t.mat); mark(c, t.a.node); }
    for (auto& t : f.units) mark(c, t.a.node);
    for (auto& t : f.offs) { mark(c, t.vec); mark(c, t.a.node); }
    for (auto& t : f.sps) { mark(c, t.mat); mark(c, t.a.node); }
}
static void mark_quad(pmb_ctx* c, const QuadForm& q) {
    for (auto& g : q.grams) { mark(c, g.a.node); mark_aff(c, g.f); mark_aff(c, g.h); }
    for (auto& b : q.bilins) { mark(c, b.a.node); mark(c, b.mat); }
    for (auto& t : q.coos) mark(c, t.a.node);
    mark_aff(c, q.aff);
}

static GramScratch make_gram_scratch(pmb_ctx* c, bool syrk, int64_t k, int64_t n, int64_t p, int64_t batch) {
    GramScratch sc{};
    sc.plan = plan_gram_split(syrk, k, n, p, batch, c->sm_count);
    if (sc.plan.nslices > 1) {
        sc.counters = dalloc<int32_t>(c, sc.plan.split_tiles, true);
        sc.workspace = dalloc<double>(c, sc.plan.workspace_doubles);
        sc.qworkspace = dalloc<double>(c, sc.plan.split_tiles * sc.plan.nslices * (sc.plan.bm + 1));
    }
    sc.qworkspace_ok = 1;
    return sc;
}

// ---- numeric nodes -> buffers + steps -----------------------------------------------------------
static int64_t src_ld(const Node& n) { return n.kind == PMB_KIND_NUM_MATRIX ? n.rows : 1; }

static void plan_numeric(pmb_ctx* c, int id) {
    Node& n = c->nodes[id];
    n.numel = numel_of(n);
    std::vector<Step>& dst = n.is_const ? c->init_steps : c->steps;
    const int64_t nb = n.is_const ? 1 : c->batch;
    if (n.nop == NOP_CONST) {
        n.dptr = upload(c, n.cvals);
        n.bstride = 0; n.owned = true;
        return;
    }
    if (n.nop == NOP_PARAM) {
        n.bstride = n.numel;
        if (n.pmode == PMB_PARAM_DEVICE && n.ext) {
            n.dptr = const_cast<double*>(n.ext);
        } else {
            n.dptr = dalloc<double>(c, c->batch * n.numel, true);
            n.owned = true;
        }
        if (n.pmode == PMB_PARAM_UNIFORM && n.numel > 0) { c->fill_jobs.push_back(FillJob{n.dptr, n.numel, n.seed, n.scale, 0, 0}); c->fill_job_nodes.push_back(id); }
        return;
    }
    n.dptr = dalloc<double>(c, nb * n.numel);
    n.bstride = n.is_const ? 0 : n.numel;
    n.owned = true;
    Step s;
    s.batch = nb;
    switch (n.nop) {
        case NOP_LINCOMB: {
            std::vector<LinCombTermDev> lt;
            for (auto& t : n.lin) {
                const Node& sn = c->nodes[t.src];
                lt.push_back(LinCombTermDev{sn.dptr, sn.bstride, t.src_off, alpha_dev(c, t.a)});
            }
            s.type = ST_LINCOMB; s.cls = KC_MISC;
            s.out = n.dptr; s.out_bstride = n.bstride; s.n = n.numel;
            s.lterms = upload(c, lt); s.nl = (int)lt.size(); s.cst = nullptr;
            break;
        }
        case NOP_TRANSPOSE: {
            const Node& a = c->nodes[n.in[0]];
            s.type = ST_TRANSPOSE; s.cls = KC_MISC;
            s.out = n.dptr; s.out_bstride = n.bstride; s.src = a.dptr; s.src_bstride = a.bstride; s.n = a.rows; s.n2 = a.cols;
            break;
        }
        case NOP_GEMV: {
            const Node& a = c->nodes[n.in[0]];
            const Node& v = c->nodes[n.in[1]];
            s.type = ST_GEMV; s.cls = KC_REDUCE;
            GemvArgs g{};
            g.A = a.dptr + n.gemv_col_off * a.rows; g.lda = a.rows; g.a_bstride = a.bstride;
            g.rows = a.rows; g.cols = n.gemv_cols; g.trans = n.gemv_trans;
            g.v = v.dptr; g.v_bstride = v.bstride; g.v_off = n.gemv_v_off;
            g.y = n.dptr; g.y_bstride = n.bstride; g.map = nullptr; g.scale = AlphaDev{1.0, nullptr, 0}; g.accumulate = 0;
            s.gemv = g;
            break;
        }
        case NOP_DOT: {
            const Node& x = c->nodes[n.in[0]];
            const Node& y = c->nodes[n.in[1]];
            s.type = ST_DOT; s.cls = KC_REDUCE;
            DotArgs d{};
            d.x = x.dptr; d.x_bstride = x.bstride; d.x_off = n.dot_xoff;
            d.y = y.dptr; d.y_bstride = y.bstride; d.y_off = n.dot_yoff;
            d.n = n.dot_n; d.out = n.dptr; d.out_bstride = n.bstride; d.scale = AlphaDev{1.0, nullptr, 0}; d.accumulate = 0;
            s.dot = d;
            break;
        }
        case NOP_SPROD: {
            s.type = ST_SPROD; s.cls = KC_MISC;
            s.out = n.dptr; s.out_bstride = n.bstride; s.scale = n.sprod_c;
            const Node& a = c->nodes[n.in[0]];
            s.s0 = a.dptr; s.s0_bs = a.bstride;
            if (n.in.size() > 1) { const Node& b = c->nodes[n.in[1]]; s.s1 = b.dptr; s.s1_bs = b.bstride; }
            break;
        }
        case NOP_GEMM_TN: {
            const Node& x = c->nodes[n.in[0]];  // X (k_total x n), rows [gm_xoff, gm_xoff + gm_k)
            const Node& y = c->nodes[n.in[1]];  // Y (gm_k x *), columns [gm_yoff, gm_yoff + p)
            s.type = ST_GEMM; s.cls = KC_GEMM;
            GemmTNArgs g{};
            g.X = x.dptr + n.gm_xoff; g.ldx = x.rows; g.x_bstride = x.bstride;
            g.Y = y.dptr + n.gm_yoff * y.rows; g.ldy = y.rows; g.y_bstride = y.bstride;
            g.k = n.gm_k; g.n = n.rows; g.p = n.cols;
            g.M = n.dptr; g.ldm = n.rows; g.m_bstride = n.bstride;
            s.gemm = g;
            s.gscratch = make_gram_scratch(c, false, g.k, g.n, g.p, nb);
            break;
        }
        default: fail(PMB_EINVAL, "internal: numeric node %d has no plan", id);
    }
    dst.push_back(s);
}

// ---- AffForm -> AffBlockDev ------------------------------------------------------------------
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// compact == true: output columns are the ranks of the touched optimizer columns (scratch operand of a Gram);
// *touched receives the touched optimizer columns (ascending).
static AffBlockDev lower_aff(pmb_ctx* c, const AffForm& f, double* C, int64_t C_bstride, double* d, int64_t d_bstride, bool compact,
                             std::vector<int32_t>* touched) {
    std::map<int32_t, std::vector<AffColEntry>> cols;
    for (size_t t = 0; t < f.mats.size(); ++t) {
        const MatTerm& m = f.mats[t];
        for (int64_t j = 0; j < m.cols; ++j) cols[optcol(c, m.vars[(size_t)j])].push_back(AffColEntry{(int32_t)t, (int32_t)(m.src_col_off + j)});
    }
    for (size_t u = 0; u < f.units.size(); ++u) {
        const UnitTerm& ut = f.units[u];
        for (int64_t r = 0; r < ut.rows; ++r) cols[optcol(c, ut.vars[(size_t)r])].push_back(AffColEntry{~(int32_t)u, (int32_t)(ut.row_off + r)});
    }
    std::vector<int32_t> col_out, col_ptr{0}, col_src;
    std::vector<AffColEntry> entries;
    bool simple = f.mats.size() == 1 && f.units.empty() && f.mats[0].row_off == 0 && f.mats[0].rows == f.rows;
    int32_t rank = 0;
    for (auto& kv : cols) {
        col_out.push_back(compact ? rank : kv.first);
        if (touched) touched->push_back(kv.first);
        for (auto& e : kv.second) entries.push_back(e);
        col_ptr.push_back((int32_t)entries.size());
        if (kv.second.size() != 1) simple = false;
        col_src.push_back(kv.second[0].src);
        ++rank;
    }
    std::vector<AffMatTermDev> mats;
    for (auto& m : f.mats) {
        const Node& sn = c->nodes[m.mat];
        mats.push_back(AffMatTermDev{sn.dptr, sn.bstride, src_ld(sn), m.row_off, m.rows, alpha_dev(c, m.a)});
    }
    std::vector<AffUnitTermDev> units;
    for (auto& u : f.units) units.push_back(AffUnitTermDev{alpha_dev(c, u.a)});
    std::vector<AffOffTermDev> offs;
    for (auto& o : f.offs) {
        const Node& sn = c->nodes[o.vec];
        offs.push_back(AffOffTermDev{sn.dptr, sn.bstride, o.src_off, o.row_off, o.rows, alpha_dev(c, o.a)});
    }
    AffBlockDev b{};
    b.C = C; b.d = d; b.C_bstride = C_bstride; b.d_bstride = d_bstride; b.rows = f.rows; b.d_rows = f.rows;
    b.ntouched = (int32_t)col_out.size();
    b.col_out = upload(c, col_out);
    b.col_ptr = upload(c, col_ptr);
    b.entries = upload(c, entries);
    b.mats = upload(c, mats);
    b.units = upload(c, units);
    b.noffs = (int32_t)offs.size();
    b.offs = upload(c, offs);
    bool anyc = false;
    for (double v : f.coff) anyc = anyc || v != 0.0;
    b.d_const = anyc ? upload(c, f.coff) : nullptr;
    b.simple = simple && b.ntouched > 0 ? 1 : 0;
    b.col_src = upload(c, col_src);
    b.vec2 = 0;
    if (b.simple) {
        const AffMatTermDev& m = mats[0];
        b.vec2 = (f.rows % 2 == 0 && m.ld % 2 == 0 && m.src_bstride % 2 == 0 && C_bstride % 2 == 0 && aligned16(m.src) && aligned16(C)) ? 1 : 0;
    }
    if (b.simple && b.vec2 && c->use_tma) {
        // bulk-copy plan: merge adjacent columns into contiguous runs, cut runs into <= 16 KB pieces, pack pieces into chunks
        const AffMatTermDev& m = mats[0];
        struct Run { int64_t src, dst, len; };
        std::vector<Run> runs;
        for (size_t tc = 0; tc < col_out.size(); ++tc) {
            Run r{(int64_t)col_src[tc] * m.ld, (int64_t)col_out[tc] * f.rows, f.rows};
            if (!runs.empty() && runs.back().src + runs.back().len == r.src && runs.back().dst + runs.back().len == r.dst) runs.back().len += r.len;
            else runs.push_back(r);
        }
        const int64_t piece = kAffStageBytes / 8;
        std::vector<AffSeg> segs;
        std::vector<int32_t> chunk_first{0};
        int32_t fill = 0;
        for (auto& r : runs)
            for (int64_t o = 0; o < r.len; o += piece) {
                int32_t bytes = (int32_t)(std::min(piece, r.len - o) * 8);
                if (fill + bytes > kAffStageBytes) { chunk_first.push_back((int32_t)segs.size()); fill = 0; }
                segs.push_back(AffSeg{r.src + o, r.dst + o, bytes, fill});
                fill += bytes;
            }
        chunk_first.push_back((int32_t)segs.size());
        b.segs = upload(c, segs);
        b.chunk_first = upload(c, chunk_first);
        b.nchunks = (int32_t)chunk_first.size() - 1;
        b.tma = (m.alpha.ptr == nullptr && m.alpha.c == 1.0) ? 1 : 2;
        // regular plan?
        bool uni = (size_t)b.nchunks == segs.size() && !segs.empty();
        if (uni && segs.size() > 1) {
            const int64_t ds = segs[1].src_off - segs[0].src_off, dd = segs[1].dst_off - segs[0].dst_off;
            for (size_t i = 1; i < segs.size() && uni; ++i) {
                uni = segs[i].src_off - segs[i - 1].src_off == ds && segs[i].dst_off - segs[i - 1].dst_off == dd;
                if (i + 1 < segs.size()) uni = uni && segs[i].bytes == segs[0].bytes;
            }
            b.u_src_stride = ds; b.u_dst_stride = dd;
        }
        if (uni) {
            b.uniform = 1;
            b.u_src0 = segs[0].src_off; b.u_dst0 = segs[0].dst_off;
            b.u_bytes = segs[0].bytes; b.u_last_bytes = segs.back().bytes;
        }
    }
    return b;
}

static int64_t aff_nterms(const AffForm& f) {
    int64_t n = 0;
    for (auto& m : f.mats) n += m.rows * m.cols;
    for (auto& u : f.units) n += u.rows;
    return n;
}

static Step affine_step(const AffBlockDev& b, int64_t batch) {
    Step s;
    s.type = ST_AFFINE; s.cls = KC_AFFINE; s.batch = batch; s.aff = b;
    return s;
}

// ---- sparse constraint block (one BSR source): values-only update over a static (row, optimizer column) index ---------
static void plan_sparse_block(pmb_ctx* c, Constraint& con) {
    const SpTerm& sp = con.form.sps[0];
    const Node& mn = c->nodes[sp.mat];
    const int64_t b = mn.bsr_b, nb = mn.bsr_nblocks, nnz = nb * b * b;
    std::vector<int32_t> rows((size_t)nnz), cols((size_t)nnz);
    for (int64_t blk = 0; blk < nb; ++blk)
        for (int64_t j = 0; j < b; ++j)
            for (int64_t i = 0; i < b; ++i) {
                const size_t e = (size_t)((blk * b + j) * b + i);  // block-major, column-major inside a block
                rows[e] = (int32_t)(mn.bsr_brow[(size_t)blk] * b + i);
                cols[e] = optcol(c, sp.vars[(size_t)(mn.bsr_bcol[(size_t)blk] * b + j)]);
            }
    con.nnz = nnz;
    con.vals = dalloc<double>(c, c->batch * nnz, true);
    con.d = dalloc<double>(c, c->batch * con.rows, true);
    con.sp_rows = upload(c, rows);
    con.sp_cols = upload(c, cols);
    con.nterms = con.rows * (int64_t)sp.vars.size();  // what the reference would emit: every entry, zeros included
    std::vector<Step>& st = con.is_const ? c->init_steps : c->steps;
    // one launch: the offsets d (rows of the block) + one contiguous (scaled) copy of the value buffer, expressed as a
    // one-column "simple" block of nnz rows
    AffForm only_off;
    only_off.rows = con.rows; only_off.offs = con.form.offs; only_off.coff = con.form.coff;
    AffBlockDev v = lower_aff(c, only_off, nullptr, 0, con.d, con.rows, true, nullptr);
    if (nnz > 0) {
        v.C = con.vals; v.C_bstride = nnz; v.rows = nnz; v.ntouched = 1;
        v.col_out = upload(c, std::vector<int32_t>{0});
        v.col_src = upload(c, std::vector<int32_t>{0});
        v.col_ptr = upload(c, std::vector<int32_t>{0, 1});
        v.entries = upload(c, std::vector<AffColEntry>{AffColEntry{0, 0}});
        v.mats = upload(c, std::vector<AffMatTermDev>{AffMatTermDev{mn.dptr, mn.bstride, nnz, 0, nnz, alpha_dev(c, sp.a)}});
        v.simple = 1;
        v.vec2 = (nnz % 2 == 0 && mn.bstride % 2 == 0 && aligned16(mn.dptr) && aligned16(con.vals)) ? 1 : 0;
        if (v.vec2 && c->use_tma) {
            std::vector<AffSeg> segs;
            for (int64_t o = 0; o < nnz; o += kAffStageBytes / 8)
                segs.push_back(AffSeg{o, o, (int32_t)(std::min<int64_t>(kAffStageBytes / 8, nnz - o) * 8), 0});
            std::vector<int32_t> cf(segs.size() + 1);
            for (size_t q = 0; q <= segs.size(); ++q) cf[q] = (int32_t)q;
            v.segs = upload(c, segs); v.chunk_first = upload(c, cf); v.nchunks = (int32_t)segs.size();
            v.tma = (sp.a.node < 0 && sp.a.c == 1.0) ? 1 : 2;
            v.uniform = 1; v.u_src0 = 0; v.u_dst0 = 0; v.u_src_stride = v.u_dst_stride = kAffStageBytes / 8;
            v.u_bytes = segs[0].bytes; v.u_last_bytes = segs.back().bytes;
        }
    }
    st.push_back(affine_step(v, c->batch));
    if (!con.is_const) c->affine_bytes += 16 * nnz + 16 * con.rows;
}

// ---- objective ---------------------------------------------------------------------------------
struct DenseSide {
    const double* X = nullptr;
    int64_t ldx = 0, xbs = 0, ncols = 0, k = 0;
    const int32_t* map = nullptr;
    const double* g = nullptr;  // offsets (k), null when identically zero
    int64_t gbs = 0;
    double cx = 1.0;
    double cg = 1.0;            // the offsets are cg * g[...]
    int whole_node = -1;        // X is exactly this (whole) matrix node, read in place
    bool ident = false;         // map[i] == i for every column: kernels take the null (identity) map and their vectorised store paths
};

static DenseSide densify(pmb_ctx* c, const AffForm& f, std::vector<Step>& steps, int64_t nb) {
    DenseSide s;
    s.k = f.rows;
    const bool has_off = !f.offs.empty() || std::any_of(f.coff.begin(), f.coff.end(), [](double v) { return v != 0.0; });
    bool direct = f.mats.size() == 1 && f.units.empty() && f.mats[0].a.node < 0 && f.mats[0].row_off == 0 && f.mats[0].rows == f.rows &&
                  c->nodes[f.mats[0].mat].kind == PMB_KIND_NUM_MATRIX;
    std::vector<int32_t> map;
    if (direct) {
        for (int64_t v : f.mats[0].vars) map.push_back(optcol(c, v));
        std::vector<int32_t> sorted = map;
        std::sort(sorted.begin(), sorted.end());
        direct = std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end();
    }
    // offsets that are one scaled numeric vector are used in place (no launch): g = cg * v
    const bool g_direct = has_off && f.offs.size() == 1 && f.offs[0].a.node < 0 && f.offs[0].row_off == 0 && f.offs[0].rows == f.rows &&
                          std::all_of(f.coff.begin(), f.coff.end(), [](double v) { return v == 0.0; });
    double* g = (has_off && !(direct && g_direct)) ? dalloc<double>(c, nb * f.rows) : nullptr;
    if (direct) {
        const MatTerm& m = f.mats[0];
        const Node& sn = c->nodes[m.mat];
        s.X = sn.dptr + m.src_col_off * sn.rows; s.ldx = sn.rows; s.xbs = sn.bstride; s.ncols = m.cols; s.cx = m.a.c;
        if (m.src_col_off == 0 && m.cols == sn.cols && sn.bsr_b == 0) s.whole_node = m.mat;
        s.map = upload(c, map);
        s.ident = true;
        for (size_t i = 0; i < map.size(); ++i) s.ident = s.ident && map[i] == (int32_t)i;
        if (has_off && g_direct) {
            const Node& gn = c->nodes[f.offs[0].vec];
            s.g = gn.dptr + f.offs[0].src_off; s.gbs = gn.bstride; s.cg = f.offs[0].a.c;
            return s;
        }
        if (has_off) {
            AffForm only_off;
            only_off.rows = f.rows; only_off.offs = f.offs; only_off.coff = f.coff;
            Step os = affine_step(lower_aff(c, only_off, nullptr, 0, g, f.rows, true, nullptr), nb);
            os.cls = KC_MISC;  // offsets only: not a coefficient-assembly launch
            steps.push_back(os);
        }
    } else {
        // materialise the operand as a compact dense block (the affine-assembly kernel sums duplicate variables per column)
        std::vector<int32_t> touched;
        int64_t nt = 0;
        {
            std::map<int32_t, int> seen;
            for (auto& m : f.mats) for (int64_t v : m.vars) seen[optcol(c, v)] = 1;
            for (auto& u : f.units) for (int64_t v : u.vars) seen[optcol(c, v)] = 1;
            nt = (int64_t)seen.size();
        }
        double* X = dalloc<double>(c, nb * f.rows * nt, true);
        double* gg = g ? g : dalloc<double>(c, nb * f.rows);
        steps.push_back(affine_step(lower_aff(c, f, X, f.rows * nt, gg, f.rows, true, &touched), nb));
        s.X = X; s.ldx = f.rows; s.xbs = f.rows * nt; s.ncols = nt; s.cx = 1.0;
        s.map = upload(c, touched);
        s.ident = true;
        for (size_t i = 0; i < touched.size(); ++i) s.ident = s.ident && touched[i] == (int32_t)i;
    }
    s.g = g; s.gbs = f.rows;
    return s;
}

static AlphaDev scaled(AlphaDev a, double k) { a.c *= k; return a; }

// One scalar quadratic function (the objective, or a ScalarQuadraticFunction-in-set constraint) -> (Q, q, c0) buffers + launches.
// The device restatement of update!(::MOI.ScalarQuadraticFunction, f, varmap), moi_interop.jl:45-62.
struct QuadTarget {
    double* Q = nullptr;
    double* q = nullptr;
    double* c0 = nullptr;
};
static int count_refs(pmb_ctx* c, int id);

static QuadTarget plan_quadratic(pmb_ctx* c, QuadForm qf, bool is_const, bool want_Q, bool is_objective = false) {
    const int64_t N = c->nvars, nb = c->batch;
    QuadTarget T;
    double* qc = dalloc<double>(c, nb * (N + 1), true);
    T.q = qc;
    T.c0 = qc + nb * N;
    std::vector<Step>& st = is_const ? c->init_steps : c->steps;
    const bool has_quad = !qf.grams.empty() || !qf.bilins.empty() || !qf.coos.empty();
    if (want_Q) T.Q = dalloc<double>(c, nb * N * N, true);
    // (q, c0) <- 0, then the affine part overwrites its touched columns, then the quadratic terms accumulate
    if (is_const) {
        Step z;
        z.type = ST_MEMSET; z.cls = KC_MISC; z.out = qc; z.n = nb * (N + 1);
        st.push_back(z);
    } else {
        c->fill_jobs.push_back(FillJob{qc, nb * (N + 1), 0, 0.0, 1, 0});  // zeroed by the fused prologue launch
        c->fill_job_nodes.push_back(-1);
    }
    if (qf.aff.rows == 0) qf.aff.rows = 1;
    if (!qf.aff.mats.empty() || !qf.aff.units.empty() || !qf.aff.offs.empty() || !qf.aff.coff.empty())
        st.push_back(affine_step(lower_aff(c, qf.aff, T.q, N, T.c0, 1, false, nullptr), nb));
    if (!has_quad) return T;
    const bool overwrite = qf.grams.size() == 1 && qf.grams[0].same && qf.bilins.empty() && qf.coos.empty();
    if (!overwrite) {
        Step zq;
        zq.type = ST_MEMSET; zq.cls = KC_MISC; zq.out = T.Q; zq.n = nb * N * N;
        st.push_back(zq);
    }
    for (auto& gt : qf.grams) {
        const AlphaDev a = alpha_dev(c, gt.a);
        if (gt.f.rows != gt.h.rows) fail(PMB_EDIM, "DimensionMismatch in dot");
        DenseSide F = densify(c, gt.f, st, nb);
        if (gt.same) {
            // H[map, map] (+)= 2 a cx^2 X'X ;  q[map] += 2 a cx X'g ;  c0 += a g.g
            Step s;
            s.type = ST_GRAM; s.cls = KC_GRAM; s.batch = nb;
            GramArgs ga{};
            ga.X = F.X; ga.ldx = F.ldx; ga.x_bstride = F.xbs; ga.k = F.k; ga.n = F.ncols;
            ga.H = T.Q; ga.ldh = N; ga.h_bstride = N * N; ga.map = F.ident ? nullptr : F.map;
            ga.scale = scaled(a, 2.0 * F.cx * F.cx); ga.accumulate = overwrite ? 0 : 1;
            // certified int8-slice path on tcgen05 (kernels_gram_i8.cu): large contractions only, fp64 kernel keeps the uncertified instances
            GramI8Scratch i8{};
            // generator fusion (both int8 routes): the operand is a UNIFORM Parameter nothing else reads -> its Philox fill moves from the prologue
            // launch into the pass that slices it (same bits); the Parameter's buffer is then written on demand (pmb_param_device_ptr), or by the
            // two-kernel route's slice pass on every update with PMB_GEN_MATERIALIZE=1
            auto try_generator_fusion = [&](bool can_write_back) {
                if (F.whole_node < 0 || getenv("PMB_NO_GEN_FUSE")) return;
                const bool write_back = getenv("PMB_GEN_MATERIALIZE") != nullptr;
                if (write_back && !can_write_back) return;
                const Node& pn = c->nodes[(size_t)F.whole_node];
                if (!(pn.nop == NOP_PARAM && pn.pmode == PMB_PARAM_UNIFORM && pn.owned && F.k == pn.rows && F.ncols == pn.cols && F.xbs == pn.numel &&
                      count_refs(c, F.whole_node) == 1)) return;
                for (size_t j = 0; j < c->fill_jobs.size() && j < c->fill_job_nodes.size(); ++j)
                    if (c->fill_job_nodes[j] == F.whole_node) {
                        FillJob lazy = c->fill_jobs[j];
                        c->fill_jobs.erase(c->fill_jobs.begin() + (long)j); c->fill_job_nodes.erase(c->fill_job_nodes.begin() + (long)j);
                        i8.gen_counter = c->d_counter; i8.gen_seed = pn.seed; i8.gen_scale = pn.scale; i8.gen_inst_offset = c->inst_offset;
                        i8.gen_writeback = write_back ? 1 : 0;
                        if (!write_back) {   // the buffer is filled on demand: keep the job as a one-element list for pmb_param_device_ptr
                            const int64_t cap = (int64_t)c->sm_count * 16;
                            lazy.nblocks = (int32_t)std::max<int64_t>(1, std::min<int64_t>(((lazy.numel + 1) / 2 + 255) / 256, std::max<int64_t>(1, cap / std::max<int64_t>(1, std::min<int64_t>(nb, 64)))));
                            Node& wn = c->nodes[(size_t)F.whole_node];
                            wn.lazy_fill = upload(c, std::vector<FillJob>{lazy}); wn.lazy_fill_blocks = lazy.nblocks;
                        }
                        return;
                    }
            };
            // fused small-instance route (kernels_gram_i8_fused.cu): batches of instances with n, k <= 256 -- slicing, digit-plane products and
            // epilogue of one instance per CTA pair in one launch, no digit planes in HBM; q and c0 ride along
            GramArgs fz = ga;
            if (F.g) {
                fz.gvec = F.g; fz.g_bstride = F.gbs; fz.qout = T.q; fz.q_bstride = N; fz.c0out = T.c0;
                fz.qscale = scaled(a, 2.0 * F.cx * F.cg); fz.cscale = scaled(a, F.cg * F.cg);
            }
            if ((c->gram_mode == 1 || c->gram_mode == 5) && c->use_tma && !getenv("PMB_NO_FUSED_I8") && gram_tma_applicable(F.X, F.ldx, F.xbs, F.k, F.ncols, nb) &&
                gram_i8_fused_applicable(fz, nb) && (c->gram_mode == 5 || gram_i8_fused_pays_off(F.k, F.ncols, nb, c->sm_count))) {
                i8.stats = dalloc<unsigned long long>(c, nb * kGramI8StatsStride + 2 + 1024, true);
                i8.tol = c->gram_tol;
                i8.force_ns = c->gram_mode == 5 ? 5 : 0;
                if (!c->fused_diag && cudaHostAlloc((void**)&c->fused_diag, 32 * sizeof(unsigned long long), cudaHostAllocMapped) == cudaSuccess)
                    { for (int w = 0; w < 32; ++w) c->fused_diag[w] = 0ull; c->pinned.push_back(c->fused_diag); }
                (void)cudaGetLastError();
                i8.diag = c->fused_diag;
                i8.fused = 1;
                if (gram_i8_fused_can_generate(fz)) try_generator_fusion(false);
                ga.cert_stats = i8.stats; ga.cert_tol = i8.tol; ga.cert_force = i8.force_ns != 0; ga.cert_none = i8.stats + nb * kGramI8StatsStride;
                fz.cert_stats = ga.cert_stats; fz.cert_tol = ga.cert_tol; fz.cert_force = ga.cert_force; fz.cert_none = ga.cert_none;
            }
            if (!i8.fused && c->gram_mode != 0 && c->use_tma && gram_tma_applicable(F.X, F.ldx, F.xbs, F.k, F.ncols, nb) &&
                (c->gram_mode >= 5 || gram_i8_pays_off(F.k, F.ncols, nb, c->sm_count))) {
                i8.plan = plan_gram_i8(F.k, F.ncols, nb);
                if (i8.plan.ok) {
                    i8.slices = dalloc<int8_t>(c, i8.plan.slice_bytes);
                    i8.sigma = dalloc<double>(c, nb * F.ncols);
                    i8.stats = dalloc<unsigned long long>(c, nb * kGramI8StatsStride, true);
                    i8.tol = c->gram_tol;
                    i8.force_ns = c->gram_mode >= 5 ? c->gram_mode : 0;
                    ga.cert_stats = i8.stats; ga.cert_tol = i8.tol; ga.cert_force = i8.force_ns != 0;
                    {
                        GramArgs probe = ga; probe.gvec = F.g; probe.g_bstride = F.gbs;
                        if (gram_i8_slice_can_generate(probe, i8.plan)) try_generator_fusion(true);
                    }
                }
            }
            // upload/compute overlap: the operand is a pinned HOST Parameter nothing else reads -> gate the kernel on chunk arrival
            if (!i8.slices && is_objective && overwrite && c->h2d_in_graph && nb == 1 && F.whole_node >= 0 && c->pipe_param < 0 && c->use_tma &&
                getenv("PMB_PIPE_GATED") && !getenv("PMB_NO_PIPE")) {
                // OPT-IN (PMB_PIPE_GATED=1): the persistent kernel polls flags written by copy nodes of a sibling graph branch, which only works
                // while the runtime runs the two concurrently (it does on this driver; a profiler, sanitizer or MPS time-slicing may serialise
                // them -- the poll is bounded and traps).  Default: the operand is uploaded first (the int8 route has its own event-gated pipeline)
                const Node& pn = c->nodes[(size_t)F.whole_node];
                const int64_t tile_cols = (F.ncols + 127) / 128;
                if (pn.nop == NOP_PARAM && pn.pmode == PMB_PARAM_HOST && tile_cols >= 8 && count_refs(c, F.whole_node) == 1 &&
                    gram_tma_applicable(F.X, F.ldx, F.xbs, F.k, F.ncols, nb)) {
                    c->pipe_param = F.whole_node;
                    c->pipe_gate_cols = (int32_t)(((tile_cols + 7) / 8) * 128);       // 8 chunks of whole 128-column panels
                    c->pipe_chunks = (int32_t)((F.ncols + c->pipe_gate_cols - 1) / c->pipe_gate_cols);
                    c->pipe_gate = dalloc<int32_t>(c, c->pipe_chunks, true);
                    void* hp = nullptr;
                    CK(cudaHostAlloc(&hp, sizeof(int32_t) * (size_t)c->pipe_chunks, cudaHostAllocDefault));
                    c->pinned.push_back(hp);
                    c->pipe_ones = (int32_t*)hp;
                    for (int q = 0; q < c->pipe_chunks; ++q) c->pipe_ones[q] = 1;
                    ga.gate = c->pipe_gate; ga.gate_cols = c->pipe_gate_cols;
                }
            }
            s.gram = ga;
            s.gscratch = make_gram_scratch(c, true, ga.k, ga.n, ga.n, nb);
            bool pipe8 = false;
            if (i8.slices && is_objective && overwrite && c->h2d_in_graph && nb == 1 && F.whole_node >= 0 && c->pipe_param < 0 && !getenv("PMB_NO_PIPE")) {
                const Node& pn = c->nodes[(size_t)F.whole_node];
                bool generators = false;   // a replay of the update must see the same Parameter values: host / manual Parameters only
                for (auto& nd : c->nodes) generators = generators || (nd.nop == NOP_PARAM && nd.pmode == PMB_PARAM_UNIFORM);
                if (pn.nop == NOP_PARAM && pn.pmode == PMB_PARAM_HOST && F.ncols >= 1920 && F.ldx == pn.rows && count_refs(c, F.whole_node) == 1 && !generators) {
                    pipe8 = true;
                    c->pipe_param = F.whole_node;
                    c->pipe8 = true;
                    // ~8 chunks of equal width (multiples of 480 = lcm(96, 80) columns keep both tile grids aligned), then a finer tail:
                    // what is exposed after the last byte of the upload is the last chunk's slice + Gram range + P_x download
                    const int64_t step = 480 * std::max<int64_t>(1, (F.ncols / 8 + 240) / 480);
                    c->pipe8_bounds.clear();
                    int64_t b0 = 0;
                    for (; F.ncols - b0 > 2 * step; b0 += step) c->pipe8_bounds.push_back(b0);
                    if (const char* e = getenv("PMB_PIPE8_TAIL")) {
                        const int parts = std::max(1, atoi(e));
                        const int64_t rem = F.ncols - b0;
                        for (int q = 0; q < parts; ++q) c->pipe8_bounds.push_back(b0 + (rem * q / parts) / 16 * 16);
                    } else {
                        const int64_t rem = F.ncols - b0;
                        for (int q = 0; q < 4; ++q) c->pipe8_bounds.push_back(b0 + (rem * q / 4) / 16 * 16);
                    }
                    c->pipe8_bounds.push_back(F.ncols);
                }
            }
            s.pipe8 = pipe8;
            if (i8.fused) {   // one fused launch (H of the certified instances, q and c0 of all), then the fp64 launch for what it left
                Step fs;
                fs.type = ST_GRAM_I8_FUSED; fs.cls = KC_GRAM_I8; fs.batch = nb; fs.gram = fz; fs.i8 = i8;
                st.push_back(fs);
                s.i8 = i8;
            }
            if (i8.slices) {  // slice pass, then the 5- and 6-plane tcgen05 launches; the fp64 launch after them skips what they certified
                Step sl;
                sl.type = ST_GRAM_I8_SLICE; sl.cls = KC_GRAM_SLICE; sl.batch = nb; sl.gram = ga; sl.i8 = i8; sl.pipe8 = pipe8;
                if (F.g) {  // the slice pass reads X anyway: q[map] += 2 a cx cg X'g and c0 += a cg^2 g.g ride along (no gemv + dot launch)
                    sl.gram.gvec = F.g; sl.gram.g_bstride = F.gbs; sl.gram.qout = T.q; sl.gram.q_bstride = N; sl.gram.c0out = T.c0;
                    sl.gram.qscale = scaled(a, 2.0 * F.cx * F.cg); sl.gram.cscale = scaled(a, F.cg * F.cg);
                }
                st.push_back(sl);
                Step mm;
                mm.type = ST_GRAM_I8_MMA; mm.cls = KC_GRAM_I8; mm.batch = nb; mm.gram = ga; mm.i8 = i8; mm.pipe8 = pipe8;
                st.push_back(mm);
                s.i8 = i8;
            }
            st.push_back(s);
            c->gram_flops += F.k * F.ncols * (F.ncols + 1) + 2 * F.k * F.ncols + 2 * F.k;
            // small (latency-bound) problems only: one launch fewer.  For large problems the extra work in the MMA loop costs more
            // than the separate gemv+dot launch (measured at cfg2: Gram 2.09 -> 2.39 ms against a 34 us gemv).
            const bool fuse_lin = !i8.slices && !i8.fused && F.g && c->use_tma && !getenv("PMB_NO_FUSED_Q") && F.k % 2 == 0 && aligned16(F.g) && (nb == 1 || F.gbs % 2 == 0) &&
                                  gram_tma_applicable(F.X, F.ldx, F.xbs, F.k, F.ncols, nb) &&
                                  gram_fused_linear_supported(F.k, F.ncols, nb, c->sm_count);
            if (fuse_lin) {
                // q[map] += 2 a cx cg X'g and c0 += a cg^2 g.g are accumulated by the Gram kernel itself (diagonal tiles)
                Step& gs = st.back();
                gs.gram.gvec = F.g; gs.gram.g_bstride = F.gbs; gs.gram.qout = T.q; gs.gram.q_bstride = N; gs.gram.c0out = T.c0;
                gs.gram.qscale = scaled(a, 2.0 * F.cx * F.cg); gs.gram.cscale = scaled(a, F.cg * F.cg);
            } else if (F.g && !i8.slices && !i8.fused) {
                Step v;
                v.type = ST_GEMV; v.cls = KC_REDUCE; v.batch = nb;
                GemvArgs g{};
                g.A = F.X; g.lda = F.ldx; g.a_bstride = F.xbs; g.rows = F.k; g.cols = F.ncols; g.trans = 1;
                g.v = F.g; g.v_bstride = F.gbs; g.v_off = 0; g.y = T.q; g.y_bstride = N; g.map = F.ident ? nullptr : F.map;
                g.scale = scaled(a, 2.0 * F.cx * F.cg); g.accumulate = 1;
                v.gemv = g;
                DotArgs d{};
                d.x = F.g; d.x_bstride = F.gbs; d.y = F.g; d.y_bstride = F.gbs; d.n = F.k; d.out = T.c0; d.out_bstride = 1;
                d.scale = scaled(a, F.cg * F.cg); d.accumulate = 1;
                v.dot = d;
                v.type = ST_GEMV_DOT;  // q += 2 a cx X'g and c0 += a g.g in one launch
                st.push_back(v);
            }
        } else {
            // M = Xf' Xh (scratch);  H[mf[i], mh[j]] += a M ; H[mh[j], mf[i]] += a M ;  q[mf] += a Xf' gh ; q[mh] += a Xh' gf ; c0 += a gf.gh
            DenseSide Hs = densify(c, gt.h, st, nb);
            double* M = dalloc<double>(c, nb * F.ncols * Hs.ncols);
            Step s;
            s.type = ST_GEMM; s.cls = KC_GEMM; s.batch = nb;
            GemmTNArgs gm{};
            gm.X = F.X; gm.ldx = F.ldx; gm.x_bstride = F.xbs; gm.Y = Hs.X; gm.ldy = Hs.ldx; gm.y_bstride = Hs.xbs;
            gm.k = F.k; gm.n = F.ncols; gm.p = Hs.ncols; gm.M = M; gm.ldm = F.ncols; gm.m_bstride = F.ncols * Hs.ncols;
            s.gemm = gm;
            s.gscratch = make_gram_scratch(c, false, gm.k, gm.n, gm.p, nb);
            st.push_back(s);
            Step sc;
            sc.type = ST_SCATTER_SYM; sc.cls = KC_MISC; sc.batch = nb;
            ScatterSymArgs ss{};
            ss.M = M; ss.ldm = F.ncols; ss.m_bstride = F.ncols * Hs.ncols; ss.n = F.ncols; ss.p = Hs.ncols; ss.xmap = F.map; ss.ymap = Hs.map;
            ss.H = T.Q; ss.ldh = N; ss.h_bstride = N * N; ss.alpha = scaled(a, F.cx * Hs.cx);
            sc.ss = ss;
            st.push_back(sc);
            c->gram_flops += 2 * F.k * F.ncols * Hs.ncols + 2 * F.k * (F.ncols + Hs.ncols) + 2 * F.k;
            auto gemv_q = [&](const DenseSide& X, const DenseSide& G) {
                if (!G.g) return;
                Step v;
                v.type = ST_GEMV; v.cls = KC_REDUCE; v.batch = nb;
                GemvArgs g{};
                g.A = X.X; g.lda = X.ldx; g.a_bstride = X.xbs; g.rows = X.k; g.cols = X.ncols; g.trans = 1;
                g.v = G.g; g.v_bstride = G.gbs; g.v_off = 0; g.y = T.q; g.y_bstride = N; g.map = X.map;
                g.scale = scaled(a, X.cx * G.cg); g.accumulate = 1;
                v.gemv = g;
                st.push_back(v);
            };
            gemv_q(F, Hs);
            gemv_q(Hs, F);
            if (F.g && Hs.g) {
                Step dd;
                dd.type = ST_DOT; dd.cls = KC_REDUCE; dd.batch = nb;
                DotArgs d{};
                d.x = F.g; d.x_bstride = F.gbs; d.y = Hs.g; d.y_bstride = Hs.gbs; d.n = F.k; d.out = T.c0; d.out_bstride = 1;
                d.scale = scaled(a, F.cg * Hs.cg); d.accumulate = 1;
                dd.dot = d;
                st.push_back(dd);
            }
        }
    }
    for (auto& bt : qf.bilins) {
        const Node& qn = c->nodes[bt.mat];
        std::vector<int32_t> xm, ym;
        for (int64_t v : bt.xv) xm.push_back(optcol(c, v));
        for (int64_t v : bt.yv) ym.push_back(optcol(c, v));
        Step s;
        s.type = ST_BILINEAR; s.cls = KC_MISC; s.batch = nb;
        BilinearArgs b{};
        b.Q = qn.dptr; b.ldq = qn.rows; b.q_bstride = qn.bstride; b.nx = (int64_t)xm.size(); b.ny = (int64_t)ym.size();
        b.xmap = upload(c, xm); b.ymap = upload(c, ym);
        b.H = T.Q; b.ldh = N; b.h_bstride = N * N; b.alpha = alpha_dev(c, bt.a);
        s.bl = b;
        st.push_back(s);
    }
    for (auto& ct : qf.coos) {
        std::vector<int32_t> ii, jj;
        for (int64_t v : ct.i) ii.push_back(optcol(c, v));
        for (int64_t v : ct.j) jj.push_back(optcol(c, v));
        Step s;
        s.type = ST_COO; s.cls = KC_MISC; s.batch = nb;
        CooQuadArgs q{};
        q.i = upload(c, ii); q.j = upload(c, jj); q.c = upload(c, ct.c); q.nnz = (int64_t)ii.size();
        q.H = T.Q; q.ldh = N; q.h_bstride = N * N; q.alpha = alpha_dev(c, ct.a);
        s.coo = q;
        st.push_back(s);
    }
    return T;
}

static void plan_objective(pmb_ctx* c) {
    if (c->obj_node < 0) {  // default objective: zero AffineFunction (model.jl:18)
        c->obj_const = true; c->obj_kind = PMB_KIND_AFF;
        double* qc = dalloc<double>(c, c->batch * (c->nvars + 1), true);
        c->q = qc; c->c0 = qc + c->batch * c->nvars;
        return;
    }
    QuadTarget T = plan_quadratic(c, quad_of(c, c->obj_node), c->obj_const, c->obj_kind == PMB_KIND_QUAD, true);
    c->Q = T.Q; c->q = T.q; c->c0 = T.c0;
}

// how many places read numeric node `id` (other numeric nodes, the objective, every constraint form)
static int count_refs_aff(const AffForm& f, int id) {
    int n = 0;
    for (auto& t : f.mats) n += (t.mat == id) + (t.a.node == id);
    for (auto& t : f.units) n += (t.a.node == id);
    for (auto& t : f.offs) n += (t.vec == id) + (t.a.node == id);
    for (auto& t : f.sps) n += (t.mat == id) + (t.a.node == id);
    return n;
}
static int count_refs_quad(const QuadForm& q, int id) {
    int n = count_refs_aff(q.aff, id);
    for (auto& g : q.grams) n += (g.a.node == id) + count_refs_aff(g.f, id) + (g.same ? 0 : count_refs_aff(g.h, id));
    for (auto& b : q.bilins) n += (b.mat == id) + (b.a.node == id);
    for (auto& t : q.coos) n += (t.a.node == id);
    return n;
}
static int count_refs(pmb_ctx* c, int id) {
    int n = 0;
    for (auto& nd : c->nodes) {
        if (!nd.needed) continue;
        for (int in : nd.in) n += (in == id);
        for (auto& t : nd.lin) n += (t.src == id) + (t.a.node == id);
    }
    if (c->obj_node >= 0) n += count_refs_quad(quad_of(c, c->obj_node), id);
    for (auto& con : c->cons) n += con.is_quad ? count_refs_quad(*con.quad, id) : count_refs_aff(con.form, id);
    return n;
}

// ---- run ---------------------------------------------------------------------------------------
static void run_step(pmb_ctx* c, const Step& s) {
    cudaStream_t st = c->stream;
    switch (s.type) {
        case ST_PROLOGUE: launch_prologue(s.jobs, s.njobs, s.job_blocks, s.batch, c->inst_offset, c->d_counter, c->d_ticket, s.tick, st); break;
        case ST_GEMV_DOT: launch_gemv_t_dot(s.gemv, s.dot, s.batch, st); break;
        case ST_PACK_P: launch_pack_upper(s.src, s.n, s.src_bstride, s.out, s.out_bstride, s.batch, st); break;
        case ST_PACK_A:
            launch_pack_constraints(c->qp_blocks_dev, (int)c->cons.size(), c->nvars, c->qp_Ax, c->qp_nnzA, c->qp_l, c->qp_u, c->qp_m, s.batch,
                                    c->sm_count, st);
            break;
        case ST_MEMSET: CK(cudaMemsetAsync(s.out, 0, (size_t)s.n * sizeof(double), st)); break;
        case ST_LINCOMB: launch_lincomb(s.out, s.out_bstride, s.n, s.lterms, s.nl, s.cst, s.batch, st); break;
        case ST_TRANSPOSE: launch_transpose(s.out, s.out_bstride, s.src, s.src_bstride, s.n, s.n2, s.batch, st); break;
        case ST_GEMV: launch_gemv(s.gemv, s.batch, st); break;
        case ST_DOT: launch_dot(s.dot, s.batch, st); break;
        case ST_SPROD: launch_scalar_prod(s.out, s.out_bstride, s.scale, s.s0, s.s0_bs, s.s1, s.s1_bs, s.batch, st); break;
        case ST_AFFINE: launch_affine_assemble(s.aff, s.batch, c->sm_count, st); break;
        case ST_GRAM_I8_SLICE:
            if (!launch_gram_i8_slice(s.gram, s.i8, s.batch, c->sm_count, st)) fail(PMB_ECUDA, "int8-slice Gram: slice launch refused");
            break;
        case ST_GRAM_I8_MMA:
            if (c->fb_mode) break;   // fp64 replay
            if (!launch_gram_i8_mma(s.gram, s.i8, s.batch, c->sm_count, st)) fail(PMB_ECUDA, "int8-slice Gram: tcgen05 launch refused");
            break;
        case ST_GRAM_I8_FUSED:
            if (c->fb_mode) break;
            if (!launch_gram_i8_fused(s.gram, s.i8, s.batch, c->sm_count, st)) fail(PMB_ECUDA, "fused int8 Gram: launch refused");
            break;
        case ST_GRAM:
            if (s.i8.slices || s.i8.fused) {  // companion of the int8 steps: must be the TMA kernel, it is the one that skips certified instances
                GramArgs ga = s.gram;
                if (c->fb_mode) { ga.cert_stats = nullptr; ga.cert_force = 0; }   // fp64 replay: compute every instance
                if (!launch_gram_syrk_tma(ga, s.gscratch, s.batch, c->sm_count, st)) fail(PMB_ECUDA, "fp64 Gram (TMA) launch refused");
                break;
            }
            if (!(c->use_tma && launch_gram_syrk_tma(s.gram, s.gscratch, s.batch, c->sm_count, st)))
                launch_gram_syrk(s.gram, s.batch, c->sm_count, st);
            break;
        case ST_GEMM:
            if (!(c->use_tma && launch_gemm_tn_tma(s.gemm, s.gscratch, s.batch, c->sm_count, st)))
                launch_gemm_tn(s.gemm, s.batch, c->sm_count, st);
            break;
        case ST_SCATTER_VEC: launch_scatter_vec(s.sv, s.batch, st); break;
        case ST_BILINEAR: launch_bilinear_scatter(s.bl, s.batch, st); break;
        case ST_COO: launch_coo_quad(s.coo, s.batch, st); break;
        case ST_SCATTER_SYM: launch_scatter_sym(s.ss, s.batch, st); break;
        default: fail(PMB_EINVAL, "internal: unknown step type %d", s.type);
    }
    CK(cudaGetLastError());
}

static void collect_timing(pmb_ctx* c) {
    if (!c->ev_pending) return;
    CK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < c->steps.size(); ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->ev[2 * i], c->ev[2 * i + 1]));
        c->cls_ms[c->steps[i].cls] += ms;
        c->cls_launches[c->steps[i].cls] += 1;
    }
    c->ev_pending = false;
}

// Capture one update into a CUDA graph: two branches -- objective chain on the context stream, constraint chain on an auxiliary
// stream -- with the HOST-Parameter uploads and the host sinks of the packed QP as memcpy nodes.  fallback == true captures the plain
// fp64 replay used when the pipelined int8 graph reports an uncertifiable instance (no chunked upload, no tcgen05 launches).
// eager == true: the same multi-stream schedule is executed directly (no capture) with timing events at the milestones -- a debugging
// aid (env PMB_TRACE=1, see pmb_update) that prints when each chunk landed, was multiplied and left.
static void capture_update_graph(pmb_ctx* ctx, bool fallback, cudaGraph_t* graph_out, cudaGraphExec_t* exec_out, bool eager = false) {
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    const bool trace = eager && ctx->trace;
    auto mark = [&](const std::string& label, cudaStream_t on) {
        if (!trace) return;
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        CK(cudaEventRecord(e, on));
        marks.emplace_back(label, e);
    };
    const int64_t N = ctx->nvars;
    cudaStream_t main_s = ctx->stream, aux_s = ctx->aux_stream;
    const bool pipe8 = !fallback && ctx->pipe8 && ctx->pipe_param >= 0 && ctx->h2d_in_graph;
    const bool pipe = !fallback && !ctx->pipe8 && ctx->pipe_param >= 0 && ctx->h2d_in_graph;
    if (pipe || pipe8) {
        if (!ctx->h2d_stream) {
            CK(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ctx->ev_h2d_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_aux_h2d, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_aux_kernels, cudaEventDisableTiming));
        }
    }
    const size_t nranges = pipe8 ? ctx->pipe8_bounds.size() - 1 : 0;
    if (pipe8) {
        if (!ctx->d2h_stream) {
            CK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ctx->ev_d2h_done, cudaEventDisableTiming));
        }
        while (ctx->ev_chunk.size() < nranges) {
            cudaEvent_t e0, e1;
            CK(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            ctx->ev_chunk.push_back(e0); ctx->ev_gram.push_back(e1);
        }
        if (!ctx->repair_flag_dev) {
            ctx->repair_flag_dev = dalloc<int32_t>(ctx, 1, true);
            void* hp = nullptr;
            CK(cudaHostAlloc(&hp, sizeof(int32_t), cudaHostAllocDefault));
            ctx->pinned.push_back(hp);
            ctx->repair_flag_host = (int32_t*)hp;
            *ctx->repair_flag_host = 0;
        }
    }
    ctx->fb_mode = fallback;
    if (!eager) CK(cudaStreamBeginCapture(main_s, cudaStreamCaptureModeThreadLocal));
    try {
        mark("start", main_s);
        auto run_phase = [&](int ph, cudaStream_t on) {
            ctx->stream = on;
            for (auto& s : ctx->steps) if (s.phase == ph) run_step(ctx, s);
            ctx->stream = main_s;
        };
        auto h2d = [&](bool main_branch, cudaStream_t on) {
            if (!ctx->h2d_in_graph) return;
            for (size_t ni = 0; ni < ctx->nodes.size(); ++ni) {
                Node& n = ctx->nodes[ni];
                if ((pipe || pipe8) && (int)ni == ctx->pipe_param) continue;  // travels in chunks on its own stream (below)
                if (!n.needed || n.nop != NOP_PARAM || n.pmode != PMB_PARAM_HOST || n.used_main != main_branch) continue;
                CK(cudaMemcpyAsync(n.dptr, n.hptr, (size_t)(ctx->batch * n.numel) * sizeof(double), cudaMemcpyHostToDevice, on));
            }
        };
        auto sink = [&](double* host, const double* dev, int64_t count, cudaStream_t on) {
            if (host && dev && count > 0) CK(cudaMemcpyAsync(host, dev, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, on));
        };
        const int64_t nb = ctx->batch;
        run_phase(0, main_s);                                        // prologue: generators, zeroing, tick
        if (pipe) CK(cudaMemsetAsync(ctx->pipe_gate, 0, sizeof(int32_t) * (size_t)ctx->pipe_chunks, main_s));
        CK(cudaEventRecord(ctx->ev_fork0, main_s));
        CK(cudaStreamWaitEvent(aux_s, ctx->ev_fork0, 0));
        run_phase(4, aux_s);                                         // generators of constraint-only Parameters (when split off the prologue)
        h2d(false, aux_s);                                           // constraint-only Parameters travel on the aux branch
        if (pipe || pipe8) {
            // the Gram operand: column chunks on a dedicated copy stream -- queued behind the (small) constraint Parameters so that the
            // constraint chain can run and leave early.  fp64 path: each chunk is followed by its arrival flag (the persistent Gram
            // kernel polls it); int8 path: each chunk is followed by an event the per-chunk launches wait on.
            CK(cudaEventRecord(ctx->ev_aux_h2d, aux_s));
            CK(cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_aux_h2d, 0));
            const Node& pn = ctx->nodes[(size_t)ctx->pipe_param];
            if (pipe) {
                for (int q = 0; q < ctx->pipe_chunks; ++q) {
                    const int64_t c0 = (int64_t)q * ctx->pipe_gate_cols, c1 = std::min<int64_t>(pn.cols, c0 + ctx->pipe_gate_cols);
                    CK(cudaMemcpyAsync(pn.dptr + c0 * pn.rows, pn.hptr + c0 * pn.rows, (size_t)((c1 - c0) * pn.rows) * sizeof(double),
                                       cudaMemcpyHostToDevice, ctx->h2d_stream));
                    CK(cudaMemcpyAsync(ctx->pipe_gate + q, ctx->pipe_ones + q, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->h2d_stream));
                }
            } else {
                for (size_t r = 0; r < nranges; ++r) {
                    const int64_t c0 = ctx->pipe8_bounds[r], c1 = ctx->pipe8_bounds[r + 1];
                    CK(cudaMemcpyAsync(pn.dptr + c0 * pn.rows, pn.hptr + c0 * pn.rows, (size_t)((c1 - c0) * pn.rows) * sizeof(double),
                                       cudaMemcpyHostToDevice, ctx->h2d_stream));
                    CK(cudaEventRecord(ctx->ev_chunk[r], ctx->h2d_stream));
                    mark("h2d chunk " + std::to_string(r), ctx->h2d_stream);
                }
            }
            CK(cudaEventRecord(ctx->ev_h2d_done, ctx->h2d_stream));
        }
        h2d(true, main_s);
        run_phase(1, main_s);                                        // derived numeric nodes (shared by both chains)
        CK(cudaEventRecord(ctx->ev_fork1, main_s));
        CK(cudaStreamWaitEvent(aux_s, ctx->ev_fork1, 0));
        run_phase(3, aux_s);                                         // constraint blocks (+ stacked-A pack)
        if (pipe) {  // the persistent Gram holds every SM until the upload is over: let the constraint kernels go first
            CK(cudaEventRecord(ctx->ev_aux_kernels, aux_s));
            CK(cudaStreamWaitEvent(main_s, ctx->ev_aux_kernels, 0));
        }
        sink(ctx->sink_Ax, ctx->qp_Ax, nb * ctx->qp_nnzA, aux_s);
        sink(ctx->sink_l, ctx->qp_l, nb * ctx->qp_m, aux_s);
        sink(ctx->sink_u, ctx->qp_u, nb * ctx->qp_m, aux_s);
        mark("constraint chain + A_x sink", aux_s);
        CK(cudaEventRecord(ctx->ev_join, aux_s));
        // objective chain (+ upper-triangular P pack)
        bool px_streamed = false;
        for (size_t si = 0; si < ctx->steps.size(); ++si) {
            const Step& s = ctx->steps[si];
            if (s.phase != 2) continue;
            if (pipe8 && s.pipe8) {
                if (s.type == ST_GRAM_I8_SLICE) {
                    const Step* mm = nullptr;
                    for (size_t sj = si + 1; sj < ctx->steps.size() && !mm; ++sj)
                        if (ctx->steps[sj].type == ST_GRAM_I8_MMA && ctx->steps[sj].pipe8) mm = &ctx->steps[sj];
                    if (!mm) fail(PMB_EINVAL, "internal: pipelined slice step without its tcgen05 step");
                    const bool stream_px = ctx->sink_Px && mm->gram.packed_dev == ctx->qp_Px && ctx->qp_P_fused;
                    int64_t done_prev = 0;   // columns of P_x complete (on both tile grids, 96 and 80 wide) and already sent
                    for (size_t r = 0; r < nranges; ++r) {
                        const int64_t c0 = ctx->pipe8_bounds[r], c1 = ctx->pipe8_bounds[r + 1];
                        CK(cudaStreamWaitEvent(main_s, ctx->ev_chunk[r], 0));
                        if (!launch_gram_i8_slice(s.gram, s.i8, 1, ctx->sm_count, main_s, c0, c1) ||
                            !launch_gram_i8_mma(mm->gram, mm->i8, 1, ctx->sm_count, main_s, c0, c1))
                            fail(PMB_ECUDA, "int8-slice Gram: pipelined launch refused");
                        CK(cudaGetLastError());
                        mark("gram range " + std::to_string(r), main_s);
                        if (stream_px) {   // columns [c0, c1) of the packed upper triangle are final: they leave while later chunks arrive
                            CK(cudaEventRecord(ctx->ev_gram[r], main_s));
                            CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_gram[r], 0));
                            // a tile column belongs to the range holding its LAST column: complete are the columns below the last tile boundary
                            const int64_t done = r + 1 == nranges ? c1 : std::min(c1 / 96 * 96, c1 / 80 * 80);
                            const int64_t o0 = done_prev * (done_prev + 1) / 2, o1 = done * (done + 1) / 2;
                            if (o1 > o0)
                                CK(cudaMemcpyAsync(ctx->sink_Px + o0, ctx->qp_Px + o0, (size_t)(o1 - o0) * sizeof(double), cudaMemcpyDeviceToHost, ctx->d2h_stream));
                            done_prev = std::max(done_prev, done);
                            mark("d2h P_x range " + std::to_string(r), ctx->d2h_stream);
                        }
                    }
                    if (stream_px) { CK(cudaEventRecord(ctx->ev_d2h_done, ctx->d2h_stream)); px_streamed = true; }
                    launch_gram_i8_flag(s.i8, 1, ctx->repair_flag_dev, main_s);
                    CK(cudaMemcpyAsync(ctx->repair_flag_host, ctx->repair_flag_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, main_s));
                    continue;
                }
                if (s.type == ST_GRAM_I8_MMA) continue;   // ran per chunk above
                // the fp64 companion stays IN the stream: it leaves at once when the certificate accepted the operand, and otherwise computes
                // Q (and the packed P_x) on the fp64 pipe -- device-resident results are valid in stream order whatever the verdict; only the
                // host copy of P_x, streamed out per range before the verdict was final, is refreshed by sync_and_repair() below
            }
            ctx->stream = main_s;
            run_step(ctx, s);
        }
        ctx->stream = main_s;
        if (!px_streamed) sink(ctx->sink_Px, ctx->qp_Px, nb * ctx->qp_nnzP, main_s);
        sink(ctx->sink_q, ctx->q, nb * N, main_s);
        sink(ctx->sink_c0, ctx->c0, nb, main_s);
        CK(cudaStreamWaitEvent(main_s, ctx->ev_join, 0));
        if (pipe || pipe8) CK(cudaStreamWaitEvent(main_s, ctx->ev_h2d_done, 0));
        if (px_streamed) CK(cudaStreamWaitEvent(main_s, ctx->ev_d2h_done, 0));
        mark("end", main_s);
    } catch (...) {
        ctx->stream = main_s;
        ctx->fb_mode = false;
        if (!eager) {
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(main_s, &g);
            if (g) cudaGraphDestroy(g);
        }
        throw;
    }
    ctx->fb_mode = false;
    if (eager && !trace) return;
    if (eager) {
        CK(cudaStreamSynchronize(main_s));
        for (auto& m : marks) {
            float ms = 0.f;
            CK(cudaEventSynchronize(m.second));
            CK(cudaEventElapsedTime(&ms, marks[0].second, m.second));
            fprintf(stderr, "[pmb trace] %8.3f ms  %s\n", ms, m.first.c_str());
        }
        for (auto& m : marks) cudaEventDestroy(m.second);
        return;
    }
    CK(cudaStreamEndCapture(main_s, graph_out));
    CK(cudaGraphInstantiate(exec_out, *graph_out, 0));
    if (!fallback) ctx->pipe8_live = pipe8;
}

void build_update_graph(pmb_ctx* ctx) {
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    if (ctx->graph) { cudaGraphDestroy(ctx->graph); ctx->graph = nullptr; }
    if (ctx->graph_exec_fb) { cudaGraphExecDestroy(ctx->graph_exec_fb); ctx->graph_exec_fb = nullptr; }
    if (ctx->graph_fb) { cudaGraphDestroy(ctx->graph_fb); ctx->graph_fb = nullptr; }
    ctx->pipe8_live = false;
    if (ctx->use_graph && !ctx->steps.empty()) {
        if (!ctx->aux_stream) {
            CK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ctx->ev_fork0, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_fork1, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        capture_update_graph(ctx, false, &ctx->graph, &ctx->graph_exec);
    }
}

}  // namespace pmb

using namespace pmb;

static const char* kClassNames[KC_COUNT] = {"affine_assemble", "gram_dmma", "gemm_tn_dmma", "philox_uniform", "h2d_stage", "reduce", "misc", "pack", "gram_i8_tcgen05", "gram_i8_slice"};

extern "C" {

const char* pmb_version(void) { return "parametron-b200 0.1 (sm_100a)"; }

int pmb_ctx_create(int device, void* stream, int64_t batch, pmb_ctx** out) {
    if (!out || batch < 1) return PMB_EINVAL;
    *out = nullptr;
    pmb_ctx* c = new (std::nothrow) pmb_ctx();
    if (!c) return PMB_ENOMEM;
    auto bail = [&](int code) { delete c; return code; };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return bail(PMB_ECUDA);  // no silent CPU fallback: fail loudly
    if (device < 0 || device >= ndev) return bail(PMB_EINVAL);
    if (cudaSetDevice(device) != cudaSuccess) return bail(PMB_ECUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(PMB_ECUDA);
    if (prop.major != 10) return bail(PMB_EUNSUPPORTED);  // kernels are sm_100a-only
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->batch = batch;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(PMB_ECUDA);
        c->own_stream = true;
    }
    configure_gram_kernels();
    configure_affine_tma_kernels();
    configure_gram_tma_kernels();
    configure_gram_i8_kernels();
    configure_gram_i8_fused_kernel();
    if (const char* gm = getenv("PMB_GRAM_MODE")) c->gram_mode = atoi(gm);
    c->trace = getenv("PMB_TRACE") != nullptr;   // read once: pmb_update is the hot call
    const char* ng = getenv("PMB_NO_GRAPH");
    c->use_graph = !(ng && ng[0] == '1');
    const char* nt = getenv("PMB_NO_TMA");
    c->use_tma = !(nt && nt[0] == '1');
    *out = c;
    return PMB_OK;
}

int pmb_ctx_destroy(pmb_ctx* c) {
    if (!c) return PMB_EINVAL;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    if (c->graph_exec_fb) cudaGraphExecDestroy(c->graph_exec_fb);
    if (c->graph_fb) cudaGraphDestroy(c->graph_fb);
    if (c->graph_exec_timing) cudaGraphExecDestroy(c->graph_exec_timing);
    if (c->graph_timing) cudaGraphDestroy(c->graph_timing);
    for (auto e : c->ev_chunk) cudaEventDestroy(e);
    for (auto e : c->ev_gram) cudaEventDestroy(e);
    if (c->ev_d2h_done) cudaEventDestroy(c->ev_d2h_done);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (auto e : c->ev) cudaEventDestroy(e);
    if (c->ev_fork0) cudaEventDestroy(c->ev_fork0);
    if (c->ev_fork1) cudaEventDestroy(c->ev_fork1);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->ev_h2d_done) cudaEventDestroy(c->ev_h2d_done);
    if (c->ev_aux_h2d) cudaEventDestroy(c->ev_aux_h2d);
    if (c->ev_aux_kernels) cudaEventDestroy(c->ev_aux_kernels);
    for (void* p : c->allocs) cudaFree(p);
    for (void* p : c->pinned) cudaFreeHost(p);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return PMB_OK;
}

const char* pmb_last_error(const pmb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pmb_initialize(pmb_ctx* ctx, const int64_t* varmap) {
    API_BEGIN(ctx)
    if (ctx->initialized) fail(PMB_EINITIALIZED, "Model has already been initialized.");
    CK(cudaSetDevice(ctx->device));
    const int64_t N = ctx->nvars;
    // mapindices! (model.jl:100-107): model variable -> optimizer variable; must be a permutation of 1..N
    ctx->varmap.resize((size_t)N);
    {
        std::vector<char> seen((size_t)N, 0);
        for (int64_t i = 0; i < N; ++i) {
            int64_t v = varmap ? varmap[i] : i + 1;
            if (v < 1 || v > N || seen[(size_t)(v - 1)]) fail(PMB_EINVAL, "varmap is not a permutation of 1..%lld", (long long)N);
            seen[(size_t)(v - 1)] = 1;
            ctx->varmap[(size_t)i] = v;
        }
    }
    // update order: bucket order, insertion order within a bucket (moi_interop.jl:180-193, 236-247)
    ctx->update_order.clear();
    for (int b = 0; b < 9; ++b)
        for (size_t i = 0; i < ctx->cons.size(); ++i)
            if (ctx->cons[i].bucket == b) ctx->update_order.push_back((int)i);
    // which numeric nodes are live
    if (ctx->obj_node >= 0) mark_quad(ctx, quad_of(ctx, ctx->obj_node));
    for (auto& con : ctx->cons) {
        if (con.is_quad) mark_quad(ctx, *con.quad);
        else mark_aff(ctx, con.form);
    }
    ctx->d_counter = (unsigned long long*)dalloc<unsigned long long>(ctx, 1, true);
    ctx->d_ticket = (unsigned int*)dalloc<unsigned int>(ctx, 1, true);
    // HOST Parameters can live inside the graph (memcpy nodes) when their buffers are pinned and fixed from here on
    ctx->h2d_in_graph = ctx->use_graph;
    for (auto& n : ctx->nodes) {
        if (!n.needed || n.nop != NOP_PARAM || n.pmode != PMB_PARAM_HOST) continue;
        unsigned int flags = 0;
        if (!n.hptr || cudaHostGetFlags(&flags, n.hptr) != cudaSuccess) { ctx->h2d_in_graph = false; (void)cudaGetLastError(); }
    }
    const size_t nn = ctx->nodes.size();
    size_t tagged = 0;
    auto tag_phase = [&](int ph) {
        for (; tagged < ctx->steps.size(); ++tagged) ctx->steps[tagged].phase = ph;
    };
    for (size_t i = 0; i < nn; ++i)
        if (ctx->nodes[i].needed && is_numeric(ctx->nodes[i].kind)) plan_numeric(ctx, (int)i);
    tag_phase(1);
    plan_objective(ctx);
    tag_phase(2);
    for (int ci : ctx->update_order) {
        Constraint& con = ctx->cons[(size_t)ci];
        if (con.sparse) { plan_sparse_block(ctx, con); continue; }
        if (con.is_quad) {  // ScalarQuadraticFunction-in-set (buckets 6-8, moi_interop.jl:186-188)
            QuadTarget T = plan_quadratic(ctx, *con.quad, con.is_const, true);
            con.Qc = T.Q; con.C = T.q; con.d = T.c0; con.nterms = 0;
            continue;
        }
        con.d = dalloc<double>(ctx, ctx->batch * con.rows, true);
        con.nterms = aff_nterms(con.form);
        // Zero-copy block: `M * x` over ALL variables in optimizer order with unit coefficient, M a matrix Parameter the library fills only inside
        // pmb_update (UNIFORM generator or HOST upload) -- the coefficient block IS the Parameter's buffer (same column-major m x N layout, same
        // instance stride): the view points at it and only the offsets d are assembled.  Caller-written device Parameters keep the copy (the block
        // must stay the snapshot update! took).  PMB_NO_BLOCK_ALIAS=1: always copy
        bool aliased = false;
        if (!con.is_const && con.form.mats.size() == 1 && con.form.units.empty() && !getenv("PMB_NO_BLOCK_ALIAS")) {
            const MatTerm& mt = con.form.mats[0];
            const Node& sn = ctx->nodes[(size_t)mt.mat];
            bool ident = mt.a.node < 0 && mt.a.c == 1.0 && mt.row_off == 0 && mt.rows == con.rows && mt.src_col_off == 0 && mt.cols == N &&
                         sn.nop == NOP_PARAM && sn.kind == PMB_KIND_NUM_MATRIX && sn.bsr_b == 0 && sn.rows == con.rows && sn.cols == N && sn.owned && sn.dptr &&
                         !sn.lazy_fill && (sn.pmode == PMB_PARAM_UNIFORM || sn.pmode == PMB_PARAM_HOST) && (int64_t)mt.vars.size() == N;
            for (int64_t j = 0; ident && j < N; ++j) ident = optcol(ctx, mt.vars[(size_t)j]) == (int32_t)j;
            if (ident) {
                AffForm only_off = con.form;
                only_off.mats.clear();
                con.C = sn.dptr;
                AffBlockDev b = lower_aff(ctx, only_off, nullptr, 0, con.d, con.rows, true, nullptr);
                ctx->steps.push_back(affine_step(b, ctx->batch));
                ctx->affine_bytes += 16 * con.rows;
                aliased = true;
            }
        }
        if (aliased) continue;
        con.C = dalloc<double>(ctx, ctx->batch * con.rows * N, true);
        AffBlockDev b = lower_aff(ctx, con.form, con.C, con.rows * N, con.d, con.rows, false, nullptr);
        (con.is_const ? ctx->init_steps : ctx->steps).push_back(affine_step(b, ctx->batch));
        if (!con.is_const) ctx->affine_bytes += 16 * (int64_t)b.ntouched * con.rows + 16 * con.rows;
    }
    tag_phase(3);
    if (ctx->qp_pack) plan_qp_pack(ctx);
    for (; tagged < ctx->steps.size(); ++tagged) ctx->steps[tagged].phase = ctx->steps[tagged].type == ST_PACK_P ? 2 : 3;
    // which Parameters does the main branch (numeric nodes + objective) read?  The rest belong to the constraint chain only.
    {
        std::vector<char> saved(ctx->nodes.size());
        for (size_t i = 0; i < ctx->nodes.size(); ++i) { saved[i] = ctx->nodes[i].needed; ctx->nodes[i].needed = false; }
        if (ctx->obj_node >= 0) mark_quad(ctx, quad_of(ctx, ctx->obj_node));
        for (size_t i = 0; i < ctx->nodes.size(); ++i) ctx->nodes[i].used_main = ctx->nodes[i].needed;
        // anything feeding a derived numeric node is produced on the main branch too
        for (size_t i = 0; i < ctx->nodes.size(); ++i) {
            const Node& n = ctx->nodes[i];
            if (!saved[i] || !is_numeric(n.kind) || n.nop == NOP_PARAM || n.nop == NOP_CONST) continue;
            for (int in : n.in) ctx->nodes[(size_t)in].used_main = true;
            for (auto& t : n.lin) { ctx->nodes[(size_t)t.src].used_main = true; if (t.a.node >= 0) ctx->nodes[(size_t)t.a.node].used_main = true; }
        }
        for (size_t i = 0; i < ctx->nodes.size(); ++i) ctx->nodes[i].needed = saved[i];
    }
    if (!ctx->fill_jobs.empty()) {
        // fused prologue launch: Philox fills of every built-in generator, zeroing of (q, c0), update tick.  The generators of Parameters that only the
        // constraint chain reads move to a second launch at the head of the constraint branch when they are large (>= 8 MB per update): the
        // objective chain then starts after ITS Parameters only (with generator fusion that is just the offset vector).  Measured r02 with the
        // fused generators: cfg3 5.460 -> 5.399 ms per sweep (1.500 -> 1.517 M instance-updates/s), cfg2 unchanged; PMB_NO_PROLOGUE_SPLIT=1 keeps
        // the single launch (small models always do: a 2 us launch on the constraint branch would lengthen their critical path)
        std::vector<FillJob> main_jobs, aux_jobs;
        int64_t aux_bytes = 0;
        for (size_t j = 0; j < ctx->fill_jobs.size(); ++j) {
            const int nd = j < ctx->fill_job_nodes.size() ? ctx->fill_job_nodes[j] : -1;
            const bool aux = nd >= 0 && !ctx->nodes[(size_t)nd].used_main && !getenv("PMB_NO_PROLOGUE_SPLIT");
            (aux ? aux_jobs : main_jobs).push_back(ctx->fill_jobs[j]);
            if (aux) aux_bytes += ctx->fill_jobs[j].numel * ctx->batch * (int64_t)sizeof(double);
        }
        if (aux_bytes < ((int64_t)8 << 20) || main_jobs.empty()) { main_jobs = ctx->fill_jobs; aux_jobs.clear(); }
        auto make = [&](std::vector<FillJob>& jobs, int tick, int phase) {
            int total = 0;
            const int64_t cap = (int64_t)ctx->sm_count * 16;
            for (auto& j : jobs) {
                const int64_t items = j.kind == 0 ? (j.numel + 1) / 2 : j.numel;
                int64_t nbk = (items + 255) / 256;
                const int64_t lim = j.kind == 0 ? std::max<int64_t>(1, cap / std::max<int64_t>(1, std::min<int64_t>(ctx->batch, 64))) : 64;
                j.nblocks = (int32_t)std::max<int64_t>(1, std::min(nbk, lim));
                total += j.nblocks;
            }
            Step s;
            s.type = ST_PROLOGUE; s.cls = KC_PHILOX; s.batch = ctx->batch;
            s.jobs = upload(ctx, jobs); s.njobs = (int)jobs.size(); s.job_blocks = total; s.tick = tick;
            s.phase = phase;
            return s;
        };
        if (!aux_jobs.empty()) ctx->steps.insert(ctx->steps.begin(), make(aux_jobs, 0, 4));   // phase 4: head of the constraint branch
        ctx->steps.insert(ctx->steps.begin(), make(main_jobs, 1, 0));
    }
    ctx->launches_per_update = 0;
    ctx->memcpy_per_update = 0;
    for (auto& s : ctx->steps) if (s.type != ST_MEMSET) ctx->launches_per_update += (s.type == ST_GRAM_I8_MMA && !s.i8.force_ns) ? 2 : 1;
    for (auto& n : ctx->nodes) if (n.needed && n.nop == NOP_PARAM && n.pmode == PMB_PARAM_HOST) ctx->memcpy_per_update += 1;
    // constants: evaluated exactly once (constant folding lazyexpression.jl:189-192; isconstant moi_interop.jl:123,153)
    for (auto& s : ctx->init_steps) run_step(ctx, s);
    CK(cudaStreamSynchronize(ctx->stream));
    build_update_graph(ctx);
    ctx->initialized = true;
    API_END(ctx)
}

int pmb_update(pmb_ctx* ctx) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_update before pmb_initialize");
    // setdirty! + evaluate every Parameter exactly once (model.jl:133, parameter.jl:93-104)
    const bool graph_path = !ctx->timing && ctx->graph_exec;
    // HOST Parameters are read from their (pinned) staging buffers by copies that run asynchronously: the previous update's copies must
    // have consumed the buffers before the callbacks below overwrite them (the reference's update! is synchronous, model.jl:132-143)
    if (ctx->memcpy_per_update > 0 && ctx->updates > 0) CK(cudaStreamSynchronize(ctx->stream));
    for (auto& n : ctx->nodes) {
        if (!n.needed || n.nop != NOP_PARAM || n.pmode != PMB_PARAM_HOST) continue;
        if (!n.hptr) fail(PMB_EINVAL, "host Parameter has no buffer (pmb_param_set_host_ptr)");
        if (n.cb) n.cb(n.hptr, n.rows, n.cols, ctx->batch, n.user);
        if (!(graph_path && ctx->h2d_in_graph))
            CK(cudaMemcpyAsync(n.dptr, n.hptr, (size_t)(ctx->batch * n.numel) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!(graph_path && ctx->h2d_in_graph) && ctx->pipe_param >= 0 && ctx->pipe_gate)  // eager paths upload everything up front: open every gate
        CK(cudaMemcpyAsync(ctx->pipe_gate, ctx->pipe_ones, sizeof(int32_t) * (size_t)ctx->pipe_chunks, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->timing) {
        collect_timing(ctx);
        if (ctx->graph_exec_timing) {
            // the steps of one update, serialised on the context stream, with an event-record NODE on either side of every launch: the
            // time stamps are taken by the device between back-to-back graph nodes, so no host launch gap lands inside an event pair
            CK(cudaGraphLaunch(ctx->graph_exec_timing, ctx->stream));
        } else {
            for (size_t i = 0; i < ctx->steps.size(); ++i) {
                CK(cudaEventRecord(ctx->ev[2 * i], ctx->stream));
                run_step(ctx, ctx->steps[i]);
                CK(cudaEventRecord(ctx->ev[2 * i + 1], ctx->stream));
            }
        }
        ctx->ev_pending = true;
    } else if (ctx->graph_exec && (ctx->pipe8_live || ctx->trace)) {
        // The pipelined int8 schedule (chunked upload || per-chunk slice + Gram || streamed P_x download) is issued directly on its four
        // streams: replayed as ONE CUDA graph the copy nodes of the two directions were serialised by the graph scheduler (measured
        // 5.0 ms against 3.9 ms for the identical schedule issued eagerly; the ~60 API calls are hidden behind 3.5 ms of PCIe time).
        // PMB_TRACE=1 additionally prints the timeline of the milestones.
        capture_update_graph(ctx, false, nullptr, nullptr, true);
    } else if (ctx->graph_exec) {
        CK(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
    } else {
        for (auto& s : ctx->steps) run_step(ctx, s);
    }
    if (!graph_path) {  // eager paths: the host sinks are plain stream-ordered copies
        const int64_t nb = ctx->batch;
        auto sink = [&](double* host, const double* dev, int64_t count) {
            if (host && dev && count > 0) CK(cudaMemcpyAsync(host, dev, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        };
        sink(ctx->sink_Ax, ctx->qp_Ax, nb * ctx->qp_nnzA); sink(ctx->sink_l, ctx->qp_l, nb * ctx->qp_m); sink(ctx->sink_u, ctx->qp_u, nb * ctx->qp_m);
        sink(ctx->sink_Px, ctx->qp_Px, nb * ctx->qp_nnzP); sink(ctx->sink_q, ctx->q, nb * ctx->nvars); sink(ctx->sink_c0, ctx->c0, nb);
    }
    ctx->updates += 1;
    API_END(ctx)
}

// Synchronise the context stream; on the pipelined int8 schedule the packed P_x left for the host sink range by range BEFORE the certificate's
// verdict was final: when the verdict ended at fp64 (the in-stream fp64 launch has then recomputed Q and P_x on the device) the host copy is
// refreshed here.  Every entry point that hands results to the host goes through this helper.
static void sync_and_repair(pmb_ctx* ctx) {
    if (cudaError_t e = cudaStreamSynchronize(ctx->stream); e != cudaSuccess) {
        char buf[1024];
        const char* why = gram_i8_fused_diag_text(ctx->fused_diag, buf, sizeof buf);   // a bounded wait of the fused kernel trapped: say which
        fail(PMB_ECUDA, "cudaStreamSynchronize(ctx->stream) failed: %s%s%s", cudaGetErrorString(e), *why ? "; " : "", why);
    }
    if (ctx->pipe8_live && ctx->repair_flag_host && *ctx->repair_flag_host) {
        *ctx->repair_flag_host = 0;
        if (ctx->sink_Px && ctx->qp_Px && ctx->qp_nnzP > 0) {
            CK(cudaMemcpyAsync(ctx->sink_Px, ctx->qp_Px, (size_t)(ctx->batch * ctx->qp_nnzP) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        ctx->repairs += 1;
    }
}

int pmb_synchronize(pmb_ctx* ctx) {
    API_BEGIN(ctx)
    sync_and_repair(ctx);
    API_END(ctx)
}

int pmb_update_count(const pmb_ctx* ctx, int64_t* n) {
    if (!ctx || !n) return PMB_EINVAL;
    *n = ctx->updates;
    return PMB_OK;
}

int pmb_param_upload(pmb_ctx* ctx, pmb_node p, const double* host_values) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_param_upload before pmb_initialize");
    Node& n = node_at(ctx, p);
    if (n.nop != NOP_PARAM || !n.needed || !n.dptr) fail(PMB_EINVAL, "node %d is not a live Parameter", p);
    CK(cudaMemcpyAsync(n.dptr, host_values, (size_t)(ctx->batch * n.numel) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    API_END(ctx)
}

int pmb_param_device_ptr(pmb_ctx* ctx, pmb_node p, double** dptr) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_param_device_ptr before pmb_initialize");
    Node& n = node_at(ctx, p);
    if (n.nop != NOP_PARAM || !dptr) fail(PMB_EINVAL, "node %d is not a Parameter", p);
    if (n.lazy_fill && ctx->updates > 0)   // generated inside the slice pass, never written by updates: materialise the last update's values now (same bits)
        launch_prologue(n.lazy_fill, 1, n.lazy_fill_blocks, ctx->batch, ctx->inst_offset, ctx->d_counter, ctx->d_ticket, 0, ctx->stream);
    *dptr = n.dptr;  // null when the Parameter is not reachable from the objective or a constraint
    API_END(ctx)
}

int pmb_get_objective(pmb_ctx* ctx, pmb_objective_view* out) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_get_objective before pmb_initialize");
    if (!out) fail(PMB_EINVAL, "null out");
    out->kind = ctx->obj_kind; out->sense = ctx->obj_sense; out->is_constant = ctx->obj_const ? 1 : 0;
    out->nvars = ctx->nvars; out->Q = ctx->Q; out->q = ctx->q; out->c0 = ctx->c0;
    API_END(ctx)
}

int pmb_num_blocks(const pmb_ctx* ctx, int32_t* n) {
    if (!ctx || !n) return PMB_EINVAL;
    *n = (int32_t)ctx->cons.size();
    return PMB_OK;
}

static Constraint& block_at(pmb_ctx* ctx, int32_t pos) {
    if (!ctx->initialized) fail(PMB_ENOTINIT, "block query before pmb_initialize");
    if (pos < 0 || pos >= (int32_t)ctx->update_order.size()) fail(PMB_EINVAL, "block position %d out of range", pos);
    return ctx->cons[(size_t)ctx->update_order[(size_t)pos]];
}

int pmb_get_block(pmb_ctx* ctx, int32_t update_pos, pmb_block_view* out) {
    API_BEGIN(ctx)
    if (!out) fail(PMB_EINVAL, "null out");
    Constraint& con = block_at(ctx, update_pos);
    out->set_kind = con.set_kind; out->is_scalar = con.is_scalar ? 1 : 0; out->is_constant = con.is_const ? 1 : 0;
    out->bucket = con.bucket; out->rows = con.rows; out->nvars = ctx->nvars; out->C = con.C; out->d = con.d; out->nterms = con.nterms;
    out->Q = con.Qc;
    out->is_sparse = con.sparse ? 1 : 0; out->nnz = con.nnz; out->vals = con.vals; out->sp_rows = con.sp_rows; out->sp_cols = con.sp_cols;
    API_END(ctx)
}

int pmb_block_term_index(pmb_ctx* ctx, int32_t update_pos, int64_t capacity, int64_t* rows, int64_t* vars, int64_t* src) {
    API_BEGIN(ctx)
    Constraint& con = block_at(ctx, update_pos);
    const AffForm& f = con.form;
    if (con.sparse) fail(PMB_EUNSUPPORTED, "sparse blocks export their static index through pmb_block_view (sp_rows, sp_cols)");
    if (con.is_quad) fail(PMB_EUNSUPPORTED, "quadratic constraints have no affine term index");
    if (!f.order_exact) fail(PMB_EUNSUPPORTED, "the reference's term order of this block is not reproducible from the device lowering");
    if (capacity < con.nterms) fail(PMB_EINVAL, "capacity %lld < %lld terms", (long long)capacity, (long long)con.nterms);
    // update!(::VectorAffineFunction): row-outer, term-inner, varmap applied (moi_interop.jl:64-81)
    int64_t t = 0;
    for (int64_t r = 0; r < f.rows; ++r) {
        for (auto& seg : f.order) {
            if (seg.is_unit) {
                const UnitTerm& u = f.units[(size_t)seg.idx[0]];
                int64_t i = r - u.row_off;
                if (i < 0 || i >= u.rows) continue;
                int64_t v = ctx->varmap[(size_t)(u.vars[(size_t)i] - 1)];
                if (rows) rows[t] = r + 1;
                if (vars) vars[t] = v;
                if (src) src[t] = r + (v - 1) * f.rows;
                ++t;
            } else {
                const MatTerm& m0 = f.mats[(size_t)seg.idx[0]];
                int64_t i = r - m0.row_off;
                if (i < 0 || i >= m0.rows) continue;
                for (int64_t j = 0; j < m0.cols; ++j)
                    for (int mi : seg.idx) {
                        int64_t v = ctx->varmap[(size_t)(f.mats[(size_t)mi].vars[(size_t)j] - 1)];
                        if (rows) rows[t] = r + 1;
                        if (vars) vars[t] = v;
                        if (src) src[t] = r + (v - 1) * f.rows;
                        ++t;
                    }
            }
        }
    }
    if (t != con.nterms) fail(PMB_EINVAL, "internal: enumerated %lld terms, expected %lld", (long long)t, (long long)con.nterms);
    API_END(ctx)
}

int pmb_get_stats(pmb_ctx* ctx, pmb_stats* out) {
    API_BEGIN(ctx)
    if (!out) fail(PMB_EINVAL, "null out");
    out->kernel_launches_per_update = ctx->launches_per_update;
    out->memcpy_nodes_per_update = ctx->memcpy_per_update;
    out->device_bytes_allocated = ctx->bytes_alloc;
    out->uses_cuda_graph = ctx->graph_exec ? 1 : 0;
    out->affine_algorithmic_bytes = ctx->affine_bytes;
    out->gram_algorithmic_flops = ctx->gram_flops;
    API_END(ctx)
}

// ---- result path: value.(model, x) / objectivevalue(model), model.jl:166-173 -------------------------------------------------------
int pmb_set_primal(pmb_ctx* ctx, const double* x, int on_device) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_set_primal before pmb_initialize");
    if (!x) fail(PMB_EINVAL, "null primal");
    const int64_t cnt = ctx->batch * ctx->nvars;
    if (!ctx->primal) {
        ctx->primal = dalloc<double>(ctx, std::max<int64_t>(cnt, 1), true);
        ctx->res_tmp = dalloc<double>(ctx, std::max<int64_t>(cnt, 1) + ctx->batch, true);
    }
    CK(cudaMemcpyAsync(ctx->primal, x, (size_t)cnt * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    if (!on_device) sync_and_repair(ctx);   // the caller's host buffer may be pageable / reused
    ctx->primal_set = true;
    API_END(ctx)
}

int pmb_value(pmb_ctx* ctx, int64_t nv, const int64_t* var_indices, double* out, int out_on_device) {
    API_BEGIN(ctx)
    if (!ctx->primal_set) fail(PMB_EINVAL, "pmb_value before pmb_set_primal (no solution has been handed back)");
    if (nv < 0 || (nv > 0 && (!var_indices || !out))) fail(PMB_EINVAL, "bad arguments");
    if (nv == 0) return PMB_OK;
    std::vector<int64_t> idx((size_t)nv);
    for (int64_t i = 0; i < nv; ++i) {
        const int64_t v = var_indices[i];
        if (v < 1 || v > ctx->nvars) fail(PMB_EINVAL, "variable index %lld out of range 1..%lld", (long long)v, (long long)ctx->nvars);
        idx[(size_t)i] = ctx->varmap[(size_t)(v - 1)] - 1;     // m.model_var_to_optimizer[x.index]
    }
    if (ctx->res_idx_cap < nv) { ctx->res_idx = dalloc<int64_t>(ctx, nv); ctx->res_idx_cap = nv; }
    CK(cudaMemcpyAsync(ctx->res_idx, idx.data(), (size_t)nv * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // idx is a local
    double* dst = out;
    if (!out_on_device) {
        if (ctx->res_gather_cap < nv * ctx->batch) { ctx->res_gather = dalloc<double>(ctx, nv * ctx->batch); ctx->res_gather_cap = nv * ctx->batch; }
        dst = ctx->res_gather;
    }
    launch_gather(dst, ctx->primal, ctx->res_idx, nv, ctx->nvars, ctx->batch, ctx->stream);
    CK(cudaGetLastError());
    if (!out_on_device) {
        CK(cudaMemcpyAsync(out, dst, (size_t)(nv * ctx->batch) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        sync_and_repair(ctx);
    }
    API_END(ctx)
}

int pmb_objective_value(pmb_ctx* ctx, double* out_host) {
    API_BEGIN(ctx)
    if (!ctx->primal_set) fail(PMB_EINVAL, "pmb_objective_value before pmb_set_primal (no solution has been handed back)");
    if (!out_host) fail(PMB_EINVAL, "null out");
    const int64_t N = ctx->nvars, nb = ctx->batch;
    double* val = ctx->res_tmp + nb * N;
    if (ctx->c0) CK(cudaMemcpyAsync(val, ctx->c0, (size_t)nb * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    else CK(cudaMemsetAsync(val, 0, (size_t)nb * sizeof(double), ctx->stream));
    if (ctx->q && N > 0) {   // + q'x
        DotArgs d{};
        d.x = ctx->q; d.x_bstride = N; d.y = ctx->primal; d.y_bstride = N; d.n = N; d.out = val; d.out_bstride = 1;
        d.scale = AlphaDev{1.0, nullptr, 0}; d.accumulate = 1;
        launch_dot(d, nb, ctx->stream);
    }
    if (ctx->obj_kind == PMB_KIND_QUAD && ctx->Q && N > 0) {   // + 1/2 x'(Qx)
        GemvArgs g{};
        g.A = ctx->Q; g.lda = N; g.a_bstride = N * N; g.rows = N; g.cols = N; g.trans = 0;
        g.v = ctx->primal; g.v_bstride = N; g.v_off = 0; g.y = ctx->res_tmp; g.y_bstride = N; g.map = nullptr;
        g.scale = AlphaDev{1.0, nullptr, 0}; g.accumulate = 0;
        launch_gemv(g, nb, ctx->stream);
        DotArgs d{};
        d.x = ctx->res_tmp; d.x_bstride = N; d.y = ctx->primal; d.y_bstride = N; d.n = N; d.out = val; d.out_bstride = 1;
        d.scale = AlphaDev{0.5, nullptr, 0}; d.accumulate = 1;
        launch_dot(d, nb, ctx->stream);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_host, val, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    sync_and_repair(ctx);
    API_END(ctx)
}

int pmb_set_gram_mode(pmb_ctx* ctx, int mode, double tol) {
    API_BEGIN(ctx)
    if (ctx->initialized) fail(PMB_EINITIALIZED, "pmb_set_gram_mode after pmb_initialize");
    if (mode != 0 && mode != 1 && mode != 5 && mode != 6) fail(PMB_EINVAL, "gram mode must be 0 (fp64), 1 (certified int8), 5 or 6 (forced planes)");
    ctx->gram_mode = mode;
    if (tol > 0) ctx->gram_tol = tol;
    API_END(ctx)
}

int pmb_gram_level_counts(pmb_ctx* ctx, int64_t* counts) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_gram_level_counts before pmb_initialize");
    if (!counts) fail(PMB_EINVAL, "null counts");
    counts[0] = counts[1] = counts[2] = 0;
    if (!ctx->gram_counts) ctx->gram_counts = dalloc<unsigned long long>(ctx, 3, true);
    CK(cudaMemsetAsync(ctx->gram_counts, 0, 3 * sizeof(unsigned long long), ctx->stream));
    int64_t fp64_only = 0;
    for (auto& s : ctx->steps) {
        if (s.type != ST_GRAM) continue;
        if (s.i8.slices || s.i8.fused) launch_gram_i8_count_levels(s.i8, s.batch, ctx->gram_counts, ctx->stream);
        else fp64_only += s.batch;
    }
    unsigned long long h[3];
    CK(cudaMemcpyAsync(h, ctx->gram_counts, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    sync_and_repair(ctx);
    counts[0] = (int64_t)h[0]; counts[1] = (int64_t)h[1]; counts[2] = (int64_t)h[2] + fp64_only;
    API_END(ctx)
}

int pmb_enable_kernel_timing(pmb_ctx* ctx, int enable) {
    API_BEGIN(ctx)
    if (!ctx->initialized) fail(PMB_ENOTINIT, "pmb_enable_kernel_timing before pmb_initialize");
    collect_timing(ctx);
    if (enable && ctx->ev.empty()) {
        ctx->ev.resize(2 * ctx->steps.size());
        for (auto& e : ctx->ev) CK(cudaEventCreate(&e));
    }
    if (enable && !ctx->timing) {
        for (int i = 0; i < KC_COUNT; ++i) { ctx->cls_ms[i] = 0; ctx->cls_launches[i] = 0; }
    }
    if (enable && !ctx->graph_exec_timing && ctx->use_graph && !ctx->steps.empty()) {
        // capture the timing graph once: every step in plan order on the context stream, bracketed by external event-record nodes
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        try {
            for (size_t i = 0; i < ctx->steps.size(); ++i) {
                CK(cudaEventRecordWithFlags(ctx->ev[2 * i], ctx->stream, cudaEventRecordExternal));
                run_step(ctx, ctx->steps[i]);
                CK(cudaEventRecordWithFlags(ctx->ev[2 * i + 1], ctx->stream, cudaEventRecordExternal));
            }
        } catch (...) {
            cudaStreamEndCapture(ctx->stream, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        CK(cudaStreamEndCapture(ctx->stream, &g));
        ctx->graph_timing = g;
        if (cudaGraphInstantiate(&ctx->graph_exec_timing, g, 0) != cudaSuccess) { (void)cudaGetLastError(); ctx->graph_exec_timing = nullptr; }
    }
    ctx->timing = enable != 0;
    API_END(ctx)
}

int pmb_kernel_class_count(void) { return KC_COUNT; }
const char* pmb_kernel_class_name(int cls) { return (cls >= 0 && cls < KC_COUNT) ? kClassNames[cls] : ""; }

int pmb_kernel_time_ms(pmb_ctx* ctx, int cls, double* total_ms, int64_t* launches) {
    API_BEGIN(ctx)
    if (cls < 0 || cls >= KC_COUNT) fail(PMB_EINVAL, "bad kernel class");
    collect_timing(ctx);
    if (total_ms) *total_ms = ctx->cls_ms[cls];
    if (launches) *launches = ctx->cls_launches[cls];
    API_END(ctx)
}

int pmb_host_alloc(pmb_ctx* ctx, size_t bytes, void** out) {
    API_BEGIN(ctx)
    if (!out) fail(PMB_EINVAL, "null out");
    void* p = nullptr;
    if (bytes > 0) CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    ctx->pinned.push_back(p);
    *out = p;
    API_END(ctx)
}

int pmb_host_free(pmb_ctx* ctx, void* p) {
    API_BEGIN(ctx)
    auto it = std::find(ctx->pinned.begin(), ctx->pinned.end(), p);
    if (it == ctx->pinned.end()) fail(PMB_EINVAL, "pointer was not allocated by pmb_host_alloc");
    ctx->pinned.erase(it);
    if (p) CK(cudaFreeHost(p));
    API_END(ctx)
}

int pmb_device_alloc(pmb_ctx* ctx, size_t bytes, void** out) {
    API_BEGIN(ctx)
    if (!out) fail(PMB_EINVAL, "null out");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    void* p = nullptr;
    if (bytes > 0) {
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) fail(PMB_ENOMEM, "cudaMalloc of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        ctx->allocs.push_back(p);
        ctx->bytes_alloc += (int64_t)bytes;
    }
    *out = p;
    API_END(ctx)
}

int pmb_device_free(pmb_ctx* ctx, void* p) {
    API_BEGIN(ctx)
    if (!p) return PMB_OK;
    auto it = std::find(ctx->allocs.begin(), ctx->allocs.end(), p);
    if (it == ctx->allocs.end()) fail(PMB_EINVAL, "pointer was not allocated by pmb_device_alloc");
    ctx->allocs.erase(it);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(p));
    API_END(ctx)
}

int pmb_memcpy_d2h(pmb_ctx* ctx, void* host, const void* dev, size_t bytes) {
    API_BEGIN(ctx)
    if (bytes) CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    API_END(ctx)
}

int pmb_memcpy_h2d(pmb_ctx* ctx, void* dev, const void* host, size_t bytes) {
    API_BEGIN(ctx)
    if (bytes) CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    API_END(ctx)
}

}  // extern "C"
