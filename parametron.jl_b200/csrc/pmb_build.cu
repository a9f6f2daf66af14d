// Build phase of the C ABI: the device expression DAG.
//
// Each pmb_op_* entry point is the lowering of one rewrite rule of src/lazyexpression.jl:198-302 onto "linear forms"
// (pmb_internal.h): instead of a preallocated term-list `dest` filled by an in-place kernel of src/functions.jl, a node
// carries the static description of which Parameter buffers feed which coefficient of the dense block.  Nothing is
// evaluated here; pmb_initialize (pmb_plan.cu) turns the forms reachable from the objective / constraints into launches.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "pmb_host.h"

namespace pmb {

void fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Err{code, buf};
}

Node& node_at(pmb_ctx* c, int id) {
    if (id < 0 || id >= (int)c->nodes.size()) fail(PMB_EINVAL, "invalid node handle %d", id);
    return c->nodes[id];
}

int push_node(pmb_ctx* c, Node&& n) {
    c->nodes.push_back(std::move(n));
    return (int)c->nodes.size() - 1;
}

void need_build_phase(pmb_ctx* c, const char* what) {
    // ref: model.jl:50,61,69
    if (c->initialized) fail(PMB_EINITIALIZED, "Model has already been initialized. %s can only be called before initialization.", what);
}

bool is_numeric(int kind) { return kind == PMB_KIND_NUM_SCALAR || kind == PMB_KIND_NUM_VECTOR || kind == PMB_KIND_NUM_MATRIX; }

int64_t numel_of(const Node& n) {
    if (n.bsr_b > 0) return n.bsr_nblocks * n.bsr_b * n.bsr_b;
    if (n.kind == PMB_KIND_NUM_SCALAR) return 1;
    if (n.kind == PMB_KIND_NUM_VECTOR) return n.rows;
    return n.rows * n.cols;
}

static int veclen(pmb_ctx* c, int id) {  // length of any vector-like node
    const Node& n = node_at(c, id);
    switch (n.kind) {
        case PMB_KIND_NUM_VECTOR: case PMB_KIND_VARVEC: case PMB_KIND_AFFVEC: return (int)n.rows;
        default: fail(PMB_EINVAL, "node %d is not a vector", id);
    }
    return 0;
}

// ---- numeric nodes --------------------------------------------------------------------------
int make_const(pmb_ctx* c, int kind, int64_t rows, int64_t cols, std::vector<double>&& vals) {
    Node n;
    n.kind = kind; n.rows = rows; n.cols = cols; n.is_const = true; n.nop = NOP_CONST; n.cvals = std::move(vals);
    return push_node(c, std::move(n));
}

static int make_numeric(pmb_ctx* c, int kind, int64_t rows, int64_t cols, int nop, std::vector<int> in) {
    Node n;
    n.kind = kind; n.rows = rows; n.cols = cols; n.nop = nop; n.in = std::move(in);
    n.is_const = true;
    for (int i : n.in) n.is_const = n.is_const && node_at(c, i).is_const;
    return push_node(c, std::move(n));
}

Alpha alpha_of_scalar(pmb_ctx* c, int id) {
    const Node& n = node_at(c, id);
    if (n.kind != PMB_KIND_NUM_SCALAR) fail(PMB_EINVAL, "node %d is not a numeric scalar", id);
    if (n.nop == NOP_CONST) return Alpha{n.cvals[0], -1};
    return Alpha{1.0, id};
}

Alpha amul(pmb_ctx* c, Alpha a, Alpha b) {
    Alpha r;
    r.c = a.c * b.c;
    if (a.node < 0) r.node = b.node;
    else if (b.node < 0) r.node = a.node;
    else {
        int id = make_numeric(c, PMB_KIND_NUM_SCALAR, 0, 0, NOP_SPROD, {a.node, b.node});
        c->nodes[id].sprod_c = 1.0;
        r.node = id;
    }
    return r;
}

static bool alpha_const(pmb_ctx* c, const Alpha& a) { return a.node < 0 || node_at(c, a.node).is_const; }

// ---- affine forms ---------------------------------------------------------------------------
bool aff_is_const(pmb_ctx* c, const AffForm& f) {
    for (auto& t : f.mats) if (!alpha_const(c, t.a) || !node_at(c, t.mat).is_const) return false;
    for (auto& t : f.units) if (!alpha_const(c, t.a)) return false;
    for (auto& t : f.offs) if (!alpha_const(c, t.a) || !node_at(c, t.vec).is_const) return false;
    for (auto& t : f.sps) if (!alpha_const(c, t.a) || !node_at(c, t.mat).is_const) return false;
    return true;
}

bool quad_is_const(pmb_ctx* c, const QuadForm& q) {
    for (auto& g : q.grams) if (!alpha_const(c, g.a) || !aff_is_const(c, g.f) || !aff_is_const(c, g.h)) return false;
    for (auto& b : q.bilins) if (!alpha_const(c, b.a) || !node_at(c, b.mat).is_const) return false;
    for (auto& t : q.coos) if (!alpha_const(c, t.a)) return false;
    return aff_is_const(c, q.aff);
}

static void aff_scale(pmb_ctx* c, AffForm& f, Alpha m) {
    for (auto& t : f.mats) t.a = amul(c, t.a, m);
    for (auto& t : f.units) t.a = amul(c, t.a, m);
    for (auto& t : f.offs) t.a = amul(c, t.a, m);
    for (auto& t : f.sps) t.a = amul(c, t.a, m);
    if (!f.coff.empty()) {
        if (m.node < 0) {
            for (double& v : f.coff) v *= m.c;
        } else {  // a dynamic multiplier on folded constants: un-fold them into a constant vector source
            std::vector<double> vals = f.coff;
            int cn = make_const(c, PMB_KIND_NUM_VECTOR, f.rows, 0, std::move(vals));
            OffTerm o;
            o.a = m; o.vec = cn; o.src_off = 0; o.row_off = 0; o.rows = f.rows;
            f.offs.push_back(o);
            f.coff.clear();
        }
    }
}

// dst (rows already final) += sign * src placed at row_shift; term order: dst's terms then src's, per row
static void aff_append(AffForm& dst, const AffForm& src, double sign, int64_t row_shift) {
    const int m0 = (int)dst.mats.size(), u0 = (int)dst.units.size();
    for (auto t : src.mats) { t.a.c *= sign; t.row_off += row_shift; dst.mats.push_back(std::move(t)); }
    for (auto t : src.units) { t.a.c *= sign; t.row_off += row_shift; dst.units.push_back(std::move(t)); }
    for (auto t : src.offs) { t.a.c *= sign; t.row_off += row_shift; dst.offs.push_back(t); }
    for (auto t : src.sps) {
        if (row_shift != 0) fail(PMB_EUNSUPPORTED, "vcat of a block-sparse product has no device lowering");
        t.a.c *= sign; dst.sps.push_back(std::move(t)); dst.order_exact = false;
    }
    if (!src.coff.empty()) {
        if (dst.coff.empty()) dst.coff.assign((size_t)dst.rows, 0.0);
        for (int64_t i = 0; i < src.rows; ++i) dst.coff[(size_t)(row_shift + i)] += sign * src.coff[(size_t)i];
    }
    for (auto s : src.order) {
        for (int& i : s.idx) i += s.is_unit ? u0 : m0;
        dst.order.push_back(std::move(s));
    }
    dst.order_exact = dst.order_exact && src.order_exact;
}

// the AffForm meaning of a vector-like / scalar node
AffForm aff_of(pmb_ctx* c, int id) {
    const Node& n = node_at(c, id);
    AffForm f;
    switch (n.kind) {
        case PMB_KIND_AFFVEC: case PMB_KIND_AFF: return *n.aff;
        case PMB_KIND_VARVEC: {
            f.rows = n.rows;
            UnitTerm u; u.rows = n.rows; u.vars = n.vars;
            f.units.push_back(std::move(u));
            f.order.push_back(OrdSeg{1, {0}});
            return f;
        }
        case PMB_KIND_NUM_VECTOR: case PMB_KIND_NUM_SCALAR: {
            f.rows = n.kind == PMB_KIND_NUM_SCALAR ? 1 : n.rows;
            if (n.nop == NOP_CONST) f.coff = n.cvals;
            else { OffTerm o; o.vec = id; o.rows = f.rows; f.offs.push_back(o); }
            return f;
        }
        default: fail(PMB_EINVAL, "node %d (kind %d) has no affine meaning", id, n.kind);
    }
    return f;
}

static int push_aff(pmb_ctx* c, int kind, AffForm&& f) {
    Node n;
    n.kind = kind; n.rows = f.rows; n.cols = 0;
    n.is_const = aff_is_const(c, f);
    n.aff = std::make_shared<AffForm>(std::move(f));
    return push_node(c, std::move(n));
}

QuadForm quad_of(pmb_ctx* c, int id) {
    const Node& n = node_at(c, id);
    if (n.kind == PMB_KIND_QUAD) return *n.quad;
    QuadForm q;
    if (n.kind == PMB_KIND_AFF || n.kind == PMB_KIND_NUM_SCALAR) { q.aff = aff_of(c, id); return q; }
    fail(PMB_EINVAL, "node %d (kind %d) has no scalar quadratic meaning", id, n.kind);
    return q;
}

static void quad_scale(pmb_ctx* c, QuadForm& q, Alpha m) {
    for (auto& g : q.grams) g.a = amul(c, g.a, m);
    for (auto& b : q.bilins) b.a = amul(c, b.a, m);
    for (auto& t : q.coos) t.a = amul(c, t.a, m);
    aff_scale(c, q.aff, m);
}

static void quad_append(QuadForm& dst, const QuadForm& src, double sign) {
    for (auto g : src.grams) { g.a.c *= sign; dst.grams.push_back(std::move(g)); }
    for (auto b : src.bilins) { b.a.c *= sign; dst.bilins.push_back(std::move(b)); }
    for (auto t : src.coos) { t.a.c *= sign; dst.coos.push_back(std::move(t)); }
    if (dst.aff.rows == 0) dst.aff.rows = 1;
    aff_append(dst.aff, src.aff, sign, 0);
}

static int push_quad(pmb_ctx* c, QuadForm&& q) {
    Node n;
    n.kind = PMB_KIND_QUAD;
    if (q.aff.rows == 0) q.aff.rows = 1;
    n.is_const = quad_is_const(c, q);
    n.quad = std::make_shared<QuadForm>(std::move(q));
    return push_node(c, std::move(n));
}

static void check_vars(pmb_ctx* c, const int64_t* v, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (v[i] < 1 || v[i] > c->nvars) fail(PMB_EINVAL, "variable index %lld out of range 1..%lld", (long long)v[i], (long long)c->nvars);
}

// numeric lincomb node  out = sum_t alpha_t * src_t   (all sources share the output's shape)
static int make_lincomb(pmb_ctx* c, int kind, int64_t rows, int64_t cols, std::vector<LinTermH> terms) {
    Node n;
    n.kind = kind; n.rows = rows; n.cols = cols; n.nop = NOP_LINCOMB; n.is_const = true;
    for (auto& t : terms) {
        n.in.push_back(t.src);
        n.is_const = n.is_const && node_at(c, t.src).is_const && alpha_const(c, t.a);
        if (t.a.node >= 0) n.in.push_back(t.a.node);
    }
    n.lin = std::move(terms);
    return push_node(c, std::move(n));
}

static int make_gemv(pmb_ctx* c, int A, int v, int trans, int64_t col_off, int64_t ncols, int64_t v_off) {
    const Node& a = node_at(c, A);
    int64_t outlen = trans ? ncols : a.rows;
    int id = make_numeric(c, PMB_KIND_NUM_VECTOR, outlen, 0, NOP_GEMV, {A, v});
    Node& n = c->nodes[id];
    n.gemv_trans = trans; n.gemv_col_off = col_off; n.gemv_cols = ncols; n.gemv_v_off = v_off;
    return id;
}

static int make_dot(pmb_ctx* c, int x, int64_t xoff, int y, int64_t yoff, int64_t len) {
    int id = make_numeric(c, PMB_KIND_NUM_SCALAR, 0, 0, NOP_DOT, {x, y});
    Node& n = c->nodes[id];
    n.dot_xoff = xoff; n.dot_yoff = yoff; n.dot_n = len;
    return id;
}

static int make_transpose(pmb_ctx* c, int A) {
    const Node& a = node_at(c, A);
    return make_numeric(c, PMB_KIND_NUM_MATRIX, a.cols, a.rows, NOP_TRANSPOSE, {A});
}

// ---- the rewrite rules ------------------------------------------------------------------------
// A * x with A::Matrix.  ref: lazyexpression.jl:200-204 -> matvecmul!, functions.jl:775-822
static void need_dense(pmb_ctx* c, int id, const char* what) {
    if (node_at(c, id).bsr_b > 0) fail(PMB_EUNSUPPORTED, "%s of a block-sparse Parameter has no device lowering", what);
}
static void need_no_sparse(const AffForm& f, const char* what) {
    if (!f.sps.empty()) fail(PMB_EUNSUPPORTED, "%s over a block-sparse product has no device lowering", what);
}

static int op_matvec_matrix(pmb_ctx* c, int A, int x) {
    const Node a = node_at(c, A);  // copy: the node vector may grow below
    const Node& xn = node_at(c, x);
    if (xn.kind == PMB_KIND_NUM_VECTOR) {
        need_dense(c, A, "Matrix * numbers");
        if (a.cols != xn.rows) fail(PMB_EDIM, "DimensionMismatch: A has %lld columns, x has length %lld", (long long)a.cols, (long long)xn.rows);
        return make_gemv(c, A, x, 0, 0, a.cols, 0);
    }
    if (xn.kind != PMB_KIND_VARVEC && xn.kind != PMB_KIND_AFFVEC) fail(PMB_EUNSUPPORTED, "Matrix * (kind %d) has no device lowering", xn.kind);
    if (a.cols != xn.rows)  // functions.jl:780-781
        fail(PMB_EDIM, "DimensionMismatch: A has %lld columns, x has length %lld", (long long)a.cols, (long long)xn.rows);
    if (a.bsr_b > 0) {
        if (xn.kind != PMB_KIND_VARVEC) fail(PMB_EUNSUPPORTED, "block-sparse Matrix * x needs x::Vector{Variable}");
        AffForm y;
        y.rows = a.rows;
        y.order_exact = false;  // the reference would also emit the structural zeros as terms
        SpTerm t; t.mat = A; t.vars = xn.vars;
        y.sps.push_back(std::move(t));
        return push_aff(c, PMB_KIND_AFFVEC, std::move(y));
    }
    AffForm fx = aff_of(c, x);
    need_no_sparse(fx, "Matrix * x");
    AffForm y;
    y.rows = a.rows;
    // Variables / unit terms of x: column gather of A        (functions.jl:790-796 and :806-818 with single-term rows)
    OrdSeg seg; seg.is_unit = 0;
    bool exact = fx.order_exact && fx.mats.empty();
    for (auto& os : fx.order) {
        if (!os.is_unit) continue;
        const UnitTerm& u = fx.units[os.idx[0]];
        MatTerm t;
        t.a = u.a; t.mat = A; t.src_col_off = u.row_off; t.row_off = 0; t.rows = a.rows; t.cols = u.rows; t.vars = u.vars;
        if (u.row_off != 0 || u.rows != a.cols) exact = false;
        seg.idx.push_back((int)y.mats.size());
        y.mats.push_back(std::move(t));
    }
    // matrix terms of x: A * (alpha * M[:, cols]) needs a dense product  P = (A')' * M
    int At = -1;
    for (auto& tm : fx.mats) {
        const Node& mn = node_at(c, tm.mat);
        if (mn.kind != PMB_KIND_NUM_MATRIX) fail(PMB_EUNSUPPORTED, "Matrix * (row-vector form) has no device lowering");
        if (At < 0) At = make_transpose(c, A);
        int P = make_numeric(c, PMB_KIND_NUM_MATRIX, a.rows, tm.cols, NOP_GEMM_TN, {At, tm.mat});
        Node& pn = c->nodes[P];
        pn.gm_xoff = tm.row_off; pn.gm_yoff = tm.src_col_off; pn.gm_k = tm.rows;
        MatTerm t;
        t.a = tm.a; t.mat = P; t.src_col_off = 0; t.row_off = 0; t.rows = a.rows; t.cols = tm.cols; t.vars = tm.vars;
        seg.idx.push_back((int)y.mats.size());
        y.mats.push_back(std::move(t));
    }
    if (exact) {
        if (!seg.idx.empty()) y.order.push_back(std::move(seg));
    } else {
        for (int i : seg.idx) y.order.push_back(OrdSeg{0, {i}});
    }
    y.order_exact = exact;
    // offsets of x: g = A * gx
    for (auto& o : fx.offs) {
        int gv = make_gemv(c, A, o.vec, 0, o.row_off, o.rows, o.src_off);
        OffTerm t; t.a = o.a; t.vec = gv; t.src_off = 0; t.row_off = 0; t.rows = a.rows;
        y.offs.push_back(t);
    }
    if (!fx.coff.empty()) {
        bool any = false;
        for (double v : fx.coff) any = any || v != 0.0;
        if (any) {
            int cv = make_const(c, PMB_KIND_NUM_VECTOR, fx.rows, 0, std::vector<double>(fx.coff));
            int gv = make_gemv(c, A, cv, 0, 0, a.cols, 0);
            OffTerm t; t.vec = gv; t.rows = a.rows;
            y.offs.push_back(t);
        }
    }
    return push_aff(c, PMB_KIND_AFFVEC, std::move(y));
}

// x . y        ref: lazyexpression.jl:228-232 -> vecdot!, functions.jl:665-731
static int op_dot_impl(pmb_ctx* c, int x, int y) {
    const int kx = node_at(c, x).kind, ky = node_at(c, y).kind;
    const int64_t lx = veclen(c, x), ly = veclen(c, y);
    if (lx != ly) fail(PMB_EDIM, "DimensionMismatch: dot of vectors of length %lld and %lld", (long long)lx, (long long)ly);  // functions.jl:669
    if (kx == PMB_KIND_NUM_VECTOR && ky == PMB_KIND_NUM_VECTOR) return make_dot(c, x, 0, y, 0, lx);
    if (kx == PMB_KIND_NUM_VECTOR || ky == PMB_KIND_NUM_VECTOR) {
        // numbers . (variables | affine functions) -> AffineFunction          functions.jl:665-687
        const int v = kx == PMB_KIND_NUM_VECTOR ? x : y, fn = kx == PMB_KIND_NUM_VECTOR ? y : x;
        AffForm f = aff_of(c, fn);
        need_no_sparse(f, "dot");
        AffForm out;
        out.rows = 1;
        out.order_exact = f.mats.empty() && f.units.size() == 1 && f.units[0].row_off == 0 && f.units[0].rows == f.rows;
        for (auto& tm : f.mats) {
            const Node& mn = node_at(c, tm.mat);
            if (mn.kind != PMB_KIND_NUM_MATRIX) fail(PMB_EUNSUPPORTED, "dot over a row-vector form has no device lowering");
            int w = make_gemv(c, tm.mat, v, 1, tm.src_col_off, tm.cols, tm.row_off);
            MatTerm t; t.a = tm.a; t.mat = w; t.src_col_off = 0; t.rows = 1; t.cols = tm.cols; t.vars = tm.vars;
            out.order.push_back(OrdSeg{0, {(int)out.mats.size()}});
            out.mats.push_back(std::move(t));
        }
        for (auto& u : f.units) {
            MatTerm t; t.a = u.a; t.mat = v; t.src_col_off = u.row_off; t.rows = 1; t.cols = u.rows; t.vars = u.vars;
            out.order.push_back(OrdSeg{0, {(int)out.mats.size()}});
            out.mats.push_back(std::move(t));
        }
        for (auto& o : f.offs) {
            int dn = make_dot(c, v, o.row_off, o.vec, o.src_off, o.rows);
            OffTerm t; t.a = o.a; t.vec = dn; t.rows = 1;
            out.offs.push_back(t);
        }
        if (!f.coff.empty()) {
            int cv = make_const(c, PMB_KIND_NUM_VECTOR, f.rows, 0, std::vector<double>(f.coff));
            int dn = make_dot(c, v, 0, cv, 0, f.rows);
            OffTerm t; t.vec = dn; t.rows = 1;
            out.offs.push_back(t);
        }
        return push_aff(c, PMB_KIND_AFF, std::move(out));
    }
    // (variables | affine) . (variables | affine) -> QuadraticFunction      functions.jl:689-709 + muladd! :548-576
    QuadForm q;
    GramTerm g;
    g.f = aff_of(c, x);
    g.h = aff_of(c, y);
    need_no_sparse(g.f, "dot");
    need_no_sparse(g.h, "dot");
    g.same = x == y;
    q.grams.push_back(std::move(g));
    return push_quad(c, std::move(q));
}

static int op_addsub(pmb_ctx* c, int x, int y, double sign) {
    const Node& xn = node_at(c, x);
    const Node& yn = node_at(c, y);
    const int kx = xn.kind, ky = yn.kind;
    if (is_numeric(kx) && is_numeric(ky)) {
        if (kx != ky || xn.rows != yn.rows || xn.cols != yn.cols) fail(PMB_EDIM, "DimensionMismatch in numeric +/-");
        need_dense(c, x, "+/-");
        need_dense(c, y, "+/-");
        return make_lincomb(c, kx, xn.rows, xn.cols, {LinTermH{Alpha{1.0, -1}, x, 0}, LinTermH{Alpha{sign, -1}, y, 0}});
    }
    auto is_vec = [](int k) { return k == PMB_KIND_NUM_VECTOR || k == PMB_KIND_VARVEC || k == PMB_KIND_AFFVEC; };
    if (is_vec(kx) && is_vec(ky)) {  // vecadd!/vecsubtract!, functions.jl:751-764
        if (xn.rows != yn.rows) fail(PMB_EDIM, "DimensionMismatch: vectors of length %lld and %lld", (long long)xn.rows, (long long)yn.rows);  // :755
        AffForm out;
        out.rows = xn.rows;
        aff_append(out, aff_of(c, x), 1.0, 0);
        aff_append(out, aff_of(c, y), sign, 0);
        return push_aff(c, PMB_KIND_AFFVEC, std::move(out));
    }
    auto is_sc = [](int k) { return k == PMB_KIND_NUM_SCALAR || k == PMB_KIND_AFF || k == PMB_KIND_QUAD; };
    if (is_sc(kx) && is_sc(ky)) {  // add!/subtract!, functions.jl:452-502
        if (kx == PMB_KIND_QUAD || ky == PMB_KIND_QUAD) {
            QuadForm out;
            out.aff.rows = 1;
            quad_append(out, quad_of(c, x), 1.0);
            quad_append(out, quad_of(c, y), sign);
            return push_quad(c, std::move(out));
        }
        AffForm out;
        out.rows = 1;
        aff_append(out, aff_of(c, x), 1.0, 0);
        aff_append(out, aff_of(c, y), sign, 0);
        return push_aff(c, PMB_KIND_AFF, std::move(out));
    }
    fail(PMB_EUNSUPPORTED, "+/- of kinds %d and %d has no device lowering", kx, ky);
    return -1;
}

}  // namespace pmb

using namespace pmb;

// =================================================================================================
extern "C" {

int pmb_add_variables(pmb_ctx* ctx, int64_t n, int64_t* first_index) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "Variable(model)");
    if (n < 0) fail(PMB_EINVAL, "negative variable count");
    if (first_index) *first_index = ctx->nvars + 1;
    ctx->nvars += n;
    API_END(ctx)
}

int pmb_num_variables(const pmb_ctx* ctx, int64_t* n) {
    if (!ctx || !n) return PMB_EINVAL;
    *n = ctx->nvars;
    return PMB_OK;
}

static int kind_of_ndims(int ndims, int64_t& rows, int64_t& cols) {
    if (ndims < 0 || ndims > 2) fail(PMB_EINVAL, "ndims must be 0, 1 or 2");
    if (ndims == 0) { rows = 0; cols = 0; return PMB_KIND_NUM_SCALAR; }
    if (ndims == 1) { cols = 0; return PMB_KIND_NUM_VECTOR; }
    return PMB_KIND_NUM_MATRIX;
}

int pmb_param_create(pmb_ctx* ctx, int ndims, int64_t rows, int64_t cols, int mode, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "Parameter(model)");
    if (rows < 0 || cols < 0 || !out) fail(PMB_EINVAL, "bad Parameter shape");
    if (mode != PMB_PARAM_DEVICE && mode != PMB_PARAM_HOST && mode != PMB_PARAM_UNIFORM) fail(PMB_EINVAL, "bad Parameter mode %d", mode);
    Node n;
    n.kind = kind_of_ndims(ndims, rows, cols);
    n.rows = rows; n.cols = cols; n.is_const = false; n.nop = NOP_PARAM; n.pmode = mode;
    *out = push_node(ctx, std::move(n));
    API_END(ctx)
}

int pmb_param_create_bsr(pmb_ctx* ctx, int64_t rows, int64_t cols, int64_t b, int64_t nblocks, const int64_t* block_rows,
                         const int64_t* block_cols, int mode, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "Parameter(model)");
    if (rows < 0 || cols < 0 || b < 1 || nblocks < 0 || !out || (nblocks > 0 && (!block_rows || !block_cols)))
        fail(PMB_EINVAL, "bad block-sparse Parameter shape");
    if (rows % b != 0 || cols % b != 0) fail(PMB_EINVAL, "block size %lld does not divide %lld x %lld", (long long)b, (long long)rows, (long long)cols);
    if (mode != PMB_PARAM_DEVICE && mode != PMB_PARAM_HOST && mode != PMB_PARAM_UNIFORM) fail(PMB_EINVAL, "bad Parameter mode %d", mode);
    Node n;
    n.kind = PMB_KIND_NUM_MATRIX; n.rows = rows; n.cols = cols; n.is_const = false; n.nop = NOP_PARAM; n.pmode = mode;
    n.bsr_b = b; n.bsr_nblocks = nblocks;
    for (int64_t i = 0; i < nblocks; ++i) {
        if (block_rows[i] < 0 || block_rows[i] >= rows / b || block_cols[i] < 0 || block_cols[i] >= cols / b)
            fail(PMB_EINVAL, "block %lld at (%lld, %lld) is outside the matrix", (long long)i, (long long)block_rows[i], (long long)block_cols[i]);
        n.bsr_brow.push_back(block_rows[i]);
        n.bsr_bcol.push_back(block_cols[i]);
    }
    *out = push_node(ctx, std::move(n));
    API_END(ctx)
}

static Node& param_at(pmb_ctx* ctx, pmb_node p, int mode) {
    Node& n = node_at(ctx, p);
    if (n.nop != NOP_PARAM) fail(PMB_EINVAL, "node %d is not a Parameter", p);
    if (mode >= 0 && n.pmode != mode) fail(PMB_EINVAL, "Parameter %d has mode %d, call needs mode %d", p, n.pmode, mode);
    return n;
}

int pmb_param_set_device_ptr(pmb_ctx* ctx, pmb_node p, const double* dptr) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "pmb_param_set_device_ptr");
    param_at(ctx, p, PMB_PARAM_DEVICE).ext = dptr;
    API_END(ctx)
}

int pmb_param_set_host_ptr(pmb_ctx* ctx, pmb_node p, double* hptr, pmb_host_callback f, void* user) {
    API_BEGIN(ctx)
    Node& n = param_at(ctx, p, PMB_PARAM_HOST);
    // the host buffer may be re-pointed between updates unless it was baked into the update graph as a memcpy node
    if (ctx->initialized && ctx->h2d_in_graph && hptr != n.hptr)
        fail(PMB_EINITIALIZED, "Model has already been initialized. The pinned host buffer of a Parameter is fixed by pmb_initialize.");
    n.hptr = hptr; n.cb = f; n.user = user;
    API_END(ctx)
}

int pmb_param_set_uniform(pmb_ctx* ctx, pmb_node p, uint64_t seed, double scale) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "pmb_param_set_uniform");
    Node& n = param_at(ctx, p, PMB_PARAM_UNIFORM);
    n.seed = seed; n.scale = scale;
    API_END(ctx)
}

int pmb_set_instance_offset(pmb_ctx* ctx, int64_t first) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "pmb_set_instance_offset");
    if (first < 0 || first + ctx->batch > ((int64_t)1 << 32)) fail(PMB_EINVAL, "instance offset %lld out of range (global instance ids are 32-bit)", (long long)first);
    ctx->inst_offset = first;
    API_END(ctx)
}

int pmb_get_instance_offset(const pmb_ctx* ctx, int64_t* first) {
    if (!ctx || !first) return PMB_EINVAL;
    *first = ctx->inst_offset;
    return PMB_OK;
}

int pmb_const_numeric(pmb_ctx* ctx, int ndims, int64_t rows, int64_t cols, const double* host_values, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "constant");
    if (rows < 0 || cols < 0 || !out) fail(PMB_EINVAL, "bad constant shape");
    int kind = kind_of_ndims(ndims, rows, cols);
    int64_t ne = kind == PMB_KIND_NUM_SCALAR ? 1 : (kind == PMB_KIND_NUM_VECTOR ? rows : rows * cols);
    if (ne > 0 && !host_values) fail(PMB_EINVAL, "null constant values");
    *out = make_const(ctx, kind, rows, cols, std::vector<double>(host_values, host_values + ne));
    API_END(ctx)
}

int pmb_varvec(pmb_ctx* ctx, int64_t n, const int64_t* var_indices, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "Vector{Variable}");
    if (n < 0 || !out || (n > 0 && !var_indices)) fail(PMB_EINVAL, "bad variable vector");
    check_vars(ctx, var_indices, n);
    Node nd;
    nd.kind = PMB_KIND_VARVEC; nd.rows = n; nd.is_const = true;
    nd.vars.assign(var_indices, var_indices + n);
    *out = push_node(ctx, std::move(nd));
    API_END(ctx)
}

int pmb_const_affine(pmb_ctx* ctx, int64_t nterms, const double* coeffs, const int64_t* vars, double constant, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "AffineFunction");
    if (nterms < 0 || !out || (nterms > 0 && (!coeffs || !vars))) fail(PMB_EINVAL, "bad affine function");
    check_vars(ctx, vars, nterms);
    AffForm f;
    f.rows = 1;
    if (nterms > 0) {
        int cv = make_const(ctx, PMB_KIND_NUM_VECTOR, nterms, 0, std::vector<double>(coeffs, coeffs + nterms));
        MatTerm t; t.mat = cv; t.rows = 1; t.cols = nterms; t.vars.assign(vars, vars + nterms);
        f.mats.push_back(std::move(t));
        f.order.push_back(OrdSeg{0, {0}});
    }
    f.coff.assign(1, constant);
    *out = push_aff(ctx, PMB_KIND_AFF, std::move(f));
    API_END(ctx)
}

int pmb_const_quadratic(pmb_ctx* ctx, int64_t nquad, const double* qcoeffs, const int64_t* rowvars, const int64_t* colvars,
                        int64_t nlin, const double* lcoeffs, const int64_t* lvars, double constant, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "QuadraticFunction");
    if (nquad < 0 || nlin < 0 || !out) fail(PMB_EINVAL, "bad quadratic function");
    check_vars(ctx, rowvars, nquad);
    check_vars(ctx, colvars, nquad);
    check_vars(ctx, lvars, nlin);
    QuadForm q;
    if (nquad > 0) {
        CooTerm t;
        t.i.assign(rowvars, rowvars + nquad);
        t.j.assign(colvars, colvars + nquad);
        t.c.assign(qcoeffs, qcoeffs + nquad);
        q.coos.push_back(std::move(t));
    }
    q.aff.rows = 1;
    if (nlin > 0) {
        int cv = make_const(ctx, PMB_KIND_NUM_VECTOR, nlin, 0, std::vector<double>(lcoeffs, lcoeffs + nlin));
        MatTerm t; t.mat = cv; t.rows = 1; t.cols = nlin; t.vars.assign(lvars, lvars + nlin);
        q.aff.mats.push_back(std::move(t));
        q.aff.order.push_back(OrdSeg{0, {0}});
    }
    q.aff.coff.assign(1, constant);
    *out = push_quad(ctx, std::move(q));
    API_END(ctx)
}

int pmb_op_adjoint(pmb_ctx* ctx, pmb_node A, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    const Node& a = node_at(ctx, A);
    if (a.kind == PMB_KIND_NUM_MATRIX) {  // lazyexpression.jl:206-217: explicit transposed copy into a buffer
        need_dense(ctx, A, "adjoint");
        *out = make_transpose(ctx, A);
    } else if (a.kind == PMB_KIND_NUM_VECTOR || a.kind == PMB_KIND_VARVEC || a.kind == PMB_KIND_AFFVEC) {
        Node n;
        n.kind = PMB_KIND_ADJVEC; n.rows = 1; n.cols = a.rows; n.is_const = a.is_const; n.adj_parent = A;
        *out = push_node(ctx, std::move(n));
    } else if (a.kind == PMB_KIND_ADJVEC) {
        *out = a.adj_parent;
    } else {
        fail(PMB_EUNSUPPORTED, "adjoint of kind %d has no device lowering", a.kind);
    }
    API_END(ctx)
}

int pmb_op_matvec(pmb_ctx* ctx, pmb_node A, pmb_node x, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    const Node& a = node_at(ctx, A);
    if (a.kind == PMB_KIND_ADJVEC) {  // Adjoint{vector} * vector -> vecdot!   functions.jl:824-829
        *out = op_dot_impl(ctx, a.adj_parent, x);
    } else if (a.kind == PMB_KIND_NUM_MATRIX) {
        *out = op_matvec_matrix(ctx, A, x);
    } else {
        fail(PMB_EUNSUPPORTED, "(kind %d) * vector has no device lowering", a.kind);
    }
    API_END(ctx)
}

int pmb_op_dot(pmb_ctx* ctx, pmb_node x, pmb_node y, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    *out = op_dot_impl(ctx, x, y);
    API_END(ctx)
}

int pmb_op_bilinear(pmb_ctx* ctx, pmb_node xT, pmb_node Q, pmb_node y, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    const Node& xa = node_at(ctx, xT);
    const Node& qn = node_at(ctx, Q);
    const Node& yn = node_at(ctx, y);
    if (xa.kind != PMB_KIND_ADJVEC || node_at(ctx, xa.adj_parent).kind != PMB_KIND_VARVEC || qn.kind != PMB_KIND_NUM_MATRIX ||
        yn.kind != PMB_KIND_VARVEC)
        fail(PMB_EUNSUPPORTED, "x' * Q * y needs Variables' * Matrix * Variables (lazyexpression.jl:219-226)");
    need_dense(ctx, Q, "x' * Q * y");
    const Node& xn = node_at(ctx, xa.adj_parent);
    if (qn.rows != xn.rows || qn.cols != yn.rows)  // functions.jl:845
        fail(PMB_EDIM, "DimensionMismatch: size(Q) = (%lld, %lld), length(x) = %lld, length(y) = %lld", (long long)qn.rows, (long long)qn.cols,
             (long long)xn.rows, (long long)yn.rows);
    QuadForm q;
    BilinTerm b;
    b.mat = Q; b.xv = xn.vars; b.yv = yn.vars;
    q.bilins.push_back(std::move(b));
    *out = push_quad(ctx, std::move(q));
    API_END(ctx)
}

int pmb_op_add(pmb_ctx* ctx, pmb_node x, pmb_node y, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    *out = op_addsub(ctx, x, y, 1.0);
    API_END(ctx)
}

int pmb_op_sub(pmb_ctx* ctx, pmb_node x, pmb_node y, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    *out = op_addsub(ctx, x, y, -1.0);
    API_END(ctx)
}

int pmb_op_mul(pmb_ctx* ctx, pmb_node x, pmb_node y, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    int kx = node_at(ctx, x).kind, ky = node_at(ctx, y).kind;
    if (kx == PMB_KIND_NUM_MATRIX && (ky == PMB_KIND_NUM_VECTOR || ky == PMB_KIND_VARVEC || ky == PMB_KIND_AFFVEC)) {
        *out = op_matvec_matrix(ctx, x, y);
    } else if (kx == PMB_KIND_ADJVEC) {
        *out = op_dot_impl(ctx, node_at(ctx, x).adj_parent, y);
    } else if (kx == PMB_KIND_AFF && ky == PMB_KIND_AFF) {  // mul!(::QuadraticFunction, ::AffineFunction, ::AffineFunction) functions.jl:548-578
        QuadForm q;
        GramTerm g;
        g.f = aff_of(ctx, x);
        g.h = aff_of(ctx, y);
        g.same = x == y;
        q.grams.push_back(std::move(g));
        *out = push_quad(ctx, std::move(q));
    } else if (kx == PMB_KIND_NUM_SCALAR || ky == PMB_KIND_NUM_SCALAR) {
        const int s = kx == PMB_KIND_NUM_SCALAR ? x : y, o = kx == PMB_KIND_NUM_SCALAR ? y : x;
        const Node on = node_at(ctx, o);
        const Alpha m = alpha_of_scalar(ctx, s);
        if (is_numeric(on.kind)) {
            if (on.kind == PMB_KIND_NUM_SCALAR) {
                Alpha r = amul(ctx, m, alpha_of_scalar(ctx, o));
                if (r.node < 0) *out = make_const(ctx, PMB_KIND_NUM_SCALAR, 0, 0, {r.c});
                else if (r.c == 1.0) *out = r.node;
                else {
                    int id = make_numeric(ctx, PMB_KIND_NUM_SCALAR, 0, 0, NOP_SPROD, {r.node});
                    ctx->nodes[id].sprod_c = r.c;
                    *out = id;
                }
            } else {
                need_dense(ctx, o, "scalar * Matrix");
                *out = make_lincomb(ctx, on.kind, on.rows, on.cols, {LinTermH{m, o, 0}});
            }
        } else if (on.kind == PMB_KIND_VARVEC || on.kind == PMB_KIND_AFFVEC || on.kind == PMB_KIND_AFF) {
            // scale! functions.jl:873-925 ; mul!(::AffineFunction, ...) :578
            AffForm f = aff_of(ctx, o);
            aff_scale(ctx, f, m);
            *out = push_aff(ctx, on.kind == PMB_KIND_AFF ? PMB_KIND_AFF : PMB_KIND_AFFVEC, std::move(f));
        } else if (on.kind == PMB_KIND_QUAD) {
            QuadForm q = quad_of(ctx, o);
            quad_scale(ctx, q, m);
            *out = push_quad(ctx, std::move(q));
        } else {
            fail(PMB_EUNSUPPORTED, "scalar * (kind %d) has no device lowering", on.kind);
        }
    } else {
        fail(PMB_EUNSUPPORTED, "(kind %d) * (kind %d) has no device lowering", kx, ky);
    }
    API_END(ctx)
}

int pmb_op_vcat(pmb_ctx* ctx, int32_t n, const pmb_node* parts, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (n < 0 || !out || (n > 0 && !parts)) fail(PMB_EINVAL, "bad vcat");
    AffForm f;  // vcat!, functions.jl:969-994
    for (int i = 0; i < n; ++i) f.rows += veclen(ctx, parts[i]);
    int64_t off = 0;
    for (int i = 0; i < n; ++i) {
        AffForm p = aff_of(ctx, parts[i]);
        aff_append(f, p, 1.0, off);
        off += p.rows;
    }
    *out = push_aff(ctx, PMB_KIND_AFFVEC, std::move(f));
    API_END(ctx)
}

int pmb_op_vect(pmb_ctx* ctx, pmb_node x, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    int k = node_at(ctx, x).kind;  // lazyexpression.jl:292-298
    if (k != PMB_KIND_AFF && k != PMB_KIND_NUM_SCALAR) fail(PMB_EUNSUPPORTED, "[x] needs a scalar affine value");
    AffForm f = aff_of(ctx, x);
    *out = push_aff(ctx, PMB_KIND_AFFVEC, std::move(f));
    API_END(ctx)
}

int pmb_op_convert(pmb_ctx* ctx, pmb_node x, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    // convert(Vector, x) -> copyto!(dest, x), lazyexpression.jl:280-282: any vector value becomes a Vector{AffineFunction}
    const int k = node_at(ctx, x).kind;
    if (k != PMB_KIND_NUM_VECTOR && k != PMB_KIND_VARVEC && k != PMB_KIND_AFFVEC) fail(PMB_EUNSUPPORTED, "convert(Vector, x) needs a vector value");
    AffForm f = aff_of(ctx, x);
    *out = push_aff(ctx, PMB_KIND_AFFVEC, std::move(f));
    API_END(ctx)
}

int pmb_op_getfield(pmb_ctx* ctx, pmb_node p, int64_t offset, int ndims, int64_t rows, int64_t cols, pmb_node* out) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "@expression");
    if (!out) fail(PMB_EINVAL, "null out");
    // getproperty(p, :field) -> GetField{field}()(p), lazyexpression.jl:300-302, functions.jl:998-999.  A struct-valued Parameter is a
    // flat numeric record on the device (fields packed in declaration order, matrices column-major); the field is the sub-array at
    // element `offset`, copied into its own node by the numeric prologue (one lincomb term, alpha = 1).
    const Node& pn = node_at(ctx, p);
    if (pn.kind != PMB_KIND_NUM_VECTOR || pn.bsr_b > 0) fail(PMB_EUNSUPPORTED, "getfield needs a record (flat vector) value");
    int64_t r = rows, cc = cols;
    const int kind = kind_of_ndims(ndims, r, cc);
    const int64_t numel = kind == PMB_KIND_NUM_SCALAR ? 1 : kind == PMB_KIND_NUM_VECTOR ? r : r * cc;
    if (offset < 0 || offset + numel > pn.rows) fail(PMB_EDIM, "field [%lld, %lld) outside a record of %lld values", (long long)offset,
                                                     (long long)(offset + numel), (long long)pn.rows);
    *out = make_lincomb(ctx, kind, r, cc, {LinTermH{Alpha{1.0, -1}, p, offset}});
    API_END(ctx)
}

int pmb_node_info(pmb_ctx* ctx, pmb_node n, int32_t* kind, int64_t* rows, int64_t* cols, int32_t* is_constant) {
    API_BEGIN(ctx)
    const Node& nd = node_at(ctx, n);
    if (kind) *kind = nd.kind;
    if (rows) *rows = nd.rows;
    if (cols) *cols = nd.cols;
    if (is_constant) *is_constant = nd.is_const ? 1 : 0;
    API_END(ctx)
}

int pmb_set_objective(pmb_ctx* ctx, int sense, pmb_node expr) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "setobjective!");
    if (sense != PMB_MINIMIZE && sense != PMB_MAXIMIZE) fail(PMB_EINVAL, "bad objective sense");
    const Node& n = node_at(ctx, expr);
    if (n.kind != PMB_KIND_AFF && n.kind != PMB_KIND_QUAD && n.kind != PMB_KIND_NUM_SCALAR)
        fail(PMB_EUNSUPPORTED, "Unhandled objective type (kind %d)", n.kind);  // model.jl:139
    if (n.kind != PMB_KIND_QUAD) need_no_sparse(aff_of(ctx, expr), "an objective");
    ctx->obj_node = expr;
    ctx->obj_sense = sense;
    ctx->obj_kind = n.kind == PMB_KIND_QUAD ? PMB_KIND_QUAD : PMB_KIND_AFF;  // canonical_function_type, moi_interop.jl:96-101
    ctx->obj_const = n.is_const;
    API_END(ctx)
}

int pmb_add_constraint(pmb_ctx* ctx, pmb_node expr, int set_kind, int32_t* block_id) {
    API_BEGIN(ctx)
    need_build_phase(ctx, "add_constraint");
    if (set_kind < PMB_SET_NONNEGATIVE || set_kind > PMB_SET_ZERO) fail(PMB_EINVAL, "bad constraint set %d", set_kind);
    const Node& n = node_at(ctx, expr);
    Constraint con;
    con.node = expr; con.set_kind = set_kind; con.is_const = n.is_const;
    if (n.kind == PMB_KIND_AFFVEC || n.kind == PMB_KIND_VARVEC) {  // model.jl:88-95 vector sets
        con.is_scalar = false; con.bucket = 3 + set_kind;
    } else if (n.kind == PMB_KIND_AFF) {  // model.jl:79-86 scalar sets
        con.is_scalar = true; con.bucket = 0 + set_kind;
    } else if (n.kind == PMB_KIND_QUAD) {  // model.jl:79-86 with a QuadraticFunction value: ScalarQuadraticFunction-in-set
        con.is_scalar = true; con.bucket = 6 + set_kind; con.is_quad = true;
        con.quad = std::make_shared<QuadForm>(quad_of(ctx, expr));
        con.rows = 1;
        ctx->cons.push_back(std::move(con));
        if (block_id) *block_id = (int32_t)ctx->cons.size() - 1;
        return PMB_OK;
    } else {
        fail(PMB_EUNSUPPORTED, "constraint on a value of kind %d has no device lowering", n.kind);
    }
    con.form = aff_of(ctx, expr);
    con.rows = con.form.rows;
    if (!con.form.sps.empty()) {
        if (con.form.sps.size() != 1 || !con.form.mats.empty() || !con.form.units.empty())
            fail(PMB_EUNSUPPORTED, "a block-sparse product must be the only coefficient source of its constraint block");
        con.sparse = true;
    }
    ctx->cons.push_back(std::move(con));
    if (block_id) *block_id = (int32_t)ctx->cons.size() - 1;
    API_END(ctx)
}

}  // extern "C"
