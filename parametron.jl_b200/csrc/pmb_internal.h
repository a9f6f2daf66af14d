// Host-side data model of the device expression DAG (the restatement of src/lazyexpression.jl's LazyExpression tree
// and of the term-list value types of src/functions.jl as static "linear forms" over Parameter buffers).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/parametron_b200.h"
#include "pmb_kernels.h"

namespace pmb {

// alpha = c * (node >= 0 ? value of scalar numeric node : 1)
struct Alpha {
    double c = 1.0;
    int node = -1;
};

// G[row_off + i, var(vars[j])] += alpha * M[i, src_col_off + j]     i < rows, j < cols      (matvecmul!, functions.jl:775-798)
struct MatTerm {
    Alpha a;
    int mat = -1;
    int64_t src_col_off = 0, row_off = 0, rows = 0, cols = 0;
    std::vector<int64_t> vars;  // 1-based model variable indices, length cols
};
// G[row_off + i, var(vars[i])] += alpha                                               (Vector{Variable} operand)
struct UnitTerm {
    Alpha a;
    int64_t row_off = 0, rows = 0;
    std::vector<int64_t> vars;
};
// g[row_off + i] += alpha * v[src_off + i]                                             (add!/subtract! of numbers)
struct OffTerm {
    Alpha a;
    int vec = -1;
    int64_t src_off = 0, row_off = 0, rows = 0;
};
// block-sparse source (BSR values buffer):  G[row(e), var(vars[col(e)])] += alpha * vals[e]  for every stored coefficient e
// (the dense meaning of matvecmul! over a matrix whose other entries are structural zeros, functions.jl:775-798)
struct SpTerm {
    Alpha a;
    int mat = -1;               // BSR Parameter node
    std::vector<int64_t> vars;  // 1-based model variable per matrix column
};
struct OrdSeg {
    int is_unit = 0;
    std::vector<int> idx;  // indices into mats (members share row_off/rows/cols) or a single index into units
};
// Vector{AffineFunction} (or a scalar AffineFunction when rows == 1 && scalar)
struct AffForm {
    int64_t rows = 0;
    std::vector<MatTerm> mats;
    std::vector<UnitTerm> units;
    std::vector<OffTerm> offs;
    std::vector<SpTerm> sps;                   // block-sparse sources (a block with one of these is emitted in sparse form)
    std::vector<double> coff;                  // folded constant offsets (size rows, or empty == zeros)
    // Reference term-append order, per row: segments are visited in order; a matrix segment with several members emits
    // its members interleaved per source column (matvecmul! over Vector{AffineFunction}, functions.jl:800-822).
    std::vector<OrdSeg> order;
    bool order_exact = true;                   // false when the reference's per-row term order is not expressible by `order`
};
struct GramTerm {  // alpha * sum_r f_r * h_r        (vecdot! on two affine vectors, functions.jl:702-709)
    Alpha a;
    AffForm f, h;
    bool same = false;
};
struct BilinTerm {  // alpha * x' Q y                 (bilinearmul!, functions.jl:840-858)
    Alpha a;
    int mat = -1;
    std::vector<int64_t> xv, yv;
};
struct CooTerm {  // alpha * sum_e c_e * x_{i_e} * x_{j_e}   (folded constant QuadraticTerms)
    Alpha a;
    std::vector<int64_t> i, j;
    std::vector<double> c;
};
struct QuadForm {
    std::vector<GramTerm> grams;
    std::vector<BilinTerm> bilins;
    std::vector<CooTerm> coos;
    AffForm aff;  // rows == 1
};

enum NumOp { NOP_NONE = -1, NOP_PARAM, NOP_CONST, NOP_TRANSPOSE, NOP_LINCOMB, NOP_GEMV, NOP_DOT, NOP_SPROD, NOP_GEMM_TN };

struct LinTermH {
    Alpha a;
    int src;
    int64_t src_off;
};

struct Node {
    int kind = -1;
    int64_t rows = 0, cols = 0;  // numeric: rows x cols (vector: cols == 0; scalar: rows == cols == 0)
    bool is_const = true;
    // numeric nodes
    int nop = NOP_NONE;
    std::vector<int> in;
    std::vector<LinTermH> lin;
    std::vector<double> cvals;
    int gemv_trans = 0;
    int64_t gemv_col_off = 0, gemv_cols = 0, gemv_row_off = 0, gemv_rows = 0, gemv_v_off = 0;
    int64_t dot_xoff = 0, dot_yoff = 0, dot_n = 0;
    int64_t gm_xoff = 0, gm_yoff = 0, gm_k = 0;  // NOP_GEMM_TN: node (rows x cols) = X[xoff.., :]' * Y[:, yoff..] over gm_k rows
    double sprod_c = 1.0;
    // block-sparse (BSR) matrices: value buffer = nblocks x (b x b column-major); static block coordinates
    int64_t bsr_b = 0, bsr_nblocks = 0;
    std::vector<int64_t> bsr_brow, bsr_bcol;   // 0-based block row / block column of each stored block
    // parameters
    int pmode = -1;
    uint64_t seed = 0;
    double scale = 1.0;
    double* hptr = nullptr;
    pmb_host_callback cb = nullptr;
    void* user = nullptr;
    const double* ext = nullptr;
    // device storage
    double* dptr = nullptr;
    int64_t numel = 0, bstride = 0;
    bool owned = false, needed = false;
    const pmb::FillJob* lazy_fill = nullptr;   // UNIFORM Parameter generated inside the int8 slice pass and not written by updates: the one-job list pmb_param_device_ptr runs
    int lazy_fill_blocks = 0;
    bool used_main = false;   // read by a numeric node or the objective chain (else: constraint chain only)
    // symbolic nodes
    std::vector<int64_t> vars;
    int adj_parent = -1;
    std::shared_ptr<AffForm> aff;
    std::shared_ptr<QuadForm> quad;
};

enum KernelClass { KC_AFFINE = 0, KC_GRAM, KC_GEMM, KC_PHILOX, KC_H2D, KC_REDUCE, KC_MISC, KC_PACK, KC_GRAM_I8, KC_GRAM_SLICE, KC_COUNT };

enum StepType {
    ST_MEMSET, ST_LINCOMB, ST_TRANSPOSE, ST_GEMV, ST_DOT, ST_SPROD, ST_AFFINE, ST_GRAM, ST_GEMM,
    ST_SCATTER_VEC, ST_BILINEAR, ST_COO, ST_SCATTER_SYM, ST_PROLOGUE, ST_GEMV_DOT, ST_PACK_P, ST_PACK_A, ST_GRAM_I8_SLICE, ST_GRAM_I8_MMA, ST_GRAM_I8_FUSED
};

struct Step {
    int type = 0, cls = KC_MISC;
    int phase = 1;  // 0: prologue, 1: numeric nodes, 2: objective chain (main branch), 3: constraint chain (aux branch)
    int64_t batch = 1;
    // generic pointer/size slots
    double* out = nullptr;
    const double* src = nullptr;
    int64_t n = 0, n2 = 0, out_bstride = 0, src_bstride = 0;
    uint64_t seed = 0;
    double scale = 1.0;
    const LinCombTermDev* lterms = nullptr;
    int nl = 0;
    const double* cst = nullptr;
    const double* s0 = nullptr; int64_t s0_bs = 0;
    const double* s1 = nullptr; int64_t s1_bs = 0;
    AffBlockDev aff{};
    GramScratch gscratch{};
    const FillJob* jobs = nullptr; int njobs = 0, job_blocks = 0, tick = 1;  // ST_PROLOGUE (tick = 0: the constraint branch's generators, no update tick)  // ST_GRAM / ST_GEMM: k-split plan, arrival counters and partial-tile workspace
    GramArgs gram{};
    GramI8Scratch i8{};   // ST_GRAM: certified int8-slice path (slices == nullptr: fp64 only)
    bool pipe8 = false;   // this Gram's operand is uploaded in column chunks; slice + tcgen05 launches run per chunk (build_update_graph)
    GemmTNArgs gemm{};
    GemvArgs gemv{};
    DotArgs dot{};
    ScatterVecArgs sv{};
    BilinearArgs bl{};
    CooQuadArgs coo{};
    ScatterSymArgs ss{};
};

struct Constraint {
    int node = -1;
    int set_kind = 0;
    int bucket = 0;
    bool is_scalar = false, is_const = false;
    int64_t rows = 0;
    double* C = nullptr;
    double* d = nullptr;
    int64_t nterms = 0;
    // scalar quadratic constraint: (Qc, C = q as a 1 x N row, d = c0)
    bool is_quad = false;
    std::shared_ptr<pmb::QuadForm> quad;
    double* Qc = nullptr;
    // sparse form (the block's only coefficient source is one BSR matrix)
    bool sparse = false;
    int64_t nnz = 0;
    double* vals = nullptr;
    int32_t* sp_rows = nullptr;
    int32_t* sp_cols = nullptr;
    // term-list export (built lazily by pmb_block_export_terms)
    bool export_ready = false;
    ExportArgs ex{};
    const TermGatherDev* ex_gather = nullptr;   // rows with different term layouts: one gather entry per term instead of the per-row template
    AffForm form;
};

}  // namespace pmb

struct pmb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t batch = 1;
    int64_t inst_offset = 0;   // global id of local instance 0 (instance-shard: pmb_set_instance_offset)
    int64_t nvars = 0;
    bool initialized = false;
    std::string err;
    std::vector<pmb::Node> nodes;
    // objective
    int obj_node = -1;
    int obj_sense = PMB_MINIMIZE;
    bool obj_const = true;
    int obj_kind = PMB_KIND_AFF;
    double *Q = nullptr, *q = nullptr, *c0 = nullptr;
    std::vector<pmb::Constraint> cons;   // insertion order
    std::vector<int> update_order;       // indices into cons, bucket order then insertion order
    std::vector<int64_t> varmap;         // model var (1-based) -> optimizer var (1-based)
    // plan
    std::vector<pmb::Step> steps;        // one update
    std::vector<pmb::Step> init_steps;   // constants, run once
    std::vector<void*> pinned;           // pmb_host_alloc blocks still owned by the context
    std::vector<void*> allocs;
    int64_t bytes_alloc = 0;
    unsigned long long* d_counter = nullptr;
    unsigned int* d_ticket = nullptr;
    std::vector<pmb::FillJob> fill_jobs;
    std::vector<int> fill_job_nodes;     // Parameter node of each fill job (-1: the zeroing job); decides the branch of the job (used_main)
    // two-branch update graph: objective chain on the context stream, constraint chain on an auxiliary stream
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork0 = nullptr, ev_fork1 = nullptr, ev_join = nullptr;
    bool h2d_in_graph = false;
    // upload/compute overlap of the objective's Gram operand (a pinned HOST Parameter used by nothing else): the matrix travels in
    // column chunks on its own stream while the flag-gated Gram kernel already consumes the chunks that have landed
    int pipe_param = -1;
    int32_t* pipe_gate = nullptr;        // device flags, one per chunk
    int32_t* pipe_ones = nullptr;        // pinned host array of ones (source of the flag copies)
    int32_t pipe_chunks = 0, pipe_gate_cols = 0;
    cudaStream_t h2d_stream = nullptr;
    // int8 variant of the same overlap (pipe8): per arrived chunk one slice launch + one ranged tcgen05 Gram launch, and the finished
    // columns of the packed P_x leave on a third stream while later chunks are still arriving.  The pipelined graph has no fp64
    // launch: if the certificate refuses an instance, a flag reaches the host and pmb_synchronize replays a plain fp64 graph.
    bool pipe8 = false;
    std::vector<int64_t> pipe8_bounds;   // column boundaries of the chunks: 0 = b[0] < b[1] < ... = n, multiples of 480 = lcm(96, 80)
    cudaStream_t d2h_stream = nullptr;
    std::vector<cudaEvent_t> ev_chunk, ev_gram;
    cudaEvent_t ev_d2h_done = nullptr;
    int32_t* repair_flag_dev = nullptr;
    int32_t* repair_flag_host = nullptr;
    cudaGraph_t graph_fb = nullptr;
    cudaGraphExec_t graph_exec_fb = nullptr;
    bool fb_mode = false;                // run_step: fp64 replay (skip the tcgen05 launches, the fp64 kernel computes every instance)
    bool pipe8_live = false;             // the current update graph is the pipelined one
    int64_t repairs = 0;
    cudaEvent_t ev_h2d_done = nullptr, ev_aux_h2d = nullptr, ev_aux_kernels = nullptr;           // HOST Parameters are copied by memcpy nodes of the graph (pinned, pointer fixed at initialize)
    // host sinks of the packed QP (pmb_qp_set_host_sink): D2H memcpy nodes at the end of each branch
    double *sink_Px = nullptr, *sink_q = nullptr, *sink_c0 = nullptr, *sink_Ax = nullptr, *sink_l = nullptr, *sink_u = nullptr;
    // solver-side ingest (pmb_enable_qp_pack): packed upper-triangular P, stacked CSC A, row bounds
    bool qp_pack = false;
    bool qp_P_fused = false;   // P_x is written by the Gram epilogue (no pack launch; host sink written directly)
    int64_t qp_m = 0, qp_nnzP = 0, qp_nnzA = 0;
    double *qp_Px = nullptr, *qp_Ax = nullptr, *qp_l = nullptr, *qp_u = nullptr;
    int64_t *qp_Pp = nullptr, *qp_Pi = nullptr, *qp_Ap = nullptr, *qp_Ai = nullptr;
    pmb::PackBlockDev* qp_blocks_dev = nullptr;  // collected while planning, lowered to one ST_PROLOGUE step
    int64_t updates = 0;
    // result path (pmb_set_primal / pmb_value / pmb_objective_value)
    double* primal = nullptr;            // [batch][N], optimizer variable order
    bool primal_set = false;
    double* res_tmp = nullptr;           // [batch][N] scratch (Q x) + [batch] result
    int64_t* res_idx = nullptr; int64_t res_idx_cap = 0;   // gather table of pmb_value
    double* res_gather = nullptr; int64_t res_gather_cap = 0;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraph_t graph_timing = nullptr;            // pmb_enable_kernel_timing: the steps serialised with event-record nodes around each launch
    cudaGraphExec_t graph_exec_timing = nullptr;
    bool use_graph = false;
    bool use_tma = true;
    bool trace = false;                  // PMB_TRACE was set when the context was created (timeline of the pipelined schedule)
    // Gram precision mode (pmb_set_gram_mode): 0 = fp64 DMMA only, 1 = certified int8 slices where they pay off (default),
    // 5 / 6 = force that many digit planes on every SYRK step (experiments: no certificate, no fp64 fallback)
    int gram_mode = 1;
    double gram_tol = 9e-11;   // worst-case bound; the fp64 reference itself rounds at ~1e-13, so 9/10 of the 1e-10 budget is the certificate's
    unsigned long long* fused_diag = nullptr;    // host-mapped: the fused int8 kernel records here which wait timed out (GramI8Scratch::diag)
    unsigned long long* gram_counts = nullptr;   // device: instances per level {5 planes, 6 planes, fp64} of the last counted update
    bool timing = false;
    std::vector<cudaEvent_t> ev;         // 2 per step
    bool ev_pending = false;
    double cls_ms[pmb::KC_COUNT] = {0};
    int64_t cls_launches[pmb::KC_COUNT] = {0};
    int64_t launches_per_update = 0, memcpy_per_update = 0;
    int64_t affine_bytes = 0, gram_flops = 0;
};
