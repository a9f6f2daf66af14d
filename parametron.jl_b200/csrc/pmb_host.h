// Shared host-side helpers of the C ABI implementation (pmb_build.cu: DAG builder, pmb_plan.cu: launch planner).
#pragma once
#include <new>
#include <string>

#include "pmb_internal.h"

namespace pmb {

struct Err {
    int code;
    std::string msg;
};

[[noreturn]] void fail(int code, const char* fmt, ...);
Node& node_at(pmb_ctx* c, int id);
int push_node(pmb_ctx* c, Node&& n);
void need_build_phase(pmb_ctx* c, const char* what);
bool is_numeric(int kind);
int64_t numel_of(const Node& n);
int make_const(pmb_ctx* c, int kind, int64_t rows, int64_t cols, std::vector<double>&& vals);
Alpha alpha_of_scalar(pmb_ctx* c, int id);
Alpha amul(pmb_ctx* c, Alpha a, Alpha b);
bool aff_is_const(pmb_ctx* c, const AffForm& f);
bool quad_is_const(pmb_ctx* c, const QuadForm& q);
AffForm aff_of(pmb_ctx* c, int id);
QuadForm quad_of(pmb_ctx* c, int id);
void build_update_graph(pmb_ctx* c);  // pmb_plan.cu

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) pmb::fail(PMB_ECUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); \
    } while (0)

// ---- device memory ---------------------------------------------------------------------------
template <class T>
inline T* dalloc(pmb_ctx* c, int64_t count, bool zero = false) {
    if (count <= 0) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)count * sizeof(T));
    if (e != cudaSuccess) fail(PMB_ENOMEM, "cudaMalloc of %lld bytes failed: %s", (long long)(count * sizeof(T)), cudaGetErrorString(e));
    c->allocs.push_back(p);
    c->bytes_alloc += count * (int64_t)sizeof(T);
    if (zero) CK(cudaMemsetAsync(p, 0, (size_t)count * sizeof(T), c->stream));
    return (T*)p;
}

template <class T>
inline T* upload(pmb_ctx* c, const std::vector<T>& v) {
    if (v.empty()) return nullptr;
    T* d = dalloc<T>(c, (int64_t)v.size());
    // synchronous copy: `v` is usually a temporary
    CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

inline AlphaDev alpha_dev(pmb_ctx* c, const Alpha& a) {
    AlphaDev d;
    d.c = a.c; d.ptr = nullptr; d.bstride = 0;
    if (a.node >= 0) {
        const Node& n = c->nodes[a.node];
        d.ptr = n.dptr; d.bstride = n.bstride;
    }
    return d;
}

inline int32_t optcol(pmb_ctx* c, int64_t var) { return (int32_t)(c->varmap[(size_t)(var - 1)] - 1); }

}  // namespace pmb

#define API_BEGIN(ctx)                 \
    if (!(ctx)) return PMB_EINVAL;     \
    try {
#define API_END(ctx)                                         \
    }                                                        \
    catch (const pmb::Err& e_) {                             \
        const_cast<pmb_ctx*>(ctx)->err = e_.msg;             \
        return e_.code;                                      \
    }                                                        \
    catch (const std::bad_alloc&) {                          \
        const_cast<pmb_ctx*>(ctx)->err = "out of host memory"; \
        return PMB_ENOMEM;                                   \
    }                                                        \
    return PMB_OK;

