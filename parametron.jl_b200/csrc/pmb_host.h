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

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) pmb::fail(PMB_ECUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); \
    } while (0)
