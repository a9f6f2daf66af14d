# ParametronB200 -- routes Parametron.jl's parameter-update hot path (`update!` + coefficient assembly, src/model.jl:132-143) through
# libpmb200.so (hand-written sm_100a CUDA kernels) while the Julia-side API stays what it is:
#
#     using Parametron, ParametronB200, OSQP
#     model = Model(B200Optimizer(OSQP.Optimizer()))        # the ONLY line of README Example 1 that changes
#     x = [Variable(model) for _ = 1 : n]
#     A = Parameter(rand!, zeros(n, n), model); ...
#     residual = @expression A * x - b
#     @objective(model, Minimize, residual ⋅ residual)
#     @constraint(model, C * x == d)
#     solve!(model)                                         # initialize! lowers the expression DAG to the device, update! = one graph launch
#
# How it hooks in (no edit of Parametron's sources -- more specific methods on our own optimizer type):
#   * `Parametron.initialize!(m::Model{T, <:B200Optimizer})` -- runs the stock MOI.copy_to / mapindices! (model.jl:100-122), then walks the
#     optimized LazyExpression trees stored in `m.objective.expr` / `constraint.expr` (the output of optimize_toplevel,
#     lazyexpression.jl:184-302) and emits one `pmb_op_*` opcode per node (lowering.jl), then `pmb_initialize(varmap)`.
#   * `Parametron.update!(m::Model{T, <:B200Optimizer})` -- `setdirty!`, `pmb_update` (Parameters are evaluated once each through a C
#     callback, then ONE device pass), then the assembled coefficients are handed to the solver: device-aware solvers read the device
#     views (`device_objective`, `device_block`); any other MOI optimizer gets the reference's own MOI functions, filled from the
#     device-written term records (moi_interop.jl:35-81 layout) and pushed with the same `MOI.set` calls as `moi_interop.jl:134,171`.
#
# WRITTEN, NOT EXECUTED: Julia is not installed in the build image or on the GPU box.  The C ABI underneath is tested through the
# equivalent ctypes binding; tests/test_julia_binding.py checks every `ccall` here against it.
module ParametronB200

using LinearAlgebra
import MathOptInterface
const MOI = MathOptInterface
import Parametron
import Parametron: Model, Parameter, Variable, LazyExpression, AffineFunction, QuadraticFunction
const Functions = Parametron.Functions

export B200Optimizer, device_objective, device_block, device_context

include("libpmb.jl")
include("lowering.jl")
include("model_hooks.jl")

end # module
