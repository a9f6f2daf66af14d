# libpmb.jl -- the thin `ccall` layer over libpmb200.so (include/parametron_b200.h).  One Julia function per C entry point; every signature
# here is cross-checked against the tested Python ctypes binding (parametron.jl_b200/_capi.py) by tests/test_julia_binding.py.
#
# WRITTEN, NOT EXECUTED: neither this container nor the GPU box has Julia (SURVEY.md section 8c).
module LibPMB

const lib = get(ENV, "PARAMETRON_B200_LIB", joinpath(@__DIR__, "..", "..", "..", "parametron.jl_b200", "lib", "libpmb200.so"))



# status codes (include/parametron_b200.h)
const PMB_OK, PMB_EINVAL, PMB_EDIM, PMB_EINITIALIZED, PMB_ENOTINIT, PMB_ECUDA, PMB_EUNSUPPORTED, PMB_ENOMEM = 0, -1, -2, -3, -4, -5, -6, -7
const KIND_NUM_SCALAR, KIND_NUM_VECTOR, KIND_NUM_MATRIX, KIND_VARVEC, KIND_AFFVEC, KIND_AFF, KIND_QUAD, KIND_ADJVEC = 0:7
const MINIMIZE, MAXIMIZE = 0, 1
const SET_NONNEGATIVE, SET_NONPOSITIVE, SET_ZERO = 0, 1, 2
const PARAM_DEVICE, PARAM_HOST, PARAM_UNIFORM = 0, 1, 2

struct PmbError <: Exception
    code::Cint
    msg::String
end

mutable struct Ctx
    ptr::Ptr{Cvoid}
    roots::Vector{Any}   # keeps Parameters / buffers alive while the library holds their pointers
end

function check(ctx::Ctx, code::Cint)
    code == PMB_OK && return nothing
    msg = unsafe_string(ccall((:pmb_last_error, lib), Cstring, (Ptr{Cvoid},), ctx.ptr))
    code == PMB_EDIM && throw(DimensionMismatch(msg))          # functions.jl:669,755,780-781,845,877
    code == PMB_EINITIALIZED && error(msg)                     # "Model has already been initialized." model.jl:50,61,69
    throw(PmbError(code, msg))
end

function Ctx(; device::Integer = 0, stream::Ptr{Cvoid} = C_NULL, batch::Integer = 1)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    code = ccall((:pmb_ctx_create, lib), Cint, (Cint, Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), device, stream, batch, out)
    code == PMB_OK || error("pmb_ctx_create failed ($code): a B200 and libpmb200.so are required, there is no CPU fallback")
    ctx = Ctx(out[], Any[])
    finalizer(c -> ccall((:pmb_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.ptr), ctx)
    ctx
end

# ---- variables / parameters -------------------------------------------------------------------------------------------
function add_variables!(ctx::Ctx, n::Integer)                       # Variable(m::Model), model.jl:49-53
    first = Ref{Int64}(0)
    check(ctx, ccall((:pmb_add_variables, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Int64}), ctx.ptr, n, first))
    first[]
end

_shape(val) = ndims(val) == 0 ? (0, 0, 0) : ndims(val) == 1 ? (1, length(val), 0) : (2, size(val, 1), size(val, 2))

# the in-place update function of a Parameter, called exactly once per update! (parameter.jl:93-102)
function _host_callback(val::Ptr{Float64}, rows::Int64, cols::Int64, batch::Int64, user::Ptr{Cvoid})::Cvoid
    p = unsafe_pointer_to_objref(user)
    p()                                                            # evaluates f at most once per update! (dirty flag, parameter.jl:93-102); in-place: fills p.val[]
    nothing
end

function register_parameter!(ctx::Ctx, p, val::Array{Float64})      # Parameter(f, val, model), parameter.jl:57
    node = Ref{Int32}(0)
    nd, r, c = _shape(val)
    check(ctx, ccall((:pmb_param_create, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Cint, Ref{Int32}), ctx.ptr, nd, r, c, PARAM_HOST, node))
    cb = @cfunction(_host_callback, Cvoid, (Ptr{Float64}, Int64, Int64, Int64, Ptr{Cvoid}))
    push!(ctx.roots, p); push!(ctx.roots, val)
    check(ctx, ccall((:pmb_param_set_host_ptr, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Cvoid}, Any), ctx.ptr, node[], val, cb, p))
    node[]                                                          # Julia arrays are column-major: no re-layout
end

# scalar (number-valued) Parameter: out-of-place in the reference (parameter.jl:48,73); `holder` is the 1-element staging buffer the callback fills
function _host_callback_scalar(val::Ptr{Float64}, rows::Int64, cols::Int64, batch::Int64, user::Ptr{Cvoid})::Cvoid
    p = unsafe_pointer_to_objref(user)
    unsafe_store!(val, Float64(p()))                               # p() evaluates the Parameter once (it was marked dirty by setdirty!)
    nothing
end
function register_scalar_parameter!(ctx::Ctx, p, holder::Vector{Float64})
    node = Ref{Int32}(0)
    check(ctx, ccall((:pmb_param_create, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Cint, Ref{Int32}), ctx.ptr, 0, 0, 0, PARAM_HOST, node))
    cb = @cfunction(_host_callback_scalar, Cvoid, (Ptr{Float64}, Int64, Int64, Int64, Ptr{Cvoid}))
    push!(ctx.roots, p); push!(ctx.roots, holder)
    check(ctx, ccall((:pmb_param_set_host_ptr, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Cvoid}, Any), ctx.ptr, node[], holder, cb, p))
    node[]
end

function register_device_parameter!(ctx::Ctx, dptr::Ptr{Float64}, dims::Tuple)   # Parameter(model; val = CuArray), parameter.jl:88
    node = Ref{Int32}(0)
    nd = length(dims); r = nd >= 1 ? dims[1] : 0; c = nd == 2 ? dims[2] : 0
    check(ctx, ccall((:pmb_param_create, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Cint, Ref{Int32}), ctx.ptr, nd, r, c, PARAM_DEVICE, node))
    check(ctx, ccall((:pmb_param_set_device_ptr, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), ctx.ptr, node[], dptr))
    node[]
end

function constant(ctx::Ctx, val::Union{Float64, Array{Float64}})    # constant folding, lazyexpression.jl:189-192
    node = Ref{Int32}(0)
    nd, r, c = _shape(val)
    buf = val isa Float64 ? [val] : val
    check(ctx, ccall((:pmb_const_numeric, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}, Ref{Int32}), ctx.ptr, nd, r, c, buf, node))
    node[]
end

function varvec(ctx::Ctx, indices::Vector{Int64})                   # Vector{Variable}
    node = Ref{Int32}(0)
    check(ctx, ccall((:pmb_varvec, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ref{Int32}), ctx.ptr, length(indices), indices, node))
    node[]
end

# ---- one @expression rewrite rule = one call (lazyexpression.jl:198-302) ------------------------------------------------
for (jl, c) in ((:matvec, :pmb_op_matvec), (:dot_, :pmb_op_dot), (:add, :pmb_op_add), (:sub, :pmb_op_sub), (:mul, :pmb_op_mul))
    @eval function $jl(ctx::Ctx, a::Int32, b::Int32)
        out = Ref{Int32}(0)
        check(ctx, ccall(($(QuoteNode(c)), lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{Int32}), ctx.ptr, a, b, out))
        out[]
    end
end
for (jl, c) in ((:adjoint_, :pmb_op_adjoint), (:vect_, :pmb_op_vect))
    @eval function $jl(ctx::Ctx, a::Int32)
        out = Ref{Int32}(0)
        check(ctx, ccall(($(QuoteNode(c)), lib), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}), ctx.ptr, a, out))
        out[]
    end
end
function bilinear(ctx::Ctx, xT::Int32, Q::Int32, y::Int32)           # transpose(x) * Q * y, lazyexpression.jl:219-226
    out = Ref{Int32}(0)
    check(ctx, ccall((:pmb_op_bilinear, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Ref{Int32}), ctx.ptr, xT, Q, y, out))
    out[]
end
function vcat_(ctx::Ctx, parts::Vector{Int32})                       # vcat!, functions.jl:969-994
    out = Ref{Int32}(0)
    check(ctx, ccall((:pmb_op_vcat, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ref{Int32}), ctx.ptr, length(parts), parts, out))
    out[]
end

# ---- model ----------------------------------------------------------------------------------------------------------------
set_objective!(ctx::Ctx, sense::Integer, node::Int32) =
    check(ctx, ccall((:pmb_set_objective, lib), Cint, (Ptr{Cvoid}, Cint, Int32), ctx.ptr, sense, node))
add_constraint!(ctx::Ctx, node::Int32, set::Integer) =
    check(ctx, ccall((:pmb_add_constraint, lib), Cint, (Ptr{Cvoid}, Int32, Cint, Ptr{Int32}), ctx.ptr, node, set, C_NULL))
# initialize!(m): varmap[i] = idxmap[MOI.VariableIndex(i)].value (mapindices!, model.jl:100-107)
initialize!(ctx::Ctx, varmap::Vector{Int64}) = check(ctx, ccall((:pmb_initialize, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}), ctx.ptr, varmap))
update!(ctx::Ctx) = check(ctx, ccall((:pmb_update, lib), Cint, (Ptr{Cvoid},), ctx.ptr))      # update!(m::Model), model.jl:132-143
synchronize(ctx::Ctx) = check(ctx, ccall((:pmb_synchronize, lib), Cint, (Ptr{Cvoid},), ctx.ptr))

struct ObjectiveView
    kind::Int32; sense::Int32; is_constant::Int32; nvars::Int64
    Q::Ptr{Float64}; q::Ptr{Float64}; c0::Ptr{Float64}
end
function objective_view(ctx::Ctx)                                    # replaces MOI.set(optimizer, ObjectiveFunction{F}(), f), moi_interop.jl:134
    v = Ref{ObjectiveView}()
    check(ctx, ccall((:pmb_get_objective, lib), Cint, (Ptr{Cvoid}, Ref{ObjectiveView}), ctx.ptr, v))
    v[]
end

struct BlockView
    set_kind::Int32; is_scalar::Int32; is_constant::Int32; bucket::Int32
    rows::Int64; nvars::Int64; C::Ptr{Float64}; d::Ptr{Float64}; nterms::Int64
    is_sparse::Int32; nnz::Int64; vals::Ptr{Float64}; sp_rows::Ptr{Int32}; sp_cols::Ptr{Int32}
    Q::Ptr{Float64}
end
function block_view(ctx::Ctx, update_pos::Integer)                   # replaces MOI.set(optimizer, ConstraintFunction(), ci, f), :171
    v = Ref{BlockView}()
    check(ctx, ccall((:pmb_get_block, lib), Cint, (Ptr{Cvoid}, Int32, Ref{BlockView}), ctx.ptr, update_pos, v))
    v[]
end

# precision mode of the Hessian contraction (build phase): 0 = fp64 DMMA only, 1 = certified int8 digit planes on tcgen05 (default)
set_gram_mode!(ctx::Ctx, mode::Integer, tol::Real = 0.0) =
    check(ctx, ccall((:pmb_set_gram_mode, lib), Cint, (Ptr{Cvoid}, Cint, Cdouble), ctx.ptr, mode, tol))

# result path, model.jl:166-173.  `x` is the solver's VariablePrimal vector in OPTIMIZER index order (host Vector or device pointer).
set_primal!(ctx::Ctx, x::Vector{Float64}) = check(ctx, ccall((:pmb_set_primal, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), ctx.ptr, x, 0))
function value(ctx::Ctx, vars::Vector{Int64}; batch::Integer = 1)     # value.(m, xs): VariablePrimal[m.model_var_to_optimizer[x.index]]
    out = Matrix{Float64}(undef, length(vars), batch)
    check(ctx, ccall((:pmb_value, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Cint), ctx.ptr, length(vars), vars, out, 0))
    out
end
function objectivevalue(ctx::Ctx; batch::Integer = 1)                 # objectivevalue(m), model.jl:173
    out = Vector{Float64}(undef, batch)
    check(ctx, ccall((:pmb_objective_value, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.ptr, out))
    out
end

# With CUDA.jl a device-aware solver wraps the views without copies, e.g.
#   v = objective_view(ctx); Q = unsafe_wrap(CuArray, CuPtr{Float64}(UInt(v.Q)), (v.nvars, v.nvars))

# ---- term-list export for unmodified MOI optimizers (moi_interop.jl:35-81) and scratch device memory ---------------------------------------
function device_alloc(ctx::Ctx, bytes::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:pmb_device_alloc, lib), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.ptr, bytes, out))
    out[]
end
device_free(ctx::Ctx, p::Ptr{Cvoid}) = check(ctx, ccall((:pmb_device_free, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.ptr, p))
memcpy_d2h(ctx::Ctx, host::Ptr{Cvoid}, dev::Ptr{Cvoid}, bytes::Integer) =
    check(ctx, ccall((:pmb_memcpy_d2h, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx.ptr, host, dev, bytes))
# 24-byte MOI.VectorAffineTerm records of block `update_pos` (row-outer / term-inner, optimizer indices) + its constants, to DEVICE buffers
block_export_terms!(ctx::Ctx, update_pos::Integer, terms_dev::Ptr{Cvoid}, constants_dev::Ptr{Cvoid}) =
    check(ctx, ccall((:pmb_block_export_terms, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}), ctx.ptr, update_pos, terms_dev, constants_dev))
# the objective as MOI.ScalarQuadraticFunction records: N (N + 1) / 2 ScalarQuadraticTerm (24 B), N ScalarAffineTerm (16 B), the constant
objective_export_terms!(ctx::Ctx, quad_dev::Ptr{Cvoid}, aff_dev::Ptr{Cvoid}, constant_dev::Ptr{Cvoid}) =
    check(ctx, ccall((:pmb_objective_export_terms, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx.ptr, quad_dev, aff_dev, constant_dev))
function num_blocks(ctx::Ctx)
    n = Ref{Int32}(0)
    code = ccall((:pmb_num_blocks, lib), Cint, (Ptr{Cvoid}, Ref{Int32}), ctx.ptr, n)
    check(ctx, code)
    n[]
end
# instance-shard data parallelism: this context owns global instances [first, first + batch) (build phase)
set_instance_offset!(ctx::Ctx, first::Integer) = check(ctx, ccall((:pmb_set_instance_offset, lib), Cint, (Ptr{Cvoid}, Int64), ctx.ptr, first))
function convert_(ctx::Ctx, a::Int32)                                 # convert(Vector, x) -> copyto!, lazyexpression.jl:280-282
    out = Ref{Int32}(0)
    check(ctx, ccall((:pmb_op_convert, lib), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}), ctx.ptr, a, out))
    out[]
end
function const_affine(ctx::Ctx, coeffs::Vector{Float64}, vars::Vector{Int64}, constant::Float64)   # a constant AffineFunction, terms in reference order
    out = Ref{Int32}(0)
    check(ctx, ccall((:pmb_const_affine, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Int64}, Cdouble, Ref{Int32}), ctx.ptr, length(coeffs), coeffs, vars, constant, out))
    out[]
end
function const_quadratic(ctx::Ctx, qc::Vector{Float64}, rv::Vector{Int64}, cv::Vector{Int64}, lc::Vector{Float64}, lv::Vector{Int64}, constant::Float64)
    out = Ref{Int32}(0)
    check(ctx, ccall((:pmb_const_quadratic, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Int64}, Cdouble, Ref{Int32}),
                     ctx.ptr, length(qc), qc, rv, cv, length(lc), lc, lv, constant, out))
    out[]
end
function node_kind(ctx::Ctx, node::Int32)
    kind = Ref{Int32}(0); rows = Ref{Int64}(0); cols = Ref{Int64}(0); isconst = Ref{Int32}(0)
    check(ctx, ccall((:pmb_node_info, lib), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}, Ref{Int64}, Ref{Int64}, Ref{Int32}), ctx.ptr, node, kind, rows, cols, isconst))
    (kind[], rows[], cols[], isconst[] != 0)
end

end # module LibPMB
