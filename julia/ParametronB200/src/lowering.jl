# lowering.jl -- LazyExpression tree -> device DAG.  One opcode per rewrite rule of Parametron's `optimize` table
# (src/lazyexpression.jl:198-302); the tree walked here is what `@expression` leaves behind AFTER optimize_toplevel, i.e. nodes whose
# function is one of the in-place kernels of src/functions.jl with the preallocated destination as first argument.

mutable struct DeviceState
    ctx::LibPMB.Ctx
    nodes::IdDict{Any, Int32}          # Parameter / LazyExpression / constant array -> node id (every Parameter is registered once)
    nvars::Int
    block_of::IdDict{Any, Int32}       # Parametron Constraint -> update position of its block
    scratch::Dict{Symbol, Ptr{Cvoid}}  # device scratch for the term-list hand-off
end

DeviceState(ctx) = DeviceState(ctx, IdDict{Any, Int32}(), 0, IdDict{Any, Int32}(), Dict{Symbol, Ptr{Cvoid}}())

struct Unsupported <: Exception
    msg::String
end

varindex(x::Variable) = Int64(x.index)

# ---- leaves --------------------------------------------------------------------------------------------------------------------
function lower(st::DeviceState, p::Parameter)
    get!(st.nodes, p) do
        val = p()                                               # current value: fixes the shape (parameter.jl:48-50 does the same)
        if val isa Number
            holder = [Float64(val)]                             # scalars travel through a 1-element staging buffer
            LibPMB.register_scalar_parameter!(st.ctx, p, holder)
        elseif val isa Array{Float64}
            LibPMB.register_parameter!(st.ctx, p, val)          # in-place Parameters: the library reads the Parameter's own buffer
        else
            throw(Unsupported("Parameter of type $(typeof(val)) has no device layout (numbers, Vector{Float64}, Matrix{Float64} are supported; " *
                              "struct-valued Parameters need a flat record, see pmb_op_getfield)"))
        end
    end
end

lower(st::DeviceState, x::Number) = LibPMB.constant(st.ctx, Float64(x))
lower(st::DeviceState, x::AbstractVector{<:Number}) = LibPMB.constant(st.ctx, collect(Float64, x))
lower(st::DeviceState, x::AbstractMatrix{<:Number}) = LibPMB.constant(st.ctx, Matrix{Float64}(x))
lower(st::DeviceState, x::AbstractVector{Variable}) = LibPMB.varvec(st.ctx, Int64[varindex(v) for v in x])
lower(st::DeviceState, x::Variable) = LibPMB.const_affine(st.ctx, [1.0], [varindex(x)], 0.0)
lower(st::DeviceState, x::Union{Transpose{Variable}, Adjoint{Variable}}) = LibPMB.adjoint_(st.ctx, lower(st, parent(x)))

function lower(st::DeviceState, f::AffineFunction)               # a constant AffineFunction: terms in the reference's order
    LibPMB.const_affine(st.ctx, Float64[t.coeff for t in f.linear], Int64[varindex(t.var) for t in f.linear], Float64(f.constant[]))
end
function lower(st::DeviceState, f::QuadraticFunction)
    LibPMB.const_quadratic(st.ctx, Float64[t.coeff for t in f.quadratic], Int64[varindex(t.rowvar) for t in f.quadratic],
                           Int64[varindex(t.colvar) for t in f.quadratic], Float64[t.coeff for t in f.affine.linear],
                           Int64[varindex(t.var) for t in f.affine.linear], Float64(f.affine.constant[]))
end

# ---- interior nodes ------------------------------------------------------------------------------------------------------------
# Optimized nodes carry (dest, operands...): the destination is the reference's preallocated buffer and is skipped.
const BINARY_INPLACE = IdDict{Any, Symbol}(
    Functions.matvecmul! => :matvec,         # A * x                     lazyexpression.jl:200-204, functions.jl:775-829
    Functions.vecdot! => :dot_,              # x ⋅ y                     :228-232, functions.jl:665-731
    Functions.add! => :add, Functions.vecadd! => :add,                   # :238-247
    Functions.subtract! => :sub, Functions.vecsubtract! => :sub,         # :249-258
    Functions.mul! => :mul, Functions.scale! => :mul)                    # :260-274, :284-290

function lower(st::DeviceState, e::LazyExpression)
    haskey(st.nodes, e) && return st.nodes[e]
    f, args = e.f, e.args
    node = if f === identity
        lower(st, args[1])
    elseif f isa Parametron.FunctionWrapper                     # a WrappedExpression: unwrap (FunctionWrappersQuickFix.jl:86-102 keeps the callable in `obj`)
        lower(st, f.obj[])
    elseif haskey(BINARY_INPLACE, f)
        op = getfield(LibPMB, BINARY_INPLACE[f])
        op(st.ctx, lower(st, args[2]), lower(st, args[3]))
    elseif f === Functions.bilinearmul!                         # x' * Q * y, :219-226: args = (dest, Q, x', y)
        LibPMB.bilinear(st.ctx, lower(st, args[3]), lower(st, args[2]), lower(st, args[4]))
    elseif f === Functions.vcat!                                # :276-278
        LibPMB.vcat_(st.ctx, Int32[lower(st, a) for a in args[2:end]])
    elseif f === copyto!                                        # convert(Vector, x), :280-282
        LibPMB.convert_(st.ctx, lower(st, args[2]))
    elseif f === convert                                        # convert(E, expr) left unoptimized for scalar functions (moi_interop.jl:121,151)
        lower(st, args[2])
    elseif f isa Functions.GetField
        throw(Unsupported("p.field on a struct-valued Parameter: register the record with ParametronB200.LibPMB and pmb_op_getfield"))
    elseif length(args) == 2 && args[1] isa AbstractMatrix && !(args[2] isa AbstractVector)
        LibPMB.adjoint_(st.ctx, lower(st, args[2]))             # the closure `optimize` builds for A' (:206-217): (dest matrix, A)
    elseif length(args) == 2 && args[1] isa AbstractVector && length(args[1]) == 1
        LibPMB.vect_(st.ctx, lower(st, args[2]))                # the closure for [x] (:292-298): (1-element dest, x)
    else
        lower_generic(st, f, args)
    end
    st.nodes[e] = node
end

# Calls `optimize` left alone (no rule matched, e.g. number * number, or a user who built LazyExpression by hand): same opcodes, chosen by
# the evaluated operand types exactly as optimize_toplevel does (`typeof(evalarg(arg))`, :186-187)
function lower_generic(st::DeviceState, f, args)
    vals = map(Parametron.evalarg, args)
    if f === (*) && length(args) == 3
        LibPMB.bilinear(st.ctx, lower(st, args[1]), lower(st, args[2]), lower(st, args[3]))
    elseif f === (*) && length(args) == 2
        (vals[1] isa AbstractMatrix && vals[2] isa AbstractVector) ? LibPMB.matvec(st.ctx, lower(st, args[1]), lower(st, args[2])) :
                                                                     LibPMB.mul(st.ctx, lower(st, args[1]), lower(st, args[2]))
    elseif f === (+)
        foldl((a, b) -> LibPMB.add(st.ctx, a, b), (lower(st, a) for a in args))      # n-ary + associates to the left (:234-236)
    elseif f === (-) && length(args) == 2
        LibPMB.sub(st.ctx, lower(st, args[1]), lower(st, args[2]))
    elseif f === (-) && length(args) == 1
        LibPMB.mul(st.ctx, lower(st, -1.0), lower(st, args[1]))
    elseif f === LinearAlgebra.dot
        LibPMB.dot_(st.ctx, lower(st, args[1]), lower(st, args[2]))
    elseif f === adjoint || f === transpose
        LibPMB.adjoint_(st.ctx, lower(st, args[1]))
    elseif f === vcat
        LibPMB.vcat_(st.ctx, Int32[lower(st, a) for a in args])
    elseif f === Base.vect && length(args) == 1
        LibPMB.vect_(st.ctx, lower(st, args[1]))
    else
        throw(Unsupported("no device lowering for a call of $(f) on $(map(typeof, vals)) (model.jl:139 \"Unhandled ...\")"))
    end
end
