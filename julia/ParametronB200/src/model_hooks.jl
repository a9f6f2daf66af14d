# model_hooks.jl -- `initialize!` / `update!` / `solve!` of a Model whose optimizer is wrapped in B200Optimizer.

"""
    B200Optimizer(inner; device = 0, batch = 1, gram_mode = 1, device_resident = false)

Wraps any `MOI.AbstractOptimizer`.  A `Parametron.Model` built on it assembles its coefficients on the B200 (`update!`); everything else
(copy_to, optimize!, result queries) is forwarded to `inner`.  `device_resident = true` skips the term-list hand-off: the solver is
expected to read the device views (`device_objective`, `device_block`).
"""
mutable struct B200Optimizer{O <: MOI.AbstractOptimizer} <: MOI.AbstractOptimizer
    inner::O
    device::Int
    batch::Int
    gram_mode::Int
    device_resident::Bool
    state::Union{Nothing, DeviceState}
end
B200Optimizer(inner::MOI.AbstractOptimizer; device = 0, batch = 1, gram_mode = 1, device_resident = false) =
    B200Optimizer(inner, device, batch, gram_mode, device_resident, nothing)

# ---- MOI forwarding (the external solve itself stays untouched) ------------------------------------------------------------------
MOI.copy_to(o::B200Optimizer, src::MOI.ModelLike; kwargs...) = MOI.copy_to(o.inner, src; kwargs...)
MOI.optimize!(o::B200Optimizer) = MOI.optimize!(o.inner)
MOI.is_empty(o::B200Optimizer) = MOI.is_empty(o.inner)
MOI.empty!(o::B200Optimizer) = (o.state = nothing; MOI.empty!(o.inner))
MOI.get(o::B200Optimizer, args...) = MOI.get(o.inner, args...)
MOI.set(o::B200Optimizer, args...) = MOI.set(o.inner, args...)
MOI.supports(o::B200Optimizer, args...) = MOI.supports(o.inner, args...)
MOI.supports_constraint(o::B200Optimizer, args...) = MOI.supports_constraint(o.inner, args...)

device_context(m::Model) = m.optimizer.state.ctx

# ---- set kinds (model.jl:75-95; bucket order moi_interop.jl:180-193) -------------------------------------------------------------
set_kind(::Union{MOI.GreaterThan, MOI.Nonnegatives}) = LibPMB.SET_NONNEGATIVE
set_kind(::Union{MOI.LessThan, MOI.Nonpositives}) = LibPMB.SET_NONPOSITIVE
set_kind(::Union{MOI.EqualTo, MOI.Zeros}) = LibPMB.SET_ZERO

# every constraint of the model in UPDATE order: the 14 buckets in declaration order, insertion order inside a bucket (moi_interop.jl:236-247)
function each_constraint(f, m::Model)
    for name in fieldnames(typeof(m.constraints)), c in getfield(m.constraints, name)
        f(c)
    end
end
has_data(c::Parametron.Constraint{Nothing}) = false          # SingleVariable-in-Integer / ZeroOne: nothing to update (moi_interop.jl:176)
has_data(c::Parametron.Constraint) = true

"""
`initialize!` (model.jl:117-122): the stock copy to the optimizer and index mapping, then the device lowering.
"""
function Parametron.initialize!(m::Model{T, <:B200Optimizer}) where {T}
    T === Float64 || throw(Unsupported("the device path is fp64 (Model{Float64})"))
    indexmap = MOI.copy_to(m.optimizer, m.backend)
    Parametron.mapindices!(m, indexmap)
    opt = m.optimizer
    ctx = LibPMB.Ctx(device = opt.device, batch = opt.batch)
    LibPMB.set_gram_mode!(ctx, opt.gram_mode)
    st = DeviceState(ctx)
    st.nvars = length(m.model_var_to_optimizer)
    LibPMB.add_variables!(ctx, st.nvars)
    # objective (model.jl:60-66): the wrapped, optimized expression stored by Objective(T, expr) (moi_interop.jl:119-129)
    sense = MOI.get(m.backend, MOI.ObjectiveSense()) == MOI.MaxSense ? LibPMB.MAXIMIZE : LibPMB.MINIMIZE
    LibPMB.set_objective!(ctx, sense, lower(st, m.objective.expr))
    # constraints, bucket by bucket: the library keeps the same order, so the running position IS the block's update position
    pos = Int32(0)
    each_constraint(m) do c
        has_data(c) || return
        LibPMB.add_constraint!(ctx, lower(st, c.expr), set_kind(c.set))
        st.block_of[c] = pos
        pos += Int32(1)
    end
    LibPMB.initialize!(ctx, Int64[vi.value for vi in m.model_var_to_optimizer])       # mapindices!, model.jl:100-107
    opt.state = st
    m.initialized = true
    nothing
end

# ---- hand-off to an unmodified MOI optimizer: fill the reference's own MOI functions from device-written records -------------------
function scratch!(st::DeviceState, key::Symbol, bytes::Integer)
    get!(st.scratch, key) do
        LibPMB.device_alloc(st.ctx, bytes)
    end
end

function pull!(st::DeviceState, dst::Vector, dev::Ptr{Cvoid})
    GC.@preserve dst LibPMB.memcpy_d2h(st.ctx, Ptr{Cvoid}(pointer(dst)), dev, sizeof(dst))
end

# MOI.ScalarQuadraticFunction (moi_interop.jl:45-62): N (N + 1) / 2 ScalarQuadraticTerm{Float64} (24 B: coefficient, variable_index_1,
# variable_index_2 -- the record pmb_objective_export_terms writes), N ScalarAffineTerm{Float64} (16 B), the constant
function handoff_objective!(m::Model, st::DeviceState, f::MOI.ScalarQuadraticFunction{Float64})
    N = st.nvars
    np = N * (N + 1) ÷ 2
    resize!(f.quadratic_terms, np); resize!(f.affine_terms, N)
    dq = scratch!(st, :obj_quad, max(24np, 8)); da = scratch!(st, :obj_aff, max(16N, 8)); dc = scratch!(st, :obj_const, 8)
    LibPMB.objective_export_terms!(st.ctx, dq, da, dc)
    c = Vector{Float64}(undef, 1)
    pull!(st, f.quadratic_terms, dq); pull!(st, f.affine_terms, da); pull!(st, c, dc)
    LibPMB.synchronize(st.ctx)
    f.constant = c[1]
    MOI.set(m.optimizer.inner, MOI.ObjectiveFunction{typeof(f)}(), f)                 # moi_interop.jl:134
end
function handoff_objective!(m::Model, st::DeviceState, f::MOI.ScalarAffineFunction{Float64})
    N = st.nvars
    resize!(f.terms, N)
    da = scratch!(st, :obj_aff, max(16N, 8)); dc = scratch!(st, :obj_const, 8)
    LibPMB.objective_export_terms!(st.ctx, Ptr{Cvoid}(C_NULL), da, dc)
    c = Vector{Float64}(undef, 1)
    pull!(st, f.terms, da); pull!(st, c, dc)
    LibPMB.synchronize(st.ctx)
    f.constant = c[1]
    MOI.set(m.optimizer.inner, MOI.ObjectiveFunction{typeof(f)}(), f)
end

# MOI.VectorAffineFunction (moi_interop.jl:64-81): VectorAffineTerm{Float64} = (output_index::Int64, ScalarAffineTerm(coefficient, variable)),
# 24 bytes, row-outer / term-inner -- exactly the record pmb_block_export_terms writes, so the device buffer is copied straight into f.terms
function handoff_block!(m::Model, st::DeviceState, c::Parametron.Constraint, f::MOI.VectorAffineFunction{Float64})
    pos = st.block_of[c]
    v = LibPMB.block_view(st.ctx, pos)
    resize!(f.terms, v.nterms); resize!(f.constants, v.rows)
    dt = scratch!(st, Symbol(:blk_t_, pos), max(24 * v.nterms, 8)); dk = scratch!(st, Symbol(:blk_c_, pos), max(8 * v.rows, 8))
    LibPMB.block_export_terms!(st.ctx, pos, dt, dk)
    pull!(st, f.terms, dt); pull!(st, f.constants, dk)
    LibPMB.synchronize(st.ctx)
    MOI.set(m.optimizer.inner, MOI.ConstraintFunction(), c.optimizerindex, f)         # moi_interop.jl:171
end
# scalar affine constraint (moi_interop.jl:35-43): the block has one row; its 24-byte records are repacked into ScalarAffineTerms on the host
function handoff_block!(m::Model, st::DeviceState, c::Parametron.Constraint, f::MOI.ScalarAffineFunction{Float64})
    pos = st.block_of[c]
    v = LibPMB.block_view(st.ctx, pos)
    recs = Vector{MOI.VectorAffineTerm{Float64}}(undef, v.nterms)
    k = Vector{Float64}(undef, 1)
    dt = scratch!(st, Symbol(:blk_t_, pos), max(24 * v.nterms, 8)); dk = scratch!(st, Symbol(:blk_c_, pos), 8)
    LibPMB.block_export_terms!(st.ctx, pos, dt, dk)
    pull!(st, recs, dt); pull!(st, k, dk)
    LibPMB.synchronize(st.ctx)
    resize!(f.terms, v.nterms)
    for (i, r) in enumerate(recs)
        f.terms[i] = r.scalar_term
    end
    f.constant = k[1]
    MOI.set(m.optimizer.inner, MOI.ConstraintFunction(), c.optimizerindex, f)
end
handoff_block!(m::Model, st::DeviceState, c::Parametron.Constraint, f) =
    throw(Unsupported("no term-list hand-off for $(typeof(f)); use device_resident = true and read device_block"))

"""
`update!` (model.jl:132-143): mark every Parameter dirty, evaluate each once (C callback), ONE device pass re-assembles the objective and
every non-constant constraint block in bucket order; then the hand-off.
"""
function Parametron.update!(m::Model{T, <:B200Optimizer}) where {T}
    st = m.optimizer.state
    Parametron.setdirty!(m)
    LibPMB.update!(st.ctx)
    if !m.optimizer.device_resident
        m.objective.isconstant || handoff_objective!(m, st, m.objective.f)
        each_constraint(m) do c
            (has_data(c) && !c.isconstant) && handoff_block!(m, st, c, c.f)
        end
    end
    nothing
end

# ---- device-resident hand-off (replaces MOI.set, moi_interop.jl:134,171) for solvers that read the GPU directly ----------------------
"""(Q, q, c0) device pointers of the last `update!`: `Q` is the symmetric MOI Hessian, N x N column-major (`LibPMB.ObjectiveView`)."""
device_objective(m::Model) = LibPMB.objective_view(device_context(m))
"""(C, d) device pointers of a constraint's block (`C x + d ∈ set`, m_i x N column-major, optimizer-variable columns)."""
device_block(m::Model, c::Parametron.Constraint) = LibPMB.block_view(device_context(m), m.optimizer.state.block_of[c])
