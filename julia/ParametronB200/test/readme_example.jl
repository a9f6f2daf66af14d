# README Example 1 of Parametron.jl (README.md:31-56) on the B200 path: only the optimizer line differs.
# Run on a host with Julia 1.x, Parametron 0.9, MathOptInterface 0.8, OSQP, a B200 and libpmb200.so (PARAMETRON_B200_LIB).
using Parametron, ParametronB200, OSQP, Random, LinearAlgebra, Test
using OSQP.MathOptInterfaceOSQP: Optimizer

n, m = 8, 2
rng = MersenneTwister(1234)
optimizer = B200Optimizer(Optimizer())                      # <- was: Optimizer()
MOI = ParametronB200.MOI
MOI.set(optimizer, OSQPSettings.Verbose(), false)
MOI.set(optimizer, OSQPSettings.EpsAbs(), 1e-8)
MOI.set(optimizer, OSQPSettings.EpsRel(), 1e-16)
model = Model(optimizer)
x = [Variable(model) for _ = 1 : n]
A = Parameter(A -> rand!(rng, A), zeros(n, n), model)
b = Parameter(b -> rand!(rng, b), zeros(n), model)
C = Parameter(zeros(m, n), model) do C
    rand!(rng, C); C .*= 2
end
d = Parameter(d -> rand!(rng, d), zeros(m), model)
residual = @expression A * x - b
@objective(model, Minimize, residual ⋅ residual)
@constraint(model, C * x == d)

for trial = 1 : 5
    solve!(model)
    xsol = value.(model, x)
    @test C() * xsol ≈ d() atol = 1e-4                       # test/model.jl:83-125 restated: feasibility and the closed-form optimum
    Ao, bo, Co, dvo = A(), b(), C(), d()
    KKT = [2Ao'Ao Co'; Co zeros(m, m)]
    ref = (KKT \ [2Ao'bo; dvo])[1:n]
    @test xsol ≈ ref rtol = 1e-4
end
