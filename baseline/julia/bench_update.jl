# bench_update.jl -- the reference's own CPU timing of the hot path (SURVEY.md section 8d-1, BASELINE.md section 3-1).
# The literal README model (README.md:31-56) at a given size; times `Parametron.update!(model)` -- every Parameter re-evaluated once, the
# objective and the constraint block re-materialised as term lists and pushed with MOI.set (src/model.jl:132-143) -- single-threaded, as
# the reference is.  NOT executable in the build image or on the GPU box (no Julia there); shipped for any host with Julia 1.x,
# Parametron 0.9, MathOptInterface 0.8 and OSQP.
#
#   julia --project baseline/julia/bench_update.jl [n=8] [m=2] [repeats=20]
#
# Memory: the objective's term list holds k n^2 QuadraticTerms of 24 bytes (functions.jl:551-559) and update! copies it once more
# (moi_interop.jl:53-60): n = 512 needs ~6.5 GB, n = 4096 (BASELINE configs[1]) 3.3 TB -- run the sizes that fit and extrapolate with
# the cubic fit below (bench.py reports the same fit for the C port of this path, `cpu_baseline.termlist_full_runs_cubic_fit`).
using Parametron, OSQP, Random, LinearAlgebra, Printf
using OSQP.MathOptInterfaceOSQP: Optimizer

function readme_model(n, m; rng = MersenneTwister(1234))
    model = Model(Optimizer())
    x = [Variable(model) for _ = 1 : n]
    A = Parameter(A -> rand!(rng, A), zeros(n, n), model)
    b = Parameter(b -> rand!(rng, b), zeros(n), model)
    C = Parameter(C -> rand!(rng, C), zeros(m, n), model)
    d = Parameter(d -> (rand!(rng, d); d .*= 2), zeros(m), model)
    residual = @expression A * x - b
    @objective(model, Minimize, residual ⋅ residual)
    @constraint(model, C * x == d)
    model
end

function time_update(n, m, repeats)
    model = readme_model(n, m)
    Parametron.initialize!(model)
    Parametron.update!(model)                                  # first call allocates the term lists (test/model.jl:117-122 excludes it too)
    best = Inf
    for _ = 1 : repeats
        best = min(best, @elapsed Parametron.update!(model))
    end
    best
end

function main()
    n = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 8
    m = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 2
    repeats = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 20
    t = time_update(n, m, repeats)
    @printf("{\"impl\": \"Parametron.jl (Julia %s, 1 thread)\", \"n\": %d, \"m\": %d, \"seconds_per_update\": %.6e, \"updates_per_s\": %.6e}\n",
            VERSION, n, m, t, 1 / t)
    if n <= 8                                                   # sweep + cubic fit, extrapolated to BASELINE configs[1] (labelled as such)
        sizes = [64, 128, 256, 384]
        ts = [time_update(s, s ÷ 4, 3) for s in sizes]
        X = hcat(Float64.(sizes) .^ 3, Float64.(sizes) .^ 2)
        a, c = X \ ts
        @printf("{\"fit\": \"t = a n^3 + c n^2\", \"a\": %.4e, \"c\": %.4e, \"extrapolated_seconds_at_n4096\": %.4e}\n", a, c, a * 4096.0^3 + c * 4096.0^2)
    end
end

main()
