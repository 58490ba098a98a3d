# julia/test/runtests.jl -- the reference's own hot-path tests, re-run through the native path (AD = AutoB200()).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI (no Julia in the build image, SURVEY.md F2); it is what a maintainer with Julia,
# AdvancedVI.jl v0.7 and a B200 runs:
#     LIBAVI_B200=/path/to/libavi_b200.so julia --project=julia julia/test/runtests.jl
# Each testset names the reference test it mirrors; the Python mirror (tests/test_gpu_*.py) runs the same call
# sequences against the same library on every round.  Everything is Float32 (the native path's only element type;
# `AutoB200` raises an ArgumentError otherwise, tested below).
using Test
using Random, LinearAlgebra, Statistics
using ADTypes, DiffResults, Distributions, LogDensityProblems, Optimisers, StableRNGs
using AdvancedVI

include(joinpath(@__DIR__, "..", "AdvancedVIB200.jl"))
using .AdvancedVIB200

const AD = AutoB200()
const SEED = 0x38bef07cf9cc549d          # test/algorithms/klminrepgraddescent.jl:43

# test/models/normal.jl:56-75 (normal_meanfield) as a native target
function normal_meanfield_native(; n_dims=5, σ0=0.3f0)
    μ = fill(5.0f0, n_dims)
    σ = fill(σ0, n_dims)
    return (model=AdvancedVIB200.MvNormalDiag(μ, σ), n_dims=n_dims, μ_true=μ, L_true=Diagonal(σ))
end

# the same target as a plain LogDensityProblem (goes through the per-sample host callback)
struct PlainNormal{C}
    μ::Vector{Float32}
    σ::Vector{Float32}
    cap::C
end
LogDensityProblems.dimension(p::PlainNormal) = length(p.μ)
LogDensityProblems.capabilities(::Type{PlainNormal{C}}) where {C} = C()
LogDensityProblems.logdensity(p::PlainNormal, θ) = sum(logpdf.(Normal.(p.μ, p.σ), θ))
function LogDensityProblems.logdensity_and_gradient(p::PlainNormal, θ)
    return LogDensityProblems.logdensity(p, θ), -(θ .- p.μ) ./ p.σ .^ 2
end

@testset "AdvancedVIB200" begin
    (; model, n_dims, μ_true, L_true) = normal_meanfield_native()
    q0 = MeanFieldGaussian(zeros(Float32, n_dims), Diagonal(ones(Float32, n_dims)))

    # test/algorithms/klminrepgraddescent.jl:8-12
    @testset "basic n_samples=$(n_samples)" for n_samples in [1, 10]
        alg = KLMinRepGradDescent(AD; n_samples, operator=ClipScale())
        optimize(alg, 1, model, q0; show_progress=false)
    end

    # :14-20
    @testset "callback" begin
        alg = KLMinRepGradDescent(AD; operator=ClipScale())
        T = 10
        callback(; iteration, kwargs...) = (iteration_check=iteration,)
        _, info, _ = optimize(alg, T, model, q0; callback, show_progress=false)
        @test [i.iteration_check for i in info] == 1:T
    end

    # :22-38
    @testset "estimate_objective" begin
        alg = KLMinRepGradDescent(AD; operator=ClipScale())
        q_true = MeanFieldGaussian(Vector(μ_true), Diagonal(L_true))
        @test isfinite(estimate_objective(alg, q_true, model))
        @test isfinite(estimate_objective(alg, q_true, model; n_samples=1))
        @test isfinite(estimate_objective(alg, q_true, model; n_samples=3))
        @test estimate_objective(alg, q_true, model; n_samples=10^5) ≈ 0 atol = 1e-2
    end

    # :40-57
    @testset "determinism" begin
        alg = KLMinRepGradDescent(AD; operator=ClipScale())
        T = 10
        q_out, _, _ = optimize(StableRNG(SEED), alg, T, model, q0; show_progress=false)
        q_rep, _, _ = optimize(StableRNG(SEED), alg, T, model, q0; show_progress=false)
        @test q_out.location == q_rep.location
        @test q_out.scale == q_rep.scale
    end

    # :59-64
    @testset "warn MvLocationScale with IdentityOperator" begin
        @test_warn "IdentityOperator" begin
            alg′ = KLMinRepGradDescent(AD; operator=IdentityOperator())
            optimize(alg′, 1, model, q0; show_progress=false)
        end
    end

    # :66-87 (through the objective interface instead of AdvancedVI._value_and_gradient!, which AutoB200 replaces)
    @testset "STL variance reduction n_montecarlo=$(n_montecarlo)" for n_montecarlo in [1, 10]
        q_true = MeanFieldGaussian(Vector(μ_true), Diagonal(L_true))
        params, re = Optimisers.destructure(q_true)
        obj = RepGradELBO(n_montecarlo; entropy=StickingTheLandingEntropy())
        out = DiffResults.DiffResult(zero(eltype(params)), similar(params))
        st = AdvancedVI.init(Random.default_rng(), obj, AD, q_true, model, params, re)
        AdvancedVI.estimate_gradient!(Random.default_rng(), obj, AD, out, st, params, re)
        @test norm(DiffResults.gradient(out)) ≈ 0 atol = 1e-4
    end

    # :90-103: the native path is Float32 only and says so
    @testset "type stability / element type" begin
        alg = KLMinRepGradDescent(AD; n_samples=10, operator=ClipScale())
        q_out, info, _ = optimize(alg, 1, model, q0; show_progress=false)
        @test eltype(q_out.location) == Float32
        @test eltype(q_out.scale) == Float32
        @test typeof(first(info).elbo) == Float32
        q64 = MeanFieldGaussian(zeros(Float64, n_dims), Diagonal(ones(Float64, n_dims)))
        @test_throws ArgumentError optimize(alg, 1, model, q64; show_progress=false)
    end

    # :105-121
    @testset "convergence $(entropy)" for entropy in [ClosedFormEntropy(), StickingTheLandingEntropy()]
        alg = KLMinRepGradDescent(AD; entropy, optimizer=Descent(1.0f-3), operator=ClipScale())
        q_out, _, _ = optimize(alg, 1000, model, q0; show_progress=false)
        Δλ0 = sum(abs2, q0.location - μ_true) + sum(abs2, q0.scale - L_true)
        Δλ = sum(abs2, q_out.location - μ_true) + sum(abs2, q_out.scale - L_true)
        @test Δλ ≤ Δλ0 / 2
    end

    # test/general/optimize.jl:27-40
    @testset "warm start" begin
        alg = KLMinRepGradDescent(AD; optimizer=Optimisers.Adam(1.0f-2), operator=ClipScale())
        T = 200
        q_ref, _, _ = optimize(StableRNG(SEED), alg, T, model, q0; show_progress=false)
        rng = StableRNG(SEED)
        _, _, state = optimize(rng, alg, T ÷ 2, model, q0; show_progress=false)
        q_avg, _, _ = optimize(rng, alg, T - T ÷ 2, model, q0; show_progress=false, state)
        @test q_avg.location == q_ref.location
        @test q_avg.scale == q_ref.scale
    end

    # test/algorithms/klminscoregraddescent.jl:82-97 and klminrepgradproxdescent.jl (same skeleton)
    @testset "KLMinScoreGradDescent / KLMinRepGradProxDescent run and improve" begin
        for alg in (KLMinScoreGradDescent(AD; n_samples=100, optimizer=Descent(1.0f-3), operator=ClipScale()),
                    KLMinRepGradProxDescent(AD; n_samples=10))
            q_out, info, _ = optimize(alg, 300, model, q0; show_progress=false)
            @test isfinite(last(info).elbo)
            @test sum(abs2, q_out.location - μ_true) < sum(abs2, q0.location - μ_true)
        end
    end

    # any other LogDensityProblem goes through the host callback; a capability-0 target gets its gradient from the
    # fallback (src/algorithms/repgradelbo.jl:50-62 differentiates through logdensity instead)
    @testset "host-callback targets, capability $(cap)" for cap in (LogDensityProblems.LogDensityOrder{1}(),
                                                                     LogDensityProblems.LogDensityOrder{0}())
        plain = PlainNormal(Vector(μ_true), fill(0.3f0, n_dims), cap)
        alg = KLMinRepGradDescent(AD; n_samples=4, optimizer=Descent(1.0f-3), operator=ClipScale())
        q_out, info, _ = optimize(alg, 200, plain, q0; show_progress=false)
        @test isfinite(last(info).elbo)
        @test sum(abs2, q_out.location - μ_true) < sum(abs2, q0.location - μ_true)
    end

    # test/general/subsampledobj.jl:62-89 on the native logistic regression: the mean over an epoch of minibatch
    # gradients (same Monte-Carlo samples) equals the full-batch gradient; estimate_objective of the subsampled
    # objective (subsampledobjective.jl:47-58) agrees with the full one; `subsample` never alters `prob`
    @testset "SubsampledObjective batchsize=$(batchsize)" for batchsize in [1, 3, 4]
        n_data, d = 8, 3
        rng = StableRNG(SEED)
        X = randn(rng, Float32, n_data, d); y = Float32.(rand(rng, n_data) .< 0.5)
        prob = AdvancedVIB200.LogReg(X, y; gemm=0)
        D = d + 1
        q = MeanFieldGaussian(zeros(Float32, D), Diagonal(fill(0.5f0, D)))
        params, re = Optimisers.destructure(q)
        full_obj = RepGradELBO(10)
        sub = ReshufflingBatchSubsampling(1:n_data, batchsize)
        sub_obj = SubsampledObjective(full_obj, sub)
        out = DiffResults.DiffResult(zero(eltype(params)), similar(params))

        full_state = AdvancedVI.init(StableRNG(SEED), full_obj, AD, q, prob, params, re)
        AdvancedVI.estimate_gradient!(StableRNG(SEED), full_obj, AD, out, full_state, params, re)
        grad_ref = copy(DiffResults.gradient(out))

        sub_state = AdvancedVI.init(StableRNG(SEED), sub_obj, AD, q, prob, params, re)
        grads = map(1:length(sub)) do _
            _, sub_state, _ = AdvancedVI.estimate_gradient!(StableRNG(SEED), sub_obj, AD, out, sub_state, params, re)
            copy(DiffResults.gradient(out))
        end
        @test mean(grads) ≈ grad_ref rtol = 1e-3

        z = randn(rng, Float32, D)
        lp_before = LogDensityProblems.logdensity(prob, z)
        view = AdvancedVI.subsample(prob, 1:batchsize)
        @test view isa AdvancedVIB200.NativeProblemView
        @test LogDensityProblems.logdensity(prob, z) == lp_before        # `prob` still evaluates on all rows

        full_val = estimate_objective(StableRNG(SEED), full_obj, q, prob; n_samples=10^5)
        sub_val = estimate_objective(StableRNG(SEED), sub_obj, q, prob; n_samples=10^5)
        @test full_val ≈ sub_val rtol = 0.1
    end

    # test/families/location_scale.jl:22-25 + docs/src/families.md:72-101: MvLocationScale with a non-Gaussian base
    # distribution.  The native path draws the base itself (avi_obj_set_base), so the check is distributional: the ELBO
    # estimate at q == q is finite and a short run moves the location towards the target's mean.
    @testset "base distribution $(basedist)" for basedist in [Laplace(0.0f0, 1.0f0), TDist(5.0f0)]
        q_t = MvLocationScale(zeros(Float32, n_dims), Diagonal(ones(Float32, n_dims)), basedist)
        alg = KLMinRepGradDescent(AD; n_samples=10, optimizer=Optimisers.Adam(1.0f-2), operator=ClipScale())
        q_avg, info, _ = optimize(StableRNG(SEED), alg, 200, model, q_t; show_progress=false)
        @test q_avg isa MvLocationScale && q_avg.dist == basedist
        @test all(isfinite, [i.elbo for i in info])
        @test sum(abs2, q_avg.location - μ_true) < sum(abs2, q_t.location - μ_true)
        @test_throws ArgumentError AdvancedVIB200.base_code(Normal(1.0f0, 2.0f0))
    end

    # src/algorithms/fisherminbatchmatch.jl:81-111: the sampling stage of FisherMinBatchMatch over a native target
    @testset "rand_batch_match_samples_with_objective!" begin
        q_fr = FullRankGaussian(zeros(Float32, n_dims), LowerTriangular(Matrix{Float32}(I, n_dims, n_dims)))
        u, z, g, fisher, logπ_avg = AdvancedVI.rand_batch_match_samples_with_objective!(StableRNG(SEED), q_fr, 64, model)
        @test size(u) == size(z) == size(g) == (n_dims, 64)
        @test z ≈ q_fr.scale * u .+ q_fr.location rtol = 1.0f-5
        @test g ≈ -(z .- μ_true) ./ diag(L_true) .^ 2 rtol = 1.0f-4              # test/models/normal.jl:8-11
        @test fisher ≈ sum(abs2, -u - q_fr.scale' * g) / 64 rtol = 1.0f-4
        @test logπ_avg ≈ mean(LogDensityProblems.logdensity(model, z[:, b]) for b in 1:64) rtol = 1.0f-4
    end
end
