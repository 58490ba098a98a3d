# AdvancedVIB200.jl -- the @ccall glue that plugs libavi_b200.so into AdvancedVI.jl (v0.7) unchanged.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (SURVEY.md F2).  The same
# call sequence is exercised by the Python mirror (advancedvi.jl_b200/api.py) and its GPU tests; this file is
# what a maintainer adds on the Julia side.  It dispatches on a new AD-type marker, `AutoB200`, which is the
# one free field every ParamSpaceSGD algorithm threads into the objective methods
# (src/algorithms/constructors.jl:46,52), so `KLMinRepGradDescent(AutoB200(); ...)`, `optimize`, callbacks,
# `SubsampledObjective` and Turing keep working as they are.
module AdvancedVIB200

using AdvancedVI, ADTypes, DiffResults, LogDensityProblems, Random
using AdvancedVI: RepGradELBO, ScoreGradELBO, MvLocationScale, ClosedFormEntropy, MonteCarloEntropy,
                  StickingTheLandingEntropy, ClosedFormEntropyZeroGradient, StickingTheLandingEntropyZeroGradient
using LinearAlgebra: Diagonal, LowerTriangular

const libavi = get(ENV, "LIBAVI_B200", "libavi_b200.so")

struct AutoB200 <: ADTypes.AbstractADType
    device::Int
end
AutoB200() = AutoB200(0)

# ---- handles --------------------------------------------------------------------------------------
mutable struct Ctx
    h::Ptr{Cvoid}
    function Ctx(device::Integer)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(@ccall(libavi.avi_ctx_create(device::Int32, r::Ptr{Ptr{Cvoid}})::Int32), C_NULL)
        finalizer(c -> @ccall(libavi.avi_ctx_destroy(c.h::Ptr{Cvoid})::Int32), new(r[]))
    end
end
const CTX = Dict{Int,Ctx}()
ctx(dev) = get!(() -> Ctx(dev), CTX, dev)

function check(code::Int32, h)
    code == 0 && return nothing
    msg = unsafe_string(@ccall libavi.avi_last_error(h::Ptr{Cvoid})::Cstring)
    error("libavi_b200 error $code: $msg")
end

"""A native target: a LogDensityProblem that also carries a device model handle.
`LogDensityProblems.dimension/capabilities/logdensity/logdensity_and_gradient` are defined on it, so the
reference's own CPU path accepts the very same object."""
mutable struct NativeProblem
    h::Ptr{Cvoid}
    c::Ctx
    D::Int
end
LogDensityProblems.dimension(p::NativeProblem) = p.D
LogDensityProblems.capabilities(::Type{NativeProblem}) = LogDensityProblems.LogDensityOrder{1}()
function LogDensityProblems.logdensity_and_gradient(p::NativeProblem, z::AbstractVector)
    zf = Vector{Float32}(z); lp = Ref{Float32}(0); g = Vector{Float32}(undef, p.D)
    check(@ccall(libavi.avi_model_logdensity_and_gradient_host(p.h::Ptr{Cvoid}, zf::Ptr{Float32}, 1::Int32,
                 lp::Ptr{Float32}, g::Ptr{Float32})::Int32), p.c.h)
    return lp[], g
end
LogDensityProblems.logdensity(p::NativeProblem, z) = first(LogDensityProblems.logdensity_and_gradient(p, z))

"Hierarchical logistic regression of docs/src/tutorials/subsampling.md:26-38 (variant = :subsampling) or README.md:47-58 (:basic)."
function LogReg(X::Matrix{Float32}, y::Vector{Float32}; n_data=size(X, 1), variant=:subsampling, gaussian=false,
                gemm=1, device=0)
    c = ctx(device); r = Ref{Ptr{Cvoid}}(C_NULL)
    check(@ccall(libavi.avi_model_glm_create(c.h::Ptr{Cvoid}, X::Ptr{Float32}, y::Ptr{Float32}, size(X, 1)::Int64,
                 size(X, 2)::Int32, n_data::Int64, (gaussian ? 1 : 0)::Int32,
                 (variant === :subsampling ? 0 : 1)::Int32, gemm::Int32, r::Ptr{Ptr{Cvoid}})::Int32), c.h)
    p = NativeProblem(r[], c, size(X, 2) + 1)
    finalizer(q -> @ccall(libavi.avi_model_destroy(q.h::Ptr{Cvoid})::Int32), p)
end

# AdvancedVI.subsample(prob, batch) (src/AdvancedVI.jl:303-313): 1-based Julia indices -> 0-based rows
function AdvancedVI.subsample(p::NativeProblem, batch)
    idx = Int32.(batch .- 1)
    check(@ccall(libavi.avi_model_subsample(p.h::Ptr{Cvoid}, idx::Ptr{Int32}, length(idx)::Int64)::Int32), p.c.h)
    return p
end

"Any other LogDensityProblem (DynamicPPL, BridgeStan, ...) goes through the per-sample host callback."
function hostcallback_problem(prob, device)
    D = LogDensityProblems.dimension(prob)
    cap = LogDensityProblems.capabilities(prob) isa LogDensityProblems.LogDensityOrder{0} ? 0 : 1
    function cb(user::Ptr{Cvoid}, z::Ptr{Float32}, Dn::Int32, lp::Ptr{Float32}, g::Ptr{Float32})::Int32
        zz = unsafe_wrap(Array, z, Dn)
        if g == C_NULL || cap == 0
            unsafe_store!(lp, Float32(LogDensityProblems.logdensity(prob, zz)))
        else
            l, gr = LogDensityProblems.logdensity_and_gradient(prob, zz)
            unsafe_store!(lp, Float32(l)); unsafe_copyto!(g, pointer(Float32.(gr)), Dn)
        end
        return Int32(0)
    end
    c = ctx(device); r = Ref{Ptr{Cvoid}}(C_NULL)
    fp = @cfunction($cb, Int32, (Ptr{Cvoid}, Ptr{Float32}, Int32, Ptr{Float32}, Ptr{Float32}))
    check(@ccall(libavi.avi_model_hostcallback_create(c.h::Ptr{Cvoid}, D::Int32, cap::Int32, fp::Ptr{Cvoid},
                 C_NULL::Ptr{Cvoid}, r::Ptr{Ptr{Cvoid}})::Int32), c.h)
    return NativeProblem(r[], c, D), fp      # keep fp alive with the state
end
native(prob::NativeProblem, dev) = (prob, nothing)
native(prob, dev) = hostcallback_problem(prob, dev)

# ---- objective state ---------------------------------------------------------------------------------
mutable struct B200ObjState
    h::Ptr{Cvoid}
    prob::NativeProblem
    keepalive::Any
end
family_code(q::MvLocationScale{<:Diagonal}) = 0
family_code(q::MvLocationScale{<:LowerTriangular}) = 1
family_code(q::MvLocationScaleLowRank) = 2
entropy_code(::ClosedFormEntropy) = 0
entropy_code(::MonteCarloEntropy) = 1
entropy_code(::StickingTheLandingEntropy) = 2
entropy_code(::ClosedFormEntropyZeroGradient) = 3
entropy_code(::StickingTheLandingEntropyZeroGradient) = 4

function make_state(rng, kind, entropy, n_samples, adtype::AutoB200, q, prob, params)
    eltype(params) === Float32 || throw(ArgumentError("AutoB200 supports Float32 only (got $(eltype(params)))"))
    p, keep = native(prob, adtype.device)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    if q isa MvLocationScaleLowRank   # location_scale_low_rank.jl: lambda = [location; scale_diag; vec(scale_factors)]
        check(@ccall(libavi.avi_obj_create_lowrank(p.c.h::Ptr{Cvoid}, p.h::Ptr{Cvoid}, size(q.scale_factors, 2)::Int32,
                     kind::Int32, entropy::Int32, n_samples::Int32, r::Ptr{Ptr{Cvoid}})::Int32), p.c.h)
    else
        check(@ccall(libavi.avi_obj_create(p.c.h::Ptr{Cvoid}, p.h::Ptr{Cvoid}, family_code(q)::Int32, kind::Int32,
                     entropy::Int32, n_samples::Int32, r::Ptr{Ptr{Cvoid}})::Int32), p.c.h)
    end
    st = B200ObjState(r[], p, keep)
    finalizer(s -> @ccall(libavi.avi_obj_destroy(s.h::Ptr{Cvoid})::Int32), st)
    # the Julia rng is used only to draw the Philox key: same seed => identical run (klminrepgraddescent.jl:40-57)
    check(@ccall(libavi.avi_obj_seed(st.h::Ptr{Cvoid}, rand(rng, UInt64)::UInt64, 0::UInt64)::Int32), p.c.h)
    return st
end

# init (src/algorithms/abstractobjective.jl:25-35; repgradelbo.jl:41-70; scoregradelbo.jl:34-50)
AdvancedVI.init(rng::Random.AbstractRNG, obj::RepGradELBO, adtype::AutoB200, q, prob, params, restructure) =
    make_state(rng, 0, entropy_code(obj.entropy), obj.n_samples, adtype, q, prob, params)
AdvancedVI.init(rng::Random.AbstractRNG, obj::ScoreGradELBO, adtype::AutoB200, q, prob, params, restructure) =
    make_state(rng, 1, 0, obj.n_samples, adtype, q, prob, params)

# set_objective_state_problem (repgradelbo.jl:31-39): used by SubsampledObjective every iteration
function AdvancedVI.set_objective_state_problem(st::B200ObjState, prob_sub::NativeProblem)
    check(@ccall(libavi.avi_obj_set_model(st.h::Ptr{Cvoid}, prob_sub.h::Ptr{Cvoid})::Int32), prob_sub.c.h)
    st.prob = prob_sub
    return st
end

# estimate_gradient! (abstractobjective.jl:67-86; repgradelbo.jl:151-177; scoregradelbo.jl:96-117)
function AdvancedVI.estimate_gradient!(rng::Random.AbstractRNG, obj::Union{RepGradELBO,ScoreGradELBO},
                                       adtype::AutoB200, out::DiffResults.MutableDiffResult, st::B200ObjState,
                                       params, restructure, args...)
    g = DiffResults.gradient(out)
    v = Ref{Float32}(0); e = Ref{Float32}(0)
    check(@ccall(libavi.avi_obj_estimate_gradient(st.h::Ptr{Cvoid}, params::Ptr{Float32}, length(params)::Int64,
                 g::Ptr{Float32}, v::Ptr{Float32}, e::Ptr{Float32})::Int32), st.prob.c.h)
    DiffResults.value!(out, v[])
    return out, st, (elbo=e[],)
end

# estimate_objective (repgradelbo.jl:112-122; scoregradelbo.jl:58-65)
function estimate_objective_b200(rng, st::B200ObjState, params::Vector{Float32}, n_samples, kind, entropy)
    r = Ref{Float32}(0)
    check(@ccall(libavi.avi_obj_estimate_objective(st.h::Ptr{Cvoid}, params::Ptr{Float32}, length(params)::Int64,
                 n_samples::Int32, kind::Int32, entropy::Int32, rand(rng, UInt64)::UInt64, r::Ptr{Float32})::Int32),
          st.prob.c.h)
    return r[]
end

# ---- gaussian_expectation_gradient_and_hessian! (src/algorithms/gauss_expected_grad_hess.jl:20-58) -------------
# Method for native targets: the sampling stage of KLMinWassFwdBwd / KLMinNaturalGradDescent /
# KLMinSqrtNaturalGradDescent (first-order Stein branch) runs on the device; the d x d updates stay in Julia.
function AdvancedVI.gaussian_expectation_gradient_and_hessian!(rng::Random.AbstractRNG,
        q::MvLocationScale{<:LinearAlgebra.AbstractTriangular,<:Normal}, n_samples::Int,
        grad_buf::AbstractVector{Float32}, hess_buf::AbstractMatrix{Float32}, prob::NativeProblem)
    params, _ = Optimisers.destructure(q)
    st = make_state(rng, 0, 0, 1, AutoB200(prob.c.device), q, prob, params)
    lp = Ref{Float32}(0)
    g, H = Vector{Float32}(undef, length(grad_buf)), Matrix{Float32}(undef, size(hess_buf)...)
    check(@ccall(libavi.avi_obj_gauss_expected_grad_hess(st.h::Ptr{Cvoid}, params::Ptr{Float32}, length(params)::Int64,
                 n_samples::Int32, lp::Ptr{Float32}, g::Ptr{Float32}, H::Ptr{Float32})::Int32), prob.c.h)
    grad_buf .= g; hess_buf .= H
    return lp[], grad_buf, hess_buf
end

# ---- optional fast path: the whole `step` on the device (src/algorithms/common.jl:40-120) ------------------
# `init`/`step`/`output` methods for the three ParamSpaceSGD algorithm types when their adtype is AutoB200 and the
# objective is not subsampled: parameters, optimiser state and the averaged iterate stay on the GPU (avi_opt_*),
# one call runs `chunk` iterations without the callback; with a callback `chunk = 1` reproduces the reference loop.
const B200Alg = Union{KLMinRepGradDescent{<:Union{RepGradELBO},AutoB200},
                      KLMinRepGradProxDescent{<:Any,AutoB200},
                      KLMinScoreGradDescent{<:Union{ScoreGradELBO},AutoB200}}
rule_code(o::Optimisers.Descent) = (0, Float32[o.eta])
rule_code(o::Optimisers.Adam) = (1, Float32[o.eta, o.beta[1], o.beta[2], o.epsilon])
rule_code(o::AdvancedVI.DoG) = (2, Float32[o.alpha])
rule_code(o::AdvancedVI.DoWG) = (3, Float32[o.alpha])
op_code(::AdvancedVI.IdentityOperator) = (0, 0.0f0)
op_code(o::AdvancedVI.ClipScale) = (1, Float32(o.epsilon))
op_code(::AdvancedVI.ProximalLocationScaleEntropy) = (2, 0.0f0)
avg_code(::AdvancedVI.NoAveraging) = (0, 0.0f0)
avg_code(a::AdvancedVI.PolynomialAveraging) = (1, Float32(a.eta))

mutable struct B200OptState
    h::Ptr{Cvoid}
    obj_st::B200ObjState
end

function AdvancedVI.init(rng::Random.AbstractRNG, alg::B200Alg, q_init, prob)
    params, re = Optimisers.destructure(q_init)
    obj_st = AdvancedVI.init(rng, alg.objective, alg.adtype, q_init, prob, params, re)
    (rule, hyper), (op, op_param), (avg, avg_param) = rule_code(alg.optimizer), op_code(alg.operator), avg_code(alg.averager)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(@ccall(libavi.avi_opt_create(obj_st.h::Ptr{Cvoid}, rule::Int32, hyper::Ptr{Float32}, length(hyper)::Int32,
                 op::Int32, op_param::Float32, avg::Int32, avg_param::Float32, params::Ptr{Float32},
                 length(params)::Int64, r::Ptr{Ptr{Cvoid}})::Int32), obj_st.prob.c.h)
    st = B200OptState(r[], obj_st)
    finalizer(s -> @ccall(libavi.avi_opt_destroy(s.h::Ptr{Cvoid})::Int32), st)
    return (prob=prob, q=q_init, iteration=0, opt=st, re=re)
end

function AdvancedVI.step(rng::Random.AbstractRNG, alg::B200Alg, state, callback, objargs...; kwargs...)
    v, e, nd = Ref{Float32}(0), Ref{Float32}(0), Ref{Int32}(0)
    c = state.opt.obj_st.prob.c.h
    check(@ccall(libavi.avi_opt_steps(state.opt.h::Ptr{Cvoid}, 1::Int32, v::Ptr{Float32}, e::Ptr{Float32},
                 nd::Ptr{Int32})::Int32), c)
    nd[] == 1 || throw(ErrorException("The objective value is $(v[]). This indicates that the optimization run diverged."))  # common.jl:83-89
    info = (elbo=e[],)
    state = merge(state, (iteration=state.iteration + 1,))
    if !isnothing(callback)
        P = length(first(Optimisers.destructure(state.q)))
        lam, lam_avg, grad = (Vector{Float32}(undef, P) for _ in 1:3)
        check(@ccall(libavi.avi_opt_get(state.opt.h::Ptr{Cvoid}, lam::Ptr{Float32}, lam_avg::Ptr{Float32},
                     grad::Ptr{Float32})::Int32), c)
        info′ = callback(; rng, iteration=state.iteration, restructure=state.re, params=lam, averaged_params=lam_avg,
                         gradient=grad, state=state)
        info = !isnothing(info′) ? merge(info′, info) : info
    end
    return state, false, info
end

function AdvancedVI.output(alg::B200Alg, state)   # common.jl:63-67: re(value(averager, avg_st))
    P = length(first(Optimisers.destructure(state.q)))
    lam_avg = Vector{Float32}(undef, P)
    check(@ccall(libavi.avi_opt_get(state.opt.h::Ptr{Cvoid}, C_NULL::Ptr{Float32}, lam_avg::Ptr{Float32},
                 C_NULL::Ptr{Float32})::Int32), state.opt.obj_st.prob.c.h)
    return state.re(lam_avg)
end

export AutoB200, LogReg, NativeProblem
end # module
