# AdvancedVIB200.jl -- the @ccall glue that plugs libavi_b200.so into AdvancedVI.jl (v0.7) unchanged.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (SURVEY.md F2).  The same
# call sequences are exercised by the Python mirror (advancedvi.jl_b200/api.py) and its GPU tests; this file is
# what a maintainer adds on the Julia side, and julia/test/runtests.jl re-runs the reference's own hot-path tests
# through it.  It dispatches on a new AD-type marker, `AutoB200`, which is the one free field every ParamSpaceSGD
# algorithm threads into the objective methods (src/algorithms/constructors.jl:46,52), so
# `KLMinRepGradDescent(AutoB200(); ...)`, `optimize`, callbacks, `SubsampledObjective` and Turing keep working as
# they are.  Every name used below is bound by one of the `using` lines (checked by hand against the file: Optimisers,
# LinearAlgebra as a module name, Distributions.Normal, the AdvancedVI algorithm / operator / averager types).
module AdvancedVIB200

using AdvancedVI, ADTypes, DiffResults, LogDensityProblems, Random
using Optimisers
using LinearAlgebra
using LinearAlgebra: Diagonal, LowerTriangular
using Distributions: Normal, Laplace, TDist, dof, params
using AdvancedVI: RepGradELBO, ScoreGradELBO, SubsampledObjective, MvLocationScale, MvLocationScaleLowRank,
                  ClosedFormEntropy, MonteCarloEntropy, StickingTheLandingEntropy, ClosedFormEntropyZeroGradient,
                  StickingTheLandingEntropyZeroGradient, KLMinRepGradDescent, KLMinRepGradProxDescent,
                  KLMinScoreGradDescent, IdentityOperator, ClipScale, ProximalLocationScaleEntropy, NoAveraging,
                  PolynomialAveraging, DoG, DoWG

const libavi = get(ENV, "LIBAVI_B200", "libavi_b200.so")

"""
    AutoB200(; device=0, fallback_gradient=central_fd_gradient)

AD-type marker selecting the native path.  `fallback_gradient(f, z) -> Vector` is only used for targets whose
`LogDensityProblems.capabilities` is order 0 under `RepGradELBO`: the reference differentiates through
`logdensity` with its AD backend (src/algorithms/repgradelbo.jl:50-62); the native path has no AD backend, so the
host-callback target obtains the per-sample gradient from this function (default: central finite differences in
Float64; pass e.g. `(f, z) -> ForwardDiff.gradient(f, z)` for an exact one).
"""
struct AutoB200{F} <: ADTypes.AbstractADType
    device::Int
    fallback_gradient::F
end
function central_fd_gradient(f, z::AbstractVector)
    g = similar(z, Float64)
    zz = Vector{Float64}(z)
    for i in eachindex(zz)
        h = 1e-5 * max(1.0, abs(zz[i]))
        zi = zz[i]
        zz[i] = zi + h; fp = f(zz)
        zz[i] = zi - h; fm = f(zz)
        zz[i] = zi
        g[i] = (fp - fm) / (2h)
    end
    return g
end
AutoB200(; device::Integer=0, fallback_gradient=central_fd_gradient) = AutoB200(Int(device), fallback_gradient)

# ---- handles --------------------------------------------------------------------------------------
function check(code::Int32, h)
    code == 0 && return nothing
    msg = unsafe_string(@ccall libavi.avi_last_error(h::Ptr{Cvoid})::Cstring)
    error("libavi_b200 error $code: $msg")
end

mutable struct Ctx
    h::Ptr{Cvoid}
    device::Int
    function Ctx(device::Integer)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(@ccall(libavi.avi_ctx_create(device::Int32, r::Ptr{Ptr{Cvoid}})::Int32), C_NULL)
        c = new(r[], Int(device))
        finalizer(x -> @ccall(libavi.avi_ctx_destroy(x.h::Ptr{Cvoid})::Int32), c)
        return c
    end
end
const CTX = Dict{Int,Ctx}()
ctx(dev::Integer) = get!(() -> Ctx(dev), CTX, Int(dev))

"""A native target: a LogDensityProblem that also carries a device model handle.
`LogDensityProblems.dimension/capabilities/logdensity/logdensity_and_gradient` are defined on it, so the
reference's own CPU path accepts the very same object."""
mutable struct NativeProblem
    h::Ptr{Cvoid}
    c::Ctx
    D::Int
    keepalive::Any      # host-callback targets: the @cfunction closure and the wrapped problem
end

"""`AdvancedVI.subsample(prob, batch)` of a native target: a NEW problem object (the reference never alters `prob`,
src/AdvancedVI.jl:303-313) that shares the parent's device data and carries the 0-based row indices.  The rows are
gathered on the device when the view is evaluated; the parent is switched back to its full data afterwards."""
struct NativeProblemView
    parent::NativeProblem
    idx::Vector{Int32}
end
const NativeTarget = Union{NativeProblem,NativeProblemView}
handle(p::NativeProblem) = p.h
handle(p::NativeProblemView) = p.parent.h
context(p::NativeProblem) = p.c
context(p::NativeProblemView) = p.parent.c

enter_view(::NativeProblem) = nothing
leave_view(::NativeProblem) = nothing
function enter_view(p::NativeProblemView)
    check(@ccall(libavi.avi_model_subsample(p.parent.h::Ptr{Cvoid}, p.idx::Ptr{Int32}, length(p.idx)::Int64)::Int32), p.parent.c.h)
end
function leave_view(p::NativeProblemView)
    check(@ccall(libavi.avi_model_subsample(p.parent.h::Ptr{Cvoid}, C_NULL::Ptr{Int32}, 0::Int64)::Int32), p.parent.c.h)
end

LogDensityProblems.dimension(p::NativeProblem) = p.D
LogDensityProblems.dimension(p::NativeProblemView) = p.parent.D
LogDensityProblems.capabilities(::Type{NativeProblem}) = LogDensityProblems.LogDensityOrder{1}()
LogDensityProblems.capabilities(::Type{NativeProblemView}) = LogDensityProblems.LogDensityOrder{1}()
function LogDensityProblems.logdensity_and_gradient(p::NativeTarget, z::AbstractVector)
    D = LogDensityProblems.dimension(p)
    zf = Vector{Float32}(z); lp = Ref{Float32}(0); g = Vector{Float32}(undef, D)
    enter_view(p)
    try
        check(@ccall(libavi.avi_model_logdensity_and_gradient_host(handle(p)::Ptr{Cvoid}, zf::Ptr{Float32}, 1::Int32,
                     lp::Ptr{Float32}, g::Ptr{Float32})::Int32), context(p).h)
    finally
        leave_view(p)
    end
    return lp[], g
end
LogDensityProblems.logdensity(p::NativeTarget, z) = first(LogDensityProblems.logdensity_and_gradient(p, z))

# AdvancedVI.subsample(prob, batch) (src/AdvancedVI.jl:303-313): 1-based Julia indices -> 0-based rows
AdvancedVI.subsample(p::NativeProblem, batch) = NativeProblemView(p, Int32.(collect(batch) .- 1))
AdvancedVI.subsample(p::NativeProblemView, batch) = NativeProblemView(p.parent, p.idx[collect(batch)])

"Hierarchical logistic regression of docs/src/tutorials/subsampling.md:26-38 (variant = :subsampling) or README.md:47-58 (:basic); `gaussian=true`: the Gaussian GLM of BASELINE.json config 4.  gemm: 0 exact fp32 SIMT, 1 TF32 tensor cores, 2 3xTF32 (fp32-grade)."
function LogReg(X::Matrix{Float32}, y::Vector{Float32}; n_data=size(X, 1), variant=:subsampling, gaussian=false,
                gemm=1, device=0)
    c = ctx(device); r = Ref{Ptr{Cvoid}}(C_NULL)
    check(@ccall(libavi.avi_model_glm_create(c.h::Ptr{Cvoid}, X::Ptr{Float32}, y::Ptr{Float32}, size(X, 1)::Int64,
                 size(X, 2)::Int32, n_data::Int64, (gaussian ? 1 : 0)::Int32,
                 (variant === :subsampling ? 0 : 1)::Int32, gemm::Int32, r::Ptr{Ptr{Cvoid}})::Int32), c.h)
    p = NativeProblem(r[], c, size(X, 2) + 1, nothing)
    finalizer(q -> @ccall(libavi.avi_model_destroy(q.h::Ptr{Cvoid})::Int32), p)
    return p
end

"`logpdf(MvNormal(mu, Diagonal(sigma.^2)), z)` as a native target (test/models/normal.jl:8-11)."
function MvNormalDiag(mu::Vector{Float32}, sigma::Vector{Float32}; device=0)
    c = ctx(device); r = Ref{Ptr{Cvoid}}(C_NULL)
    check(@ccall(libavi.avi_model_mvnormal_diag_create(c.h::Ptr{Cvoid}, mu::Ptr{Float32}, sigma::Ptr{Float32},
                 length(mu)::Int32, r::Ptr{Ptr{Cvoid}})::Int32), c.h)
    p = NativeProblem(r[], c, length(mu), nothing)
    finalizer(q -> @ccall(libavi.avi_model_destroy(q.h::Ptr{Cvoid})::Int32), p)
    return p
end

"""Any other LogDensityProblem (DynamicPPL, BridgeStan, ...) goes through the per-sample host callback
(src/algorithms/repgradelbo.jl:84-86 shape: one call per Monte-Carlo sample).  A capability-0 target is presented to
the library as first-order, its gradient coming from `adtype.fallback_gradient` (see `AutoB200`)."""
function hostcallback_problem(prob, adtype::AutoB200)
    D = LogDensityProblems.dimension(prob)
    cap0 = LogDensityProblems.capabilities(typeof(prob)) isa LogDensityProblems.LogDensityOrder{0}
    if cap0
        @info "The capability of the supplied `LogDensityProblem` is order 0: AutoB200 has no AD backend to differentiate through `LogDensityProblems.logdensity`; per-sample gradients come from `AutoB200(; fallback_gradient)` (default: central finite differences)."
    end
    f = Base.Fix1(LogDensityProblems.logdensity, prob)
    function cb(user::Ptr{Cvoid}, z::Ptr{Float32}, Dn::Int32, lp::Ptr{Float32}, g::Ptr{Float32})::Int32
        try
            zz = Vector{Float64}(unsafe_wrap(Array, z, Int(Dn)))
            if g == C_NULL
                unsafe_store!(lp, Float32(f(zz)))
            elseif cap0
                unsafe_store!(lp, Float32(f(zz)))
                gr = Vector{Float32}(adtype.fallback_gradient(f, zz))
                unsafe_copyto!(g, pointer(gr), Int(Dn))
            else
                l, grd = LogDensityProblems.logdensity_and_gradient(prob, zz)
                gr = Vector{Float32}(grd)
                unsafe_store!(lp, Float32(l)); unsafe_copyto!(g, pointer(gr), Int(Dn))
            end
            return Int32(0)
        catch
            return Int32(1)
        end
    end
    c = ctx(adtype.device); r = Ref{Ptr{Cvoid}}(C_NULL)
    fp = @cfunction($cb, Int32, (Ptr{Cvoid}, Ptr{Float32}, Int32, Ptr{Float32}, Ptr{Float32}))
    check(@ccall(libavi.avi_model_hostcallback_create(c.h::Ptr{Cvoid}, D::Int32, 1::Int32, fp::Ptr{Cvoid},
                 C_NULL::Ptr{Cvoid}, r::Ptr{Ptr{Cvoid}})::Int32), c.h)
    p = NativeProblem(r[], c, D, (fp, prob))      # keeps the closure and the wrapped problem alive
    finalizer(q -> @ccall(libavi.avi_model_destroy(q.h::Ptr{Cvoid})::Int32), p)
    return p
end
native(prob::NativeTarget, ::AutoB200) = prob
native(prob, adtype::AutoB200) = hostcallback_problem(prob, adtype)

# ---- objective state ---------------------------------------------------------------------------------
mutable struct B200ObjState
    h::Ptr{Cvoid}
    prob::NativeTarget
end
family_code(q::MvLocationScale{<:Diagonal}) = 0
family_code(q::MvLocationScale{<:LowerTriangular}) = 1
entropy_code(::ClosedFormEntropy) = 0
entropy_code(::MonteCarloEntropy) = 1
entropy_code(::StickingTheLandingEntropy) = 2
entropy_code(::ClosedFormEntropyZeroGradient) = 3
entropy_code(::StickingTheLandingEntropyZeroGradient) = 4
kind_code(::RepGradELBO) = 0
kind_code(::ScoreGradELBO) = 1
# base distribution `dist` of MvLocationScale(location, scale, dist) (location_scale.jl:15-19; docs/src/families.md:72-101)
# -> (AVI_BASE_* code, parameter) of avi_obj_set_base
base_code(d::Normal) = params(d) == (0, 1) ? (Int32(0), 0.0f0) :
    throw(ArgumentError("AutoB200: the Normal base distribution must be Normal(0, 1) (got $d)"))
base_code(d::Laplace) = params(d) == (0, 1) ? (Int32(1), 0.0f0) :
    throw(ArgumentError("AutoB200: the Laplace base distribution must be Laplace(0, 1) (got $d)"))
base_code(d::TDist) = (Int32(2), Float32(dof(d)))
base_code(d) = throw(ArgumentError("AutoB200 supports the base distributions Normal(0, 1), Laplace(0, 1) and TDist(nu) (got $d)"))
objective_entropy_code(obj::RepGradELBO) = entropy_code(obj.entropy)
objective_entropy_code(::ScoreGradELBO) = 0

function make_state(key::UInt64, kind, entropy, n_samples, q, p::NativeTarget, params)
    eltype(params) === Float32 || throw(ArgumentError("AutoB200 supports Float32 only (got $(eltype(params)))"))
    c = context(p)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    if q isa MvLocationScaleLowRank   # location_scale_low_rank.jl: lambda = [location; scale_diag; vec(scale_factors)]
        check(@ccall(libavi.avi_obj_create_lowrank(c.h::Ptr{Cvoid}, handle(p)::Ptr{Cvoid}, size(q.scale_factors, 2)::Int32,
                     kind::Int32, entropy::Int32, n_samples::Int32, r::Ptr{Ptr{Cvoid}})::Int32), c.h)
    else
        check(@ccall(libavi.avi_obj_create(c.h::Ptr{Cvoid}, handle(p)::Ptr{Cvoid}, family_code(q)::Int32, kind::Int32,
                     entropy::Int32, n_samples::Int32, r::Ptr{Ptr{Cvoid}})::Int32), c.h)
    end
    st = B200ObjState(r[], p)
    finalizer(s -> @ccall(libavi.avi_obj_destroy(s.h::Ptr{Cvoid})::Int32), st)
    if q isa MvLocationScale
        bc, bp = base_code(q.dist)
        bc == 0 || check(@ccall(libavi.avi_obj_set_base(st.h::Ptr{Cvoid}, bc::Int32, bp::Float32)::Int32), c.h)
    end
    # the Julia rng is used only to draw the Philox key: same seed => identical run (klminrepgraddescent.jl:40-57)
    check(@ccall(libavi.avi_obj_seed(st.h::Ptr{Cvoid}, key::UInt64, 0::UInt64)::Int32), c.h)
    return st
end

# init (src/algorithms/abstractobjective.jl:25-35; repgradelbo.jl:41-70; scoregradelbo.jl:34-50)
function AdvancedVI.init(rng::Random.AbstractRNG, obj::Union{RepGradELBO,ScoreGradELBO}, adtype::AutoB200, q, prob,
                         params, restructure)
    return make_state(rand(rng, UInt64), kind_code(obj), objective_entropy_code(obj), obj.n_samples, q,
                      native(prob, adtype), params)
end

# set_objective_state_problem (repgradelbo.jl:31-39): used by SubsampledObjective every iteration
function AdvancedVI.set_objective_state_problem(st::B200ObjState, prob_sub::NativeTarget)
    handle(prob_sub) == handle(st.prob) || check(@ccall(libavi.avi_obj_set_model(st.h::Ptr{Cvoid},
                                                        handle(prob_sub)::Ptr{Cvoid})::Int32), context(prob_sub).h)
    st.prob = prob_sub
    return st
end

# estimate_gradient! (abstractobjective.jl:67-86; repgradelbo.jl:151-177; scoregradelbo.jl:96-117)
function AdvancedVI.estimate_gradient!(rng::Random.AbstractRNG, obj::Union{RepGradELBO,ScoreGradELBO},
                                       adtype::AutoB200, out::DiffResults.MutableDiffResult, st::B200ObjState,
                                       params, restructure, args...)
    g = DiffResults.gradient(out)
    v = Ref{Float32}(0); e = Ref{Float32}(0)
    enter_view(st.prob)
    try
        check(@ccall(libavi.avi_obj_estimate_gradient(st.h::Ptr{Cvoid}, params::Ptr{Float32}, length(params)::Int64,
                     g::Ptr{Float32}, v::Ptr{Float32}, e::Ptr{Float32})::Int32), context(st.prob).h)
    finally
        leave_view(st.prob)
    end
    out = DiffResults.value!(out, v[])
    return out, st, (elbo=e[],)
end

# estimate_objective (abstractobjective.jl:38-55; repgradelbo.jl:112-122; scoregradelbo.jl:58-65).  The reference
# methods take no adtype, so the native ones dispatch on the TARGET being native.  The algorithm-level method
# (common.jl:29-38) and SubsampledObjective's epoch mean (subsampledobjective.jl:47-58) are the reference's own generic
# code: they call `subsample` (-> NativeProblemView) and this method, nothing else is needed.
function AdvancedVI.estimate_objective(rng::Random.AbstractRNG, obj::Union{RepGradELBO,ScoreGradELBO},
                                       q::Union{MvLocationScale,MvLocationScaleLowRank}, prob::NativeTarget;
                                       n_samples::Int=obj.n_samples)
    params, _ = Optimisers.destructure(q)
    st = make_state(UInt64(0), kind_code(obj), objective_entropy_code(obj), obj.n_samples, q, prob, params)
    r = Ref{Float32}(0)
    enter_view(prob)
    try
        check(@ccall(libavi.avi_obj_estimate_objective(st.h::Ptr{Cvoid}, params::Ptr{Float32}, length(params)::Int64,
                     n_samples::Int32, kind_code(obj)::Int32, objective_entropy_code(obj)::Int32,
                     rand(rng, UInt64)::UInt64, r::Ptr{Float32})::Int32), context(prob).h)
    finally
        leave_view(prob)
    end
    return r[]
end

# ---- gaussian_expectation_gradient_and_hessian! (src/algorithms/gauss_expected_grad_hess.jl:20-58) -------------
# Method for native targets: the sampling stage of KLMinWassFwdBwd / KLMinNaturalGradDescent /
# KLMinSqrtNaturalGradDescent (first-order Stein branch) runs on the device; the d x d updates stay in Julia.
function AdvancedVI.gaussian_expectation_gradient_and_hessian!(rng::Random.AbstractRNG,
        q::MvLocationScale{<:LinearAlgebra.AbstractTriangular,<:Normal}, n_samples::Int,
        grad_buf::AbstractVector{Float32}, hess_buf::AbstractMatrix{Float32}, prob::NativeProblem)
    params, _ = Optimisers.destructure(q)
    st = make_state(rand(rng, UInt64), 0, 0, 1, q, prob, params)
    lp = Ref{Float32}(0)
    g, H = Vector{Float32}(undef, length(grad_buf)), Matrix{Float32}(undef, size(hess_buf)...)
    check(@ccall(libavi.avi_obj_gauss_expected_grad_hess(st.h::Ptr{Cvoid}, params::Ptr{Float32}, length(params)::Int64,
                 n_samples::Int32, lp::Ptr{Float32}, g::Ptr{Float32}, H::Ptr{Float32})::Int32), prob.c.h)
    grad_buf .= g; hess_buf .= H
    return lp[], grad_buf, hess_buf
end

# sampling stage of FisherMinBatchMatch (src/algorithms/fisherminbatchmatch.jl:81-111) over a native target
function AdvancedVI.rand_batch_match_samples_with_objective!(rng::Random.AbstractRNG,
        q::MvLocationScale{<:LinearAlgebra.AbstractTriangular,<:Normal}, n_samples::Int, prob::NativeProblem,
        u_buf::AbstractMatrix{Float32}=Matrix{Float32}(undef, length(q.location), n_samples),
        grad_buf::AbstractMatrix{Float32}=Matrix{Float32}(undef, length(q.location), n_samples))
    params, _ = Optimisers.destructure(q)
    st = make_state(rand(rng, UInt64), 0, 0, 1, q, prob, params)
    d = length(q.location)
    u, z, g = (Matrix{Float32}(undef, d, n_samples) for _ in 1:3)
    fisher, lp = Ref{Float32}(0), Ref{Float32}(0)
    check(@ccall(libavi.avi_obj_batch_match_samples(st.h::Ptr{Cvoid}, params::Ptr{Float32}, length(params)::Int64,
                 n_samples::Int32, u::Ptr{Float32}, z::Ptr{Float32}, g::Ptr{Float32}, fisher::Ptr{Float32},
                 lp::Ptr{Float32})::Int32), prob.c.h)
    u_buf .= u; grad_buf .= g
    return u_buf, z, grad_buf, fisher[], lp[]
end

# ---- optional fast path: the whole `step` on the device (src/algorithms/common.jl:40-120) ------------------
# `init`/`step`/`output` methods for the three ParamSpaceSGD algorithm types when their adtype is AutoB200 and the
# objective is not subsampled: parameters, optimiser state and the averaged iterate stay on the GPU (avi_opt_*); one
# call = one iteration = one kernel launch for the GLM targets.  `state.q` is refreshed only when a callback is
# installed (that is when the parameters cross to the host anyway); `output` always reads the device.
const B200Alg = Union{KLMinRepGradDescent{<:RepGradELBO,<:AutoB200},
                      KLMinRepGradProxDescent{<:RepGradELBO,<:AutoB200},
                      KLMinScoreGradDescent{<:ScoreGradELBO,<:AutoB200}}
rule_code(o::Optimisers.Descent) = (0, Float32[o.eta])
rule_code(o::Optimisers.Adam) = (1, Float32[o.eta, o.beta[1], o.beta[2], o.epsilon])
rule_code(o::DoG) = (2, Float32[o.alpha])
rule_code(o::DoWG) = (3, Float32[o.alpha])
op_code(::IdentityOperator) = (0, 0.0f0)
op_code(o::ClipScale) = (1, Float32(o.epsilon))
op_code(::ProximalLocationScaleEntropy) = (2, 0.0f0)
avg_code(::NoAveraging) = (0, 0.0f0)
avg_code(a::PolynomialAveraging) = (1, Float32(a.eta))

mutable struct B200OptState
    h::Ptr{Cvoid}
    obj_st::B200ObjState
    P::Int
end

function AdvancedVI.init(rng::Random.AbstractRNG, alg::B200Alg, q_init, prob)
    if q_init isa MvLocationScale && alg.operator isa IdentityOperator
        @warn "IdentityOperator is used with a variational family <:MvLocationScale. Optimization can easily fail under this combination due to singular scale matrices. Consider using the operator `ClipScale` in the algorithm instead."   # common.jl:42-46
    end
    params, re = Optimisers.destructure(q_init)
    obj_st = AdvancedVI.init(rng, alg.objective, alg.adtype, q_init, prob, params, re)
    (rule, hyper), (op, op_param), (avg, avg_param) = rule_code(alg.optimizer), op_code(alg.operator), avg_code(alg.averager)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(@ccall(libavi.avi_opt_create(obj_st.h::Ptr{Cvoid}, rule::Int32, hyper::Ptr{Float32}, length(hyper)::Int32,
                 op::Int32, op_param::Float32, avg::Int32, avg_param::Float32, params::Ptr{Float32},
                 length(params)::Int64, r::Ptr{Ptr{Cvoid}})::Int32), context(obj_st.prob).h)
    st = B200OptState(r[], obj_st, length(params))
    finalizer(s -> @ccall(libavi.avi_opt_destroy(s.h::Ptr{Cvoid})::Int32), st)
    return (prob=prob, q=q_init, iteration=0, opt=st, re=re)
end

function AdvancedVI.step(rng::Random.AbstractRNG, alg::B200Alg, state, callback, objargs...; kwargs...)
    v, e, nd = Ref{Float32}(0), Ref{Float32}(0), Ref{Int32}(0)
    c = context(state.opt.obj_st.prob).h
    check(@ccall(libavi.avi_opt_steps(state.opt.h::Ptr{Cvoid}, 1::Int32, v::Ptr{Float32}, e::Ptr{Float32},
                 nd::Ptr{Int32})::Int32), c)
    nd[] == 1 || throw(ErrorException("The objective value is $(v[]). This indicates that the optimization run diverged."))  # common.jl:83-89
    info = (elbo=e[],)
    state = merge(state, (iteration=state.iteration + 1,))
    if !isnothing(callback)
        P = state.opt.P
        lam, lam_avg, grad = Vector{Float32}(undef, P), Vector{Float32}(undef, P), Vector{Float32}(undef, P)
        check(@ccall(libavi.avi_opt_get(state.opt.h::Ptr{Cvoid}, lam::Ptr{Float32}, lam_avg::Ptr{Float32},
                     grad::Ptr{Float32})::Int32), c)
        state = merge(state, (q=state.re(lam),))
        info′ = callback(; rng, iteration=state.iteration, restructure=state.re, params=lam, averaged_params=lam_avg,
                         gradient=grad, state=state)
        info = !isnothing(info′) ? merge(info′, info) : info
    end
    return state, false, info
end

function AdvancedVI.output(alg::B200Alg, state)   # common.jl:63-67: re(value(averager, avg_st))
    lam_avg = Vector{Float32}(undef, state.opt.P)
    check(@ccall(libavi.avi_opt_get(state.opt.h::Ptr{Cvoid}, C_NULL::Ptr{Float32}, lam_avg::Ptr{Float32},
                 C_NULL::Ptr{Float32})::Int32), context(state.opt.obj_st.prob).h)
    return state.re(lam_avg)
end

export AutoB200, LogReg, MvNormalDiag, NativeProblem, NativeProblemView
end # module
