"""ELBO objectives: RepGradELBO (+ entropy estimators), ScoreGradELBO, SubsampledObjective.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference differentiates a forward closure with an AD backend
(src/AdvancedVI.jl:47-98).  Here the forward closures are restated 1:1
(`repgrad_forward`, `scoregrad_forward`) and the gradients are the closed forms of
SURVEY.md Appendix A; tests/test_oracle_gradients.py checks every closed form against
central finite differences of the restated forward with eps held fixed, which is what
an AD backend would return.

eps is an explicit argument everywhere: `rand(rng, q, M)` of the reference
(src/algorithms/repgradelbo.jl:107) becomes `q.rand_from_eps(eps)`.
"""

from __future__ import annotations

import numpy as np

from .family import MvLocationScale

ENTROPIES = ("ClosedFormEntropy", "MonteCarloEntropy", "StickingTheLandingEntropy",
             "ClosedFormEntropyZeroGradient", "StickingTheLandingEntropyZeroGradient")


# -- src/algorithms/entropy.jl -------------------------------------------------------------
def estimate_entropy(kind: str, samples: np.ndarray, q: MvLocationScale, q_stop: MvLocationScale):
    if kind == "ClosedFormEntropyZeroGradient":          # entropy.jl:13-15
        return q_stop.entropy()
    if kind == "ClosedFormEntropy":                      # entropy.jl:27-29
        return q.entropy()
    if kind == "MonteCarloEntropy":                      # entropy.jl:42-46
        return np.mean(-q.logpdf(samples))
    if kind == "StickingTheLandingEntropy":              # entropy.jl:59-65
        return np.mean(-q_stop.logpdf(samples))
    if kind == "StickingTheLandingEntropyZeroGradient":  # entropy.jl:80-90
        return np.mean(-q_stop.logpdf(samples)) - q.entropy() + q_stop.entropy()
    raise ValueError(kind)


# -- src/algorithms/repgradelbo.jl ---------------------------------------------------------
def estimate_energy_with_samples(prob, samples: np.ndarray, per_sample: bool = False):
    """mean(logdensity(prob, z_m) for each column)  (repgradelbo.jl:84-86).
    per_sample=True loops over columns exactly like the reference (M separate calls)."""
    if per_sample:
        return np.mean([prob.logdensity(samples[:, m]) for m in range(samples.shape[1])])
    return np.mean(prob.logdensity_batch(samples))


def repgrad_forward(params, q_template: MvLocationScale, q_stop: MvLocationScale, prob,
                    eps: np.ndarray, entropy: str):
    """estimate_repgradelbo_ad_forward (repgradelbo.jl:142-149): returns -ELBO."""
    q = q_template.restructure(params)
    samples = q.rand_from_eps(eps)                               # reparam_with_entropy :104-110
    ent = estimate_entropy(entropy, samples, q, q_stop)
    energy = estimate_energy_with_samples(prob, samples)
    return -(energy + ent)


def _scale_grad_pack(q: MvLocationScale, g_mu, g_scale):
    if q.is_meanfield:
        return np.concatenate([g_mu, g_scale])
    return np.concatenate([g_mu, np.tril(g_scale).reshape(-1, order="F")])


def repgrad_value_and_gradient(params, q_template: MvLocationScale, prob, eps: np.ndarray,
                               entropy: str, per_sample: bool = False):
    """estimate_gradient! for RepGradELBO (repgradelbo.jl:151-177) with q_stop =
    restructure(params) (:162).  Returns (value = -ELBO, gradient, elbo).
    Closed forms: SURVEY.md Appendix A.1-A.3."""
    q = q_template.restructure(params)
    M = eps.shape[1]
    Z = q.rand_from_eps(eps)
    if per_sample:   # reference-shaped: one logdensity_and_gradient call per column
        lg = [prob.logdensity_and_gradient(Z[:, m]) for m in range(M)]
        logp = np.array([a for a, _ in lg]); G = np.stack([b for _, b in lg], axis=1)
    else:
        logp, G = prob.logdensity_and_gradient_batch(Z)
    ent = estimate_entropy(entropy, Z, q, q)
    value = -(np.mean(logp) + ent)
    sd = q.scale_diag()
    stl = entropy in ("StickingTheLandingEntropy", "StickingTheLandingEntropyZeroGradient")
    if stl:
        # w_m = g_m - grad_z log q_stop(z_m) = g_m - L^{-T} score(eps_m), score = d log phi / du of the base
        # distribution: -u for Normal(0, 1), i.e. the g_m + L^{-T} eps_m of Appendix A.3
        sc = -q.dist.score(eps)
        if q.is_meanfield:
            W = G + sc / sd[:, None]
        else:
            from scipy.linalg import solve_triangular
            W = G + solve_triangular(q.scale.T, sc, lower=False)
    else:
        W = G
    g_mu = -np.mean(W, axis=1)
    if q.is_meanfield:
        g_sc = -np.mean(W * eps, axis=1)
        inv = 1.0 / sd
    else:
        g_sc = -np.tril(W @ eps.T) / M
        inv = np.diag(1.0 / sd)
    if entropy in ("ClosedFormEntropy", "MonteCarloEntropy"):
        g_sc = g_sc - inv                       # -grad H(q)
    elif entropy == "StickingTheLandingEntropyZeroGradient":
        g_sc = g_sc + inv                       # value has -H(q) added: A.3 + grad H(q)
    return value, _scale_grad_pack(q, g_mu, g_sc), -value


def repgrad_estimate_objective(q: MvLocationScale, prob, eps, entropy="ClosedFormEntropy"):
    """estimate_objective(rng, obj::RepGradELBO, q, prob) (repgradelbo.jl:112-118)."""
    samples = q.rand_from_eps(eps)
    ent = estimate_entropy(entropy, samples, q, q)
    return -(estimate_energy_with_samples(prob, samples) + ent)


# -- src/algorithms/scoregradelbo.jl -------------------------------------------------------
def scoregrad_forward(params, q_template: MvLocationScale, samples_stop, logprob_stop):
    """estimate_scoregradelbo_ad_forward (scoregradelbo.jl:87-94): VarGrad value."""
    q = q_template.restructure(params)
    f = q.logpdf(samples_stop) - logprob_stop
    return (np.mean(f * f) - np.mean(f) ** 2) / 2


def scoregrad_value_and_gradient(params, q_template: MvLocationScale, prob, eps):
    """estimate_gradient! for ScoreGradELBO (scoregradelbo.jl:96-117).
    Returns (value = VarGrad, gradient, elbo).  Closed form: Appendix A.4."""
    q = q_template.restructure(params)
    M = eps.shape[1]
    Z = q.rand_from_eps(eps)                              # :107 (not differentiated)
    logpi = prob.logdensity_batch(Z)                      # :108
    logq = q.logpdf(Z)
    f = logq - logpi
    value = (np.mean(f * f) - np.mean(f) ** 2) / 2
    c = f - np.mean(f)
    u = q.standardize(Z)                                  # == eps up to rounding
    s = -q.dist.score(u)                                  # -d log phi / du of the base distribution: u for Normal(0, 1)
    sd = q.scale_diag()
    if q.is_meanfield:
        g_mu = np.mean(c[None, :] * s, axis=1) / sd
        g_sc = np.mean(c[None, :] * (s * u - 1.0), axis=1) / sd
    else:
        from scipy.linalg import solve_triangular
        # d logq/d mu = L^{-T} s ; d logq/d L = tril(L^{-T} s u') - diag(1/L_ii)
        Linv_T_u = solve_triangular(q.scale.T, s, lower=False)
        g_mu = np.mean(c[None, :] * Linv_T_u, axis=1)
        S = (s * c[None, :]) @ u.T / M
        g_sc = np.tril(solve_triangular(q.scale.T, S, lower=False)) - np.mean(c) * np.diag(1.0 / sd)
    elbo = np.mean(logpi - logq)                          # :113-114
    return value, _scale_grad_pack(q, g_mu, g_sc), elbo


def scoregrad_estimate_objective(q: MvLocationScale, prob, eps):
    """estimate_objective(rng, obj::ScoreGradELBO, ...) (scoregradelbo.jl:58-65)."""
    Z = q.rand_from_eps(eps)
    return -np.mean(prob.logdensity_batch(Z) - q.logpdf(Z))


# -- src/algorithms/common.jl:29-38 --------------------------------------------------------
def estimate_objective(q: MvLocationScale, prob, eps, entropy="MonteCarloEntropy"):
    """estimate_objective(rng, alg::ParamSpaceSGD, q, prob; n_samples, entropy): always a
    fresh RepGradELBO with MonteCarloEntropy by default, ignoring subsampling."""
    return repgrad_estimate_objective(q, prob, eps, entropy)


def gaussian_expectation_gradient_and_hessian(q: MvLocationScale, prob, u):
    """gaussian_expectation_gradient_and_hessian! (src/algorithms/gauss_expected_grad_hess.jl:20-58) for given standard
    normal draws u (D x n).  First-order targets (:33-58, Stein / Price identity): z = C u + m,
    hess = C' \\ mean_b(u_b grad log pi(z_b)').  Targets that also provide `hessian_batch` (second order, :59-80):
    plain sample average of the Hessians.  Returns (logpi_avg, grad, hess)."""
    from scipy.linalg import solve_triangular
    n = u.shape[1]
    m, Cs = q.location, (np.diag(q.scale) if q.is_meanfield else q.scale)
    z = Cs @ u + m[:, None]                                             # :44-45
    logp, G = prob.logdensity_and_gradient_batch(z)
    logpi_avg = float(np.sum(logp / n))                                 # :49
    grad = np.sum(G / n, axis=1)                                        # :51-54
    if getattr(prob, "capability", 1) >= 2 and hasattr(prob, "hessian_batch"):
        hess = np.sum(prob.hessian_batch(z) / n, axis=0)                # :62-77
        return logpi_avg, grad, hess
    A = u @ (G / n).T                                                   # :55  sum_b u_b (grad_b / n)'
    hess = solve_triangular(Cs.T, A, lower=False)                       # :57  C' \\ A
    return logpi_avg, grad, hess


def rand_batch_match_samples_with_objective(q: MvLocationScale, prob, u):
    """rand_batch_match_samples_with_objective! (src/algorithms/fisherminbatchmatch.jl:81-111) for given standard normal
    draws u (D x n): z = C u + mu (:91), grad_buf[:, b] = grad log pi(z_b) and the running log pi sum (:93-98), the
    Fisher-divergence estimate sum |-u - C' grad|^2 / n (:100-108).  Returns (u, z, grad, fisher, logpi_avg)."""
    n = u.shape[1]
    mu, Cs = q.location, (np.diag(q.scale) if q.is_meanfield else q.scale)
    z = Cs @ u + mu[:, None]
    logp, G = prob.logdensity_and_gradient_batch(z)
    fisher = float(np.sum((-u - Cs.T @ G) ** 2) / n)
    return u, z, G, fisher, float(np.sum(logp) / n)


def _lowrank_logq_param_grads(q, Z):
    """Per-sample gradients of log q_lambda(z) w.r.t. lambda at FIXED z for the low-rank Gaussian:
    w = Sigma^-1 (z - mu);  d/d mu = w,  d/d D_i = D_i (w_i^2 - (Sigma^-1)_ii),  d/d U = w (w' U) - Sigma^-1 U.
    Returns (w (d, M), gD (d, M), gU (d, r, M))."""
    w = q.cov_solve(Z - q.location[:, None])
    sinv_diag, sinv_U = q.cov_inv_diag_and_factor()
    gD = q.scale_diag[:, None] * (w * w - sinv_diag[:, None])
    gU = w[:, None, :] * (q.scale_factors.T @ w)[None, :, :] - sinv_U[:, :, None]
    return w, gD, gU


def repgrad_lowrank_value_and_gradient(params, q_template, prob, u_diag, u_fact, entropy="ClosedFormEntropy"):
    """estimate_gradient! for RepGradELBO over MvLocationScaleLowRank (repgradelbo.jl:142-177 with the sampling
    path of location_scale_low_rank.jl:79-86), closed forms.  With g_m = grad log pi(z_m) and the path Jacobian
    dz/d(location, scale_diag, scale_factors) = (I, diag(u_diag), . u_fact'):
      ClosedFormEntropy          d = -mean J' g - dH          (dH: MvLocationScaleLowRank.entropy_gradient)
      StickingTheLandingEntropy  d = -mean J' (g + w),  w = Sigma^-1 (z - mu)   (q frozen inside log q, entropy.jl:59-65)
      MonteCarloEntropy          the STL path term plus mean d log q / d lambda at fixed z (entropy.jl:42-46)
    Returns (value = -ELBO estimate, gradient in destructure order, elbo).  Device path: family_lr.cu."""
    q = q_template.restructure(params)
    M = u_diag.shape[1]
    Z = q.rand_from_eps(u_diag, u_fact)
    logp, G = prob.logdensity_and_gradient_batch(Z)
    if entropy == "ClosedFormEntropy":
        ent = q.entropy()
        gD_H, gU_H = q.entropy_gradient()
        Wg = G
        g_loc_x, g_diag_x, g_fact_x = 0.0, -gD_H, -gU_H
    elif entropy == "ClosedFormEntropyZeroGradient":     # entropy.jl:13-15: the value of H, no gradient through it
        ent = q.entropy()
        Wg = G
        g_loc_x, g_diag_x, g_fact_x = 0.0, 0.0, 0.0
    elif entropy in ("StickingTheLandingEntropy", "MonteCarloEntropy", "StickingTheLandingEntropyZeroGradient"):
        ent = -float(np.mean(q.logpdf(Z)))
        w, gD, gU = _lowrank_logq_param_grads(q, Z)
        Wg = G + w
        if entropy == "MonteCarloEntropy":   # + d/d lambda of mean log q_lambda(z) at fixed z (the value has -H_MC = +mean log q)
            g_loc_x, g_diag_x, g_fact_x = np.mean(w, axis=1), np.mean(gD, axis=1), np.mean(gU, axis=2)
        elif entropy == "StickingTheLandingEntropyZeroGradient":
            # entropy.jl:80-90: STL - H(q) + H(q_stop): the value is STL's, the gradient gains +grad H(q)
            gD_H, gU_H = q.entropy_gradient()
            g_loc_x, g_diag_x, g_fact_x = 0.0, gD_H, gU_H
        else:
            g_loc_x, g_diag_x, g_fact_x = 0.0, 0.0, 0.0
    else:
        raise ValueError(entropy)
    elbo = float(np.mean(logp) + ent)
    g_loc = -np.mean(Wg, axis=1) + g_loc_x
    g_diag = -np.mean(Wg * u_diag, axis=1) + g_diag_x
    g_fact = -(Wg @ u_fact.T) / M + g_fact_x
    return -elbo, np.concatenate([g_loc, g_diag, g_fact.reshape(-1, order="F")]), elbo


def scoregrad_lowrank_value_and_gradient(params, q_template, prob, u_diag, u_fact):
    """estimate_gradient! for ScoreGradELBO (VarGrad, scoregradelbo.jl:87-117) over MvLocationScaleLowRank:
    f_m = log q_lambda(z_m) - log pi(z_m) with z, log pi constants; value = (mean f^2 - (mean f)^2) / 2,
    gradient = mean (f_m - fbar) d log q_lambda(z_m) / d lambda.  Returns (value, gradient, elbo).  Device path:
    family_lr.cu."""
    q = q_template.restructure(params)
    Z = q.rand_from_eps(u_diag, u_fact)
    logp = prob.logdensity_batch(Z) if hasattr(prob, "logdensity_batch") else prob.logdensity_and_gradient_batch(Z)[0]
    f = q.logpdf(Z) - logp
    c = f - np.mean(f)
    w, gD, gU = _lowrank_logq_param_grads(q, Z)
    g_loc = np.mean(c[None, :] * w, axis=1)
    g_diag = np.mean(c[None, :] * gD, axis=1)
    g_fact = np.mean(c[None, None, :] * gU, axis=2)
    value = (np.mean(f * f) - np.mean(f) ** 2) / 2
    return value, np.concatenate([g_loc, g_diag, g_fact.reshape(-1, order="F")]), float(np.mean(logp - q.logpdf(Z)))


def repgrad_general_base_value_and_gradient(params, q_template: MvLocationScale, prob, u: np.ndarray, entropy: str):
    """RepGradELBO over MvLocationScale(location, scale, dist) with a NON-Gaussian base distribution
    (docs/src/families.md:72-101: TDist, Laplace): u are draws of the base distribution, z = scale u + location.
    Closed forms: the energy part is A.1 with eps -> u; ClosedFormEntropy adds -grad logdet(scale);
    StickingTheLandingEntropy replaces g by g - grad_z log q_stop(z) = g - scale^-T score(u), score = d log phi / du
    (for Normal(0, 1): score = -u, i.e. the g + L^-T eps of A.3).  Groundwork: no device path."""
    q = q_template.restructure(params)
    M = u.shape[1]
    Z = q.rand_from_eps(u)
    logp, G = prob.logdensity_and_gradient_batch(Z)
    sd = q.scale_diag()
    if entropy == "ClosedFormEntropy":
        ent, W = q.entropy(), G
    elif entropy == "StickingTheLandingEntropy":
        ent = -float(np.mean(q.logpdf(Z)))
        sc = q.dist.score(u)
        if q.is_meanfield:
            W = G - sc / sd[:, None]
        else:
            from scipy.linalg import solve_triangular
            W = G - solve_triangular(q.scale.T, sc, lower=False)
    else:
        raise ValueError(entropy)
    value = -(np.mean(logp) + ent)
    g_mu = -np.mean(W, axis=1)
    if q.is_meanfield:
        g_sc = -np.mean(W * u, axis=1)
        inv = 1.0 / sd
    else:
        g_sc = -np.tril(W @ u.T) / M
        inv = np.diag(1.0 / sd)
    if entropy == "ClosedFormEntropy":
        g_sc = g_sc - inv
    return value, _scale_grad_pack(q, g_mu, g_sc), -value
