"""Independent pinning of the oracle (CPU): the reference differentiates its forward closures with an AD backend
(src/AdvancedVI.jl:47-98, estimate_repgradelbo_ad_forward / estimate_scoregradelbo_ad_forward) and relies on
Distributions.jl / Optimisers.jl for the densities and the update rules.  Neither can run here, but PyTorch's
autograd (an AD backend), torch.distributions / scipy.stats (the same closed-form densities from an unrelated code
base) and torch.optim (the same published update rules) can.  These tests re-state the reference's forward model from its
SOURCE TEXT in torch -- not from the oracle -- and require the oracle's closed forms to agree with what autograd returns,
to 1e-10 in float64.  They complement tests/test_oracle_gradients.py (central finite differences).

What this pins that the round-1 verdict listed as unpinned: the docs-only logistic-regression arithmetic
(docs/src/tutorials/subsampling.md:26-38, README.md:47-58), the Adam / Descent arithmetic of Optimisers.jl
(SURVEY.md Appendix B), and the closed-form gradients of Appendix A as AD results rather than FD approximations."""
import numpy as np
import pytest
import torch

from oracle import family as F, models as Mo, objectives as O, optim as Op, philox as P

torch.set_default_dtype(torch.float64)
LOG2PI = float(np.log(2 * np.pi))


# ---- the reference's models, restated from their source text with torch.distributions ------------------------------
def logreg_logdensity_torch(theta, X, y, n_data, variant):
    """docs/src/tutorials/subsampling.md:26-38 (variant "subsampling": Normal(0, 3) prior on sigma = exp(eta), likelihood
    scaled by n_data / n) and README.md:47-58 under the exp bijector of README.md:91-106 (variant "basic": LogNormal(0, 3)
    prior on sigma plus the log-Jacobian eta)."""
    D = torch.distributions
    n, d = X.shape
    beta, eta = theta[:d], theta[d]
    sigma = torch.exp(eta)
    logprior_beta = D.Normal(torch.zeros(d), sigma).log_prob(beta).sum()          # MvNormal(Zeros(d), sigma): isotropic, std sigma
    logit = X @ beta
    loglike = -torch.nn.functional.binary_cross_entropy_with_logits(logit, y, reduction="sum")   # sum logpdf(BernoulliLogit(l), y)
    if variant == "subsampling":
        return n_data / n * loglike + logprior_beta + D.Normal(0.0, 3.0).log_prob(sigma)
    return loglike + logprior_beta + D.LogNormal(0.0, 3.0).log_prob(sigma) + eta


def gaussglm_logdensity_torch(theta, X, y, n_data):
    D = torch.distributions
    n, d = X.shape
    beta, eta = theta[:d], theta[d]
    sigma = torch.exp(eta)
    return (n_data / n * D.Normal(X @ beta, 1.0).log_prob(y).sum() + D.Normal(torch.zeros(d), sigma).log_prob(beta).sum()
            + D.Normal(0.0, 3.0).log_prob(sigma))


def _data(n=30, d=5, fam="bernoulli_logit"):
    X, y = Mo.synth_glm_data(n, d, seed=3, family=fam)
    return X.astype(np.float64), y.astype(np.float64)


@pytest.mark.parametrize("variant", ["subsampling", "basic"])
def test_logreg_value_and_gradient_against_autograd(variant):
    X, y = _data()
    n_data = 100 if variant == "subsampling" else X.shape[0]
    prob = Mo.LogReg(X, y, n_data=n_data, variant=variant)
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y)
    for k in range(4):
        z = 0.4 * P.normal_matrix(11, k, X.shape[1] + 1, 1)[:, 0]
        th = torch.tensor(z, requires_grad=True)
        lp = logreg_logdensity_torch(th, Xt, yt, n_data, variant)
        lp.backward()
        l, g = prob.logdensity_and_gradient(z)
        assert abs(l - lp.item()) <= 1e-11 * abs(lp.item())
        assert np.allclose(g, th.grad.numpy(), rtol=1e-10, atol=1e-12)


def test_gaussglm_value_and_gradient_against_autograd():
    X, y = _data(fam="gaussian")
    prob = Mo.GaussGLM(X, y, n_data=77)
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y)
    z = 0.3 * P.normal_matrix(12, 0, X.shape[1] + 1, 1)[:, 0]
    th = torch.tensor(z, requires_grad=True)
    lp = gaussglm_logdensity_torch(th, Xt, yt, 77)
    lp.backward()
    l, g = prob.logdensity_and_gradient(z)
    assert abs(l - lp.item()) <= 1e-11 * abs(lp.item()) and np.allclose(g, th.grad.numpy(), rtol=1e-10, atol=1e-12)


# ---- the family and the objectives, restated from src/families/location_scale.jl and src/algorithms/*.jl ------------
def q_pieces(lam, D, meanfield):
    mu = lam[:D]
    if meanfield:
        return mu, torch.diag(lam[D:])
    return mu, torch.tril(lam[D:].reshape(D, D).T)          # vec(L) is column-major; AD through LowerTriangular


def logpdf_q(z, mu, L):
    """location_scale.jl:59-63: sum(logpdf.(Normal(0, 1), scale \\ (z - location))) - logdet(scale), per column."""
    u = torch.linalg.solve_triangular(L, z - mu[:, None], upper=False)
    return torch.distributions.Normal(0.0, 1.0).log_prob(u).sum(0) - torch.log(torch.diagonal(L)).sum()


def entropy_q(L):
    """location_scale.jl:52-57: D * entropy(Normal(0, 1)) + logdet(scale)."""
    D = L.shape[0]
    return D * torch.distributions.Normal(0.0, 1.0).entropy() + torch.log(torch.diagonal(L)).sum()


def repgrad_forward_torch(lam, lam_stop, eps, target, D, meanfield, entropy):
    """estimate_repgradelbo_ad_forward (repgradelbo.jl:142-149) with the estimators of entropy.jl:11-90."""
    mu, L = q_pieces(lam, D, meanfield)
    mu_s, L_s = q_pieces(lam_stop, D, meanfield)
    z = L @ eps + mu[:, None]                                  # rand: scale * eps .+ location (location_scale.jl:71-87)
    energy = torch.stack([target(z[:, m]) for m in range(z.shape[1])]).mean()      # repgradelbo.jl:84-86
    if entropy == "ClosedFormEntropy":
        ent = entropy_q(L)
    elif entropy == "ClosedFormEntropyZeroGradient":
        ent = entropy_q(L_s)
    elif entropy == "MonteCarloEntropy":
        ent = -logpdf_q(z, mu, L).mean()
    elif entropy == "StickingTheLandingEntropy":
        ent = -logpdf_q(z, mu_s, L_s).mean()
    else:   # StickingTheLandingEntropyZeroGradient, entropy.jl:80-90
        ent = -logpdf_q(z, mu_s, L_s).mean() - entropy_q(L) + entropy_q(L_s)
    return -(energy + ent)


@pytest.mark.parametrize("family", ["meanfield", "fullrank"])
@pytest.mark.parametrize("entropy", O.ENTROPIES)
def test_repgrad_closed_forms_equal_autograd(family, entropy):
    """SURVEY.md Appendix A.1-A.3 == what an AD backend returns for the reference's forward closure."""
    X, y = _data(n=20, d=3)
    D, M = 4, 5
    prob = Mo.LogReg(X, y, n_data=50)
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y)
    target = lambda th: logreg_logdensity_torch(th, Xt, yt, 50, "subsampling")
    mu = 0.1 * np.arange(D) - 0.2
    q = (F.MeanFieldGaussian(mu, 0.5 + 0.1 * np.arange(D)) if family == "meanfield"
         else F.FullRankGaussian(mu, np.tril(0.1 * np.ones((D, D))) + np.diag(0.5 + 0.1 * np.arange(D))))
    lam = q.destructure()
    eps = P.normal_matrix(3, 7, D, M)
    v, g, elbo = O.repgrad_value_and_gradient(lam, q, prob, eps, entropy)
    lt = torch.tensor(lam, requires_grad=True)
    val = repgrad_forward_torch(lt, torch.tensor(lam), torch.from_numpy(eps), target, D, family == "meanfield", entropy)
    val.backward()
    assert abs(v - val.item()) <= 1e-11 * abs(val.item()) and elbo == -v
    assert np.allclose(g, lt.grad.numpy(), rtol=1e-9, atol=1e-11)
    if family == "fullrank":     # the strictly-upper entries of vec(L) get exactly 0 from AD through LowerTriangular
        upper = ~np.tril(np.ones((D, D), bool)).reshape(-1, order="F")
        assert np.all(g[D:][upper] == 0.0) and np.all(lt.grad.numpy()[D:][upper] == 0.0)


@pytest.mark.parametrize("family", ["meanfield", "fullrank"])
def test_scoregrad_closed_form_equals_autograd(family):
    """SURVEY.md Appendix A.4 (VarGrad, scoregradelbo.jl:87-94): samples and log pi are constants, AD goes through
    logpdf(q_lambda, z) only."""
    X, y = _data(n=20, d=3)
    D, M = 4, 7
    prob = Mo.LogReg(X, y, variant="basic")
    mu = 0.1 * np.arange(D) - 0.2
    q = (F.MeanFieldGaussian(mu, 0.5 + 0.1 * np.arange(D)) if family == "meanfield"
         else F.FullRankGaussian(mu, np.tril(0.1 * np.ones((D, D))) + np.diag(0.5 + 0.1 * np.arange(D))))
    lam = q.destructure()
    eps = P.normal_matrix(3, 2, D, M)
    v, g, elbo = O.scoregrad_value_and_gradient(lam, q, prob, eps)
    Z = torch.from_numpy(q.rand_from_eps(eps))
    logpi = torch.from_numpy(prob.logdensity_batch(Z.numpy()))
    lt = torch.tensor(lam, requires_grad=True)
    mu_t, L_t = q_pieces(lt, D, family == "meanfield")
    f = logpdf_q(Z, mu_t, L_t) - logpi
    val = (torch.mean(f * f) - torch.mean(f) ** 2) / 2
    val.backward()
    assert abs(v - val.item()) <= 1e-10 * abs(val.item())
    assert np.allclose(g, lt.grad.numpy(), rtol=1e-8, atol=1e-10)
    assert abs(elbo - float(torch.mean(logpi - f.detach() - logpi))) <= 1e-10 * abs(elbo)


# ---- Optimisers.jl rules against torch.optim (the same published algorithms, an unrelated implementation) -----------
def test_adam_and_descent_match_torch_optim():
    """Optimisers.Adam(eta, (b1, b2), eps): m, v moments, bias-corrected step eta * mhat / (sqrt(vhat) + eps)
     == torch.optim.Adam; Optimisers.Descent(eta) == torch.optim.SGD."""
    rng = np.random.default_rng(0)
    A = rng.standard_normal((6, 6)); A = A @ A.T + np.eye(6)
    b = rng.standard_normal(6)
    grad = lambda x: A @ x - b
    for mk_o, mk_t in ((lambda: Op.Adam(3e-2), lambda p: torch.optim.Adam([p], lr=3e-2, betas=(0.9, 0.999), eps=1e-8)),
                       (lambda: Op.Descent(1e-2), lambda p: torch.optim.SGD([p], lr=1e-2))):
        x = rng.standard_normal(6)
        p = torch.tensor(x.copy(), requires_grad=True)
        opt_t = mk_t(p)
        rule = mk_o()
        st = rule.init(x)
        for _ in range(25):
            st, dx = rule.apply(st, x, grad(x))
            x = x - dx                                      # Optimisers.update!: x .- dx
            opt_t.zero_grad()
            p.grad = torch.from_numpy(grad(p.detach().numpy()))
            opt_t.step()
            assert np.allclose(x, p.detach().numpy(), rtol=1e-12, atol=1e-14)


# ---- non-Gaussian base distributions (docs/src/families.md:72-101) through torch.distributions -----------------------
def _torch_base(name, nu):
    D = torch.distributions
    return D.Laplace(0.0, 1.0) if name == "laplace" else D.StudentT(nu)


@pytest.mark.parametrize("base,nu", [("laplace", None), ("tdist", 5.0)])
@pytest.mark.parametrize("family", ["meanfield", "fullrank"])
@pytest.mark.parametrize("entropy", ["ClosedFormEntropy", "MonteCarloEntropy", "StickingTheLandingEntropy"])
def test_nongaussian_base_closed_forms_equal_autograd(base, nu, family, entropy):
    """MvLocationScale(location, scale, dist): logpdf = sum logpdf(dist, scale \\ (z - location)) - logdet(scale),
    entropy = D entropy(dist) + logdet(scale) (location_scale.jl:52-63) with torch's Laplace / StudentT."""
    X, y = _data(n=20, d=3)
    D, M = 4, 5
    prob = Mo.LogReg(X, y, n_data=50)
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y)
    target = lambda th: logreg_logdensity_torch(th, Xt, yt, 50, "subsampling")
    dist_t = _torch_base(base, nu)
    dist_o = F.LaplaceDist() if base == "laplace" else F.TDistBase(nu)
    mu = 0.1 * np.arange(D) - 0.2
    scale = (0.5 + 0.1 * np.arange(D)) if family == "meanfield" else np.tril(0.1 * np.ones((D, D))) + np.diag(0.5 + 0.1 * np.arange(D))
    q = F.MvLocationScale(mu, scale, dist_o)
    lam = q.destructure()
    u = P.laplace_matrix(3, 7, D, M) if base == "laplace" else P.student_t_matrix(3, 7, D, M, nu)
    v, g, _ = O.repgrad_value_and_gradient(lam, q, prob, u, entropy)

    def logpdf_b(z, m_, L_):
        w = torch.linalg.solve_triangular(L_, z - m_[:, None], upper=False)
        return dist_t.log_prob(w).sum(0) - torch.log(torch.diagonal(L_)).sum()
    lt = torch.tensor(lam, requires_grad=True)
    m_, L_ = q_pieces(lt, D, family == "meanfield")
    ms, Ls = q_pieces(torch.tensor(lam), D, family == "meanfield")
    z = L_ @ torch.from_numpy(u) + m_[:, None]
    energy = torch.stack([target(z[:, k]) for k in range(M)]).mean()
    if entropy == "ClosedFormEntropy":
        ent = D * dist_t.entropy() + torch.log(torch.diagonal(L_)).sum()
    elif entropy == "MonteCarloEntropy":
        ent = -logpdf_b(z, m_, L_).mean()
    else:
        ent = -logpdf_b(z, ms, Ls).mean()
    val = -(energy + ent)
    val.backward()
    assert abs(v - val.item()) <= 1e-10 * abs(val.item())
    assert np.allclose(g, lt.grad.numpy(), rtol=1e-8, atol=1e-10)


# ---- the low-rank family against torch.distributions.LowRankMultivariateNormal --------------------------------------
def test_lowrank_family_and_gradients_against_torch():
    """MvLocationScaleLowRank (src/families/location_scale_low_rank.jl): covariance diag(D^2) + U U'.  logpdf / entropy
    (Woodbury + matrix determinant lemma in the oracle, :35-67) against LowRankMultivariateNormal, and the RepGradELBO
    closed forms (closed-form / Monte-Carlo / sticking-the-landing entropy) against autograd through rand (:79-86)."""
    d, r, M = 6, 2, 5
    rng = np.random.default_rng(2)
    loc = 0.3 * rng.standard_normal(d)
    sd = 0.5 + 0.3 * rng.random(d)
    U = 0.2 * rng.standard_normal((d, r))
    q = F.MvLocationScaleLowRank(loc, sd, U)
    Z = rng.standard_normal((d, 9))
    ref = torch.distributions.LowRankMultivariateNormal(torch.from_numpy(loc), torch.from_numpy(U), torch.from_numpy(sd ** 2))
    assert np.allclose(q.logpdf(Z), ref.log_prob(torch.from_numpy(Z.T)).numpy(), rtol=1e-11)
    assert np.isclose(q.entropy(), ref.entropy().item(), rtol=1e-12)

    X, y = _data(n=20, d=d - 1)
    prob = Mo.LogReg(X, y, n_data=50)
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y)
    target = lambda th: logreg_logdensity_torch(th, Xt, yt, 50, "subsampling")
    lam = q.destructure()
    u1, u2 = P.normal_matrix(5, 0, d, M), P.normal_matrix(5, 0, r, M, stream=P.STREAM_EPS_FACTORS)
    for entropy in ("ClosedFormEntropy", "MonteCarloEntropy", "StickingTheLandingEntropy"):
        v, g, _ = O.repgrad_lowrank_value_and_gradient(lam, q, prob, u1, u2, entropy)
        lt = torch.tensor(lam, requires_grad=True)
        m_, d_, U_ = lt[:d], lt[d:2 * d], lt[2 * d:].reshape(r, d).T          # destructure order, column-major factors
        z = d_[:, None] * torch.from_numpy(u1) + U_ @ torch.from_numpy(u2) + m_[:, None]
        energy = torch.stack([target(z[:, k]) for k in range(M)]).mean()
        live = torch.distributions.LowRankMultivariateNormal(m_, U_, d_ ** 2)
        if entropy == "ClosedFormEntropy":
            ent = live.entropy()
        elif entropy == "MonteCarloEntropy":
            ent = -live.log_prob(z.T).mean()
        else:
            ent = -ref.log_prob(z.T).mean()
        val = -(energy + ent)
        val.backward()
        assert abs(v - val.item()) <= 1e-10 * abs(val.item()), entropy
        assert np.allclose(g, lt.grad.numpy(), rtol=1e-7, atol=1e-9), entropy


# ---- the committed golden vectors are reproduced by the independent AD restatement (they are not circular) ----------
def test_golden_vectors_are_reproduced_by_autograd():
    """tests/golden/elbo_golden.npz was written by the oracle; here the value slot, the gradient and the ELBO of every
    case are recomputed from the stored inputs (X, y, lambda, eps) by autograd over the torch restatement of the
    reference's forward closures -- nothing from oracle/ -- and must equal the stored vectors to 1e-9."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as MG
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "elbo_golden.npz"))
    for name, (n, d, M, key, dseed, fam, objective, entropy) in MG.CASES.items():
        X, y = torch.from_numpy(gold[f"{name}/X"].astype(np.float64)), torch.from_numpy(gold[f"{name}/y"].astype(np.float64))
        lam, eps = gold[f"{name}/lam"], torch.from_numpy(gold[f"{name}/eps"])
        D = d + 1
        target = lambda th: logreg_logdensity_torch(th, X, y, n, "subsampling")     # Mo.LogReg(X, y): n_data = n
        lt = torch.tensor(lam, requires_grad=True)
        if objective == "rep":
            val = repgrad_forward_torch(lt, torch.tensor(lam), eps, target, D, fam == "mf", entropy)
            elbo = -val.item()
        else:
            mu_s, L_s = q_pieces(torch.tensor(lam), D, fam == "mf")
            Z = L_s @ eps + mu_s[:, None]
            logpi = torch.stack([target(Z[:, m]) for m in range(M)])
            mu_t, L_t = q_pieces(lt, D, fam == "mf")
            f = logpdf_q(Z, mu_t, L_t) - logpi
            val = (torch.mean(f * f) - torch.mean(f) ** 2) / 2
            elbo = float(torch.mean(logpi - logpdf_q(Z, mu_s, L_s)))
        val.backward()
        assert abs(val.item() - float(gold[f"{name}/value"])) <= 1e-9 * max(1.0, abs(val.item())), name
        assert abs(elbo - float(gold[f"{name}/elbo"])) <= 1e-9 * max(1.0, abs(elbo)), name
        assert np.allclose(lt.grad.numpy(), gold[f"{name}/grad"], rtol=1e-8, atol=1e-10), name
