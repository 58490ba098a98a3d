"""Closed-form gradients of the oracle vs central finite differences
of the restated forward closures with eps held fixed -- i.e. what the reference's AD backend
returns for estimate_repgradelbo_ad_forward / estimate_scoregradelbo_ad_forward."""
import numpy as np
import pytest

from oracle import family as F, models as Mo, objectives as O, philox as P


def fd_grad(f, x, h=1e-6):
    g = np.zeros_like(x)
    for i in range(len(x)):
        e = np.zeros_like(x); e[i] = h
        g[i] = (f(x + e) - f(x - e)) / (2 * h)
    return g


def make_problem(name, D):
    if name == "normal_diag":
        return Mo.NormalDiag(np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D))
    if name == "normal_dense":
        L = np.tril(np.eye(D) + 0.3 * np.ones((D, D)))
        return Mo.NormalDense(np.linspace(-1, 1, D), L)
    d = D - 1
    X, y = Mo.synth_glm_data(40, d, seed=5, family="gaussian" if name == "gaussglm" else "bernoulli_logit")
    if name == "gaussglm":
        return Mo.GaussGLM(X, y, n_data=100)
    return Mo.LogReg(X, y, n_data=100 if name == "logreg_subsampling" else None,
                     variant="subsampling" if name == "logreg_subsampling" else "basic")


def make_q(kind, D):
    mu = 0.1 * np.arange(D) - 0.2
    if kind == "meanfield":
        return F.MeanFieldGaussian(mu, 0.5 + 0.1 * np.arange(D))
    L = np.tril(0.1 * np.ones((D, D))) + np.diag(0.5 + 0.1 * np.arange(D))
    return F.FullRankGaussian(mu, L)


PROBLEMS = ["normal_diag", "normal_dense", "logreg_subsampling", "logreg_basic", "gaussglm"]


@pytest.mark.parametrize("problem", PROBLEMS)
def test_target_gradients_match_finite_differences(problem):
    D = 5
    prob = make_problem(problem, D)
    z = 0.3 * P.normal_matrix(1, 0, D, 1)[:, 0]
    l, g = prob.logdensity_and_gradient(z)
    assert np.allclose(g, fd_grad(prob.logdensity, z), rtol=1e-6, atol=1e-7)
    Z = 0.3 * P.normal_matrix(1, 0, D, 3)
    lb, Gb = prob.logdensity_and_gradient_batch(Z)
    assert np.isclose(lb[0], l) and np.allclose(Gb[:, 0], g)


@pytest.mark.parametrize("family", ["meanfield", "fullrank"])
@pytest.mark.parametrize("entropy", O.ENTROPIES)
@pytest.mark.parametrize("problem", ["normal_dense", "logreg_subsampling"])
def test_repgrad_closed_form_vs_fd(family, entropy, problem):
    D, M = 4, 3
    prob = make_problem(problem, D)
    q = make_q(family, D)
    eps = P.normal_matrix(3, 7, D, M)
    lam = q.destructure()
    q_stop = q.restructure(lam)
    v, g, elbo = O.repgrad_value_and_gradient(lam, q, prob, eps, entropy)
    f = lambda p: O.repgrad_forward(p, q, q_stop, prob, eps, entropy)
    assert np.isclose(v, f(lam), rtol=1e-12)
    g_fd = fd_grad(f, lam)
    if family == "fullrank":      # AD through LowerTriangular: strictly-upper entries get 0
        mask = np.concatenate([np.ones(D, bool), np.tril(np.ones((D, D), bool)).reshape(-1, order="F")])
        g_fd = np.where(mask, g_fd, 0.0)
    assert np.allclose(g, g_fd, rtol=1e-5, atol=1e-6)
    assert elbo == -v


@pytest.mark.parametrize("family", ["meanfield", "fullrank"])
@pytest.mark.parametrize("problem", ["normal_diag", "logreg_basic"])
def test_scoregrad_closed_form_vs_fd(family, problem):
    D, M = 4, 6
    prob = make_problem(problem, D)
    q = make_q(family, D)
    eps = P.normal_matrix(3, 2, D, M)
    lam = q.destructure()
    Z = q.rand_from_eps(eps)
    logpi = prob.logdensity_batch(Z)
    v, g, elbo = O.scoregrad_value_and_gradient(lam, q, prob, eps)
    f = lambda p: O.scoregrad_forward(p, q, Z, logpi)
    assert np.isclose(v, f(lam), rtol=1e-10)
    g_fd = fd_grad(f, lam)
    if family == "fullrank":
        mask = np.concatenate([np.ones(D, bool), np.tril(np.ones((D, D), bool)).reshape(-1, order="F")])
        g_fd = np.where(mask, g_fd, 0.0)
    assert np.allclose(g, g_fd, rtol=1e-4, atol=1e-5)
    assert np.isclose(elbo, np.mean(logpi - q.logpdf(Z)))


def test_per_sample_and_batched_evaluation_agree():
    """The reference calls logdensity once per column (repgradelbo.jl:84-86); the batched
    evaluation used by the GPU path and the best-effort CPU baseline is the same arithmetic."""
    prob = make_problem("logreg_subsampling", 6)
    q = make_q("meanfield", 6)
    eps = P.normal_matrix(1, 1, 6, 5)
    a = O.repgrad_value_and_gradient(q.destructure(), q, prob, eps, "ClosedFormEntropy", per_sample=True)
    b = O.repgrad_value_and_gradient(q.destructure(), q, prob, eps, "ClosedFormEntropy", per_sample=False)
    assert np.isclose(a[0], b[0], rtol=1e-13) and np.allclose(a[1], b[1], rtol=1e-12)


# --- low-rank family: logpdf-based estimators (groundwork) ---------------------------------------
def _lowrank_setup():
    from oracle import family as F, models as Mo, philox as P
    d, r, M = 6, 2, 5
    prob = Mo.NormalDiag(np.linspace(-1, 1, d), np.linspace(0.5, 1.5, d))
    q = F.LowRankGaussian(0.1 * np.arange(d), 0.5 + 0.1 * np.arange(d), 0.3 * P.normal_matrix(61, 0, d, r))
    return q, prob, P.normal_matrix(62, 0, d, M), P.normal_matrix(63, 0, r, M)


def _fd(f, x, h=1e-6):
    return np.array([(f(x + h * e) - f(x - h * e)) / (2 * h) for e in np.eye(len(x))])


def test_lowrank_woodbury_pieces():
    q, _, u1, u2 = _lowrank_setup()
    Sigma = q.cov()
    R = q.rand_from_eps(u1, u2) - q.location[:, None]
    assert np.allclose(q.cov_solve(R), np.linalg.solve(Sigma, R), rtol=1e-10)
    diag, SinvU = q.cov_inv_diag_and_factor()
    assert np.allclose(diag, np.diag(np.linalg.inv(Sigma)), rtol=1e-10)
    assert np.allclose(SinvU, np.linalg.solve(Sigma, q.scale_factors), rtol=1e-10)


@pytest.mark.parametrize("entropy", ["StickingTheLandingEntropy", "MonteCarloEntropy", "StickingTheLandingEntropyZeroGradient"])
def test_lowrank_repgrad_logpdf_entropies_vs_fd(entropy):
    """Closed forms vs central differences of the forward closure the reference differentiates
    (repgradelbo.jl:142-149 with entropy.jl:42-46 / :59-65: q frozen inside log q for STL)."""
    from oracle import objectives as O
    q, prob, u1, u2 = _lowrank_setup()
    lam = q.destructure()
    v, g, e = O.repgrad_lowrank_value_and_gradient(lam, q, prob, u1, u2, entropy)

    def forward(x):
        qx = q.restructure(x)
        Z = qx.rand_from_eps(u1, u2)
        logp, _ = prob.logdensity_and_gradient_batch(Z)
        q_in_logpdf = qx if entropy == "MonteCarloEntropy" else q    # live q vs q_stop
        extra = -qx.entropy() + q.entropy() if entropy == "StickingTheLandingEntropyZeroGradient" else 0.0   # entropy.jl:86-89
        return -(np.mean(logp) - np.mean(q_in_logpdf.logpdf(Z)) + extra)
    assert np.isclose(v, forward(lam))
    assert np.allclose(g, _fd(forward, lam), rtol=2e-6, atol=2e-7)


def test_lowrank_scoregrad_vs_fd():
    from oracle import objectives as O
    q, prob, u1, u2 = _lowrank_setup()
    lam = q.destructure()
    v, g, e = O.scoregrad_lowrank_value_and_gradient(lam, q, prob, u1, u2)
    Z = q.rand_from_eps(u1, u2)
    logp, _ = prob.logdensity_and_gradient_batch(Z)

    def forward(x):                       # scoregradelbo.jl:87-94: samples and log pi are constants
        f = q.restructure(x).logpdf(Z) - logp
        return (np.mean(f * f) - np.mean(f) ** 2) / 2
    assert np.isclose(v, forward(lam))
    assert np.allclose(g, _fd(forward, lam), rtol=2e-6, atol=2e-7)
    assert np.isclose(e, np.mean(logp - q.logpdf(Z)))


# ---- non-Gaussian base distributions of MvLocationScale (docs/src/families.md:72-101) ------------------------------
def _base(name):
    return F.LaplaceDist() if name == "laplace" else F.TDistBase(5.0)


def _base_draws(name, key, step, D, M):
    return P.laplace_matrix(key, step, D, M) if name == "laplace" else P.student_t_matrix(key, step, D, M, 5.0)


@pytest.mark.parametrize("base", ["laplace", "tdist"])
@pytest.mark.parametrize("family", ["meanfield", "fullrank"])
@pytest.mark.parametrize("entropy", O.ENTROPIES)
def test_repgrad_closed_form_vs_fd_nongaussian_base(base, family, entropy):
    """The closed forms with eps -> u (draws of the base) and eps -> -score(u) in the sticking-the-landing term equal
    what AD returns for the restated forward, for every entropy estimator."""
    D, M = 4, 3
    prob = make_problem("logreg_subsampling", D)
    q0 = make_q(family, D)
    q = F.MvLocationScale(q0.location, q0.scale, _base(base))
    u = _base_draws(base, 3, 7, D, M)
    lam = q.destructure()
    q_stop = q.restructure(lam)
    v, g, elbo = O.repgrad_value_and_gradient(lam, q, prob, u, entropy)
    f = lambda p: O.repgrad_forward(p, q, q_stop, prob, u, entropy)
    assert np.isclose(v, f(lam), rtol=1e-12)
    g_fd = fd_grad(f, lam)
    if family == "fullrank":
        mask = np.concatenate([np.ones(D, bool), np.tril(np.ones((D, D), bool)).reshape(-1, order="F")])
        g_fd = np.where(mask, g_fd, 0.0)
    assert np.allclose(g, g_fd, rtol=2e-5, atol=2e-6)
    if entropy in ("ClosedFormEntropy", "StickingTheLandingEntropy"):   # the round-1 groundwork function agrees
        v2, g2, _ = O.repgrad_general_base_value_and_gradient(lam, q, prob, u, entropy)
        assert np.isclose(v, v2, rtol=1e-13) and np.allclose(g, g2, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("base", ["laplace", "tdist"])
@pytest.mark.parametrize("family", ["meanfield", "fullrank"])
def test_scoregrad_closed_form_vs_fd_nongaussian_base(base, family):
    D, M = 4, 6
    prob = make_problem("logreg_basic", D)
    q0 = make_q(family, D)
    q = F.MvLocationScale(q0.location, q0.scale, _base(base))
    u = _base_draws(base, 3, 2, D, M)
    lam = q.destructure()
    Z = q.rand_from_eps(u)
    logpi = prob.logdensity_batch(Z)
    v, g, elbo = O.scoregrad_value_and_gradient(lam, q, prob, u)
    f = lambda p: O.scoregrad_forward(p, q, Z, logpi)
    assert np.isclose(v, f(lam), rtol=1e-10)
    g_fd = fd_grad(f, lam)
    if family == "fullrank":
        mask = np.concatenate([np.ones(D, bool), np.tril(np.ones((D, D), bool)).reshape(-1, order="F")])
        g_fd = np.where(mask, g_fd, 0.0)
    # (Laplace: the score is discontinuous at 0 -- no draw sits there, and FD with h = 1e-6 stays on one side)
    assert np.allclose(g, g_fd, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("base,nu", [("laplace", None), ("tdist", 3.0), ("tdist", 7.5)])
def test_base_draws_follow_their_distribution(base, nu):
    """The counter-based samplers the device mirrors (inversion for Laplace, Bailey's polar transform for Student-t):
    Kolmogorov-Smirnov against scipy's CDF, moments, and independence of the coordinates that share a Philox block."""
    from scipy import stats
    key = 0x38BEF07CF9CC549D
    x = P.laplace_matrix(key, 0, 64, 4000) if base == "laplace" else P.student_t_matrix(key, 0, 64, 4000, nu)
    dist = stats.laplace() if base == "laplace" else stats.t(nu)
    assert stats.kstest(x.ravel()[:40000], dist.cdf).pvalue > 1e-3
    assert abs(np.median(x)) < 0.02
    for a, b in ((0, 1), (0, 2), (1, 3)):
        assert abs(np.corrcoef(np.abs(x[a]), np.abs(x[b]))[0, 1]) < 0.06
    # entropy / logpdf of the oracle's base classes against scipy
    d = F.LaplaceDist() if base == "laplace" else F.TDistBase(nu)
    assert np.isclose(d.entropy(), dist.entropy(), rtol=1e-12)
    assert np.allclose(d.logpdf(x[:3, :50]), dist.logpdf(x[:3, :50]), rtol=1e-12, atol=1e-12)
