"""Pin the oracle against the reference's own known-answer tests.

Each test names the reference test it re-expresses (paths under /root/reference/test).
No GPU needed.  Random draws use the oracle's Philox stream; the reference's StableRNG
stream values are not part of what these tests pin.
"""
import numpy as np
import pytest
from scipy import stats

from oracle import family as F, models as Mo, objectives as O, optim as Op, philox as P
from oracle import reshuffling as R


def normal_meanfield():
    """test/models/normal.jl:56-75 (5-D, mu = 5, sigma = 0.3)."""
    n = 5
    return Mo.NormalDiag(np.full(n, 5.0), np.full(n, 0.3)), np.full(n, 5.0), np.full(n, 0.3)


def normal_fullrank():
    """test/models/normal.jl:36-54."""
    n = 5
    L = 0.3 * np.eye(n)
    return Mo.NormalDense(np.full(n, 5.0), L), np.full(n, 5.0), L


# --- test/families/location_scale.jl ------------------------------------------------------
@pytest.mark.parametrize("covtype", ["meanfield", "fullrank"])
def test_family_logpdf_entropy_moments(covtype):
    """location_scale.jl (test) :34-97: logpdf / entropy vs MvNormal, mean/var/cov, sample
    moments; dense case L = tril(I + ones/2) (:13)."""
    d = 5
    mu = np.full(d, 1.0)
    Ld = np.tril(np.eye(d) + np.ones((d, d)) / 2)
    q = F.MeanFieldGaussian(mu, np.diag(Ld).copy()) if covtype == "meanfield" else F.FullRankGaussian(mu, Ld)
    Sigma = np.diag(np.diag(Ld) ** 2) if covtype == "meanfield" else Ld @ Ld.T
    ref = stats.multivariate_normal(mu, Sigma)
    z = P.normal_matrix(3, 0, d, 7)
    assert np.allclose(q.logpdf(z), ref.logpdf(z.T), rtol=1e-10)
    assert np.isclose(q.entropy(), ref.entropy(), rtol=1e-12)
    assert np.allclose(q.mean(), mu) and np.allclose(q.var(), np.diag(Sigma)) and np.allclose(q.cov(), Sigma)
    Z = q.rand_from_eps(P.normal_matrix(5, 0, d, 10 ** 6))
    assert np.allclose(Z.mean(axis=1), mu, rtol=1e-2)
    assert np.allclose(Z.var(axis=1), np.diag(Sigma), rtol=1e-2)
    assert np.allclose(np.cov(Z), Sigma, rtol=1e-2, atol=1e-2)


def test_meanfield_destructure_length_and_roundtrip():
    """location_scale.jl (test) :146-155: Diagonal destructure has length 2d and round-trips."""
    d = 5
    q = F.MeanFieldGaussian(np.arange(d, dtype=float), np.arange(1, d + 1, dtype=float))
    lam = q.destructure()
    assert lam.shape == (2 * d,)
    q2 = q.restructure(lam)
    assert np.array_equal(q2.location, q.location) and np.array_equal(q2.scale, q.scale)
    qf = F.FullRankGaussian(np.zeros(d), np.tril(np.ones((d, d))))
    lamf = qf.destructure()
    assert lamf.shape == (d + d * d,)
    assert np.array_equal(qf.restructure(lamf).scale, qf.scale)


# --- test/algorithms/klmin*descent.jl: estimate_objective ~ 0 at q = pi ---------------------
@pytest.mark.parametrize("make", [normal_meanfield, normal_fullrank])
def test_estimate_objective_zero_at_truth(make):
    """klminrepgraddescent.jl:23-38 / klminscoregraddescent.jl:23-38: with q = pi and
    MonteCarloEntropy, log pi(z) - log q(z) = 0 pointwise, so the estimate is ~ 0."""
    prob, mu, L = make()
    q = F.MvLocationScale(mu, L)
    eps = P.normal_matrix(0x38BEF07CF9CC549D, 0, 5, 10 ** 5)
    assert abs(O.estimate_objective(q, prob, eps)) < 1e-8          # reference: atol 1e-2
    assert abs(O.scoregrad_estimate_objective(q, prob, eps)) < 1e-8
    # ClosedFormEntropy variant is only ~0 in expectation (atol of the reference: 1e-2)
    assert abs(O.repgrad_estimate_objective(q, prob, eps, "ClosedFormEntropy")) < 1e-2


@pytest.mark.parametrize("make", [normal_meanfield, normal_fullrank])
@pytest.mark.parametrize("M", [1, 10])
def test_stl_gradient_vanishes_at_truth(make, M):
    """klminrepgraddescent.jl:66-87: STL gradient at q = pi has norm ~ 0 (atol 1e-5)."""
    prob, mu, L = make()
    q = F.MvLocationScale(mu, L)
    eps = P.normal_matrix(1, 0, 5, M)
    _, g, _ = O.repgrad_value_and_gradient(q.destructure(), q, prob, eps, "StickingTheLandingEntropy")
    assert np.linalg.norm(g) < 1e-10


def test_closed_form_expectations_c1():
    """target N(0, I_2), q = N(mu, diag s^2):
    E[-ELBO] = -1 + (|mu|^2 + |s|^2)/2 - sum log s; E grad_mu = mu; E grad_s = s - 1/s."""
    prob = Mo.NormalDiag(np.zeros(2), np.ones(2))
    mu, s = np.array([0.3, -0.2]), np.array([0.7, 1.4])
    q = F.MeanFieldGaussian(mu, s)
    eps = P.normal_matrix(4, 0, 2, 400000)
    v, g, elbo = O.repgrad_value_and_gradient(q.destructure(), q, prob, eps, "ClosedFormEntropy")
    assert np.isclose(v, -1 + 0.5 * (mu @ mu + s @ s) - np.log(s).sum(), atol=1e-2)
    assert np.allclose(g[:2], mu, atol=5e-3) and np.allclose(g[2:], s - 1 / s, atol=1e-2)
    assert elbo == -v


# --- convergence (klminrepgraddescent.jl:105-121, klminscoregraddescent.jl:82-97) -----------
@pytest.mark.parametrize("objective,M", [("ClosedFormEntropy", 1), ("StickingTheLandingEntropy", 1),
                                         ("score", 100)])
def test_convergence_descent_clipscale(objective, M):
    prob, mu_true, L_true = normal_meanfield()
    q0 = F.MeanFieldGaussian(np.zeros(5), np.ones(5))
    rule, op, avg = Op.Descent(1e-3), Op.ClipScale(), Op.PolynomialAveraging()
    st = Op.sgd_init(q0, rule, avg)

    def grad_fn(params, t):
        eps = P.normal_matrix(9, t, 5, M)
        if objective == "score":
            v, g, e = O.scoregrad_value_and_gradient(params, q0, prob, eps)
        else:
            v, g, e = O.repgrad_value_and_gradient(params, q0, prob, eps, objective)
        return v, g, dict(elbo=e)

    for _ in range(1000):
        Op.sgd_step(st, q0, grad_fn, rule, op, avg)
    q = Op.sgd_output(st, q0, avg)
    d0 = np.sum((q0.location - mu_true) ** 2) + np.sum((q0.scale - L_true) ** 2)
    d1 = np.sum((q.location - mu_true) ** 2) + np.sum((q.scale - L_true) ** 2)
    assert d1 <= d0 / 2


# --- test/general/subsampledobj.jl:62-89 ---------------------------------------------------
@pytest.mark.parametrize("batchsize", [1, 3, 4])
def test_epoch_mean_of_minibatch_gradients_equals_full_gradient(batchsize):
    n_data = 8
    mus = P.normal_matrix(4, 0, n_data, 1)[:, 0]
    prob = Mo.SubsampledNormals(mus)
    q0 = F.MeanFieldGaussian(np.array([mus.mean()]), np.array([np.sqrt(1 / n_data)]))
    params = q0.destructure()
    eps = P.normal_matrix(0x38BEF07CF9CC549D, 1, 1, 10)          # same MC samples for all batches
    _, g_ref, _ = O.repgrad_value_and_gradient(params, q0, prob, eps, "ClosedFormEntropy")
    sub = R.ReshufflingBatchSubsampling(np.arange(n_data), batchsize)
    state = R.subsampled_init(sub, key=5)
    grads = []
    for _ in range(len(sub)):
        def grad_fn(prob_sub):
            v, g, e = O.repgrad_value_and_gradient(params, q0, prob_sub, eps, "ClosedFormEntropy")
            return v, g, dict(elbo=e)
        _, g, state, info = R.subsampled_estimate_gradient(sub, state, prob, grad_fn)
        grads.append(g)
        assert set(info) == {"epoch", "step", "elbo"}
    if n_data % batchsize == 0:
        assert np.allclose(np.mean(grads, axis=0), g_ref, rtol=1e-10)
    else:
        # batchsize 3: the short trailing batch is swapped for the first batch of the next
        # epoch (reshuffling.jl:46-52), so only the scaling is checked here
        assert np.all(np.isfinite(np.mean(grads, axis=0)))


def test_reshuffling_state_machine():
    """reshuffling.jl:38-60: every index once per epoch; drop_trailing swaps the short batch
    for the first batch of the next epoch and reports the new epoch number."""
    sub = R.ReshufflingBatchSubsampling(np.arange(8), 3)
    assert len(sub) == 3
    st = R.sub_init(sub, key=1)
    seen, infos = [], []
    for _ in range(3):
        b, st, info = R.sub_step(sub, st)
        seen.extend(b.tolist()); infos.append(info)
    assert sorted(seen) == list(range(8))
    assert [i["step"] for i in infos] == [1, 2, 3] and [i["epoch"] for i in infos] == [1, 1, 2]
    st = R.sub_init(sub, key=1)
    sizes = []
    for _ in range(7):
        b, st, info = R.sub_step(sub, st, True)
        sizes.append(len(b))
    assert all(s == 3 for s in sizes)
    # determinism + shard independence of the permutation stream
    assert np.array_equal(R.philox_shuffle(np.arange(100), 3, 2), R.philox_shuffle(np.arange(100), 3, 2))
    assert sorted(R.philox_shuffle(np.arange(100), 3, 2).tolist()) == list(range(100))


# --- test/general/averaging.jl:25-37 --------------------------------------------------------
def test_polynomial_averaging_weights():
    eta, d, n = 1, 3, 10
    xs = P.normal_matrix(6, 0, d, n)
    avg = Op.PolynomialAveraging(eta)
    st = avg.init(xs[:, 0])
    for t in range(n):
        st = avg.apply(st, xs[:, t])
    alpha = [(eta + 1) / (t + eta) * (1 if t == n else np.prod([(j - 1) / (j + eta) for j in range(t + 1, n + 1)]))
             for t in range(1, n + 1)]
    assert np.allclose(avg.value(st), xs @ np.array(alpha))
    na = Op.NoAveraging()
    s = na.init(xs[:, 0])
    for t in range(n):
        s = na.apply(s, xs[:, t])
    assert np.array_equal(na.value(s), xs[:, -1])


# --- test/general/clip_scale.jl:3-25, proximal_location_scale_entropy.jl:3-56 ---------------
@pytest.mark.parametrize("covtype", ["meanfield", "fullrank"])
def test_clipscale_and_prox(covtype):
    d = 5
    q = F.MeanFieldGaussian(np.zeros(d), np.ones(d)) if covtype == "meanfield" else F.FullRankGaussian(np.zeros(d), np.eye(d))
    e = np.sqrt(0.5)
    # a scale below the clip level must be raised to it
    lam = q.destructure().copy()
    if covtype == "meanfield":
        lam[d:] = 0.1
    else:
        lam[d:] = (0.1 * np.eye(d)).reshape(-1)
    q2 = q.restructure(Op.ClipScale(e).apply(q, Op.Descent(1e-2), None, lam))
    assert np.all(q2.var() >= e ** 2 - 1e-15)
    # prox stationarity: 1/L'_ii = (L'_ii - L_ii)/gamma on the diagonal, off-diagonals untouched
    gamma = 1e-2
    lam = q.destructure()
    q3 = q.restructure(Op.ProximalLocationScaleEntropy().apply(q, Op.Descent(gamma), None, lam))
    s0, s1 = q.scale_diag(), q3.scale_diag()
    assert np.allclose(1.0 / s1, (s1 - s0) / gamma)
    if covtype == "fullrank":
        assert np.array_equal(q3.scale - np.diag(s1), q.scale - np.diag(s0))


# --- test/general/rules.jl:3-28 --------------------------------------------------------------
@pytest.mark.parametrize("rule", [Op.DoWG(), Op.DoG(), Op.DoWG(1e-5), Op.DoG(1e-5)])
def test_rules_reduce_least_squares_loss(rule):
    T, d, n = 10 ** 4, 10, 1000
    w = P.normal_matrix(8, 0, d, 1)[:, 0]
    w_true = P.normal_matrix(8, 1, d, 1)[:, 0]
    X = P.uniform23(P.uniform_u32(8, n * d, P.STREAM_DATA)).reshape(n, d)
    idx = P.uniform_u32(8, T, P.STREAM_DATA, offset=1 << 20) % n

    def loss(Xs, w):
        return np.mean((Xs @ w - Xs @ w_true) ** 2)

    l0 = loss(X, w)
    st = rule.init(w)
    for t in range(T):
        xi = X[idx[t]]
        g = 2 * xi * (xi @ w - xi @ w_true)
        st, dxp = rule.apply(st, w, g)
        w = w - dxp
    assert loss(X, w) < l0 / 10


def test_adam_first_step_is_eta_sign():
    """Optimisers.Adam: bias-corrected first step = eta * g / (|g| + eps)."""
    x = np.array([1.0, -2.0, 3.0]); g = np.array([0.5, -4.0, 1e-3])
    rule = Op.Adam(1e-2)
    st, dxp = rule.apply(rule.init(x), x, g)
    assert np.allclose(dxp, 1e-2 * g / (np.abs(g) + 1e-8), rtol=1e-6)
    assert np.allclose(st[2], (0.81, 0.998001))


def test_nonfinite_objective_raises():
    """src/algorithms/common.jl:83-89."""
    q0 = F.MeanFieldGaussian(np.zeros(2), np.ones(2))
    st = Op.sgd_init(q0, Op.Descent(1e-3), Op.NoAveraging())
    with pytest.raises(RuntimeError, match="diverged"):
        Op.sgd_step(st, q0, lambda p, t: (np.nan, np.zeros(4), {}), Op.Descent(1e-3), Op.ClipScale(), Op.NoAveraging())


# --- test/general/gauss_expected_grad_hess.jl ----------------------------------------------------
class _TestQuad:
    """TestQuad of the reference test (:1-27): log pi(x) = -x' S x / 2, gradient -S x, Hessian -S."""

    def __init__(self, S, capability):
        self.S, self.capability = np.asarray(S, float), capability

    def logdensity_and_gradient_batch(self, Z):
        return -0.5 * np.sum(Z * (self.S @ Z), axis=0), -(self.S @ Z)

    def hessian_batch(self, Z):
        return np.broadcast_to(-self.S, (Z.shape[1],) + self.S.shape)


@pytest.mark.parametrize("capability", [1, 2])
def test_gauss_expected_grad_hess_known_answer(capability):
    """gauss_expected_grad_hess.jl (test) :29-56: d = 2, Sigma = [2 -0.1; -0.1 2], q = N(1, 0.1^2 I), n = 10^6:
    E grad = -Sigma mu and E Hessian = -Sigma within atol 1e-1, for first- and second-order targets."""
    S = np.array([[2.0, -0.1], [-0.1, 2.0]])
    q = F.FullRankGaussian(np.ones(2), 0.1 * np.eye(2))
    u = P.normal_matrix(77, 0, 2, 10 ** 6)
    lp, g, H = O.gaussian_expectation_gradient_and_hessian(q, _TestQuad(S, capability), u)
    assert np.allclose(g, -S @ q.location, atol=1e-1)
    assert np.allclose(H, -S, atol=1e-1)
    assert np.isfinite(lp)


# --- test/families/location_scale_low_rank.jl (groundwork for SURVEY 8f rank 4: oracle only) ----------------
@pytest.mark.parametrize("rank", [1, 2])
def test_lowrank_family_logpdf_entropy_moments(rank):
    """location_scale_low_rank.jl (test) :3-90, basedist = :gaussian: logpdf / entropy / mean / var / cov vs
    MvNormal(location, Diagonal(scale_diag^2) + U U'), sample moments within 1e-2."""
    d = 10
    loc = P.normal_matrix(41, 0, d, 1)[:, 0]
    U = P.normal_matrix(42, 0, d, rank)
    q = F.LowRankGaussian(loc, np.ones(d), U)
    Sigma = np.eye(d) + U @ U.T
    ref = stats.multivariate_normal(loc, Sigma)
    z = q.rand_from_eps(P.normal_matrix(43, 0, d, 5), P.normal_matrix(44, 0, rank, 5))
    assert np.allclose(q.logpdf(z), ref.logpdf(z.T), rtol=1e-10)
    assert np.isclose(q.entropy(), ref.entropy(), rtol=1e-12)
    assert len(q) == d and np.allclose(q.mean(), loc) and np.allclose(q.var(), np.diag(Sigma)) and np.allclose(q.cov(), Sigma)
    Z = q.rand_from_eps(P.normal_matrix(45, 0, d, 10 ** 6), P.normal_matrix(46, 0, rank, 10 ** 6))
    assert np.allclose(Z.mean(axis=1), loc, rtol=1e-2, atol=1e-2)
    assert np.allclose(Z.var(axis=1), np.diag(Sigma), rtol=1e-2)
    lam = q.destructure()
    assert lam.shape == (2 * d + d * rank,)
    q2 = q.restructure(lam)
    assert np.array_equal(q2.scale_factors, q.scale_factors) and np.array_equal(q2.scale_diag, q.scale_diag)


def test_lowrank_repgrad_closed_form_vs_finite_differences():
    """The closed-form RepGradELBO + ClosedFormEntropy gradient over the low-rank family equals central finite
    differences of the forward closure (what the reference's AD backend returns)."""
    d, r, M = 6, 2, 5
    prob = Mo.NormalDiag(np.linspace(-1, 1, d), np.linspace(0.5, 1.5, d))
    q = F.LowRankGaussian(0.1 * np.arange(d), 0.5 + 0.1 * np.arange(d), 0.3 * P.normal_matrix(51, 0, d, r))
    u1, u2 = P.normal_matrix(52, 0, d, M), P.normal_matrix(53, 0, r, M)
    lam = q.destructure()
    v, g, e = O.repgrad_lowrank_value_and_gradient(lam, q, prob, u1, u2)

    def f(x):
        return O.repgrad_lowrank_value_and_gradient(x, q, prob, u1, u2)[0]
    h = 1e-6
    fd = np.array([(f(lam + h * np.eye(len(lam))[k]) - f(lam - h * np.eye(len(lam))[k])) / (2 * h) for k in range(len(lam))])
    assert np.allclose(g, fd, rtol=1e-6, atol=1e-7)
    assert np.isclose(v, -e)


# --- non-standard / non-Gaussian base distributions (groundwork for SURVEY 8f rank 4: oracle only) -----------------
@pytest.mark.parametrize("covtype", ["meanfield", "fullrank"])
def test_family_nonstandard_gaussian_base(covtype):
    """location_scale.jl (test) :22-31, basedist = :gaussian_nonstd: MvLocationScale(location, scale, Normal(3, 3))
    equals MvNormal(location + scale * fill(3, d), 9 scale scale')."""
    d = 10
    loc = P.normal_matrix(71, 0, d, 1)[:, 0]
    Ld = np.tril(np.eye(d) + np.ones((d, d)) / 2)
    scale = np.ones(d) if covtype == "meanfield" else Ld
    q = F.MvLocationScale(loc, scale, F.NormalDist(3.0, 3.0))
    C = np.diag(scale) if covtype == "meanfield" else scale
    ref = stats.multivariate_normal(loc + C @ np.full(d, 3.0), 9.0 * C @ C.T)
    z = q.rand_from_eps(q.dist.from_normal(P.normal_matrix(72, 0, d, 6)))
    assert np.allclose(q.logpdf(z), ref.logpdf(z.T), rtol=1e-10)
    assert np.isclose(q.entropy(), ref.entropy(), rtol=1e-12)
    assert np.allclose(q.mean(), ref.mean) and np.allclose(q.cov(), ref.cov) and np.allclose(q.var(), np.diag(ref.cov))


@pytest.mark.parametrize("dist,ref", [(F.LaplaceDist(), stats.laplace()), (F.TDistBase(5.0), stats.t(5.0))])
def test_base_distributions_against_scipy(dist, ref):
    """Laplace() and TDist(nu) bases (docs/src/families.md:72-101): logpdf, entropy, variance, score and the
    sampling transform against scipy."""
    u = np.linspace(-4, 4, 41)
    assert np.allclose(dist.logpdf(u), ref.logpdf(u), rtol=1e-12)
    assert np.isclose(dist.entropy(), ref.entropy(), rtol=1e-12) and np.isclose(dist.var(), ref.var())
    h = 1e-6
    mid = u[np.abs(u) > 1e-3]
    assert np.allclose(dist.score(mid), (dist.logpdf(mid + h) - dist.logpdf(mid - h)) / (2 * h), rtol=1e-5, atol=1e-7)
    x = dist.from_normal(P.normal_matrix(73, 0, 1, 10 ** 6)[0])
    assert abs(x.mean()) < 1e-2 and np.isclose(x.var(), ref.var(), rtol=3e-2)
    assert stats.kstest(x[:20000], ref.cdf).pvalue > 1e-3


@pytest.mark.parametrize("covtype", ["meanfield", "fullrank"])
@pytest.mark.parametrize("dist", [F.LaplaceDist(), F.TDistBase(5.0)], ids=["laplace", "tdist5"])
@pytest.mark.parametrize("entropy", ["ClosedFormEntropy", "StickingTheLandingEntropy"])
def test_nongaussian_base_repgrad_closed_form_vs_fd(covtype, dist, entropy):
    """Closed-form RepGradELBO gradient for Student-t / Laplace location-scale families vs central differences of
    the forward closure (q frozen inside log q for STL)."""
    d, M = 5, 6
    prob = Mo.NormalDiag(np.linspace(-1, 1, d), np.linspace(0.5, 1.5, d))
    scale = 0.5 + 0.1 * np.arange(d) if covtype == "meanfield" else np.tril(0.05 * np.ones((d, d))) + np.diag(0.5 + 0.1 * np.arange(d))
    q = F.MvLocationScale(0.1 * np.arange(d), scale, dist)
    u = dist.from_normal(P.normal_matrix(74, 0, d, M))
    lam = q.destructure()
    v, g, e = O.repgrad_general_base_value_and_gradient(lam, q, prob, u, entropy)

    def forward(x):
        qx = q.restructure(x)
        Z = qx.rand_from_eps(u)
        logp, _ = prob.logdensity_and_gradient_batch(Z)
        ent = qx.entropy() if entropy == "ClosedFormEntropy" else -np.mean(q.logpdf(Z))
        return -(np.mean(logp) + ent)
    hh = 1e-6
    fd = np.array([(forward(lam + hh * ek) - forward(lam - hh * ek)) / (2 * hh) for ek in np.eye(len(lam))])
    if covtype == "fullrank":   # the strict upper triangle is not a parameter of a LowerTriangular scale
        D = d
        mask = np.concatenate([np.ones(D, bool), np.tril(np.ones((D, D), bool)).reshape(-1, order="F")])
        fd = fd * mask
    assert np.isclose(v, forward(lam))
    assert np.allclose(g, fd, rtol=5e-6, atol=5e-7)


def test_batch_match_fisher_divergence_vanishes_at_truth():
    """rand_batch_match_samples_with_objective! (src/algorithms/fisherminbatchmatch.jl:81-111): at q == pi (a Gaussian
    with scale C) grad log pi(z) = -C^-T u, so -u - C' grad log pi(z) == 0 and the Fisher-divergence estimate is zero
    for ANY draws (:100-108); away from it it is positive; log pi_avg is the sample mean of the log-density."""
    D, n = 6, 40
    mu = np.linspace(-1.0, 2.0, D)
    C = np.tril(0.1 * np.ones((D, D))) + np.diag(0.5 + 0.1 * np.arange(D))
    prob = Mo.NormalDense(mu, C)
    u = P.normal_matrix(7, 3, D, n)
    uo, z, g, fisher, lp = O.rand_batch_match_samples_with_objective(F.FullRankGaussian(mu, C), prob, u)
    assert uo is u and np.allclose(z, C @ u + mu[:, None])
    assert fisher < 1e-24
    assert np.isclose(lp, np.mean(prob.logdensity_and_gradient_batch(z)[0]))
    _, _, _, fisher_off, _ = O.rand_batch_match_samples_with_objective(F.FullRankGaussian(mu + 0.3, 1.2 * C), prob, u)
    assert fisher_off > 1e-2
