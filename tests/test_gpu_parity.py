"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical Philox eps.

Every test here needs a B200 (`-m gpu`).  Tolerances are stated next to each comparison:
fp32 SIMT arithmetic vs the fp64 oracle ~1e-5 relative; the TF32 tensor-core contractions
(10-bit mantissa operands, fp32 accumulate) 5e-4 on the ELBO and 2e-3 on the gradient norm
(BASELINE.md section 4).
"""
import numpy as np
import pytest

from oracle import family as F, models as Mo, objectives as O, optim as Op, philox as P

pytestmark = pytest.mark.gpu

KEY = 0x38BEF07CF9CC549D   # the reference's test seed (test/algorithms/klminrepgraddescent.jl:43)


@pytest.fixture(scope="module")
def ctx(avi):
    c = avi.Context(0)
    yield c
    c.close()


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def make_q(avi, kind, D, oracle=False):
    mu = (0.1 * np.arange(D) - 0.2)
    if kind == "meanfield":
        s = 0.5 + 0.05 * np.arange(D)
        return (F.MeanFieldGaussian(mu, s) if oracle else avi.MeanFieldGaussian(mu.astype(np.float32), s.astype(np.float32)))
    Lm = np.tril(0.05 * np.ones((D, D))) + np.diag(0.5 + 0.05 * np.arange(D))
    return (F.FullRankGaussian(mu, Lm) if oracle else avi.FullRankGaussian(mu.astype(np.float32), Lm.astype(np.float32)))


ENT = {"ClosedFormEntropy": "ClosedFormEntropy", "MonteCarloEntropy": "MonteCarloEntropy",
       "StickingTheLandingEntropy": "StickingTheLandingEntropy",
       "ClosedFormEntropyZeroGradient": "ClosedFormEntropyZeroGradient",
       "StickingTheLandingEntropyZeroGradient": "StickingTheLandingEntropyZeroGradient"}


# --- K1: rand(rng, q, M) ------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
@pytest.mark.parametrize("D,M", [(2, 4), (5, 10), (37, 33), (130, 7)])
def test_rand_matches_oracle(avi, ctx, kind, D, M):
    """location_scale.jl:71-87 with eps = BoxMuller(Philox(key, step, m, i)): bit-identical uniforms,
    fp32 vs fp64 transcendental rounding only (abs 2e-6 on eps)."""
    prob = avi.MvNormalDiag(ctx, np.zeros(D), np.ones(D))
    q, qo = make_q(avi, kind, D), make_q(avi, kind, D, oracle=True)
    obj = avi.Objective(KEY, avi.RepGradELBO(M), q, prob)
    obj.seed(KEY, 3)
    Z, E = obj.rand(q)
    eps = P.normal_matrix(KEY, 3, D, M)
    assert np.abs(E - eps).max() < 4e-6
    assert np.abs(Z - qo.rand_from_eps(eps)).max() < 2e-5
    obj.close(); prob.close()


# --- RepGradELBO / ScoreGradELBO on the Gaussian test target -------------------------------------
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
@pytest.mark.parametrize("entropy", list(ENT))
@pytest.mark.parametrize("D,M", [(5, 10), (33, 70)])
def test_repgrad_matches_oracle(avi, ctx, kind, entropy, D, M):
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    prob, probo = avi.MvNormalDiag(ctx, mu_t, sg_t), Mo.NormalDiag(mu_t, sg_t)
    q, qo = make_q(avi, kind, D), make_q(avi, kind, D, oracle=True)
    obj = avi.Objective(KEY, avi.RepGradELBO(M, getattr(avi, entropy)()), q, prob)
    v, g, e = obj.estimate_gradient(q.destructure())
    eps = P.normal_matrix(KEY, 0, D, M)
    vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, eps, entropy)
    assert abs(v - vo) <= 2e-5 * max(1, abs(vo)) and abs(e - eo) <= 2e-5 * max(1, abs(eo))
    assert relerr(g, go) < 2e-5
    assert obj.step_counter() == 1          # the step counter advanced: next call draws fresh eps
    v2, g2, _ = obj.estimate_gradient(q.destructure())
    vo2, go2, _ = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, P.normal_matrix(KEY, 1, D, M), entropy)
    assert abs(v2 - vo2) <= 2e-5 * max(1, abs(vo2)) and relerr(g2, go2) < 2e-5
    obj.close(); prob.close()


@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
@pytest.mark.parametrize("D,M", [(5, 10), (33, 70)])
def test_scoregrad_matches_oracle(avi, ctx, kind, D, M):
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    prob, probo = avi.MvNormalDiag(ctx, mu_t, sg_t), Mo.NormalDiag(mu_t, sg_t)
    q, qo = make_q(avi, kind, D), make_q(avi, kind, D, oracle=True)
    obj = avi.Objective(KEY, avi.ScoreGradELBO(M), q, prob)
    for step in range(2):   # the second call exercises the centring shift carried between calls
        v, g, e = obj.estimate_gradient(q.destructure())
        vo, go, eo = O.scoregrad_value_and_gradient(qo.destructure(), qo, probo, P.normal_matrix(KEY, step, D, M))
        assert abs(v - vo) <= 1e-4 * max(1, abs(vo)) and abs(e - eo) <= 2e-5 * max(1, abs(eo))
        assert relerr(g, go) < 2e-4
    obj.close(); prob.close()


# --- GLM targets ---------------------------------------------------------------------------------
def glm_pair(avi, ctx, name, n, d, gemm, n_data=None):
    fam = "gaussian" if name == "gaussglm" else "bernoulli_logit"
    X, y = Mo.synth_glm_data(n, d, seed=5, family=fam)
    if name == "gaussglm":
        return avi.GaussGLM(ctx, X, y, n_data=n_data, gemm=gemm), Mo.GaussGLM(X, y, n_data=n_data)
    variant = "basic" if name == "logreg_basic" else "subsampling"
    return (avi.LogReg(ctx, X, y, n_data=n_data, variant=variant, gemm=gemm),
            Mo.LogReg(X, y, n_data=n_data, variant=variant))


@pytest.mark.parametrize("name", ["logreg_subsampling", "logreg_basic", "gaussglm"])
@pytest.mark.parametrize("gemm,tol_l,tol_g", [("fp32", 2e-6, 2e-5), ("tf32", 5e-4, 2e-3), ("tf32x3", 4e-6, 4e-5)])
@pytest.mark.parametrize("n,d,M", [(40, 4, 3), (300, 37, 17), (1000, 160, 130)])
def test_glm_logdensity_and_gradient(avi, ctx, name, gemm, tol_l, tol_g, n, d, M):
    prob, probo = glm_pair(avi, ctx, name, n, d, gemm, n_data=3 * n)
    Z = 0.3 * P.normal_matrix(1, 0, d + 1, M)
    lp, G = prob.logdensity_and_gradient(Z.astype(np.float32))
    lpo, Go = probo.logdensity_and_gradient_batch(Z.astype(np.float32).astype(np.float64))
    assert np.abs(lp - lpo).max() <= tol_l * np.abs(lpo).max()
    assert relerr(G, Go) < tol_g
    prob.close()


@pytest.mark.parametrize("gemm,tol_v,tol_g", [("fp32", 1e-5, 5e-5), ("tf32", 5e-4, 2e-3), ("tf32x3", 1e-5, 5e-5)])
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
@pytest.mark.parametrize("entropy", ["ClosedFormEntropy", "StickingTheLandingEntropy"])
def test_repgrad_logreg_matches_oracle(avi, ctx, gemm, tol_v, tol_g, kind, entropy):
    n, d, M = 500, 63, 40
    prob, probo = glm_pair(avi, ctx, "logreg_subsampling", n, d, gemm)
    D = d + 1
    mu = 0.05 * np.cos(np.arange(D))
    if kind == "meanfield":
        q, qo = avi.MeanFieldGaussian(mu.astype(np.float32), np.full(D, 0.3, np.float32)), F.MeanFieldGaussian(mu, np.full(D, 0.3))
    else:
        Lm = 0.3 * np.eye(D) + np.tril(0.01 * np.ones((D, D)), -1)
        q, qo = avi.FullRankGaussian(mu.astype(np.float32), Lm.astype(np.float32)), F.FullRankGaussian(mu, Lm)
    obj = avi.Objective(KEY, avi.RepGradELBO(M, getattr(avi, entropy)()), q, prob)
    v, g, e = obj.estimate_gradient(q.destructure())
    vo, go, eo = O.repgrad_value_and_gradient(qo.destructure().astype(np.float32).astype(np.float64), qo, probo,
                                              P.normal_matrix(KEY, 0, D, M), entropy)
    assert abs(v - vo) <= tol_v * abs(vo)
    assert relerr(g, go) < tol_g
    obj.close(); prob.close()


def test_scoregrad_gaussglm_matches_oracle(avi, ctx):
    n, d, M = 400, 31, 64
    prob, probo = glm_pair(avi, ctx, "gaussglm", n, d, "tf32")
    D = d + 1
    q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.2, np.float32))
    qo = F.MeanFieldGaussian(np.zeros(D), np.full(D, 0.2, np.float32).astype(np.float64))
    obj = avi.Objective(KEY, avi.ScoreGradELBO(M), q, prob)
    v, g, e = obj.estimate_gradient(q.destructure())
    vo, go, eo = O.scoregrad_value_and_gradient(qo.destructure(), qo, probo, P.normal_matrix(KEY, 0, D, M))
    assert abs(e - eo) <= 5e-4 * abs(eo)
    assert abs(v - vo) <= 2e-2 * abs(vo)          # variance of f: differences of O(1e3) numbers in fp32/TF32
    assert relerr(g, go) < 2e-2
    obj.close(); prob.close()


def test_subsample_matches_oracle(avi, ctx):
    """AdvancedVI.subsample: rows idx, likelihood scaled by n_data / batch (subsampling.md:99-102)."""
    n, d, M = 256, 20, 8
    prob, probo = glm_pair(avi, ctx, "logreg_subsampling", n, d, "tf32")
    idx = np.array([5, 17, 3, 200, 255, 0, 99, 42, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16])
    Z = (0.3 * P.normal_matrix(2, 0, d + 1, M)).astype(np.float32)
    lp, G = prob.subsample(idx).logdensity_and_gradient(Z)
    lpo, Go = probo.subsample(idx).logdensity_and_gradient_batch(Z.astype(np.float64))
    assert np.abs(lp - lpo).max() <= 5e-4 * np.abs(lpo).max() and relerr(G, Go) < 2e-3
    lp2, _ = prob.subsample(None).logdensity_and_gradient(Z)
    assert np.abs(lp2 - probo.logdensity_batch(Z.astype(np.float64))).max() <= 5e-4 * np.abs(lp2).max()
    prob.close()
    # the 3xTF32 mode rebuilds the split operands of the batch (minibatch gather of [hi | lo | hi] rows)
    prob3 = avi.LogReg(ctx, probo.X.astype(np.float32), probo.y.astype(np.float32), gemm="tf32x3")
    lp3, G3 = prob3.subsample(idx).logdensity_and_gradient(Z)
    assert np.abs(lp3 - lpo).max() <= 4e-6 * np.abs(lpo).max() and relerr(G3, Go) < 4e-5
    prob3.close()


def test_hostcallback_target(avi, ctx):
    """Any LogDensityProblem through the per-sample callback; a deliberately wrong gradient must come
    out unchanged (test/general/mixedad_logdensity.jl:16-61: capability >= 1 => the target's own gradient)."""
    D, M = 4, 6
    mu = np.arange(D, dtype=np.float64)

    def fn(z):
        return -0.5 * float(np.sum((z - mu) ** 2)), np.full(D, 7.0)     # wrong on purpose
    prob = avi.HostCallbackProblem(ctx, D, fn, capability=1)
    q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
    obj = avi.Objective(KEY, avi.RepGradELBO(M), q, prob)
    v, g, e = obj.estimate_gradient(q.destructure())
    eps = P.normal_matrix(KEY, 0, D, M)
    assert np.allclose(g[:D], -7.0, rtol=1e-6)
    assert np.allclose(g[D:], -7.0 * eps.mean(axis=1) - 1.0, rtol=1e-5, atol=1e-5)
    # capability 0 cannot drive RepGradELBO, but ScoreGradELBO only needs logdensity
    prob0 = avi.HostCallbackProblem(ctx, D, lambda z: -0.5 * float(np.sum((z - mu) ** 2)), capability=0)
    with pytest.raises(avi.AviError):
        avi.Objective(KEY, avi.RepGradELBO(M), q, prob0)
    o2 = avi.Objective(KEY, avi.ScoreGradELBO(M), q, prob0)
    v2, g2, e2 = o2.estimate_gradient(q.destructure())
    probo = type("T", (), {"logdensity_batch": staticmethod(lambda Z: -0.5 * np.sum((Z - mu[:, None]) ** 2, axis=0))})
    qo = F.MeanFieldGaussian(np.zeros(D), np.ones(D))
    vo, go, eo = O.scoregrad_value_and_gradient(qo.destructure(), qo, probo, eps)
    assert abs(e2 - eo) < 1e-4 * abs(eo) and relerr(g2, go) < 1e-3
    obj.close(); o2.close(); prob.close(); prob0.close()


def test_hostcallback_capability0_fallback_gradient(avi, ctx):
    """A capability-0 target under RepGradELBO (src/algorithms/repgradelbo.jl:50-62: the reference differentiates
    through `logdensity`): the glue supplies the per-sample gradient (central finite differences by default, or any
    callable), and the result matches the oracle, which uses the analytic gradient of the same target."""
    D, M = 5, 12
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    probo = Mo.NormalDiag(mu_t, sg_t)
    calls = []

    def logdensity_only(z):
        calls.append(1)
        return float(probo.logdensity(np.asarray(z, np.float64)))
    q = avi.MeanFieldGaussian((0.1 * np.arange(D)).astype(np.float32), np.full(D, 0.4, np.float32))
    qo = F.MeanFieldGaussian((0.1 * np.arange(D)).astype(np.float32).astype(np.float64), np.full(D, 0.4, np.float32).astype(np.float64))
    eps = P.normal_matrix(KEY, 0, D, M)
    vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, eps, "StickingTheLandingEntropy")
    for fb in ("central_fd", lambda f, z: -(z - mu_t) / sg_t ** 2):
        prob = avi.HostCallbackProblem(ctx, D, logdensity_only, capability=0, fallback_gradient=fb)
        assert prob.capability == 1
        obj = avi.Objective(KEY, avi.RepGradELBO(M, avi.StickingTheLandingEntropy()), q, prob)
        v, g, e = obj.estimate_gradient(q.destructure())
        assert abs(v - vo) <= 2e-5 * abs(vo)
        assert relerr(g, go) < (2e-4 if fb == "central_fd" else 2e-5)     # finite differences in Float64, h = 1e-5
        obj.close(); prob.close()
    assert len(calls) >= M * (2 * D + 1)       # one logdensity call per sample plus 2 D for its central differences


# --- the reference's known-answer tests on the GPU path ---------------------------------------------
@pytest.mark.parametrize("alg_name", ["rep", "score", "prox"])
def test_estimate_objective_zero_at_truth(avi, ctx, alg_name):
    """klminrepgraddescent.jl:23-38, klminscoregraddescent.jl:23-38, klminrepgradproxdescent.jl:23-38."""
    D = 5
    prob = avi.MvNormalDiag(ctx, np.full(D, 5.0), np.full(D, 0.3))
    q_true = avi.MeanFieldGaussian(np.full(D, 5.0, np.float32), np.full(D, 0.3, np.float32))
    alg = {"rep": avi.KLMinRepGradDescent, "score": avi.KLMinScoreGradDescent, "prox": avi.KLMinRepGradProxDescent}[alg_name]()
    est = avi.estimate_objective(KEY, alg, q_true, prob, n_samples=10 ** 5)
    assert abs(est) < 1e-2
    prob.close()


@pytest.mark.parametrize("M", [1, 10])
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
def test_stl_gradient_vanishes_at_truth(avi, ctx, M, kind):
    """klminrepgraddescent.jl:66-87 (atol 1e-5 there in Float64; fp32 here: the rounding of z = mu + L eps
    (~3e-7) is amplified by 1/sigma^2 = 11 in g = -(z - mu)/sigma^2, so 5e-5)."""
    D = 5
    prob = avi.MvNormalDiag(ctx, np.full(D, 5.0), np.full(D, 0.3))
    if kind == "meanfield":
        q = avi.MeanFieldGaussian(np.full(D, 5.0, np.float32), np.full(D, 0.3, np.float32))
    else:
        q = avi.FullRankGaussian(np.full(D, 5.0, np.float32), (0.3 * np.eye(D)).astype(np.float32))
    obj = avi.Objective(KEY, avi.RepGradELBO(M, avi.StickingTheLandingEntropy()), q, prob)
    _, g, _ = obj.estimate_gradient(q.destructure())
    assert np.linalg.norm(g) < 5e-5
    obj.close(); prob.close()


def _oracle_run(qo, probo, T, M, rule, op, avg, kind, entropy, key):
    st = Op.sgd_init(qo, rule, avg)

    def grad_fn(params, t):
        eps = P.normal_matrix(key, t - 1, len(qo), M)
        if kind == "score":
            v, g, e = O.scoregrad_value_and_gradient(params, qo, probo, eps)
        else:
            v, g, e = O.repgrad_value_and_gradient(params, qo, probo, eps, entropy)
        return v, g, dict(elbo=e)
    elbos = [Op.sgd_step(st, qo, grad_fn, rule, op, avg)["elbo"] for _ in range(T)]
    return st, np.array(elbos)


RULES = {
    "descent": (lambda a: a.Descent(1e-2), lambda: Op.Descent(1e-2)),
    "adam": (lambda a: a.Adam(1e-2), lambda: Op.Adam(1e-2)),
    # alpha = 1e-2 instead of the default 1e-6: with r0 ~ 1e-6 the first displacements |x - x0| sit at the
    # fp32 rounding level of x itself, so an fp32 run (the reference's Float32 run too) cannot track fp64
    "dog": (lambda a: a.DoG(1e-2), lambda: Op.DoG(1e-2)),
    "dowg": (lambda a: a.DoWG(1e-2), lambda: Op.DoWG(1e-2)),
}


@pytest.mark.parametrize("rule", list(RULES))
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
def test_fused_step_trajectory_matches_oracle(avi, ctx, rule, kind):
    """20 fused iterations (update! + ClipScale + PolynomialAveraging, common.jl:91-94) follow the fp64
    oracle trajectory driven by the same eps: lambda and the averaged iterate within 1e-4."""
    D, M, T = 6, 8, 20
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    prob, probo = avi.MvNormalDiag(ctx, mu_t, sg_t), Mo.NormalDiag(mu_t, sg_t)
    q, qo = make_q(avi, kind, D), make_q(avi, kind, D, oracle=True)
    alg = avi.KLMinRepGradDescent(optimizer=RULES[rule][0](avi), n_samples=M, operator=avi.ClipScale())
    qa, info, state = avi.optimize(KEY, alg, T, prob, q)
    st, elbos = _oracle_run(qo, probo, T, M, RULES[rule][1](), Op.ClipScale(), Op.PolynomialAveraging(), "rep",
                            "ClosedFormEntropy", KEY)
    lam, avg, _ = state.params()
    assert [i["iteration"] for i in info] == list(range(1, T + 1))
    assert np.allclose([i["elbo"] for i in info], elbos, rtol=2e-4, atol=2e-4)
    assert relerr(lam, st.params) < 1e-4 and relerr(avg, st.avg_st[0]) < 1e-4
    assert relerr(qa.destructure(), st.avg_st[0]) < 1e-4
    state.close(); state.obj.close(); prob.close()


def test_prox_descent_trajectory_matches_oracle(avi, ctx):
    D, M, T = 5, 4, 15
    prob, probo = avi.MvNormalDiag(ctx, np.full(D, 5.0), np.full(D, 0.3)), Mo.NormalDiag(np.full(D, 5.0), np.full(D, 0.3))
    q, qo = make_q(avi, "meanfield", D), make_q(avi, "meanfield", D, oracle=True)
    alg = avi.KLMinRepGradProxDescent(optimizer=avi.DoWG(1e-2), n_samples=M)
    _, info, state = avi.optimize(KEY, alg, T, prob, q)
    st, elbos = _oracle_run(qo, probo, T, M, Op.DoWG(1e-2), Op.ProximalLocationScaleEntropy(), Op.PolynomialAveraging(),
                            "rep", "ClosedFormEntropyZeroGradient", KEY)
    lam, avg, _ = state.params()
    assert relerr(lam, st.params) < 2e-4 and relerr(avg, st.avg_st[0]) < 2e-4
    state.close(); state.obj.close(); prob.close()


@pytest.mark.parametrize("objective,M", [("cfe", 1), ("stl", 1), ("score", 100)])
def test_convergence(avi, ctx, objective, M):
    """klminrepgraddescent.jl:105-121 / klminscoregraddescent.jl:82-97: T = 1000, Descent(1e-3), ClipScale."""
    D = 5
    prob = avi.MvNormalDiag(ctx, np.full(D, 5.0), np.full(D, 0.3))
    q0 = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
    if objective == "score":
        alg = avi.KLMinScoreGradDescent(optimizer=avi.Descent(1e-3), n_samples=M, operator=avi.ClipScale())
    else:
        ent = avi.ClosedFormEntropy() if objective == "cfe" else avi.StickingTheLandingEntropy()
        alg = avi.KLMinRepGradDescent(optimizer=avi.Descent(1e-3), entropy=ent, n_samples=M, operator=avi.ClipScale())
    q, info, state = avi.optimize(KEY, alg, 1000, prob, q0)
    d0 = np.sum((q0.location - 5.0) ** 2) + np.sum((q0.scale - 0.3) ** 2)
    d1 = np.sum((q.location - 5.0) ** 2) + np.sum((q.scale - 0.3) ** 2)
    assert d1 <= d0 / 2 and len(info) == 1000
    state.close(); state.obj.close(); prob.close()


def test_determinism_and_warm_start(avi, ctx):
    """Same seed => bitwise identical result (klminrepgraddescent.jl:40-57); T/2 + T/2 with the returned
    state == T straight, bitwise (test/general/optimize.jl:27-40), also through export/import of the state."""
    D, T = 5, 10
    prob = avi.MvNormalDiag(ctx, np.full(D, 5.0), np.full(D, 0.3))
    q0 = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))

    def alg():
        return avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=10, operator=avi.ClipScale())
    qa, ia, sa = avi.optimize(KEY, alg(), T, prob, q0)
    qb, ib, sb = avi.optimize(KEY, alg(), T, prob, q0)
    assert np.array_equal(qa.location, qb.location) and np.array_equal(qa.scale, qb.scale)
    assert [i["elbo"] for i in ia] == [i["elbo"] for i in ib]
    qc, ic, sc = avi.optimize(KEY, alg(), T // 2, prob, q0)
    blob = sc.export_bytes()
    qd, id_, sd = avi.optimize(KEY, alg(), T // 2, prob, q0, state=sc)
    assert np.array_equal(qa.location, qd.location) and np.array_equal(qa.scale, qd.scale)
    assert id_[0]["iteration"] == T // 2 + 1
    # restore the half-way state into a fresh optimiser
    qe, _, se = avi.optimize(KEY + 1, alg(), 1, prob, q0)      # different key: must be overwritten by import
    se.import_bytes(blob)
    qf, _, _ = avi.optimize(KEY, alg(), T // 2, prob, q0, state=se)
    assert np.array_equal(qa.location, qf.location) and np.array_equal(qa.scale, qf.scale)
    for s in (sa, sb, sc, se):
        s.close(); s.obj.close()
    prob.close()


def test_callback_and_divergence(avi, ctx):
    """Callback sees iteration == 1:T and its return value is merged into info (test/general/optimize.jl:18-25);
    a non-finite objective raises (src/algorithms/common.jl:83-89)."""
    D = 3
    prob = avi.MvNormalDiag(ctx, np.zeros(D), np.ones(D))
    q0 = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
    seen = []

    def cb(**kw):
        seen.append(kw["iteration"])
        assert kw["params"].shape == (2 * D,) and kw["gradient"].shape == (2 * D,)
        return dict(test_value=kw["iteration"] * 2)
    alg = avi.KLMinRepGradDescent(optimizer=avi.Descent(1e-2), n_samples=2, operator=avi.ClipScale())
    _, info, st = avi.optimize(KEY, alg, 5, prob, q0, callback=cb)
    assert seen == [1, 2, 3, 4, 5] and [i["test_value"] for i in info] == [2, 4, 6, 8, 10]
    st.close(); st.obj.close()
    with pytest.warns(UserWarning, match="IdentityOperator"):
        bad = avi.KLMinRepGradDescent(optimizer=avi.Descent(1e6), n_samples=2)
        with pytest.raises(RuntimeError, match="diverged"):
            avi.optimize(KEY, bad, 50, prob, q0)
    prob.close()


def test_subsampled_objective_epoch_mean_gradient(avi, ctx):
    """test/general/subsampledobj.jl:62-89: the mean over an epoch of minibatch gradients (same MC samples)
    equals the full-batch gradient."""
    n, d, M, bs = 64, 6, 16, 8
    X, y = Mo.synth_glm_data(n, d, seed=9)
    prob = avi.LogReg(ctx, X, y, gemm="fp32")
    D = d + 1
    q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.5, np.float32))
    obj = avi.Objective(KEY, avi.RepGradELBO(M), q, prob)
    lam = q.destructure()
    obj.seed(KEY, 0)
    _, g_full, _ = obj.estimate_gradient(lam)
    sub = avi.ReshufflingBatchSubsampling(np.arange(n), bs)
    grads = []
    for k, batch in sub.reshuffle_batches(KEY, 0):
        prob.subsample(batch)
        obj.seed(KEY, 0)                      # same Monte-Carlo samples for every batch
        grads.append(obj.estimate_gradient(lam)[1])
    assert sorted(np.concatenate([b for _, b in sub.reshuffle_batches(KEY, 0)]).tolist()) == list(range(n))
    assert relerr(np.mean(grads, axis=0), g_full) < 2e-5
    obj.close(); prob.close()


def test_optimize_subsampled_matches_oracle(avi, ctx):
    """SubsampledObjective + ReshufflingBatchSubsampling inside the fused loop (device-side gather per
    iteration) against the oracle state machine fed the same Philox permutation."""
    from oracle import reshuffling as R
    n, d, M, bs, T = 50, 5, 8, 8, 17            # 50 % 8 != 0: exercises the drop-trailing swap
    X, y = Mo.synth_glm_data(n, d, seed=11)
    prob, probo = avi.LogReg(ctx, X, y, gemm="fp32"), Mo.LogReg(X, y)
    D = d + 1
    q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.5, np.float32))
    qo = F.MeanFieldGaussian(np.zeros(D), np.full(D, 0.5))
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=M, operator=avi.ClipScale(),
                                  subsampling=avi.ReshufflingBatchSubsampling(np.arange(n), bs))
    _, info, state = avi.optimize(KEY, alg, T, prob, q)
    sub = R.ReshufflingBatchSubsampling(np.arange(n), bs)
    sst = R.subsampled_init(sub, KEY)
    rule, op, avg = Op.Adam(1e-2), Op.ClipScale(), Op.PolynomialAveraging()
    st = Op.sgd_init(qo, rule, avg)
    infos = []
    for t in range(T):
        def grad_fn(params, it):
            nonlocal sst
            def inner(ps):
                return O.repgrad_value_and_gradient(params, qo, ps, P.normal_matrix(KEY, it - 1, D, M), "ClosedFormEntropy")[:2] + (dict(),)
            v, g, sst, inf = R.subsampled_estimate_gradient(sub, sst, probo, inner)
            return v, g, inf
        infos.append(Op.sgd_step(st, qo, grad_fn, rule, op, avg))
    assert [(i["epoch"], i["step"]) for i in info] == [(i["epoch"], i["step"]) for i in infos]
    lam, avgp, _ = state.params()
    assert relerr(lam, st.params) < 1e-4 and relerr(avgp, st.avg_st[0]) < 1e-4
    state.close(); state.obj.close(); prob.close()


# --- BASELINE.json config 2 at full size: size-independent properties ---------------------------------
def test_c2_full_size_properties(avi, ctx):
    """n = 10000, d = 1024, M = 256 (the benchmark workload), TF32 tensor-core path vs exact fp32 SIMT path
    on the same inputs; linearity of the gradient sums in the likelihood weight; determinism."""
    n, d, M = 10000, 1024, 256
    X, y = Mo.synth_glm_data(n, d, seed=1)
    D = d + 1
    q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32) * 0.1)
    lam = q.destructure()
    res = {}
    for gemm in ("tf32", "fp32", "tf32x3"):
        prob = avi.LogReg(ctx, X, y, gemm=gemm)
        obj = avi.Objective(1, avi.RepGradELBO(M), q, prob)
        res[gemm] = obj.estimate_gradient(lam)
        if gemm == "tf32":
            obj.seed(1, 0)
            again = obj.estimate_gradient(lam)
            assert again[0] == res[gemm][0] and np.array_equal(again[1], res[gemm][1])     # bitwise repeatable
        obj.close(); prob.close()
    (v1, g1, _), (v0, g0, _), (v3, g3, _) = res["tf32"], res["fp32"], res["tf32x3"]
    assert abs(v1 - v0) <= 5e-4 * abs(v0)
    assert relerr(g1, g0) < 2e-3
    assert abs(v3 - v0) <= 1e-5 * abs(v0) and relerr(g3, g0) < 5e-5      # 3xTF32 == fp32-grade
    # oracle on a bounded slice of the same workload: 16 samples
    probo = Mo.LogReg(X, y)
    prob = avi.LogReg(ctx, X, y, gemm="tf32")
    Z = (0.1 * P.normal_matrix(1, 0, D, 16)).astype(np.float32)
    lp, G = prob.logdensity_and_gradient(Z)
    lpo, Go = probo.logdensity_and_gradient_batch(Z.astype(np.float64))
    assert np.abs(lp - lpo).max() <= 5e-4 * np.abs(lpo).max() and relerr(G, Go) < 2e-3
    prob.close()


# --- rand_batch_match_samples_with_objective! (src/algorithms/fisherminbatchmatch.jl:81-111) ----------------------
@pytest.mark.parametrize("n_samples", [48, 4500])      # one chunk / two chunks of 4096
def test_batch_match_sampling_stage_matches_oracle(avi, ctx, n_samples):
    """The sampling stage of FisherMinBatchMatch on the device vs the oracle on the same Philox draws: u, z = C u + mu,
    the per-sample gradients, the Fisher-divergence estimate sum |-u - C' grad|^2 / n and mean log pi (fp32 SIMT
    arithmetic vs fp64: 2e-5 relative; the draws within 4e-6 absolute)."""
    n, d = 300, 11
    X, y = Mo.synth_glm_data(n, d, seed=9)
    D = d + 1
    prob, probo = avi.LogReg(ctx, X, y, gemm="fp32"), Mo.LogReg(X, y)
    mu = 0.1 * P.normal_matrix(5, 0, D, 1)[:, 0]
    Lm = np.tril(0.05 * P.normal_matrix(6, 0, D, D)) + 0.4 * np.eye(D)
    q = avi.FullRankGaussian(mu.astype(np.float32), Lm.astype(np.float32))
    qo = F.FullRankGaussian(mu.astype(np.float32).astype(np.float64), Lm.astype(np.float32).astype(np.float64))
    obj = avi.Objective(KEY, avi.RepGradELBO(8), q, prob)
    for step in range(2):      # the call advances the step: the second one uses fresh draws
        u, z, g, fisher, lp = obj.rand_batch_match_samples_with_objective(q, n_samples)
        uo, zo, go, fo, lpo = O.rand_batch_match_samples_with_objective(qo, probo, P.normal_matrix(KEY, step, D, n_samples))
        assert u.shape == (D, n_samples) and np.abs(u - uo).max() < 4e-6
        assert np.abs(z - zo).max() < 2e-5 and relerr(g, go) < 2e-5
        assert abs(fisher - fo) <= 2e-5 * abs(fo) and abs(lp - lpo) <= 2e-5 * abs(lpo)
    # the reference's capability check (fisherminbatchmatch.jl:63-70) and family requirement
    qmf = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
    omf = avi.Objective(KEY, avi.RepGradELBO(8), qmf, prob)
    with pytest.raises(avi.AviError):
        omf.rand_batch_match_samples_with_objective(qmf, 4)
    omf.close(); obj.close(); prob.close()


# --- gaussian_expectation_gradient_and_hessian! (src/algorithms/gauss_expected_grad_hess.jl) ---------------------
@pytest.mark.parametrize("n_samples", [64, 5000])      # one chunk / several chunks of 4096 (+ SIMT tail path)
def test_gauss_expected_grad_hess_matches_oracle(avi, ctx, n_samples):
    """Stein branch on the device vs the oracle on the same Philox draws: logistic-regression target, dense scale."""
    n, d = 300, 11
    X, y = Mo.synth_glm_data(n, d, seed=9)
    D = d + 1
    prob, probo = avi.LogReg(ctx, X, y, gemm="tf32x3"), Mo.LogReg(X, y)
    mu = 0.1 * P.normal_matrix(5, 0, D, 1)[:, 0]
    Lm = np.tril(0.05 * P.normal_matrix(6, 0, D, D)) + 0.4 * np.eye(D)
    q = avi.FullRankGaussian(mu.astype(np.float32), Lm.astype(np.float32))
    qo = F.FullRankGaussian(mu.astype(np.float32).astype(np.float64), Lm.astype(np.float32).astype(np.float64))
    obj = avi.Objective(KEY, avi.RepGradELBO(8), q, prob)
    lp, g, H = obj.gaussian_expectation_gradient_and_hessian(q, n_samples)
    lpo, go, Ho = O.gaussian_expectation_gradient_and_hessian(qo, probo, P.normal_matrix(KEY, 0, D, n_samples))
    assert abs(lp - lpo) <= 2e-5 * abs(lpo)
    assert relerr(g, go) < 5e-5 and relerr(H, Ho) < 2e-4
    # the call advanced the step: a second call uses fresh draws
    lp2, g2, H2 = obj.gaussian_expectation_gradient_and_hessian(q, n_samples)
    lpo2, go2, Ho2 = O.gaussian_expectation_gradient_and_hessian(qo, probo, P.normal_matrix(KEY, 1, D, n_samples))
    assert relerr(g2, go2) < 5e-5 and relerr(H2, Ho2) < 2e-4 and not np.array_equal(g, g2)
    obj.close(); prob.close()


def test_gauss_expected_grad_hess_known_answer(avi, ctx):
    """test/general/gauss_expected_grad_hess.jl:29-56 on the device (first-order capability): q = N(1, 0.1^2 I); for a
    Gaussian target N(mu_t, diag sigma_t^2): E grad = -(m - mu_t) / sigma_t^2, E Hessian = -diag(1 / sigma_t^2), atol 1e-1."""
    D = 3
    mu_t, sg_t = np.array([0.5, -0.2, 1.5]), np.array([0.8, 1.0, 0.7])
    prob = avi.MvNormalDiag(ctx, mu_t, sg_t)
    q = avi.FullRankGaussian(np.ones(D, np.float32), (0.1 * np.eye(D)).astype(np.float32))
    g, H = np.zeros(D, np.float32), np.zeros((D, D), np.float32)
    lp, g, H = avi.gaussian_expectation_gradient_and_hessian(KEY, q, 10 ** 6, g, H, prob)
    assert np.allclose(g, -(1.0 - mu_t) / sg_t ** 2, atol=1e-1)
    assert np.allclose(H, -np.diag(1.0 / sg_t ** 2), atol=1e-1)
    # a mean-field family has no triangular scale: rejected like the reference's method signature (:20-27)
    qm = avi.MeanFieldGaussian(np.ones(D, np.float32), np.full(D, 0.1, np.float32))
    with pytest.raises(avi.AviError, match="full-rank"):
        avi.gaussian_expectation_gradient_and_hessian(KEY, qm, 16, g, H, prob)
    prob.close()


# --- MvLocationScaleLowRank / LowRankGaussian (src/families/location_scale_low_rank.jl) ---------
def _lowrank_pair(avi, D, r, zero_factors=False):
    mu = (0.1 * np.arange(D) - 0.2)
    sd = 0.5 + 0.05 * np.arange(D)
    U = np.zeros((D, r)) if zero_factors else 0.3 * P.normal_matrix(91, 0, D, r)
    mu32, sd32, U32 = mu.astype(np.float32), sd.astype(np.float32), U.astype(np.float32)
    return (avi.LowRankGaussian(mu32, sd32, U32),
            F.LowRankGaussian(mu32.astype(np.float64), sd32.astype(np.float64), U32.astype(np.float64)))


def _lowrank_draws(D, r, M, step=0):
    return P.normal_matrix(KEY, step, D, M), P.normal_matrix(KEY, step, r, M, stream=P.STREAM_EPS_FACTORS)


@pytest.mark.parametrize("D,r,M", [(5, 1, 7), (33, 3, 70), (130, 32, 16)])
def test_lowrank_rand_and_gradient_match_oracle(avi, ctx, D, r, M):
    """rand (location_scale_low_rank.jl:79-86) and estimate_gradient! for RepGradELBO + ClosedFormEntropy over the
    low-rank family vs the oracle on the same two Philox streams (u_diag, u_fact)."""
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    prob, probo = avi.MvNormalDiag(ctx, mu_t, sg_t), Mo.NormalDiag(mu_t, sg_t)
    q, qo = _lowrank_pair(avi, D, r)
    obj = avi.Objective(KEY, avi.RepGradELBO(M), q, prob)
    assert obj.P == 2 * D + D * r
    u1, u2 = _lowrank_draws(D, r, M)
    Z, E = obj.rand(q)
    assert np.abs(E - u1).max() < 5e-6
    assert relerr(Z, qo.rand_from_eps(u1, u2)) < 2e-6
    v, g, e = obj.estimate_gradient(q.destructure())
    vo, go, eo = O.repgrad_lowrank_value_and_gradient(qo.destructure(), qo, probo, u1, u2)
    assert abs(v - vo) <= 2e-5 * max(1.0, abs(vo)) and abs(e - eo) <= 2e-5 * max(1.0, abs(eo))
    assert relerr(g, go) < 5e-5
    # blocks separately: location / scale_diag / scale_factors
    assert relerr(g[:D], go[:D]) < 5e-5 and relerr(g[D:2 * D], go[D:2 * D]) < 5e-5 and relerr(g[2 * D:], go[2 * D:]) < 1e-4
    obj.close(); prob.close()


def test_lowrank_logreg_gradient_matches_oracle(avi, ctx):
    n, d, r, M = 200, 20, 4, 32
    X, y = Mo.synth_glm_data(n, d, seed=8)
    D = d + 1
    prob, probo = avi.LogReg(ctx, X, y, gemm="tf32x3"), Mo.LogReg(X, y)
    q, qo = _lowrank_pair(avi, D, r)
    obj = avi.Objective(KEY, avi.RepGradELBO(M), q, prob)
    v, g, e = obj.estimate_gradient(q.destructure())
    u1, u2 = _lowrank_draws(D, r, M)
    vo, go, eo = O.repgrad_lowrank_value_and_gradient(qo.destructure(), qo, probo, u1, u2)
    assert abs(v - vo) <= 2e-5 * abs(vo) and relerr(g, go) < 5e-5
    obj.close(); prob.close()


@pytest.mark.parametrize("D,r,M", [(6, 3, 8), (37, 5, 33), (130, 8, 64)])
@pytest.mark.parametrize("ent", ["StickingTheLandingEntropy", "MonteCarloEntropy", "ClosedFormEntropyZeroGradient",
                                 "StickingTheLandingEntropyZeroGradient"])
def test_lowrank_logq_entropies_match_oracle(avi, ctx, D, r, M, ent):
    """RepGradELBO over the low-rank family with the entropy estimators that need log q(z) (entropy.jl:42-65) or drop the
    entropy gradient (:13-15): w = Sigma^-1 (z - mu) through the r x r capacitance inverse, vs the oracle."""
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    prob, probo = avi.MvNormalDiag(ctx, mu_t, sg_t), Mo.NormalDiag(mu_t, sg_t)
    q, qo = _lowrank_pair(avi, D, r)
    obj = avi.Objective(KEY, avi.RepGradELBO(M, getattr(avi, ent)()), q, prob)
    u1, u2 = _lowrank_draws(D, r, M)
    v, g, e = obj.estimate_gradient(q.destructure())
    vo, go, eo = O.repgrad_lowrank_value_and_gradient(qo.destructure(), qo, probo, u1, u2, entropy=ent)
    assert abs(v - vo) <= 5e-5 * max(1.0, abs(vo)) and abs(e - eo) <= 5e-5 * max(1.0, abs(eo)), (v, vo)
    assert relerr(g[:D], go[:D]) < 2e-4 and relerr(g[D:2 * D], go[D:2 * D]) < 2e-4 and relerr(g[2 * D:], go[2 * D:]) < 3e-4
    # a second call: the step counter moves both Philox streams
    v2, g2, _ = obj.estimate_gradient(q.destructure())
    u1b, u2b = _lowrank_draws(D, r, M, step=1)
    vo2, go2, _ = O.repgrad_lowrank_value_and_gradient(qo.destructure(), qo, probo, u1b, u2b, entropy=ent)
    assert abs(v2 - vo2) <= 5e-5 * max(1.0, abs(vo2)) and relerr(g2, go2) < 3e-4
    obj.close(); prob.close()


@pytest.mark.parametrize("D,r,M", [(6, 3, 8), (37, 5, 33), (130, 8, 64)])
def test_lowrank_scoregrad_matches_oracle(avi, ctx, D, r, M):
    """ScoreGradELBO (VarGrad, scoregradelbo.jl:87-117) over the low-rank family vs the oracle."""
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    prob, probo = avi.MvNormalDiag(ctx, mu_t, sg_t), Mo.NormalDiag(mu_t, sg_t)
    q, qo = _lowrank_pair(avi, D, r)
    obj = avi.Objective(KEY, avi.ScoreGradELBO(M), q, prob)
    u1, u2 = _lowrank_draws(D, r, M)
    v, g, e = obj.estimate_gradient(q.destructure())
    vo, go, eo = O.scoregrad_lowrank_value_and_gradient(qo.destructure(), qo, probo, u1, u2)
    assert abs(v - vo) <= 2e-4 * max(1.0, abs(vo)) and abs(e - eo) <= 5e-5 * max(1.0, abs(eo)), (v, vo, e, eo)
    assert relerr(g[:D], go[:D]) < 5e-4 and relerr(g[D:2 * D], go[D:2 * D]) < 5e-4 and relerr(g[2 * D:], go[2 * D:]) < 5e-4
    obj.close(); prob.close()


def test_lowrank_stl_logreg_and_fused_trajectory(avi, ctx):
    """Sticking-the-landing over the low-rank family on the logistic-regression target (gradient from the tensor-core
    kernels, fp32-grade) and 15 fused Adam + ClipScale iterations against the oracle loop."""
    n, d, r, M, T = 200, 20, 4, 32, 15
    X, y = Mo.synth_glm_data(n, d, seed=8)
    D = d + 1
    prob, probo = avi.LogReg(ctx, X, y, gemm="tf32x3"), Mo.LogReg(X, y)
    q, qo = _lowrank_pair(avi, D, r)
    ent = "StickingTheLandingEntropy"
    obj = avi.Objective(KEY, avi.RepGradELBO(M, avi.StickingTheLandingEntropy()), q, prob)
    v, g, e = obj.estimate_gradient(q.destructure())
    u1, u2 = _lowrank_draws(D, r, M)
    vo, go, eo = O.repgrad_lowrank_value_and_gradient(qo.destructure(), qo, probo, u1, u2, entropy=ent)
    assert abs(v - vo) <= 5e-5 * abs(vo) and relerr(g, go) < 2e-4
    obj.close()
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), entropy=avi.StickingTheLandingEntropy(), n_samples=M,
                                  operator=avi.ClipScale())
    qa, info, state = avi.optimize(KEY, alg, T, prob, q)
    rule, op, avg = Op.Adam(1e-2), Op.ClipScale(), Op.PolynomialAveraging()
    st = Op.sgd_init(qo, rule, avg)

    def grad_fn(params, t):
        a1, a2 = _lowrank_draws(D, r, M, step=t - 1)
        vv, gg, ee = O.repgrad_lowrank_value_and_gradient(params, qo, probo, a1, a2, entropy=ent)
        return vv, gg, dict(elbo=ee)
    elbos = [Op.sgd_step(st, qo, grad_fn, rule, op, avg)["elbo"] for _ in range(T)]
    lam, lam_avg, _ = state.params()
    assert np.allclose([i["elbo"] for i in info], elbos, rtol=5e-4, atol=5e-4)
    assert relerr(lam, st.params) < 5e-4 and relerr(lam_avg, st.avg_st[0]) < 5e-4
    state.close(); state.obj.close(); prob.close()


def test_lowrank_fused_trajectory_and_unsupported_combinations(avi, ctx):
    """docs/src/families.md:185-190: LowRankGaussian(mu, ones, zeros(d, 3)) with KLMinRepGradDescent, Adam and
    ClipScale: 20 fused iterations follow the fp64 oracle; what the low-rank family does not implement (the STL
    zero-gradient entropy / proximal algorithm, rank > 32, estimate_objective) is reported as unsupported."""
    D, r, M, T = 6, 3, 8, 20
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    prob, probo = avi.MvNormalDiag(ctx, mu_t, sg_t), Mo.NormalDiag(mu_t, sg_t)
    q, qo = _lowrank_pair(avi, D, r, zero_factors=True)
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=M, operator=avi.ClipScale())
    qa, info, state = avi.optimize(KEY, alg, T, prob, q)
    rule, op, avg = Op.Adam(1e-2), Op.ClipScale(), Op.PolynomialAveraging()
    st = Op.sgd_init(qo, rule, avg)

    def grad_fn(params, t):
        u1, u2 = _lowrank_draws(D, r, M, step=t - 1)
        v, g, e = O.repgrad_lowrank_value_and_gradient(params, qo, probo, u1, u2)
        return v, g, dict(elbo=e)
    elbos = [Op.sgd_step(st, qo, grad_fn, rule, op, avg)["elbo"] for _ in range(T)]
    lam, lam_avg, _ = state.params()
    assert np.allclose([i["elbo"] for i in info], elbos, rtol=2e-4, atol=2e-4)
    assert relerr(lam, st.params) < 1e-4 and relerr(lam_avg, st.avg_st[0]) < 1e-4
    assert isinstance(qa, avi.MvLocationScaleLowRank) and qa.scale_factors.shape == (D, r)
    state.close(); state.obj.close()
    with pytest.raises(avi.AviError, match="rank"):
        avi.Objective(KEY, avi.RepGradELBO(M), avi.LowRankGaussian(q.location, q.scale_diag, np.zeros((D, 33), np.float32)), prob)
    with pytest.raises(avi.AviError, match="MvLocationScale only"):      # the proximal operator has no low-rank form
        avi.optimize(KEY, avi.KLMinRepGradProxDescent(optimizer=avi.DoWG(), n_samples=M), 1, prob, q)
    prob.close()


@pytest.mark.parametrize("D,r", [(6, 3), (37, 5)])
def test_lowrank_estimate_objective_matches_oracle(avi, ctx, D, r):
    """estimate_objective (repgradelbo.jl:112-122, scoregradelbo.jl:58-65, common.jl:29-38) over the low-rank family,
    away from q = pi: closed-form entropy from the capacitance matrix, Monte-Carlo entropy from log q(z) through it."""
    mu_t, sg_t = np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D)
    prob, probo = avi.MvNormalDiag(ctx, mu_t, sg_t), Mo.NormalDiag(mu_t, sg_t)
    q, qo = _lowrank_pair(avi, D, r)
    n = 61
    u1, u2 = P.normal_matrix(KEY, 0, D, n), P.normal_matrix(KEY, 0, r, n, stream=P.STREAM_EPS_FACTORS)
    Z = qo.rand_from_eps(u1, u2)
    energy = float(np.mean(probo.logdensity_and_gradient_batch(Z)[0]))
    want_closed, want_mc = -(energy + qo.entropy()), -(energy - float(np.mean(qo.logpdf(Z))))
    assert abs(want_closed) > 1.0 and abs(want_mc) > 1.0
    for ent, want in (("ClosedFormEntropy", want_closed), ("ClosedFormEntropyZeroGradient", want_closed),
                      ("MonteCarloEntropy", want_mc), ("StickingTheLandingEntropy", want_mc)):
        got = avi.estimate_objective(KEY, avi.RepGradELBO(n, getattr(avi, ent)()), q, prob)
        assert abs(got - want) <= 3e-5 * abs(want), (ent, got, want)
    got = avi.estimate_objective(KEY, avi.ScoreGradELBO(n), q, prob)
    assert abs(got - want_mc) <= 3e-5 * abs(want_mc)
    alg = avi.KLMinRepGradDescent(n_samples=3)          # algorithm level: RepGradELBO + MonteCarloEntropy, caller's n_samples
    got = avi.estimate_objective(KEY, alg, q, prob, n_samples=n)
    assert abs(got - want_mc) <= 3e-5 * abs(want_mc)
    prob.close()
