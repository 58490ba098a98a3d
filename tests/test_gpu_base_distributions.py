"""MvLocationScale(location, scale, dist) with a NON-Gaussian base distribution on the device (SURVEY.md 8f rank 4;
reference: src/families/location_scale.jl:15-19 constructor, :52-57 entropy, :59-63 logpdf, :71-87 rand; the bases the
family documentation runs, docs/src/families.md:72-101: TDist(nu) and Laplace(0, 1)).

Both sides consume the same Philox words: Laplace by inversion of the CDF, Student-t by Bailey's polar transform
(oracle/philox.py: laplace_matrix / student_t_matrix; csrc/base_dist.cuh).  Tolerances: draws 5e-6 relative to
max(1, |u|) (fp32 transcendental rounding); value / ELBO 3e-5, gradient 5e-5 relative (fp32 SIMT arithmetic vs the fp64
oracle; heavier tails than the Gaussian cases of tests/test_gpu_parity.py), ScoreGrad 3e-4 as there."""
import numpy as np
import pytest

from oracle import family as F, models as Mo, objectives as O, optim as Op, philox as P

pytestmark = pytest.mark.gpu

KEY = 0x38BEF07CF9CC549D
BASES = [("laplace", None), ("tdist", 5.0), ("tdist", 2.5)]
ENTROPIES = ["ClosedFormEntropy", "MonteCarloEntropy", "StickingTheLandingEntropy", "ClosedFormEntropyZeroGradient",
             "StickingTheLandingEntropyZeroGradient"]


@pytest.fixture(scope="module")
def ctx(avi):
    c = avi.Context(0)
    yield c
    c.close()


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def draws(base, nu, step, D, M, key=KEY):
    return P.laplace_matrix(key, step, D, M) if base == "laplace" else P.student_t_matrix(key, step, D, M, nu)


def make_q(avi, kind, D, base, nu):
    dist_d = avi.Laplace() if base == "laplace" else avi.TDist(nu)
    dist_o = F.LaplaceDist() if base == "laplace" else F.TDistBase(nu)
    mu = (0.1 * np.cos(np.arange(D)) - 0.05).astype(np.float32)
    if kind == "meanfield":
        s = (0.3 + 0.02 * (np.arange(D) % 7)).astype(np.float32)
    else:
        s = (np.tril(0.02 * np.sin(np.arange(D * D)).reshape(D, D), -1) + np.diag(0.3 + 0.02 * (np.arange(D) % 7))).astype(np.float32)
    return avi.MvLocationScale(mu, s, dist_d), F.MvLocationScale(mu.astype(np.float64), s.astype(np.float64), dist_o)


def targets(avi, ctx, D):
    m, s = 0.5 + 0.2 * np.arange(D), 0.6 + 0.05 * np.arange(D)
    return avi.MvNormalDiag(ctx, m, s), Mo.NormalDiag(m.astype(np.float32).astype(np.float64), s.astype(np.float32).astype(np.float64))


@pytest.mark.parametrize("base,nu", BASES)
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
@pytest.mark.parametrize("D,M", [(5, 10), (37, 33), (130, 7)])
def test_rand_matches_oracle(avi, ctx, base, nu, kind, D, M):
    """rand(rng, q, M): z = scale * u + location with u iid from the base (location_scale.jl:71-87)."""
    prob, _ = targets(avi, ctx, D)
    q, qo = make_q(avi, kind, D, base, nu)
    obj = avi.Objective(KEY, avi.RepGradELBO(M), q, prob)
    obj.seed(KEY, 3)
    Z, U = obj.rand(q)
    u = draws(base, nu, 3, D, M)
    assert np.max(np.abs(U - u) / np.maximum(1.0, np.abs(u))) < 5e-6
    zo = qo.rand_from_eps(u)
    assert np.max(np.abs(Z - zo) / np.maximum(1.0, np.abs(zo))) < 2e-5
    obj.close(); prob.close()


@pytest.mark.parametrize("base,nu", BASES)
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
@pytest.mark.parametrize("entropy", ENTROPIES)
def test_repgrad_matches_oracle(avi, ctx, base, nu, kind, entropy):
    """estimate_gradient! of RepGradELBO for every entropy estimator (entropy.jl:11-90): energy terms with eps -> u,
    sticking-the-landing with eps -> -score(u), closed-form entropy D * entropy(dist) + logdet(scale)."""
    D, M = 33, 70
    prob, probo = targets(avi, ctx, D)
    q, qo = make_q(avi, kind, D, base, nu)
    obj = avi.Objective(KEY, avi.RepGradELBO(M, getattr(avi, entropy)()), q, prob)
    for step in range(2):
        v, g, e = obj.estimate_gradient(q.destructure())
        vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, draws(base, nu, step, D, M), entropy)
        assert abs(v - vo) <= 3e-5 * max(1, abs(vo)) and abs(e - eo) <= 3e-5 * max(1, abs(eo)), (step, v, vo)
        assert relerr(g, go) < 5e-5, (step, relerr(g, go))
    obj.close(); prob.close()


@pytest.mark.parametrize("base,nu", BASES)
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
def test_scoregrad_matches_oracle(avi, ctx, base, nu, kind):
    """ScoreGradELBO (VarGrad, scoregradelbo.jl:87-117): grad log q through the base's score."""
    D, M = 12, 90
    prob, probo = targets(avi, ctx, D)
    q, qo = make_q(avi, kind, D, base, nu)
    obj = avi.Objective(KEY, avi.ScoreGradELBO(M), q, prob)
    for step in range(2):
        v, g, e = obj.estimate_gradient(q.destructure())
        vo, go, eo = O.scoregrad_value_and_gradient(qo.destructure(), qo, probo, draws(base, nu, step, D, M))
        assert abs(v - vo) <= 3e-4 * max(1, abs(vo)) and abs(e - eo) <= 3e-5 * max(1, abs(eo)), (step, v, vo, e, eo)
        assert relerr(g, go) < 3e-4, (step, relerr(g, go))
    obj.close(); prob.close()


@pytest.mark.parametrize("base,nu", BASES)
@pytest.mark.parametrize("kind", ["meanfield", "fullrank"])
def test_estimate_objective_matches_oracle(avi, ctx, base, nu, kind):
    """estimate_objective (repgradelbo.jl:112-122, scoregradelbo.jl:58-65): logpdf(q, z) = sum logpdf(dist, u) - logdet."""
    D, n = 9, 57
    prob, probo = targets(avi, ctx, D)
    q, qo = make_q(avi, kind, D, base, nu)
    u = draws(base, nu, 0, D, n)
    for ent in ENTROPIES:
        got = avi.estimate_objective(KEY, avi.RepGradELBO(n, getattr(avi, ent)()), q, prob)
        want = O.repgrad_estimate_objective(qo, probo, u, ent)
        assert abs(got - want) <= 3e-5 * abs(want), (ent, got, want)
    got = avi.estimate_objective(KEY, avi.ScoreGradELBO(n), q, prob)
    want = O.scoregrad_estimate_objective(qo, probo, u)
    assert abs(got - want) <= 3e-5 * abs(want)
    prob.close()


@pytest.mark.parametrize("base,nu,kind", [("laplace", None, "meanfield"), ("tdist", 5.0, "fullrank"),
                                          ("tdist", 5.0, "meanfield"), ("laplace", None, "fullrank")])
def test_optimiser_trajectory_matches_oracle(avi, ctx, base, nu, kind):
    """`step` (common.jl:69-120) with Adam + ClipScale + PolynomialAveraging on a logistic-regression target (exact-fp32
    contraction): 8 iterations follow the fp64 oracle driven by the same base draws; same seed => same run, bitwise."""
    d, n, M, T = 20, 200, 16, 8
    X, y = Mo.synth_glm_data(n, d, seed=8)
    prob, probo = avi.LogReg(ctx, X, y, gemm="fp32"), Mo.LogReg(X, y)
    D = d + 1
    q, qo = make_q(avi, kind, D, base, nu)
    ent = "StickingTheLandingEntropy" if kind == "fullrank" else "ClosedFormEntropy"
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(5e-3), entropy=getattr(avi, ent)(), n_samples=M, operator=avi.ClipScale())
    qa, info, st = avi.optimize(KEY, alg, T, prob, q)
    qb, info_b, st_b = avi.optimize(KEY, alg, T, prob, q)
    assert np.array_equal(qa.destructure(), qb.destructure()) and [i["elbo"] for i in info] == [i["elbo"] for i in info_b]
    assert repr(qa.dist) == repr(q.dist)
    rule, op, avgr = Op.Adam(5e-3), Op.ClipScale(), Op.PolynomialAveraging()
    so = Op.sgd_init(qo, rule, avgr)

    def grad_fn(params, t):
        v, g, e = O.repgrad_value_and_gradient(params, qo, probo, draws(base, nu, t - 1, D, M), ent)
        return v, g, dict(elbo=e)
    for t in range(T):
        elbo = Op.sgd_step(so, qo, grad_fn, rule, op, avgr)["elbo"]
        assert abs(info[t]["elbo"] - elbo) <= 1e-4 * abs(elbo), (t, info[t]["elbo"], elbo)
    lam, avg, _ = st.params()
    assert relerr(lam, so.params) < 3e-4 and relerr(avg, so.avg_st[0]) < 3e-4
    for s in (st, st_b):
        s.close(); s.obj.close()
    prob.close()


def test_tensor_core_target_and_invalid_combinations(avi, ctx):
    """A non-Gaussian base over the TF32 tensor-core target runs the multi-kernel path (the single-launch iteration is a
    Normal(0, 1) kernel) and matches the oracle within the TF32 tolerances; the low-rank family rejects other bases."""
    n, d, M = 700, 96, 64
    X, y = Mo.synth_glm_data(n, d, seed=6)
    prob, probo = avi.LogReg(ctx, X, y, gemm="tf32"), Mo.LogReg(X, y)
    D = d + 1
    q, qo = make_q(avi, "meanfield", D, "tdist", 4.0)
    for ent in ("ClosedFormEntropy", "StickingTheLandingEntropy"):
        obj = avi.Objective(KEY, avi.RepGradELBO(M, getattr(avi, ent)()), q, prob)
        v, g, e = obj.estimate_gradient(q.destructure())
        vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, draws("tdist", 4.0, 0, D, M), ent)
        assert abs(v - vo) <= 5e-4 * abs(vo) and relerr(g, go) < 2e-3
        obj.close()
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=M, operator=avi.ClipScale())
    l0 = ctx.launch_count()
    _, info, st = avi.optimize(KEY, alg, 4, prob, q)
    assert (ctx.launch_count() - l0) / 4 > 2          # not the single-launch iteration
    assert np.isfinite(info[-1]["elbo"])
    st.close(); st.obj.close()
    from advancedvi_jl_b200 import _lib as L
    qlr = avi.LowRankGaussian(np.zeros(D, np.float32), np.ones(D, np.float32), 0.1 * np.ones((D, 2), np.float32))
    olr = avi.Objective(KEY, avi.RepGradELBO(8), qlr, prob)
    assert L.lib.avi_obj_set_base(olr.h, 1, 0.0) == 3          # AVI_ERR_UNSUPPORTED
    assert L.lib.avi_obj_set_base(olr.h, 2, -1.0) != 0         # invalid nu
    assert L.lib.avi_obj_set_base(olr.h, 0, 0.0) == 0          # Normal(0, 1) is fine everywhere
    olr.close()
    with pytest.raises(ValueError):
        avi.TDist(0.0)
    prob.close()
