"""Edge cases of the boundary: smallest shapes, ragged sizes, invalid arguments, resource lifecycle."""
import numpy as np
import pytest

from oracle import family as F, models as Mo, objectives as O, philox as P

pytestmark = pytest.mark.gpu
KEY = 12345


@pytest.fixture(scope="module")
def ctx(avi):
    c = avi.Context(0)
    yield c
    c.close()


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("gemm", ["fp32", "tf32", "tf32x3"])
@pytest.mark.parametrize("n,d,M", [(1, 1, 1), (2, 3, 1), (33, 1, 5), (31, 33, 129), (257, 5, 2), (129, 127, 257)])
def test_glm_ragged_shapes(avi, ctx, gemm, n, d, M):
    """n, d, M that are not multiples of any tile size (TMA out-of-bounds zero fill, masked epilogues)."""
    X, y = Mo.synth_glm_data(max(n, 2), d, seed=7)
    X, y = X[:n], y[:n]
    prob, probo = avi.LogReg(ctx, X, y, n_data=5 * n, gemm=gemm), Mo.LogReg(X, y, n_data=5 * n)
    Z = (0.2 * P.normal_matrix(3, 0, d + 1, M)).astype(np.float32)
    lp, G = prob.logdensity_and_gradient(Z)
    lpo, Go = probo.logdensity_and_gradient_batch(Z.astype(np.float64))
    tol_l, tol_g = (5e-4, 3e-3) if gemm == "tf32" else (5e-6, 5e-5)
    assert np.abs(lp - lpo).max() <= tol_l * max(np.abs(lpo).max(), 1.0)
    assert relerr(G, Go) < tol_g
    # and through the objective (mean-field fused path + full-rank path)
    D = d + 1
    q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.3, np.float32))
    qo = F.MeanFieldGaussian(np.zeros(D), np.full(D, 0.3, np.float32).astype(np.float64))
    obj = avi.Objective(KEY, avi.RepGradELBO(M), q, prob)
    v, g, e = obj.estimate_gradient(q.destructure())
    vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, P.normal_matrix(KEY, 0, D, M), "ClosedFormEntropy")
    assert abs(v - vo) <= (5e-4 if gemm == "tf32" else 2e-5) * max(abs(vo), 1.0)
    assert relerr(g, go) < (3e-3 if gemm == "tf32" else 1e-4)
    obj.close(); prob.close()


def test_dimension_one_target(avi, ctx):
    """D = 1 (the shape of test/models/subsamplednormals.jl): every padded lane is masked."""
    prob, probo = avi.MvNormalDiag(ctx, [2.0], [0.5]), Mo.NormalDiag([2.0], [0.5])
    for kind in ("meanfield", "fullrank"):
        if kind == "meanfield":
            q, qo = avi.MeanFieldGaussian(np.zeros(1, np.float32), np.ones(1, np.float32)), F.MeanFieldGaussian(np.zeros(1), np.ones(1))
        else:
            q, qo = avi.FullRankGaussian(np.zeros(1, np.float32), np.ones((1, 1), np.float32)), F.FullRankGaussian(np.zeros(1), np.ones((1, 1)))
        for spec, name in ((avi.RepGradELBO(7, avi.StickingTheLandingEntropy()), "stl"), (avi.ScoreGradELBO(7), "score")):
            obj = avi.Objective(KEY, spec, q, prob)
            v, g, e = obj.estimate_gradient(q.destructure())
            eps = P.normal_matrix(KEY, 0, 1, 7)
            if name == "stl":
                vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, eps, "StickingTheLandingEntropy")
            else:
                vo, go, eo = O.scoregrad_value_and_gradient(qo.destructure(), qo, probo, eps)
            assert abs(v - vo) <= 1e-4 * max(1, abs(vo)) and relerr(g, go) < 2e-4
            obj.close()
    prob.close()


def test_invalid_arguments_are_reported(avi, ctx):
    """Error behaviour of the boundary: status code + message, never a crash or a silent fallback."""
    with pytest.raises(avi.AviError, match="sigma must be positive"):
        avi.MvNormalDiag(ctx, [0.0, 0.0], [1.0, -1.0])
    with pytest.raises(ValueError):
        avi.LogReg(ctx, np.zeros((4, 2), np.float32), np.zeros(3, np.float32))
    X, y = Mo.synth_glm_data(16, 3, seed=1)
    prob = avi.LogReg(ctx, X, y, gemm="tf32")
    with pytest.raises(avi.AviError, match="out of range"):
        prob.subsample([0, 99])
    with pytest.raises(avi.AviError):
        prob.subsample([])
    q = avi.MeanFieldGaussian(np.zeros(4, np.float32), np.ones(4, np.float32))
    with pytest.raises(avi.AviError, match="n_samples"):
        avi.Objective(KEY, avi.RepGradELBO(0), q, prob)
    with pytest.raises(ValueError, match="dimension"):
        avi.Objective(KEY, avi.RepGradELBO(2), avi.MeanFieldGaussian(np.zeros(3, np.float32), np.ones(3, np.float32)), prob)
    obj = avi.Objective(KEY, avi.RepGradELBO(2), q, prob)
    with pytest.raises(avi.AviError, match="num_params"):
        obj.estimate_gradient(np.zeros(5, np.float32))
    # ProximalLocationScaleEntropy only supports Descent / DoG / DoWG (proximal_location_scale_entropy.jl:29-44)
    alg = avi.KLMinRepGradProxDescent(optimizer=avi.Adam(1e-3), n_samples=2)
    with pytest.raises(avi.AviError, match="ProximalLocationScaleEntropy"):
        avi.optimize(KEY, alg, 1, prob, q)
    obj.close(); prob.close()


def test_many_handles_and_reuse(avi, ctx):
    """Create / destroy many handles, swap the target of a live objective (set_objective_state_problem)."""
    X, y = Mo.synth_glm_data(64, 4, seed=2)
    q = avi.MeanFieldGaussian(np.zeros(5, np.float32), np.full(5, 0.4, np.float32))
    probs = [avi.LogReg(ctx, X[i::2], y[i::2], gemm="tf32") for i in range(2)]
    obj = avi.Objective(KEY, avi.RepGradELBO(9), q, probs[0])
    res = []
    for k in range(6):
        obj.set_problem(probs[k % 2])
        obj.seed(KEY, 0)
        res.append(obj.estimate_gradient(q.destructure()))
    assert res[0][0] == res[2][0] == res[4][0] and res[1][0] == res[3][0] == res[5][0]
    assert np.array_equal(res[0][1], res[4][1]) and not np.array_equal(res[0][1], res[1][1])
    for _ in range(20):
        p = avi.MvNormalDiag(ctx, np.zeros(5), np.ones(5))
        o = avi.Objective(KEY, avi.ScoreGradELBO(3), q, p)
        o.estimate_gradient(q.destructure())
        o.close(); p.close()
    obj.close()
    for p in probs:
        p.close()


def test_steps_begin_enqueue_end_matches_blocking_call(avi, ctx):
    """avi_opt_steps_begin / _enqueue / _end (no host round trip between iterations) == avi_opt_steps, bitwise;
    misuse is reported as a state error."""
    from advancedvi_jl_b200 import api as A
    X, y = Mo.synth_glm_data(200, 12, seed=4)
    prob = avi.LogReg(ctx, X, y, gemm="tf32")
    D = 13
    q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.5, np.float32))
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=16, operator=avi.ClipScale())
    _, info, st_ref = avi.optimize(KEY, alg, 10, prob, q)
    lam_ref, avg_ref, _ = st_ref.params()

    obj = avi.Objective(KEY, alg.objective, q, prob)
    st = A._OptState(alg, obj, q)
    with pytest.raises(avi.AviError, match="without avi_opt_steps_begin"):
        st.steps_enqueue(1)
    st.steps_begin(10)
    with pytest.raises(avi.AviError, match="already open"):
        st.steps_begin(10)
    st.steps_enqueue(3)
    st.steps_enqueue(7)
    with pytest.raises(avi.AviError, match="reserved"):
        st.steps_enqueue(1)
    vals, elbos, done = st.steps_end()
    assert done == 10 and st.iteration == 10
    assert np.array_equal(elbos[:10], np.array([i["elbo"] for i in info], np.float32))
    lam, avg, _ = st.params()
    assert np.array_equal(lam, lam_ref) and np.array_equal(avg, avg_ref)
    st.close(); obj.close(); st_ref.close(); st_ref.obj.close(); prob.close()
