"""The full-rank iteration of the optimiser loop (csrc/opt.cu: k_fr_update_t, csrc/family_fr.cu): its update kernel
keeps the transposed 3xTF32 split of L that the NEXT iteration's z = L eps + mu contraction reads, so the sampling stage
skips its own pass over the D x D matrix.  That is only valid while nothing else touched lambda or the split: these
tests interleave everything that may (the objective evaluated at another lambda, forward-only estimates, state
import, a second optimiser over the same objective buffers) and require BITWISE the trajectory of an undisturbed run
-- and the oracle's trajectory within the fp32 tolerances of tests/test_gpu_parity.py.

Reference behaviour being reproduced: `step` of src/algorithms/common.jl:69-120 with `rand` of
src/families/location_scale.jl:71-77 (scale * eps + location for a LowerTriangular scale)."""
import numpy as np
import pytest

from oracle import family as F, models as Mo, objectives as O, optim as Op, philox as P

pytestmark = pytest.mark.gpu

KEY = 0x38BEF07CF9CC549D


@pytest.fixture(scope="module")
def ctx(avi):
    c = avi.Context(0)
    yield c
    c.close()


def _setup(avi, ctx, d=36, n=320, gemm="tf32"):
    X, y = Mo.synth_glm_data(n, d, seed=7)
    prob = avi.LogReg(ctx, X, y, gemm=gemm)
    D = d + 1
    rng = np.random.default_rng(5)
    Lm = np.tril(0.03 * rng.standard_normal((D, D))).astype(np.float32)
    Lm[np.diag_indices(D)] = (0.3 + 0.01 * np.arange(D)).astype(np.float32)
    mu = (0.05 * rng.standard_normal(D)).astype(np.float32)
    return X, y, prob, avi.FullRankGaussian(mu, Lm), D


def _alg(avi, M=24, ent=None):
    return avi.KLMinRepGradDescent(optimizer=avi.Adam(5e-3), n_samples=M, entropy=ent, operator=avi.ClipScale())


@pytest.mark.parametrize("entropy", ["ClosedFormEntropy", "StickingTheLandingEntropy"])
def test_fullrank_trajectory_survives_interleaved_calls(avi, ctx, entropy):
    """8 iterations in one call == 3 + (other users of the objective's buffers) + 2 + (state export / import) + 3."""
    X, y, prob, q0, D = _setup(avi, ctx)
    ent = getattr(avi, entropy)()
    qa, ia, sa = avi.optimize(KEY, _alg(avi, ent=ent), 8, prob, q0)
    lam_a, avg_a, _ = sa.params()

    qb, ib, sb = avi.optimize(KEY, _alg(avi, ent=ent), 3, prob, q0)
    # (1) the same objective handle draws at a DIFFERENT lambda (rewrites the split of L; does not advance the step) and a
    # forward-only estimate runs over the same buffers
    other = avi.FullRankGaussian(np.ones(D, np.float32), (0.7 * np.eye(D)).astype(np.float32))
    sb.obj.rand(other)
    sb.obj.estimate_objective(KEY + 9, other, 64)
    _, ib2, _ = avi.optimize(KEY, _alg(avi, ent=ent), 2, prob, q0, state=sb)
    # (2) a second optimiser state over a second objective on the same target, run in between
    qx, _, sx = avi.optimize(KEY + 3, _alg(avi, ent=ent), 2, prob, other)
    # (3) export / import of the optimiser state (lambda is rewritten from the host)
    blob = sb.export_bytes()
    sb.import_bytes(blob)
    _, ib3, _ = avi.optimize(KEY, _alg(avi, ent=ent), 3, prob, q0, state=sb)
    lam_b, avg_b, _ = sb.params()
    assert np.array_equal(lam_a, lam_b) and np.array_equal(avg_a, avg_b)
    assert [i["elbo"] for i in ia] == [i["elbo"] for i in ib + ib2 + ib3]
    for s in (sa, sb, sx):
        s.close(); s.obj.close()
    prob.close()


def test_fullrank_trajectory_matches_oracle_exact_fp32(avi, ctx):
    """The same loop against the fp64 oracle on identical eps, exact-fp32 GLM contraction (the family's own contractions
    run 3xTF32): lambda after 6 Adam + ClipScale steps within 2e-4 relative, every ELBO within 5e-5."""
    X, y, prob, q0, D = _setup(avi, ctx, gemm="fp32")
    M, T = 24, 6
    q, info, st = avi.optimize(KEY, _alg(avi, M), T, prob, q0)
    lam, avg, _ = st.params()
    qo = F.FullRankGaussian(q0.location.astype(np.float64), q0.scale.astype(np.float64))
    po = Mo.LogReg(X, y)
    rule, op, avgr = Op.Adam(5e-3), Op.ClipScale(), Op.PolynomialAveraging()
    so = Op.sgd_init(qo, rule, avgr)

    def grad_fn(params, t):
        v, g, e = O.repgrad_value_and_gradient(params, qo, po, P.normal_matrix(KEY, t - 1, D, M), "ClosedFormEntropy")
        return v, g, dict(elbo=e)
    for t in range(T):
        elbo = Op.sgd_step(so, qo, grad_fn, rule, op, avgr)["elbo"]
        assert abs(info[t]["elbo"] - elbo) <= 5e-5 * abs(elbo), (t, info[t]["elbo"], elbo)
    assert np.linalg.norm(lam - so.params) <= 2e-4 * np.linalg.norm(so.params)
    assert np.linalg.norm(avg - so.avg_st[0]) <= 2e-4 * np.linalg.norm(so.avg_st[0])
    st.close(); st.obj.close(); prob.close()


def test_fullrank_halted_step_keeps_split_consistent(avi, ctx):
    """A non-finite value slot leaves lambda (and therefore the maintained split of L) untouched: after the divergence is
    reported, the state still holds the last good iterate and a fresh run from it reproduces itself bitwise."""
    D = 6
    prob = avi.MvNormalDiag(ctx, np.zeros(D), np.ones(D))
    q0 = avi.FullRankGaussian(np.zeros(D, np.float32), np.eye(D, dtype=np.float32))
    alg = avi.KLMinRepGradDescent(optimizer=avi.Descent(1e38), n_samples=4, operator=avi.ClipScale())
    with pytest.raises(Exception):
        avi.optimize(KEY, alg, 6, prob, q0)
    # a sane run afterwards on the same context and target is unaffected
    alg2 = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=4, operator=avi.ClipScale())
    qa, ia, sa = avi.optimize(KEY, alg2, 5, prob, q0)
    qb, ib, sb = avi.optimize(KEY, alg2, 5, prob, q0)
    assert np.array_equal(qa.scale, qb.scale) and [i["elbo"] for i in ia] == [i["elbo"] for i in ib]
    for s in (sa, sb):
        s.close(); s.obj.close()
    prob.close()
