"""Committed golden vectors (tests/golden/elbo_golden.npz, made by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them exactly; the CUDA path must match them within the stated tolerances."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as MG   # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "elbo_golden.npz"))


@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_oracle_reproduces_golden(name):
    cur = MG.build(name)
    for k, v in cur.items():
        ref = GOLD[f"{name}/{k}"]
        assert np.array_equal(np.asarray(v), ref), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("gemm,tol_v,tol_g", [("fp32", 2e-5, 5e-5), ("tf32x3", 2e-5, 5e-5), ("tf32", 5e-4, 3e-3)])
@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_gpu_matches_golden(avi, name, gemm, tol_v, tol_g):
    """Same lambda, same Philox key => same eps on the device; value slot / gradient / elbo vs the committed fp64
    vectors.  Tolerances: fp32-grade paths 2e-5 (value) / 5e-5 (gradient norm), single-pass TF32 5e-4 / 3e-3."""
    n, d, M, key, dseed, fam, objective, entropy = MG.CASES[name]
    X, y, lam = GOLD[f"{name}/X"], GOLD[f"{name}/y"], GOLD[f"{name}/lam"]
    D = d + 1
    ctx = avi.Context(0)
    prob = avi.LogReg(ctx, X, y, gemm=gemm)
    if fam == "mf":
        q = avi.MeanFieldGaussian(lam[:D].astype(np.float32), lam[D:].astype(np.float32))
    else:
        q = avi.FullRankGaussian(lam[:D].astype(np.float32), lam[D:].reshape(D, D, order="F").astype(np.float32))
    ent = {None: None, "ClosedFormEntropy": avi.ClosedFormEntropy(), "MonteCarloEntropy": avi.MonteCarloEntropy(),
           "StickingTheLandingEntropy": avi.StickingTheLandingEntropy()}[entropy]
    spec = avi.RepGradELBO(M, ent) if objective == "rep" else avi.ScoreGradELBO(M)
    obj = avi.Objective(key, spec, q, prob)
    Z, eps = obj.rand(q)
    assert np.abs(eps - GOLD[f"{name}/eps"]).max() < 5e-6          # identical Philox stream, fp32 Box-Muller
    v, g, e = obj.estimate_gradient(q.destructure())
    gv, gg, ge = float(GOLD[f"{name}/value"]), GOLD[f"{name}/grad"], float(GOLD[f"{name}/elbo"])
    scale_v = max(abs(gv), abs(ge), 1.0)
    tv = 20 * tol_v if objective == "score" else tol_v             # VarGrad value: a variance of O(|elbo|)-sized terms
    assert abs(v - gv) <= tv * scale_v and abs(e - ge) <= tol_v * scale_v, (v, gv, e, ge)
    assert np.linalg.norm(g - gg) <= (10 * tol_g if objective == "score" else tol_g) * np.linalg.norm(gg)
    obj.close(); prob.close(); ctx.close()
