"""Generates tests/golden/elbo_golden.npz: committed input/output vectors of the ELBO-gradient path.

The reference (Julia) cannot run in this image, so these vectors are produced by the CPU oracle (oracle/, fp64),
which is itself pinned against the reference's known-answer tests (tests/test_oracle_known_answers.py).  They are
regression pins: the oracle must keep reproducing them bit for bit (fp64, same numpy arithmetic), and the CUDA path
must reproduce them within the fp32 / TF32 tolerances written in tests/test_golden_vectors.py.  An independent check of
the stored OUTPUTS exists too: tests/test_oracle_independent_ad.py recomputes them from the stored inputs with PyTorch
autograd over a restatement of the reference's forward closures that does not use oracle/.

    python tests/golden/make_golden.py        (from the repository root; overwrites the fixture)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import family as F, models as Mo, objectives as O, philox as P   # noqa: E402

CASES = {
    # name: (n, d, M, key, data seed, family, objective, entropy)
    "logreg_mf_rep_cfe": (64, 7, 16, 21, 5, "mf", "rep", "ClosedFormEntropy"),
    "logreg_mf_rep_stl": (64, 7, 16, 22, 5, "mf", "rep", "StickingTheLandingEntropy"),
    "logreg_mf_score": (64, 7, 32, 23, 5, "mf", "score", None),
    "logreg_fr_rep_cfe": (48, 5, 16, 24, 6, "fr", "rep", "ClosedFormEntropy"),
    "logreg_fr_rep_mc": (48, 5, 16, 25, 6, "fr", "rep", "MonteCarloEntropy"),
}


def build(name):
    n, d, M, key, dseed, fam, objective, entropy = CASES[name]
    X, y = Mo.synth_glm_data(n, d, seed=dseed)
    D = d + 1
    mu = 0.1 * P.normal_matrix(key + 100, 0, D, 1)[:, 0]
    if fam == "mf":
        q = F.MeanFieldGaussian(mu, np.full(D, 0.4) + 0.05 * np.arange(D))
    else:
        L = np.tril(0.05 * P.normal_matrix(key + 200, 0, D, D)) + 0.5 * np.eye(D)
        q = F.FullRankGaussian(mu, L)
    eps = P.normal_matrix(key, 0, D, M)
    prob = Mo.LogReg(X, y)
    lam = q.destructure()
    if objective == "rep":
        v, g, e = O.repgrad_value_and_gradient(lam, q, prob, eps, entropy)
    else:
        v, g, e = O.scoregrad_value_and_gradient(lam, q, prob, eps)
    return dict(X=X, y=y, lam=lam, eps=eps, value=np.float64(v), grad=np.asarray(g, np.float64), elbo=np.float64(e))


def main():
    out = {}
    for name in CASES:
        for k, v in build(name).items():
            out[f"{name}/{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "elbo_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
