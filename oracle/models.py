"""Target log-densities (the LogDensityProblems plugin side of the path).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Interface mirrors LogDensityProblems as the reference uses it
(src/algorithms/repgradelbo.jl:50, :85; src/mixedad_logdensity.jl:13-28):
``dimension()``, ``capability`` (0 or 1), ``logdensity(z)``,
``logdensity_and_gradient(z)`` for ONE sample z in R^D, plus ``subsample(idx)``
(src/AdvancedVI.jl:303-313).  ``*_batch`` variants evaluate all columns of a (D, M)
matrix at once (same arithmetic; used as the "best effort" CPU baseline).

Models:
  * NormalDiag / NormalDense   -- test/models/normal.jl:8-11, :36-75
  * SubsampledNormals          -- test/models/subsamplednormals.jl:17-20, :45-48
  * LogReg(variant="subsampling") -- docs/src/tutorials/subsampling.md:26-38, :99-102
  * LogReg(variant="basic")    -- README.md:47-58 wrapped by the exp-bijector
                                  TransformedLogDensityProblem README.md:91-106
  * GaussGLM                   -- builder-defined (BASELINE.json config 4):
                                  LogReg(subsampling) with y_i ~ N(x_i' beta, 1)
Distributions.jl formulas are restated from their definitions;
the reference has no golden vectors for the regression models ("parity unpinned").
"""

from __future__ import annotations

import numpy as np

LOG2PI = float(np.log(2.0 * np.pi))
LOG3 = float(np.log(3.0))


def log1pexp(x):
    """log(1 + exp(x)), stable (LogExpFunctions.log1pexp as used by BernoulliLogit)."""
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


class _Target:
    capability = 1

    def dimension(self):
        raise NotImplementedError

    def logdensity(self, z):
        return self.logdensity_and_gradient(z)[0]

    def logdensity_and_gradient(self, z):
        l, g = self.logdensity_and_gradient_batch(np.asarray(z)[:, None])
        return l[0], g[:, 0]

    def logdensity_batch(self, Z):
        return self.logdensity_and_gradient_batch(Z)[0]

    def subsample(self, idx):            # AdvancedVI.subsample default = identity
        return self                      # (src/AdvancedVI.jl:313)


class NormalDiag(_Target):
    """logpdf(MvNormal(mu, Diagonal(sigma.^2)), z)  (test/models/normal.jl:8-11, :56-75)."""

    def __init__(self, mu, sigma, capability=1):
        self.mu = np.asarray(mu, dtype=np.float64)
        self.sigma = np.asarray(sigma, dtype=np.float64)
        self.capability = capability

    def dimension(self):
        return len(self.mu)

    def logdensity_and_gradient_batch(self, Z):
        r = (Z - self.mu[:, None]) / self.sigma[:, None]
        logp = (-0.5 * len(self.mu) * LOG2PI - np.sum(np.log(self.sigma))
                - 0.5 * np.sum(r * r, axis=0))
        G = -r / self.sigma[:, None]
        return logp, G


class NormalDense(_Target):
    """logpdf(MvNormal(mu, Sigma), z) with Sigma = L L' (test/models/normal.jl:36-54)."""

    def __init__(self, mu, L, capability=1):
        self.mu = np.asarray(mu, dtype=np.float64)
        self.L = np.asarray(L, dtype=np.float64)
        self.capability = capability

    def dimension(self):
        return len(self.mu)

    def logdensity_and_gradient_batch(self, Z):
        from scipy.linalg import solve_triangular
        r = solve_triangular(self.L, Z - self.mu[:, None], lower=True)
        logp = (-0.5 * len(self.mu) * LOG2PI - np.sum(np.log(np.diag(self.L)))
                - 0.5 * np.sum(r * r, axis=0))
        G = -solve_triangular(self.L.T, r, lower=False)
        return logp, G


class SubsampledNormals(_Target):
    """1-D product of unit-variance normals with likelihood adjustment
    (test/models/subsamplednormals.jl:17-20); subsample -> :45-48."""

    def __init__(self, mus, likeadj=1.0, capability=1, n_data=None):
        self.mus = np.asarray(mus, dtype=np.float64)
        self.likeadj = float(likeadj)
        self.capability = capability

    def dimension(self):
        return 1

    def logdensity_and_gradient_batch(self, Z):
        x = Z[0][None, :]                                   # only(x)
        d = x - self.mus[:, None]
        logp = self.likeadj * np.sum(-0.5 * LOG2PI - 0.5 * d * d, axis=0)
        G = (self.likeadj * np.sum(-d, axis=0))[None, :]
        return logp, G

    def subsample(self, idx):
        idx = np.asarray(idx)
        return SubsampledNormals(self.mus[idx], len(self.mus) / len(idx), self.capability)


class LogReg(_Target):
    """Hierarchical logistic regression, theta = [beta; eta], sigma = exp(eta).

    variant="subsampling" (docs/src/tutorials/subsampling.md:26-38):
        n_data/n * sum_i logpdf(BernoulliLogit(l_i), y_i) + logpdf(MvNormal(0, sigma), beta)
        + logpdf(Normal(0, 3), sigma)
    variant="basic" (README.md:47-58 under the exp bijector, README.md:91-106):
        sum_i ... + logpdf(MvNormal(0, sigma), beta) + logpdf(LogNormal(0, 3), sigma) + eta
    MvNormal(Zeros(d), sigma::Real) is the isotropic constructor with sigma the standard
    deviation (Appendix B).  `capability` defaults to 1: the native CUDA target supplies
    its own gradient, the semantics of MixedADLogDensityProblem
    (src/mixedad_logdensity.jl:23-34).
    """

    family = "bernoulli_logit"

    def __init__(self, X, y, n_data=None, variant="subsampling", capability=1):
        self.X = np.asarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.n_data = int(self.X.shape[0] if n_data is None else n_data)
        assert variant in ("subsampling", "basic")
        self.variant = variant
        self.capability = capability

    def dimension(self):
        return self.X.shape[1] + 1

    # likelihood pieces, overridden by GaussGLM
    def _loglik_and_resid(self, logits):
        y = self.y[:, None]
        ll = y * logits - log1pexp(logits)       # logpdf(BernoulliLogit(l), y)
        return np.sum(ll, axis=0), y - sigmoid(logits)

    def likeadj(self):
        n = self.X.shape[0]
        return self.n_data / n if self.variant == "subsampling" else 1.0

    def logdensity_and_gradient_batch(self, Z):
        d = self.X.shape[1]
        B, eta = Z[:d], Z[d]
        sigma2 = np.exp(2.0 * eta)
        bn2 = np.sum(B * B, axis=0)
        w = self.likeadj()
        logits = self.X @ B                                   # the GEMM (M GEMVs in the reference)
        ll, resid = self._loglik_and_resid(logits)
        logprior_beta = -0.5 * d * LOG2PI - d * eta - 0.5 * bn2 / sigma2
        if self.variant == "subsampling":
            logprior_sigma = -LOG3 - 0.5 * LOG2PI - sigma2 / 18.0
            dprior_eta = -sigma2 / 9.0
        else:
            logprior_sigma = -LOG3 - 0.5 * LOG2PI - eta * eta / 18.0     # (-eta + eta cancel)
            dprior_eta = -eta / 9.0
        logp = w * ll + logprior_beta + logprior_sigma
        G = np.empty_like(Z)
        G[:d] = w * (self.X.T @ resid) - B / sigma2
        G[d] = -d + bn2 / sigma2 + dprior_eta
        return logp, G

    def logdensity(self, z):
        """One sample, one GEMV over X -- the shape of the reference's inner loop
        (src/algorithms/repgradelbo.jl:84-86)."""
        return self.logdensity_and_gradient(z)[0]

    def subsample(self, idx):
        """docs/src/tutorials/subsampling.md:99-102."""
        idx = np.asarray(idx)
        return type(self)(self.X[idx], self.y[idx], n_data=self.n_data,
                          variant=self.variant, capability=self.capability)


class GaussGLM(LogReg):
    """LogReg(subsampling) with a unit-variance Gaussian likelihood (builder-defined)."""

    family = "gaussian"

    def _loglik_and_resid(self, logits):
        y = self.y[:, None]
        r = y - logits
        return np.sum(-0.5 * LOG2PI - 0.5 * r * r, axis=0), r


# ---------------------------------------------------------------------------------------
# synthetic benchmark data: generated by Philox so that the same data
# can be rebuilt on the GPU box without shipping files.
def synth_glm_data(n: int, d: int, seed: int, family: str = "bernoulli_logit"):
    """X[:, :d-1] ~ N(0,1)/sqrt(d), X[:, d-1] = 1 (intercept); beta* ~ N(0,1);
    y ~ Bernoulli(sigmoid(X beta*)) or N(X beta*, 1).  Returns float32 X (n, d), y (n,)."""
    from .philox import normal_matrix, uniform_u32, STREAM_DATA
    # column-major generation: feature j of row i = normal_matrix[i, j]
    X = normal_matrix(seed, 0, n, d, stream=STREAM_DATA, dtype=np.float64) / np.sqrt(d)
    X[:, d - 1] = 1.0
    beta = normal_matrix(seed, 1, d, 1, stream=STREAM_DATA)[:, 0]
    logits = X @ beta
    if family == "bernoulli_logit":
        u = (uniform_u32(seed, n, STREAM_DATA, offset=1 << 30).astype(np.float64) + 0.5) / 2.0 ** 32
        y = (u < sigmoid(logits)).astype(np.float32)
    else:
        y = (logits + normal_matrix(seed, 2, n, 1, stream=STREAM_DATA)[:, 0]).astype(np.float32)
    return X.astype(np.float32), y
