"""Location-scale variational family (mean-field / full-rank Gaussian).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
src/families/location_scale.jl of the reference; matrices are column-major in the
reference, here samples are COLUMNS of a (D, M) numpy array exactly as there
(`eachsample = eachcol`, src/utils.jl:6).
"""

from __future__ import annotations

import numpy as np

LOG2PI = float(np.log(2.0 * np.pi))
H0 = 0.5 * (LOG2PI + 1.0)   # entropy(Normal(0,1)), Appendix B of SURVEY.md


class StdNormal:
    """Normal(0, 1): the base distribution of MeanFieldGaussian / FullRankGaussian / LowRankGaussian."""
    name = "Normal(0,1)"

    def mean(self): return 0.0
    def var(self): return 1.0
    def entropy(self): return H0
    def logpdf(self, u): return -0.5 * (u * u + LOG2PI)
    def score(self, u): return -u                      # d/du logpdf
    def from_normal(self, eps): return eps             # sampling transform of standard normal draws


class NormalDist(StdNormal):
    """Normal(m, s) base (the :gaussian_nonstd case of test/families/location_scale.jl:22-25)."""

    def __init__(self, m, s):
        self.m, self.s, self.name = float(m), float(s), f"Normal({m},{s})"

    def mean(self): return self.m
    def var(self): return self.s ** 2
    def entropy(self): return H0 + float(np.log(self.s))
    def logpdf(self, u): return -0.5 * (((u - self.m) / self.s) ** 2 + LOG2PI) - np.log(self.s)
    def score(self, u): return -(u - self.m) / self.s ** 2
    def from_normal(self, eps): return self.m + self.s * eps


class LaplaceDist(StdNormal):
    """Laplace(0, 1) base (docs/src/families.md:88-101): density exp(-|u|) / 2."""
    name = "Laplace(0,1)"

    def mean(self): return 0.0
    def var(self): return 2.0
    def entropy(self): return 1.0 + float(np.log(2.0))
    def logpdf(self, u): return -np.abs(u) - np.log(2.0)
    def score(self, u): return -np.sign(u)

    def from_normal(self, eps):
        """Inverse-CDF transform of the uniform Phi(eps) (a device sampler would start from the Philox uniform)."""
        from scipy.special import ndtr
        p = ndtr(eps) - 0.5
        return -np.sign(p) * np.log1p(-2.0 * np.abs(p))


class TDistBase(StdNormal):
    """TDist(nu) base (docs/src/families.md:72-86)."""

    def __init__(self, nu):
        self.nu, self.name = float(nu), f"TDist({nu})"

    def mean(self): return 0.0
    def var(self): return self.nu / (self.nu - 2.0) if self.nu > 2 else np.inf

    def entropy(self):
        from scipy.special import betaln, digamma
        nu = self.nu
        return float((nu + 1) / 2 * (digamma((nu + 1) / 2) - digamma(nu / 2)) + 0.5 * np.log(nu) + betaln(nu / 2, 0.5))

    def logpdf(self, u):
        from scipy.special import gammaln
        nu = self.nu
        c = gammaln((nu + 1) / 2) - gammaln(nu / 2) - 0.5 * np.log(nu * np.pi)
        return c - (nu + 1) / 2 * np.log1p(u * u / nu)

    def score(self, u): return -(self.nu + 1.0) * u / (self.nu + u * u)

    def from_normal(self, eps):
        from scipy.special import ndtr
        from scipy.stats import t as student_t
        return student_t.ppf(ndtr(eps), self.nu)


STD_NORMAL = StdNormal()


class MvLocationScale:
    """src/families/location_scale.jl:15-19; dist = Normal(0, 1) unless another base distribution is given
    (`MvLocationScale(location, scale, dist)`; the device path implements Normal(0, 1) only).

    `scale` is a 1-D array (Diagonal -> MeanFieldGaussian, :139-141) or a 2-D lower
    triangular array (LowerTriangular -> FullRankGaussian, :124-128).
    """

    def __init__(self, location, scale, dist=STD_NORMAL):
        self.location = np.array(location, copy=True)
        self.scale = np.array(scale, copy=True)
        self.dist = dist
        assert self.scale.ndim in (1, 2)
        if self.scale.ndim == 2:
            assert self.scale.shape == (len(self.location),) * 2

    @property
    def is_meanfield(self) -> bool:
        return self.scale.ndim == 1

    def __len__(self):                       # :45
        return len(self.location)

    @property
    def dtype(self):
        return self.location.dtype

    def scale_diag(self) -> np.ndarray:
        return self.scale if self.is_meanfield else np.diag(self.scale)

    # -- Optimisers.destructure ------------------------------------------------------
    def destructure(self) -> np.ndarray:
        """Mean-field: flat = [location; diag(scale)] (location_scale.jl:39-43).
        Full-rank: generic Functors path = [location; vec(scale)] column-major with the
        zero upper triangle present (unpinned by reference tests).
        """
        if self.is_meanfield:
            return np.concatenate([self.location, self.scale])
        return np.concatenate([self.location, self.scale.reshape(-1, order="F")])

    def restructure(self, flat: np.ndarray) -> "MvLocationScale":
        """RestructureMeanField (location_scale.jl:32-37) / generic restructure."""
        D = len(self.location)
        flat = np.asarray(flat)
        if self.is_meanfield:
            assert flat.shape == (2 * D,)
            return MvLocationScale(flat[:D], flat[D:], self.dist)
        assert flat.shape == (D + D * D,)
        return MvLocationScale(flat[:D], flat[D:].reshape(D, D, order="F"), self.dist)

    # -- StatsBase.entropy (location_scale.jl:52-57) -------------------------------
    def entropy(self):
        D = len(self.location)
        return D * self.dtype.type(self.dist.entropy()) + np.sum(np.log(self.scale_diag()))

    # -- Distributions.logpdf (location_scale.jl:59-63) ----------------------------
    def standardize(self, z: np.ndarray) -> np.ndarray:
        """z_std = scale \\ (z - location); z is (D,) or (D, M)."""
        r = z - (self.location if z.ndim == 1 else self.location[:, None])
        if self.is_meanfield:
            return r / (self.scale if z.ndim == 1 else self.scale[:, None])
        from scipy.linalg import solve_triangular
        return solve_triangular(self.scale, r, lower=True)

    def logpdf(self, z: np.ndarray):
        """sum(logpdf(Normal(0,1), z_std)) - logdet(scale); vectorised over columns."""
        u = self.standardize(z)
        return np.sum(self.dist.logpdf(u), axis=0) - np.sum(np.log(self.scale_diag()))

    # -- Distributions.rand (location_scale.jl:71-87) ------------------------------
    def rand_from_eps(self, eps: np.ndarray) -> np.ndarray:
        """scale * eps .+ location (dense, :76) / diag(scale) .* eps .+ location (:86)."""
        if self.is_meanfield:
            return self.scale[:, None] * eps + self.location[:, None]
        return self.scale @ eps + self.location[:, None]

    # -- mean / var / cov (location_scale.jl:98-113), mean(Normal(0,1)) = 0, var = 1 --
    def mean(self):
        if self.dist.mean() == 0.0:
            return self.location.copy()
        m = np.full(len(self.location), self.dist.mean())
        return self.location + (self.scale * m if self.is_meanfield else self.scale @ m)   # :98-101

    def var(self):
        if self.is_meanfield:
            return self.dist.var() * self.scale ** 2
        return self.dist.var() * np.sum(self.scale ** 2, axis=1)          # var(dist) diag(C C')

    def cov(self):
        if self.is_meanfield:
            return self.dist.var() * np.diag(self.scale ** 2)
        return self.dist.var() * (self.scale @ self.scale.T)


def MeanFieldGaussian(mu, diag_scale) -> MvLocationScale:
    """location_scale.jl:139-141."""
    diag_scale = np.asarray(diag_scale)
    assert diag_scale.ndim == 1
    return MvLocationScale(mu, diag_scale)


def FullRankGaussian(mu, L) -> MvLocationScale:
    """location_scale.jl:124-128; L must be lower triangular."""
    L = np.asarray(L)
    assert L.ndim == 2 and np.allclose(L, np.tril(L))
    return MvLocationScale(mu, L)


class MvLocationScaleLowRank:
    """src/families/location_scale_low_rank.jl:16-24 with dist = Normal(0, 1) (`LowRankGaussian`, :132-135):
    z = scale_diag .* u_diag + scale_factors * u_fact + location, covariance diag(scale_diag^2) + U U'.
    Groundwork for SURVEY 8f rank 4: no device kernels use it yet."""

    def __init__(self, location, scale_diag, scale_factors):
        self.location = np.array(location, copy=True)
        self.scale_diag = np.array(scale_diag, copy=True)
        self.scale_factors = np.array(scale_factors, copy=True)
        d = len(self.location)
        assert self.scale_diag.shape == (d,) and self.scale_factors.ndim == 2 and self.scale_factors.shape[0] == d

    def __len__(self):                                   # :28
        return len(self.location)

    @property
    def rank(self) -> int:
        return self.scale_factors.shape[1]

    # Functors.@functor order (location, scale_diag, scale_factors) (:26); matrices flatten column-major
    def destructure(self) -> np.ndarray:
        return np.concatenate([self.location, self.scale_diag, self.scale_factors.reshape(-1, order="F")])

    def restructure(self, flat) -> "MvLocationScaleLowRank":
        d, r = len(self.location), self.rank
        flat = np.asarray(flat)
        assert flat.shape == (2 * d + d * r,)
        return MvLocationScaleLowRank(flat[:d], flat[d:2 * d], flat[2 * d:].reshape(d, r, order="F"))

    def entropy(self):                                   # :34-43
        d = len(self.location)
        D2 = self.scale_diag * self.scale_diag
        UtDinvU = self.scale_factors.T @ (self.scale_factors / D2[:, None])
        logdet_sigma = 2.0 * np.sum(np.log(self.scale_diag)) + np.linalg.slogdet(np.eye(self.rank) + UtDinvU)[1]
        return d * H0 + logdet_sigma / 2.0

    def logpdf(self, z: np.ndarray):                     # :45-70 (differentiable O(d^3) path; mean(dist) = 0)
        from scipy.linalg import cholesky, solve_triangular
        scale2 = np.diag(self.scale_diag ** 2) + self.scale_factors @ self.scale_factors.T
        Lc = cholesky(scale2, lower=True)
        r = z - (self.location if z.ndim == 1 else self.location[:, None])
        u = solve_triangular(Lc, r, lower=True)
        return np.sum(-0.5 * (u * u + LOG2PI), axis=0) - np.sum(np.log(np.diag(Lc)))

    def rand_from_eps(self, u_diag: np.ndarray, u_fact: np.ndarray) -> np.ndarray:   # :79-86
        """u_diag: (d, M), u_fact: (r, M) independent standard normal draws."""
        return self.scale_diag[:, None] * u_diag + self.scale_factors @ u_fact + self.location[:, None]

    def mean(self):                                      # :99-105
        return self.location.copy()

    def var(self):                                       # :107-111
        return self.scale_diag ** 2 + np.sum(self.scale_factors ** 2, axis=1)

    def cov(self):                                       # :113-117
        return np.diag(self.scale_diag ** 2) + self.scale_factors @ self.scale_factors.T

    def entropy_gradient(self):
        """d entropy / d (scale_diag, scale_factors): with W = D^-2 U and B = I + U' W,
        dH/dU = W B^-1 and dH/dD_i = 1 / D_i - (U B^-1 U')_ii / D_i^3 (the r x r capacitance matrix is all a
        device kernel would have to factor)."""
        D = self.scale_diag
        W = self.scale_factors / (D * D)[:, None]
        Binv = np.linalg.inv(np.eye(self.rank) + self.scale_factors.T @ W)
        gU = W @ Binv
        gD = 1.0 / D - np.einsum("ik,kl,il->i", self.scale_factors, Binv, self.scale_factors) / D ** 3
        return gD, gU


    # -- Woodbury pieces a device kernel needs for logpdf-based estimators (only r x r systems are solved) --
    def _capacitance_inv(self):
        W = self.scale_factors / (self.scale_diag ** 2)[:, None]
        return W, np.linalg.inv(np.eye(self.rank) + self.scale_factors.T @ W)

    def cov_solve(self, R: np.ndarray) -> np.ndarray:
        """Sigma^-1 R for columns R (d, M): D^-2 R - W B^-1 U' D^-2 R with W = D^-2 U, B = I + U' W."""
        W, Binv = self._capacitance_inv()
        DR = R / (self.scale_diag ** 2)[:, None]
        return DR - W @ (Binv @ (self.scale_factors.T @ DR))

    def cov_inv_diag_and_factor(self):
        """diag(Sigma^-1) and Sigma^-1 U (= W B^-1): what d log q / d scale_diag, d scale_factors need."""
        W, Binv = self._capacitance_inv()
        SinvU = W @ Binv
        diag = 1.0 / self.scale_diag ** 2 - np.einsum("ik,ik->i", SinvU, W)
        return diag, SinvU


def LowRankGaussian(mu, D, U) -> MvLocationScaleLowRank:
    """location_scale_low_rank.jl:119-135."""
    return MvLocationScaleLowRank(mu, D, U)
