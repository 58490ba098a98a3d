"""Host-side mirror of AdvancedVI.jl's interface for the ELBO-gradient path, over the C ABI.

The reference's host language (Julia) is not available in this image, so this module plays the
role the Julia glue (`julia/AdvancedVIB200.jl`) plays for a Julia user: same names, same argument
meaning and error behaviour as the reference (paths below are under /root/reference):

  MvLocationScale / MeanFieldGaussian / FullRankGaussian   src/families/location_scale.jl:15-141
  RepGradELBO / ScoreGradELBO / SubsampledObjective         src/algorithms/{repgradelbo,scoregradelbo,subsampledobjective}.jl
  entropy estimators                                        src/algorithms/entropy.jl
  KLMinRepGradDescent (ADVI) / KLMinRepGradProxDescent / KLMinScoreGradDescent (BBVI)
                                                            src/algorithms/constructors.jl:44-233
  optimize / estimate_objective                             src/optimize.jl:42-94, src/algorithms/common.jl:29-120
  ReshufflingBatchSubsampling                               src/reshuffling.jl:13-60
  Descent / Adam / DoG / DoWG, ClipScale, ProximalLocationScaleEntropy, PolynomialAveraging

All arithmetic runs in libavi_b200.so on the GPU (float32).  This file holds only host logic:
argument checking, flattening q, the minibatch state machine and the optimisation loop.
"""

from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib as L

__all__ = [
    "Context", "MvNormalDiag", "LogReg", "GaussGLM", "HostCallbackProblem",
    "MvLocationScale", "MeanFieldGaussian", "FullRankGaussian", "MvLocationScaleLowRank", "LowRankGaussian",
    "Normal", "Laplace", "TDist",
    "ClosedFormEntropy", "MonteCarloEntropy", "StickingTheLandingEntropy",
    "ClosedFormEntropyZeroGradient", "StickingTheLandingEntropyZeroGradient",
    "RepGradELBO", "ScoreGradELBO", "SubsampledObjective", "ReshufflingBatchSubsampling",
    "Descent", "Adam", "DoG", "DoWG", "IdentityOperator", "ClipScale", "ProximalLocationScaleEntropy",
    "NoAveraging", "PolynomialAveraging",
    "KLMinRepGradDescent", "KLMinRepGradProxDescent", "KLMinScoreGradDescent", "ADVI", "BBVI",
    "optimize", "estimate_objective", "Objective", "AviError", "HostUpdate", "HostStep", "central_fd_gradient",
    "gaussian_expectation_gradient_and_hessian",
]

AviError = L.AviError


def _f32(a, name="array"):
    a = np.asarray(a)
    if a.dtype != np.float32:
        if a.dtype == np.float64 or np.issubdtype(a.dtype, np.integer) or a.dtype == np.bool_:
            return np.ascontiguousarray(a, dtype=np.float32)
        raise TypeError(f"{name}: unsupported element type {a.dtype}; the B200 path is Float32 only")
    return np.ascontiguousarray(a)


class _PtrCache:
    """float* of caller-owned arrays that come back call after call (the parameter vector and the gradient buffer of an
    optimisation loop): building a ctypes pointer costs ~5 us, as much as the device-side staging of the call.  The
    cache holds a reference to the array, so an id() cannot be recycled while its entry is alive."""

    def __init__(self, limit=8):
        self.d, self.limit = {}, limit

    def __call__(self, a):
        k = id(a)
        c = self.d.get(k)
        if c is None or c[0] is not a or c[2] != a.__array_interface__["data"][0]:
            if len(self.d) >= self.limit:
                self.d.clear()
            c = (a, a.ctypes.data_as(L.c_float_p), a.__array_interface__["data"][0])
            self.d[k] = c
        return c[1]


def _key_from(rng):
    """The host rng is used only to draw the 64-bit Philox key (include/avi.h: avi_obj_seed)."""
    if isinstance(rng, (int, np.integer)):
        return int(rng) & 0xFFFFFFFFFFFFFFFF
    if isinstance(rng, np.random.Generator):
        return int(rng.integers(0, 2 ** 63, dtype=np.int64)) & 0xFFFFFFFFFFFFFFFF
    raise TypeError("rng must be an int seed or a numpy.random.Generator")


# ---------------------------------------------------------------------------------------------
class Context:
    """One CUDA device + stream (+ optional multi-rank exchange)."""

    def __init__(self, device: int = 0):
        h = L.vp()
        L.check(L.lib.avi_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self.rank, self.nranks = 0, 1
        self._cb = None

    def info(self):
        sm, mem, ma, mi = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        L.check(L.lib.avi_ctx_info(self.h, C.byref(sm), C.byref(mem), C.byref(ma), C.byref(mi)), self.h)
        return dict(sm_count=sm.value, hbm_bytes=mem.value, cc=(ma.value, mi.value))

    def launch_count(self) -> int:
        return int(L.lib.avi_ctx_launch_count(self.h))

    def synchronize(self):
        L.check(L.lib.avi_ctx_synchronize(self.h), self.h)

    def stream(self) -> int:
        return int(L.lib.avi_ctx_stream(self.h) or 0)

    def timing(self, enable: bool):
        L.check(L.lib.avi_ctx_timing(self.h, int(enable)), self.h)

    def kernel_time(self, name: str):
        """(total_ms, launches) of the named hot kernel since timing was enabled."""
        ms, cnt = C.c_double(), C.c_int64()
        L.check(L.lib.avi_ctx_timing_get(self.h, name.encode(), C.byref(ms), C.byref(cnt)), self.h)
        return ms.value, cnt.value

    def set_allreduce(self, fn, rank: int, nranks: int):
        """fn(dev_ptr: int, count: int, stream: int) -> None sums `count` floats at dev_ptr over ranks."""
        def tramp(user, buf, count, stream):
            try:
                fn(int(buf), int(count), int(stream or 0))
                return 0
            except Exception:   # noqa: BLE001 - reported through the status code
                import traceback
                traceback.print_exc()
                return 1
        self._cb = L.ALLREDUCE_FN(tramp)
        L.check(L.lib.avi_ctx_set_allreduce(self.h, self._cb, None, rank, nranks), self.h)
        self.rank, self.nranks = rank, nranks

    def connect_peers(self, max_floats: int, rank: int, nranks: int, allgather_bytes):
        """Native NVLink exchange: allgather_bytes(b: bytes) -> list[bytes] over the ranks."""
        buf = C.create_string_buffer(64)
        L.check(L.lib.avi_comm_buffer(self.h, int(max_floats), buf), self.h)
        handles = allgather_bytes(buf.raw)
        assert len(handles) == nranks and all(len(x) == 64 for x in handles)
        L.check(L.lib.avi_comm_connect(self.h, rank, nranks, b"".join(handles)), self.h)
        self.rank, self.nranks = rank, nranks

    def comm_barrier(self):
        """Device-side rendezvous of the connected ranks, enqueued on this context's stream."""
        L.check(L.lib.avi_comm_barrier(self.h), self.h)

    def disconnect_peers(self):
        """Unmap the peers' exchange buffers; call on every rank and synchronise the ranks before re-connecting."""
        L.check(L.lib.avi_comm_disconnect(self.h), self.h)
        self.rank, self.nranks = 0, 1

    def close(self):
        if self.h:
            L.lib.avi_ctx_destroy(self.h)
            self.h = None


# ---------------------------------------------------------------------------------------------
# targets: the LogDensityProblems side (dimension / capabilities / logdensity_and_gradient)
class _Problem:
    h = None
    ctx: Context = None

    def dimension(self) -> int:
        return int(L.lib.avi_model_dimension(self.h))

    @property
    def capability(self) -> int:
        return int(L.lib.avi_model_capability(self.h))

    def logdensity_and_gradient(self, Z, want_grad=True):
        """Z: (D, M) or (D,) -> (logp (M,), G (D, M)).  Batched LogDensityProblems.logdensity_and_gradient."""
        Z = _f32(Z, "Z")
        single = Z.ndim == 1
        Zc = np.ascontiguousarray((Z[:, None] if single else Z).T)       # sample-major (M, D)
        M, D = Zc.shape
        if D != self.dimension():
            raise ValueError("dimension mismatch")
        logp = np.empty(M, np.float32)
        G = np.empty((M, D), np.float32) if want_grad else None
        L.check(L.lib.avi_model_logdensity_and_gradient_host(self.h, L.fptr(Zc), M, L.fptr(logp), L.fptr(G)), self.ctx.h)
        if single:
            return float(logp[0]), (G[0] if want_grad else None)
        return logp, (np.ascontiguousarray(G.T) if want_grad else None)

    def logdensity(self, Z):
        return self.logdensity_and_gradient(Z, want_grad=False)[0]

    def subsample(self, batch):
        """AdvancedVI.subsample(prob, batch) (src/AdvancedVI.jl:303-313); 0-based row indices, None = all rows.
        Mutates and returns the same object, which the reference allows."""
        if batch is None:
            L.check(L.lib.avi_model_subsample(self.h, None, 0), self.ctx.h)
        else:
            idx = np.ascontiguousarray(batch, dtype=np.int32)
            L.check(L.lib.avi_model_subsample(self.h, L.iptr(idx), len(idx)), self.ctx.h)
        return self

    def close(self):
        if self.h:
            L.lib.avi_model_destroy(self.h)
            self.h = None


class MvNormalDiag(_Problem):
    """logpdf(MvNormal(mu, Diagonal(sigma.^2)), z) -- test/models/normal.jl:8-11, :56-75."""

    def __init__(self, ctx: Context, mu, sigma):
        mu, sigma = _f32(mu), _f32(sigma)
        if mu.shape != sigma.shape or mu.ndim != 1:
            raise ValueError("mu and sigma must be vectors of equal length")
        h = L.vp()
        L.check(L.lib.avi_model_mvnormal_diag_create(ctx.h, L.fptr(mu), L.fptr(sigma), len(mu), C.byref(h)), ctx.h)
        self.h, self.ctx = h, ctx


_GEMM = {"fp32": L.GEMM_SIMT_FP32, "simt": L.GEMM_SIMT_FP32, "tf32": L.GEMM_TF32, "tf32x3": L.GEMM_TF32X3}


class LogReg(_Problem):
    """Hierarchical logistic regression, theta = [beta; log sigma].
    variant="subsampling": docs/src/tutorials/subsampling.md:26-38; variant="basic": README.md:47-58 + :91-106.
    X is (n, d) (any layout; copied once to the device), y in {0, 1}^n."""

    likelihood = L.GLM_BERNOULLI_LOGIT

    def __init__(self, ctx: Context, X, y, n_data=None, variant="subsampling", gemm="tf32"):
        X = _f32(X, "X")
        y = _f32(y, "y")
        if X.ndim != 2 or y.shape != (X.shape[0],):
            raise ValueError("X must be (n, d) and y (n,)")
        n, d = X.shape
        Xf = np.asfortranarray(X)                # column-major n x d, as the Julia host holds it
        h = L.vp()
        L.check(L.lib.avi_model_glm_create(ctx.h, Xf.ctypes.data_as(L.c_float_p), L.fptr(y), n, d,
                                           int(n if n_data is None else n_data), self.likelihood,
                                           L.GLM_SUBSAMPLING if variant == "subsampling" else L.GLM_BASIC,
                                           _GEMM[gemm], C.byref(h)), ctx.h)
        self.h, self.ctx, self.n, self.d = h, ctx, n, d

    def set_data_shard(self, nshards: int, rows_global: int, include_prior: bool):
        L.check(L.lib.avi_model_set_data_shard(self.h, nshards, rows_global, int(include_prior)), self.ctx.h)

    def set_fused_step(self, mode: int):
        """0: one kernel per stage, 1: library default, 2: always the single whole-iteration kernel."""
        L.check(L.lib.avi_model_set_fused_step(self.h, int(mode)), self.ctx.h)


class GaussGLM(LogReg):
    """LogReg(subsampling) with y_i ~ Normal(x_i' beta, 1) (builder-defined; BASELINE.json config 4)."""
    likelihood = L.GLM_GAUSSIAN


def central_fd_gradient(f, z, rel_step=1e-5):
    """Central finite differences of a scalar function in Float64: the default gradient of capability-0 targets."""
    z = np.asarray(z, np.float64).copy()
    g = np.empty_like(z)
    for i in range(z.size):
        h = rel_step * max(1.0, abs(z[i]))
        zi = z[i]
        z[i] = zi + h; fp = float(f(z))
        z[i] = zi - h; fm = float(f(z))
        z[i] = zi
        g[i] = (fp - fm) / (2.0 * h)
    return g


class HostCallbackProblem(_Problem):
    """Any LogDensityProblem: fn(z: (D,) float32) -> logp  (capability 0)  or  (logp, grad)  (capability 1).
    One call per Monte-Carlo sample, as the reference does (src/algorithms/repgradelbo.jl:84-86).

    Capability 0 under RepGradELBO: the reference differentiates through `logdensity` with its AD backend
    (src/algorithms/repgradelbo.jl:50-62).  The native path has no AD backend, so such a target needs
    `fallback_gradient`: "central_fd" (central finite differences in Float64) or a callable (f, z) -> gradient; the
    target is then presented to the library as first-order.  Without it RepGradELBO is refused with an explanation
    (ScoreGradELBO and estimate_objective never need a gradient)."""

    def __init__(self, ctx: Context, D: int, fn, capability: int = 1, fallback_gradient=None):
        if capability == 0 and fallback_gradient is not None:
            grad_of = central_fd_gradient if fallback_gradient == "central_fd" else fallback_gradient

            def scalar(z):          # (called with Float64 points: differences of Float32 evaluations would drown h)
                r = fn(z)
                return r[0] if isinstance(r, tuple) else r

            def first_order(z):
                return scalar(z), grad_of(scalar, z.astype(np.float64))
            fn_used, cap_used = first_order, 1
        else:
            fn_used, cap_used = fn, capability

        def tramp(user, z, Dn, logp, grad):
            try:
                zz = np.ctypeslib.as_array(z, shape=(Dn,)).copy()
                r = fn_used(zz)
                if cap_used >= 1:
                    lp, g = r
                    if grad:
                        np.ctypeslib.as_array(grad, shape=(Dn,))[:] = np.asarray(g, dtype=np.float32)
                else:
                    lp = r[0] if isinstance(r, tuple) else r
                logp[0] = float(lp)
                return 0
            except Exception:   # noqa: BLE001
                import traceback
                traceback.print_exc()
                return 1
        self._cb = L.LOGDENSITY_FN(tramp)
        h = L.vp()
        L.check(L.lib.avi_model_hostcallback_create(ctx.h, D, cap_used, self._cb, None, C.byref(h)), ctx.h)
        self.h, self.ctx = h, ctx


# ---------------------------------------------------------------------------------------------
# variational family (host container; the arithmetic is in the library)
class Normal:
    """Normal(0, 1): the base distribution of MeanFieldGaussian / FullRankGaussian."""
    code, param = 0, 0.0

    def __repr__(self):
        return "Normal(0, 1)"


class Laplace(Normal):
    """Laplace(0, 1) base distribution (docs/src/families.md:88-101)."""
    code, param = 1, 0.0

    def __repr__(self):
        return "Laplace(0, 1)"


class TDist(Normal):
    """TDist(nu) base distribution (docs/src/families.md:72-86)."""
    code = 2

    def __init__(self, nu: float):
        if not nu > 0:
            raise ValueError("TDist: nu must be positive")
        self.param = float(nu)

    def __repr__(self):
        return f"TDist({self.param})"


class MvLocationScale:
    """MvLocationScale(location, scale, dist) (src/families/location_scale.jl:15-19); dist = Normal(0, 1) unless given
    (Laplace(), TDist(nu): docs/src/families.md:72-101); float32 only."""

    def __init__(self, location, scale, dist=None):
        self.dist = dist if dist is not None else Normal()
        if not isinstance(self.dist, Normal):
            raise TypeError("dist must be Normal(), Laplace() or TDist(nu)")
        for a in (location, scale):
            if np.asarray(a).dtype not in (np.float32,):
                raise TypeError("MvLocationScale: the B200 path supports Float32 only "
                                f"(got {np.asarray(a).dtype}); convert with .astype(np.float32)")
        self.location = np.array(location, dtype=np.float32)
        self.scale = np.array(scale, dtype=np.float32)
        if self.scale.ndim == 2 and self.scale.shape != (len(self.location),) * 2:
            raise ValueError("scale must be D x D")

    @property
    def is_meanfield(self):
        return self.scale.ndim == 1

    @property
    def family(self):
        return L.MEANFIELD if self.is_meanfield else L.FULLRANK

    def __len__(self):
        return len(self.location)

    def destructure(self):
        """[location; diag(scale)] (location_scale.jl:39-43) or [location; vec(scale)] (generic Functors path)."""
        if self.is_meanfield:
            return np.concatenate([self.location, self.scale])
        return np.concatenate([self.location, self.scale.reshape(-1, order="F")])

    def restructure(self, flat):
        D = len(self.location)
        flat = np.asarray(flat, dtype=np.float32)
        if self.is_meanfield:
            return MvLocationScale(flat[:D].copy(), flat[D:].copy(), self.dist)
        return MvLocationScale(flat[:D].copy(), flat[D:].reshape(D, D, order="F").copy(), self.dist)


class MvLocationScaleLowRank:
    """src/families/location_scale_low_rank.jl:16-24 with dist = Normal(0, 1) (`LowRankGaussian`): covariance
    diag(scale_diag^2) + scale_factors scale_factors'; float32 only.  Flat parameters follow the Functors order
    (location, scale_diag, scale_factors) with the factor matrix column-major."""

    family = L.LOWRANK

    def __init__(self, location, scale_diag, scale_factors):
        for a in (location, scale_diag, scale_factors):
            if np.asarray(a).dtype not in (np.float32,):
                raise TypeError("MvLocationScaleLowRank: the B200 path supports Float32 only "
                                f"(got {np.asarray(a).dtype}); convert with .astype(np.float32)")
        self.location = np.array(location, dtype=np.float32)
        self.scale_diag = np.array(scale_diag, dtype=np.float32)
        self.scale_factors = np.array(scale_factors, dtype=np.float32)
        d = len(self.location)
        if self.scale_diag.shape != (d,) or self.scale_factors.ndim != 2 or self.scale_factors.shape[0] != d:
            raise ValueError("scale_diag must have D entries and scale_factors must be D x rank")

    @property
    def rank(self):
        return self.scale_factors.shape[1]

    def __len__(self):
        return len(self.location)

    def destructure(self):
        return np.concatenate([self.location, self.scale_diag, self.scale_factors.reshape(-1, order="F")])

    def restructure(self, flat):
        d, r = len(self.location), self.rank
        flat = np.asarray(flat, dtype=np.float32)
        return MvLocationScaleLowRank(flat[:d].copy(), flat[d:2 * d].copy(), flat[2 * d:].reshape(d, r, order="F").copy())

    def cov(self):                                   # location_scale_low_rank.jl:113-117
        return np.diag(self.scale_diag.astype(np.float64) ** 2) + self.scale_factors.astype(np.float64) @ self.scale_factors.T


def LowRankGaussian(mu, D, U):
    """location_scale_low_rank.jl:119-135."""
    return MvLocationScaleLowRank(mu, D, U)


def MeanFieldGaussian(mu, diag_scale):
    """location_scale.jl:139-141."""
    if np.asarray(diag_scale).ndim != 1:
        raise ValueError("MeanFieldGaussian needs the diagonal of the scale")
    return MvLocationScale(mu, diag_scale)


def FullRankGaussian(mu, Lmat):
    """location_scale.jl:124-128."""
    Lmat = np.asarray(Lmat)
    if Lmat.ndim != 2 or not np.array_equal(Lmat, np.tril(Lmat)):
        raise ValueError("FullRankGaussian needs a lower-triangular scale")
    return MvLocationScale(mu, Lmat)


# ---------------------------------------------------------------------------------------------
class _Entropy:
    code = None

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self))


class ClosedFormEntropy(_Entropy):                       # entropy.jl:25-29
    code = L.ENT_CLOSEDFORM


class MonteCarloEntropy(_Entropy):                       # entropy.jl:40-46
    code = L.ENT_MONTECARLO


class StickingTheLandingEntropy(_Entropy):               # entropy.jl:57-65
    code = L.ENT_STL


class ClosedFormEntropyZeroGradient(_Entropy):           # entropy.jl:11-15
    code = L.ENT_CLOSEDFORM_ZEROGRAD


class StickingTheLandingEntropyZeroGradient(_Entropy):   # entropy.jl:78-90
    code = L.ENT_STL_ZEROGRAD


class RepGradELBO:
    """repgradelbo.jl:21-24."""
    kind = L.REPGRAD

    def __init__(self, n_samples: int, entropy: _Entropy = None):
        self.n_samples = int(n_samples)
        self.entropy = ClosedFormEntropy() if entropy is None else entropy


class ScoreGradELBO:
    """scoregradelbo.jl:15-17 (the VarGrad objective; it has no entropy option)."""
    kind = L.SCOREGRAD

    def __init__(self, n_samples: int):
        self.n_samples = int(n_samples)
        self.entropy = ClosedFormEntropy()   # unused by the library for SCOREGRAD


class ReshufflingBatchSubsampling:
    """reshuffling.jl:13-16; `dataset` is a vector of 0-based row indices."""

    def __init__(self, dataset, batchsize: int):
        self.dataset = np.ascontiguousarray(dataset, dtype=np.int32)
        self.batchsize = int(batchsize)
        if self.batchsize < 1:
            raise ValueError("batchsize must be >= 1")
        if self.dataset.size and self.dataset.min() < 0:
            raise ValueError("dataset holds 0-based row indices: negative entries are invalid")

    def __len__(self):                                    # reshuffling.jl:23-25
        return -(-len(self.dataset) // self.batchsize)

    def reshuffle_batches(self, key: int, shuffle_index: int):     # :27-32
        perm = self.dataset.copy()
        L.check(L.lib.avi_shuffle(key, shuffle_index, len(perm), L.iptr(perm)))
        b = self.batchsize
        return [(k + 1, perm[k * b:(k + 1) * b]) for k in range(len(self))]


class _SubState:
    """reshuffling.jl:18-21; n_shuffles stands for the position in the rng stream."""

    def __init__(self, epoch, iterator, n_shuffles, key):
        self.epoch, self.iterator, self.n_shuffles, self.key = epoch, iterator, n_shuffles, key

    def copy(self):
        return _SubState(self.epoch, list(self.iterator), self.n_shuffles, self.key)


def _sub_init(sub, key):                                  # reshuffling.jl:34-36
    return _SubState(1, sub.reshuffle_batches(key, 0), 1, key)


def _sub_step(sub, st, drop_trailing=False):              # reshuffling.jl:38-60
    epoch, it, nsh = st.epoch, list(st.iterator), st.n_shuffles
    (k, batch), it = it[0], it[1:]
    if not it:
        it = sub.reshuffle_batches(st.key, nsh)
        nsh += 1
        if drop_trailing and len(batch) < sub.batchsize:
            (k, batch), it = it[0], it[1:]
        epoch += 1
    return batch, _SubState(epoch, it, nsh, st.key), dict(epoch=epoch, step=k)


class SubsampledObjective:
    """subsampledobjective.jl:10-14."""

    def __init__(self, objective, subsampling: ReshufflingBatchSubsampling):
        self.objective, self.subsampling = objective, subsampling
        self.n_samples = objective.n_samples
        self.entropy = objective.entropy
        self.kind = objective.kind


# ---------------------------------------------------------------------------------------------
class Descent:
    code = L.RULE_DESCENT

    def __init__(self, eta=0.1):
        self.hyper = [float(eta)]


class Adam:
    code = L.RULE_ADAM

    def __init__(self, eta=1e-3, beta=(0.9, 0.999), epsilon=1e-8):
        self.hyper = [float(eta), float(beta[0]), float(beta[1]), float(epsilon)]


class DoG:                                                # rules.jl:48-64
    code = L.RULE_DOG

    def __init__(self, alpha=1e-6):
        self.hyper = [float(alpha)]


class DoWG:                                               # rules.jl:17-34
    code = L.RULE_DOWG

    def __init__(self, alpha=1e-6):
        self.hyper = [float(alpha)]


class IdentityOperator:
    code, param = L.OP_IDENTITY, 0.0


class ClipScale:                                          # clip_scale.jl:8-16
    code = L.OP_CLIPSCALE

    def __init__(self, epsilon=1e-5):
        self.param = float(epsilon)


class ProximalLocationScaleEntropy:                       # proximal_location_scale_entropy.jl:20
    code, param = L.OP_PROXENTROPY, 0.0


class NoAveraging:
    code, param = L.AVG_NONE, 0.0


class PolynomialAveraging:                                # averaging.jl:26-40
    code = L.AVG_POLYNOMIAL

    def __init__(self, eta=8):
        self.param = float(eta)


# ---------------------------------------------------------------------------------------------
class _ParamSpaceSGD:
    def __init__(self, objective, optimizer, averager, operator):
        self.objective, self.optimizer, self.averager, self.operator = objective, optimizer, averager, operator


def KLMinRepGradDescent(optimizer=None, entropy=None, n_samples: int = 1, averager=None, operator=None,
                        subsampling=None):
    """constructors.jl:44-77 (adtype is implied: the native closed-form gradient)."""
    entropy = ClosedFormEntropy() if entropy is None else entropy
    if not isinstance(entropy, (ClosedFormEntropy, StickingTheLandingEntropy, MonteCarloEntropy)):
        raise ValueError("entropy must be ClosedFormEntropy, StickingTheLandingEntropy or MonteCarloEntropy")  # :60
    obj = RepGradELBO(n_samples, entropy)
    if subsampling is not None:
        obj = SubsampledObjective(obj, subsampling)
    return _ParamSpaceSGD(obj, optimizer or DoWG(), averager or PolynomialAveraging(), operator or IdentityOperator())


ADVI = KLMinRepGradDescent


def KLMinRepGradProxDescent(optimizer=None, entropy_zerograd=None, n_samples: int = 1, averager=None,
                            subsampling=None):
    """constructors.jl:122-157."""
    ent = ClosedFormEntropyZeroGradient() if entropy_zerograd is None else entropy_zerograd
    if not isinstance(ent, (ClosedFormEntropyZeroGradient, StickingTheLandingEntropyZeroGradient)):
        raise ValueError("entropy_zerograd must be a ...ZeroGradient estimator")
    obj = RepGradELBO(n_samples, ent)
    if subsampling is not None:
        obj = SubsampledObjective(obj, subsampling)
    return _ParamSpaceSGD(obj, optimizer or DoWG(), averager or PolynomialAveraging(), ProximalLocationScaleEntropy())


def KLMinScoreGradDescent(optimizer=None, n_samples: int = 1, averager=None, operator=None, subsampling=None):
    """constructors.jl:199-231."""
    obj = ScoreGradELBO(n_samples)
    if subsampling is not None:
        obj = SubsampledObjective(obj, subsampling)
    return _ParamSpaceSGD(obj, optimizer or DoWG(), averager or PolynomialAveraging(), operator or IdentityOperator())


BBVI = KLMinScoreGradDescent


# ---------------------------------------------------------------------------------------------
class Objective:
    """Objective state: init / estimate_gradient! / estimate_objective / set_objective_state_problem
    (src/algorithms/abstractobjective.jl:25-86) bound to one target and one family."""

    def __init__(self, rng, objective, q: MvLocationScale, prob: _Problem):
        base = objective.objective if isinstance(objective, SubsampledObjective) else objective
        self.ctx, self.prob, self.q_template, self.spec = prob.ctx, prob, q, base
        if len(q) != prob.dimension():
            raise ValueError("q and the target have different dimensions")
        h = L.vp()
        if q.family == L.LOWRANK:
            L.check(L.lib.avi_obj_create_lowrank(self.ctx.h, prob.h, q.rank, base.kind, base.entropy.code, base.n_samples,
                                                 C.byref(h)), self.ctx.h)
        else:
            L.check(L.lib.avi_obj_create(self.ctx.h, prob.h, q.family, base.kind, base.entropy.code, base.n_samples,
                                         C.byref(h)), self.ctx.h)
        self.h = h
        dist = getattr(q, "dist", None)
        if dist is not None and dist.code != 0:
            L.check(L.lib.avi_obj_set_base(h, dist.code, dist.param), self.ctx.h)
        self.P = int(L.lib.avi_obj_num_params(h))
        self._ptr = _PtrCache()
        self._v, self._e = C.c_float(), C.c_float()
        self.key = _key_from(rng)
        L.check(L.lib.avi_obj_seed(h, self.key, 0), self.ctx.h)

    def seed(self, key: int, step: int = 0):
        self.key = key
        L.check(L.lib.avi_obj_seed(self.h, key, step), self.ctx.h)

    def step_counter(self) -> int:
        s = C.c_uint64()
        L.check(L.lib.avi_obj_get_step(self.h, C.byref(s)), self.ctx.h)
        return s.value

    def set_problem(self, prob: _Problem):
        L.check(L.lib.avi_obj_set_model(self.h, prob.h), self.ctx.h)
        self.prob = prob

    def set_sample_shard(self, m0: int, m_local: int):
        L.check(L.lib.avi_obj_set_sample_shard(self.h, m0, m_local), self.ctx.h)

    def set_shard_axis(self, axis: int):
        L.check(L.lib.avi_obj_set_shard_axis(self.h, axis), self.ctx.h)

    def estimate_gradient(self, params, out=None):
        """-> (value, gradient, elbo): value = -ELBO (RepGrad) or the VarGrad value (ScoreGrad).
        `out`: caller-owned float32 gradient buffer of P entries, written in place like the DiffResults buffer of
        estimate_gradient! (abstractobjective.jl:67-86)."""
        if not (type(params) is np.ndarray and params.dtype == np.float32 and params.flags.c_contiguous):
            params = _f32(params, "params")
        if out is None:
            grad = np.empty(self.P, np.float32)
        else:
            grad = out
            if grad.dtype != np.float32 or grad.size != self.P or not grad.flags.c_contiguous:
                raise ValueError("out must be a contiguous float32 array of num_params entries")
        v, e = self._v, self._e
        rc = L.lib.avi_obj_estimate_gradient(self.h, self._ptr(params), params.size, self._ptr(grad), C.byref(v), C.byref(e))
        if rc:
            L.check(rc, self.ctx.h)
        return v.value, grad, e.value

    def estimate_objective(self, rng, q: MvLocationScale, n_samples: int, kind=None, entropy=None):
        params = q.destructure()
        out = C.c_float()
        kind = self.spec.kind if kind is None else kind
        ent = (self.spec.entropy if entropy is None else entropy).code
        L.check(L.lib.avi_obj_estimate_objective(self.h, L.fptr(params), len(params), int(n_samples), kind, ent,
                                                 _key_from(rng), C.byref(out)), self.ctx.h)
        return out.value

    def rand(self, q: MvLocationScale):
        """rand(rng, q, M) for the current step (not advancing it): (Z, eps), each (D, M_local)."""
        params = q.destructure()
        D = len(q)
        M = self.spec.n_samples
        Z = np.empty((M, D), np.float32)
        E = np.empty((M, D), np.float32)
        L.check(L.lib.avi_obj_rand(self.h, L.fptr(params), len(params), L.fptr(Z), L.fptr(E)), self.ctx.h)
        return np.ascontiguousarray(Z.T), np.ascontiguousarray(E.T)

    def rand_batch_match_samples_with_objective(self, q: MvLocationScale, n_samples: int):
        """rand_batch_match_samples_with_objective!(rng, q, n_samples, prob, u_buf, grad_buf), the sampling stage of
        FisherMinBatchMatch (src/algorithms/fisherminbatchmatch.jl:81-111) -> (u, z, grad (each D x n_samples), fisher,
        logpi_avg); the draws are the objective's Philox stream at its current step (which then advances)."""
        params = q.destructure()
        D = len(q)
        U, Z, G = (np.empty((n_samples, D), np.float32) for _ in range(3))
        fisher, lp = C.c_float(), C.c_float()
        L.check(L.lib.avi_obj_batch_match_samples(self.h, L.fptr(params), len(params), int(n_samples), L.fptr(U), L.fptr(Z),
                                                  L.fptr(G), C.byref(fisher), C.byref(lp)), self.ctx.h)
        return (np.ascontiguousarray(U.T), np.ascontiguousarray(Z.T), np.ascontiguousarray(G.T), fisher.value, lp.value)

    def gaussian_expectation_gradient_and_hessian(self, q: MvLocationScale, n_samples: int):
        """gaussian_expectation_gradient_and_hessian!(rng, q, n_samples, grad_buf, hess_buf, prob), first-order (Stein)
        branch of src/algorithms/gauss_expected_grad_hess.jl:20-58 -> (logpi_avg, grad (D), hess (D, D)); the draws
        are the objective's Philox stream at its current step (which then advances)."""
        params = q.destructure()
        D = len(q)
        grad, hess, lp = np.empty(D, np.float32), np.empty(D * D, np.float32), C.c_float()
        L.check(L.lib.avi_obj_gauss_expected_grad_hess(self.h, L.fptr(params), len(params), int(n_samples), C.byref(lp),
                                                       L.fptr(grad), L.fptr(hess)), self.ctx.h)
        return lp.value, grad, hess.reshape(D, D, order="F")

    def close(self):
        if self.h:
            L.lib.avi_obj_destroy(self.h)
            self.h = None


def gaussian_expectation_gradient_and_hessian(rng, q: MvLocationScale, n_samples: int, grad_buf, hess_buf, prob: _Problem):
    """Same name and argument order as the reference (gauss_expected_grad_hess.jl:20-27): fills grad_buf (D) and
    hess_buf (D, D) in place and returns (logpi_avg, grad_buf, hess_buf).  `rng` provides the Philox key."""
    o = Objective(rng, RepGradELBO(1), q, prob)
    try:
        lp, g, H = o.gaussian_expectation_gradient_and_hessian(q, n_samples)
    finally:
        o.close()
    grad_buf[...] = g
    hess_buf[...] = H
    return lp, grad_buf, hess_buf


class HostUpdate:
    """Host-side `Optimisers.update!` + operator + averager (src/algorithms/common.jl:91-94) for callers that stay on
    the estimate_gradient! boundary with the parameters in host memory: Descent / Adam, ClipScale on the scale entries
    of a mean-field lambda, PolynomialAveraging.  Compiled (avi_host_update in the library), no device work."""

    def __init__(self, optimizer, operator, averager, lambda0, scale_offset: int):
        if optimizer.code not in (L.RULE_DESCENT, L.RULE_ADAM):
            raise ValueError("HostUpdate supports Descent and Adam")
        self.rule, self.op, self.avg = optimizer, operator, averager
        self.hyper = np.asarray(optimizer.hyper, np.float32)
        self.lam = np.ascontiguousarray(lambda0, np.float32).copy()
        self.m1, self.m2 = np.zeros_like(self.lam), np.zeros_like(self.lam)
        self.lam_avg = self.lam.copy()
        self.state = np.zeros(16, np.float32)
        self.scale_offset = int(scale_offset)
        self._ptr = _PtrCache()
        # the state arrays live as long as this object: their pointers are built once
        self._fixed = (self.rule.code, L.fptr(self.hyper), len(self.hyper), self.op.code,
                       C.c_float(getattr(self.op, "param", 0.0)), self.avg.code, C.c_float(getattr(self.avg, "param", 0.0)),
                       self.lam.size, self.scale_offset, L.fptr(self.lam))
        self._tail = (L.fptr(self.m1), L.fptr(self.m2), L.fptr(self.lam_avg), L.fptr(self.state))

    def update(self, grad):
        g = grad if (type(grad) is np.ndarray and grad.dtype == np.float32 and grad.flags.c_contiguous) else \
            np.ascontiguousarray(grad, np.float32)
        rc = L.lib.avi_host_update(*self._fixed, self._ptr(g), *self._tail)
        if rc != 0:
            raise AviError(rc, "avi_host_update: unsupported rule / operator or missing state arrays")
        return self.lam


class HostStep:
    """`step` (src/algorithms/common.jl:75-104) with the parameters in HOST memory: every call is one estimate_gradient!
    through the zero-copy boundary (host lambda in, host gradient out) followed by Optimisers.update! + operator + averager
    on the host (avi_hoststep_step).  The arrays `lam`, `grad`, `lam_avg` are owned here and bound to the handle once, so a
    call crosses the FFI with two pointers.  Same arithmetic as Objective.estimate_gradient + HostUpdate.update."""

    def __init__(self, obj: "Objective", optimizer, operator, averager, lambda0, scale_offset: int):
        self.obj = obj
        self.lam = np.ascontiguousarray(lambda0, np.float32).copy()
        self.grad = np.zeros_like(self.lam)
        self.lam_avg = self.lam.copy()
        hyper = np.asarray(optimizer.hyper, np.float32)
        h = L.vp()
        L.check(L.lib.avi_hoststep_create(obj.h, optimizer.code, L.fptr(hyper), len(hyper), operator.code,
                                          float(getattr(operator, "param", 0.0)), averager.code,
                                          float(getattr(averager, "param", 0.0)), int(scale_offset), L.fptr(self.lam),
                                          L.fptr(self.grad), L.fptr(self.lam_avg), C.byref(h)), obj.ctx.h)
        self.h = h
        self._v, self._e = C.c_float(), C.c_float()
        self._pv, self._pe = C.byref(self._v), C.byref(self._e)
        self._fn = L.lib.avi_hoststep_step

    def step(self):
        """-> (value, elbo); `lam`, `grad`, `lam_avg` are updated in place."""
        rc = self._fn(self.h, self._pv, self._pe)
        if rc:
            L.check(rc, self.obj.ctx.h)
        return self._v.value, self._e.value

    def timing(self):
        """wall-clock microseconds of the last call: (estimate_gradient!, host update, enqueue, wait for the flag) --
        the last two are parts of the first"""
        a, b, c, d = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        L.lib.avi_hoststep_timing(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, d.value

    def close(self):
        if self.h:
            L.lib.avi_hoststep_destroy(self.h)
            self.h = None


class _OptState:
    """The `state` NamedTuple of src/algorithms/common.jl:52-60, device-resident."""

    def __init__(self, alg, obj: Objective, q_init: MvLocationScale):
        self.alg, self.obj, self.q_template = alg, obj, q_init
        lam0 = q_init.destructure()
        rule = alg.optimizer
        hyper = np.asarray(rule.hyper, np.float32)
        h = L.vp()
        L.check(L.lib.avi_opt_create(obj.h, rule.code, L.fptr(hyper), len(hyper), alg.operator.code,
                                     alg.operator.param, alg.averager.code, alg.averager.param, L.fptr(lam0),
                                     len(lam0), C.byref(h)), obj.ctx.h)
        self.h = h
        self.sub_state = None

    @property
    def iteration(self) -> int:
        return int(L.lib.avi_opt_iteration(self.h))

    def params(self):
        P = self.obj.P
        lam, avg, grad = (np.empty(P, np.float32) for _ in range(3))
        L.check(L.lib.avi_opt_get(self.h, L.fptr(lam), L.fptr(avg), L.fptr(grad)), self.obj.ctx.h)
        return lam, avg, grad

    # avi_opt_steps in three phases (no host round trip between iterations; see include/avi.h)
    def steps_begin(self, capacity: int):
        L.check(L.lib.avi_opt_steps_begin(self.h, int(capacity)), self.obj.ctx.h)
        self._cap = int(capacity)

    def steps_enqueue(self, n: int = 1):
        L.check(L.lib.avi_opt_steps_enqueue(self.h, int(n)), self.obj.ctx.h)

    def steps_end(self):
        """-> (value slots, elbos, iterations completed); raises like `step` when the objective diverged."""
        vals, elbos, nd = np.empty(self._cap, np.float32), np.empty(self._cap, np.float32), C.c_int32()
        L.check(L.lib.avi_opt_steps_end(self.h, L.fptr(vals), L.fptr(elbos), C.byref(nd)), self.obj.ctx.h)
        return vals, elbos, nd.value

    def export_bytes(self) -> bytes:
        n = int(L.lib.avi_opt_state_nbytes(self.h))
        buf = C.create_string_buffer(n)
        L.check(L.lib.avi_opt_state_export(self.h, buf, n), self.obj.ctx.h)
        return buf.raw

    def import_bytes(self, b: bytes):
        L.check(L.lib.avi_opt_state_import(self.h, b, len(b)), self.obj.ctx.h)

    def close(self):
        if self.h:
            L.lib.avi_opt_destroy(self.h)
            self.h = None


def _raise_diverged(value):
    # src/algorithms/common.jl:83-89
    raise RuntimeError(f"The objective value is {value}. This indicates that the optimization run diverged.")


def optimize(rng, alg: _ParamSpaceSGD, max_iter: int, prob: _Problem, q_init: MvLocationScale, *,
             callback=None, state: _OptState = None, chunk: int = 4096):
    """optimize(rng, alg, max_iter, prob, q_init; callback, state) -> (q_avg, info, state)
    (src/optimize.jl:42-94 with init/step/output of src/algorithms/common.jl:40-120).

    Without a callback the iterations run on the device in chunks (one synchronisation per chunk);
    with one, `callback(rng=, iteration=, restructure=, params=, averaged_params=, gradient=, state=)`
    is called after every iteration and its dict return value is merged into that iteration's info."""
    objective = alg.objective
    subsampled = isinstance(objective, SubsampledObjective)
    if state is None:
        if isinstance(alg.operator, IdentityOperator):
            warnings.warn("IdentityOperator is used with a variational family <:MvLocationScale. Optimization can "
                          "easily fail under this combination due to singular scale matrices. Consider using the "
                          "operator `ClipScale` in the algorithm instead.")          # common.jl:42-46
        obj = Objective(rng, objective, q_init, prob)
        state = _OptState(alg, obj, q_init)
        if subsampled:
            state.sub_state = _sub_init(objective.subsampling, obj.key)   # subsampledobjective.jl:32 (pre-step state)
    obj = state.obj
    info = []
    done = 0
    vals = np.empty(max(1, min(chunk, max_iter)), np.float32)
    elbos = np.empty_like(vals)
    ndone = C.c_int32()
    while done < max_iter:
        n = 1 if callback is not None else min(chunk, max_iter - done)
        sub_infos = []
        if subsampled:
            sub = objective.subsampling
            batches = []
            st = state.sub_state
            for _ in range(n):
                b, st, si = _sub_step(sub, st, True)                       # subsampledobjective.jl:79
                batches.append(b); sub_infos.append(si)
            bsz = len(batches[0])
            if any(len(b) != bsz for b in batches):
                raise ValueError("minibatches of one chunk must have equal sizes")
            idx = np.ascontiguousarray(np.stack(batches), dtype=np.int32)
            L.check(L.lib.avi_opt_steps_subsampled(state.h, n, L.iptr(idx), bsz, L.fptr(vals), L.fptr(elbos),
                                                   C.byref(ndone)), obj.ctx.h)
        else:
            L.check(L.lib.avi_opt_steps(state.h, n, L.fptr(vals), L.fptr(elbos), C.byref(ndone)), obj.ctx.h)
        nd = ndone.value
        it0 = state.iteration - nd
        for k in range(nd):
            d = dict(iteration=it0 + k + 1, elbo=float(elbos[k]))
            if subsampled:
                d.update(sub_infos[k])
            info.append(d)
        if subsampled:
            st = state.sub_state
            for _ in range(nd):
                _, st, _ = _sub_step(objective.subsampling, st, True)
            state.sub_state = st
        if nd < n:
            _raise_diverged(float(vals[nd]))
        if callback is not None:
            lam, avg, grad = state.params()
            extra = callback(rng=rng, iteration=state.iteration, restructure=q_init.restructure, params=lam,
                             averaged_params=avg, gradient=grad, state=state)
            if extra:
                info[-1].update(extra)
        done += nd
    _, avg, _ = state.params()
    return q_init.restructure(avg), info, state


def estimate_objective(rng, alg_or_obj, q: MvLocationScale, prob: _Problem, *, n_samples: int = None, entropy=None):
    """estimate_objective(rng, alg, q, prob; n_samples, entropy) -> -ELBO estimate.
    For an algorithm this is always a fresh RepGradELBO with MonteCarloEntropy, ignoring subsampling
    (src/algorithms/common.jl:29-38); for an objective it is that objective's own estimate
    (repgradelbo.jl:112-118, scoregradelbo.jl:58-65); for a SubsampledObjective the mean of the inner objective's
    estimates over all length(subsampling) minibatches of one freshly shuffled epoch, short trailing batch included
    (subsampledobjective.jl:47-58).  Minibatch k draws its samples with key + 1 + k (the reference's rng stream moves
    on between batches); the target is back on its full data afterwards."""
    if isinstance(alg_or_obj, SubsampledObjective):
        sub, inner = alg_or_obj.subsampling, alg_or_obj.objective
        key = _key_from(rng)
        st = _sub_init(sub, key)                                           # :51
        n = n_samples or inner.n_samples
        spec = inner if entropy is None else RepGradELBO(inner.n_samples, entropy)
        o = Objective(0, spec if (spec.kind != L.REPGRAD or prob.capability >= 1) else ScoreGradELBO(spec.n_samples), q, prob)
        total = 0.0
        try:
            for k in range(len(sub)):                                      # :52
                batch, st, _ = _sub_step(sub, st)                          # :53
                prob.subsample(batch)                                      # :54
                total += o.estimate_objective((key + 1 + k) & 0xFFFFFFFFFFFFFFFF, q, n, kind=spec.kind,
                                              entropy=spec.entropy) / len(sub)   # :56
        finally:
            prob.subsample(None)
            o.close()
        return total
    if isinstance(alg_or_obj, _ParamSpaceSGD):
        spec = RepGradELBO(n_samples or alg_or_obj.objective.n_samples, entropy or MonteCarloEntropy())
    else:
        spec = alg_or_obj
        if entropy is not None:
            spec = RepGradELBO(spec.n_samples, entropy)
    if spec.kind == L.REPGRAD and prob.capability < 1:
        # forward-only: no gradient needed, evaluate through a ScoreGrad-kind handle
        handle_spec = ScoreGradELBO(spec.n_samples)
    else:
        handle_spec = spec
    o = Objective(0, handle_spec, q, prob)
    try:
        return o.estimate_objective(rng, q, n_samples or spec.n_samples, kind=spec.kind, entropy=spec.entropy)
    finally:
        o.close()
