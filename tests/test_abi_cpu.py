"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/avi.h
declares, fails loudly without a GPU, and its host-side integer path (the minibatch permutation) is
bit-exact against the oracle.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import reshuffling as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "avi.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(avi_[a-z0-9_]+)\s*\(", txt))
    return sorted(n for n in names if not n.endswith("_fn"))


def test_library_exports_every_declared_symbol(avi):
    lib = C.CDLL(avi._lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 45
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # ... and the ctypes mirror binds every one of them with a signature
    unbound = [s for s in syms if s not in avi._lib.SIGNATURES]
    assert not unbound, unbound
    assert avi._lib.lib.avi_version() == 100


def test_no_cpu_fallback(avi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(avi.AviError, match="no CUDA device"):
        avi.Context(0)


@pytest.mark.parametrize("n", [0, 1, 2, 5, 8, 100, 1001])
@pytest.mark.parametrize("key,k", [(1, 0), (0x38BEF07CF9CC549D, 3), (2 ** 64 - 1, 2 ** 32 - 1)])
def test_shuffle_is_bit_exact_against_oracle(avi, n, key, k):
    """avi_shuffle == Random.shuffle replacement of src/reshuffling.jl:29 (oracle/reshuffling.py)."""
    perm = np.arange(n, dtype=np.int32)
    assert avi._lib.lib.avi_shuffle(key, k, n, avi._lib.iptr(perm)) == 0
    assert np.array_equal(perm, R.philox_shuffle(np.arange(n), key, k))


@pytest.mark.parametrize("n,bs", [(8, 3), (8, 4), (10, 1), (7, 7), (50, 8)])
def test_reshuffling_state_machine_matches_oracle(avi, n, bs):
    """Host mirror of ReshufflingBatchSubsampling (src/reshuffling.jl:38-60) vs the oracle, incl. the
    drop-trailing swap and the (epoch, step) info."""
    from advancedvi_jl_b200.api import _sub_init, _sub_step
    sub, subo = avi.ReshufflingBatchSubsampling(np.arange(n), bs), R.ReshufflingBatchSubsampling(np.arange(n), bs)
    assert len(sub) == len(subo)
    for drop in (False, True):
        st, sto = _sub_init(sub, 9), R.sub_init(subo, 9)
        for _ in range(3 * len(sub) + 2):
            b, st, info = _sub_step(sub, st, drop)
            bo, sto, infoo = R.sub_step(subo, sto, drop)
            assert np.array_equal(b, bo) and info == infoo


def test_family_container_and_eltype(avi):
    """Diagonal destructure has length 2d and round-trips (test/families/location_scale.jl:146-155);
    Float64 is rejected (the reference supports both, the B200 path is Float32 only)."""
    d = 5
    q = avi.MeanFieldGaussian(np.arange(d, dtype=np.float32), np.arange(1, d + 1, dtype=np.float32))
    lam = q.destructure()
    assert lam.shape == (2 * d,) and lam.dtype == np.float32
    q2 = q.restructure(lam)
    assert np.array_equal(q2.location, q.location) and np.array_equal(q2.scale, q.scale)
    Lm = np.tril(np.ones((d, d), np.float32))
    qf = avi.FullRankGaussian(np.zeros(d, np.float32), Lm)
    assert qf.destructure().shape == (d + d * d,)
    assert np.array_equal(qf.restructure(qf.destructure()).scale, Lm)
    with pytest.raises(TypeError, match="Float32"):
        avi.MeanFieldGaussian(np.zeros(d), np.ones(d))
    with pytest.raises(ValueError):
        avi.FullRankGaussian(np.zeros(d, np.float32), np.ones((d, d), np.float32))


def test_algorithm_constructors(avi):
    """constructors.jl:58-77, :136-157, :213-231: defaults and the entropy whitelist."""
    a = avi.KLMinRepGradDescent()
    assert isinstance(a.optimizer, avi.DoWG) and isinstance(a.averager, avi.PolynomialAveraging)
    assert isinstance(a.operator, avi.IdentityOperator) and a.objective.n_samples == 1
    assert isinstance(a.objective.entropy, avi.ClosedFormEntropy)
    with pytest.raises(ValueError):
        avi.KLMinRepGradDescent(entropy=avi.ClosedFormEntropyZeroGradient())
    p = avi.KLMinRepGradProxDescent()
    assert isinstance(p.operator, avi.ProximalLocationScaleEntropy)
    assert isinstance(p.objective.entropy, avi.ClosedFormEntropyZeroGradient)
    s = avi.KLMinScoreGradDescent(n_samples=7, subsampling=avi.ReshufflingBatchSubsampling(np.arange(10), 2))
    assert isinstance(s.objective, avi.SubsampledObjective) and s.objective.n_samples == 7
    assert avi.ADVI is avi.KLMinRepGradDescent and avi.BBVI is avi.KLMinScoreGradDescent


def test_host_update_matches_oracle_adam_clip_polyavg():
    """avi_host_update (host-side Optimisers.update! + ClipScale + PolynomialAveraging, no device work) against the
    oracle's restatement of src/algorithms/common.jl:91-94 over 50 steps in fp32."""
    import numpy as np
    import advancedvi_jl_b200 as avi
    from oracle import optim as Oo
    D = 37
    rng = np.random.default_rng(0)
    lam0 = np.concatenate([rng.normal(size=D), np.abs(rng.normal(size=D)) + 0.1]).astype(np.float32)
    for rule_a, rule_o in ((avi.Adam(1e-2), Oo.Adam(1e-2)), (avi.Descent(0.05), Oo.Descent(0.05))):
        hu = avi.HostUpdate(rule_a, avi.ClipScale(0.3), avi.PolynomialAveraging(8), lam0, scale_offset=D)
        x = lam0.copy(); st = rule_o.init(x); avg_o = Oo.PolynomialAveraging(8); ast = avg_o.init(x)
        clip = Oo.ClipScale(0.3)
        for t in range(50):
            g = rng.normal(size=2 * D).astype(np.float32)
            hu.update(g)
            st, dx = rule_o.apply(st, x, g)
            x = (x - dx).astype(np.float32)
            x[D:] = np.maximum(x[D:], np.float32(0.3))        # clip_scale.jl:18-29 on a mean-field lambda
            ast = avg_o.apply(ast, x)
        assert np.allclose(hu.lam, x, rtol=2e-6, atol=1e-7)
        assert np.allclose(hu.lam_avg, avg_o.value(ast), rtol=1e-5, atol=1e-6)
        assert (hu.lam[D:] >= 0.3).all()


def test_lowrank_host_container_matches_oracle_layout():
    """The host-side MvLocationScaleLowRank container flattens exactly like the oracle's restatement of the Functors
    order (location, scale_diag, scale_factors column-major; location_scale_low_rank.jl:26) and round-trips."""
    import numpy as np
    import advancedvi_jl_b200 as avi
    from oracle import family as F
    d, r = 7, 3
    rng = np.random.default_rng(1)
    mu, sd, U = (rng.normal(size=d).astype(np.float32), (np.abs(rng.normal(size=d)) + 0.1).astype(np.float32),
                 rng.normal(size=(d, r)).astype(np.float32))
    q, qo = avi.LowRankGaussian(mu, sd, U), F.LowRankGaussian(mu, sd, U)
    lam = q.destructure()
    assert lam.dtype == np.float32 and lam.shape == (2 * d + d * r,)
    assert np.array_equal(lam, qo.destructure().astype(np.float32))
    q2 = q.restructure(lam)
    assert np.array_equal(q2.location, mu) and np.array_equal(q2.scale_diag, sd) and np.array_equal(q2.scale_factors, U)
    assert q.rank == r and len(q) == d and np.allclose(q.cov(), qo.cov(), rtol=1e-6)
    import pytest
    with pytest.raises(TypeError):
        avi.LowRankGaussian(mu.astype(np.float64), sd, U)
    with pytest.raises(ValueError):
        avi.LowRankGaussian(mu, sd[:-1], U)


@pytest.mark.parametrize("base,param", [(0, 0.0), (1, 0.0), (2, 0.7), (2, 2.5), (2, 5.0), (2, 30.0), (2, 1000.0)])
def test_base_distribution_constants_match_scipy(base, param):
    """entropy(dist) and the log normaliser the kernels use for MvLocationScale(location, scale, dist)
    (src/families/location_scale.jl:52-63; docs/src/families.md:72-101), host arithmetic of csrc/base_dist.cuh
    (lgamma + an asymptotic digamma) against scipy: 2e-7 relative (they are stored as float32)."""
    import ctypes as C
    from scipy import stats
    from advancedvi_jl_b200 import _lib as L
    dist = [stats.norm(), stats.laplace(), stats.t(param)][base]
    h, c = C.c_float(), C.c_float()
    assert L.lib.avi_base_constants(base, param, C.byref(h), C.byref(c)) == 0
    assert abs(h.value - dist.entropy()) <= 2e-7 * max(1.0, abs(dist.entropy()))
    assert abs(c.value - dist.logpdf(0.0)) <= 2e-7 * max(1.0, abs(dist.logpdf(0.0)))
    assert L.lib.avi_base_constants(2, 0.0, C.byref(h), C.byref(c)) != 0       # nu must be positive
    assert L.lib.avi_base_constants(7, 1.0, C.byref(h), C.byref(c)) != 0       # unknown base


def test_check_indices_matches_a_plain_scan():
    """avi_check_indices (the range check in front of avi_opt_steps_subsampled): same verdict and same first offending
    position as the obvious loop, for clean arrays, one bad entry anywhere (also across the 4096-entry block edges),
    negative entries, and the empty array."""
    import ctypes as C
    from advancedvi_jl_b200 import _lib as L
    rng = np.random.default_rng(0)
    rows = 1000
    bad = C.c_int64()

    def check(a):
        a = np.ascontiguousarray(a, np.int32)
        rc = L.lib.avi_check_indices(a.ctypes.data_as(C.POINTER(C.c_int32)), a.size, rows, C.byref(bad))
        wrong = np.flatnonzero((a < 0) | (a >= rows))
        assert (rc == 0) == (wrong.size == 0)
        assert bad.value == (wrong[0] if wrong.size else -1)
    check(np.empty(0, np.int32))
    for n in (1, 5, 4095, 4096, 4097, 3 * 4096 + 17):
        a = rng.integers(0, rows, n)
        check(a)
        for pos in {0, n - 1, n // 2, min(n - 1, 4095), min(n - 1, 4096)}:
            for v in (rows, -1, 2 ** 31 - 1, -2 ** 31):
                b = a.copy(); b[pos] = v
                check(b)
        if n > 10:
            b = a.copy(); b[[3, n - 2]] = [-5, rows + 5]
            check(b)
    assert L.lib.avi_check_indices(None, 3, rows, C.byref(bad)) != 0
