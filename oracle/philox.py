"""Counter-based standard-normal draws shared by the oracle and the CUDA path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws epsilon with ``rand(rng, Normal{T}(0,1), D, M)``
(src/families/location_scale.jl:76, :86), filling a D x M column-major matrix from a
Julia RNG stream.  That stream cannot be reproduced on a GPU, so both
sides use the same *pure function* instead:

    eps[i, m] at step t under key k  =  BoxMuller(Philox4x32-10(ctr=(i//4, m, t, stream), key=k))

Philox4x32-10 is the Random123 generator (Salmon et al., SC'11); the known-answer
vectors of that publication pin this restatement (tests/test_oracle_philox.py).

Uniform -> normal: u = ((x >> 9) + 0.5) * 2**-23 is exactly representable in fp32, so
the oracle (fp64 log/cos/sin, rounded once) and the GPU (fp32 logf/sincospif) start
from bit-identical uniforms and differ only by the last-ulp error of the fp32
transcendental functions.
"""

from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_EPS = 0        # epsilon draws of the variational family
STREAM_SHUFFLE = 1    # minibatch reshuffling (oracle/reshuffling.py)
STREAM_DATA = 2       # synthetic benchmark data (bench.py / tests)
STREAM_EPS_FACTORS = 3   # u_fact draws of the low-rank family (location_scale_low_rank.jl:84)


def philox4x32_10(ctr: np.ndarray, key) -> np.ndarray:
    """Philox4x32 with 10 rounds.

    ctr: uint32 array of shape (..., 4); key: pair of uint32.  Returns (..., 4) uint32.
    """
    c = np.asarray(ctr, dtype=np.uint64) & MASK32
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0 = int(key[0]) & 0xFFFFFFFF
    k1 = int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0          # 64-bit products of 32-bit operands
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def split_key(key: int):
    key = int(key) & 0xFFFFFFFFFFFFFFFF
    return key & 0xFFFFFFFF, key >> 32


def uniform23(x: np.ndarray) -> np.ndarray:
    """uint32 -> (0,1), exactly representable in fp32: ((x >> 9) + 0.5) * 2**-23."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) + 0.5) * (2.0 ** -23)


def box_muller(x: np.ndarray) -> np.ndarray:
    """(..., 4) uint32 -> (..., 4) float64 standard normals.

    Pairs (x0, x1) -> (n0, n1) and (x2, x3) -> (n2, n3):
        r = sqrt(-2 ln u_a), n_a = r cos(2 pi u_b), n_b = r sin(2 pi u_b).
    """
    u = uniform23(x)
    out = np.empty(u.shape, dtype=np.float64)
    for a in (0, 2):
        r = np.sqrt(-2.0 * np.log(u[..., a]))
        th = 2.0 * np.pi * u[..., a + 1]
        out[..., a] = r * np.cos(th)
        out[..., a + 1] = r * np.sin(th)
    return out


def normal_matrix(key: int, step: int, D: int, M: int, m0: int = 0,
                  stream: int = STREAM_EPS, dtype=np.float64) -> np.ndarray:
    """eps in R^{D x M}: column m is Monte-Carlo sample (m0 + m); entry i of sample m at
    optimisation step `step` is a pure function of (key, step, m0 + m, i).

    Counter layout: (c0, c1, c2, c3) = (i // 4, m, step & 0xffffffff, stream | (step >> 32) << 8).
    """
    nq = (D + 3) // 4
    q = np.arange(nq, dtype=np.uint64)[:, None]
    m = (np.arange(M, dtype=np.uint64) + np.uint64(m0))[None, :]
    ctr = np.empty((nq, M, 4), dtype=np.uint64)
    ctr[..., 0] = q
    ctr[..., 1] = m
    ctr[..., 2] = np.uint64(int(step) & 0xFFFFFFFF)
    ctr[..., 3] = np.uint64((int(stream) & 0xFF) | (((int(step) >> 32) & 0xFFFFFF) << 8))
    x = philox4x32_10(ctr, split_key(key))
    n = box_muller(x)                                  # (nq, M, 4)
    eps = np.transpose(n, (0, 2, 1)).reshape(nq * 4, M)[:D]
    return np.ascontiguousarray(eps).astype(dtype)


STREAM_EPS_B = 4      # second word block of the Student-t base draws (two variates per Philox block)


def _eps_words(key: int, step: int, D: int, M: int, m0: int, stream: int) -> np.ndarray:
    """The raw Philox words behind `normal_matrix`: (nq, M, 4) uint32, same counter layout."""
    nq = (D + 3) // 4
    ctr = np.empty((nq, M, 4), dtype=np.uint64)
    ctr[..., 0] = np.arange(nq, dtype=np.uint64)[:, None]
    ctr[..., 1] = (np.arange(M, dtype=np.uint64) + np.uint64(m0))[None, :]
    ctr[..., 2] = np.uint64(int(step) & 0xFFFFFFFF)
    ctr[..., 3] = np.uint64((int(stream) & 0xFF) | (((int(step) >> 32) & 0xFFFFFF) << 8))
    return philox4x32_10(ctr, split_key(key))


def laplace_matrix(key: int, step: int, D: int, M: int, m0: int = 0, dtype=np.float64) -> np.ndarray:
    """iid Laplace(0, 1) draws in R^{D x M} (base distribution of docs/src/families.md:88-101) by inversion of the CDF
    of the 23-bit uniform of every Philox word: p < 1/2 -> log(2 p), else -log(2 (1 - p)).  Word k of block (i // 4, m)
    is coordinate 4 (i // 4) + k, as for the normals."""
    p = uniform23(_eps_words(key, step, D, M, m0, STREAM_EPS))          # (nq, M, 4)
    u = np.where(p < 0.5, np.log(2.0 * p), -np.log(2.0 * (1.0 - p)))
    return np.ascontiguousarray(np.transpose(u, (0, 2, 1)).reshape(-1, M)[:D]).astype(dtype)


def student_t_matrix(key: int, step: int, D: int, M: int, nu: float, m0: int = 0, dtype=np.float64) -> np.ndarray:
    """iid TDist(nu) draws in R^{D x M} (docs/src/families.md:72-86) without rejection (Bailey 1994): with U, V
    uniform, sqrt(nu (U^(-2/nu) - 1)) cos(2 pi V) is t_nu.  The sine partner is not independent of the cosine one, so a
    pair of words gives ONE variate: coordinates 4q, 4q+1 come from words (0,1), (2,3) of the eps-stream block
    (q, m) and 4q+2, 4q+3 from the block of STREAM_EPS_B."""
    def half(stream):
        x = _eps_words(key, step, D, M, m0, stream)
        U, V = uniform23(x[..., 0::2]), uniform23(x[..., 1::2])          # (nq, M, 2)
        return np.sqrt(nu * np.expm1(-(2.0 / nu) * np.log(U))) * np.cos(2.0 * np.pi * V)
    t = np.concatenate([half(STREAM_EPS), half(STREAM_EPS_B)], axis=-1)   # (nq, M, 4)
    return np.ascontiguousarray(np.transpose(t, (0, 2, 1)).reshape(-1, M)[:D]).astype(dtype)


def uniform_u32(key: int, n: int, stream: int, offset: int = 0) -> np.ndarray:
    """n raw uint32 words from stream `stream` (counter = (j // 4 + offset, 0, 0, stream))."""
    nb = (n + 3) // 4
    ctr = np.zeros((nb, 4), dtype=np.uint64)
    ctr[:, 0] = (np.arange(nb, dtype=np.uint64) + np.uint64(offset)) & MASK32
    ctr[:, 1] = (np.arange(nb, dtype=np.uint64) + np.uint64(offset)) >> np.uint64(32)
    ctr[:, 3] = np.uint64(stream)
    return philox4x32_10(ctr, split_key(key)).reshape(-1)[:n]
