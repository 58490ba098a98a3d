"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) + normal-draw sanity."""
import numpy as np

from oracle import philox as p

KAT = [  # (counter, key, expected) from the Random123 distribution's known-answer file
    ([0, 0, 0, 0], (0, 0), [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, (0xFFFFFFFF, 0xFFFFFFFF), [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], (0xA4093822, 0x299F31D0),
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_philox_known_answers():
    for ctr, key, exp in KAT:
        got = p.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert [int(v) for v in got] == exp


def test_uniform23_exact_in_fp32():
    x = np.array([0, 1, 511, 512, 0xFFFFFFFF, 0x80000000], dtype=np.uint32)
    u = p.uniform23(x)
    assert np.all(u > 0) and np.all(u < 1)
    assert np.array_equal(u.astype(np.float32).astype(np.float64), u)


def test_normal_matrix_is_pure_function_of_index():
    a = p.normal_matrix(7, 3, 10, 8)
    b = p.normal_matrix(7, 3, 10, 4, m0=4)      # samples 4..7 drawn as a shard
    assert np.array_equal(a[:, 4:], b)
    c = p.normal_matrix(7, 3, 6, 8)             # fewer coordinates: same leading rows
    assert np.array_equal(a[:6], c)
    assert not np.array_equal(a, p.normal_matrix(7, 4, 10, 8))
    assert not np.array_equal(a, p.normal_matrix(8, 3, 10, 8))


def test_normal_moments():
    e = p.normal_matrix(11, 0, 64, 20000)
    assert abs(e.mean()) < 5e-3
    assert abs(e.var() - 1) < 1e-2
    assert abs(np.mean(e ** 4) - 3) < 0.05
    assert abs(np.corrcoef(e[0], e[1])[0, 1]) < 0.03
