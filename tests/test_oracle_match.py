"""CPU: known-answer checks of the matcher restatement (the reference ships no tests or vectors for it)."""
import numpy as np

import oracle


def test_descriptor_distance_is_popcount():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (50, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (50, 32)).astype(np.uint8)
    for x, y in zip(a, b):
        assert oracle.descriptor_distance(x, y) == int(np.unpackbits(x ^ y).sum())
    assert oracle.descriptor_distance(a[0], a[0]) == 0
    assert oracle.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_self_match_is_identity():
    rng = np.random.default_rng(1)
    d = rng.integers(0, 256, (200, 32)).astype(np.uint8)
    ang = rng.uniform(0, 360, 200).astype(np.float32)
    n, m = oracle.search_by_bow_bf(d, ang, d, ang, 0.7, True)
    assert n == 200 and np.array_equal(m, np.arange(200))


def test_rotation_histogram_keeps_three_bins():
    rng = np.random.default_rng(2)
    d = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    ang = rng.uniform(0, 360, 300).astype(np.float32)
    fang = ang.copy()
    fang[:100] = (ang[:100] + 90) % 360     # a second consistent rotation cluster
    fang[100:110] = (ang[100:110] + 200) % 360
    fang[110:113] = (ang[110:113] + 300) % 360   # a fourth, small cluster must be dropped
    n, m = oracle.search_by_bow_bf(d, ang, d, fang, 0.7, True)
    # bins: 187 / 100 / 10 / 3 matches.  ComputeThreeMaxima drops the third bin too because 10 < 0.1*187
    # (ORBmatcher.cc:1642-1645), so 13 matches are removed.
    assert n == 287 and (m[100:113] == -1).all() and (m[:100] >= 0).all() and (m[113:] >= 0).all()
    n2, m2 = oracle.search_by_bow_bf(d, ang, d, fang, 0.7, False)
    assert n2 == 300


def _medoid_numpy(d):
    """MapPoint.cc:271-331 by definition: sorted row, element (int)(0.5 * (N - 1)), first minimum"""
    if len(d) == 0:
        return -1
    dist = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(axis=2)
    med = np.sort(dist, axis=1)[:, int(0.5 * (len(d) - 1))]
    return int(np.argmin(med))


def test_distinctive_descriptor_is_least_median():
    import ctypes as C
    rng = np.random.default_rng(5)
    f = oracle.lib().oracle_distinctive_descriptor
    for n in (0, 1, 2, 3, 4, 7, 32, 33, 90):
        base = rng.integers(0, 256, (3, 32)).astype(np.uint8)
        d = np.ascontiguousarray(base[rng.integers(0, 3, n)] ^ np.packbits(rng.integers(0, 100, (n, 256)) < 6, axis=1)) if n else np.zeros((0, 32), np.uint8)
        assert f(d.ctypes.data_as(C.c_void_p), n) == _medoid_numpy(d), n
    # all identical: every median is 0, the first observation wins; two observations: median = d[0] of {0, x} = 0 for both -> index 0
    d = np.repeat(rng.integers(0, 256, (1, 32)).astype(np.uint8), 9, axis=0)
    assert f(d.ctypes.data_as(C.c_void_p), 9) == 0
