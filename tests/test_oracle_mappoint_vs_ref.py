"""CPU: MapPoint members against the reference's OWN src/MapPoint.cc (+ include/MapPoint.h, src/ORBmatcher.cc) compiled unmodified into
oracle/_ref/libref_mappoint.so with stand-in KeyFrame / Frame / Map:
 * ComputeDistinctiveDescriptors (src/MapPoint.cc:271-331) vs oracle_distinctive_descriptor (the checker of k_distinctive / b200_distinctive_descriptors_host);
 * PredictScale / GetMinDistanceInvariance / GetMaxDistanceInvariance (:391-435) vs the threshold table the device projection uses (kfgeom.level_thresholds)
   and the libm evaluation of the numpy model.
Golden replay everywhere (tests/golden/mappoint_ref.npz), live where oracle/_ref exists."""
import ctypes as C
import os

import numpy as np
import pytest

import match_cases2 as m2
import oracle
from orb_slam2_aruco_b200 import kfgeom

vp = C.c_void_p
HERE = os.path.dirname(os.path.abspath(__file__))


def P(a):
    return a.ctypes.data_as(vp)


def descriptor_sets(seed=21, n_sets=60):
    """observation sets of 1..40 descriptors: noisy copies of a centre (so that medians tie often), duplicates, and pure noise"""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(n_sets):
        n = int(rng.integers(1, 41))
        centre = rng.integers(0, 256, 32).astype(np.uint8)
        d = np.repeat(centre[None], n, 0) ^ np.packbits(rng.integers(0, 100, (n, 256)) < rng.choice([1, 3, 10, 50]), axis=1)
        if s % 7 == 0 and n > 3:
            d[n // 2] = d[0]; d[n - 1] = d[1]
        out.append(np.ascontiguousarray(d))
    return out


def ratios(seed=22):
    rng = np.random.default_rng(seed)
    thr = kfgeom.level_thresholds(1.2, 8)
    near = np.concatenate([(t.view(np.uint32) + np.arange(-40, 41).astype(np.uint32)).view(np.float32) for t in thr.reshape(-1, 1)])
    return np.concatenate([np.exp(rng.uniform(-1.5, 2.5, 3000)).astype(np.float32), near])


def ref_levels(R, maxd, dists):
    lv = np.zeros(len(dists), np.int32); inv = np.zeros(2, np.float32)
    R.ref_predict_scale(C.c_float(maxd), P(dists), len(dists), C.c_float(1.2), 8, P(lv), P(inv))
    return lv, inv


@pytest.fixture(scope="module")
def golden(golden_dir):
    path = os.path.join(golden_dir, "mappoint_ref.npz")
    if not os.path.exists(path) and oracle.ref_mappoint() is not None:
        write_golden(path)
    return np.load(path)


def write_golden(path):
    """tests/golden/mappoint_ref.npz = the reference's answers for the seeded inputs above; delete the file and run this test where oracle/_ref exists
    to regenerate it"""
    R = oracle.ref_mappoint()
    out = {}
    best = []
    for d in descriptor_sets():
        o = np.zeros(32, np.uint8)
        R.ref_distinctive_descriptor(P(d), None, len(d), P(o))
        best.append(o)
    out["distinctive"] = np.stack(best)
    maxd = np.float32(7.3)
    dists = (maxd / ratios()).astype(np.float32)
    lv, inv = ref_levels(R, float(maxd), dists)
    out["levels"] = lv; out["inv"] = inv
    np.savez_compressed(path, **out)


def test_distinctive_descriptor_replays_the_reference(golden):
    f = oracle.lib().oracle_distinctive_descriptor
    for d, want in zip(descriptor_sets(), golden["distinctive"]):
        i = f(P(d), len(d))
        assert 0 <= i < len(d) and np.array_equal(d[i], want)


def test_predict_scale_replays_the_reference(golden):
    maxd = np.float32(7.3)
    r = ratios()
    dists = (maxd / r).astype(np.float32)
    ratio = (maxd / dists).astype(np.float32)                      # what PredictScale forms: mfMaxDistance / currentDist
    thr = kfgeom.level_thresholds(1.2, 8)
    assert np.array_equal((ratio[:, None] > thr[None, :]).sum(1), golden["levels"])                      # the device's threshold table
    assert np.array_equal(m2.predict_scale(np.full(len(dists), maxd), dists, kfgeom.pyramid()[3], 8), golden["levels"])   # libm, element by element
    assert golden["inv"][1] == np.float32(1.2) * maxd and len(set(golden["levels"].tolist())) == 8


@pytest.mark.skipif(oracle.ref_mappoint() is None, reason="oracle/_ref/libref_mappoint.so not built (needs /root/reference)")
def test_live_reference():
    R = oracle.ref_mappoint()
    f = oracle.lib().oracle_distinctive_descriptor
    for d in descriptor_sets(seed=77, n_sets=150):
        o = np.zeros(32, np.uint8)
        used = R.ref_distinctive_descriptor(P(d), None, len(d), P(o))
        assert used == len(d) and np.array_equal(d[f(P(d), len(d))], o)
    # bad keyframes are skipped (src/MapPoint.cc:296-297): the product's adapters drop them before the call
    d = descriptor_sets(seed=78, n_sets=1)[0]
    while len(d) < 6:
        d = np.concatenate([d, d])
    bad = np.zeros(len(d), np.uint8); bad[1::3] = 1
    o = np.zeros(32, np.uint8)
    R.ref_distinctive_descriptor(P(d), P(bad), len(d), P(o))
    keep = np.ascontiguousarray(d[bad == 0])
    assert np.array_equal(keep[f(P(keep), len(keep))], o)
    thr = kfgeom.level_thresholds(1.2, 8)
    for maxd in (np.float32(7.3), np.float32(0.91), np.float32(123.456)):
        dists = (maxd / ratios(seed=5)).astype(np.float32)
        lv, inv = ref_levels(R, float(maxd), dists)
        ratio = (maxd / dists).astype(np.float32)
        assert np.array_equal((ratio[:, None] > thr[None, :]).sum(1), lv)
        assert inv[1] == np.float32(1.2) * maxd
        sf = kfgeom.pyramid()[0]
        assert inv[0] == np.float32(0.8) * np.float32(maxd / sf[7])                                        # mfMinDistance = mfMaxDistance / mvScaleFactors[nLevels-1]


@pytest.mark.skipif(oracle.ref_mappoint() is None, reason="oracle/_ref/libref_mappoint.so not built (needs /root/reference)")
def test_golden_file_is_current(golden, tmp_path):
    path = os.path.join(str(tmp_path), "g.npz")
    write_golden(path)
    g = np.load(path)
    assert all(np.array_equal(g[k], golden[k]) for k in g.files)


@pytest.mark.skipif(oracle.ref_mappoint() is None, reason="oracle/_ref/libref_mappoint.so not built (needs /root/reference)")
def test_fuse_outcome_on_real_mappoints(monkeypatch):
    """ORBmatcher::Fuse run by the reference on REAL MapPoint objects (Replace re-points keyframe slots, AddObservation counts) against the mirror's FuseReal:
    final holder of every keyframe feature, bad flags and observation counts.  The mirror's device search is answered by the CPU oracle here."""
    import match_cases as mc
    from test_oracle_match2_vs_ref import _OracleSearches
    R = oracle.ref_mappoint()
    for seed, th in ((4, 3.0), (24, 3.0), (25, 6.0)):
        c = m2.keyframe_points_inputs(seed=seed)
        c["mp_state"] = np.where(c["mp_state"] == 3, 1, c["mp_state"]).astype(np.uint8)      # the scene's "already observed" points name no real slot
        n, m = len(c["k"]), len(c["mp_desc"])
        # several points per keypoint, so that slots are hit more than once in one call
        slot = np.zeros(n, np.int32); mb = np.zeros(m, np.uint8); hb = np.zeros(n, np.uint8); no = np.zeros(m, np.int32)
        nf = R.ref_fuse_real(P(c["k"]), P(c["d"]), n, P(mc.BOUNDS), P(mc.CAM4), P(c["T"]), P(c["held_state"]), P(c["held_nobs"]), m, P(c["mp_state"]),
                             P(c["mp_pos"]), P(c["mp_normal"]), P(c["mp_desc"]), P(c["mp_minmax"]), P(c["mp_nobs"]), C.c_float(th), P(slot), P(mb), P(hb), P(no))
        got = _OracleSearches(0.6, True).FuseReal(c["k"], c["d"], mc.BOUNDS, mc.CAM4, c["T"], c["held_state"], c["held_nobs"], c["mp_state"], c["mp_pos"],
                                                  c["mp_normal"], c["mp_desc"], c["mp_minmax"], c["mp_nobs"], th)
        assert got["n"] == nf and nf > 200
        assert np.array_equal(got["slot"], slot) and np.array_equal(got["mp_bad"], mb) and np.array_equal(got["held_bad"], hb)
        live = c["mp_state"] > 0
        assert np.array_equal(got["mp_nobs"][live], no[live])
        assert (slot >= 0).sum() > (c["held_state"] > 0).sum() and (slot[(slot >= 0) & (slot < 1000000)] >= 0).sum() > 100
