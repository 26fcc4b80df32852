"""GPU (runs last): the CUDA detector against the answers of the reference's OWN marker detector (tests/golden/aruco_ref.npz, written from
oracle/_ref/libref_aruco.so = Thirdparty/aruco/aruco compiled unmodified): same markers in the same order, same ids, refined corners within 1e-4 px."""
import os

import numpy as np
import pytest

import aruco_ref_cases as ac
from orb_slam2_aruco_b200.api import MarkerDetector

pytestmark = pytest.mark.gpu


def test_cuda_detector_replays_the_reference(built_lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "aruco_ref.npz"))
    total = 0
    for j, case in enumerate(ac.CASES):
        det = MarkerDetector(case["dict"])
        got = det.detect(ac.frame(case))
        det.close()
        ids = np.array([m.id for m in got], np.int32)
        assert np.array_equal(ids, g["c%d.id" % j]), (case, ids.tolist(), g["c%d.id" % j].tolist())
        if len(ids):
            xy = np.array([m.corners for m in got], np.float32).reshape(-1, 8)
            assert np.abs(xy - g["c%d.xy" % j]).max() <= 1e-4, (case, np.abs(xy - g["c%d.xy" % j]).max())
        total += len(ids)
    assert total > 100


def test_keyframe_searches_with_distorted_image_bounds(built_lib, monkeypatch):
    """non-integer image bounds: KeyFrame::IsInImage and the origin of KeyFrame::GetFeaturesInArea use the keyframe's int copies of the bounds
    (b200_keyframe_features_in_area); the CUDA path against the oracle, which equals the reference's own ORBmatcher.cc on these inputs (CPU test)"""
    import match_cases as mc
    import match_cases2 as m2
    import oracle
    from orb_slam2_aruco_b200.api import ORBmatcher
    B = np.array([-3.7, 643.6, -2.4, 482.9], np.float32)
    monkeypatch.setattr(m2, "BOUNDS", B)
    O, M = oracle.lib(), ORBmatcher(0.6, True)
    c = m2.keyframe_points_inputs(seed=14)
    want = m2.run_fuse(O, "oracle", c, 3.0)
    got = M.Fuse(c["k"], c["d"], B, mc.CAM4, c["T"], c["held_state"], c["held_nobs"], c["mp_state"], c["mp_pos"], c["mp_normal"], c["mp_desc"], c["mp_minmax"],
                 c["mp_nobs"], 3.0)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]) and got[0] > 200
    c = m2.keyframe_points_inputs(seed=16, sim3=True)
    want = m2.run_loop(O, "oracle", c, 10)
    st = np.where(c["mp_state"] == 2, 2, 1).astype(np.uint8)
    got = M.SearchByProjectionLoop(c["k"], c["d"], B, mc.CAM4, c["T"], st, c["mp_pos"], c["mp_normal"], c["mp_desc"], c["mp_minmax"], m2.loop_matched(c), 10)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[0] > 100
    c = m2.sim3_inputs(seed=18)
    want = m2.run_sim3(O, "oracle", c, 7.5)
    got = M.SearchBySim3(c["k1"], c["d1"], c["T1"], c["st1"], c["p1"], c["d1"], c["mm1"], c["k2"], c["d2"], c["T2"], c["st2"], c["p2"], c["d2"], c["mm2"],
                         B, mc.CAM4, c["m12"], float(c["s12"]), c["R12"], c["t12"], 7.5)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[0] > 100
