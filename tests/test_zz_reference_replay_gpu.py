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
