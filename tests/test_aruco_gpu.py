"""GPU: the CUDA ArUco detector through the C-ABI vs the CPU restatement of aruco::MarkerDetector::detect.
Marker ids, candidate order and unrefined corners: bit-exact.  Refined (CORNER_LINES) corners: within 1e-4 px
(north_star's tolerance for sub-pixel corners; the float SVD accumulates in a different order on the device)."""
import os

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200.api import MarkerDetector

pytestmark = pytest.mark.gpu
TOL = 1e-4


def check_frame(det_out, img, dict_name):
    want = oracle.aruco_detect(img, dict_name)
    got_ids = np.array([m.id for m in det_out], np.int32)
    assert np.array_equal(got_ids, want["id"]), (got_ids.tolist(), want["id"].tolist())
    if len(want):
        got = np.array([m.corners for m in det_out], np.float32).reshape(-1, 8)
        assert np.abs(got - want["xy"]).max() <= TOL, np.abs(got - want["xy"]).max()
    return len(want)


@pytest.mark.parametrize("idx,w,h,dn", [(0, 640, 480, "ARUCO_MIP_25h7"), (1, 640, 480, "ARUCO_MIP_25h7"), (3, 640, 480, "ARUCO"),
                                        (4, 640, 480, "ARUCO_MIP_36h12"), (2, 1280, 720, "ARUCO_MIP_25h7"), (5, 960, 540, "ARUCO_MIP_25h7"),
                                        (6, 1920, 1080, "ARUCO_MIP_25h7")])
def test_detect_matches_oracle(built_lib, idx, w, h, dn):
    img = synth.make_frame(idx, w, h, markers=20, dict_name=dn)
    det = MarkerDetector(dn)
    n = check_frame(det.detect(img), img, dn)
    assert n >= 15
    det.close()


def test_stage_taps_bit_exact(built_lib):
    """candidates after prefilterCandidates (order + corners) and the decoded ids, against the oracle's stages"""
    img = synth.make_frame(0, markers=20)
    det = MarkerDetector("ARUCO_MIP_25h7")
    det.detect(img)
    counts, corners, ids = det.debug(0)
    o = oracle.aruco_stages(img)
    big = sum(1 for c in o["contours"] if len(c) > 70)
    assert counts[0] == big
    assert np.array_equal(corners, o["candidates"])
    det.close()


def test_golden_cv2_pipeline(built_lib, golden_dir):
    """vectors produced by driving the real cv2 primitives in the reference's order (tests/golden/cv2_aruco_pipeline.py)"""
    g = np.load(os.path.join(golden_dir, "aruco.npz"))
    for i, ((idx, w, h), dn) in enumerate(zip(g["cases"], g["dicts"])):
        img = synth.make_frame(int(idx), int(w), int(h), markers=20, dict_name=str(dn))
        det = MarkerDetector(str(dn))
        out = det.detect(img)
        assert np.array_equal(np.array([m.id for m in out], np.int32), g["ids_%d" % i])
        _, corners, _ = det.debug(0)
        assert np.array_equal(corners, g["candidates_%d" % i])
        got = np.array([m.corners for m in out], np.float32)
        assert np.abs(got - g["corners_%d" % i]).max() < 1e-3        # cv2 used LAPACK for the big SVDs (SURVEY A-12)
        det.close()


def test_batch_and_strides(built_lib):
    imgs = synth.make_batch(6, markers=20, first=20)
    det = MarkerDetector("ARUCO_MIP_25h7")
    markers, counts = det.detect_batch(imgs)
    for f in range(6):
        want = oracle.aruco_detect(imgs[f])
        assert counts[f] == len(want)
        assert np.array_equal(markers[f, :counts[f]]["id"], want["id"])
        assert np.abs(markers[f, :counts[f]]["xy"] - want["xy"]).max() <= TOL
    big = np.zeros((2, 500, 700), np.uint8)
    big[:, 7:487, 11:651] = imgs[:2]
    m2, c2 = det.detect_batch(big[:, 7:487, 11:651])
    assert np.array_equal(c2, counts[:2]) and np.array_equal(m2[0, :c2[0]]["id"], markers[0, :counts[0]]["id"])
    det.close()


def test_no_markers_and_edge_cases(built_lib):
    det = MarkerDetector("ARUCO_MIP_25h7")
    assert det.detect(synth.make_frame(9)) == []
    assert det.detect(np.full((480, 640), 128, np.uint8)) == []
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (240, 320)).astype(np.uint8)
    check_frame(det.detect(noise), noise, "ARUCO_MIP_25h7")
    with pytest.raises(Exception):
        MarkerDetector("NOT_A_DICTIONARY").detect(noise)
    det.close()


def test_duplicate_ids_keep_larger_perimeter(built_lib):
    """the same marker id planted twice: sort + de-dup must keep one, like the reference (markerdetector_impl.cpp:8159-8311)"""
    img, truth = synth.make_frame(12, markers=20, return_truth=True)
    cells = synth.marker_cells("ARUCO_MIP_25h7", truth[0][0])
    c = np.array([[300, 200], [340, 200], [340, 240], [300, 240]], np.float64)
    img2 = img.copy()
    synth._draw_marker(img2, cells, c)
    det = MarkerDetector("ARUCO_MIP_25h7")
    check_frame(det.detect(img2), img2, "ARUCO_MIP_25h7")
    det.close()
