"""CPU: the oracle's OpenCV-primitive restatements against golden vectors produced by the real cv2 4.13
(tests/golden/make_golden.py), and the extractor restatement against vectors produced by the reference's own
ORBextractor.cc (oracle/_ref).  Bit-exact everywhere."""
import os

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth


@pytest.fixture(scope="module")
def prim(golden_dir):
    return np.load(os.path.join(golden_dir, "primitives.npz"))


def test_synth_is_reproducible(prim):
    assert np.array_equal(synth.make_frame(1, 320, 240), prim["img"])


def test_resize_matches_cv2(prim):
    for i in range(4):
        want = prim["resize_%d" % i]
        got = oracle.resize_linear(prim["img"], want.shape[1], want.shape[0])
        assert np.array_equal(got, want)


def test_blur_matches_cv2(prim):
    assert np.array_equal(oracle.gaussian_blur7(prim["img"]), prim["blur"])


def test_border_matches_cv2(prim):
    assert np.array_equal(oracle.border_reflect101(prim["border_src"], 19), prim["border"])


def test_fast_matches_cv2(prim):
    img = prim["img"]
    for k, (y0, x0, hh, ww) in enumerate(prim["fast_rois"]):
        roi = np.ascontiguousarray(img[y0:y0 + hh, x0:x0 + ww])
        for thr in (20, 7):
            assert np.array_equal(oracle.fast_nms(roi, thr), prim["fast_%d_%d" % (k, thr)])


def test_fast_atan2_matches_cv2(prim):
    got = np.array([oracle.fast_atan2(y, x) for y, x in prim["atan2_in"]], np.float32)
    assert np.array_equal(got, prim["atan2_out"])


def test_extractor_matches_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "orb_ref.npz"))
    for i, (idx, w, h, nf) in enumerate(g["cases"]):
        img = synth.make_frame(int(idx), int(w), int(h))
        k, d = oracle.orb_extract(img, int(nf))
        assert len(k) == len(g["kps_%d" % i])
        assert np.array_equal(k, g["kps_%d" % i])
        assert np.array_equal(d, g["desc_%d" % i])


def test_level_geometry_matches_survey_table():
    # sizes and quotas at 640x480 / 1000 features
    lw, lh, q, sf = oracle.orb_levels(640, 480, 1000, 1.2, 8)
    assert list(lw) == [640, 533, 444, 370, 309, 257, 214, 179]
    assert list(lh) == [480, 400, 333, 278, 231, 193, 161, 134]
    assert list(q) == [217, 181, 151, 126, 105, 87, 73, 60]


def test_empty_and_flat_images():
    k, d = oracle.orb_extract(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = oracle.orb_extract(np.full((480, 640), 77, np.uint8))
    assert len(k) == 0
