"""CPU: the ArUco restatement (oracle/aruco_oracle.cpp + cvprim_aruco.h) against golden vectors made with the real
cv2 4.13 primitives (tests/golden/make_golden.py, cv2_aruco_pipeline.py) and against planted ground truth."""
import os

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "aruco.npz"))


def test_adaptive_threshold_and_half_resize(g):
    for bs in (5, 11, 15):
        assert np.array_equal(oracle.adaptive_threshold(g["warp_src"], bs), g["athr_%d" % bs])
    assert np.array_equal(oracle.resize_half(g["warp_src"]), g["half_even"])        # exact 1/2: 2x2 area mean
    assert np.array_equal(oracle.resize_half(g["half_odd_src"]), g["half_odd"])     # odd size: generic bilinear


def test_approx_poly_dp_and_convexity(g):
    o = 0; q = 0
    for n, m, cv in zip(g["approx_in_sizes"], g["approx_out_sizes"], g["approx_convex"]):
        pts = g["approx_in"][o:o + n]; want = g["approx_out"][q:q + m]
        got, convex = oracle.approx_poly(pts, n * 0.05)
        assert np.array_equal(got, want)
        if m >= 3:
            assert convex == bool(cv)
        o += n; q += m


def test_perspective_warp_otsu(g):
    dst = np.array([[0, 0], [34, 0], [34, 34], [0, 34]], np.float32)
    for quad, M, patch, ot in zip(g["warp_quads"], g["warp_M"], g["warp_out"], g["otsu"]):
        M2 = oracle.perspective_transform(quad, dst)
        assert np.array_equal(M2, M)                                  # bit-identical doubles
        got = oracle.warp_perspective(g["warp_src"], M, 35)
        assert np.array_equal(got, patch)
        assert oracle.otsu(patch) == int(ot)


def test_svd_solve_small_systems_bit_exact(g):
    for A, b, x in zip(g["svd_A"], g["svd_b"], g["svd_x"]):
        n = int(x[2])
        got = oracle.solve_svd(A[:n], b[:n])
        assert np.array_equal(got.view(np.uint32), x[:2].view(np.uint32))


def test_pipeline_stage_by_stage(g):
    for i, ((idx, w, h), dn) in enumerate(zip(g["cases"], g["dicts"])):
        img = synth.make_frame(int(idx), int(w), int(h), markers=20, dict_name=str(dn))
        o = oracle.aruco_stages(img, str(dn))
        tsum, ncont, npts = g["thres_sum_%d" % i]
        assert int(o["thres"].astype(np.uint64).sum()) == tsum
        assert len(o["contours"]) == ncont and sum(len(c) for c in o["contours"]) == npts
        assert np.array_equal(np.array([len(c) for c in o["contours"]], np.int32), g["contour_sizes_%d" % i])
        crc = np.array([int((c.astype(np.int64) * np.arange(1, 2 * len(c) + 1).reshape(-1, 2)).sum() % 1000003) for c in o["contours"]], np.int32)
        assert np.array_equal(crc, g["contour_crc_%d" % i])           # every point of every contour, in order
        assert np.array_equal(o["candidates"], g["candidates_%d" % i])
        assert np.array_equal(o["patches"], g["patches_%d" % i])
        assert np.array_equal(o["prerefine"], g["prerefine_%d" % i])
        assert np.array_equal(o["markers"]["id"], g["ids_%d" % i])
        # refined corners: cv2 routes the >=25-row SVDs to LAPACK, the oracle uses OpenCV's built-in Jacobi (SURVEY A-12)
        assert np.abs(o["markers"]["xy"].reshape(-1, 4, 2) - g["corners_%d" % i]).max() < 1e-3


def test_planted_markers_are_found(g):
    """known answer: every planted id is returned and the refined corners sit on the rendered ones"""
    for i, ((idx, w, h), dn) in enumerate(zip(g["cases"], g["dicts"])):
        img = synth.make_frame(int(idx), int(w), int(h), markers=20, dict_name=str(dn))
        m = oracle.aruco_detect(img, str(dn))
        truth = dict(zip(g["truth_ids_%d" % i].tolist(), g["truth_corners_%d" % i]))
        found = [int(x) for x in m["id"] if int(x) in truth]
        assert len(found) >= 18, (i, len(found))
        assert (np.diff(m["id"]) >= 0).all()
        for r in m:
            if int(r["id"]) in truth:
                assert np.abs(r["xy"].reshape(4, 2) - truth[int(r["id"])]).max() < 1.5


def test_no_markers_in_plain_frames():
    assert len(oracle.aruco_detect(synth.make_frame(9))) == 0
    assert len(oracle.aruco_detect(np.full((480, 640), 128, np.uint8))) == 0
