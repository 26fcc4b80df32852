"""CPU: the IPPE pose oracle (oracle/ippe_oracle.cpp, restating Thirdparty/aruco/aruco/ippe.cpp) against golden vectors from
cv2.solvePnPGeneric(SOLVEPNP_IPPE) (tests/golden/make_ippe_golden.py) and against the ground-truth poses."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

POSE_TOL = 1e-6          # rvec / tvec against OpenCV's implementation of the same algorithm (double arithmetic, different op order)


def solve(corners, msize, cam):
    out = np.zeros(14)
    oracle.lib().oracle_ippe_marker_pose(np.ascontiguousarray(corners, np.float32).ctypes.data_as(C.c_void_p), C.c_float(msize),
                                         np.ascontiguousarray(cam, np.float64).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ippe.npz"))


def test_matches_cv2_ippe(g):
    worst = 0.0
    for c, cam, want in zip(g["corners"], g["cams"], g["poses"]):
        got = solve(c, float(g["msize"]), cam)
        assert got[6] <= got[13]                                            # sorted by reprojection error
        worst = max(worst, np.abs(got[[0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12]] - want[[0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12]]).max())
        assert abs(got[6] - want[6]) <= 1e-4 * max(1, want[6]) and abs(got[13] - want[13]) <= 1e-4 * max(1, want[13])
    assert worst <= POSE_TOL, worst


def rodrigues(r):
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def test_recovers_ground_truth(g):
    """0.3 px corner noise: one of the two poses is the true one (rotation within a few degrees, translation within 5 % of the depth);
    mostly it is the one with the smaller reprojection error"""
    first = 0
    for c, cam, tr in zip(g["corners"], g["cams"], g["truth"]):
        got = solve(c, float(g["msize"]), cam)
        Rt = rodrigues(tr[:3])
        ang = [np.degrees(np.arccos(np.clip((np.trace(rodrigues(got[o:o + 3]).T @ Rt) - 1) / 2, -1, 1))) for o in (0, 7)]
        dt = [np.abs(got[o:o + 3] - tr[3:6]).max() for o in (3, 10)]
        best = int(np.argmin(ang))
        assert ang[best] < 12 and dt[best] < 0.05 * tr[5], (ang, dt)
        first += best == 0
    assert first > 0.7 * len(g["truth"])


def test_frontal_marker_known_answer():
    """a fronto-parallel 0.2 m square at z = 1 m, fx = fy = 500, object axes aligned with the image axes: R = I, t = (0, 0, 1)"""
    cam = [500, 500, 320, 240, 0, 0, 0, 0, 0]
    corners = np.array([[270, 290], [370, 290], [370, 190], [270, 190]], np.float32)
    got = solve(corners, 0.2, cam)
    assert np.allclose(got[3:6], [0, 0, 1], atol=1e-9)
    assert np.allclose(got[0:3], 0, atol=1e-7)
    assert got[6] < 1e-4
