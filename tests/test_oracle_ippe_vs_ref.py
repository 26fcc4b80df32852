"""CPU: the marker pose oracle (oracle/ippe_oracle.cpp) against the reference's OWN pose solver - Thirdparty/aruco/aruco/ippe.cpp (IPPE::PoseSolver,
aruco::solvePnP), compiled unmodified on oracle/ippeshim into oracle/_ref/libref_ippe.so and driven like aruco::Marker::calculateExtrinsics
(marker.cpp:322-343) drives it.
 * replay of tests/golden/ippe_ref.npz (the reference's answers, written by tests/golden/make_ippe_ref_golden.py) - runs everywhere;
 * live on fresh markers where oracle/_ref exists (the dev container)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

POSE_TOL = 1e-7      # both are double pipelines over the same statements; libm calls and the 3 x 3 eigen solver differ in the last bits
ERR_TOL = 1e-6       # reprojection errors are float sums of float differences
IDX = [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12]


def solve(corners, msize, cam):
    out = np.zeros(14)
    oracle.lib().oracle_ippe_marker_pose(np.ascontiguousarray(corners, np.float32).ctypes.data_as(C.c_void_p), C.c_float(float(msize)),
                                         np.ascontiguousarray(cam, np.float64).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def rodrigues(r):
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def test_oracle_replays_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ippe_ref.npz"))
    assert len(g["corners"]) > 450
    worst = 0.0
    for c, cam, size, want, T, errs in zip(g["corners"], g["cams"], g["sizes"], g["poses"], g["T"], g["errs"]):
        got = solve(c, size, cam)
        worst = max(worst, (np.abs(got[IDX] - want[IDX]) / np.maximum(1, np.abs(want[IDX]))).max())
        assert abs(got[6] - want[6]) <= ERR_TOL * max(1, want[6]) and abs(got[13] - want[13]) <= ERR_TOL * max(1, want[13])
        # aruco::solvePnP (src/Frame.cc:170): the 4 x 4 float matrices are the Rodrigues matrices of the two rotation vectors with their translations
        for s, o in ((0, 0), (1, 7)):
            M = T[s].reshape(4, 4)
            assert np.abs(M[:3, :3] - rodrigues(got[o:o + 3])).max() < 2e-6 and np.abs(M[:3, 3] - got[o + 3:o + 6]).max() < 2e-6 * max(1, np.abs(got[o + 3:o + 6]).max())
            assert abs(errs[s] - got[6 + o]) <= ERR_TOL * max(1, errs[s])
    assert worst <= POSE_TOL, worst


@pytest.mark.skipif(oracle.ref_ippe() is None, reason="oracle/_ref/libref_ippe.so not built (needs /root/reference)")
def test_live_reference_on_fresh_markers():
    from orb_slam2_aruco_b200 import synth
    R = oracle.ref_ippe()
    vp = C.c_void_p
    cam = np.array([517.3, 516.5, 318.6, 255.3, 0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32)
    n = 0
    for i in (3, 9):
        for m in oracle.aruco_detect(synth.make_frame(i, markers=20)):
            xy = np.ascontiguousarray(m["xy"], np.float32)
            want = np.zeros(14)
            R.ref_ippe_marker_pose(xy.ctypes.data_as(vp), C.c_float(0.187), cam.ctypes.data_as(vp), want.ctypes.data_as(vp))
            got = solve(xy, 0.187, cam)
            assert (np.abs(got[IDX] - want[IDX]) / np.maximum(1, np.abs(want[IDX]))).max() <= POSE_TOL
            assert abs(got[6] - want[6]) <= ERR_TOL * max(1, want[6]) and abs(got[13] - want[13]) <= ERR_TOL * max(1, want[13])
            assert want[6] <= want[13]
            n += 1
    assert n >= 30


@pytest.mark.skipif(oracle.ref_ippe() is None, reason="oracle/_ref/libref_ippe.so not built (needs /root/reference)")
def test_golden_file_is_current(golden_dir):
    g = np.load(os.path.join(golden_dir, "ippe_ref.npz"))
    R = oracle.ref_ippe()
    vp = C.c_void_p
    for i in (0, 57, 300):
        out = np.zeros(14)
        R.ref_ippe_marker_pose(np.ascontiguousarray(g["corners"][i]).ctypes.data_as(vp), C.c_float(float(g["sizes"][i])),
                               np.ascontiguousarray(g["cams"][i]).ctypes.data_as(vp), out.ctypes.data_as(vp))
        assert np.array_equal(out, g["poses"][i])
