"""GPU: marker pose (csrc/pose.cu through b200_aruco_pose_host / MarkerDetector.detect(image, cameraParams, size)) against the
CPU oracle (oracle/ippe_oracle.cpp) and the cv2 IPPE golden vectors."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200._lib import B200Error, MARKER_DTYPE
from orb_slam2_aruco_b200.api import CameraParameters, MarkerDetector

pytestmark = pytest.mark.gpu
TOL = 2e-6          # float outputs of a double pipeline whose acos / sin / cos / hypot differ in the last bit between libm and CUDA


def oracle_pose(corners, msize, cam9):
    out = np.zeros(14)
    oracle.lib().oracle_ippe_marker_pose(np.ascontiguousarray(corners, np.float32).ctypes.data_as(C.c_void_p), C.c_float(msize),
                                         np.ascontiguousarray(cam9, np.float64).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def close(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() <= TOL * max(1.0, np.abs(b).max())


def test_golden_cases_match_oracle_and_cv2(built_lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "ippe.npz"))
    det = MarkerDetector("ARUCO_MIP_25h7")
    msize = float(g["msize"])
    for lo in (0, 200):                                       # the two camera models of the fixture
        cam = g["cams"][lo]
        cp = CameraParameters([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], cam[4:9])
        mk = np.zeros(200, MARKER_DTYPE)
        mk["xy"] = g["corners"][lo:lo + 200].reshape(200, 8)
        poses = det.estimate_poses(mk, msize, cp)
        for i in range(200):
            want = oracle_pose(g["corners"][lo + i], msize, cam)
            p = poses[i]
            assert close(p["rvec"], want[0:3]) and close(p["tvec"], want[3:6]) and close(p["rvec2"], want[7:10]) and close(p["tvec2"], want[10:13]), i
            assert p["err1"] == np.float32(want[6]) or abs(p["err1"] - want[6]) <= 1e-5 * max(1, want[6])
            assert p["err2"] == np.float32(want[13]) or abs(p["err2"] - want[13]) <= 1e-5 * max(1, want[13])
            cv = g["poses"][lo + i]
            assert np.abs(p["rvec"] - cv[0:3]).max() < 1e-5 and np.abs(p["tvec"] - cv[3:6]).max() < 1e-5
    det.close()


def test_detect_with_camera_fills_extrinsics(built_lib):
    img = synth.make_frame(0, markers=20)
    det = MarkerDetector("ARUCO_MIP_25h7")
    cp = CameraParameters([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], [0.2624, -0.9531, -0.0054, 0.0026, 1.1633])
    ms = det.detect(img, cp, 0.187)
    plain = det.detect(img)
    assert len(ms) == len(plain) >= 15 and all(m.Rvec is None and m.ssize == -1 for m in plain)
    for m in ms:
        want = oracle_pose(m.corners, 0.187, cp.cam9())
        assert close(m.Rvec, want[0:3]) and close(m.Tvec, want[3:6]) and m.ssize == pytest.approx(0.187)
        assert m.err1 <= m.err2 and m.Tvec[2] > 0
    det.close()


def test_invalid_arguments(built_lib):
    det = MarkerDetector("ARUCO_MIP_25h7")
    mk = np.zeros(1, MARKER_DTYPE)
    mk["xy"][0] = [270, 290, 370, 290, 370, 190, 270, 190]
    cp = CameraParameters(np.diag([500, 500, 1]))
    with pytest.raises(B200Error):
        det.estimate_poses(mk, 0.0, cp)                       # marker.cpp:328: markerSize <= 0
    with pytest.raises(B200Error):
        det.estimate_poses(mk, 0.1, CameraParameters(np.zeros((3, 3))))
    p = det.estimate_poses(mk, 0.2, CameraParameters([[500, 0, 320], [0, 500, 240], [0, 0, 1]]))[0]
    assert np.allclose(p["tvec"], [0, 0, 1], atol=1e-6) and np.allclose(p["rvec"], 0, atol=1e-6)
    assert len(det.estimate_poses(mk[:0], 0.2, cp)) == 0
    det.close()


def test_poses_equal_the_reference_s_own_solver(built_lib, golden_dir):
    """tests/golden/ippe_ref.npz: answers of Thirdparty/aruco/aruco/ippe.cpp itself (compiled unmodified into oracle/_ref/libref_ippe.so, written by
    tests/golden/make_ippe_ref_golden.py) - k_pose must give the same two poses and errors (float outputs of a double pipeline: 2e-6)"""
    g = np.load(os.path.join(golden_dir, "ippe_ref.npz"))
    det = MarkerDetector("ARUCO_MIP_25h7")
    corners, cams, sizes, want = g["corners"], g["cams"], g["sizes"], g["poses"]
    done = 0
    for cam in np.unique(cams, axis=0):
        for size in np.unique(sizes):
            sel = np.nonzero((cams == cam).all(1) & (sizes == size))[0]
            if len(sel) == 0:
                continue
            cp = CameraParameters([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], cam[4:9])
            mk = np.zeros(len(sel), MARKER_DTYPE)
            mk["xy"] = corners[sel].reshape(len(sel), 8)
            poses = det.estimate_poses(mk, float(size), cp)
            for j, i in enumerate(sel):
                p, w = poses[j], want[i]
                assert close(p["rvec"], w[0:3]) and close(p["tvec"], w[3:6]) and close(p["rvec2"], w[7:10]) and close(p["tvec2"], w[10:13]), i
                assert abs(p["err1"] - w[6]) <= 1e-5 * max(1, w[6]) and abs(p["err2"] - w[13]) <= 1e-5 * max(1, w[13])
                done += 1
    assert done == len(corners)
    det.close()
