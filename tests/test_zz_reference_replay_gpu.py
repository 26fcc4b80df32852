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


def test_adapter_solvepnp_with_the_original_camera(built_lib, tmp_path):
    """src/Frame.cc:155-177 calls aruco::solvePnP(v3d, v2d, mK, mDistCoef) per marker with the ORIGINAL camera (detect() used the resized one) and tests
    v2pose[0].second / v2pose[1].second < 0.7: the adapter's solvePnPSquare against the IPPE oracle for that camera"""
    import ctypes as C
    import subprocess
    import oracle
    from orb_slam2_aruco_b200 import synth
    from test_adapters_gpu import build_adapter_smoke
    exe = build_adapter_smoke(str(tmp_path))
    img = synth.make_frame(31, markers=20)
    raw = os.path.join(str(tmp_path), "f.raw"); out = os.path.join(str(tmp_path), "o.bin")
    img.tofile(raw)
    r = subprocess.run([exe, raw, "640", "480", "ARUCO_MIP_25h7", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(out, "rb").read()
    nk, nmk, nm, dist = np.frombuffer(buf[:16], np.int32)
    o = 16 + 28 * nk + 32 * nk
    mk = np.frombuffer(buf[o:o + 36 * nmk], oracle.MARKER_DTYPE)
    o += 36 * nmk + 4 * nk + 36 * nmk + 36
    errs = np.frombuffer(buf[o:o + 8 * nmk], np.float32).reshape(nmk, 2)
    assert len(buf) >= o + 8 * nmk and nmk >= 15
    # Frame::UndistortArucoCorners through the C++ adapter: the same bits as the oracle's cv::undistortPoints restatement
    un = np.frombuffer(buf[o + 8 * nmk:], np.float32).reshape(-1, 2)
    assert len(un) == 4 * nmk
    kin = np.zeros(4 * nmk, oracle.KP_DTYPE); kin["x"] = mk["xy"].reshape(-1, 2)[:, 0]; kin["y"] = mk["xy"].reshape(-1, 2)[:, 1]
    want_un = np.zeros(4 * nmk, oracle.KP_DTYPE)
    cam64 = np.array([np.float32(v) for v in (517.3, 516.5, 318.6, 255.3, 0.2624, -0.9531, -0.0054, 0.0026, 1.1633)], np.float64)
    oracle.lib().oracle_undistort_keypoints(kin.ctypes.data_as(C.c_void_p), len(kin), cam64.ctypes.data_as(C.c_void_p), want_un.ctypes.data_as(C.c_void_p))
    assert np.array_equal(un[:, 0].view(np.uint32), want_un["x"].view(np.uint32)) and np.array_equal(un[:, 1].view(np.uint32), want_un["y"].view(np.uint32))
    cam9 = np.array([np.float32(v) for v in (517.3, 516.5, 318.6, 255.3, 0.2624, -0.9531, -0.0054, 0.0026, 1.1633)], np.float64)
    for i in range(nmk):
        out14 = np.zeros(14)
        oracle.lib().oracle_ippe_marker_pose(np.ascontiguousarray(mk["xy"][i]).ctypes.data_as(C.c_void_p), C.c_float(0.187),
                                             cam9.ctypes.data_as(C.c_void_p), out14.ctypes.data_as(C.c_void_p))
        e1, e2 = out14[6], out14[13]                                # rvec1 tvec1 err1 rvec2 tvec2 err2
        assert abs(errs[i, 0] - e1) <= 1e-4 * max(1e-3, e2) and abs(errs[i, 1] - e2) <= 1e-4 * max(1e-3, e2)
        if abs(e1 / e2 - 0.7) > 1e-3:
            assert (errs[i, 0] / errs[i, 1] < 0.7) == (e1 / e2 < 0.7)


def test_adapter_detect_resizes_the_camera_like_the_reference(built_lib, tmp_path):
    """CamSize 1280 x 720 (src/Frame.cc:132) against a 640 x 480 frame: detect() resizes the camera (cameraparameters.cpp:158-173) before the pose step.
    With that camera the two IPPE solutions often have almost equal reprojection errors on these frames, so a marker may come back with the solutions
    in either order; each reported pose must be one of the oracle's two."""
    import ctypes as C
    import subprocess
    import oracle
    from orb_slam2_aruco_b200 import synth
    from test_adapters_gpu import build_adapter_smoke
    exe = build_adapter_smoke(str(tmp_path))
    img = synth.make_frame(31, markers=20)
    raw = os.path.join(str(tmp_path), "f.raw"); out = os.path.join(str(tmp_path), "o.bin")
    img.tofile(raw)
    r = subprocess.run([exe, raw, "640", "480", "ARUCO_MIP_25h7", out, "1280", "720"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(out, "rb").read()
    nk, nmk, nm, dist = np.frombuffer(buf[:16], np.int32)
    o = 16 + 28 * nk + 32 * nk
    mk = np.frombuffer(buf[o:o + 36 * nmk], oracle.MARKER_DTYPE); o += 36 * nmk + 4 * nk
    poses = np.frombuffer(buf[o:o + 36 * nmk], np.float32).reshape(nmk, 9); o += 36 * nmk
    cam_used = np.frombuffer(buf[o:o + 36], np.float32)
    ax, ay = np.float32(640) / np.float32(1280), np.float32(480) / np.float32(720)
    want_cam = np.array([np.float32(517.3) * ax, np.float32(516.5) * ay, np.float32(318.6) * ax, np.float32(255.3) * ay,
                         0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32)
    assert np.allclose(cam_used, want_cam, rtol=1e-6, atol=0)
    cam9 = cam_used.astype(np.float64)
    for i in range(nmk):
        out14 = np.zeros(14)
        oracle.lib().oracle_ippe_marker_pose(np.ascontiguousarray(mk["xy"][i]).ctypes.data_as(C.c_void_p), C.c_float(0.187),
                                             cam9.ctypes.data_as(C.c_void_p), out14.ctypes.data_as(C.c_void_p))
        first = np.abs(poses[i, :6] - out14[:6]).max() <= 1e-4 * max(1, np.abs(out14[:6]).max())
        second = np.abs(poses[i, :6] - out14[7:13]).max() <= 1e-4 * max(1, np.abs(out14[7:13]).max())
        tie = abs(out14[6] - out14[13]) <= 1e-3 * out14[13]
        assert first or (tie and second), (i, poses[i, :6], out14)


def test_undistort_aruco_corners_bit_exact(built_lib, golden_dir):
    """Frame::UndistortArucoCorners (src/Frame.cc:388-416): the marker corners of a frame through the keypoint undistortion kernel, against the oracle
    (cv::undistortPoints restatement, bit-exact against cv2 golden vectors)"""
    import ctypes as C
    import oracle
    from orb_slam2_aruco_b200.api import CameraParameters, FrameGrid
    g = np.load(os.path.join(golden_dir, "aruco_ref.npz"))
    xy = np.ascontiguousarray(g["c0.xy"], np.float32).reshape(-1, 2)                  # the reference detector's corners of the first golden frame
    cam = np.array([517.3, 516.5, 318.6, 255.3, 0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32)
    fg = FrameGrid(640, 480, CameraParameters([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], cam[4:9]))
    mk = np.zeros(len(xy) // 4, oracle.MARKER_DTYPE); mk["xy"] = xy.reshape(-1, 8)
    got = fg.undistort_aruco_corners(mk)
    kin = np.zeros(len(xy), oracle.KP_DTYPE); kin["x"] = xy[:, 0]; kin["y"] = xy[:, 1]
    want = np.zeros(len(xy), oracle.KP_DTYPE)
    cam64 = np.ascontiguousarray(cam, np.float64)
    oracle.lib().oracle_undistort_keypoints(kin.ctypes.data_as(C.c_void_p), len(kin), cam64.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
    assert got.shape == xy.shape and len(xy) >= 60
    assert np.array_equal(got[:, 0].view(np.uint32), want["x"].view(np.uint32)) and np.array_equal(got[:, 1].view(np.uint32), want["y"].view(np.uint32))
    assert np.abs(got - xy).max() > 0.5                              # the distortion moves corners visibly
