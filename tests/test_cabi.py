"""CPU: the C-ABI shared library builds, loads and exports every symbol include/*.h declares; without a GPU every
compute entry point fails loudly (B200_ENODEV) instead of falling back to a CPU path."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(b200_\w+)\s*\(", txt)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(built_lib):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(built_lib, n), "libb200slam.so does not export " + n


def test_product_does_not_touch_the_oracle():
    """the shipped path must not import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "orb_slam2_aruco_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hpp", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, re.M), f
                assert "liboracle" not in txt and "oracle/" not in txt.replace("oracle/orb_oracle.cpp", "").replace("(oracle/", "("), f


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from orb_slam2_aruco_b200 import _lib
    from orb_slam2_aruco_b200.api import ORBextractor, ORBmatcher
    with pytest.raises(_lib.B200Error) as e:
        ORBextractor(1000, 1.2, 8, 20, 7)(np.zeros((120, 160), np.uint8))
    assert e.value.code == _lib.ENODEV
    with pytest.raises(_lib.B200Error) as e:
        ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.zeros(32, np.uint8))
    assert e.value.code == _lib.ENODEV
    # the widened rows (marker pose, frame grid) fail the same way: nothing computes on the CPU
    from orb_slam2_aruco_b200.api import CameraParameters, FrameGrid, MarkerDetector
    cp = CameraParameters([[500, 0, 320], [0, 500, 240], [0, 0, 1]], [0.1, 0, 0, 0, 0])
    mk = np.zeros(1, _lib.MARKER_DTYPE)
    with pytest.raises(_lib.B200Error) as e:
        MarkerDetector("ARUCO_MIP_25h7").estimate_poses(mk, 0.1, cp)
    assert e.value.code == _lib.ENODEV
    with pytest.raises(_lib.B200Error) as e:
        FrameGrid(640, 480, cp)
    assert e.value.code == _lib.ENODEV


def test_bad_arguments_are_rejected(built_lib):
    h = C.c_void_p()
    assert built_lib.b200_orb_create(C.byref(h), 1000, 1.0, 8, 20, 7, 640, 480, 1, 0) == -1     # scale factor must be > 1
    assert built_lib.b200_orb_create(None, 1000, 1.2, 8, 20, 7, 640, 480, 1, 0) == -1
    assert built_lib.b200_orb_destroy(None) == 0
    assert built_lib.b200_orb_max_keypoints(None) == -1


def test_keyframe_side_entry_points_validate_and_refuse_without_gpu(built_lib):
    """b200_match_for_triangulation_host / b200_match_kf_radius_host / b200_kf_project_host / b200_kf_search_points_host: sizes are checked before the
    device is touched (B200_EINVAL = -1), and with valid sizes nothing computes on the CPU (B200_ENODEV)"""
    import torch
    from orb_slam2_aruco_b200 import _lib
    L = built_lib
    z = np.zeros(64, np.float32); zi = np.zeros(64, np.int32); zb = np.zeros(64, np.uint8)
    p = lambda a: a.ctypes.data
    assert L.b200_match_kf_radius_host(None, None, -1, None, None, None, None, 0, None, 8, C.c_double(0), None, None, 0) == _lib.EINVAL
    assert L.b200_match_kf_radius_host(None, None, 0, None, None, None, None, 1, None, 17, C.c_double(0), None, None, 0) == _lib.EINVAL      # nlevels > 16
    assert L.b200_kf_project_host(None, None, None, None, None, None, None, None, None, None, -1, 3.0, None, None, 8, None, None, None, 0) == _lib.EINVAL
    assert L.b200_kf_search_points_host(None, None, 0, None, None, None, None, None, None, None, None, None, None, None, None, 1, 3.0, None, None, None, 0,
                                        C.c_double(0), None, None, None, None, None, 0) == _lib.EINVAL
    assert L.b200_match_for_triangulation_host(None, None, 1, None, None, 1, None, None, None, None, -1, None, None, None, None, 8, 1, 50, None, 0) == _lib.EINVAL
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert L.b200_match_kf_radius_host(p(z), p(zb), 1, p(z), p(z), p(zi), p(zb), 1, p(z), 8, C.c_double(5.99), p(zi), p(zi), 0) == _lib.ENODEV
    assert L.b200_kf_project_host(p(z), p(z), p(z), None, None, p(z), p(z), p(z), None, p(z), 1, 3.0, p(z), p(z), 8, p(zb), p(z), p(zi), 0) == _lib.ENODEV
    assert L.b200_match_for_triangulation_host(p(z), p(zb), 1, p(z), p(zb), 1, p(zi), p(zi), p(zi), p(zi), 1, p(z), p(z), p(z), p(z), 8, 1, 50, p(zi), 0) == _lib.ENODEV


def test_camera_parameters_resize_like_the_reference():
    """CameraParameters::resize (cameraparameters.cpp:158-173), applied by detect() whenever CamSize differs from the image - in the reference always,
    its CamSize being the hard-coded 1280 x 720 of src/Frame.cc:132: float factors, fx cx by the width ratio, fy cy by the height ratio"""
    from orb_slam2_aruco_b200.api import CameraParameters
    cp = CameraParameters([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], [0.2624, -0.9531, -0.0054, 0.0026, 1.1633], (1280, 720))
    r = cp.resized(640, 480)
    ax, ay = np.float32(640) / np.float32(1280), np.float32(480) / np.float32(720)
    assert r.CamSize == (640, 480) and cp.CamSize == (1280, 720)
    assert r.cam9()[:4].tolist() == [np.float32(517.3) * ax, np.float32(516.5) * ay, np.float32(318.6) * ax, np.float32(255.3) * ay]
    assert np.array_equal(r.cam9()[4:], cp.cam9()[4:])
    assert cp.resized(1280, 720) is cp and r.resized(640, 480) is r
    assert CameraParameters([[500, 0, 320], [0, 500, 240], [0, 0, 1]]).resized(64, 48).CamSize is None       # no CamSize: used as given


def test_undistort_aruco_corners_without_distortion_needs_no_device(built_lib):
    """Frame::UndistortArucoCorners (src/Frame.cc:388-416) returns early when k1 == 0: the corners are used as they are - no device involved; with
    distortion the call needs the GPU like everything else"""
    import torch
    from orb_slam2_aruco_b200 import _lib
    xy = np.arange(16, dtype=np.float32).reshape(8, 2) * 7.25
    out = np.zeros_like(xy)
    cam = np.array([500, 500, 320, 240, 0, 0, 0, 0, 0], np.float32)
    assert built_lib.b200_frame_undistort_points_host(xy.ctypes.data, 8, cam.ctypes.data, out.ctypes.data, 0) == 0 and np.array_equal(out, xy)
    assert built_lib.b200_frame_undistort_points_host(xy.ctypes.data, -1, cam.ctypes.data, out.ctypes.data, 0) == _lib.EINVAL
    if not torch.cuda.is_available():
        cam[4] = 0.1
        assert built_lib.b200_frame_undistort_points_host(xy.ctypes.data, 8, cam.ctypes.data, out.ctypes.data, 0) == _lib.ENODEV


def test_rt_matrix_equals_cv2_rodrigues():
    """the 4 x 4 pose matrices aruco::solvePnP returns (getRTMatrix, ippe.cpp:16-60) are built on the host from the device's Rodrigues vectors"""
    cv2 = pytest.importorskip("cv2")
    from orb_slam2_aruco_b200.api import MarkerDetector
    rng = np.random.default_rng(2)
    for _ in range(50):
        r = rng.normal(0, 1.5, 3); t = rng.normal(0, 2, 3)
        T = MarkerDetector.rt_matrix(r, t)
        R, _ = cv2.Rodrigues(r.reshape(3, 1))
        assert T.dtype == np.float32 and np.abs(T[:3, :3] - R).max() < 1e-6 and np.array_equal(T[:3, 3], t.astype(np.float32)) and T[3].tolist() == [0, 0, 0, 1]
    assert np.array_equal(MarkerDetector.rt_matrix(np.zeros(3), np.zeros(3)), np.eye(4, dtype=np.float32))
