"""GPU: the source-compatible C++ adapter classes (include/b200slam_adapters.hpp: ORB_SLAM2::ORBextractor,
aruco::MarkerDetector, ORB_SLAM2::ORBmatcher) compiled with g++ against libb200slam.so and driven like Frame/Tracking do."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200._lib import KP_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_adapter_smoke(tmp):
    exe = os.path.join(tmp, "adapter_smoke")
    pkg = os.path.join(ROOT, "orb_slam2_aruco_b200")
    cmd = ["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adapter_smoke.cpp"),
           "-o", exe, "-L", pkg, "-l:libb200slam.so", "-Wl,-rpath," + pkg]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adapters_compile_without_gpu(built_lib, tmp_path):
    """(runs everywhere) the adapter header is valid C++ and links against the C-ABI"""
    build_adapter_smoke(str(tmp_path))


@pytest.mark.gpu
def test_adapters_match_oracle(built_lib, tmp_path):
    exe = build_adapter_smoke(str(tmp_path))
    img = synth.make_frame(31, markers=20)
    raw = os.path.join(str(tmp_path), "f.raw"); out = os.path.join(str(tmp_path), "o.bin")
    img.tofile(raw)
    r = subprocess.run([exe, raw, "640", "480", "ARUCO_MIP_25h7", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(out, "rb").read()
    nk, nmk, nm, dist = np.frombuffer(buf[:16], np.int32)
    o = 16
    kps = np.frombuffer(buf[o:o + 28 * nk], KP_DTYPE); o += 28 * nk
    desc = np.frombuffer(buf[o:o + 32 * nk], np.uint8).reshape(nk, 32); o += 32 * nk
    mk = np.frombuffer(buf[o:o + 36 * nmk], oracle.MARKER_DTYPE); o += 36 * nmk
    matches = np.frombuffer(buf[o:o + 4 * nk], np.int32); o += 4 * nk
    poses = np.frombuffer(buf[o:o + 36 * nmk], np.float32).reshape(nmk, 9); o += 36 * nmk       # Rvec Tvec err1 err2 ssize
    cam_used = np.frombuffer(buf[o:o + 36], np.float32)              # the camera the adapter's detect() gave the pose step
    k2, d2 = oracle.orb_extract(img)
    assert nk == len(k2) and np.array_equal(desc, d2)
    for name in k2.dtype.names:
        assert np.array_equal(kps[name].view(np.uint32), k2[name].view(np.uint32))
    want = oracle.aruco_detect(img)
    assert np.array_equal(mk["id"], want["id"]) and np.abs(mk["xy"] - want["xy"]).max() <= 1e-4
    n2, m2 = oracle.search_by_bow_bf(d2, k2["angle"], d2, k2["angle"], 0.7, True)
    assert nm == n2 and np.array_equal(matches, m2)
    assert dist == oracle.descriptor_distance(d2[0], d2[1])
    # detect(image, cameraParams, markerSize) (src/Frame.cc:142): extrinsics of every marker against the IPPE oracle
    import ctypes as C
    # CamSize = the image size here: detect() uses the camera as given (the resize of cameraparameters.cpp:158-173 is exercised in
    # tests/test_zz_reference_replay_gpu.py with the reference's hard-coded 1280 x 720)
    cam9 = np.array([np.float32(v) for v in (517.3, 516.5, 318.6, 255.3, 0.2624, -0.9531, -0.0054, 0.0026, 1.1633)], np.float64)
    assert len(cam_used) == 9 and np.array_equal(cam_used.astype(np.float64), cam9)
    for i in range(nmk):
        out14 = np.zeros(14)
        oracle.lib().oracle_ippe_marker_pose(np.ascontiguousarray(mk["xy"][i]).ctypes.data_as(C.c_void_p), C.c_float(0.187),
                                             cam9.ctypes.data_as(C.c_void_p), out14.ctypes.data_as(C.c_void_p))
        assert np.abs(poses[i, :6] - out14[:6]).max() <= 2e-6 * max(1, np.abs(out14[:6]).max())
        assert poses[i, 6] <= poses[i, 7] and poses[i, 8] == np.float32(0.187)


@pytest.mark.gpu
@pytest.mark.skipif(oracle.ref_aruco() is None, reason="oracle/_ref/libref_aruco.so not built (needs /root/reference)")
def test_marker_contour_points_and_dict_info_equal_the_reference(built_lib, tmp_path):
    """aruco::Marker::contourPoints / dict_info (marker.h:57-59) as the adapter's detect() fills them against the reference's own detector
    (markerdetector_impl.cpp:6752-6772): same border, same point order, same dictionary name; the same run checks the public member
    ORBextractor::mvImagePyramid (ORBextractor.h:85) inside the C++ program"""
    import ctypes as C
    exe = build_adapter_smoke(str(tmp_path))
    for seed, dname in ((31, "ARUCO_MIP_25h7"), (77, "ARUCO")):
        img = synth.make_frame(seed, markers=20, dict_name=dname)
        raw = os.path.join(str(tmp_path), "f.raw"); out = os.path.join(str(tmp_path), "o.bin")
        img.tofile(raw)
        r = subprocess.run([exe, raw, "640", "480", dname, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        buf = open(out + ".contours", "rb").read()
        cap = 64
        mk = np.zeros(cap, oracle.MARKER_DTYPE); names = np.zeros((cap, 32), np.uint8); ofs = np.zeros(cap + 1, np.int32); xy = np.zeros((1 << 16, 2), np.int32)
        vp = C.c_void_p
        n = oracle.ref_aruco().ref_aruco_detect_contours(img.ctypes.data_as(vp), 640, 480, 640, dname.encode(), mk.ctypes.data_as(vp), cap, names.ctypes.data_as(vp),
                                                         ofs.ctypes.data_as(vp), xy.ctypes.data_as(vp), len(xy))
        assert n >= 15
        o = 0
        for i in range(n):
            name = buf[o:o + 32].split(b"\0")[0].decode(); o += 32
            ln = int(np.frombuffer(buf[o:o + 4], np.int32)[0]); o += 4
            pts = np.frombuffer(buf[o:o + 8 * ln], np.int32).reshape(ln, 2); o += 8 * ln
            assert name == bytes(names[i]).split(b"\0")[0].decode() == dname
            assert ln == ofs[i + 1] - ofs[i] > 70 and np.array_equal(pts, xy[ofs[i]:ofs[i + 1]])
        assert o == len(buf)
