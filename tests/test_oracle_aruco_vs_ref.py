"""CPU: the marker detector oracle (oracle/aruco_oracle.cpp) against the reference's OWN detector - Thirdparty/aruco/aruco/markerdetector.cpp,
markerdetector_impl.cpp (all of MarkerDetector_Impl::detect), marker.cpp, markerlabeler.cpp, dictionary.cpp, dictionary_based.cpp compiled unmodified
into oracle/_ref/libref_aruco.so on oracle/arucoshim and configured as src/Frame.cc:133-139.  Same markers, same order, same ids and bit-identical
refined corners.  Golden replay everywhere (tests/golden/aruco_ref.npz), live on more frames where oracle/_ref exists."""
import os

import numpy as np
import pytest

import aruco_ref_cases as ac
import oracle
from orb_slam2_aruco_b200 import synth


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "aruco_ref.npz"))


def test_oracle_replays_the_reference(golden):
    total = 0
    for j, case in enumerate(ac.CASES):
        got = oracle.aruco_detect(ac.frame(case), case["dict"])
        assert np.array_equal(got["id"], golden["c%d.id" % j]), case
        assert np.array_equal(got["xy"].view(np.uint32), golden["c%d.xy" % j].view(np.uint32)), case       # refined corners: identical float bits
        total += len(got)
        assert len(got) >= case["markers"] - 4
    assert total > 100


@pytest.mark.skipif(oracle.ref_aruco() is None, reason="oracle/_ref/libref_aruco.so not built (needs /root/reference)")
def test_live_reference():
    n = 0
    for seed, w, h, name in [(400, 640, 480, "ARUCO_MIP_25h7"), (401, 640, 480, "ARUCO_MIP_25h7"), (402, 640, 480, "ARUCO"), (403, 1280, 720, "ARUCO_MIP_25h7"),
                             (404, 640, 480, "ARUCO_MIP_36h12"), (405, 640, 480, "ARUCO_MIP_25h7"), (406, 800, 600, "ARUCO_MIP_25h7"), (407, 1920, 1080, "ARUCO")]:
        img = np.ascontiguousarray(synth.make_frame(seed, w, h, markers=20, dict_name=name))
        a = oracle.ref_aruco_detect(img, name); b = oracle.aruco_detect(img, name)
        assert np.array_equal(a["id"], b["id"]) and np.array_equal(a["xy"].view(np.uint32), b["xy"].view(np.uint32)), (seed, w, h, name)
        n += len(a)
    assert n > 130
    # frames without markers, flat frames, a frame of pure noise
    for img in (np.zeros((240, 320), np.uint8), np.full((240, 320), 200, np.uint8), np.random.default_rng(1).integers(0, 256, (480, 640)).astype(np.uint8),
                np.ascontiguousarray(synth.make_frame(410, markers=0))):
        a = oracle.ref_aruco_detect(img); b = oracle.aruco_detect(img)
        assert np.array_equal(a["id"], b["id"]) and np.array_equal(a["xy"].view(np.uint32), b["xy"].view(np.uint32))


@pytest.mark.skipif(oracle.ref_aruco() is None, reason="oracle/_ref/libref_aruco.so not built (needs /root/reference)")
def test_golden_file_is_current(golden):
    case = ac.CASES[0]
    m = oracle.ref_aruco_detect(ac.frame(case), case["dict"])
    assert np.array_equal(m["id"], golden["c0.id"]) and np.array_equal(m["xy"].view(np.uint32), golden["c0.xy"].view(np.uint32))


@pytest.mark.skipif(oracle.ref_aruco() is None, reason="oracle/_ref/libref_aruco.so not built (needs /root/reference)")
def test_live_reference_sweep():
    """a sweep over sizes and dictionaries (a 240-frame / 4116-marker run of the same loop, seeds 1000-1239, had no mismatch either)"""
    sizes = [(640, 480), (640, 480), (640, 480), (800, 600), (1280, 720), (320, 240), (960, 540)]
    dicts = ["ARUCO_MIP_25h7", "ARUCO_MIP_25h7", "ARUCO", "ARUCO_MIP_36h12", "ARUCO_MIP_16h3", "TAG36h11"]
    n = 0
    for seed in range(1000, 1036):
        (w, h), dn = sizes[seed % len(sizes)], dicts[seed % len(dicts)]
        img = np.ascontiguousarray(synth.make_frame(seed, w, h, markers=20 if seed % 5 else 8, dict_name=dn))
        a = oracle.ref_aruco_detect(img, dn); b = oracle.aruco_detect(img, dn)
        assert np.array_equal(a["id"], b["id"]) and np.array_equal(a["xy"].view(np.uint32), b["xy"].view(np.uint32)), (seed, w, h, dn)
        n += len(a)
    assert n > 500


@pytest.mark.skipif(oracle.ref_aruco() is None, reason="oracle/_ref/libref_aruco.so not built (needs /root/reference)")
def test_camera_resize_equals_the_reference_s_cameraparameters():
    """CameraParameters::setParams + resize of the reference's own cameraparameters.cpp (what MarkerDetector::detect applies when CamSize differs from the
    image) against the adapters' CameraParameters.resized: identical float bits"""
    import ctypes as C
    from orb_slam2_aruco_b200.api import CameraParameters
    R = oracle.ref_aruco()
    rng = np.random.default_rng(3)
    for cam_size, size in (((1280, 720), (640, 480)), ((1280, 720), (1920, 1080)), ((640, 480), (1280, 720)), ((1280, 720), (1280, 720)), ((1000, 750), (333, 257))):
        for _ in range(20):
            cam = np.array([rng.uniform(300, 900), rng.uniform(300, 900), rng.uniform(200, 700), rng.uniform(100, 500)], np.float32)
            d = rng.normal(0, 0.3, 5).astype(np.float32)
            out = np.zeros(4, np.float32)
            assert R.ref_camera_resize(cam.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), cam_size[0], cam_size[1], size[0], size[1],
                                       out.ctypes.data_as(C.c_void_p)) == 0
            cp = CameraParameters([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], d, cam_size).resized(*size)
            assert np.array_equal(cp.cam9()[:4].view(np.uint32), out.view(np.uint32)) and np.array_equal(cp.cam9()[4:], d)
