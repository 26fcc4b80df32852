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
    matches = np.frombuffer(buf[o:o + 4 * nk], np.int32)
    k2, d2 = oracle.orb_extract(img)
    assert nk == len(k2) and np.array_equal(desc, d2)
    for name in k2.dtype.names:
        assert np.array_equal(kps[name].view(np.uint32), k2[name].view(np.uint32))
    want = oracle.aruco_detect(img)
    assert np.array_equal(mk["id"], want["id"]) and np.abs(mk["xy"] - want["xy"]).max() <= 1e-4
    n2, m2 = oracle.search_by_bow_bf(d2, k2["angle"], d2, k2["angle"], 0.7, True)
    assert nm == n2 and np.array_equal(matches, m2)
    assert dist == oracle.descriptor_distance(d2[0], d2[1])
