"""GPU: every switchable form of a stage (kept behind an environment variable with its measured numbers) must give the SAME bytes as
the default form: keypoint records, descriptor bits, marker records and match indices of a small and a medium batch, compared through checksums computed
in child processes (the switches are read once per process), plus the one-call single-frame entry point against the three calls it replaces."""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200._lib import MARKER_DTYPE, check, lib, ptr
from orb_slam2_aruco_b200.api import MarkerDetector

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200.api import FrontEnd, MarkerDetector, ORBextractor, ORBmatcher
h = hashlib.sha256()
for (w, hh, nf, n) in ((640, 480, 1000, 3), (640, 480, 1000, 40), (1280, 720, 2000, 2)):
    base = synth.make_batch(min(n, 5), w, hh, 20, first=9100)
    imgs = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
    ex = ORBextractor(nf, 1.2, 8, 20, 7, w, hh, n)
    fe = FrontEnd(ex, MarkerDetector("ARUCO_MIP_25h7", w, hh, n), ORBmatcher(0.7, True))
    rk, rd = ex(synth.make_view(base[0], 1))
    out = fe.process_batch(imgs, rd[:1000], rk[:1000])
    for f in range(n):
        c, m = int(out["counts"][f]), int(out["marker_counts"][f])
        h.update(np.ascontiguousarray(out["kps"][f, :c]).tobytes()); h.update(np.ascontiguousarray(out["desc"][f, :c]).tobytes())
        h.update(np.ascontiguousarray(out["markers"][f, :m]["id"]).tobytes()); h.update(np.round(out["markers"][f, :m]["xy"].astype(np.float64), 3).tobytes())
        h.update(np.ascontiguousarray(out["matches"][f, :c]).tobytes())
    print(w, n, int(out["counts"].sum()), int(out["marker_counts"].sum()), int(out["n_matches"].sum()))
print(h.hexdigest())
""" % ROOT

VARIANTS = [{"B200_FAST_DENSE": "1"}, {"B200_FAST_WARP": "1"}, {"B200_FAST_THREADS": "64"}, {"B200_FAST_TP": "80"}, {"B200_ATHRESH_SMEM": "1"},
            {"B200_CONTOURS_RING": "1"}, {"B200_CONTOURS_RING": "0"}, {"B200_CONTOURS_SHARED": "1"}, {"B200_EMIT_CKPT": "1"}, {"B200_WALK_STEPS": "1"},
            {"B200_MATCH_POPC": "1"}]


def run_child(env):
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=dict(os.environ, **env), timeout=600)
    assert r.returncode == 0, (env, r.stderr[-800:])
    return r.stdout


def test_every_switchable_form_gives_the_default_form_s_bytes(built_lib):
    want = run_child({})
    assert len(want.strip().splitlines()) == 4
    for env in VARIANTS:
        got = run_child(env)
        assert got == want, (env, got, want)        # (refined corners are rounded to 1e-3 px: the forms differ in nothing that feeds them)


def test_one_call_frame_entry_point_equals_the_three_calls(built_lib):
    """b200_aruco_detect_frame_host (markers + poses + contour points, one synchronisation) against detect_host + pose_host + get_contours"""
    img = synth.make_frame(31, 640, 480, 20, "ARUCO_MIP_25h7")
    det = MarkerDetector("ARUCO_MIP_25h7", 640, 480, 1)
    det.detect(img)                                                    # creates the handle
    L = lib()
    cap = L.b200_aruco_max_markers(det._h)
    cam9 = np.array([517.3, 516.5, 318.6, 255.3, 0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32)
    m1 = np.zeros(cap, MARKER_DTYPE); n1 = np.zeros(1, np.int32)
    check(L.b200_aruco_detect_host(det._h, ptr(img), 1, 640, 480, 640, 640 * 480, ptr(m1), ptr(n1)))
    n = int(n1[0])
    assert n >= 15
    ofs1 = np.zeros(n + 1, np.int32); xy1 = np.zeros((40000, 2), np.int32)
    tot1 = L.b200_aruco_get_contours(det._h, 0, n, ptr(ofs1), ptr(xy1), 40000)
    assert tot1 > 0
    p1 = np.zeros((cap, 14), np.float32)
    check(L.b200_aruco_pose_host(ptr(m1), n, C.c_float(0.187), ptr(cam9), ptr(p1), 0))
    m2 = np.zeros(cap, MARKER_DTYPE); n2 = np.zeros(1, np.int32); p2 = np.zeros((cap, 14), np.float32)
    ofs2 = np.zeros(cap + 1, np.int32); xy2 = np.zeros((40000, 2), np.int32)
    tot2 = L.b200_aruco_detect_frame_host(det._h, ptr(img), 640, 480, 640, ptr(m2), ptr(n2), C.c_float(0.187), ptr(cam9), ptr(p2), ptr(ofs2), ptr(xy2), 40000)
    assert tot2 == tot1 and int(n2[0]) == n
    assert m1[:n].tobytes() == m2[:n].tobytes() and np.array_equal(ofs1, ofs2[:n + 1]) and np.array_equal(xy1[:tot1], xy2[:tot1])
    assert p1[:n].tobytes() == p2[:n].tobytes()
    # without camera and contours: markers only, returns 0; a contour buffer that is too small reports the total it would need
    m3 = np.zeros(cap, MARKER_DTYPE); n3 = np.zeros(1, np.int32)
    assert L.b200_aruco_detect_frame_host(det._h, ptr(img), 640, 480, 640, ptr(m3), ptr(n3), C.c_float(0.0), None, None, None, None, 0) == 0
    assert m3[:n].tobytes() == m1[:n].tobytes()
    small = np.zeros((100, 2), np.int32)
    assert L.b200_aruco_detect_frame_host(det._h, ptr(img), 640, 480, 640, ptr(m3), ptr(n3), C.c_float(0.0), None, None, ptr(ofs2), ptr(small), 100) == tot1
    assert np.array_equal(small, xy1[:100])
