"""CPU: the frame-grid oracle (oracle/frame_oracle.cpp, restating Frame::UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid /
GetFeaturesInArea of src/Frame.cc) against cv2.undistortPoints golden vectors and brute-force definitions."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

vp = C.c_void_p


def P(a):
    return a.ctypes.data_as(vp)


def undistort(kps, cam9):
    out = np.zeros_like(kps)
    oracle.lib().oracle_undistort_keypoints(P(kps), len(kps), P(np.ascontiguousarray(cam9, np.float64)), P(out))
    return out


def grid(un, bounds):
    cs = np.zeros(64 * 48 + 1, np.int32); ci = np.zeros(max(len(un), 1), np.int32)
    oracle.lib().oracle_assign_grid(P(un), len(un), P(bounds), P(cs), P(ci))
    return cs, ci


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "frame.npz"))


def test_undistort_matches_cv2_bit_exact(g):
    kps = np.zeros(len(g["pts"]), oracle.KP_DTYPE)
    kps["x"], kps["y"] = g["pts"][:, 0], g["pts"][:, 1]
    kps["octave"] = np.arange(len(kps)) % 8
    un = undistort(kps, g["cam9"])
    assert np.array_equal(un["x"].view(np.uint32), g["und"][:, 0].view(np.uint32))
    assert np.array_equal(un["y"].view(np.uint32), g["und"][:, 1].view(np.uint32))
    assert np.array_equal(un["octave"], kps["octave"])
    cam0 = g["cam9"].copy(); cam0[4] = 0                      # k1 == 0: the reference copies the keypoints (Frame.cc:359-363)
    assert np.array_equal(undistort(kps, cam0), kps)


def test_bounds_grid_and_area_queries(g):
    kps = np.zeros(len(g["pts"]), oracle.KP_DTYPE)
    kps["x"], kps["y"] = g["pts"][:, 0], g["pts"][:, 1]
    kps["octave"] = np.arange(len(kps)) % 8
    cam = np.ascontiguousarray(g["cam9"], np.float64)
    un = undistort(kps, cam)
    b = np.zeros(4, np.float32)
    oracle.lib().oracle_image_bounds(640, 480, P(cam), P(b))
    corners = g["und"][-4:]
    assert b[0] == min(corners[0, 0], corners[2, 0]) and b[1] == max(corners[1, 0], corners[3, 0])
    assert b[2] == min(corners[0, 1], corners[1, 1]) and b[3] == max(corners[2, 1], corners[3, 1])
    cs, ci = grid(un, b)
    inv_w, inv_h = np.float32(64) / (b[1] - b[0]), np.float32(48) / (b[3] - b[2])
    def rnd(v):                                                      # C round(): half away from zero, on the float product widened to double
        v = v.astype(np.float64)
        return (np.sign(v) * np.floor(np.abs(v) + 0.5)).astype(int)
    px, py = rnd((un["x"] - b[0]) * inv_w), rnd((un["y"] - b[2]) * inv_h)
    inside = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    assert cs[-1] == inside.sum()
    for c in (0, 100, 1500, 64 * 48 - 1):
        want = np.nonzero(inside & (px * 48 + py == c))[0]
        assert np.array_equal(ci[cs[c]:cs[c + 1]], want)                           # index order inside a cell
    rng = np.random.default_rng(3)
    out = np.zeros(4096, np.int32)
    for _ in range(50):
        x, y, r = rng.uniform(0, 640), rng.uniform(0, 480), rng.uniform(5, 60)
        lo, hi = int(rng.integers(-1, 4)), int(rng.integers(-1, 8))
        n = oracle.lib().oracle_features_in_area(P(un), P(cs), P(ci), P(b), C.c_float(x), C.c_float(y), C.c_float(r), lo, hi, P(out), 4096)
        m = inside & (np.abs(un["x"] - np.float32(x)) < np.float32(r)) & (np.abs(un["y"] - np.float32(y)) < np.float32(r))
        if lo > 0 or hi >= 0:
            m &= un["octave"] >= lo
            if hi >= 0:
                m &= un["octave"] <= hi
        # brute force finds the same SET; the cell window may clip features whose cell lies outside it, exactly like the reference
        got = set(out[:n].tolist())
        assert got <= set(np.nonzero(m)[0].tolist())
        assert len(got) >= 0.9 * m.sum() - 2
