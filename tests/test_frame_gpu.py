"""GPU: keypoint undistortion, image bounds, the 64x48 feature grid and GetFeaturesInArea (csrc/frame.cu through the C-ABI) against
the CPU oracle (oracle/frame_oracle.cpp) and the cv2.undistortPoints golden vectors: bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200._lib import KP_DTYPE
from orb_slam2_aruco_b200.api import CameraParameters, FrameGrid, ORBextractor

pytestmark = pytest.mark.gpu
vp = C.c_void_p


def P(a):
    return a.ctypes.data_as(vp)


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(a.shape[0], -1).copy()).cuda() if a.dtype == KP_DTYPE else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_undistort_bounds_grid_area_bit_exact(built_lib, golden_dir):
    import torch
    g = np.load(os.path.join(golden_dir, "frame.npz"))
    cam = g["cam9"]
    cp = CameraParameters([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], cam[4:9])
    fg = FrameGrid(640, 480, cp)
    cam64 = np.ascontiguousarray(cam, np.float64)
    b = np.zeros(4, np.float32)
    oracle.lib().oracle_image_bounds(640, 480, P(cam64), P(b))
    assert np.array_equal(fg.bounds.view(np.uint32), b.view(np.uint32))
    # real keypoints of two frames + the golden points as a third "frame"
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    kps, desc, counts = ex.extract_batch(synth.make_batch(2, first=60))
    cap = max(kps.shape[1], len(g["pts"]))
    allk = np.zeros((3, cap), KP_DTYPE)
    allk[:2, :kps.shape[1]] = kps
    allk[2, :len(g["pts"])]["x"] = g["pts"][:, 0]; allk[2, :len(g["pts"])]["y"] = g["pts"][:, 1]
    allk[2, :len(g["pts"])]["octave"] = np.arange(len(g["pts"])) % 8
    cnt = np.array([counts[0], counts[1], len(g["pts"])], np.int32)
    d_k = torch.from_numpy(allk.view(np.uint8).reshape(3, cap, 28).copy()).cuda()
    d_un = torch.zeros_like(d_k)
    d_cnt = torch.from_numpy(cnt).cuda()
    fg.undistort(d_k, d_cnt, d_un)
    d_cs = torch.zeros((3, 64 * 48 + 1), dtype=torch.int32, device="cuda")
    d_ci = torch.zeros((3, cap), dtype=torch.int32, device="cuda")
    fg.assign(d_un, d_cnt, d_cs, d_ci)
    torch.cuda.synchronize()
    un = d_un.cpu().numpy().reshape(3, cap * 28).view(KP_DTYPE).reshape(3, cap)
    cs, ci = d_cs.cpu().numpy(), d_ci.cpu().numpy()
    n2 = len(g["pts"])
    assert np.array_equal(un[2, :n2]["x"].view(np.uint32), g["und"][:, 0].view(np.uint32))       # == cv2.undistortPoints
    assert np.array_equal(un[2, :n2]["y"].view(np.uint32), g["und"][:, 1].view(np.uint32))
    rng = np.random.default_rng(9)
    for f in range(3):
        n = int(cnt[f])
        want = np.zeros(n, oracle.KP_DTYPE)
        oracle.lib().oracle_undistort_keypoints(P(np.ascontiguousarray(allk[f, :n])), n, P(cam64), P(want))
        assert un[f, :n].tobytes() == want.tobytes()
        wcs = np.zeros(64 * 48 + 1, np.int32); wci = np.zeros(max(n, 1), np.int32)
        oracle.lib().oracle_assign_grid(P(want), n, P(b), P(wcs), P(wci))
        assert np.array_equal(cs[f], wcs) and np.array_equal(ci[f, :wcs[-1]], wci[:wcs[-1]])
        # GetFeaturesInArea: 64 random windows with level filters (the SearchByProjection pattern: r = 2.5..4 x scale, levels l-1..l)
        nq = 64
        q = np.stack([rng.uniform(0, 640, nq), rng.uniform(0, 480, nq), rng.uniform(3, 70, nq)], 1).astype(np.float32)
        lv = np.stack([rng.integers(-1, 5, nq), rng.integers(-1, 8, nq)], 1).astype(np.int32)
        d_out = torch.zeros((nq, 512), dtype=torch.int32, device="cuda"); d_n = torch.zeros(nq, dtype=torch.int32, device="cuda")
        fg.features_in_area(d_un[f], d_cs[f], d_ci[f], torch.from_numpy(q).cuda(), torch.from_numpy(lv).cuda(), d_out, d_n)
        torch.cuda.synchronize()
        out, on = d_out.cpu().numpy(), d_n.cpu().numpy()
        buf = np.zeros(512, np.int32)
        for k in range(nq):
            m = oracle.lib().oracle_features_in_area(P(want), P(wcs), P(wci), P(b), C.c_float(q[k, 0]), C.c_float(q[k, 1]), C.c_float(q[k, 2]),
                                                    int(lv[k, 0]), int(lv[k, 1]), P(buf), 512)
            assert on[k] == m and np.array_equal(out[k, :min(m, 512)], buf[:min(m, 512)]), (f, k)
    ex.close()


def test_no_distortion_copies_and_edge_cases(built_lib):
    import torch
    cp = CameraParameters([[500, 0, 320], [0, 500, 240], [0, 0, 1]])             # k1 == 0
    fg = FrameGrid(640, 480, cp)
    assert fg.bounds.tolist() == [0.0, 640.0, 0.0, 480.0]
    k = np.zeros((1, 8), KP_DTYPE)
    k["x"][0] = [0, 639.9, 320, 5, 5, 5, 700, -3]; k["y"][0] = [0, 479.9, 240, 5, 5, 5, 10, 10]
    d_k = torch.from_numpy(k.view(np.uint8).reshape(1, 8, 28).copy()).cuda(); d_un = torch.zeros_like(d_k)
    d_cnt = torch.tensor([8], dtype=torch.int32, device="cuda")
    fg.undistort(d_k, d_cnt, d_un)
    d_cs = torch.zeros((1, 64 * 48 + 1), dtype=torch.int32, device="cuda"); d_ci = torch.zeros((1, 8), dtype=torch.int32, device="cuda")
    fg.assign(d_un, d_cnt, d_cs, d_ci)
    torch.cuda.synchronize()
    assert torch.equal(d_un, d_k)
    cs, ci = d_cs.cpu().numpy()[0], d_ci.cpu().numpy()[0]
    assert cs[-1] == 6                                                            # (700, 10) and (639.9 -> column 64) fall outside, (-3 -> round(-0.3) = 0) stays
    c = 0 * 48 + 1                                                                # the three keypoints at (5, 5): round(0.5) = 1 in both axes? x: 5*0.1 = 0.5 -> 1; y: 5*0.1 -> 1
    assert ci[cs[1 * 48 + 1]:cs[1 * 48 + 2]].tolist() == [3, 4, 5]                # push order kept
    d_cnt0 = torch.tensor([0], dtype=torch.int32, device="cuda")
    fg.assign(d_un, d_cnt0, d_cs, d_ci)
    torch.cuda.synchronize()
    assert int(d_cs.cpu().numpy()[0, -1]) == 0
