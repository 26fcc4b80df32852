#!/usr/bin/env python
"""Generate the committed golden fixtures (run in the dev container: needs python cv2 4.13 and, for the
extractor vectors, oracle/_ref built from /root/reference).

  primitives.npz   inputs + outputs of the real OpenCV primitives the reference calls
                   (resize, copyMakeBorder, GaussianBlur, FAST, fastAtan2) -- pins oracle/cvprim.h
  orb_ref.npz      keypoints + descriptors of the reference's OWN src/ORBextractor.cc (compiled unmodified
                   on oracle/cvshim, oracle/_ref/libref_orb.so) on synthetic frames -- pins the restatement
                   and the CUDA path on the GPU box, where /root/reference does not exist
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import cv2  # noqa: E402

import oracle  # noqa: E402
from orb_slam2_aruco_b200 import synth  # noqa: E402

cv2.setNumThreads(1)


def primitives():
    rng = np.random.default_rng(7)
    out = {}
    base = synth.make_frame(1, 320, 240)
    out["img"] = base
    for i, (dw, dh) in enumerate([(267, 200), (222, 167), (100, 77), (319, 239)]):
        out["resize_%d" % i] = cv2.resize(base, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out["blur"] = cv2.GaussianBlur(base, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    small = base[40:61, 50:63].copy()
    out["border_src"] = small
    out["border"] = cv2.copyMakeBorder(small, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    # FAST on cell-sized ROIs at both thresholds
    rois = []
    for k in range(6):
        y0, x0 = int(rng.integers(0, 200)), int(rng.integers(0, 280))
        hh, ww = int(rng.integers(7, 40)), int(rng.integers(7, 40))
        rois.append((y0, x0, hh, ww))
    out["fast_rois"] = np.array(rois, np.int32)
    for k, (y0, x0, hh, ww) in enumerate(rois):
        roi = np.ascontiguousarray(base[y0:y0 + hh, x0:x0 + ww])
        for thr in (20, 7):
            det = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=True)
            kp = det.detect(roi)
            out["fast_%d_%d" % (k, thr)] = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in kp], np.int32).reshape(-1, 3)
    yx = rng.integers(-200000, 200000, size=(4000, 2)).astype(np.float32)
    yx[:50] = rng.integers(-3, 4, size=(50, 2))
    out["atan2_in"] = yx
    out["atan2_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)
    print("primitives.npz", os.path.getsize(os.path.join(HERE, "primitives.npz")))


def orb_ref():
    assert oracle.ref() is not None, "build oracle/_ref first (make -C oracle ref)"
    out = {}
    cases = [(0, 640, 480, 1000), (5, 320, 240, 500), (9, 413, 307, 700)]
    out["cases"] = np.array(cases, np.int32)
    for i, (idx, w, h, nf) in enumerate(cases):
        img = synth.make_frame(idx, w, h)
        k, d = oracle.ref_orb_extract(img, nf)
        out["img_crc_%d" % i] = np.array([int(img.astype(np.uint64).sum()), int((img.astype(np.uint64) * np.arange(img.size, dtype=np.uint64).reshape(img.shape) % 65521).sum())], np.uint64)
        out["kps_%d" % i] = k
        out["desc_%d" % i] = d
    np.savez_compressed(os.path.join(HERE, "orb_ref.npz"), **out)
    print("orb_ref.npz", os.path.getsize(os.path.join(HERE, "orb_ref.npz")))




def aruco():
    """cv2-driven reference pipeline (cv2_aruco_pipeline.py) on marker frames + primitive vectors"""
    import cv2_aruco_pipeline as P
    out = {}
    cases = [(0, 640, 480, "ARUCO_MIP_25h7"), (1, 640, 480, "ARUCO_MIP_25h7"), (2, 1280, 720, "ARUCO_MIP_25h7"), (3, 640, 480, "ARUCO"),
             (4, 640, 480, "ARUCO_MIP_36h12")]
    out["cases"] = np.array([(a, b, c) for a, b, c, _ in cases], np.int32)
    out["dicts"] = np.array([d for _, _, _, d in cases])
    for i, (idx, w, h, dn) in enumerate(cases):
        nbits, _, codes = synth.dictionaries()[dn]
        img, truth = synth.make_frame(idx, w, h, markers=20, dict_name=dn, return_truth=True)
        r = P.detect(img, codes, nbits)
        out["thres_sum_%d" % i] = np.array([int(r["thres"].astype(np.uint64).sum()), len(r["contours"]), sum(len(c) for c in r["contours"])], np.int64)
        out["contour_sizes_%d" % i] = np.array([len(c) for c in r["contours"]], np.int32)
        out["contour_crc_%d" % i] = np.array([int((c.astype(np.int64) * np.arange(1, 2 * len(c) + 1).reshape(-1, 2)).sum() % 1000003) for c in r["contours"]], np.int32)
        out["candidates_%d" % i] = r["candidates"]
        out["patches_%d" % i] = r["patches"]
        out["prerefine_%d" % i] = r["prerefine"]
        out["ids_%d" % i] = np.array([m[0] for m in r["markers"]], np.int32)
        out["corners_%d" % i] = np.array([m[1] for m in r["markers"]], np.float32).reshape(-1, 4, 2)
        out["truth_ids_%d" % i] = np.array(sorted(t[0] for t in truth), np.int32)
        out["truth_corners_%d" % i] = np.array([t[1] for t in sorted(truth, key=lambda t: t[0])], np.float32)
    # primitive vectors
    img = synth.make_frame(0, 640, 480, markers=20)
    th = cv2.adaptiveThreshold(img, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, 5, 7)
    cs, _ = cv2.findContours(th, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    big = [c.reshape(-1, 2) for c in cs if len(c) > 70][:60]
    out["approx_in_sizes"] = np.array([len(c) for c in big], np.int32)
    out["approx_in"] = np.concatenate(big).astype(np.int32)
    res = [cv2.approxPolyDP(c.reshape(-1, 1, 2), len(c) * 0.05, True).reshape(-1, 2) for c in big]
    out["approx_out_sizes"] = np.array([len(c) for c in res], np.int32)
    out["approx_out"] = np.concatenate(res).astype(np.int32)
    out["approx_convex"] = np.array([bool(cv2.isContourConvex(c.reshape(-1, 1, 2))) if len(c) >= 3 else False for c in res])
    rng = np.random.default_rng(11)
    small = synth.make_frame(6, 160, 120, markers=0)
    out["warp_src"] = small
    quads, Ms, patches = [], [], []
    for k in range(12):
        c = (np.array([[40, 30], [100, 30], [100, 90], [40, 90]], np.float32) + rng.uniform(-25, 25, (4, 2)).astype(np.float32))
        dst = np.array([[0, 0], [34, 0], [34, 34], [0, 34]], np.float32)
        M = cv2.getPerspectiveTransform(c, dst)
        quads.append(c); Ms.append(M); patches.append(cv2.warpPerspective(small, M, (35, 35), flags=cv2.INTER_LINEAR))
    out["warp_quads"] = np.array(quads); out["warp_M"] = np.array(Ms); out["warp_out"] = np.array(patches)
    out["otsu"] = np.array([cv2.threshold(p, 125, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)[0] for p in patches], np.int32)
    for bs in (5, 11, 15):
        out["athr_%d" % bs] = cv2.adaptiveThreshold(small, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, bs, 7)
    out["half_even"] = cv2.resize(small, (80, 60))
    odd = synth.make_frame(7, 135, 67)
    out["half_odd_src"] = odd
    out["half_odd"] = cv2.resize(odd, (67, 33))
    A_list, b_list, x_list = [], [], []
    for k in range(40):
        n = int(rng.integers(2, 24))                     # < 25 rows: cv2 uses its built-in Jacobi SVD (LAPACK takes over above)
        x = np.arange(n, dtype=np.float32) + np.float32(rng.integers(0, 600))
        y = (rng.uniform(-1, 1) * x + rng.uniform(0, 400) + rng.normal(0, 0.5, n)).round().astype(np.float32)
        A = np.zeros((24, 2), np.float32); b = np.zeros(24, np.float32)
        A[:n, 0] = x; A[:n, 1] = 1; b[:n] = y
        _, X = cv2.solve(A[:n], b[:n].reshape(-1, 1), flags=cv2.DECOMP_SVD)
        A_list.append(A); b_list.append(b); x_list.append(np.append(X.reshape(-1), n))
    out["svd_A"] = np.array(A_list); out["svd_b"] = np.array(b_list); out["svd_x"] = np.array(x_list, np.float32)
    np.savez_compressed(os.path.join(HERE, "aruco.npz"), **out)
    print("aruco.npz", os.path.getsize(os.path.join(HERE, "aruco.npz")))


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    which = sys.argv[1:] or ["primitives", "orb_ref", "aruco"]
    if "primitives" in which:
        primitives()
    if "orb_ref" in which:
        orb_ref()
    if "aruco" in which:
        aruco()
