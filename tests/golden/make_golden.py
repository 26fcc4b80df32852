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


if __name__ == "__main__":
    primitives()
    orb_ref()
