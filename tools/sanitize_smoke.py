#!/usr/bin/env python
"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): every kernel family at a small and a medium batch size, so that both forms of
the batch-size dependent stages run (ring form and end-to-end contour walkers, shared-memory and global quadtree keys, short and long threshold strips).
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200.api import FrontEnd, MarkerDetector, ORBextractor, ORBmatcher

nbig = int(sys.argv[1]) if len(sys.argv) > 1 else 72
for (w, h, nf, n) in ((640, 480, 1000, 2), (640, 480, 1000, nbig), (1280, 720, 2000, 2), (333, 257, 500, 3)):
    base = synth.make_batch(min(n, 4), w, h, 20 if w >= 640 else 2, first=9000)
    imgs = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
    ex = ORBextractor(nf, 1.2, 8, 20, 7, w, h, n)
    det = MarkerDetector("ARUCO_MIP_25h7", w, h, n)
    fe = FrontEnd(ex, det, ORBmatcher(0.7, True))
    rk, rd = ex(np.roll(base[0], (3, 5), axis=(0, 1)))
    out = fe.process_batch(imgs, rd[:1000], rk[:1000])
    k, d = ex(imgs[0])
    m = det.detect(imgs[0])
    print(w, h, n, int(out["counts"].sum()), int(out["marker_counts"].sum()), int(out["n_matches"].sum()), len(k), len(m))
print("done")
