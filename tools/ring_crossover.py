import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200.api import MarkerDetector
imgs = np.concatenate([synth.make_batch(32, markers=20, first=7000)] * 8)
det = MarkerDetector("ARUCO_MIP_25h7", 640, 480, 256)
for n in (1, 8, 32, 48, 64, 96, 128, 256):
    for _ in range(5): det.detect_batch(imgs[:n])
    ts = []
    for _ in range(40):
        t0 = time.perf_counter(); det.detect_batch(imgs[:n]); ts.append(time.perf_counter() - t0)
    print(n, "%.3f ms" % (1e3 * sorted(ts)[len(ts) // 2]))
