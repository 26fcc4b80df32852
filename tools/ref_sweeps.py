"""Re-runs the end-of-round sweeps quoted in DESIGN.md section 2: every CPU restatement (oracle/*_oracle.cpp) against the reference's OWN sources compiled
into oracle/_ref (needs /root/reference to have been present at build time).  CPU only, about two minutes.

    python tools/ref_sweeps.py [--frames 80]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import match_cases as mc
import match_cases2 as m2
import oracle
from orb_slam2_aruco_b200 import synth


def extractor(n):
    sizes = [(640, 480, 1000), (640, 480, 1000), (752, 480, 1200), (1280, 720, 2000), (320, 240, 500), (960, 540, 1500), (1024, 768, 2000), (641, 479, 1000)]
    bad, nk = [], 0
    for seed in range(3000, 3000 + n):
        w, h, nf = sizes[seed % len(sizes)]
        img = synth.make_frame(seed, w, h, markers=(20 if seed % 3 == 0 else 0))
        k, d = oracle.orb_extract(img, nf); k2, d2 = oracle.ref_orb_extract(img, nf)
        nk += len(k)
        if not (len(k) == len(k2) and np.array_equal(k, k2) and np.array_equal(d, d2)):
            bad.append((seed, w, h, nf))
    return "extractor: %d frames, %d keypoints" % (n, nk), bad


def detector(n):
    sizes = [(640, 480), (640, 480), (640, 480), (800, 600), (1280, 720), (320, 240), (960, 540)]
    dicts = ["ARUCO_MIP_25h7", "ARUCO_MIP_25h7", "ARUCO", "ARUCO_MIP_36h12", "ARUCO_MIP_16h3", "TAG36h11"]
    bad, nm = [], 0
    for seed in range(1000, 1000 + n):
        (w, h), dn = sizes[seed % len(sizes)], dicts[seed % len(dicts)]
        img = np.ascontiguousarray(synth.make_frame(seed, w, h, markers=20 if seed % 5 else 8, dict_name=dn))
        a = oracle.ref_aruco_detect(img, dn); b = oracle.aruco_detect(img, dn)
        nm += len(a)
        if not (np.array_equal(a["id"], b["id"]) and np.array_equal(a["xy"].view(np.uint32), b["xy"].view(np.uint32))):
            bad.append((seed, w, h, dn))
    return "detector: %d frames, %d markers" % (n, nm), bad


def matcher(n):
    R, O = oracle.ref_match(), oracle.lib()
    bad, tot = [], 0

    def cmp(tag, seed, a, b):
        nonlocal tot
        tot += a[0]
        if a[0] != b[0] or not all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:])):
            bad.append((tag, seed))
    for seed in range(100, 100 + n):
        c = mc.bow_inputs(seed=seed)
        for ratio, ori in ((0.6, True), (0.9, False)):
            cmp("bow", seed, mc.run_bow(R, "ref", c, ratio, ori), mc.run_bow(O, "oracle", c, ratio, ori))
            cmp("bow_kfkf", seed, mc.run_bow_kfkf(R, "ref", c, ratio, ori), mc.run_bow_kfkf(O, "oracle", c, ratio, ori))
        c = m2.keyframe_points_inputs(seed=seed)
        cmp("fuse", seed, m2.run_fuse(R, "ref", c, 3.0)[:3], m2.run_fuse(O, "oracle", c, 3.0))
        c = m2.keyframe_points_inputs(seed=seed + 500, sim3=True)
        cmp("fuse_scw", seed, m2.run_fuse_sim3(R, "ref", c, 4.0)[:3], m2.run_fuse_sim3(O, "oracle", c, 4.0))
        cmp("loop", seed, m2.run_loop(R, "ref", c, 10)[:2], m2.run_loop(O, "oracle", c, 10))
        c = m2.sim3_inputs(seed=seed)
        cmp("sim3", seed, m2.run_sim3(R, "ref", c, 7.5)[:2], m2.run_sim3(O, "oracle", c, 7.5))
        c = m2.triangulation_inputs(seed=seed)
        cmp("triangulation", seed, m2.run_triangulation(R, "ref", c, 1), m2.run_triangulation(O, "oracle", c, 1))
    return "matcher: %d seeds x 9 member calls, %d matches" % (n, tot), bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=80)
    args = ap.parse_args()
    if oracle.ref() is None or oracle.ref_aruco() is None or oracle.ref_match() is None:
        raise SystemExit("oracle/_ref is not built (needs /root/reference): python -c 'import oracle; oracle.build()'")
    failed = False
    for f, n in ((extractor, args.frames), (detector, 3 * args.frames), (matcher, max(1, 3 * args.frames // 8))):
        t = time.time()
        what, bad = f(n)
        print("%s: %s (%.0f s)" % (what, "no mismatch" if not bad else "MISMATCHES %s" % bad, time.time() - t))
        failed |= bool(bad)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
