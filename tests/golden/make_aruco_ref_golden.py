"""Writes tests/golden/aruco_ref.npz: what the reference's OWN marker detector (Thirdparty/aruco/aruco markerdetector_impl.cpp & co., compiled unmodified
into oracle/_ref/libref_aruco.so, configured as src/Frame.cc:133-139) returns for seeded synthetic frames (orb_slam2_aruco_b200/synth.py): ids and the
CORNER_LINES-refined corners, raw float bits.  Replayed by tests/test_oracle_aruco_vs_ref.py (oracle) and tests/test_aruco_gpu.py (CUDA) on boxes
without the reference.

    python tests/golden/make_aruco_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import aruco_ref_cases as ac
import oracle


def main():
    if oracle.ref_aruco() is None:
        raise SystemExit("oracle/_ref/libref_aruco.so not built (needs /root/reference): make -C oracle ref")
    out = {}
    for j, case in enumerate(ac.CASES):
        img = ac.frame(case)
        m = oracle.ref_aruco_detect(img, case["dict"])
        out["c%d.id" % j] = m["id"].copy(); out["c%d.xy" % j] = m["xy"].copy()
        print(j, case, "->", len(m), "markers")
    path = os.path.join(HERE, "aruco_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
