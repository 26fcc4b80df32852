"""Writes tests/golden/ippe_ref.npz: answers of the reference's OWN pose solver (Thirdparty/aruco/aruco/ippe.cpp compiled unmodified into
oracle/_ref/libref_ippe.so) for the marker corners the detector finds on synthetic frames under three cameras, plus the 400 projected squares of
ippe.npz.  Run in the dev container (needs /root/reference):  python tests/golden/make_ippe_ref_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from orb_slam2_aruco_b200 import synth  # noqa: E402

CAMS = np.array([[517.3, 516.5, 318.6, 255.3, 0.2624, -0.9531, -0.0054, 0.0026, 1.1633],
                 [458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0],
                 [600.0, 600.0, 320.0, 240.0, 0.0, 0.0, 0.0, 0.0, 0.0]], np.float32)


def main():
    R = oracle.ref_ippe()
    assert R is not None, "build oracle/_ref first (python -c 'import oracle; oracle.build()')"
    vp = C.c_void_p
    corners, cams, sizes = [], [], []
    for i in range(6):
        for m in oracle.aruco_detect(synth.make_frame(40 + i, markers=20)):
            corners.append(m["xy"]); cams.append(CAMS[i % 3]); sizes.append(0.187 if i % 2 == 0 else 0.05)
    g = np.load(os.path.join(HERE, "ippe.npz"))
    for c, cam in zip(g["corners"], g["cams"]):
        corners.append(c.reshape(8)); cams.append(cam.astype(np.float32)); sizes.append(float(g["msize"]))
    corners = np.ascontiguousarray(corners, np.float32); cams = np.ascontiguousarray(cams, np.float32); sizes = np.asarray(sizes, np.float32)
    poses = np.zeros((len(corners), 14)); T = np.zeros((len(corners), 2, 16), np.float32); errs = np.zeros((len(corners), 2))
    for i in range(len(corners)):
        R.ref_ippe_marker_pose(corners[i].ctypes.data_as(vp), C.c_float(float(sizes[i])), cams[i].ctypes.data_as(vp), poses[i].ctypes.data_as(vp))
        R.ref_ippe_solvepnp(corners[i].ctypes.data_as(vp), C.c_float(float(sizes[i])), cams[i].ctypes.data_as(vp), T[i].ctypes.data_as(vp), errs[i].ctypes.data_as(vp))
    np.savez_compressed(os.path.join(HERE, "ippe_ref.npz"), corners=corners, cams=cams, sizes=sizes, poses=poses, T=T, errs=errs)
    print("wrote %d cases" % len(corners))


if __name__ == "__main__":
    main()
