"""Seeded frames for the detector-vs-reference parity cases (shared by tests/golden/make_aruco_ref_golden.py and the tests that replay it)."""
import numpy as np

from orb_slam2_aruco_b200 import synth

CASES = [dict(seed=300, w=640, h=480, markers=20, dict="ARUCO_MIP_25h7"), dict(seed=301, w=640, h=480, markers=20, dict="ARUCO_MIP_25h7"),
         dict(seed=302, w=640, h=480, markers=12, dict="ARUCO"), dict(seed=303, w=640, h=480, markers=20, dict="ARUCO_MIP_36h12"),
         dict(seed=304, w=1280, h=720, markers=20, dict="ARUCO_MIP_25h7"), dict(seed=305, w=1920, h=1080, markers=20, dict="ARUCO_MIP_25h7"),
         dict(seed=306, w=640, h=480, markers=0, dict="ARUCO_MIP_25h7"), dict(seed=307, w=320, h=240, markers=6, dict="ARUCO_MIP_25h7")]


def frame(case):
    return np.ascontiguousarray(synth.make_frame(case["seed"], case["w"], case["h"], markers=case["markers"], dict_name=case["dict"]))
