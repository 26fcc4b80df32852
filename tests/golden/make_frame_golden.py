#!/usr/bin/env python
"""Golden vectors for Frame::UndistortKeyPoints (src/Frame.cc:357-388): cv2.undistortPoints(pts, K, D, None, K) of python cv2 4.13 on
random pixel positions (inside and slightly outside a 640x480 image) for the TUM1 camera of the reference's examples.

  python tests/golden/make_frame_golden.py      -> tests/golden/frame.npz
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260002)
    K = np.array([[517.306408, 0, 318.643040], [0, 516.469215, 255.313989], [0, 0, 1]], np.float32)
    D = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
    pts = np.concatenate([rng.uniform([-5, -5], [645, 485], (4000, 2)), [[0, 0], [640, 0], [0, 480], [640, 480]]]).astype(np.float32)
    und = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "frame.npz"), pts=pts, und=und, cam9=np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], *D], np.float32),
                        cv2_version=cv2.__version__)
    print("wrote", len(pts), "points")


if __name__ == "__main__":
    main()
