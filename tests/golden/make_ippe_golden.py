#!/usr/bin/env python
"""Golden vectors for the marker pose step: cv2.solvePnPGeneric(SOLVEPNP_IPPE) -- the IPPE algorithm of
Thirdparty/aruco/aruco/ippe.cpp as shipped inside OpenCV (python cv2 4.13) -- on noisy projections of a 0.187 m square
(the reference's marker size, src/Frame.cc:131) seen from random poses, with and without lens distortion.

  python tests/golden/make_ippe_golden.py      -> tests/golden/ippe.npz
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260001)
    K = np.array([[517.306408, 0, 318.643040], [0, 516.469215, 255.313989], [0, 0, 1]], np.float64)      # TUM1.yaml of the reference
    dists = [np.zeros(5), np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314])]
    msize = np.float32(0.187)
    h = float(msize) / 2
    obj = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float32)
    corners, cams, poses, truth = [], [], [], []
    for dist in dists:
        n = 0
        while n < 200:
            rv = rng.normal(0, 0.6, 3)
            rv[0] += np.pi * (rng.random() < 0.5)
            tv = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4), rng.uniform(0.6, 3.0)])
            img, _ = cv2.projectPoints(obj.astype(np.float64), rv, tv, K, dist)
            img = img.reshape(4, 2)
            if (img < 0).any() or (img[:, 0] > 640).any() or (img[:, 1] > 480).any():
                continue                                   # the detector only reports markers inside the frame
            img = (img + rng.normal(0, 0.3, (4, 2))).astype(np.float32)
            ok, rvecs, tvecs, errs = cv2.solvePnPGeneric(obj, img, K.astype(np.float32), dist.astype(np.float32), flags=cv2.SOLVEPNP_IPPE)
            assert ok and len(rvecs) == 2
            corners.append(img)
            cams.append(np.array([np.float32(K[0, 0]), np.float32(K[1, 1]), np.float32(K[0, 2]), np.float32(K[1, 2])] +
                                 [np.float32(d) for d in dist], np.float64))
            poses.append(np.concatenate([rvecs[0].ravel(), tvecs[0].ravel(), [errs[0, 0]], rvecs[1].ravel(), tvecs[1].ravel(), [errs[1, 0]]]))
            truth.append(np.concatenate([rv, tv]))
            n += 1
    np.savez_compressed(os.path.join(HERE, "ippe.npz"), corners=np.array(corners), cams=np.array(cams), poses=np.array(poses),
                        truth=np.array(truth), msize=msize, cv2_version=cv2.__version__)
    print("wrote", len(corners), "cases")


if __name__ == "__main__":
    main()
