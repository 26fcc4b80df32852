"""Seeded canonical patches for the marker identification stage (DictionaryBased::detect): the warped candidate patches of synthetic frames (true
markers at every rotation plus the many non-marker quads), clean renderings rotated and degraded, and noise.  Shared by
tests/golden/make_dict_golden.py and tests/test_oracle_dict_vs_ref.py.  CPU only."""
import ctypes as C

import numpy as np

import oracle
from orb_slam2_aruco_b200 import synth

DICTS = ["ARUCO", "ARUCO_MIP_16h3", "ARUCO_MIP_25h7", "ARUCO_MIP_36h12", "ARTAG", "ARTOOLKITPLUS", "ARTOOLKITPLUSBCH", "TAG16h5", "TAG25h7", "TAG25h9",
         "TAG36h11", "TAG36h10", "CHILITAGS"]
DECODE_DICTS = ["ARUCO_MIP_25h7", "ARUCO", "ARUCO_MIP_36h12"]
RENDER_IDS = {"ARUCO_MIP_25h7": [0, 7, 99], "ARUCO": [0, 511, 1022], "ARUCO_MIP_36h12": [0, 100, 249]}


def patches_for(dict_name, frames=2, seed=0):
    rng = np.random.default_rng(seed + len(dict_name))
    nbits = synth.dictionaries()[dict_name][0]
    n = int(round(nbits ** 0.5)); ws = 5 * (n + 2)
    out = []
    for f in range(frames):                                       # what the detector really feeds the decoder
        img = synth.make_frame(200 + f, markers=20, dict_name=dict_name)
        out += list(oracle.aruco_stages(img, dict_name)["patches"])
    ids = rng.choice(len(synth.dictionaries()[dict_name][2]), 12, replace=False)
    for i in ids:                                                 # clean renderings: rotations, contrast, blur, noise, a flipped cell, a grey border
        cells = synth.marker_cells(dict_name, int(i))
        base = np.kron(cells, np.ones((5, 5), np.uint8)) * 255
        for r in range(4):
            p = np.rot90(base, r).astype(np.float64)
            out.append(p.astype(np.uint8))
            lo, hi = rng.uniform(10, 110), rng.uniform(130, 250)
            q = lo + p / 255.0 * (hi - lo) + rng.normal(0, rng.choice([2, 12, 30]), p.shape)
            out.append(np.clip(q, 0, 255).astype(np.uint8))
            k = np.array([1, 2, 1]) / 4.0
            q = np.apply_along_axis(lambda v: np.convolve(v, k, "same"), 0, np.apply_along_axis(lambda v: np.convolve(v, k, "same"), 1, p))
            out.append(np.clip(q + rng.normal(0, 8, p.shape), 0, 255).astype(np.uint8))
        c2 = cells.copy(); y, x = rng.integers(1, n + 1, 2); c2[y, x] ^= 1
        out.append((np.kron(c2, np.ones((5, 5), np.uint8)) * 255).astype(np.uint8))
        c3 = cells.copy(); c3[0, rng.integers(0, n + 2)] = 1
        out.append((np.kron(c3, np.ones((5, 5), np.uint8)) * 255).astype(np.uint8))
    for _ in range(40):
        out.append(rng.integers(0, 256, (ws, ws)).astype(np.uint8))
    out.append(np.zeros((ws, ws), np.uint8)); out.append(np.full((ws, ws), 255, np.uint8)); out.append(np.full((ws, ws), 126, np.uint8))
    return np.ascontiguousarray(np.stack(out), np.uint8)


def ref_marker_image(R, dict_name, marker_id, bit_size):
    buf = np.zeros(1 << 16, np.uint8)
    side = R.ref_marker_image(dict_name.encode(), int(marker_id), int(bit_size), buf.ctypes.data_as(C.c_void_p), buf.size)
    assert side > 0
    return buf[:side * side].reshape(side, side).copy()
