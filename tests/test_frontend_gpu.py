"""GPU: the combined front-end call (one upload -> extractor + detector + matcher, b200_frontend_host) against the
three oracles, pageable and pinned buffers, batch larger than one upload chunk."""
import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200.api import FrontEnd, MarkerDetector, ORBextractor, ORBmatcher

pytestmark = pytest.mark.gpu


def check(out, imgs, rd, rk):
    for f in range(len(imgs)):
        n = int(out["counts"][f])
        k2, d2 = oracle.orb_extract(imgs[f])
        assert n == len(k2) and np.array_equal(out["desc"][f, :n], d2)
        for name in k2.dtype.names:
            assert np.array_equal(out["kps"][f, :n][name].view(np.uint32), k2[name].view(np.uint32)), name
        want = oracle.aruco_detect(imgs[f])
        m = int(out["marker_counts"][f])
        assert m == len(want) and np.array_equal(out["markers"][f, :m]["id"], want["id"])
        if m:
            assert np.abs(out["markers"][f, :m]["xy"] - want["xy"]).max() <= 1e-4
        nm, mm = oracle.search_by_bow_bf(rd, rk["angle"], d2, k2["angle"], 0.7, True)
        assert int(out["n_matches"][f]) == nm and np.array_equal(out["matches"][f, :n], mm)


def test_frontend_matches_three_oracles(built_lib):
    imgs = synth.make_batch(5, markers=20, first=300)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 8)
    fe = FrontEnd(ex, MarkerDetector("ARUCO_MIP_25h7"), ORBmatcher(0.7, True))
    rk, rd = ex(np.roll(imgs[0], (3, 5), axis=(0, 1)))
    out = fe.process_batch(imgs, rd, rk)
    check(out, imgs, rd, rk)
    assert out["n_matches"][0] > 100
    # pinned buffers take the no-staging path
    import torch
    pin = torch.from_numpy(imgs).pin_memory()
    out2 = fe.alloc_outputs(5, pinned=True)
    fe.process_batch(pin.numpy(), rd, rk, out=out2)
    for key in ("counts", "marker_counts", "n_matches"):
        assert np.array_equal(out[key], out2[key]), key
    for f in range(5):                                   # slots beyond the counts are unspecified
        n, m = int(out["counts"][f]), int(out["marker_counts"][f])
        assert np.array_equal(out["kps"][f, :n], out2["kps"][f, :n]) and np.array_equal(out["desc"][f, :n], out2["desc"][f, :n])
        assert np.array_equal(out["markers"][f, :m], out2["markers"][f, :m]) and np.array_equal(out["matches"][f, :n], out2["matches"][f, :n])


def test_frontend_multi_chunk_batch(built_lib):
    """40 frames = two upload chunks (32 + 8) overlapping with compute"""
    base = synth.make_batch(4, markers=20, first=400)
    imgs = np.concatenate([base] * 10)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 40)
    fe = FrontEnd(ex, MarkerDetector("ARUCO_MIP_25h7", 640, 480, 32), ORBmatcher(0.7, True))
    rk, rd = ex(np.roll(base[1], (3, 5), axis=(0, 1)))
    out = fe.process_batch(imgs, rd, rk)
    check({k: v[:4] for k, v in out.items()}, base, rd, rk)
    for rep in range(1, 10):
        for key in ("counts", "marker_counts", "n_matches"):
            assert np.array_equal(out[key][4 * rep:4 * rep + 4], out[key][:4]), key
        assert np.array_equal(out["desc"][4 * rep:4 * rep + 4], out["desc"][:4])


def test_full_size_c3_batch_properties(built_lib):
    """BASELINE config[2] size (256 x 640x480, 20 markers, match vs 1000 descriptors) through b200_frontend_host: idempotence,
    size-independent properties and spot checks of three frames against the three oracles"""
    imgs = synth.make_batch(256, markers=20, first=2000)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 256)
    fe = FrontEnd(ex, MarkerDetector("ARUCO_MIP_25h7", 640, 480, 256), ORBmatcher(0.7, True))
    rk, rd = ex(np.roll(imgs[0], (3, 5), axis=(0, 1)))
    rd, rk = np.ascontiguousarray(rd[:1000]), np.ascontiguousarray(rk[:1000])
    out = {k: v.copy() for k, v in fe.process_batch(imgs, rd, rk).items()}
    out2 = fe.process_batch(imgs, rd, rk)
    for key in ("counts", "marker_counts", "n_matches"):
        assert np.array_equal(out[key], out2[key]), key                   # deterministic (slots past the counts are unspecified)
    for f in range(256):
        n, m = int(out["counts"][f]), int(out["marker_counts"][f])
        assert out["kps"][f, :n].tobytes() == out2["kps"][f, :n].tobytes() and np.array_equal(out["desc"][f, :n], out2["desc"][f, :n])
        assert np.array_equal(out["matches"][f, :n], out2["matches"][f, :n])
        assert out["markers"][f, :m].tobytes() == out2["markers"][f, :m].tobytes()
    assert (out["counts"] > 900).all() and (out["marker_counts"] >= 12).all() and (out["marker_counts"] <= 20).all()
    for f in range(256):
        m = out["markers"][f, :out["marker_counts"][f]]
        assert (np.diff(m["id"]) > 0).all()                               # sorted by id, de-duplicated
        mt = out["matches"][f, :out["counts"][f]]
        used = mt[mt >= 0]
        assert len(used) == out["n_matches"][f] and len(np.unique(used)) == len(used)       # a reference row matches at most once
    check({k: v[[0, 127, 255]] for k, v in out.items()}, imgs[[0, 127, 255]], rd, rk)
    assert out["n_matches"][0] > 100
