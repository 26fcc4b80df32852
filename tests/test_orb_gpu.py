"""GPU: the CUDA extractor, called through the C-ABI, against the CPU oracle on the same seeded frames.
Bit-exact for every field: keypoint coordinates/size/angle/response/octave are compared as raw bits, the
256 descriptor bits as bytes (north_star allows 1e-4 px for coordinates; we hold them to 0)."""
import os

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200.api import ORBextractor

pytestmark = pytest.mark.gpu


def assert_same(k, d, k2, d2):
    assert len(k) == len(k2), (len(k), len(k2))
    for f in k.dtype.names:
        assert np.array_equal(k[f].view(np.uint32), k2[f].view(np.uint32)), f
    assert np.array_equal(d, d2)


@pytest.fixture(scope="module")
def ex(built_lib):
    return ORBextractor(1000, 1.2, 8, 20, 7)


def test_pyramid_levels_bit_exact(ex):
    img = synth.make_frame(21)
    ex(img)
    for level in range(8):
        want = oracle.border_reflect101(oracle.orb_pyramid_level(img, level), 19)
        assert np.array_equal(ex.pyramid_level(0, level), want), level


def test_fast_candidates_bit_exact(ex):
    img = synth.make_frame(22)
    ex(img)
    for level in range(8):
        want = oracle.orb_candidates(img, level)
        got = ex.candidates(0, level)
        assert np.array_equal(got, want), level


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_single_frame_bit_exact(ex, idx):
    img = synth.make_frame(idx)
    k, d = ex(img)
    k2, d2 = oracle.orb_extract(img)
    assert_same(k, d, k2, d2)


@pytest.mark.parametrize("w,h,nf,nl,scale", [(1280, 720, 2000, 8, 1.2), (960, 540, 1000, 8, 1.2), (333, 257, 500, 8, 1.2),
                                             (200, 150, 300, 6, 1.2), (752, 480, 1200, 5, 1.5), (1920, 1080, 4000, 8, 1.2)])
def test_other_geometries(built_lib, w, h, nf, nl, scale):
    e = ORBextractor(nf, scale, nl, 20, 7)
    img = synth.make_frame(30 + nl, w, h)
    k, d = e(img)
    k2, d2 = oracle.orb_extract(img, nf, scale, nl)
    assert_same(k, d, k2, d2)
    e.close()


def test_reference_golden_vectors(built_lib, golden_dir):
    """vectors produced by the reference's own ORBextractor.cc (oracle/_ref) in the dev container"""
    g = np.load(os.path.join(golden_dir, "orb_ref.npz"))
    for i, (idx, w, h, nf) in enumerate(g["cases"]):
        e = ORBextractor(int(nf), 1.2, 8, 20, 7)
        k, d = e(synth.make_frame(int(idx), int(w), int(h)))
        assert_same(k, d, g["kps_%d" % i], g["desc_%d" % i])
        e.close()


def test_batch_equals_per_frame(ex):
    imgs = synth.make_batch(12, first=40)
    kps, desc, counts = ex.extract_batch(imgs)
    for f in range(len(imgs)):
        k2, d2 = oracle.orb_extract(imgs[f])
        assert_same(kps[f, :counts[f]], desc[f, :counts[f]], k2, d2)


def test_strided_input(ex):
    big = np.zeros((3, 500, 700), np.uint8)
    imgs = synth.make_batch(3, first=60)
    big[:, 10:490, 30:670] = imgs
    view = big[:, 10:490, 30:670]
    kps, desc, counts = ex.extract_batch(view)
    for f in range(3):
        k2, d2 = oracle.orb_extract(imgs[f])
        assert_same(kps[f, :counts[f]], desc[f, :counts[f]], k2, d2)


def test_edge_cases(ex):
    k, d = ex(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.full((480, 640), 90, np.uint8))                # flat: no corners anywhere
    assert len(k) == 0
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (480, 640)).astype(np.uint8)   # every cell saturated
    assert_same(*ex(noise), *oracle.orb_extract(noise))
    low = (synth.make_frame(70) // 16 * 3 + 100).astype(np.uint8)   # low contrast: exercises the minThFAST fallback
    assert_same(*ex(low), *oracle.orb_extract(low))
    with pytest.raises(AssertionError):
        ex(np.zeros((480, 640), np.float32))


def test_few_features_and_getters(built_lib):
    e = ORBextractor(50, 1.2, 8, 20, 7)
    img = synth.make_frame(80)
    assert_same(*e(img), *oracle.orb_extract(img, 50))
    lw, lh, q, sf = oracle.orb_levels(640, 480, 50, 1.2, 8)
    assert np.array_equal(e.GetScaleFactors().view(np.uint32), sf.view(np.uint32))
    assert np.array_equal(e.GetFeaturesPerLevel(), q)
    assert e.GetLevels() == 8
    assert np.allclose(e.GetInverseScaleFactors() * e.GetScaleFactors(), 1, atol=1e-6)
    e.close()


def test_full_size_batch_properties(built_lib):
    """BASELINE config[1] size (256 x 640x480): size-independent properties + spot checks against the oracle"""
    e = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 256)
    imgs = synth.make_batch(256, first=1000)
    kps, desc, counts = e.extract_batch(imgs)
    kps2, desc2, counts2 = e.extract_batch(imgs)                 # idempotent / deterministic
    assert np.array_equal(counts, counts2) and np.array_equal(desc, desc2) and np.array_equal(kps, kps2)
    assert (counts > 900).all() and (counts <= e.cap).all()
    for f in range(256):
        k = kps[f, :counts[f]]
        assert (np.diff(k["octave"]) >= 0).all()                  # levels concatenated 0..7
        assert ((k["angle"] >= 0) & (k["angle"] < 360)).all()
    for f in (0, 100, 255):
        assert_same(kps[f, :counts[f]], desc[f, :counts[f]], *oracle.orb_extract(imgs[f]))
    e.close()
