"""CPU: the restatement (oracle/orb_oracle.cpp) against the reference's own src/ORBextractor.cc compiled
unmodified on the cv shim (oracle/_ref/libref_orb.so).  Skipped where oracle/_ref was never built."""
import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth

pytestmark = pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("idx,w,h,nf", [(2, 640, 480, 1000), (3, 640, 480, 2000), (11, 960, 540, 1000), (12, 333, 257, 500),
                                       (13, 200, 150, 300), (14, 1280, 720, 2000), (15, 1920, 1080, 4000), (16, 640, 480, 1000)])
def test_restatement_equals_reference(idx, w, h, nf):
    img = synth.make_frame(idx, w, h, markers=20 if idx == 16 else 0)      # 15: the C5 frame size; 16: a frame with planted markers (C3)
    k, d = oracle.orb_extract(img, nf)
    k2, d2 = oracle.ref_orb_extract(img, nf)
    assert len(k) == len(k2)
    assert np.array_equal(k, k2)
    assert np.array_equal(d, d2)


def test_low_texture_and_noise():
    rng = np.random.default_rng(5)
    for img in (rng.integers(0, 256, (240, 320)).astype(np.uint8), (synth.make_frame(3, 320, 240) // 16 * 4 + 100).astype(np.uint8)):
        k, d = oracle.orb_extract(img, 500)
        k2, d2 = oracle.ref_orb_extract(img, 500)
        assert np.array_equal(k, k2) and np.array_equal(d, d2)


def test_pyramid_with_border_equals_reference():
    img = synth.make_frame(4, 320, 240)
    for level in (0, 1, 4, 7):
        ref = oracle.ref_orb_pyramid_level(img, level)
        mine = oracle.border_reflect101(oracle.orb_pyramid_level(img, level), 19)
        assert np.array_equal(ref, mine)


def test_canonical_cos_sin_leaves_no_descriptor_bit_different_from_the_box_s_libm():
    """Canonical choice 2 of DESIGN.md section 2: oracle/_ref binds cosf / sinf of ORBextractor.cc:113 to (float)cos((double)x).  The same reference
    source linked against THIS box's libm (oracle/_ref/libref_orb_native_libm.so) must give the same keypoints and the same descriptor bits, or the
    canonicalisation stops being harmless: asserted here, not only reported by tools/ref_sweeps.py."""
    import ctypes as C
    import os
    path = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libref_orb_native_libm.so")
    if not os.path.exists(path):
        pytest.skip("libref_orb_native_libm.so not built")
    native = C.CDLL(path)
    bits = kps = 0
    for idx in range(20, 28):
        img = np.ascontiguousarray(synth.make_frame(idx, 640, 480, markers=10 if idx & 1 else 0))
        k, d = oracle.ref_orb_extract(img, 1000)
        cap = 1000 + 3 * 8 + 64
        raw = np.zeros((cap, 7), np.float32); d2 = np.zeros((cap, 32), np.uint8)
        n = native.ref_orb_extract(img.ctypes.data_as(C.c_void_p), 640, 480, 640, 1000, C.c_float(1.2), 8, 20, 7,
                                   raw.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p), cap)
        assert n == len(k)
        assert np.array_equal(raw[:n, 0], k["x"]) and np.array_equal(raw[:n, 1], k["y"]) and np.array_equal(raw[:n, 3], k["angle"])
        bits += int(np.unpackbits(d ^ d2[:n]).sum()); kps += n
    assert kps > 7000 and bits == 0
