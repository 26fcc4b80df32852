"""CPU: the marker identification stage against the reference's OWN Thirdparty/aruco/aruco/dictionary.cpp + dictionary_based.cpp + markerlabeler.cpp
(compiled unmodified into oracle/_ref/libref_dict.so on oracle/arucoshim): dictionary code tables (the product's csrc/aruco_dicts.inc, which the oracle
includes too), DictionaryBased::detect on canonical patches (Otsu threshold, cell votes, border test, rotations, lookup) vs oracle decode_patch, and
Dictionary::getMarkerImage_id vs the renderer of the synthetic frames.  Golden replay everywhere (tests/golden/dict_ref.npz), live where oracle/_ref exists."""
import hashlib
import os

import numpy as np
import pytest

import dict_cases as dc
import oracle
from orb_slam2_aruco_b200 import synth


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "dict_ref.npz"))


def test_code_tables_equal_the_reference_s(golden):
    for name in dc.DICTS:
        nb, tau, codes = oracle.dictionary_codes(name)
        assert [nb, tau, len(codes)] == golden["meta.%s" % name].tolist(), name
        assert hashlib.sha256(codes.tobytes()).digest() == golden["table.%s" % name].tobytes(), name
        pnb, ptau, pcodes = synth.dictionaries()[name]                   # the product's table file as the Python side parses it
        first = {}
        for i, c in enumerate(pcodes):
            first.setdefault(c, i)                                       # a repeated code keeps its first id (std::map::insert; k_decode: atomicMin)
        want = np.zeros(max(first.values()) + 1, np.uint64)
        for c, i in first.items():
            want[i] = c
        assert (pnb, ptau) == (nb, tau) and np.array_equal(want, codes), name


def test_decode_replays_the_reference(golden):
    for name in dc.DECODE_DICTS:
        patches, want = golden["patches.%s" % name], golden["answers.%s" % name]
        got = np.array([oracle.decode_patch(p, name) for p in patches], np.int32)
        assert np.array_equal(got, want), name
        assert want[:, 0].sum() > 60 and (want[:, 0] == 0).sum() > 60 and set(want[want[:, 0] == 1, 2]) == {0, 1, 2, 3}
        assert np.array_equal(dc.patches_for(name), patches)              # the committed patches are the seeded ones


def test_marker_rendering_equals_get_marker_image(golden):
    for name in dc.DECODE_DICTS:
        for i, want in zip(dc.RENDER_IDS[name], golden["render.%s" % name]):
            cells = synth.marker_cells(name, i)
            assert np.array_equal(np.kron(cells, np.ones((4, 4), np.uint8)) * 255, want)


@pytest.mark.skipif(oracle.ref_dict() is None, reason="oracle/_ref/libref_dict.so not built (needs /root/reference)")
def test_live_reference():
    R = oracle.ref_dict()
    for name in dc.DICTS:
        a, b = oracle.dictionary_codes(name), oracle.dictionary_codes(name, impl=R)
        assert a[:2] == b[:2] and np.array_equal(a[2], b[2]), name
    rng = np.random.default_rng(5)
    for name in dc.DECODE_DICTS:
        patches = dc.patches_for(name, frames=4, seed=9)
        for p in patches:
            assert oracle.decode_patch(p, name) == oracle.decode_patch(p, name, impl=R)
        for i in rng.choice(len(synth.dictionaries()[name][2]), 10, replace=False):
            cells = synth.marker_cells(name, int(i))
            assert np.array_equal(np.kron(cells, np.ones((3, 3), np.uint8)) * 255, dc.ref_marker_image(R, name, int(i), 3))
            assert oracle.decode_patch(dc.ref_marker_image(R, name, int(i), 5), name, impl=R) == (1, int(i), 0)
