"""CPU: the bag-of-words transform against the reference's OWN vendored DBoW2 (Thirdparty/DBoW2 compiled unmodified into oracle/_ref/libref_voc.so):
TemplatedVocabulary::loadFromTextFile + transform(features, BowVector, FeatureVector, levelsup) on synthetic vocabularies (the snapshot ships no
ORBvoc.txt) vs oracle/bow_oracle.cpp - word ids, TF-IDF weights after L1 normalisation (exact doubles), node ids, feature order.  Golden replay
everywhere (tests/golden/voc_ref.npz holds the reference's answers; trees and descriptors are re-made from their seeds), live where oracle/_ref exists."""
import os

import numpy as np
import pytest

import oracle
import voc_cases as vc


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "voc_ref.npz"))


def test_oracle_replays_the_reference(golden):
    for j, (k, L, irregular, levelsup) in enumerate(vc.CASES):
        rng = np.random.default_rng(500 + j)
        tree = vc.make_tree(rng, k, L, irregular)
        feats = vc.features(rng, tree)
        got = vc.vectors_oracle(tree, L, feats, levelsup)
        want = {name: golden["c%d.%s" % (j, name)] for name in got}
        assert vc.same(got, want), (k, L, irregular, levelsup)
        assert len(want["bow_words"]) > 20 and len(want["fv_items"]) > 300 and abs(want["bow_values"].sum() - 1.0) < 1e-12


@pytest.mark.skipif(oracle.ref_voc() is None, reason="oracle/_ref/libref_voc.so not built (needs /root/reference)")
def test_live_reference(tmp_path):
    # irregular trees keep L - levelsup <= 2, the shallowest level make_tree puts a leaf on: a descent that ends ABOVE level L - levelsup never
    # assigns *nid in the reference (TemplatedVocabulary.h:1221-1256; the caller's NodeId is uninitialised), the oracle and the product answer 0
    R = oracle.ref_voc()
    for j, (k, L, irregular, levelsup) in enumerate([(10, 4, False, 4), (7, 5, True, 3), (10, 3, False, 0), (4, 6, True, 4)]):
        rng = np.random.default_rng(900 + j)
        tree = vc.make_tree(rng, k, L, irregular)
        feats = vc.features(rng, tree, n=1000)
        path = os.path.join(str(tmp_path), "v%d.txt" % j)
        vc.write_voc_text(path, k, L, tree)
        a = vc.vectors_ref(R, path, feats, levelsup); b = vc.vectors_oracle(tree, L, feats, levelsup)
        assert vc.same(a, b), (k, L, irregular, levelsup)


@pytest.mark.skipif(oracle.ref_voc() is None, reason="oracle/_ref/libref_voc.so not built (needs /root/reference)")
def test_golden_file_is_current(golden, tmp_path):
    R = oracle.ref_voc()
    k, L, irregular, levelsup = vc.CASES[0]
    rng = np.random.default_rng(500)
    tree = vc.make_tree(rng, k, L, irregular)
    feats = vc.features(rng, tree)
    path = os.path.join(str(tmp_path), "v.txt")
    vc.write_voc_text(path, k, L, tree)
    a = vc.vectors_ref(R, path, feats, levelsup)
    assert vc.same(a, {name: golden["c0.%s" % name] for name in a})


@pytest.mark.skipif(oracle.ref_voc() is None, reason="oracle/_ref/libref_voc.so not built (needs /root/reference)")
def test_trailing_newline_quirk(tmp_path):
    """Documented non-determinism of the reference: a vocabulary file that ends with a newline makes loadFromTextFile append one
    more node under the root, built from an empty line - parent 0, leaf flag and descriptor bytes never assigned (uninitialised memory in the
    reference's build).  The canonical vocabulary (oracle, product, adapters) is the one the file lists; here the reference is only shown to grow."""
    import ctypes as C
    R = oracle.ref_voc()
    rng = np.random.default_rng(3)
    tree = vc.make_tree(rng, 5, 2, False)
    a = os.path.join(str(tmp_path), "a.txt"); b = os.path.join(str(tmp_path), "b.txt")
    vc.write_voc_text(a, 5, 2, tree); vc.write_voc_text(b, 5, 2, tree, trailing_newline=True)
    ha, hb = R.ref_voc_load(a.encode()), R.ref_voc_load(b.encode())
    assert R.ref_voc_size(ha) == int(tree[1].sum())
    assert R.ref_voc_size(hb) in (int(tree[1].sum()), int(tree[1].sum()) + 1)      # + 1 when the garbage leaf flag happens to be positive
    R.ref_voc_free(ha); R.ref_voc_free(hb)
