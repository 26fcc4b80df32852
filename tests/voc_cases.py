"""Synthetic vocabularies in the ORBvoc.txt format and runners of the reference's own DBoW2 (oracle/_ref/libref_voc.so) and of the oracle
(oracle/bow_oracle.cpp) over them.  Shared by tests/golden/make_voc_golden.py and tests/test_oracle_voc_vs_ref.py.  CPU only."""
import ctypes as C

import numpy as np

import oracle
from test_bow import P, make_tree, oracle_descend

CASES = [(10, 3, False, 4), (10, 3, False, 1), (6, 4, True, 2), (9, 5, True, 4), (10, 2, False, 4)]     # k, L, irregular, levelsup (Frame.cc:353 uses 4)


def write_voc_text(path, k, L, tree, trailing_newline=False):
    """TemplatedVocabulary::loadFromTextFile format (TemplatedVocabulary.h:1338-1425): header `k L scoring weighting` (0 0 = L1_NORM, TF_IDF), then one
    line per node in id order: parent id, is-leaf flag, the 32 descriptor bytes, the weight"""
    parent, leaf, desc, weight = tree
    lines = ["%d %d 0 0" % (k, L)]
    for i in range(1, len(parent)):
        lines.append("%d %d %s %.17g" % (parent[i], leaf[i], " ".join(str(int(b)) for b in desc[i]), weight[i]))
    with open(path, "w") as fh:
        # no newline after the last node: the reference's `while(!f.eof())` loop (TemplatedVocabulary.h:1367) would otherwise parse one more, empty
        # line into a spurious extra child of the root whose leaf flag and descriptor are uninitialised memory (see test_trailing_newline_quirk)
        fh.write("\n".join(lines) + ("\n" if trailing_newline else ""))


def features(rng, tree, n=400):
    feats = rng.integers(0, 256, (n, 32)).astype(np.uint8)
    feats[:40] = tree[2][rng.integers(1, len(tree[0]), 40)]         # exact node descriptors
    feats[40:80] = feats[:40] ^ np.packbits(rng.integers(0, 100, (40, 256)) < 3, axis=1)
    return np.ascontiguousarray(feats)


def vectors_ref(R, path, feats, levelsup):
    h = R.ref_voc_load(path.encode())
    assert h
    n = len(feats)
    bw = np.zeros(n, np.int32); bv = np.zeros(n, np.float64); fn = np.zeros(n, np.int32); fs = np.zeros(n + 1, np.int32); fi = np.zeros(n, np.int32)
    c2 = np.zeros(2, np.int32)
    R.ref_voc_transform(h, P(feats), n, levelsup, P(bw), P(bv), P(fn), P(fs), P(fi), P(c2))
    words = R.ref_voc_size(h)
    R.ref_voc_free(h)
    return dict(words=np.int32(words), bow_words=bw[:c2[0]].copy(), bow_values=bv[:c2[0]].copy(), fv_nodes=fn[:c2[1]].copy(), fv_start=fs[:c2[1] + 1].copy(),
                fv_items=fi[:fs[c2[1]]].copy())


def vectors_oracle(tree, L, feats, levelsup):
    w, wt, nid = oracle_descend(tree, L, feats, levelsup)
    n = len(feats)
    bw = np.zeros(n, np.int32); bv = np.zeros(n, np.float64); fn = np.zeros(n, np.int32); fs = np.zeros(n + 1, np.int32); fi = np.zeros(n, np.int32)
    c2 = np.zeros(2, np.int32)
    oracle.lib().oracle_voc_vectors(P(w), P(wt), P(nid), n, P(bw), P(bv), P(fn), P(fs), P(fi), P(c2))
    return dict(words=np.int32(int(tree[1].sum())), bow_words=bw[:c2[0]].copy(), bow_values=bv[:c2[0]].copy(), fv_nodes=fn[:c2[1]].copy(),
                fv_start=fs[:c2[1] + 1].copy(), fv_items=fi[:fs[c2[1]]].copy())


def same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ("words", "bow_words", "bow_values", "fv_nodes", "fv_start", "fv_items"))
