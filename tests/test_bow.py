"""DBoW2 vocabulary descent: the oracle (oracle/bow_oracle.cpp) against a brute-force numpy definition on CPU, and the
CUDA path (csrc/bow.cu through the C-ABI / ORBVocabulary) against the oracle on the GPU.  The reference snapshot ships no vocabulary
file, so the trees are synthetic, written in the node order of ORBvoc.txt (and once through the text-file loader)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

vp = C.c_void_p


def P(a):
    return a.ctypes.data_as(vp)


def make_tree(rng, k, L, irregular=False):
    """nodes in breadth-first file order: parent, is_leaf, descriptor, weight (entry 0 = root)"""
    parent, leaf, level = [0], [0], [0]
    frontier = [0]
    for lvl in range(1, L + 1):
        nxt = []
        for p in frontier:
            nch = k if not irregular else int(rng.integers(1, k + 1))
            for _ in range(nch):
                nid = len(parent)
                parent.append(p); level.append(lvl)
                is_leaf = lvl == L or (irregular and lvl >= 2 and rng.random() < 0.2)
                leaf.append(1 if is_leaf else 0)
                if not is_leaf:
                    nxt.append(nid)
        frontier = nxt
    n = len(parent)
    desc = rng.integers(0, 256, (n, 32)).astype(np.uint8)
    desc[3] = desc[2]                                          # two identical siblings: the first one must win ties
    weight = np.where(np.array(leaf) > 0, rng.uniform(0.1, 8.0, n), 0.0)
    weight[np.nonzero(leaf)[0][::17]] = 0.0                    # some stopped words
    return np.array(parent, np.int32), np.array(leaf, np.uint8), desc, weight.astype(np.float64)


def oracle_descend(tree, L, feats, levelsup):
    parent, leaf, desc, weight = tree
    n = len(feats)
    w = np.zeros(n, np.int32); wt = np.zeros(n, np.float64); nid = np.zeros(n, np.int32)
    oracle.lib().oracle_voc_transform(P(parent), P(leaf), P(desc), P(weight), len(parent), L, P(np.ascontiguousarray(feats)), n, levelsup, P(w), P(wt), P(nid))
    return w, wt, nid


def brute_descend(tree, L, feats, levelsup):
    parent, leaf, desc, weight = tree
    pc = np.array([bin(i).count("1") for i in range(256)])
    children = {}
    for i in range(1, len(parent)):
        children.setdefault(int(parent[i]), []).append(i)
    word_of = {int(i): j for j, i in enumerate(np.nonzero(leaf)[0])}
    out = []
    for f in feats:
        cur, nid, lvl = 0, 0, 0
        while True:
            lvl += 1
            ch = children[cur]
            d = [int(pc[desc[c] ^ f].sum()) for c in ch]
            cur = ch[int(np.argmin(d))]                        # np.argmin: first minimum
            if lvl == L - levelsup:
                nid = cur
            if leaf[cur]:
                break
        out.append((word_of[cur], weight[cur], nid))
    return out


@pytest.mark.parametrize("k,L,irregular,levelsup", [(10, 3, False, 1), (6, 4, True, 2), (10, 2, False, 4)])
def test_oracle_matches_brute_force(k, L, irregular, levelsup):
    rng = np.random.default_rng(k * 10 + L)
    tree = make_tree(rng, k, L, irregular)
    feats = rng.integers(0, 256, (200, 32)).astype(np.uint8)
    feats[:20] = tree[2][rng.integers(1, len(tree[0]), 20)]     # some exact node descriptors
    w, wt, nid = oracle_descend(tree, L, feats, levelsup)
    for i, (bw, bwt, bn) in enumerate(brute_descend(tree, L, feats, levelsup)):
        assert (w[i], wt[i], nid[i]) == (bw, bwt, bn), i
    # BowVector / FeatureVector assembly
    n = len(feats)
    bw = np.zeros(n, np.int32); bv = np.zeros(n, np.float64); fn = np.zeros(n, np.int32); fs = np.zeros(n + 1, np.int32); fi = np.zeros(n, np.int32)
    c2 = np.zeros(2, np.int32)
    oracle.lib().oracle_voc_vectors(P(w), P(wt), P(nid), n, P(bw), P(bv), P(fn), P(fs), P(fi), P(c2))
    live = wt > 0
    assert np.array_equal(bw[:c2[0]], np.unique(w[live])) and abs(bv[:c2[0]].sum() - 1.0) < 1e-12
    assert np.array_equal(fn[:c2[1]], np.unique(nid[live])) and fs[c2[1]] == live.sum()
    for j in range(c2[1]):
        assert np.array_equal(fi[fs[j]:fs[j + 1]], np.nonzero(live & (nid == fn[j]))[0])


@pytest.mark.gpu
@pytest.mark.parametrize("k,L,irregular,levelsup", [(10, 3, False, 1), (6, 4, True, 2), (10, 4, False, 4), (20, 2, True, 1)])
def test_cuda_descent_matches_oracle(built_lib, tmp_path, k, L, irregular, levelsup):
    from orb_slam2_aruco_b200 import synth
    from orb_slam2_aruco_b200.api import ORBVocabulary, ORBextractor
    rng = np.random.default_rng(100 + k + L)
    tree = make_tree(rng, k, L, irregular)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    _, feats = ex(synth.make_frame(95))
    feats = np.concatenate([feats, tree[2][rng.integers(1, len(tree[0]), 50)]])
    if (k, L) == (10, 3):                                       # once through the text-file loader (the ORBvoc.txt format)
        path = os.path.join(str(tmp_path), "voc.txt")
        with open(path, "w") as fh:
            fh.write("%d %d 0 0\n" % (k, L))
            for i in range(1, len(tree[0])):
                fh.write("%d %d %s %.17g\n" % (tree[0][i], tree[1][i], " ".join(str(int(b)) for b in tree[2][i]), tree[3][i]))
        voc = ORBVocabulary.loadFromTextFile(path)
    else:
        voc = ORBVocabulary(k, L, *tree)
    assert voc.size() == int(tree[1].sum())
    w, wt, nid = voc.descend(feats, levelsup)
    w2, wt2, nid2 = oracle_descend(tree, L, feats, levelsup)
    assert np.array_equal(w, w2) and np.array_equal(wt, wt2) and np.array_equal(nid, nid2)
    bow, fv = voc.transform(feats, levelsup)
    n = len(feats)
    bw = np.zeros(n, np.int32); bv = np.zeros(n, np.float64); fn = np.zeros(n, np.int32); fs = np.zeros(n + 1, np.int32); fi = np.zeros(n, np.int32)
    c2 = np.zeros(2, np.int32)
    oracle.lib().oracle_voc_vectors(P(w2), P(wt2), P(nid2), n, P(bw), P(bv), P(fn), P(fs), P(fi), P(c2))
    assert list(bow.keys()) == bw[:c2[0]].tolist() and np.allclose(list(bow.values()), bv[:c2[0]], rtol=0, atol=1e-15)
    assert list(fv.keys()) == fn[:c2[1]].tolist()
    for j, node in enumerate(fn[:c2[1]]):
        assert fv[int(node)] == fi[fs[j]:fs[j + 1]].tolist()
    voc.close(); ex.close()


def _fv_arrays(fv):
    """{node: [indices]} -> the oracle's sorted arrays (nodes, start, items)"""
    nodes = np.array(sorted(fv), np.int32)
    start = np.zeros(len(nodes) + 1, np.int32)
    items = []
    for j, nd in enumerate(nodes):
        items += list(fv[int(nd)])
        start[j + 1] = len(items)
    return nodes, start, np.array(items if items else [0], np.int32)


def oracle_bow_nodes(mode, d1, a1, v1, fv1, d2, a2, v2, fv2, ratio, ori):
    d1 = np.ascontiguousarray(d1); d2 = np.ascontiguousarray(d2)
    a1 = np.ascontiguousarray(a1, np.float32); a2 = np.ascontiguousarray(a2, np.float32)
    v1 = np.ascontiguousarray(v1, np.uint8); v2 = np.ascontiguousarray(v2, np.uint8)
    n1, s1, i1 = _fv_arrays(fv1); n2, s2, i2 = _fv_arrays(fv2)
    if mode == 0:
        out = np.zeros(max(len(d2), 1), np.int32)
        n = oracle.lib().oracle_search_by_bow_nodes(P(d1), P(a1), P(v1), P(n1), P(s1), P(i1), len(n1), P(d2), P(a2), len(d2), P(n2), P(s2), P(i2), len(n2),
                                                    C.c_float(ratio), int(ori), P(out))
        return n, out[:len(d2)]
    out = np.zeros(max(len(d1), 1), np.int32)
    n = oracle.lib().oracle_search_by_bow_kfkf_nodes(P(d1), P(a1), P(v1), len(d1), P(n1), P(s1), P(i1), len(n1), P(d2), P(a2), P(v2), len(d2), P(n2), P(s2), P(i2),
                                                     len(n2), C.c_float(ratio), int(ori), P(out))
    return n, out[:len(d1)]


def test_node_oracles_reduce_to_single_node_oracles():
    """one all-inclusive node + good map points everywhere: the node-wise restatements must equal the brute-force ones that the
    extractor / matcher parity tests already pin"""
    rng = np.random.default_rng(21)
    base = rng.integers(0, 256, (12, 32)).astype(np.uint8)
    d1 = np.repeat(base, 25, axis=0) ^ np.packbits(rng.integers(0, 100, (300, 256)) < 4, axis=1)
    d2 = np.repeat(base, 20, axis=0) ^ np.packbits(rng.integers(0, 100, (240, 256)) < 4, axis=1)
    a1 = rng.uniform(0, 360, 300).astype(np.float32); a2 = rng.uniform(0, 360, 240).astype(np.float32)
    fv1 = {7: list(range(300))}; fv2 = {7: list(range(240))}
    one1 = np.ones(300, np.uint8); one2 = np.ones(240, np.uint8)
    for ratio, ori in ((0.6, True), (0.9, False), (1.5, True)):
        n, m = oracle_bow_nodes(0, d1, a1, one1, fv1, d2, a2, one2, fv2, ratio, ori)
        nb, mb = oracle.search_by_bow_bf(d1, a1, d2, a2, ratio, ori)
        assert n == nb and np.array_equal(m, mb)
        n, m = oracle_bow_nodes(1, d1, a1, one1, fv1, d2, a2, one2, fv2, ratio, ori)
        want = np.zeros(300, np.int32)
        nw = oracle.lib().oracle_search_by_bow_kfkf_bf(P(d1), P(a1), 300, P(d2), P(a2), 240, C.c_float(ratio), int(ori), P(want))
        assert n == nw and np.array_equal(m, want)
    # disjoint nodes never match; a node present on one side only is skipped by the lower_bound branches
    n, m = oracle_bow_nodes(0, d1, a1, one1, {1: list(range(150)), 5: list(range(150, 300))}, d2, a2, one2, {2: list(range(100)), 6: list(range(100, 240))}, 0.9, True)
    assert n == 0 and (m == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 2), (6, 4, 2), (10, 3, 1)])
def test_search_by_bow_over_feature_vectors_matches_oracle(built_lib, k, L, levelsup):
    """SearchByBoW(KF, F) (ORBmatcher.cc:159-292) and SearchByBoW(KF, KF) (ORBmatcher.cc:526-659) with the FeatureVectors of a
    vocabulary, random good-MapPoint masks; bit-exact against the node-wise oracles"""
    from orb_slam2_aruco_b200 import synth
    from orb_slam2_aruco_b200.api import ORBVocabulary, ORBextractor, ORBmatcher
    rng = np.random.default_rng(300 + k + L)
    tree = make_tree(rng, k, L)
    voc = ORBVocabulary(k, L, *tree)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    img = synth.make_frame(96)
    k1, d1 = ex(img); k2, d2 = ex(np.roll(img, (2, 3), axis=(0, 1)))
    # near-duplicates of keyframe descriptors on the frame side so that ties and second-best rejections occur inside nodes
    extra = d1[rng.integers(0, len(d1), 200)] ^ np.packbits(rng.integers(0, 100, (200, 256)) < 2, axis=1)
    d2 = np.concatenate([d2, extra]); a2 = np.concatenate([k2["angle"], rng.uniform(0, 360, 200).astype(np.float32)])
    a1 = k1["angle"]
    _, fv1 = voc.transform(d1, levelsup); _, fv2 = voc.transform(d2, levelsup)
    assert len(set(fv1) & set(fv2)) > 1
    v1 = (rng.random(len(d1)) < 0.8).astype(np.uint8); v2 = (rng.random(len(d2)) < 0.8).astype(np.uint8)
    total = 0
    for ratio, ori in ((0.6, True), (0.75, False), (0.9, True), (1.5, True)):
        m = ORBmatcher(ratio, ori)
        n, got = m.SearchByBoW_nodes(d1, a1, v1, fv1, d2, a2, fv2)
        nw, want = oracle_bow_nodes(0, d1, a1, v1, fv1, d2, a2, v2, fv2, ratio, ori)
        assert n == nw and np.array_equal(got, want), (ratio, ori)
        total += n
        n, got = m.SearchByBoW_KF_nodes(d1, a1, v1, fv1, d2, a2, v2, fv2)
        nw, want = oracle_bow_nodes(1, d1, a1, v1, fv1, d2, a2, v2, fv2, ratio, ori)
        assert n == nw and np.array_equal(got, want), (ratio, ori)
        total += n
    assert total > 50
    # empty sides
    n, got = ORBmatcher(0.9, True).SearchByBoW_nodes(d1, a1, v1, fv1, d2[:0], a2[:0], {})
    assert n == 0 and len(got) == 0
    n, got = ORBmatcher(0.9, True).SearchByBoW_KF_nodes(d1, a1, np.zeros_like(v1), fv1, d2, a2, v2, fv2)
    assert n == 0 and (got == -1).all()
    voc.close(); ex.close()
