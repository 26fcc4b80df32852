"""GPU: CUDA matcher through the C-ABI vs the CPU restatement of ORBmatcher (bit-exact match indices)."""
import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import synth
from orb_slam2_aruco_b200.api import ORBmatcher

pytestmark = pytest.mark.gpu


def test_hamming_matrix(built_lib):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (37, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (53, 32)).astype(np.uint8)
    want = np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(axis=2)
    assert np.array_equal(ORBmatcher().distance_matrix(a, b), want)
    assert ORBmatcher.DescriptorDistance(a[0], a[0]) == 0
    assert ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def _frames_desc(n, first):
    out = []
    for i in range(n):
        k, d = oracle.orb_extract(synth.make_frame(first + i))
        out.append((k, d))
    return out


def test_search_by_bow_bruteforce_real_descriptors(built_lib):
    # reference set: a shifted copy of the scene; frames: the scene + two unrelated ones
    kr, dr = oracle.orb_extract(np.roll(synth.make_frame(200), (3, 5), axis=(0, 1)))
    frames = [oracle.orb_extract(synth.make_frame(i)) for i in (200, 200, 201)]
    cap = max(len(k) for k, _ in frames) + 7
    fd = np.zeros((3, cap, 32), np.uint8); fa = np.zeros((3, cap), np.float32); nf = np.zeros(3, np.int32)
    for i, (k, d) in enumerate(frames):
        fd[i, :len(d)] = d; fa[i, :len(k)] = k["angle"]; nf[i] = len(k)
    for ratio, ori in ((0.7, True), (0.9, False), (0.6, True)):
        m = ORBmatcher(ratio, ori)
        nm, matches = m.SearchByBoW_batch(dr, kr["angle"], fd, fa, nf)
        for i, (k, d) in enumerate(frames):
            n2, m2 = oracle.search_by_bow_bf(dr, kr["angle"], d, k["angle"], ratio, ori)
            assert nm[i] == n2
            assert np.array_equal(matches[i, :len(d)], m2)
            assert (matches[i, len(d):] == -1).all()
    assert nm[0] > 100          # the shifted scene really matches


def test_greedy_conflicts_and_ties(built_lib):
    """many near-duplicate descriptors: earlier reference rows take candidates away from later ones and the
    top-K lists run dry, which forces the rescan path"""
    rng = np.random.default_rng(4)
    base = rng.integers(0, 256, (8, 32)).astype(np.uint8)
    ref = np.repeat(base, 40, axis=0)                  # 320 reference descriptors, 40 copies of each
    frame = np.repeat(base, 25, axis=0)                # 200 frame descriptors
    flip = rng.integers(0, 256, frame.shape) < 6       # sprinkle bit noise
    frame = frame ^ np.packbits(rng.integers(0, 100, (200, 256)) < 3, axis=1)
    ra = rng.uniform(0, 360, len(ref)).astype(np.float32); fa = rng.uniform(0, 360, len(frame)).astype(np.float32)
    for ratio, ori in ((0.7, True), (1.5, False), (1.5, True)):
        n2, m2 = oracle.search_by_bow_bf(ref, ra, frame, fa, ratio, ori)
        n, m = ORBmatcher(ratio, ori).SearchByBoW(ref, ra, frame, fa)
        assert n == n2 and np.array_equal(m, m2)


def test_empty_sets(built_lib):
    d = np.zeros((0, 32), np.uint8); a = np.zeros(0, np.float32)
    rng = np.random.default_rng(1)
    x = rng.integers(0, 256, (10, 32)).astype(np.uint8); xa = np.zeros(10, np.float32)
    n, m = ORBmatcher(0.7).SearchByBoW(d, a, x, xa)
    assert n == 0 and (m == -1).all()
    n, m = ORBmatcher(0.7).SearchByBoW(x, xa, x[:1], xa[:1])
    assert n == 0 or m[0] == 0      # a single candidate has bestDist2 = 256


def test_candidate_lists(built_lib):
    rng = np.random.default_rng(2)
    q = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    t = rng.integers(0, 256, (500, 32)).astype(np.uint8)
    t[:100] = q[:100] ^ np.packbits(rng.integers(0, 100, (100, 256)) < 5, axis=1)
    lens = rng.integers(0, 60, 300)
    lens[5] = 0
    ofs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = rng.integers(0, 500, ofs[-1]).astype(np.int32)
    got = ORBmatcher().match_candidates(q, t, ofs, cand)
    want = oracle.match_candidates(q, t, ofs, cand)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_search_for_initialization_matches_oracle(built_lib):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:409-524): two views of one scene (second one shifted), windowed
    candidates from the device feature grid, sequential accept / steal / histogram logic; and a second call that continues
    from the updated vbPrevMatched like Tracking::MonocularInitialization does"""
    import ctypes as C
    from orb_slam2_aruco_b200 import synth
    from orb_slam2_aruco_b200.api import ORBextractor, search_for_initialization
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    a = synth.make_frame(70)
    vp = C.c_void_p
    bounds = np.array([0, 640, 0, 480], np.float32)
    for shift, window in (((4, 7), 100), ((25, 12), 30), ((1, 1), 10)):
        b = np.roll(a, shift, axis=(0, 1))
        k1, d1 = ex(a); k2, d2 = ex(b)
        prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
        for rep in range(2):
            want_prev = prev.copy(); want = np.zeros(len(k1), np.int32)
            nw = oracle.lib().oracle_search_for_initialization(np.ascontiguousarray(k1).ctypes.data_as(vp), np.ascontiguousarray(d1).ctypes.data_as(vp), len(k1),
                                                               np.ascontiguousarray(k2).ctypes.data_as(vp), np.ascontiguousarray(d2).ctypes.data_as(vp), len(k2),
                                                               bounds.ctypes.data_as(vp), want_prev.ctypes.data_as(vp), window, C.c_float(0.9), 1,
                                                               want.ctypes.data_as(vp))
            n, m12, new_prev = search_for_initialization(k1, d1, k2, d2, bounds, prev, window, 0.9, True)
            assert n == nw and np.array_equal(m12, want) and np.array_equal(new_prev, want_prev)
            assert (m12[k1["octave"] > 0] == -1).all()
            prev = new_prev
        if window >= 30:
            assert n > 40                               # the shifted view really matches
    n, m12, _ = search_for_initialization(k1[:0], d1[:0], k2, d2, bounds, prev[:0])
    assert n == 0 and len(m12) == 0
    ex.close()


def test_search_by_bow_keyframe_pair_matches_oracle(built_lib):
    """SearchByBoW(KeyFrame*, KeyFrame*) (ORBmatcher.cc:526-659), brute-force node: strict < TH_LOW, factor 1/HISTO_LENGTH"""
    import ctypes as C
    from orb_slam2_aruco_b200 import synth
    from orb_slam2_aruco_b200.api import ORBextractor
    vp = C.c_void_p
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    a = synth.make_frame(80)
    k1, d1 = ex(a); k2, d2 = ex(np.roll(a, (3, 5), axis=(0, 1)))
    rng = np.random.default_rng(8)
    base = rng.integers(0, 256, (8, 32)).astype(np.uint8)
    dup1 = np.repeat(base, 30, axis=0); dup2 = np.repeat(base, 20, axis=0) ^ np.packbits(rng.integers(0, 100, (160, 256)) < 3, axis=1)
    cases = [(d1, k1["angle"], d2, k2["angle"]), (dup1, rng.uniform(0, 360, 240).astype(np.float32), dup2, rng.uniform(0, 360, 160).astype(np.float32))]
    for da, aa, db, ab in cases:
        for ratio, ori in ((0.6, True), (0.8, False), (1.5, True)):
            want = np.zeros(len(da), np.int32)
            nw = oracle.lib().oracle_search_by_bow_kfkf_bf(np.ascontiguousarray(da).ctypes.data_as(vp), np.ascontiguousarray(aa).ctypes.data_as(vp), len(da),
                                                           np.ascontiguousarray(db).ctypes.data_as(vp), np.ascontiguousarray(ab).ctypes.data_as(vp), len(db),
                                                           C.c_float(ratio), int(ori), want.ctypes.data_as(vp))
            n, m12 = ORBmatcher(ratio, ori).SearchByBoW_KF(da, aa, db, ab)
            assert n == nw and np.array_equal(m12, want), (ratio, ori)
    ex.close()


def test_search_by_projection_modes_match_oracle(built_lib):
    """SearchByProjection(Frame, mapPoints, th) and SearchByProjection(Current, Last, th, mono) on ready-made projections: the map
    points are the keypoints of a second view whose positions are perturbed, so that windows overlap and assignments compete"""
    import ctypes as C
    from orb_slam2_aruco_b200 import synth
    from orb_slam2_aruco_b200.api import ORBextractor, search_by_projection
    vp = C.c_void_p
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(vp)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    a = synth.make_frame(90)
    kf, df = ex(a)
    kq, dq = ex(np.roll(a, (2, 3), axis=(0, 1)))
    rng = np.random.default_rng(12)
    sf = 1.2 ** kq["octave"].astype(np.float32)
    bounds = np.array([0, 640, 0, 480], np.float32)
    for mode, ratio, ori, th_high in ((0, 0.8, True, 100), (0, 0.6, False, 100), (1, 0.9, True, 100), (1, 0.9, False, 100), (1, 0.9, True, 64)):
        nq = len(kq)
        xyr = np.stack([kq["x"] - 3 + rng.normal(0, 2, nq), kq["y"] - 2 + rng.normal(0, 2, nq),
                        (rng.choice([2.5, 4.0], nq) if mode == 0 else np.full(nq, 7.0)) * sf], 1).astype(np.float32)
        lev = (np.stack([kq["octave"] - 1, kq["octave"]], 1) if mode == 0 else np.stack([kq["octave"] - 1, kq["octave"] + 1], 1)).astype(np.int32)
        observed = (rng.random(nq) < 0.8).astype(np.uint8)
        occupied = (rng.random(len(kf)) < 0.1).astype(np.uint8)
        want_occ = occupied.copy(); want = np.zeros(len(kf), np.int32)
        nw = oracle.lib().oracle_search_by_projection(P(kf), P(df), len(kf), P(bounds), P(want_occ), P(xyr), P(lev), P(dq), P(kq["angle"].copy()), P(observed), nq,
                                                      mode, C.c_float(ratio), int(ori), th_high, P(want))
        n, assign, occ = search_by_projection(kf, df, bounds, occupied, xyr, lev, dq, kq["angle"].copy(), observed, mode, ratio, ori, th_high)
        assert n == nw and np.array_equal(assign, want) and np.array_equal(occ, want_occ), (mode, ratio, ori)
        assert n > 300
    n, assign, occ = search_by_projection(kf, df, bounds, np.zeros(len(kf), np.uint8), np.zeros((0, 3)), np.zeros((0, 2)), np.zeros((0, 32)), np.zeros(0), np.zeros(0), 0)
    assert n == 0 and (assign == -1).all()
    ex.close()


def test_distinctive_descriptors_match_oracle(built_lib):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:271-331), batched: register path (N <= 32) and shared-memory rows (N > 32)"""
    import ctypes as C
    rng = np.random.default_rng(17)
    sizes = [0, 1, 2, 3, 5, 8, 31, 32, 33, 64, 100, 257] + rng.integers(1, 40, 300).tolist()
    obs = []
    for n in sizes:
        base = rng.integers(0, 256, (3, 32)).astype(np.uint8)
        obs.append(np.ascontiguousarray(base[rng.integers(0, 3, n)] ^ np.packbits(rng.integers(0, 100, (n, 256)) < 5, axis=1)) if n else np.zeros((0, 32), np.uint8))
    obs.append(np.repeat(rng.integers(0, 256, (1, 32)).astype(np.uint8), 40, axis=0))          # all medians tie at 0
    m = ORBmatcher(0.7, True)
    best, desc = m.ComputeDistinctiveDescriptors(obs)
    f = oracle.lib().oracle_distinctive_descriptor
    for p, o in enumerate(obs):
        want = f(o.ctypes.data_as(C.c_void_p), len(o))
        assert best[p] == want, (p, len(o))
        assert np.array_equal(desc[p], o[want] if want >= 0 else np.zeros(32, np.uint8))
    small = [o for o in obs if len(o) <= 32]                                                   # a batch that never needs the row buffer
    b2, _ = m.ComputeDistinctiveDescriptors(small)
    assert np.array_equal(b2, [f(o.ctypes.data_as(C.c_void_p), len(o)) for o in small])
    b3, d3 = m.ComputeDistinctiveDescriptors([])
    assert len(b3) == 0 and len(d3) == 0


def test_reference_golden_vectors(built_lib, golden_dir):
    """tests/golden/match_ref.npz: answers of the reference's OWN src/ORBmatcher.cc (oracle/_ref/libref_match.so, written by
    tests/golden/make_match_golden.py) for DescriptorDistance, both SearchByBoW over feature vectors, SearchForInitialization and the three
    Tracking-thread SearchByProjection variants (on the grid queries the reference itself made) - the CUDA path must give the same bits"""
    import os
    import match_cases as mc
    from orb_slam2_aruco_b200.api import search_by_projection, search_for_initialization
    g = mc.load_golden(os.path.join(golden_dir, "match_ref.npz"))
    d = g["dist"]
    assert np.array_equal(np.diagonal(ORBmatcher().distance_matrix(d["a"], d["b"])), d["d"])
    c = g["bow"]
    fv1 = mc.fv_dict(c["n1"], c["s1"], c["i1"]); fv2 = mc.fv_dict(c["n2"], c["s2"], c["i2"])
    for run in c["runs"]:
        m = ORBmatcher(float(run["cfg"][0]), bool(run["cfg"][1]))
        n, got = m.SearchByBoW_nodes(c["d1"], c["a1"], c["v1"], fv1, c["d2"], c["a2"], fv2)
        assert n == run["n"] and np.array_equal(got, run["matches"])
        n, got = m.SearchByBoW_KF_nodes(c["d1"], c["a1"], c["v1"], fv1, c["d2"], c["a2"], c["v2"], fv2)
        assert n == run["n_kfkf"] and np.array_equal(got, run["matches12"])
    for run in g["bf"]["runs"]:                               # the benched brute-force kernels (k_match_topk + k_match_resolve)
        m = ORBmatcher(float(run["cfg"][0]), bool(run["cfg"][1]))
        n, got = m.SearchByBoW(c["d1"], c["a1"], c["d2"], c["a2"])
        assert n == run["n"] and np.array_equal(got[:len(c["d2"])], run["matches"])
        n, got = m.SearchByBoW_KF(c["d1"], c["a1"], c["d2"], c["a2"])
        assert n == run["n_kfkf"] and np.array_equal(got, run["matches12"])
    c = g["init"]
    prev = None
    for j, run in enumerate(c["runs"]):
        prev = c["prev"] if j % 2 == 0 else prev
        n, m12, prev = search_for_initialization(c["k1"], c["d1"], c["k2"], c["d2"], mc.BOUNDS, prev, int(run["cfg"][0]), float(run["cfg"][1]), bool(run["cfg"][2]))
        assert n == run["n"] and np.array_equal(m12, run["matches12"]) and np.array_equal(prev, run["prev"])
    for kind in ("points", "last", "reloc"):
        c = g[kind]
        for run in c["runs"]:
            occ, qd, qa, qo, mode = mc.projection_queries(kind, c, run["q_mp"])
            ratio = float(run["cfg"][1]) if kind == "points" else 0.9
            ori = True if kind == "points" else bool(run["cfg"][-1])
            th_high = int(run["cfg"][1]) if kind == "reloc" else 100
            n, assign, _ = search_by_projection(c["k2"], c["d2"], mc.BOUNDS, occ, run["q_xyr"], run["q_lev"], qd, qa, qo, mode, ratio, ori, th_high)
            assert n == run["n"] and np.array_equal(assign, mc.expected_assign(run["assign"], run["q_mp"])), (kind, run["cfg"])
