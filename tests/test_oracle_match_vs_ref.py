"""CPU: oracle/match_oracle.cpp against the reference's OWN src/ORBmatcher.cc.
 * replay of tests/golden/match_ref.npz (answers of oracle/_ref/libref_match.so, written by tests/golden/make_match_golden.py) - runs everywhere;
 * live comparison on larger cases where oracle/_ref exists (the dev container)."""
import os

import numpy as np
import pytest

import match_cases as mc
import oracle


@pytest.fixture(scope="module")
def golden(golden_dir):
    return mc.load_golden(os.path.join(golden_dir, "match_ref.npz"))


def test_descriptor_distance(golden):
    d = golden["dist"]
    assert [oracle.descriptor_distance(a, b) for a, b in zip(d["a"], d["b"])] == d["d"].tolist()
    assert d["d"][:8].tolist() == [0] * 8 and d["d"][8:16].tolist() == [256] * 8


def test_search_by_bow_over_feature_vectors(golden):
    c = golden["bow"]
    assert len(c["runs"]) == 4
    for run in c["runs"]:
        ratio, ori = float(run["cfg"][0]), int(run["cfg"][1])
        n, m = mc.run_bow(oracle.lib(), "oracle", c, ratio, ori)
        assert n == run["n"] and np.array_equal(m, run["matches"])
        n, m = mc.run_bow_kfkf(oracle.lib(), "oracle", c, ratio, ori)
        assert n == run["n_kfkf"] and np.array_equal(m, run["matches12"])
        assert run["n"] > 50 and run["n_kfkf"] > 50


def test_brute_force_configuration(golden):
    """one node + all MapPoints good: the reference's answers against the single-node oracle entry points the bench and the GPU tests use"""
    import ctypes as C
    c = golden["bow"]
    for run in golden["bf"]["runs"]:
        ratio, ori = float(run["cfg"][0]), int(run["cfg"][1])
        n, m = oracle.search_by_bow_bf(c["d1"], c["a1"], c["d2"], c["a2"], ratio, bool(ori))
        assert n == run["n"] and np.array_equal(m, run["matches"])
        want = np.zeros(len(c["d1"]), np.int32)
        n = oracle.lib().oracle_search_by_bow_kfkf_bf(mc.P(c["d1"]), mc.P(c["a1"]), len(c["d1"]), mc.P(c["d2"]), mc.P(c["a2"]), len(c["d2"]), C.c_float(ratio), ori, mc.P(want))
        assert n == run["n_kfkf"] and np.array_equal(want, run["matches12"])


def test_search_for_initialization(golden):
    c = golden["init"]
    prev = None
    for j, run in enumerate(c["runs"]):
        window, ratio, ori = int(run["cfg"][0]), float(run["cfg"][1]), int(run["cfg"][2])
        prev = c["prev"] if j % 2 == 0 else prev             # runs come in pairs: the second continues from the first one's vbPrevMatched
        n, m, prev = mc.run_init(oracle.lib(), "oracle", c, prev, window, ratio, ori)
        assert n == run["n"] and np.array_equal(m, run["matches12"]) and np.array_equal(prev, run["prev"])


@pytest.mark.parametrize("kind", ["points", "last", "reloc"])
def test_search_by_projection(golden, kind):
    c = golden[kind]
    for run in c["runs"]:
        occ, qd, qa, qo, mode = mc.projection_queries(kind, c, run["q_mp"])
        ratio = float(run["cfg"][1]) if kind == "points" else 0.9
        ori = True if kind == "points" else bool(run["cfg"][-1])
        th_high = int(run["cfg"][1]) if kind == "reloc" else 100
        n, assign = mc.oracle_projection(c, occ, run["q_xyr"], run["q_lev"], qd, qa, qo, mode, ratio, ori, th_high)
        assert n == run["n"] and np.array_equal(assign, mc.expected_assign(run["assign"], run["q_mp"]))
        assert run["n"] > 100 and (run["assign"] == -2).any()


@pytest.mark.skipif(oracle.ref_match() is None, reason="oracle/_ref/libref_match.so not built (needs /root/reference)")
def test_live_reference_on_full_size_frames():
    R, O = oracle.ref_match(), oracle.lib()
    assert mc.NFEATURES == 1000
    c = mc.bow_inputs(seed=5)
    for ratio, ori in ((0.6, True), (0.8, False), (1.2, True)):
        a = mc.run_bow(R, "ref", c, ratio, ori); b = mc.run_bow(O, "oracle", c, ratio, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
        a = mc.run_bow_kfkf(R, "ref", c, ratio, ori); b = mc.run_bow_kfkf(O, "oracle", c, ratio, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
    for shift, window in (((4, 7), 100), ((25, 12), 30)):
        c = mc.init_inputs(shift)
        prev = c["prev"]
        for rep in range(2):
            a = mc.run_init(R, "ref", c, prev, window, 0.9, True); b = mc.run_init(O, "oracle", c, prev, window, 0.9, True)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
            prev = a[2]
    for kind, make, run, cfgs in (("points", mc.points_inputs, mc.ref_points, [(1.0, 0.8), (3.0, 0.7)]),
                                  ("last", mc.last_inputs, mc.ref_last, [(7.0, True), (15.0, False)]),
                                  ("reloc", mc.reloc_inputs, mc.ref_reloc, [(10.0, 100, True), (3.0, 64, True)])):
        c = make(seed=9)
        for cfg in cfgs:
            n, asg, qx, ql, qm = run(R, c, *cfg)
            occ, qd, qa, qo, mode = mc.projection_queries(kind, c, qm)
            ratio = cfg[1] if kind == "points" else 0.9
            ori = True if kind == "points" else cfg[-1]
            th_high = cfg[1] if kind == "reloc" else 100
            n2, asg2 = mc.oracle_projection(c, occ, qx, ql, qd, qa, qo, mode, ratio, ori, th_high)
            assert n == n2 and n > 200 and np.array_equal(mc.expected_assign(asg, qm), asg2)


@pytest.mark.skipif(oracle.ref_match() is None, reason="oracle/_ref/libref_match.so not built (needs /root/reference)")
def test_golden_file_is_current(golden):
    """the committed vectors are what the reference answers today for the same seeded inputs"""
    R = oracle.ref_match()
    c = golden["bow"]
    run = c["runs"][0]
    n, m = mc.run_bow(R, "ref", c, float(run["cfg"][0]), int(run["cfg"][1]))
    assert n == run["n"] and np.array_equal(m, run["matches"])
    c = golden["last"]
    run = c["runs"][0]
    n, asg, qx, ql, qm = mc.ref_last(R, c, float(run["cfg"][0]), int(run["cfg"][1]))
    assert n == run["n"] and np.array_equal(asg, run["assign"]) and np.array_equal(qx, run["q_xyr"]) and np.array_equal(qm, run["q_mp"])


@pytest.mark.skipif(oracle.ref_match() is None, reason="oracle/_ref/libref_match.so not built (needs /root/reference)")
def test_reference_arm_batch_matcher_equals_oracle_batch_matcher():
    """bench.py --impl reference times the reference's own SearchByBoW (one node) through ref_search_by_bow_bf_batch; the oracle's batch driver, which
    the GPU tests and smoke() check the CUDA matcher against, gives the same matches on extractor-shaped buffers"""
    import ctypes as C
    from orb_slam2_aruco_b200 import synth
    vp = C.c_void_p
    imgs = synth.make_batch(3, 640, 480, markers=20, first=3)
    k, desc, cnt = oracle.orb_extract_batch(imgs, 1000, nthreads=3)
    kps28 = np.ascontiguousarray(k).view(np.uint8).reshape(3, -1, 28); cap = kps28.shape[1]
    rk, rd = oracle.orb_extract(np.roll(imgs[0], (3, 5), axis=(0, 1)), 1000)
    rd = np.ascontiguousarray(rd); ra = np.ascontiguousarray(rk["angle"])
    m1 = np.zeros((3, cap), np.int32); n1 = np.zeros(3, np.int32); m2 = m1.copy(); n2 = n1.copy()
    oracle.ref_match().ref_search_by_bow_bf_batch(rd.ctypes.data_as(vp), ra.ctypes.data_as(vp), len(rd), desc.ctypes.data_as(vp), kps28.ctypes.data_as(vp),
                                                  cnt.ctypes.data_as(vp), 3, cap, C.c_float(0.7), 1, m1.ctypes.data_as(vp), n1.ctypes.data_as(vp), 2)
    oracle.lib().oracle_search_by_bow_bf_batch(rd.ctypes.data_as(vp), ra.ctypes.data_as(vp), len(rd), desc.ctypes.data_as(vp), kps28.ctypes.data_as(vp),
                                               cnt.ctypes.data_as(vp), 3, cap, C.c_float(0.7), 1, C.c_float(np.float32(30.0) / np.float32(360.0)),
                                               m2.ctypes.data_as(vp), n2.ctypes.data_as(vp), 2)
    assert np.array_equal(n1, n2) and n1[0] > 300
    for f in range(3):
        assert np.array_equal(m1[f, :cnt[f]], m2[f, :cnt[f]])
