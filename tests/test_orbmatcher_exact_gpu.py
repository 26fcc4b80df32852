"""GPU: ORB_SLAM2::ORBmatcher with the reference's EXACT member signatures (include/b200slam_orbmatcher.hpp) against the reference's own src/ORBmatcher.cc.

oracle/ref_match_wrap.cpp calls the ten members the way src/Tracking.cc / LocalMapping.cc / LoopClosing.cc call them - matcher.SearchByBoW(&kf, F, vpMapPointMatches),
matcher.SearchByProjection(CurrentFrame, LastFrame, th, true), matcher.Fuse(&kf, vpMapPoints, th) ... on Frame / KeyFrame / MapPoint objects.  The SAME file is compiled
twice: with the reference's ORBmatcher.h + ORBmatcher.cc (oracle/_ref/libref_match.so) and with the product header over libb200slam.so
(oracle/_ref/libadapter_match.so).  Every answer the two libraries write into the caller's objects must be equal: match counts, vpMapPointMatches, F.mvpMapPoints,
vnMatches12 / vbPrevMatched, vMatchedPairs, vpMatched, vpReplacePoint, the Replace / AddObservation outcomes of Fuse.  Where the reference library is absent the
committed answers of tests/golden/match_ref.npz / match_ref2.npz (written from it) are replayed instead."""
import os

import numpy as np
import pytest

import match_cases as mc
import match_cases2 as m2
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def adapter(built_lib):
    a = oracle.adapter_match()
    if a is None:
        pytest.skip("oracle/_ref/libadapter_match.so not built (needs /root/reference for the DBoW2 headers of the stand-ins)")
    return a


def same(a, b, k):
    assert a[0] == b[0], (a[0], b[0])
    for x, y in zip(a[1:k], b[1:k]):
        assert np.array_equal(x, y)


def test_descriptor_distance_is_a_host_popcount(adapter):
    rng = np.random.default_rng(3)
    for _ in range(50):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert adapter.ref_descriptor_distance(mc.P(a), mc.P(b)) == int(np.unpackbits(a ^ b).sum())


def test_tracking_thread_members_replay_the_golden_answers(adapter, golden_dir):
    g = mc.load_golden(os.path.join(golden_dir, "match_ref.npz"))
    c = g["bow"]
    for run in c["runs"]:
        ratio, ori = float(run["cfg"][0]), int(run["cfg"][1])
        n, m = mc.run_bow(adapter, "ref", c, ratio, ori)                  # Tracking.cc:917-920, 1778
        assert n == run["n"] and np.array_equal(m, run["matches"])
        n, m = mc.run_bow_kfkf(adapter, "ref", c, ratio, ori)             # LoopClosing.cc:400
        assert n == run["n_kfkf"] and np.array_equal(m, run["matches12"])
    c = g["init"]
    prev = None
    for j, run in enumerate(c["runs"]):
        prev = c["prev"] if j % 2 == 0 else prev
        n, m, prev = mc.run_init(adapter, "ref", c, prev, int(run["cfg"][0]), float(run["cfg"][1]), int(run["cfg"][2]))      # Tracking.cc:531-532
        assert n == run["n"] and np.array_equal(m, run["matches12"]) and np.array_equal(prev, run["prev"])
    for kind, fn in (("points", mc.ref_points), ("last", mc.ref_last), ("reloc", mc.ref_reloc)):                             # Tracking.cc:1515, 1011-1017, 1858 / 1875
        c = g[kind]
        for run in c["runs"]:
            cfg = [float(v) for v in run["cfg"]]
            if kind == "last":
                got = fn(adapter, c, cfg[0], int(cfg[1]))
            elif kind == "reloc":
                got = fn(adapter, c, cfg[0], int(cfg[1]), int(cfg[2]))
            else:
                got = fn(adapter, c, cfg[0], cfg[1])
            assert got[0] == run["n"] and np.array_equal(got[1], run["assign"]), (kind, run["cfg"])


def test_mapping_and_loop_closing_members_replay_the_golden_answers(adapter, golden_dir):
    g = mc.load_golden(os.path.join(golden_dir, "match_ref2.npz"))
    c = g["tri"]
    for run in c["runs"]:
        n, m = m2.run_triangulation(adapter, "ref", c, int(run["cfg"][0]))[:2]                                               # LocalMapping.cc:283
        assert n == run["n"] and np.array_equal(m, run["matches12"])
    c = g["fuse"]
    for run in c["runs"]:
        n, idx, act = m2.run_fuse(adapter, "ref", c, float(run["cfg"][0]))[:3]                                               # LocalMapping.cc:857, 882
        assert n == run["n"] and np.array_equal(idx, run["fused_idx"]) and np.array_equal(act, run["action"])
    c = g["scw"]
    for run in c["runs"]:
        n, rep, add = m2.run_fuse_sim3(adapter, "ref", c, float(run["cfg"][0]))[:3]                                          # LoopClosing.cc:1086
        assert n == run["n"] and np.array_equal(rep, run["replace_idx"]) and np.array_equal(add, run["added_idx"])
        n, matched = m2.run_loop(adapter, "ref", c, int(run["cfg"][1]))[:2]                                                  # LoopClosing.cc:458, 629
        assert n == run["n_loop"] and np.array_equal(matched, run["matched"])
    c = g["sim3"]
    for run in c["runs"]:
        n, m12 = m2.run_sim3(adapter, "ref", c, float(run["cfg"][0]))[:2]                                                    # LoopClosing.cc:424, 577
        assert n == run["n"] and np.array_equal(m12, run["matches12"])


@pytest.mark.skipif(oracle.ref_match() is None, reason="oracle/_ref/libref_match.so not built (needs /root/reference)")
def test_live_against_the_reference_on_full_size_frames(adapter):
    """fresh seeds, 1000-feature frames and keyframes: both libraries are driven through the identical wrapper code on identical stand-in objects"""
    R = oracle.ref_match()
    for seed in (5, 17):
        c = mc.bow_inputs(seed=seed)
        for ratio, ori in ((0.6, True), (0.8, False), (1.2, True)):
            same(mc.run_bow(R, "ref", c, ratio, ori), mc.run_bow(adapter, "ref", c, ratio, ori), 2)
            same(mc.run_bow_kfkf(R, "ref", c, ratio, ori), mc.run_bow_kfkf(adapter, "ref", c, ratio, ori), 2)
        one = mc.one_node(c)                                            # the benched brute-force configuration
        same(mc.run_bow(R, "ref", one, 0.7, True), mc.run_bow(adapter, "ref", one, 0.7, True), 2)
    for shift, window in (((4, 7), 100), ((25, 12), 30)):
        c = mc.init_inputs(shift)
        prev = c["prev"]
        for rep in range(2):
            a = mc.run_init(R, "ref", c, prev, window, 0.9, True)
            same(a, mc.run_init(adapter, "ref", c, prev, window, 0.9, True), 3)
            prev = a[2]
    for make, run, cfgs in ((mc.points_inputs, mc.ref_points, [(1.0, 0.8), (3.0, 0.7)]), (mc.last_inputs, mc.ref_last, [(7.0, True), (15.0, False)]),
                            (mc.reloc_inputs, mc.ref_reloc, [(10.0, 100, True), (3.0, 64, True)])):
        for seed in (9, 23):
            c = make(seed=seed)
            for cfg in cfgs:
                a = run(R, c, *cfg)
                assert a[0] > 100
                same(a, run(adapter, c, *cfg), 2)
    for seed in (61, 62):
        c = m2.triangulation_inputs(seed=seed)
        for ori in (0, 1):
            same(m2.run_triangulation(R, "ref", c, ori), m2.run_triangulation(adapter, "ref", c, ori), 2)
        c = m2.keyframe_points_inputs(seed=seed + 10)
        for th in (3.0, 5.0):
            a = m2.run_fuse(R, "ref", c, th)
            assert a[0] > 50
            same(a, m2.run_fuse(adapter, "ref", c, th), 3)
        c = m2.keyframe_points_inputs(seed=seed + 20, sim3=True)
        same(m2.run_fuse_sim3(R, "ref", c, 4.0), m2.run_fuse_sim3(adapter, "ref", c, 4.0), 3)
        same(m2.run_loop(R, "ref", c, 10), m2.run_loop(adapter, "ref", c, 10), 2)
        c = m2.sim3_inputs(seed=seed + 30)
        a = m2.run_sim3(R, "ref", c, 7.5)
        assert a[0] > 50
        same(a, m2.run_sim3(adapter, "ref", c, 7.5), 2)


def test_three_threads_call_the_matcher_concurrently(adapter, golden_dir):
    """the reference calls ORBmatcher from the Tracking, LocalMapping and LoopClosing threads at once: three host threads, each on its own
    per-thread stream inside the library, must each get the single-threaded answers"""
    import threading
    g1 = mc.load_golden(os.path.join(golden_dir, "match_ref.npz"))
    g2 = mc.load_golden(os.path.join(golden_dir, "match_ref2.npz"))
    errors = []

    def tracking():
        c = g1["last"]
        for _ in range(6):
            for run in c["runs"]:
                got = mc.ref_last(adapter, c, float(run["cfg"][0]), int(run["cfg"][1]))
                if got[0] != run["n"] or not np.array_equal(got[1], run["assign"]):
                    errors.append("tracking")

    def mapping():
        c = g2["fuse"]
        for _ in range(6):
            for run in c["runs"]:
                n, idx, act = m2.run_fuse(adapter, "ref", c, float(run["cfg"][0]))[:3]
                if n != run["n"] or not np.array_equal(idx, run["fused_idx"]):
                    errors.append("mapping")

    def closing():
        c = g2["sim3"]
        for _ in range(6):
            for run in c["runs"]:
                n, m12 = m2.run_sim3(adapter, "ref", c, float(run["cfg"][0]))[:2]
                if n != run["n"] or not np.array_equal(m12, run["matches12"]):
                    errors.append("closing")
    ts = [threading.Thread(target=f) for f in (tracking, mapping, closing)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
