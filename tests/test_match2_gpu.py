"""GPU: the KeyFrame-side matcher members through the C-ABI (k_triang, k_triang_finish, k_radius_best, and the projection resolve kernel for the
loop-closing SearchByProjection) against the reference's own answers (tests/golden/match_ref2.npz) and against the CPU oracle on full-size inputs."""
import ctypes as C
import os

import numpy as np
import pytest

import match_cases as mc
import match_cases2 as m2
import oracle
from orb_slam2_aruco_b200 import kfgeom
from orb_slam2_aruco_b200.api import ORBmatcher

pytestmark = pytest.mark.gpu
P, A = mc.P, mc.A


@pytest.fixture(scope="module")
def golden(golden_dir):
    return mc.load_golden(os.path.join(golden_dir, "match_ref2.npz"))


def test_reference_answers_through_the_device(golden, built_lib):
    m2.replay_product(ORBmatcher(0.6, True), golden)


def test_radius_search_matches_oracle(built_lib):
    rng = np.random.default_rng(11)
    k, d, kq, dq = mc.two_views(97, (1, 2))
    nq = len(kq)
    q3 = A(np.stack([kq["x"] + rng.normal(0, 2, nq), kq["y"] + rng.normal(0, 2, nq), rng.choice([2.5, 4.0, 12.0, 40.0], nq) * 1.2 ** kq["octave"]], 1), np.float32)
    q3[:5, 0] += 2000                                             # windows outside the grid
    ql = A(np.clip(kq["octave"] + rng.integers(-1, 2, nq), 0, 7), np.int32)
    M = ORBmatcher()
    for chi2 in (0.0, 5.99, 0.5):
        bi, bd = M.kf_radius_search(k, d, mc.BOUNDS, q3, ql, dq, chi2=chi2)
        wi, wd = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
        oracle.lib().oracle_kf_radius_search(P(k), P(d), len(k), P(mc.BOUNDS), P(q3), P(ql), P(dq), nq, C.c_float(1.2), 8, C.c_double(chi2), P(wi), P(wd))
        assert np.array_equal(bi, wi) and np.array_equal(bd, wd)
        assert (wi >= 0).sum() > (50 if chi2 == 0.5 else 300) and (wi[:5] == -1).all()
    # no queries / an empty keyframe
    bi, bd = M.kf_radius_search(k, d, mc.BOUNDS, q3[:0], ql[:0], dq[:0])
    assert len(bi) == 0
    bi, bd = M.kf_radius_search(k[:0], d[:0], mc.BOUNDS, q3[:7], ql[:7], dq[:7])
    assert (bi == -1).all() and (bd == 256).all()


def test_triangulation_matches_oracle_on_full_size_keyframes(built_lib):
    assert mc.NFEATURES == 1000
    for seed in (3, 12):
        c = m2.triangulation_inputs(seed=seed)
        for ori in (False, True):
            want_n, want = m2.run_triangulation(oracle.lib(), "oracle", c, ori)
            n, m12 = ORBmatcher(0.6, ori).SearchForTriangulation(c["k1"], c["d1"], c["has1"], m2.fv_of(c, "1"), c["T1"], c["k2"], c["d2"], c["has2"],
                                                                 m2.fv_of(c, "2"), c["T2"], mc.CAM4, c["F12"])
            assert n == want_n and np.array_equal(m12, want) and n > 40
    # nothing in common: no groups
    n, m12 = ORBmatcher(0.6, True).SearchForTriangulation(c["k1"], c["d1"], c["has1"], {1: [0, 1]}, c["T1"], c["k2"], c["d2"], c["has2"], {2: [0, 1]}, c["T2"],
                                                          mc.CAM4, c["F12"])
    assert n == 0 and (m12 == -1).all()


def test_fuse_and_sim3_match_oracle_on_full_size_keyframes(built_lib):
    O, M = oracle.lib(), ORBmatcher(0.6, True)
    c = m2.keyframe_points_inputs(seed=4)
    for th in (3.0, 8.0):
        want = m2.run_fuse(O, "oracle", c, th)
        got = M.Fuse(c["k"], c["d"], mc.BOUNDS, mc.CAM4, c["T"], c["held_state"], c["held_nobs"], c["mp_state"], c["mp_pos"], c["mp_normal"], c["mp_desc"],
                     c["mp_minmax"], c["mp_nobs"], th)
        assert got[0] == want[0] and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]) and got[0] > 200
    c = m2.keyframe_points_inputs(seed=6, sim3=True)
    for th in (4.0, 12.0):
        want = m2.run_fuse_sim3(O, "oracle", c, th)
        st = np.where(c["mp_state"] == 0, 1, c["mp_state"]).astype(np.uint8)
        got = M.FuseSim3(c["k"], c["d"], mc.BOUNDS, mc.CAM4, c["T"], c["held_state"], st, c["mp_pos"], c["mp_normal"], c["mp_desc"], c["mp_minmax"], th)
        assert got[0] == want[0] and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]) and got[0] > 200
        want = m2.run_loop(O, "oracle", c, int(th))
        st = np.where(c["mp_state"] == 2, 2, 1).astype(np.uint8)
        got = M.SearchByProjectionLoop(c["k"], c["d"], mc.BOUNDS, mc.CAM4, c["T"], st, c["mp_pos"], c["mp_normal"], c["mp_desc"], c["mp_minmax"],
                                       m2.loop_matched(c), int(th))
        assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[0] > 100
    c = m2.sim3_inputs(seed=8)
    for th in (7.5, 2.0):
        want = m2.run_sim3(O, "oracle", c, th)
        got = M.SearchBySim3(c["k1"], c["d1"], c["T1"], c["st1"], c["p1"], c["d1"], c["mm1"], c["k2"], c["d2"], c["T2"], c["st2"], c["p2"], c["d2"], c["mm2"],
                             mc.BOUNDS, mc.CAM4, c["m12"], float(c["s12"]), c["R12"], c["t12"], th)
        assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[0] > 100


def test_device_projection_equals_the_reference_arithmetic(golden, built_lib):
    """k_kf_project against the numpy model of the reference's cv::Mat statements (itself pinned bit for bit to the queries the reference made, CPU test)
    and directly against those traced queries"""
    M = ORBmatcher()
    for seed, sim3 in ((4, False), (6, True)):
        c = m2.keyframe_points_inputs(seed=seed, sim3=sim3)
        pose = kfgeom.pose_from_S(c["T"]) if sim3 else kfgeom.pose_from_T(c["T"])
        for th in (3.0, 10.0):
            v, q3, lv = M.project_points(pose, mc.CAM4, mc.BOUNDS, c["mp_pos"], c["mp_normal"], c["mp_minmax"], th)
            wv, wq, wl = m2.host_project_points(pose, mc.CAM4, mc.BOUNDS, c["mp_pos"], c["mp_normal"], c["mp_minmax"], th)
            assert np.array_equal(v, wv) and v.sum() > 500 and (~v).sum() > 50
            assert np.array_equal(q3[v].view(np.uint32), wq[v].view(np.uint32)) and np.array_equal(lv[v], wl[v])
    c = m2.sim3_inputs(seed=8)
    sR12, sR21, t21 = kfgeom.sim3_between(c["s12"], c["R12"], c["t12"])
    for T, sR, tt, pos, mm in ((c["T1"], sR21, t21, c["p1"], c["mm1"]), (c["T2"], sR12, c["t12"], c["p2"], c["mm2"])):
        v, q3, lv = M.project_points(kfgeom.pose_from_T(T), mc.CAM4, mc.BOUNDS, pos, None, mm, 7.5, sim3=(sR, tt))
        wv, wq, wl = m2.host_project_points_sim3(kfgeom.pose_from_T(T), sR, tt, mc.CAM4, mc.BOUNDS, pos, mm, 7.5)
        assert np.array_equal(v, wv) and v.sum() > 500
        assert np.array_equal(q3[v].view(np.uint32), wq[v].view(np.uint32)) and np.array_equal(lv[v], wl[v])
    # the traced queries of the reference's own Fuse
    g = golden["fuse"]
    for run in g["runs"]:
        v, q3, lv = M.project_points(kfgeom.pose_from_T(g["T"]), mc.CAM4, mc.BOUNDS, g["mp_pos"], g["mp_normal"], g["mp_minmax"], float(run["cfg"][0]))
        qs = np.nonzero(v & (g["mp_state"] == 1))[0]
        assert np.array_equal(q3[qs].view(np.uint32), A(run["q_xyr"], np.float32).view(np.uint32)) and np.array_equal(lv[qs], run["q_lev"][:, 1])
    v, q3, lv = M.project_points(kfgeom.pose_from_T(g["T"]), mc.CAM4, mc.BOUNDS, g["mp_pos"][:0], g["mp_normal"][:0], g["mp_minmax"][:0], 3.0)
    assert len(v) == 0


def test_fused_projection_and_search_equals_the_two_steps(built_lib):
    M = ORBmatcher()
    c = m2.keyframe_points_inputs(seed=4)
    pose = kfgeom.pose_from_T(c["T"])
    skip = c["mp_state"] != 1
    for chi2, nrm in ((5.99, c["mp_normal"]), (0.0, None)):
        v, q3, lv = M.project_points(pose, mc.CAM4, mc.BOUNDS, c["mp_pos"], nrm, c["mp_minmax"], 3.0)
        v &= ~skip
        qs = np.nonzero(v)[0]
        wi, wd = M.kf_radius_search(c["k"], c["d"], mc.BOUNDS, q3[qs], lv[qs], c["mp_desc"][qs], chi2=chi2)
        fv, bi, bd = M.search_points(c["k"], c["d"], mc.BOUNDS, pose, mc.CAM4, c["mp_pos"], nrm, c["mp_minmax"], c["mp_desc"], skip, 3.0, chi2=chi2)
        assert np.array_equal(fv, v) and np.array_equal(bi[qs], wi) and np.array_equal(bd[qs], wd) and (wi >= 0).sum() > 200
        assert (bi[~v] == -1).all() and (bd[~v] == 256).all()
    fv, bi, bd = M.search_points(c["k"][:0], c["d"][:0], mc.BOUNDS, pose, mc.CAM4, c["mp_pos"], None, c["mp_minmax"], c["mp_desc"], skip, 3.0)
    assert (bi == -1).all() and fv.sum() > 300
