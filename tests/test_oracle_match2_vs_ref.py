"""CPU: the KeyFrame-side matcher members (SearchForTriangulation, Fuse, Fuse(Scw), loop-closing SearchByProjection, SearchBySim3).
 * oracle/match2_oracle.cpp against the reference's OWN src/ORBmatcher.cc: replay of tests/golden/match_ref2.npz (answers of
   oracle/_ref/libref_match.so, written by tests/golden/make_match_golden.py) everywhere, live on full-size cases where oracle/_ref exists;
 * the arithmetic of the device projection kernel k_kf_project, as a numpy model (tests/match_cases2.py host_project_points, on the product's per-call
   glue orb_slam2_aruco_b200/kfgeom.py), against the grid queries the reference made, bit for bit; the PredictScale threshold table against libm;
 * the ORBmatcher mirror's sequential outcome logic with its device calls swapped for CPU models (a test double, not a product path)."""
import ctypes as C
import os

import numpy as np
import pytest

import match_cases as mc
import match_cases2 as m2
import oracle
from orb_slam2_aruco_b200 import api, kfgeom

P, A = mc.P, mc.A


@pytest.fixture(scope="module")
def golden(golden_dir):
    return mc.load_golden(os.path.join(golden_dir, "match_ref2.npz"))


def test_oracle_replays_the_reference(golden):
    O = oracle.lib()
    c = golden["tri"]
    for run in c["runs"]:
        n, m = m2.run_triangulation(O, "oracle", c, int(run["cfg"][0]))
        assert n == run["n"] and np.array_equal(m, run["matches12"])
    c = golden["fuse"]
    for run in c["runs"]:
        n, idx, act = m2.run_fuse(O, "oracle", c, float(run["cfg"][0]))
        assert n == run["n"] and np.array_equal(idx, run["fused_idx"]) and np.array_equal(act, run["action"])
    c = golden["scw"]
    for run in c["runs"]:
        n, rep, add = m2.run_fuse_sim3(O, "oracle", c, float(run["cfg"][0]))
        assert n == run["n"] and np.array_equal(rep, run["replace_idx"]) and np.array_equal(add, run["added_idx"])
        n, matched = m2.run_loop(O, "oracle", c, int(run["cfg"][1]))
        assert n == run["n_loop"] and np.array_equal(matched, run["matched"])
    c = golden["sim3"]
    for run in c["runs"]:
        n, m12 = m2.run_sim3(O, "oracle", c, float(run["cfg"][0]))
        assert n == run["n"] and np.array_equal(m12, run["matches12"])


def _same_queries(valid, q3, level, run, suffix=""):
    qs = np.nonzero(valid)[0]
    assert len(qs) == len(run["q_xyr" + suffix])
    assert np.array_equal(q3[qs].view(np.uint32), A(run["q_xyr" + suffix], np.float32).view(np.uint32))          # u, v, radius: identical bits
    assert np.array_equal(level[qs], run["q_lev" + suffix][:, 1]) and np.array_equal(level[qs] - 1, run["q_lev" + suffix][:, 0])
    hit = run["q_mp" + suffix] >= 0                                # queries that found candidates name their map point
    assert np.array_equal(qs[hit], run["q_mp" + suffix][hit])
    return qs


def test_projection_model_makes_the_reference_s_queries(golden):
    c = golden["fuse"]
    for run in c["runs"]:
        valid, q3, level = m2.host_project_points(kfgeom.pose_from_T(c["T"]), mc.CAM4, mc.BOUNDS, c["mp_pos"], c["mp_normal"], c["mp_minmax"], float(run["cfg"][0]))
        # Fuse skips NULL / bad / already observed points; a point Replace()d earlier in the same call cannot come again (each is listed once)
        assert len(_same_queries(valid & (c["mp_state"] == 1), q3, level, run)) > 300
    c = golden["scw"]
    for run in c["runs"]:
        pose = kfgeom.pose_from_S(c["T"])
        valid, q3, level = m2.host_project_points(pose, mc.CAM4, mc.BOUNDS, c["mp_pos"], c["mp_normal"], c["mp_minmax"], float(run["cfg"][0]))
        _same_queries(valid & (c["mp_state"] != 2) & (c["mp_state"] != 3), q3, level, run)
        valid, q3, level = m2.host_project_points(pose, mc.CAM4, mc.BOUNDS, c["mp_pos"], c["mp_normal"], c["mp_minmax"], int(run["cfg"][1]))
        found = np.zeros(len(valid), bool); lm = m2.loop_matched(c); found[lm[lm >= 0]] = True
        _same_queries(valid & (c["mp_state"] != 2) & ~found, q3, level, run, "_loop")
    c = golden["sim3"]
    for run in c["runs"]:
        th = float(run["cfg"][0])
        sR12, sR21, t21 = kfgeom.sim3_between(c["s12"], c["R12"], c["t12"])
        v1, q1, l1 = m2.host_project_points_sim3(kfgeom.pose_from_T(c["T1"]), sR21, t21, mc.CAM4, mc.BOUNDS, c["p1"], c["mm1"], th)
        v2, q2, l2 = m2.host_project_points_sim3(kfgeom.pose_from_T(c["T2"]), sR12, c["t12"], mc.CAM4, mc.BOUNDS, c["p2"], c["mm2"], th)
        done1 = c["m12"] >= 0
        done2 = np.zeros(len(v2), bool); j = c["m12"][done1]; done2[j[c["st2"][j] > 0]] = True
        a = np.nonzero(v1 & (c["st1"] == 1) & ~done1)[0]; b = np.nonzero(v2 & (c["st2"] == 1) & ~done2)[0]
        assert len(a) + len(b) == len(run["q_xyr"]) and len(a) > 200 and len(b) > 200
        assert np.array_equal(np.concatenate([q1[a], q2[b]]).view(np.uint32), A(run["q_xyr"], np.float32).view(np.uint32))
        assert np.array_equal(np.concatenate([l1[a], l2[b]]), run["q_lev"][:, 1])


class _OracleSearches(api.ORBmatcher):
    """the mirror with its device searches answered by the CPU oracle: isolates the host logic (this is a test double, not a product path)"""

    def kf_radius_search(self, kps_un, desc, bounds4, q_xyr, q_level, q_desc, chi2=0.0, scale_factor=1.2, nlevels=8):
        k, d = A(kps_un), A(desc, np.uint8)
        q3, ql, qd = A(q_xyr, np.float32), A(q_level, np.int32), A(q_desc, np.uint8)
        bi = np.zeros(max(len(q3), 1), np.int32); bd = np.zeros(max(len(q3), 1), np.int32)
        oracle.lib().oracle_kf_radius_search(P(k), P(d), len(k), P(A(bounds4, np.float32)), P(q3), P(ql), P(qd), len(q3), C.c_float(scale_factor), nlevels,
                                             C.c_double(chi2), P(bi), P(bd))
        return bi[:len(q3)], bd[:len(q3)]

    def project_points(self, pose, cam4, bounds4, pos, normal, minmax, th, sim3=None, scale_factor=1.2, nlevels=8):
        if sim3 is None:
            return m2.host_project_points(pose, cam4, bounds4, pos, normal, minmax, th, scale_factor, nlevels)
        return m2.host_project_points_sim3(pose, sim3[0], sim3[1], cam4, bounds4, pos, minmax, th, scale_factor, nlevels)

    def search_points(self, kps_un, desc, bounds4, pose, cam4, pos, normal, minmax, q_desc, skip, th, chi2=0.0, sim3=None, scale_factor=1.2, nlevels=8):
        valid, q3, level = self.project_points(pose, cam4, bounds4, pos, normal, minmax, th, sim3, scale_factor, nlevels)
        valid = valid & ~np.asarray(skip, bool)
        qs = np.nonzero(valid)[0]
        bi = np.full(len(valid), -1, np.int32); bd = np.full(len(valid), 256, np.int32)
        bi[qs], bd[qs] = self.kf_radius_search(kps_un, desc, bounds4, q3[qs], level[qs], A(q_desc, np.uint8)[qs], chi2, scale_factor, nlevels)
        return valid, bi, bd

    def SearchForTriangulation(self, k1, d1, has1, fv1, T1, k2, d2, has2, fv2, T2, cam4, F12, scale_factor=1.2, nlevels=8):
        # the triangulation entry point is one device call; its host part is the group building and the epipole, checked here against the oracle's
        gq, qi, gc, ci = self.common_node_groups(fv1, ~np.asarray(has1, bool), fv2, ~np.asarray(has2, bool))
        assert len(qi) == sum((not has1[i]) for nd in fv1 if nd in fv2 for i in fv1[nd])
        e = kfgeom.epipole(T1, T2, cam4)
        assert np.isfinite(e).all() and 0 < e[0] < 640 and 0 < e[1] < 480
        n1, s1, i1 = mc.fv_arrays(fv1); n2, s2, i2 = mc.fv_arrays(fv2)
        c = dict(k1=A(k1), d1=A(d1), has1=A(has1, np.uint8), n1=n1, s1=s1, i1=i1, T1=A(T1, np.float32), k2=A(k2), d2=A(d2), has2=A(has2, np.uint8),
                 n2=n2, s2=s2, i2=i2, T2=A(T2, np.float32), F12=A(F12, np.float32))
        return m2.run_triangulation(oracle.lib(), "oracle", c, self.mbCheckOrientation)


def test_mirror_host_logic_with_oracle_searches(golden, monkeypatch):
    def by_projection(kps_un, desc, bounds4, occupied, q_xyr, q_levels, q_desc, q_angle, q_observed, mode, nnratio=0.8, check_ori=True, th_high=100, device=0):
        c = dict(k2=A(kps_un), d2=A(desc, np.uint8))
        n, assign = mc.oracle_projection(c, A(occupied, np.uint8), A(q_xyr, np.float32), A(q_levels, np.int32), A(q_desc, np.uint8), A(q_angle, np.float32),
                                         A(q_observed, np.uint8), mode, nnratio, check_ori, th_high)
        return n, assign, None
    monkeypatch.setattr(api, "search_by_projection", by_projection)
    m2.replay_product(_OracleSearches(0.6, True), golden)


@pytest.mark.skipif(oracle.ref_match() is None, reason="oracle/_ref/libref_match.so not built (needs /root/reference)")
def test_live_reference_on_full_size_frames():
    R, O = oracle.ref_match(), oracle.lib()
    assert mc.NFEATURES == 1000
    c = m2.triangulation_inputs(seed=3)
    for ori in (0, 1):
        a = m2.run_triangulation(R, "ref", c, ori); b = m2.run_triangulation(O, "oracle", c, ori)
        assert a[0] == b[0] and a[0] > 40 and np.array_equal(a[1], b[1])
    c = m2.keyframe_points_inputs(seed=4)
    for th in (3.0, 8.0):
        a = m2.run_fuse(R, "ref", c, th); b = m2.run_fuse(O, "oracle", c, th)
        assert a[0] == b[0] and a[0] > 200 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    c = m2.keyframe_points_inputs(seed=6, sim3=True)
    for th in (4.0, 12.0):
        a = m2.run_fuse_sim3(R, "ref", c, th); b = m2.run_fuse_sim3(O, "oracle", c, th)
        assert a[0] == b[0] and a[0] > 200 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        a = m2.run_loop(R, "ref", c, int(th)); b = m2.run_loop(O, "oracle", c, int(th))
        assert a[0] == b[0] and a[0] > 40 and np.array_equal(a[1], b[1])
    c = m2.sim3_inputs(seed=8)
    for th in (7.5, 2.0):
        a = m2.run_sim3(R, "ref", c, th); b = m2.run_sim3(O, "oracle", c, th)
        assert a[0] == b[0] and a[0] > 40 and np.array_equal(a[1], b[1])


@pytest.mark.skipif(oracle.ref_match() is None, reason="oracle/_ref/libref_match.so not built (needs /root/reference)")
def test_golden_file_is_current(golden):
    R = oracle.ref_match()
    c = golden["fuse"]; run = c["runs"][0]
    n, idx, act, qx, ql, qm = m2.run_fuse(R, "ref", c, float(run["cfg"][0]))
    assert n == run["n"] and np.array_equal(idx, run["fused_idx"]) and np.array_equal(qx, run["q_xyr"])
    c = golden["sim3"]; run = c["runs"][0]
    n, m12, qx, ql, qm = m2.run_sim3(R, "ref", c, float(run["cfg"][0]))
    assert n == run["n"] and np.array_equal(m12, run["matches12"]) and np.array_equal(qx, run["q_xyr"])


def test_predict_scale_thresholds_equal_libm():
    """level = #(ratio > threshold) is MapPoint::PredictScale evaluated with this host's logf, for every ratio: random ones and the neighbourhoods
    of the thresholds, where a non-monotone logf would show"""
    thr = kfgeom.level_thresholds(1.2, 8)
    log_sf = kfgeom.pyramid(1.2, 8)[3]
    rng = np.random.default_rng(0)
    r = np.exp(rng.uniform(-3, 4, 20000)).astype(np.float32)
    near = np.concatenate([(t.view(np.uint32) + np.arange(-300, 301).astype(np.uint32)).view(np.float32) for t in thr.reshape(-1, 1)])
    r = np.concatenate([r, near, np.array([0.0, np.inf, np.nan], np.float32)])
    want = m2.predict_scale(r, np.ones(len(r), np.float32), log_sf, 8)
    got = (r[:, None] > thr[None, :]).sum(1)
    ok = np.isfinite(r) & (r > 0)
    assert np.array_equal(got[ok], want[ok])
    assert got[-3] == 0 and got[-1] == 0                           # ratio 0 / NaN: level 0 (such points fail the distance tests anyway)


@pytest.mark.skipif(oracle.ref_match() is None, reason="oracle/_ref/libref_match.so not built (needs /root/reference)")
def test_live_reference_with_distorted_image_bounds(monkeypatch):
    """non-integer image bounds (a distorted camera, src/Frame.cc:390-416): the KeyFrame keeps them as int (include/KeyFrame.h:211-214), so IsInImage
    and the origin of KeyFrame::GetFeaturesInArea truncate while the grid itself was built from the float bounds"""
    monkeypatch.setattr(m2, "BOUNDS", np.array([-3.7, 643.6, -2.4, 482.9], np.float32))
    R, O = oracle.ref_match(), oracle.lib()
    c = m2.keyframe_points_inputs(seed=14)
    a = m2.run_fuse(R, "ref", c, 3.0); b = m2.run_fuse(O, "oracle", c, 3.0)
    assert a[0] == b[0] and a[0] > 200 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    c = m2.keyframe_points_inputs(seed=16, sim3=True)
    a = m2.run_fuse_sim3(R, "ref", c, 4.0); b = m2.run_fuse_sim3(O, "oracle", c, 4.0)
    assert a[0] == b[0] and a[0] > 200 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    a = m2.run_loop(R, "ref", c, 10); b = m2.run_loop(O, "oracle", c, 10)
    assert a[0] == b[0] and a[0] > 100 and np.array_equal(a[1], b[1])
    c = m2.sim3_inputs(seed=18)
    a = m2.run_sim3(R, "ref", c, 7.5); b = m2.run_sim3(O, "oracle", c, 7.5)
    assert a[0] == b[0] and a[0] > 100 and np.array_equal(a[1], b[1])
