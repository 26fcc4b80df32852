"""Seeded inputs for the KeyFrame-side matcher members (SearchForTriangulation, the two Fuse overloads, the loop-closing SearchByProjection,
SearchBySim3) and runners of one implementation over them: "ref" = the reference's own src/ORBmatcher.cc (oracle/_ref/libref_match.so, which also
returns the trace of the grid queries it made), "oracle" = oracle/match2_oracle.cpp.  Shared by tests/golden/make_match_golden.py (-> match_ref2.npz)
and by the tests that replay that file.  CPU only."""
import ctypes as C

import numpy as np

import match_cases as mc
import oracle

P, A, BOUNDS, CAM4 = mc.P, mc.A, mc.BOUNDS, mc.CAM4


def _nodes_of(d):
    lab = (d[:, 0] >> 6).astype(np.int32) * 4 + (d[:, 1] >> 6)
    fv = {}
    for i, l in enumerate(lab):
        fv.setdefault(int(l) + 100, []).append(i)
    return fv


def _minmax(rng, pos, T, octave):
    ow = -(T[:3, :3].astype(np.float64).T @ T[:3, 3].astype(np.float64))
    dist = np.linalg.norm(pos - ow, axis=1)
    maxd = dist * 1.2 ** (octave + rng.uniform(-0.4, 0.4, len(pos)))
    maxd[rng.random(len(pos)) < 0.05] *= 0.3                      # too far for the scale pyramid
    return A(np.stack([maxd / 1.2 ** 7, maxd], 1), np.float32), ow


def _normals(rng, pos, ow):
    nrm = pos - ow
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm += rng.normal(0, 0.3, nrm.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[rng.random(len(pos)) < 0.05] *= -1                         # seen from behind: fails the 60 degree test
    return A(nrm, np.float32)


def triangulation_inputs(seed=51):
    """two keyframes whose keypoints differ by an image shift; F12 = the fundamental matrix of that shift (epipole at infinity along the shift),
    slightly perturbed; the poses only place the epipole used by the distance test (inside the image, so some candidates fall to it)"""
    rng = np.random.default_rng(seed)
    shift = (3, 5)                                                 # rows, columns
    k1, d1, k2, d2 = mc.two_views(93, shift)
    extra = rng.integers(0, len(d1), 150)
    d2 = A(np.concatenate([d2, d1[extra] ^ np.packbits(rng.integers(0, 100, (150, 256)) < 2, axis=1)]))
    ke = k1[extra].copy(); ke["x"] += shift[1] + rng.normal(0, 1.0, 150).astype(np.float32); ke["y"] += shift[0] + rng.normal(0, 1.0, 150).astype(np.float32)
    k2 = A(np.concatenate([k2, ke]))
    fv1, fv2 = _nodes_of(d1), _nodes_of(d2)
    fv1.pop(sorted(fv1)[2]); fv2.pop(sorted(fv2)[-3])
    n1, s1, i1 = mc.fv_arrays(fv1); n2, s2, i2 = mc.fv_arrays(fv2)
    s = np.array([shift[1], shift[0], 0.0])
    sx = np.array([[0, -s[2], s[1]], [s[2], 0, -s[0]], [-s[1], s[0], 0]])
    F12 = A(-sx * 0.01 + rng.normal(0, 2e-6, (3, 3)), np.float32)
    T1 = mc._pose(rng, 1.0)
    T2 = np.eye(4, dtype=np.float32); T2[:3, 3] = [0.02, -0.03, 0.5]
    T2 = A(T2 @ T1)
    return dict(k1=k1, d1=d1, has1=(rng.random(len(d1)) < 0.3).astype(np.uint8), n1=n1, s1=s1, i1=i1, T1=A(T1),
                k2=k2, d2=d2, has2=(rng.random(len(d2)) < 0.3).astype(np.uint8), n2=n2, s2=s2, i2=i2, T2=T2, F12=F12)


def run_triangulation(L, prefix, c, ori):
    f = getattr(L, prefix + "_search_for_triangulation")
    m = np.zeros(max(len(c["d1"]), 1), np.int32)
    n = f(P(c["k1"]), P(c["d1"]), P(c["has1"]), len(c["d1"]), P(c["n1"]), P(c["s1"]), P(c["i1"]), len(c["n1"]), P(c["T1"]),
          P(c["k2"]), P(c["d2"]), P(c["has2"]), len(c["d2"]), P(c["n2"]), P(c["s2"]), P(c["i2"]), len(c["n2"]), P(c["T2"]), P(BOUNDS), P(CAM4), P(c["F12"]),
          int(ori), P(m))
    return n, m[:len(c["d1"])]


def keyframe_points_inputs(seed=53, sim3=False):
    """a keyframe (second view) and a list of map points that project near the keypoints of the first view; used by Fuse, Fuse(Scw) and the
    loop-closing SearchByProjection.  sim3: the pose is handed over as a Sim3 matrix s [R | t]"""
    rng = np.random.default_rng(seed)
    k, d, kq, dq = mc.two_views(94, (2, 3))
    n, m = len(k), len(kq)
    T = mc._pose(rng, 1.0)
    uv = np.stack([kq["x"] - 3 + rng.normal(0, 1.5, m), kq["y"] - 2 + rng.normal(0, 1.5, m)], 1).astype(np.float64)
    uv[rng.random(m) < 0.03] += 900
    depth = rng.uniform(2, 8, m)
    depth[rng.random(m) < 0.03] *= -1
    pos = mc._world_points(rng, T, uv, depth)
    minmax, ow = _minmax(rng, pos, T, kq["octave"])
    S = T.copy()
    if sim3:
        S[:3, :] *= np.float32(1.07)
    return dict(k=k, d=d, T=A(S), held_state=A(rng.choice([0, 0, 0, 1, 1, 2], n), np.uint8), held_nobs=A(rng.choice([1, 2, 3, 6], n), np.int32),
                mp_state=A(rng.choice([0, 1, 1, 1, 1, 1, 1, 1, 2, 3], m), np.uint8), mp_pos=pos, mp_normal=_normals(rng, pos.astype(np.float64), ow),
                mp_desc=dq, mp_minmax=minmax, mp_nobs=A(rng.choice([1, 2, 3, 6], m), np.int32),
                matched=A(rng.choice([-1, -1, -1, -1, -2, 0], n) * 1, np.int32))


def _trace(cap):
    return np.zeros((cap, 3), np.float32), np.zeros((cap, 2), np.int32), np.zeros(cap, np.int32), np.zeros(1, np.int32)


def run_fuse(L, prefix, c, th):
    m = len(c["mp_desc"])
    idx, act = np.zeros(m, np.int32), np.zeros(m, np.int32)
    args = [P(c["k"]), P(c["d"]), len(c["k"]), P(BOUNDS), P(CAM4), P(c["T"]), P(c["held_state"]), P(c["held_nobs"]), m, P(c["mp_state"]), P(c["mp_pos"]),
            P(c["mp_normal"]), P(c["mp_desc"]), P(c["mp_minmax"]), P(c["mp_nobs"]), C.c_float(th), P(idx), P(act)]
    if prefix == "ref":
        qx, ql, qm, nq = _trace(m)
        n = L.ref_fuse(*args, P(qx), P(ql), P(qm), P(nq))
        return n, idx, act, qx[:nq[0]], ql[:nq[0]], qm[:nq[0]]
    return L.oracle_fuse(*args), idx, act


def run_fuse_sim3(L, prefix, c, th):
    m = len(c["mp_desc"])
    rep, add = np.zeros(m, np.int32), np.zeros(m, np.int32)
    st = np.where(c["mp_state"] == 0, 1, c["mp_state"]).astype(np.uint8)        # vpPoints holds no NULLs here
    args = [P(c["k"]), P(c["d"]), len(c["k"]), P(BOUNDS), P(CAM4), P(c["T"]), P(c["held_state"]), m, P(st), P(c["mp_pos"]), P(c["mp_normal"]),
            P(c["mp_desc"]), P(c["mp_minmax"]), C.c_float(th), P(rep), P(add)]
    if prefix == "ref":
        qx, ql, qm, nq = _trace(m)
        n = L.ref_fuse_sim3(*args, P(qx), P(ql), P(qm), P(nq))
        return n, rep, add, qx[:nq[0]], ql[:nq[0]], qm[:nq[0]]
    return L.oracle_fuse_sim3(*args), rep, add


def loop_matched(c):
    """vpMatched before the call: -1 none, -2 some other point, the rest distinct indices into vpPoints"""
    rng = np.random.default_rng(7)
    matched = c["matched"].copy()
    z = np.nonzero(matched == 0)[0]
    matched[z] = rng.choice(len(c["mp_desc"]), len(z), replace=False)
    return matched


def run_loop(L, prefix, c, th):
    m = len(c["mp_desc"])
    matched = loop_matched(c)
    st = np.where(c["mp_state"] == 2, 2, 1).astype(np.uint8)
    args = [P(c["k"]), P(c["d"]), len(c["k"]), P(BOUNDS), P(CAM4), P(c["T"]), m, P(st), P(c["mp_pos"]), P(c["mp_normal"]), P(c["mp_desc"]), P(c["mp_minmax"]),
            int(th), P(matched)]
    if prefix == "ref":
        qx, ql, qm, nq = _trace(m)
        n = L.ref_search_by_projection_loop(*args, P(qx), P(ql), P(qm), P(nq))
        return n, matched, qx[:nq[0]], ql[:nq[0]], qm[:nq[0]]
    return L.oracle_search_by_projection_loop(*args), matched


def sim3_inputs(seed=57):
    """two keyframes related by a sideways translation (all points near one depth, so keypoints differ by the image shift), every feature owning
    a map point; Sim3 (s12, R12, t12) = the relative pose, slightly off"""
    rng = np.random.default_rng(seed)
    shift = (3, 5)
    k1, d1, k2, d2 = mc.two_views(95, shift)
    Z0 = 4.0
    T1 = mc._pose(rng, 1.0).astype(np.float64)
    Trel = np.eye(4); Trel[:3, 3] = [shift[1] * Z0 / CAM4[0], shift[0] * Z0 / CAM4[1], 0.0]
    T2 = Trel @ T1

    def points(k, T, seed2):
        r = np.random.default_rng(seed2)
        uv = np.stack([k["x"], k["y"]], 1).astype(np.float64) + r.normal(0, 0.7, (len(k), 2))
        depth = Z0 * (1 + r.normal(0, 0.01, len(k)))
        depth[r.random(len(k)) < 0.03] *= -1
        pos = mc._world_points(r, T.astype(np.float32), uv, depth)
        ow = -(T[:3, :3].T @ T[:3, 3])
        dist = np.linalg.norm(pos - ow, axis=1)
        maxd = dist * 1.2 ** (k["octave"] + r.uniform(-0.4, 0.4, len(k)))
        maxd[r.random(len(k)) < 0.05] *= 0.3
        return pos, A(np.stack([maxd / 1.2 ** 7, maxd], 1), np.float32), A(r.choice([0, 1, 1, 1, 1, 1, 1, 2], len(k)), np.uint8)
    p1, mm1, st1 = points(k1, T1, seed + 1)
    p2, mm2, st2 = points(k2, T2, seed + 2)
    R12 = T1[:3, :3] @ T2[:3, :3].T
    t12 = T1[:3, 3] - R12 @ T2[:3, 3]
    w = rng.normal(0, 0.002, 3)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R12 = (np.eye(3) + Kx) @ R12
    m12 = np.full(len(k1), -1, np.int32)
    pre = rng.choice(len(k1), 40, replace=False)
    m12[pre] = rng.choice(len(k2), 40, replace=False)             # matches found earlier (by SearchByBoW): both ends are skipped
    return dict(k1=k1, d1=d1, T1=A(T1, np.float32), st1=st1, p1=p1, mm1=mm1, k2=k2, d2=d2, T2=A(T2, np.float32), st2=st2, p2=p2, mm2=mm2,
                s12=np.float32(1.01), R12=A(R12, np.float32), t12=A(t12 + rng.normal(0, 0.002, 3), np.float32), m12=m12)


def run_sim3(L, prefix, c, th):
    m12 = c["m12"].copy()
    args = [P(c["k1"]), P(c["d1"]), len(c["k1"]), P(c["T1"]), P(c["st1"]), P(c["p1"]), P(c["d1"]), P(c["mm1"]),
            P(c["k2"]), P(c["d2"]), len(c["k2"]), P(c["T2"]), P(c["st2"]), P(c["p2"]), P(c["d2"]), P(c["mm2"]), P(BOUNDS), P(CAM4),
            C.c_float(float(c["s12"])), P(c["R12"]), P(c["t12"]), C.c_float(th), P(m12)]
    if prefix == "ref":
        cap = len(c["k1"]) + len(c["k2"])
        qx, ql, qm, nq = _trace(cap)
        n = L.ref_search_by_sim3(*args, P(qx), P(ql), P(qm), P(nq))
        return n, m12, qx[:nq[0]], ql[:nq[0]], qm[:nq[0]]
    return L.oracle_search_by_sim3(*args), m12


# ---- numpy model of the device projection kernel (k_kf_project): the reference's per-point statements in cv::Mat CV_32F arithmetic.  Used by the CPU
# ---- tests to pin that arithmetic to the queries the reference itself made, and as the projection of the oracle-backed test double ------------
from orb_slam2_aruco_b200 import kfgeom  # noqa: E402

f32 = np.float32


def predict_scale(max_distance, dist, log_sf, nlevels):
    """MapPoint::PredictScale (src/MapPoint.cc:403-435) element by element through libm's logf, the way the reference evaluates it"""
    with np.errstate(all="ignore"):
        ratio = (np.asarray(max_distance, f32) / np.asarray(dist, f32)).astype(f32)
    out = np.zeros(len(ratio), np.int32)
    for i, r in enumerate(ratio):
        if not np.isfinite(r) or r <= 0:
            continue
        n = int(np.ceil(f32(f32(kfgeom._libm.logf(float(r))) / log_sf)))
        out[i] = min(max(n, 0), nlevels - 1)
    return out


def _pixel(pc, cam4):
    cam4 = np.asarray(cam4, f32)
    with np.errstate(all="ignore"):
        invz = (f32(1.0) / pc[:, 2]).astype(f32)
        u = cam4[0] * (pc[:, 0] * invz) + cam4[2]
        v = cam4[1] * (pc[:, 1] * invz) + cam4[3]
    return u.astype(f32), v.astype(f32), invz


def _in_image(u, v, bounds4):
    b = [f32(int(x)) for x in np.asarray(bounds4, f32)]           # KeyFrame keeps mnMinX .. mnMaxY as int (include/KeyFrame.h:211-214)
    return (u >= b[0]) & (u < b[1]) & (v >= b[2]) & (v < b[3])


def host_project_points(pose, cam4, bounds4, pos, normal, minmax, th, scale_factor=1.2, nlevels=8):
    """the part of Fuse / Fuse(Scw) / SearchByProjection(Scw) between "Get 3D Coords" and GetFeaturesInArea for all points at once.
    pose = (R, t, Ow).  Returns (valid [N] bool, q_xyr [N, 3] float32, level [N] int32); rows with valid = False were discarded by one of the tests."""
    R, t, Ow = pose
    pos = np.ascontiguousarray(pos, f32).reshape(-1, 3); normal = np.ascontiguousarray(normal, f32).reshape(-1, 3)
    minmax = np.ascontiguousarray(minmax, f32).reshape(-1, 2)
    sf, _, _, log_sf = kfgeom.pyramid(scale_factor, nlevels)
    pc = (kfgeom._mul_points(R, pos) + t).astype(f32)
    valid = ~(pc[:, 2] < 0)
    u, v, _ = _pixel(pc, cam4)
    valid &= _in_image(u, v, bounds4)
    maxd, mind = f32(1.2) * minmax[:, 1], f32(0.8) * minmax[:, 0]
    PO = (pos - Ow).astype(f32)
    dist = np.sqrt(kfgeom._dot_rows(PO, PO)).astype(f32)
    valid &= ~(dist < mind) & ~(dist > maxd)
    valid &= ~(kfgeom._dot_rows(PO, normal) < 0.5 * dist.astype(np.float64))
    level = np.zeros(len(pos), np.int32)
    idx = np.nonzero(valid)[0]
    level[idx] = predict_scale(minmax[idx, 1], dist[idx], log_sf, nlevels)
    radius = (f32(th) * sf[level]).astype(f32)
    return valid, np.ascontiguousarray(np.stack([u, v, radius], 1), f32), level


def host_project_points_sim3(pose_a, sR, tt, cam4, bounds4, pos, minmax, th, scale_factor=1.2, nlevels=8):
    """one direction of SearchBySim3 (src/ORBmatcher.cc:1158-1195): world point -> camera a -> camera b = sR * p + tt -> pixel in keyframe b"""
    R, t, _ = pose_a
    pos = np.ascontiguousarray(pos, f32).reshape(-1, 3); minmax = np.ascontiguousarray(minmax, f32).reshape(-1, 2)
    sf, _, _, log_sf = kfgeom.pyramid(scale_factor, nlevels)
    pa = (kfgeom._mul_points(R, pos) + t).astype(f32)
    pb = (kfgeom._mul_points(sR, pa) + tt).astype(f32)
    valid = ~(pb[:, 2] < 0)
    u, v, _ = _pixel(pb, cam4)
    valid &= _in_image(u, v, bounds4)
    maxd, mind = f32(1.2) * minmax[:, 1], f32(0.8) * minmax[:, 0]
    dist = np.sqrt(kfgeom._dot_rows(pb, pb)).astype(f32)
    valid &= ~(dist < mind) & ~(dist > maxd)
    level = np.zeros(len(pos), np.int32)
    idx = np.nonzero(valid)[0]
    level[idx] = predict_scale(minmax[idx, 1], dist[idx], log_sf, nlevels)
    radius = (f32(th) * sf[level]).astype(f32)
    return valid, np.ascontiguousarray(np.stack([u, v, radius], 1), f32), level


# ---- replay of tests/golden/match_ref2.npz through the product's ORBmatcher mirror -----------------------------------------------------------
def fv_of(c, side):
    return mc.fv_dict(c["n" + side], c["s" + side], c["i" + side])


def replay_product(M, golden):
    """M = orb_slam2_aruco_b200.api.ORBmatcher (device searches) or the same object with its two device calls swapped for the oracle's (CPU test
    of the host glue).  Every answer is compared with what the reference's own ORBmatcher.cc returned for the same inputs."""
    c = golden["tri"]
    for run in c["runs"]:
        M.mbCheckOrientation = bool(run["cfg"][0])
        n, m12 = M.SearchForTriangulation(c["k1"], c["d1"], c["has1"], fv_of(c, "1"), c["T1"], c["k2"], c["d2"], c["has2"], fv_of(c, "2"), c["T2"], CAM4, c["F12"])
        assert n == run["n"] and np.array_equal(m12, run["matches12"]) and n > 30
    c = golden["fuse"]
    for run in c["runs"]:
        n, idx, act = M.Fuse(c["k"], c["d"], BOUNDS, CAM4, c["T"], c["held_state"], c["held_nobs"], c["mp_state"], c["mp_pos"], c["mp_normal"], c["mp_desc"],
                             c["mp_minmax"], c["mp_nobs"], float(run["cfg"][0]))
        assert n == run["n"] and np.array_equal(idx, run["fused_idx"]) and np.array_equal(act, run["action"])
        assert all((run["action"] == a).sum() > 5 for a in (1, 2, 3))
    c = golden["scw"]
    for run in c["runs"]:
        st = np.where(c["mp_state"] == 0, 1, c["mp_state"]).astype(np.uint8)
        n, rep, add = M.FuseSim3(c["k"], c["d"], BOUNDS, CAM4, c["T"], c["held_state"], st, c["mp_pos"], c["mp_normal"], c["mp_desc"], c["mp_minmax"],
                                 float(run["cfg"][0]))
        assert n == run["n"] and np.array_equal(rep, run["replace_idx"]) and np.array_equal(add, run["added_idx"]) and n > 100
        st = np.where(c["mp_state"] == 2, 2, 1).astype(np.uint8)
        n, matched = M.SearchByProjectionLoop(c["k"], c["d"], BOUNDS, CAM4, c["T"], st, c["mp_pos"], c["mp_normal"], c["mp_desc"], c["mp_minmax"],
                                              loop_matched(c), int(run["cfg"][1]))
        assert n == run["n_loop"] and np.array_equal(matched, run["matched"]) and n > 50
    c = golden["sim3"]
    for run in c["runs"]:
        n, m12 = M.SearchBySim3(c["k1"], c["d1"], c["T1"], c["st1"], c["p1"], c["d1"], c["mm1"], c["k2"], c["d2"], c["T2"], c["st2"], c["p2"], c["d2"], c["mm2"],
                                BOUNDS, CAM4, c["m12"], float(c["s12"]), c["R12"], c["t12"], float(run["cfg"][0]))
        assert n == run["n"] and np.array_equal(m12, run["matches12"]) and n > 50
