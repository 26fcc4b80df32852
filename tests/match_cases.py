"""Seeded inputs for the matcher parity cases and thin runners of one implementation ("ref" = the reference's own src/ORBmatcher.cc in
oracle/_ref/libref_match.so, "oracle" = oracle/match_oracle.cpp) over them.  Shared by tests/golden/make_match_golden.py (which stores the
inputs and the reference's answers in tests/golden/match_ref.npz) and by the tests that replay that file.  CPU only."""
import ctypes as C

import numpy as np

import oracle
from orb_slam2_aruco_b200 import synth

vp = C.c_void_p
KP = oracle.KP_DTYPE
BOUNDS = np.array([0, 640, 0, 480], np.float32)
CAM4 = np.array([517.3, 516.5, 318.6, 255.3], np.float32)


def P(a):
    return a.ctypes.data_as(vp)


def A(a, dtype=None):
    return np.ascontiguousarray(a, dtype)


def fv_arrays(fv):
    """{node: [indices]} -> sorted arrays (nodes, start, items)"""
    nodes = np.array(sorted(fv), np.int32)
    start = np.zeros(len(nodes) + 1, np.int32)
    items = []
    for j, nd in enumerate(nodes):
        items += list(fv[int(nd)])
        start[j + 1] = len(items)
    return nodes, start, np.array(items if items else [0], np.int32)


def fv_dict(nodes, start, items):
    return {int(nd): items[start[j]:start[j + 1]].tolist() for j, nd in enumerate(nodes)}


# ---- inputs ---------------------------------------------------------------------------------------------------------------------------------
NFEATURES = 1000                     # the golden file is written with 500 to stay small


def two_views(seed, shift):
    a = synth.make_frame(seed)
    k1, d1 = oracle.orb_extract(a, NFEATURES)
    k2, d2 = oracle.orb_extract(np.roll(a, shift, axis=(0, 1)), NFEATURES)
    return A(k1), A(d1), A(k2), A(d2)


def bow_inputs(seed=31):
    """two views, near-duplicate descriptors on the second side, feature vectors from a random node labelling that keeps similar
    descriptors together (nodes = the top bits of the first descriptor bytes), random good-MapPoint masks"""
    rng = np.random.default_rng(seed)
    k1, d1, k2, d2 = two_views(96, (2, 3))
    extra = d1[rng.integers(0, len(d1), 200)] ^ np.packbits(rng.integers(0, 100, (200, 256)) < 2, axis=1)
    d2 = A(np.concatenate([d2, extra]))
    a1 = A(k1["angle"]); a2 = A(np.concatenate([k2["angle"], rng.uniform(0, 360, 200).astype(np.float32)]))

    def nodes_of(d):
        lab = (d[:, 0] >> 6).astype(np.int32) * 4 + (d[:, 1] >> 6)
        fv = {}
        for i, l in enumerate(lab):
            fv.setdefault(int(l) + 100, []).append(i)
        return fv
    fv1, fv2 = nodes_of(d1), nodes_of(d2)
    fv1.pop(sorted(fv1)[1]); fv2.pop(sorted(fv2)[-2])            # nodes present on one side only exercise the lower_bound branches
    v1 = (rng.random(len(d1)) < 0.8).astype(np.uint8); v2 = (rng.random(len(d2)) < 0.8).astype(np.uint8)
    n1, s1, i1 = fv_arrays(fv1); n2, s2, i2 = fv_arrays(fv2)
    return dict(d1=d1, a1=a1, v1=v1, n1=n1, s1=s1, i1=i1, d2=d2, a2=a2, v2=v2, n2=n2, s2=s2, i2=i2)


def one_node(c):
    """the same descriptors with ONE all-inclusive vocabulary node and good MapPoints everywhere: the brute-force configuration of the bench"""
    n1, n2 = len(c["d1"]), len(c["d2"])
    one = np.array([7], np.int32)
    return dict(c, v1=np.ones(n1, np.uint8), v2=np.ones(n2, np.uint8), n1=one, s1=np.array([0, n1], np.int32), i1=np.arange(n1, dtype=np.int32),
                n2=one, s2=np.array([0, n2], np.int32), i2=np.arange(n2, dtype=np.int32))


def init_inputs(shift=(4, 7)):
    k1, d1, k2, d2 = two_views(70, shift)
    return dict(k1=k1, d1=d1, k2=k2, d2=d2, prev=A(np.stack([k1["x"], k1["y"]], 1), np.float32))


def _pose(rng, scale):
    w = rng.normal(0, 0.02 * scale, 3)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = R; T[:3, 3] = rng.normal(0, 0.05 * scale, 3)
    return T


def _world_points(rng, T, uv, depth):
    """world positions that project to uv (pixels) at the given depth under pose T (float64 algebra, stored as float32)"""
    xc = np.stack([(uv[:, 0] - CAM4[2]) / CAM4[0] * depth, (uv[:, 1] - CAM4[3]) / CAM4[1] * depth, depth], 1)
    return A(((xc - T[:3, 3].astype(np.float64)) @ T[:3, :3].astype(np.float64)), np.float32)      # R^T (xc - t)


def points_inputs(seed=41):
    """SearchByProjection(Frame, vpMapPoints, th): map points = the keypoints of a second view, projections perturbed so windows overlap"""
    rng = np.random.default_rng(seed)
    k2, d2, kq, dq = two_views(90, (2, 3))
    n, m = len(k2), len(kq)
    return dict(k2=k2, d2=d2, frame_obs=A(rng.choice([-1, -1, -1, -1, -1, -1, -1, 0, 1, 3], n), np.int32),
                mp_in_view=(rng.random(m) < 0.9).astype(np.uint8), mp_bad=(rng.random(m) < 0.05).astype(np.uint8),
                mp_level=A(kq["octave"], np.int32), mp_viewcos=A(rng.choice([0.9, 0.9985, 0.999], m), np.float32),
                mp_projxy=A(np.stack([kq["x"] - 3 + rng.normal(0, 2, m), kq["y"] - 2 + rng.normal(0, 2, m)], 1), np.float32),
                mp_desc=dq, mp_nobs=A(rng.choice([0, 1, 2, 5], m), np.int32))


def last_inputs(seed=43):
    """SearchByProjection(Current, Last, th, mono): the last frame's map points lie where the current pose sees them near the matching
    keypoints; some are behind the camera, outside the image, outliers or absent"""
    rng = np.random.default_rng(seed)
    k2, d2, kl, dl = two_views(91, (3, 2))
    n, m = len(k2), len(kl)
    Tc, Tl = _pose(rng, 1.0), _pose(rng, 1.0)
    uv = np.stack([kl["x"] - 2 + rng.normal(0, 2, m), kl["y"] - 3 + rng.normal(0, 2, m)], 1).astype(np.float64)
    uv[rng.random(m) < 0.03] += 900                               # outside the image bounds
    depth = rng.uniform(2, 8, m)
    depth[rng.random(m) < 0.03] *= -1                             # behind the camera
    return dict(k2=k2, d2=d2, frame_obs=A(rng.choice([-1, -1, -1, -1, -1, -1, -1, 0, 1, 3], n), np.int32), tcw_cur=A(Tc), tcw_last=A(Tl), k_last=kl,
                mp_present=(rng.random(m) < 0.85).astype(np.uint8), mp_outlier=(rng.random(m) < 0.05).astype(np.uint8),
                mp_pos=_world_points(rng, Tc, uv, depth), mp_desc=dl, mp_nobs=A(rng.choice([0, 1, 2, 5], m), np.int32))


def reloc_inputs(seed=47):
    """SearchByProjection(Current, KeyFrame, sAlreadyFound, th, ORBdist)"""
    rng = np.random.default_rng(seed)
    k2, d2, kk, dk = two_views(92, (2, 4))
    n, m = len(k2), len(kk)
    Tc = _pose(rng, 1.0)
    uv = np.stack([kk["x"] - 4 + rng.normal(0, 2, m), kk["y"] - 2 + rng.normal(0, 2, m)], 1).astype(np.float64)
    uv[rng.random(m) < 0.03] -= 900
    depth = rng.uniform(2, 8, m)
    pos = _world_points(rng, Tc, uv, depth)
    ow = -(Tc[:3, :3].astype(np.float64).T @ Tc[:3, 3].astype(np.float64))
    dist = np.linalg.norm(pos - ow, axis=1)
    maxd = dist * 1.2 ** (kk["octave"] + rng.uniform(-0.4, 0.4, m))
    maxd[rng.random(m) < 0.05] *= 0.3                             # too far for the scale pyramid
    mind = maxd / 1.2 ** 7
    return dict(k2=k2, d2=d2, frame_obs=A(rng.choice([-1, -1, -1, -1, -1, -1, -1, 0, 1, 3], n), np.int32), tcw_cur=A(Tc), k_kf=kk,
                mp_state=A(rng.choice([0, 1, 1, 1, 1, 1, 1, 1, 2, 3], m), np.uint8), mp_pos=pos, mp_desc=dk,
                mp_minmax=A(np.stack([mind, maxd], 1), np.float32))


# ---- runners: impl = ctypes library + prefix ("ref" / "oracle") --------------------------------------------------------------------------------
def run_bow(L, prefix, c, ratio, ori):
    m = np.zeros(max(len(c["d2"]), 1), np.int32)
    if prefix == "ref":
        n = L.ref_search_by_bow_nodes(P(c["d1"]), P(c["a1"]), P(c["v1"]), len(c["d1"]), P(c["n1"]), P(c["s1"]), P(c["i1"]), len(c["n1"]),
                                      P(c["d2"]), P(c["a2"]), len(c["d2"]), P(c["n2"]), P(c["s2"]), P(c["i2"]), len(c["n2"]), C.c_float(ratio), int(ori), P(m))
    else:
        n = L.oracle_search_by_bow_nodes(P(c["d1"]), P(c["a1"]), P(c["v1"]), P(c["n1"]), P(c["s1"]), P(c["i1"]), len(c["n1"]),
                                         P(c["d2"]), P(c["a2"]), len(c["d2"]), P(c["n2"]), P(c["s2"]), P(c["i2"]), len(c["n2"]), C.c_float(ratio), int(ori), P(m))
    return n, m[:len(c["d2"])]


def run_bow_kfkf(L, prefix, c, ratio, ori):
    m = np.zeros(max(len(c["d1"]), 1), np.int32)
    f = L.ref_search_by_bow_kfkf_nodes if prefix == "ref" else L.oracle_search_by_bow_kfkf_nodes
    n = f(P(c["d1"]), P(c["a1"]), P(c["v1"]), len(c["d1"]), P(c["n1"]), P(c["s1"]), P(c["i1"]), len(c["n1"]),
          P(c["d2"]), P(c["a2"]), P(c["v2"]), len(c["d2"]), P(c["n2"]), P(c["s2"]), P(c["i2"]), len(c["n2"]), C.c_float(ratio), int(ori), P(m))
    return n, m[:len(c["d1"])]


def run_init(L, prefix, c, prev, window, ratio, ori):
    f = L.ref_search_for_initialization if prefix == "ref" else L.oracle_search_for_initialization
    prev = prev.copy(); m = np.zeros(max(len(c["k1"]), 1), np.int32)
    n = f(P(c["k1"]), P(c["d1"]), len(c["k1"]), P(c["k2"]), P(c["d2"]), len(c["k2"]), P(BOUNDS), P(prev), window, C.c_float(ratio), int(ori), P(m))
    return n, m[:len(c["k1"])], prev


def _trace_bufs(cap):
    return np.zeros((cap, 3), np.float32), np.zeros((cap, 2), np.int32), np.zeros(cap, np.int32), np.zeros(1, np.int32)


def ref_points(L, c, th, ratio):
    cap = len(c["mp_desc"]); assign = np.zeros(len(c["k2"]), np.int32); qx, ql, qm, nq = _trace_bufs(cap)
    n = L.ref_search_by_projection_points(P(c["k2"]), P(c["d2"]), len(c["k2"]), P(BOUNDS), P(c["frame_obs"]), cap, P(c["mp_in_view"]), P(c["mp_bad"]),
                                          P(c["mp_level"]), P(c["mp_viewcos"]), P(c["mp_projxy"]), P(c["mp_desc"]), P(c["mp_nobs"]), C.c_float(th),
                                          C.c_float(ratio), P(assign), P(qx), P(ql), P(qm), P(nq))
    return n, assign, qx[:nq[0]], ql[:nq[0]], qm[:nq[0]]


def ref_last(L, c, th, ori):
    cap = len(c["k_last"]); assign = np.zeros(len(c["k2"]), np.int32); qx, ql, qm, nq = _trace_bufs(cap)
    n = L.ref_search_by_projection_last(P(c["k2"]), P(c["d2"]), len(c["k2"]), P(BOUNDS), P(c["frame_obs"]), P(CAM4), P(c["tcw_cur"]), P(c["tcw_last"]),
                                        P(c["k_last"]), cap, P(c["mp_present"]), P(c["mp_outlier"]), P(c["mp_pos"]), P(c["mp_desc"]), P(c["mp_nobs"]),
                                        C.c_float(th), int(ori), P(assign), P(qx), P(ql), P(qm), P(nq))
    return n, assign, qx[:nq[0]], ql[:nq[0]], qm[:nq[0]]


def ref_reloc(L, c, th, orb_dist, ori):
    cap = len(c["k_kf"]); assign = np.zeros(len(c["k2"]), np.int32); qx, ql, qm, nq = _trace_bufs(cap)
    n = L.ref_search_by_projection_reloc(P(c["k2"]), P(c["d2"]), len(c["k2"]), P(BOUNDS), P(c["frame_obs"]), P(CAM4), P(c["tcw_cur"]), P(c["k_kf"]), cap,
                                         P(c["mp_state"]), P(c["mp_pos"]), P(c["mp_desc"]), P(c["mp_minmax"]), C.c_float(th), int(orb_dist), int(ori),
                                         P(assign), P(qx), P(ql), P(qm), P(nq))
    return n, assign, qx[:nq[0]], ql[:nq[0]], qm[:nq[0]]


def projection_queries(kind, c, q_mp):
    """what the product / oracle entry point takes for the traced queries of one reference call: (occupied, q_desc, q_angle, q_observed, mode)"""
    q_desc = A(c["mp_desc"][q_mp])
    if kind == "points":
        return (c["frame_obs"] > 0).astype(np.uint8), q_desc, np.zeros(len(q_mp), np.float32), (c["mp_nobs"][q_mp] > 0).astype(np.uint8), 0
    if kind == "last":
        return (c["frame_obs"] > 0).astype(np.uint8), q_desc, A(c["k_last"]["angle"][q_mp]), (c["mp_nobs"][q_mp] > 0).astype(np.uint8), 1
    # relocalisation: any map point pointer hides the keypoint (ORBmatcher.cc:1540-1541), and so does every new assignment
    return (c["frame_obs"] >= 0).astype(np.uint8), q_desc, A(c["k_kf"]["angle"][q_mp]), np.ones(len(q_mp), np.uint8), 1


def oracle_projection(c, occupied, qx, ql, q_desc, q_angle, q_obs, mode, ratio, ori, th_high):
    occ = occupied.copy(); assign = np.zeros(len(c["k2"]), np.int32)
    n = oracle.lib().oracle_search_by_projection(P(c["k2"]), P(c["d2"]), len(c["k2"]), P(BOUNDS), P(occ), P(A(qx)), P(A(ql)), P(q_desc), P(q_angle), P(q_obs),
                                                 len(qx), mode, C.c_float(ratio), int(ori), th_high, P(assign))
    return n, assign


def expected_assign(ref_assign, q_mp):
    """the reference's answer (map point index per frame keypoint, -2 = the preset point, -1 = none) in the product's terms: query index"""
    pos = {int(m): q for q, m in enumerate(q_mp)}
    return np.array([pos[int(a)] if a >= 0 else -1 for a in ref_assign], np.int32)


# ---- the golden file ------------------------------------------------------------------------------------------------------------------------
def load_golden(path):
    """tests/golden/match_ref.npz -> {case: {name: array, "runs": [{cfg, n, ...}]}}; structured keypoint arrays get their dtype back"""
    g = np.load(path)
    cases = {}
    for key in g.files:
        parts = key.split(".")
        case = cases.setdefault(parts[0], {"runs": {}})
        if len(parts) == 2:
            case[parts[1]] = g[key]
        else:
            case["runs"].setdefault(".".join(parts[1:-1]), {})[parts[-1]] = g[key]
    for case in cases.values():
        for name in ("k1", "k2", "k_last", "k_kf", "k"):
            if name in case:
                case[name] = A(case[name]).view(KP).reshape(-1) if case[name].dtype != KP else A(case[name])
        case["runs"] = [case["runs"][k] for k in sorted(case["runs"])]
    return cases
