"""CPU: orb_slam2_aruco_b200/mapfile.py against the reference's OWN map file code - src/Map.cc Map::Save / Map::Load (:219-533), compiled unmodified on
oracle/mapshim into oracle/_ref/libref_map.so (oracle/ref_map_wrap.cpp).
 * a map written by Map::Save is parsed by MapFile.load: every field equals what was handed to the reference, MapFile.save reproduces the file byte for byte;
 * a file written by MapFile.save is read by Map::Load: the reference reconstructs the same points, keyframes, poses (through its quaternion conversion),
   features, observations, spanning tree and covisibility weights, and calls its recomputation hooks (UndistortKeyPoints, AssignFeaturesToGrid, ComputeBoW,
   ComputeDistinctiveDescriptors) the number of times mapfile.rebuild() batches them;
 * tests/golden/map_ref.bin (written by Map::Save here, committed) replays the first check where the reference library is absent (the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest

from orb_slam2_aruco_b200 import mapfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_map.so")
GOLDEN = os.path.join(HERE, "golden", "map_ref.bin")
P = lambda a: a.ctypes.data_as(C.c_void_p)


def scene(seed=11, n_mp=60, n_kf=4):
    """a small map: points, keyframes with rotations that exercise all four branches of the rotation -> quaternion conversion, a spanning tree and
    covisibility lists; a keyframe may list the same point at two features (AddObservation keeps the first)"""
    rng = np.random.default_rng(seed)
    mp_id = (np.arange(n_mp) * 3 + 7).astype(np.uint64)
    mp_pos = rng.normal(0, 2, (n_mp, 3)).astype(np.float32)
    kf_id = (np.arange(n_kf) * 5 + 2).astype(np.uint64)
    kf_time = (rng.random(n_kf) * 100).astype(np.float64)
    T = np.zeros((n_kf, 4, 4), np.float32)
    for k in range(n_kf):
        ang = [0.3, 3.0, 3.1, 3.05][k % 4]                       # trace > 0, and the three "largest diagonal" branches
        axis = [[0.2, 0.5, 0.8], [1, 0.05, 0.02], [0.03, 1, 0.04], [0.02, 0.03, 1]][k % 4]
        a = np.asarray(axis, np.float64); a /= np.linalg.norm(a)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        T[k, :3, :3] = R.astype(np.float32); T[k, :3, 3] = rng.normal(0, 1, 3).astype(np.float32); T[k, 3, 3] = 1
    n_kp = rng.integers(20, 40, n_kf).astype(np.int32)
    tot = int(n_kp.sum())
    kp5 = rng.random((tot, 5)).astype(np.float32) * 100
    octave = rng.integers(0, 8, tot).astype(np.int32)
    desc = rng.integers(0, 256, (tot, 32), dtype=np.uint8)
    kp_mp = np.where(rng.random(tot) < 0.6, rng.integers(0, n_mp, tot), -1).astype(np.int64)
    parent = np.array([-1] + [int(rng.integers(0, k)) for k in range(1, n_kf)], np.int32)
    con_ofs = [0]; con_kf = []; con_w = []
    for k in range(n_kf):
        others = [j for j in range(n_kf) if j != k and rng.random() < 0.7]
        con_kf += others; con_w += [int(rng.integers(15, 200)) for _ in others]; con_ofs.append(len(con_kf))
    return dict(mp_id=mp_id, mp_pos=mp_pos, kf_id=kf_id, kf_time=kf_time, T=T, n_kp=n_kp, kp5=kp5, octave=octave, desc=desc, kp_mp=kp_mp, parent=parent,
                con_ofs=np.asarray(con_ofs, np.int32), con_kf=np.asarray(con_kf + [0], np.int32), con_w=np.asarray(con_w + [0], np.int32))


def check_parsed(mf, s):
    """the parsed file against the scene handed to Map::Save (std::set order == creation order: the wrapper's arena)"""
    assert np.array_equal(mf.map_points["id"], s["mp_id"]) and np.array_equal(mf.map_points["pos"], s["mp_pos"])
    assert len(mf.keyframes) == len(s["kf_id"])
    o = 0
    for k, kf in enumerate(mf.keyframes):
        assert kf["id"] == int(s["kf_id"][k]) and kf["timestamp"] == float(s["kf_time"][k])
        assert np.array_equal(kf["quat"], mapfile.rotation_to_quaternion(s["T"][k, :3, :3])) and np.array_equal(kf["t"], s["T"][k, :3, 3])
        f = kf["features"]; n = int(s["n_kp"][k])
        assert len(f) == n
        for j, name in enumerate(("x", "y", "size", "angle", "response")):
            assert np.array_equal(f[name], s["kp5"][o:o + n, j])
        assert np.array_equal(f["octave"], s["octave"][o:o + n]) and np.array_equal(f["desc"], s["desc"][o:o + n]) and (f["cols"] == 32).all()
        want = np.where(s["kp_mp"][o:o + n] >= 0, s["kp_mp"][o:o + n], -1).astype(np.int64)
        assert np.array_equal(np.where(f["mp_idx"] == mapfile.ULONG_MAX, -1, f["mp_idx"].astype(np.int64)), want)
        assert int(mf.parents[k]) == (int(s["kf_id"][s["parent"][k]]) if s["parent"][k] >= 0 else int(mapfile.ULONG_MAX))
        c = mf.connections[k]
        # GetConnectedKeyFrames() is a std::set<KeyFrame*>: pointer (= creation) order
        idx = sorted(s["con_kf"][s["con_ofs"][k]:s["con_ofs"][k + 1]].tolist())
        w = {int(a): int(b) for a, b in zip(s["con_kf"][s["con_ofs"][k]:s["con_ofs"][k + 1]], s["con_w"][s["con_ofs"][k]:s["con_ofs"][k + 1]])}
        assert c["id"].tolist() == [int(s["kf_id"][j]) for j in idx] and c["weight"].tolist() == [w[j] for j in idx]
        o += n


def ref_save(path, s):
    L = C.CDLL(LIB)
    rc = L.ref_map_save(path.encode(), len(s["mp_id"]), P(s["mp_id"]), P(s["mp_pos"]), len(s["kf_id"]), P(s["kf_id"]), P(s["kf_time"]), P(s["T"]), P(s["n_kp"]),
                        P(s["kp5"]), P(s["octave"]), P(s["desc"]), P(s["kp_mp"]), P(s["parent"]), P(s["con_ofs"]), P(s["con_kf"]), P(s["con_w"]))
    assert rc == 0


def test_golden_file_written_by_map_save_parses(tmp_path):
    """runs everywhere: the committed bytes are Map::Save's"""
    s = scene()
    mf = mapfile.MapFile.load(GOLDEN)
    check_parsed(mf, s)
    out = os.path.join(str(tmp_path), "again.bin")
    mf.save(out)
    assert open(out, "rb").read() == open(GOLDEN, "rb").read()


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_map.so not built (needs /root/reference)")
def test_map_save_of_the_reference_is_parsed_and_reproduced(tmp_path):
    for seed in (11, 12, 13):
        s = scene(seed, n_mp=40 + seed, n_kf=3 + seed % 3)
        path = os.path.join(str(tmp_path), "m%d.bin" % seed)
        ref_save(path, s)
        mf = mapfile.MapFile.load(path)
        check_parsed(mf, s)
        out = path + ".again"
        mf.save(out)
        assert open(out, "rb").read() == open(path, "rb").read()
    ref_save(os.path.join(str(tmp_path), "g.bin"), scene())
    assert open(os.path.join(str(tmp_path), "g.bin"), "rb").read() == open(GOLDEN, "rb").read(), "tests/golden/map_ref.bin is stale"


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_map.so not built (needs /root/reference)")
def test_map_load_of_the_reference_reads_what_mapfile_writes(tmp_path):
    s = scene(21, n_mp=50, n_kf=5)
    src = os.path.join(str(tmp_path), "src.bin")
    ref_save(src, s)
    mf = mapfile.MapFile.load(src)
    mine = os.path.join(str(tmp_path), "mine.bin")
    mapfile.MapFile(mf.map_points, mf.keyframes, mf.parents, mf.connections).save(mine)
    L = C.CDLL(LIB)
    n_mp, n_kf = len(s["mp_id"]), len(s["kf_id"]); tot = int(s["n_kp"].sum())
    o_nmp = np.zeros(1, np.int32); o_nkf = np.zeros(1, np.int32)
    mp_id = np.zeros(n_mp, np.uint64); mp_pos = np.zeros((n_mp, 3), np.float32); obs = np.zeros(n_mp, np.int32)
    kf_id = np.zeros(n_kf, np.uint64); kf_time = np.zeros(n_kf, np.float64); kf_T = np.zeros((n_kf, 4, 4), np.float32); n_kp = np.zeros(n_kf, np.int32)
    kp5 = np.zeros((tot, 5), np.float32); octave = np.zeros(tot, np.int32); desc = np.zeros((tot, 32), np.uint8); kp_mp = np.zeros(tot, np.int64)
    parent = np.zeros(n_kf, np.int64); con_ofs = np.zeros(n_kf + 1, np.int32); con_id = np.zeros(64, np.uint64); con_w = np.zeros(64, np.int32); hooks = np.zeros(4, np.int32)
    rc = L.ref_map_load(mine.encode(), n_mp, P(o_nmp), P(mp_id), P(mp_pos), P(obs), n_kf, P(o_nkf), P(kf_id), P(kf_time), P(kf_T), P(n_kp), tot, P(kp5), P(octave),
                        P(desc), P(kp_mp), P(parent), P(con_ofs), 64, P(con_id), P(con_w), P(hooks))
    assert rc == 0 and o_nmp[0] == n_mp and o_nkf[0] == n_kf
    order = np.argsort(mp_id)
    assert np.array_equal(mp_id[order], np.sort(s["mp_id"])) and np.array_equal(mp_pos[order], s["mp_pos"][np.argsort(s["mp_id"])])
    # observations as mapfile counts them (one per keyframe and point, the first feature) == what the reference's AddObservation kept
    want_obs = {int(i): 0 for i in s["mp_id"]}
    for plist, pid in zip(mf.observations(), mf.map_points["id"]):
        want_obs[int(pid)] = len(plist)
    assert {int(i): int(n) for i, n in zip(mp_id, obs)} == want_obs
    starts = np.concatenate([[0], np.cumsum(n_kp)])
    for k_ref in range(n_kf):
        k = int(np.nonzero(s["kf_id"] == kf_id[k_ref])[0][0])
        a, b = int(starts[k_ref]), int(starts[k_ref + 1]); a0 = int(s["n_kp"][:k].sum()); n = int(s["n_kp"][k])
        assert b - a == n and kf_time[k_ref] == s["kf_time"][k]
        assert np.array_equal(kp5[a:b], s["kp5"][a0:a0 + n]) and np.array_equal(octave[a:b], s["octave"][a0:a0 + n]) and np.array_equal(desc[a:b], s["desc"][a0:a0 + n])
        want = np.where(s["kp_mp"][a0:a0 + n] >= 0, s["mp_id"][np.maximum(s["kp_mp"][a0:a0 + n], 0)].astype(np.int64), -1)
        assert np.array_equal(kp_mp[a:b], want)
        # the pose goes through quaternion and back (src/Map.cc:285-292, 455-468): mapfile.pose() is the same float matrix
        assert np.array_equal(kf_T[k_ref], mf.pose(k))
        assert parent[k_ref] == (int(s["kf_id"][s["parent"][k]]) if s["parent"][k] >= 0 else -1)
        got = {int(i): int(w) for i, w in zip(con_id[con_ofs[k_ref]:con_ofs[k_ref + 1]], con_w[con_ofs[k_ref]:con_ofs[k_ref + 1]])}
        lo, hi = s["con_ofs"][k], s["con_ofs"][k + 1]
        assert got == {int(s["kf_id"][j]): int(w) for j, w in zip(s["con_kf"][lo:hi], s["con_w"][lo:hi])}
    # the recomputation Load triggers: once per keyframe (undistort, grid, BoW), once per map point (distinctive descriptor) - what mapfile.rebuild() batches
    assert hooks.tolist()[2:] == [n_kf, n_mp]
