"""The reference's map file (src/Map.cc:219-330 Save, 339-533 Load): byte layout against a hand-packed file, round trips,
the quaternion converters, and - on the GPU - the batched recomputation half of Map::Load against the oracles."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import mapfile as mf

ULONG_MAX = 0xFFFFFFFFFFFFFFFF


def _hand_packed(tmp_path):
    """the byte stream Map::Save emits for 2 map points and 2 keyframes, written field by field with struct like the f.write calls"""
    rng = np.random.default_rng(3)
    desc = rng.integers(0, 256, (3, 32)).astype(np.uint8)
    b = struct.pack("<Q", 2)
    b += struct.pack("<Qfff", 11, 1.0, 2.0, 3.0) + struct.pack("<Qfff", 12, -1.5, 0.25, 8.0)
    b += struct.pack("<Q", 2)
    b += struct.pack("<Qd4f3fi", 0, 1305031102.175304, 0.0, 0.0, 0.0, 1.0, 0.1, 0.2, 0.3, 2)
    b += struct.pack("<fffffi", 10.5, 20.25, 31.0, 45.0, 77.0, 0) + struct.pack("<i", 32) + desc[0].tobytes() + struct.pack("<Q", 1)
    b += struct.pack("<fffffi", 300.0, 200.0, 37.2, 90.5, 21.0, 1) + struct.pack("<i", 32) + desc[1].tobytes() + struct.pack("<Q", ULONG_MAX)
    b += struct.pack("<Qd4f3fi", 5, 1305031102.5, 0.5, 0.5, 0.5, 0.5, -1.0, 0.0, 2.0, 1)
    b += struct.pack("<fffffi", 12.0, 22.0, 31.0, 50.0, 60.0, 0) + struct.pack("<i", 32) + desc[2].tobytes() + struct.pack("<Q", 1)
    b += struct.pack("<QQ", ULONG_MAX, 1) + struct.pack("<Qi", 5, 17)
    b += struct.pack("<QQ", 0, 1) + struct.pack("<Qi", 0, 17)
    path = os.path.join(str(tmp_path), "map.bin")
    with open(path, "wb") as f:
        f.write(b)
    return path, b, desc


def test_load_reads_the_layout_map_save_writes(tmp_path):
    path, raw, desc = _hand_packed(tmp_path)
    m = mf.MapFile.load(path)
    assert m.map_points["id"].tolist() == [11, 12] and np.array_equal(m.map_points["pos"][1], np.float32([-1.5, 0.25, 8.0]))
    assert [kf["id"] for kf in m.keyframes] == [0, 5] and m.keyframes[0]["timestamp"] == 1305031102.175304
    f0 = m.keyframes[0]["features"]
    assert f0["x"].tolist() == [10.5, 300.0] and f0["octave"].tolist() == [0, 1] and np.array_equal(f0["desc"], desc[:2])
    assert f0["mp_idx"].tolist() == [1, ULONG_MAX]
    assert m.parents.tolist() == [ULONG_MAX, 0] and m.connections[0]["id"].tolist() == [5] and m.connections[1]["weight"].tolist() == [17]
    assert m.observations() == [[], [(0, 0), (1, 0)]]
    # and save() gives back the very same bytes
    out = os.path.join(str(tmp_path), "again.bin")
    m.save(out)
    assert open(out, "rb").read() == raw
    # pose assembly of LoadKeyFrame: identity quaternion and a 120 degree turn about (1,1,1)
    assert np.array_equal(m.pose(0), np.float32([[1, 0, 0, 0.1], [0, 1, 0, 0.2], [0, 0, 1, 0.3], [0, 0, 0, 1]]))
    assert np.allclose(m.pose(1)[:3, :3], [[0, 0, 1], [1, 0, 0], [0, 1, 0]], atol=1e-7)


def test_a_keyframe_listing_one_map_point_twice_contributes_its_first_feature_only(tmp_path):
    """MapPoint::AddObservation returns early when the keyframe already observes the point (src/MapPoint.cc:122-124): the second feature of the
    same keyframe adds no descriptor row to ComputeDistinctiveDescriptors (round-1 advisor finding)"""
    path, raw, desc = _hand_packed(tmp_path)
    m = mf.MapFile.load(path)
    f0 = m.keyframes[0]["features"]
    f0["mp_idx"][1] = 1                                   # keyframe 0 now lists map point 1 at features 0 AND 1
    assert m.observations() == [[], [(0, 0), (1, 0)]]
    f0["mp_idx"][0] = ULONG_MAX                            # ... and with the first one gone, the second becomes the observation
    assert m.observations() == [[], [(0, 1), (1, 0)]]


def test_truncated_and_corrupt_files_are_refused(tmp_path):
    path, raw, _ = _hand_packed(tmp_path)
    for cut in (4, 30, 100, len(raw) - 3):
        p = os.path.join(str(tmp_path), "cut.bin")
        open(p, "wb").write(raw[:cut])
        with pytest.raises(ValueError):
            mf.MapFile.load(p)
    bad = bytearray(raw)
    bad[8 + 40 + 8 + 48 + 24] = 31                               # cols of the first feature
    p = os.path.join(str(tmp_path), "cols.bin")
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        mf.MapFile.load(p)
    empty = os.path.join(str(tmp_path), "empty.bin")
    mf.MapFile().save(empty)
    m = mf.MapFile.load(empty)
    assert len(m.map_points) == 0 and m.keyframes == [] and os.path.getsize(empty) == 16


def test_quaternion_converters_round_trip():
    rng = np.random.default_rng(4)
    for _ in range(200):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        R = mf.quaternion_to_rotation(q.astype(np.float32))
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(R) - 1) < 1e-6
        q2 = mf.rotation_to_quaternion(R)
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-6          # q and -q are the same rotation
    # the three trace <= 0 branches of Eigen's conversion
    for axis in range(3):
        R = -np.eye(3); R[axis, axis] = 1
        q = mf.rotation_to_quaternion(R)
        want = np.zeros(4); want[axis] = 1
        assert np.allclose(q, want)


def _synthetic_map(rng, n_kf=5, n_mp=400):
    from orb_slam2_aruco_b200 import synth
    from orb_slam2_aruco_b200.api import ORBextractor
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    kfs = []
    for k in range(n_kf):
        kps, desc = ex(np.roll(synth.make_frame(70), (2 * k, 3 * k), axis=(0, 1)))
        f = np.zeros(len(kps), mf.FEATURE_DTYPE)
        for name in ("x", "y", "size", "angle", "response", "octave"):
            f[name] = kps[name]
        f["cols"] = 32; f["desc"] = desc
        f["mp_idx"] = ULONG_MAX
        seen = rng.choice(len(kps), size=min(len(kps), 300), replace=False)
        f["mp_idx"][seen] = rng.choice(n_mp, size=len(seen), replace=False)
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        kfs.append({"id": 3 * k, "timestamp": 100.0 + k, "quat": q.astype(np.float32), "t": rng.normal(size=3).astype(np.float32), "features": f})
    ex.close()
    mp = np.zeros(n_mp, mf.MAPPOINT_DTYPE)
    mp["id"] = np.arange(n_mp) * 2; mp["pos"] = rng.normal(size=(n_mp, 3))
    parents = np.array([ULONG_MAX] + [3 * (k - 1) for k in range(1, n_kf)], np.uint64)
    cons = [np.array([(3 * j, 20 + j) for j in range(n_kf) if j != k], mf.CONN_DTYPE) for k in range(n_kf)]
    return mf.MapFile(mp, kfs, parents, cons)


@pytest.mark.gpu
def test_map_load_recomputation_matches_oracles(built_lib, tmp_path, golden_dir):
    """Map::Load's per-keyframe / per-map-point recomputation (Map.cc:415, 512-519), batched on the device, against the frame, BoW and
    distinctive-descriptor oracles - through a save / load round trip of a synthetic map of real ORB features"""
    from test_bow import make_tree, oracle_descend
    from orb_slam2_aruco_b200.api import CameraParameters, ORBVocabulary
    from orb_slam2_aruco_b200._lib import KP_DTYPE
    rng = np.random.default_rng(12)
    m0 = _synthetic_map(rng)
    path = os.path.join(str(tmp_path), "m.bin")
    m0.save(path)
    m = mf.MapFile.load(path)
    assert len(m.keyframes) == 5 and all(np.array_equal(a["features"], b["features"]) for a, b in zip(m.keyframes, m0.keyframes))
    cam = np.load(os.path.join(golden_dir, "frame.npz"))["cam9"]
    cp = CameraParameters([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], cam[4:9])
    tree = make_tree(rng, 10, 3)
    voc = ORBVocabulary(10, 3, *tree)
    r = mf.rebuild(m, cp, 640, 480, vocabulary=voc, levelsup=2)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    cam64 = np.ascontiguousarray(cam, np.float64)
    for k, kf in enumerate(m.keyframes):
        f = kf["features"]
        n = len(f)
        kin = np.zeros(n, KP_DTYPE)
        for name in ("x", "y", "size", "angle", "response", "octave"):
            kin[name] = f[name]
        kin["class_id"] = -1
        want = np.zeros(n, KP_DTYPE)
        oracle.lib().oracle_undistort_keypoints(P(kin), n, P(cam64), P(want))
        assert np.array_equal(r["keys_un"][k].view(np.uint8), want.view(np.uint8))
        wcs = np.zeros(64 * 48 + 1, np.int32); wci = np.zeros(max(n, 1), np.int32)
        oracle.lib().oracle_assign_grid(P(want), n, P(r["bounds"]), P(wcs), P(wci))
        assert np.array_equal(r["cell_start"][k], wcs) and np.array_equal(r["cell_items"][k], wci[:wcs[-1]])
        w, wt, nid = oracle_descend(tree, 3, np.ascontiguousarray(f["desc"]), 2)
        bow, fv = ORBVocabulary.vectors(w, wt, nid)
        assert r["bow"][k] == bow and r["featvec"][k] == fv
    obs = m.observations()
    assert max(len(o) for o in obs) >= 2
    for p, o in enumerate(obs):
        rows = np.ascontiguousarray(np.stack([m.keyframes[k]["features"]["desc"][i] for k, i in o])) if o else np.zeros((0, 32), np.uint8)
        want = oracle.lib().oracle_distinctive_descriptor(P(rows), len(o))
        assert r["best_obs"][p] == want
        if want >= 0:
            assert np.array_equal(r["descriptors"][p], rows[want])
    voc.close()
