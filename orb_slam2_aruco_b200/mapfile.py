"""The reference's on-disk map: `Map::Save` / `Map::Load` (src/Map.cc:219-330 / 339-426) as a byte-exact reader and
writer, plus the data-parallel part of `Map::Load` - what `LoadKeyFrame` and the tail of `Load` recompute for every keyframe and map
point - on the B200 path:

    InitKeyFrame::UndistortKeyPoints + AssignFeaturesToGrid   (Map.cc:512-513)  -> FrameGrid          (csrc/frame.cu)
    KeyFrame::ComputeBoW                                      (Map.cc:519)      -> ORBVocabulary      (csrc/bow.cu)
    MapPoint::ComputeDistinctiveDescriptors                   (Map.cc:415)      -> ORBmatcher.ComputeDistinctiveDescriptors (csrc/match.cu)

File layout (little endian, `unsigned long` = 8 bytes, no padding; Map.cc:219-330):
    u64 nMapPoints; nMapPoints x { u64 mnId; f32 x, y, z }
    u64 nKeyFrames; nKeyFrames x { u64 mnId; f64 mTimeStamp; f32 quat[4] (x, y, z, w); f32 t[3]; i32 N;
                                   N x { f32 pt.x, pt.y, size, angle, response; i32 octave; i32 cols (= 32); u8 desc[32]; u64 mapPointIdx } }
    nKeyFrames x { u64 parent mnId (ULONG_MAX: none); u64 nb_con; nb_con x { u64 mnId; i32 weight } }
`mapPointIdx` is the position of the map point in the map-point block above (Map::GetMapPointsIdx), ULONG_MAX for none.
The pose / graph bookkeeping (SetPose, ChangeParent, AddConnection, UpdateNormalAndDepth) is the reference's host glue and stays there.
"""
import numpy as np

ULONG_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)

MAPPOINT_DTYPE = np.dtype([("id", "<u8"), ("pos", "<f4", (3,))])
FEATURE_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
                          ("cols", "<i4"), ("desc", "u1", (32,)), ("mp_idx", "<u8")])
KF_HEAD_DTYPE = np.dtype([("id", "<u8"), ("timestamp", "<f8"), ("quat", "<f4", (4,)), ("t", "<f4", (3,)), ("N", "<i4")])
CONN_DTYPE = np.dtype([("id", "<u8"), ("weight", "<i4")])
assert MAPPOINT_DTYPE.itemsize == 20 and FEATURE_DTYPE.itemsize == 68 and KF_HEAD_DTYPE.itemsize == 48 and CONN_DTYPE.itemsize == 12


def quaternion_to_rotation(q):
    """Converter::toCvMat(const std::vector<float>&) (src/Converter.cc:92-103): Eigen::Quaterniond(x, y, z, w) -> 3x3, computed in double
    without normalising (like Eigen's toRotationMatrix) and stored as float"""
    x, y, z, w = (float(v) for v in q)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]], np.float64).astype(np.float32)


def rotation_to_quaternion(R):
    """Converter::toQuaternion (src/Converter.cc:150-162): Eigen::Quaterniond(Matrix3d) -> float (x, y, z, w)"""
    m = np.asarray(R, np.float64)
    q = [0.0, 0.0, 0.0]
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        q = [(m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t]
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        w = (m[k, j] - m[j, k]) * t
        q[j] = (m[j, i] + m[i, j]) * t
        q[k] = (m[k, i] + m[i, k]) * t
    return np.array([q[0], q[1], q[2], w], np.float32)


class MapFile:
    """What Map::Save writes: `map_points` (MAPPOINT_DTYPE, in set order), `keyframes` = list of dicts {id, timestamp, quat, t, features
    (FEATURE_DTYPE [N])}, `parents` [nKF] u64 and `connections` = list of CONN_DTYPE arrays, both in keyframe order."""

    def __init__(self, map_points=None, keyframes=None, parents=None, connections=None):
        self.map_points = np.zeros(0, MAPPOINT_DTYPE) if map_points is None else np.asarray(map_points, MAPPOINT_DTYPE)
        self.keyframes = list(keyframes or [])
        self.parents = np.full(len(self.keyframes), ULONG_MAX, np.uint64) if parents is None else np.asarray(parents, np.uint64)
        self.connections = [np.zeros(0, CONN_DTYPE) for _ in self.keyframes] if connections is None else [np.asarray(c, CONN_DTYPE) for c in connections]
        if len(self.parents) != len(self.keyframes) or len(self.connections) != len(self.keyframes):
            raise ValueError("one parent and one connection list per keyframe")

    # Map::Save (src/Map.cc:219-265), SaveMapPoint (267-276), SaveKeyFrame (278-330)
    def save(self, path):
        with open(path, "wb") as f:
            f.write(np.uint64(len(self.map_points)).tobytes())
            f.write(np.ascontiguousarray(self.map_points).tobytes())
            f.write(np.uint64(len(self.keyframes)).tobytes())
            for kf in self.keyframes:
                feats = np.ascontiguousarray(kf["features"], FEATURE_DTYPE)
                if len(feats) and (feats["cols"] != 32).any():
                    raise ValueError("descriptor rows are 32 bytes (Map.cc:311)")
                head = np.zeros(1, KF_HEAD_DTYPE)
                head["id"] = kf["id"]; head["timestamp"] = kf["timestamp"]; head["quat"] = kf["quat"]; head["t"] = kf["t"]; head["N"] = len(feats)
                f.write(head.tobytes())
                f.write(feats.tobytes())
            for parent, con in zip(self.parents, self.connections):
                f.write(np.uint64(parent).tobytes())
                f.write(np.uint64(len(con)).tobytes())
                f.write(np.ascontiguousarray(con).tobytes())

    # Map::Load (src/Map.cc:339-426), LoadMapPoint (428-445), LoadKeyFrame (447-533): the parsing half
    @classmethod
    def load(cls, path):
        buf = np.fromfile(path, np.uint8)
        pos = 0

        def take(dtype, count):
            nonlocal pos
            nbytes = dtype.itemsize * count
            if pos + nbytes > len(buf):
                raise ValueError("map file truncated at byte %d" % pos)
            out = buf[pos:pos + nbytes].view(dtype).copy()
            pos += nbytes
            return out

        u64 = np.dtype("<u8")
        n_mp = int(take(u64, 1)[0])
        map_points = take(MAPPOINT_DTYPE, n_mp)
        n_kf = int(take(u64, 1)[0])
        keyframes = []
        for _ in range(n_kf):
            head = take(KF_HEAD_DTYPE, 1)[0]
            if head["N"] < 0:
                raise ValueError("negative feature count")
            feats = take(FEATURE_DTYPE, int(head["N"]))
            if len(feats) and (feats["cols"] != 32).any():
                raise ValueError("descriptor rows are 32 bytes (Map.cc:493-496)")
            bad = (feats["mp_idx"] != ULONG_MAX) & (feats["mp_idx"] >= np.uint64(n_mp))
            if bad.any():
                raise ValueError("map point index out of range")
            keyframes.append({"id": int(head["id"]), "timestamp": float(head["timestamp"]), "quat": head["quat"].copy(), "t": head["t"].copy(),
                              "features": feats})
        parents = np.zeros(n_kf, np.uint64)
        connections = []
        for i in range(n_kf):
            parents[i] = take(u64, 1)[0]
            connections.append(take(CONN_DTYPE, int(take(u64, 1)[0])))
        return cls(map_points, keyframes, parents, connections)

    def pose(self, i):
        """Tcw of keyframe i as LoadKeyFrame assembles it (Map.cc:455-468)"""
        T = np.zeros((4, 4), np.float32)
        T[:3, :3] = quaternion_to_rotation(self.keyframes[i]["quat"])
        T[:3, 3] = self.keyframes[i]["t"]
        T[3, 3] = 1
        return T

    def observations(self):
        """per map point: [(keyframe position, feature index)] in keyframe load order - MapPoint::AddObservation as LoadKeyFrame calls it
        (Map.cc:521-529).  The reference iterates its std::map<KeyFrame*, size_t> in POINTER order, which no file can pin; load order is
        the deterministic choice, and it only matters when two observations tie on the median distance.  AddObservation returns early when the
        keyframe already observes the point (MapPoint.cc:122-124): a keyframe listing the same map point at two features contributes its FIRST
        feature only.  mp_idx is resolved in file order here; Map::Load resolves it through GetAllMapPoints(), a std::set<MapPoint*> in pointer
        order (Map.cc:428-433, 521-529), which equals file order only while the allocator hands out ascending addresses."""
        obs = [[] for _ in range(len(self.map_points))]
        for k, kf in enumerate(self.keyframes):
            idx = kf["features"]["mp_idx"]
            seen = set()
            for i in np.nonzero(idx != ULONG_MAX)[0]:
                p = int(idx[i])
                if p in seen:
                    continue
                seen.add(p)
                obs[p].append((k, int(i)))
        return obs


def rebuild(mapfile, camera_params, width, height, vocabulary=None, device=0, levelsup=4):
    """The recomputation half of Map::Load on the B200 path, batched over the whole map.  All keyframes in ONE launch each: undistorted
    keypoints + 64x48 grid (InitKeyFrame::UndistortKeyPoints / AssignFeaturesToGrid, Map.cc:512-513); with a vocabulary the BowVector /
    FeatureVector of every keyframe (KeyFrame::ComputeBoW, Map.cc:519) from one tree descent over all descriptors; all map points in one
    launch for the distinctive descriptor (Map.cc:415).  Returns {"keys_un": [KP_DTYPE arrays], "cell_start": [nKF, 64*48+1],
    "cell_items": [int arrays], "bounds": [4], "bow": [...], "featvec": [...], "best_obs": [P], "descriptors": [P, 32]}."""
    import torch
    from ._lib import KP_DTYPE
    from .api import FrameGrid, ORBmatcher
    nkf = len(mapfile.keyframes)
    counts = np.array([len(kf["features"]) for kf in mapfile.keyframes], np.int32)
    cap = max(int(counts.max()) if nkf else 0, 1)
    grid = FrameGrid(width, height, camera_params, device=device)
    out = {"bounds": grid.bounds.copy(), "keys_un": [], "cell_items": [], "cell_start": np.zeros((nkf, FrameGrid.COLS * FrameGrid.ROWS + 1), np.int32),
           "bow": [None] * nkf, "featvec": [None] * nkf}
    if nkf:
        kps = np.zeros((nkf, cap), KP_DTYPE)
        for k, kf in enumerate(mapfile.keyframes):
            f = kf["features"]
            for name in ("x", "y", "size", "angle", "response", "octave"):
                kps[k, :len(f)][name] = f[name]
            kps[k, :len(f)]["class_id"] = -1
        dev = torch.device("cuda", device)
        d_k = torch.from_numpy(kps.view(np.uint8).reshape(nkf, cap, KP_DTYPE.itemsize).copy()).to(dev)
        d_un = torch.zeros_like(d_k)
        d_cnt = torch.from_numpy(counts).to(dev)
        d_cs = torch.zeros((nkf, FrameGrid.COLS * FrameGrid.ROWS + 1), dtype=torch.int32, device=dev)
        d_ci = torch.zeros((nkf, cap), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            grid.undistort(d_k, d_cnt, d_un)
            grid.assign(d_un, d_cnt, d_cs, d_ci)
            torch.cuda.synchronize()
        un = d_un.cpu().numpy().reshape(nkf, cap * KP_DTYPE.itemsize).view(KP_DTYPE).reshape(nkf, cap)
        ci = d_ci.cpu().numpy()
        out["cell_start"] = d_cs.cpu().numpy()
        for k in range(nkf):
            out["keys_un"].append(un[k, :counts[k]].copy())
            out["cell_items"].append(ci[k, :out["cell_start"][k, -1]].copy())
        if vocabulary is not None and counts.sum():
            # one descent for every descriptor of the map, split back per keyframe (transform is per feature; the maps are per keyframe)
            alld = np.concatenate([np.ascontiguousarray(kf["features"]["desc"]) for kf in mapfile.keyframes])
            word, weight, node = vocabulary.descend(alld, levelsup)
            ofs = np.concatenate([[0], np.cumsum(counts)])
            for k in range(nkf):
                sl = slice(ofs[k], ofs[k + 1])
                out["bow"][k], out["featvec"][k] = vocabulary.vectors(word[sl], weight[sl], node[sl])
    obs = mapfile.observations()
    rows = [np.stack([mapfile.keyframes[k]["features"]["desc"][i] for k, i in o]) if o else np.zeros((0, 32), np.uint8) for o in obs]
    out["best_obs"], out["descriptors"] = ORBmatcher(device=device).ComputeDistinctiveDescriptors(rows)
    return out
