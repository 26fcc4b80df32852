"""Host glue of the KeyFrame-side ORBmatcher members: the projection of map points into a keyframe as the reference computes it with cv::Mat
on CV_32F (src/ORBmatcher.cc:294-358, 831-895, 983-1061, 1106-1195): matrix products accumulate in double and are stored as float, sums and
differences are float, MapPoint::PredictScale (src/MapPoint.cc:403-435) goes through the C library's logf like the reference's std::log(float).
Everything here is O(points) arithmetic that prepares the radius queries for b200_match_kf_radius_host / b200_match_by_projection_host; the
searches themselves run on the device."""
import ctypes
import ctypes.util

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.logf.restype = ctypes.c_float
_libm.logf.argtypes = [ctypes.c_float]


def pyramid(scale_factor=1.2, nlevels=8):
    """mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2 (src/ORBextractor.cc:418-428) and mfLogScaleFactor (src/Frame.cc:96), float arithmetic"""
    sf = np.ones(nlevels, f32); s2 = np.ones(nlevels, f32)
    f = f32(scale_factor)
    for i in range(1, nlevels):
        sf[i] = sf[i - 1] * f
        s2[i] = sf[i] * sf[i]
    return sf, s2, (f32(1.0) / s2).astype(f32), f32(_libm.logf(float(f)))


def _mul_points(M, X):
    """rows of X (N x 3, float32) through the 3 x 3 float32 matrix M: double accumulation in index order, float result"""
    M64, X64 = M.astype(np.float64), X.astype(np.float64)
    out = np.empty((len(X), 3), f32)
    for r in range(3):
        s = M64[r, 0] * X64[:, 0]
        s = s + M64[r, 1] * X64[:, 1]
        s = s + M64[r, 2] * X64[:, 2]
        out[:, r] = s.astype(f32)
    return out


def _dot_rows(A, B):
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    s = A64[:, 0] * B64[:, 0]
    s = s + A64[:, 1] * B64[:, 1]
    return s + A64[:, 2] * B64[:, 2]


def pose_from_T(T):
    """GetRotation / GetTranslation / GetCameraCenter of a row-major 4 x 4 Tcw -> (R, t, Ow)"""
    T = np.asarray(T, f32).reshape(4, 4)
    R, t = np.ascontiguousarray(T[:3, :3]), np.ascontiguousarray(T[:3, 3])
    return R, t, _mul_points(-R.T, t[None])[0]


def pose_from_S(S):
    """the Sim3 decomposition of src/ORBmatcher.cc:302-307 -> (Rcw, tcw, Ow)"""
    S = np.asarray(S, f32).reshape(4, 4)
    sR = S[:3, :3]
    scw = f32(np.sqrt(_dot_rows(sR[:1], sR[:1])[0]))
    R = (sR.astype(np.float64) / np.float64(scw)).astype(f32)
    t = (S[:3, 3].astype(np.float64) / np.float64(scw)).astype(f32)
    return R, t, _mul_points(-R.T, t[None])[0]


def predict_scale(max_distance, dist, log_sf, nlevels):
    """MapPoint::PredictScale, element-wise; entries with a non-positive or non-finite ratio get level 0 (they are discarded by the callers' tests)"""
    with np.errstate(all="ignore"):
        ratio = (np.asarray(max_distance, f32) / np.asarray(dist, f32)).astype(f32)
    out = np.zeros(len(ratio), np.int32)
    for i, r in enumerate(ratio):
        if not np.isfinite(r) or r <= 0:
            continue
        n = int(np.ceil(f32(f32(_libm.logf(float(r))) / log_sf)))
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


def project_points(pose, cam4, bounds4, pos, normal, minmax, th, scale_factor=1.2, nlevels=8):
    """the part of Fuse / Fuse(Scw) / SearchByProjection(Scw) between "Get 3D Coords" and GetFeaturesInArea for all points at once.
    pose = (R, t, Ow).  Returns (valid [N] bool, q_xyr [N, 3] float32, level [N] int32); rows with valid = False were discarded by one of the tests."""
    R, t, Ow = pose
    pos = np.ascontiguousarray(pos, f32).reshape(-1, 3); normal = np.ascontiguousarray(normal, f32).reshape(-1, 3)
    minmax = np.ascontiguousarray(minmax, f32).reshape(-1, 2)
    sf, _, _, log_sf = pyramid(scale_factor, nlevels)
    pc = (_mul_points(R, pos) + t).astype(f32)
    valid = ~(pc[:, 2] < 0)
    u, v, _ = _pixel(pc, cam4)
    valid &= _in_image(u, v, bounds4)
    maxd, mind = f32(1.2) * minmax[:, 1], f32(0.8) * minmax[:, 0]
    PO = (pos - Ow).astype(f32)
    dist = np.sqrt(_dot_rows(PO, PO)).astype(f32)
    valid &= ~(dist < mind) & ~(dist > maxd)
    valid &= ~(_dot_rows(PO, normal) < 0.5 * dist.astype(np.float64))
    level = np.zeros(len(pos), np.int32)
    idx = np.nonzero(valid)[0]
    level[idx] = predict_scale(minmax[idx, 1], dist[idx], log_sf, nlevels)
    radius = (f32(th) * sf[level]).astype(f32)
    return valid, np.ascontiguousarray(np.stack([u, v, radius], 1), f32), level


def project_points_sim3(pose_a, sR, tt, cam4, bounds4, pos, minmax, th, scale_factor=1.2, nlevels=8):
    """one direction of SearchBySim3 (src/ORBmatcher.cc:1158-1195): world point -> camera a -> camera b = sR * p + tt -> pixel in keyframe b"""
    R, t, _ = pose_a
    pos = np.ascontiguousarray(pos, f32).reshape(-1, 3); minmax = np.ascontiguousarray(minmax, f32).reshape(-1, 2)
    sf, _, _, log_sf = pyramid(scale_factor, nlevels)
    pa = (_mul_points(R, pos) + t).astype(f32)
    pb = (_mul_points(sR, pa) + tt).astype(f32)
    valid = ~(pb[:, 2] < 0)
    u, v, _ = _pixel(pb, cam4)
    valid &= _in_image(u, v, bounds4)
    maxd, mind = f32(1.2) * minmax[:, 1], f32(0.8) * minmax[:, 0]
    dist = np.sqrt(_dot_rows(pb, pb)).astype(f32)
    valid &= ~(dist < mind) & ~(dist > maxd)
    level = np.zeros(len(pos), np.int32)
    idx = np.nonzero(valid)[0]
    level[idx] = predict_scale(minmax[idx, 1], dist[idx], log_sf, nlevels)
    radius = (f32(th) * sf[level]).astype(f32)
    return valid, np.ascontiguousarray(np.stack([u, v, radius], 1), f32), level


def sim3_between(s12, R12, t12):
    """sR12, sR21, t21 of src/ORBmatcher.cc:1123-1126"""
    R12 = np.asarray(R12, f32).reshape(3, 3); t12 = np.asarray(t12, f32).reshape(3)
    sR12 = (np.float64(f32(s12)) * R12.astype(np.float64)).astype(f32)
    sR21 = ((1.0 / np.float64(f32(s12))) * R12.T.astype(np.float64)).astype(f32)
    t21 = _mul_points(-sR21, t12[None])[0]
    return sR12, sR21, t21


def epipole(T1, T2, cam4):
    """the epipole of camera 1 in image 2 (src/ORBmatcher.cc:668-674)"""
    _, _, Ow1 = pose_from_T(T1)
    R2, t2, _ = pose_from_T(T2)
    C2 = (_mul_points(R2, Ow1[None])[0] + t2).astype(f32)
    cam4 = np.asarray(cam4, f32)
    with np.errstate(all="ignore"):
        invz = f32(1.0) / C2[2]
        return np.array([cam4[0] * C2[0] * invz + cam4[2], cam4[1] * C2[1] * invz + cam4[3]], f32)
