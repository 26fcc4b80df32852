"""Host glue of the KeyFrame-side ORBmatcher members: the per-CALL quantities the reference computes once before its loop over map points, in
its cv::Mat CV_32F arithmetic (matrix products accumulate in double and are stored as float): pose decomposition (src/ORBmatcher.cc:302-307,
831-841), the Sim3 pair of SearchBySim3 (:1123-1126), the epipole of SearchForTriangulation (:668-674), the scale pyramid tables, and the
PredictScale thresholds.  Everything per POINT (projection, tests, predicted level, radius: k_kf_project; grid queries and Hamming search:
k_features_in_area, k_radius_best, k_proj_resolve) runs on the device."""
import functools
import struct
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


@functools.lru_cache(maxsize=8)
def level_thresholds(scale_factor=1.2, nlevels=8):
    """MapPoint::PredictScale (src/MapPoint.cc:403-435) is ceil(logf(ratio) / logf(scaleFactor)) clamped to [0, nlevels): monotone in ratio.
    thresholds[n] = the largest float ratio that still yields a level <= n, found by bisection over float bit patterns with THIS host's logf (the
    one the reference would call), so that level = #(ratio > thresholds[n]) equals the reference's integer without a device logf."""
    log_sf = f32(_libm.logf(float(f32(scale_factor))))

    def level(bits):
        r = struct.unpack("<f", struct.pack("<I", bits))[0]
        return int(np.ceil(f32(f32(_libm.logf(r)) / log_sf)))
    out = np.zeros(max(nlevels - 1, 1), f32)
    for n in range(nlevels - 1):
        lo, hi = 0x00800000, 0x7f000000                           # smallest normal float (level far below 0) .. 2^127 (far above any level)
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if level(mid) <= n:
                lo = mid
            else:
                hi = mid
        out[n] = struct.unpack("<f", struct.pack("<I", lo))[0]
    return out


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
