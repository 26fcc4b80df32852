"""oracle -- TEST INFRASTRUCTURE: ctypes bindings of the CPU checkers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (orb_slam2_aruco_b200) never does.

  liboracle.so        our CPU restatement (oracle/*_oracle.cpp over oracle/cvprim*.h)
  _ref/libref_orb.so  the reference's own src/ORBextractor.cc compiled unmodified on oracle/cvshim
  _ref/libref_match.so  the reference's own src/ORBmatcher.cc compiled unmodified on oracle/matchshim
                      (built only where /root/reference exists; the prebuilt .so travels to the GPU box)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ref = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(quiet=True):
    """Compile liboracle.so (always) and _ref/libref_orb.so (only if /root/reference is present)."""
    r = subprocess.run(["make", "-C", HERE, "all", "CXX=g++"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.oracle_fast_atan2.restype = C.c_float
        _lib.oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _lib


def ref():
    """The reference's own extractor on the shim, or None when it was never built."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libref_orb.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
    return _ref


_ref_match = None


def ref_match():
    """The reference's own src/ORBmatcher.cc on oracle/matchshim (oracle/ref_match_wrap.cpp), or None when it was never built."""
    global _ref_match
    if _ref_match is None:
        path = os.path.join(HERE, "_ref", "libref_match.so")
        if not os.path.exists(path):
            return None
        _ref_match = C.CDLL(path)
    return _ref_match


_adapter_match = None


def adapter_match():
    """The PRODUCT's signature-exact ORB_SLAM2::ORBmatcher (include/b200slam_orbmatcher.hpp over libb200slam.so) behind the same wrapper and stand-in objects as
    ref_match() (oracle/ref_match_wrap.cpp built with -DB200_ADAPTER_MATCHER), or None when it was never built.  Needs a GPU to compute anything."""
    global _adapter_match
    if _adapter_match is None:
        path = os.path.join(HERE, "_ref", "libadapter_match.so")
        if not os.path.exists(path):
            return None
        _adapter_match = C.CDLL(path)
    return _adapter_match


_ref_ippe = None


def ref_ippe():
    """The reference's own pose solver (Thirdparty/aruco/aruco/ippe.cpp on oracle/ippeshim, oracle/ref_ippe_wrap.cpp), or None when never built."""
    global _ref_ippe
    if _ref_ippe is None:
        path = os.path.join(HERE, "_ref", "libref_ippe.so")
        if not os.path.exists(path):
            return None
        _ref_ippe = C.CDLL(path)
    return _ref_ippe


_ref_mappoint = None


def ref_mappoint():
    """The reference's own src/MapPoint.cc + include/MapPoint.h (and ORBmatcher.cc) on stand-in KeyFrame / Frame / Map, or None when never built."""
    global _ref_mappoint
    if _ref_mappoint is None:
        path = os.path.join(HERE, "_ref", "libref_mappoint.so")
        if not os.path.exists(path):
            return None
        _ref_mappoint = C.CDLL(path)
    return _ref_mappoint


_ref_aruco = None


def ref_aruco():
    """The reference's own marker detector (markerdetector_impl.cpp & co. on oracle/arucoshim, oracle/ref_aruco_wrap.cpp), or None when never built."""
    global _ref_aruco
    if _ref_aruco is None:
        path = os.path.join(HERE, "_ref", "libref_aruco.so")
        if not os.path.exists(path):
            return None
        _ref_aruco = C.CDLL(path)
    return _ref_aruco


def ref_aruco_detect(img, dict_name="ARUCO_MIP_25h7", cap=256):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(cap, MARKER_DTYPE)
    n = ref_aruco().ref_aruco_detect(_p(img), img.shape[1], img.shape[0], img.shape[1], dict_name.encode(), _p(out), cap)
    assert 0 <= n <= cap
    return out[:n].copy()


_ref_voc = None


def ref_voc():
    """The reference's own vendored DBoW2 on oracle/vocshim (oracle/ref_voc_wrap.cpp), or None when it was never built."""
    global _ref_voc
    if _ref_voc is None:
        path = os.path.join(HERE, "_ref", "libref_voc.so")
        if not os.path.exists(path):
            return None
        _ref_voc = C.CDLL(path)
        _ref_voc.ref_voc_load.restype = C.c_void_p
        _ref_voc.ref_voc_load.argtypes = [C.c_char_p]
        _ref_voc.ref_voc_free.argtypes = [C.c_void_p]
        _ref_voc.ref_voc_size.argtypes = [C.c_void_p]
        _ref_voc.ref_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        _ref_voc.ref_voc_score.restype = C.c_double
        _ref_voc.ref_voc_score.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return _ref_voc


_ref_dict = None


def ref_dict():
    """The reference's own dictionary.cpp / dictionary_based.cpp / markerlabeler.cpp on oracle/arucoshim (oracle/ref_dict_wrap.cpp), or None when never built."""
    global _ref_dict
    if _ref_dict is None:
        path = os.path.join(HERE, "_ref", "libref_dict.so")
        if not os.path.exists(path):
            return None
        _ref_dict = C.CDLL(path)
    return _ref_dict


def decode_patch(patch, dict_name="ARUCO_MIP_25h7", impl=None):
    """decode stage on one canonical patch -> (ok, id, nrot); impl = ref_dict() for the reference's own code, default the oracle restatement"""
    patch = np.ascontiguousarray(patch, np.uint8)
    i = C.c_int32(-1); r = C.c_int32(-1)
    f = lib().oracle_aruco_decode_patch if impl is None else impl.ref_dictionary_detect
    ok = f(_p(patch), patch.shape[0], dict_name.encode(), C.byref(i), C.byref(r))
    return (1, i.value, r.value) if ok == 1 else (0, -1, -1)


def dictionary_codes(dict_name, impl=None):
    """(nbits, tau, codes by id) of a dictionary from the oracle / product table or (impl = ref_dict()) from the reference's dictionary.cpp"""
    codes = np.zeros(8192, np.uint64); nb = C.c_int32(); tau = C.c_int32()
    f = lib().oracle_dictionary_codes if impl is None else impl.ref_dictionary_codes
    n = f(dict_name.encode(), _p(codes), len(codes), C.byref(nb), C.byref(tau))
    assert 0 <= n <= len(codes), n
    return nb.value, tau.value, codes[:n].copy()


# ---- primitives -------------------------------------------------------------------------------
def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def border_reflect101(src, border):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h + 2 * border, w + 2 * border), np.uint8)
    lib().oracle_border_reflect101(_p(src), w, h, w, _p(dst), w + 2 * border, border)
    return dst


def gaussian_blur7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().oracle_gaussian_blur7(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), src.shape[1])
    return dst


def fast_nms(img, thr):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.empty((cap, 3), np.int32)
    n = lib().oracle_fast_nms(_p(img), img.shape[1], img.shape[0], img.shape[1], int(thr), _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return float(lib().oracle_fast_atan2(float(y), float(x)))


# ---- extractor --------------------------------------------------------------------------------
def orb_levels(w, h, nfeatures=1000, scale=1.2, nlevels=8):
    lw = np.empty(nlevels, np.int32); lh = np.empty(nlevels, np.int32); q = np.empty(nlevels, np.int32)
    sf = np.empty(nlevels, np.float32)
    lib().oracle_orb_levels(w, h, nfeatures, C.c_float(scale), nlevels, _p(lw), _p(lh), _p(q), _p(sf))
    return lw, lh, q, sf


def orb_pyramid_level(img, level, scale=1.2, nlevels=8):
    img = np.ascontiguousarray(img, np.uint8)
    lw, lh, _, _ = orb_levels(img.shape[1], img.shape[0], 1000, scale, nlevels)
    out = np.empty((lh[level], lw[level]), np.uint8)
    lib().oracle_orb_pyramid_level(_p(img), img.shape[1], img.shape[0], img.shape[1], C.c_float(scale), nlevels, level, _p(out))
    return out


def orb_candidates(img, level, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
    img = np.ascontiguousarray(img, np.uint8)
    cap = 200000
    out = np.empty((cap, 3), np.int32)
    n = lib().oracle_orb_candidates(_p(img), img.shape[1], img.shape[0], img.shape[1], nfeatures, C.c_float(scale), nlevels,
                                    ini_th, min_th, level, _p(out), cap)
    assert n >= 0
    return out[:n].copy()


def orb_extract(img, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
    """-> (keypoints structured array [n], descriptors uint8 [n,32])"""
    img = np.ascontiguousarray(img, np.uint8)
    cap = nfeatures + 3 * nlevels + 64
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    if img.size == 0:
        return kps[:0], desc[:0]
    n = lib().oracle_orb_extract(_p(img), img.shape[1], img.shape[0], img.shape[1], nfeatures, C.c_float(scale), nlevels,
                                 ini_th, min_th, _p(kps), _p(desc), cap)
    assert n >= 0
    return kps[:n].copy(), desc[:n].copy()


def orb_extract_batch(imgs, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, nthreads=1):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    cap = nfeatures + 3 * nlevels + 64
    kps = np.zeros((n, cap), KP_DTYPE)
    desc = np.zeros((n, cap, 32), np.uint8)
    counts = np.zeros(n, np.int32)
    lib().oracle_orb_extract_batch(_p(imgs), n, w, h, w, C.c_long(w * h), nfeatures, C.c_float(scale), nlevels, ini_th, min_th,
                                   _p(kps), _p(desc), _p(counts), cap, nthreads)
    return kps, desc, counts


def ref_orb_extract(img, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
    """The reference's own ORBextractor::operator() (src/ORBextractor.cc:1043) through oracle/_ref."""
    r = ref()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_orb.so not built (needs /root/reference)")
    img = np.ascontiguousarray(img, np.uint8)
    cap = nfeatures + 3 * nlevels + 64
    raw = np.zeros((cap, 7), np.float32)
    desc = np.zeros((cap, 32), np.uint8)
    n = r.ref_orb_extract(_p(img), img.shape[1], img.shape[0], img.shape[1], nfeatures, C.c_float(scale), nlevels,
                          ini_th, min_th, _p(raw), _p(desc), cap)
    assert n >= 0
    kps = np.zeros(n, KP_DTYPE)
    for i, f in enumerate(("x", "y", "size", "angle", "response")):
        kps[f] = raw[:n, i]
    kps["octave"] = raw[:n, 5].astype(np.int32)
    kps["class_id"] = raw[:n, 6].astype(np.int32)
    return kps, desc[:n].copy()


def ref_orb_pyramid_level(img, level, scale=1.2, nlevels=8):
    r = ref()
    img = np.ascontiguousarray(img, np.uint8)
    lw, lh, _, _ = orb_levels(img.shape[1], img.shape[0], 1000, scale, nlevels)
    out = np.empty((lh[level] + 38, lw[level] + 38), np.uint8)
    wl = C.c_int(); hl = C.c_int()
    r.ref_orb_pyramid_level(_p(img), img.shape[1], img.shape[0], img.shape[1], C.c_float(scale), nlevels, level, _p(out),
                            C.byref(wl), C.byref(hl))
    assert (wl.value, hl.value) == (lw[level], lh[level])
    return out


# ---- matcher ----------------------------------------------------------------------------------
def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return int(lib().oracle_descriptor_distance(_p(a), _p(b)))


def search_by_bow_bf(kf_desc, kf_angle, f_desc, f_angle, nnratio=0.7, check_ori=True, factor=None):
    """ORBmatcher::SearchByBoW(KF, F) with one all-inclusive vocabulary node -> (nmatches, matches[n_f])"""
    kf_desc = np.ascontiguousarray(kf_desc, np.uint8).reshape(-1, 32)
    f_desc = np.ascontiguousarray(f_desc, np.uint8).reshape(-1, 32)
    kf_angle = np.ascontiguousarray(kf_angle, np.float32); f_angle = np.ascontiguousarray(f_angle, np.float32)
    if factor is None:
        factor = np.float32(30.0) / np.float32(360.0)
    m = np.empty(len(f_desc), np.int32)
    n = lib().oracle_search_by_bow_bf(_p(kf_desc), _p(kf_angle), len(kf_desc), _p(f_desc), _p(f_angle), len(f_desc),
                                      C.c_float(nnratio), int(check_ori), C.c_float(factor), _p(m))
    return int(n), m


def match_candidates(qd, td, ofs, cand):
    qd = np.ascontiguousarray(qd, np.uint8).reshape(-1, 32); td = np.ascontiguousarray(td, np.uint8).reshape(-1, 32)
    ofs = np.ascontiguousarray(ofs, np.int32); cand = np.ascontiguousarray(cand, np.int32)
    bi, bd, sd = (np.empty(len(qd), np.int32) for _ in range(3))
    lib().oracle_match_candidates(_p(qd), len(qd), _p(td), _p(ofs), _p(cand), _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


# ---- ArUco detector ---------------------------------------------------------------------------
MARKER_DTYPE = np.dtype([("id", "<i4"), ("xy", "<f4", (8,))])


def aruco_detect(img, dict_name="ARUCO_MIP_25h7", cap=256):
    """aruco::MarkerDetector::detect on the reference's path -> structured array (id, xy[8]) sorted by id"""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(cap, MARKER_DTYPE)
    n = lib().oracle_aruco_detect(_p(img), img.shape[1], img.shape[0], img.shape[1], dict_name.encode(), _p(out), cap)
    assert n >= 0, n
    return out[:n].copy()


def aruco_detect_batch(imgs, dict_name="ARUCO_MIP_25h7", cap=64, nthreads=1):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    out = np.zeros((n, cap), MARKER_DTYPE)
    counts = np.zeros(n, np.int32)
    r = lib().oracle_aruco_detect_batch(_p(imgs), n, w, h, w, C.c_long(w * h), dict_name.encode(), _p(out), _p(counts), cap, nthreads)
    assert r == 0
    return out, counts


def aruco_stages(img, dict_name="ARUCO_MIP_25h7"):
    """stage outputs: dict(thres, contours [list of (n,2) arrays], candidates (k,4,2), patches (k,ws,ws), prerefine, markers)"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    thres = np.zeros((h, w), np.uint8)
    maxc, maxp, maxk, cap = 20000, 2000000, 2000, 256
    sizes = np.zeros(maxc, np.int32); pts = np.zeros((maxp, 2), np.int32); nc = C.c_int()
    cand = np.zeros((maxk, 8), np.float32); nk = C.c_int()
    from orb_slam2_aruco_b200 import synth
    nbits = synth.dictionaries()[dict_name][0]
    ws = 5 * (int(round(nbits ** 0.5)) + 2)
    patches = np.zeros((maxk, ws, ws), np.uint8)
    pre = np.zeros((cap, 8), np.float32)
    out = np.zeros(cap, MARKER_DTYPE)
    n = lib().oracle_aruco_stages(_p(img), w, h, w, dict_name.encode(), _p(thres), _p(sizes), _p(pts), maxc, maxp, C.byref(nc),
                                  _p(cand), _p(patches), maxk, C.byref(nk), _p(pre), _p(out), cap)
    assert n >= 0
    contours = []
    o = 0
    for i in range(nc.value):
        contours.append(pts[o:o + sizes[i]].copy()); o += sizes[i]
    return dict(thres=thres, contours=contours, candidates=cand[:nk.value].reshape(-1, 4, 2).copy(), patches=patches[:nk.value].copy(),
                prerefine=pre[:n].reshape(-1, 4, 2).copy(), markers=out[:n].copy())


def adaptive_threshold(img, bs, c=7):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().oracle_adaptive_threshold(_p(img), img.shape[1], img.shape[0], _p(out), bs, c)
    return out


def find_contours(img):
    img = np.ascontiguousarray(img, np.uint8)
    maxc, maxp = 50000, 4000000
    sizes = np.zeros(maxc, np.int32); pts = np.zeros((maxp, 2), np.int32)
    n = lib().oracle_find_contours(_p(img), img.shape[1], img.shape[0], _p(sizes), _p(pts), maxc, maxp)
    out = []; o = 0
    for i in range(n):
        out.append(pts[o:o + sizes[i]].copy()); o += sizes[i]
    return out


def approx_poly(pts, eps):
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    out = np.zeros((len(pts) + 4, 2), np.int32)
    cv = C.c_int()
    n = lib().oracle_approx_poly(_p(pts), len(pts), C.c_double(eps), _p(out), C.byref(cv))
    return out[:n].copy(), bool(cv.value)


def resize_half(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty((img.shape[0] // 2, img.shape[1] // 2), np.uint8)
    lib().oracle_resize_half(_p(img), img.shape[1], img.shape[0], _p(out))
    return out


def perspective_transform(src, dst):
    src = np.ascontiguousarray(src, np.float32).reshape(8); dst = np.ascontiguousarray(dst, np.float32).reshape(8)
    M = np.zeros(9, np.float64)
    lib().oracle_perspective_transform(_p(src), _p(dst), _p(M))
    return M.reshape(3, 3)


def warp_perspective(img, M, size):
    img = np.ascontiguousarray(img, np.uint8)
    M = np.ascontiguousarray(M, np.float64)
    out = np.empty((size, size), np.uint8)
    lib().oracle_warp_perspective(_p(img), img.shape[1], img.shape[0], _p(out), size, _p(M))
    return out


def otsu(img):
    img = np.ascontiguousarray(img, np.uint8)
    return int(lib().oracle_otsu(_p(img), img.shape[1], img.shape[0]))


def solve_svd(A, b):
    A = np.ascontiguousarray(A, np.float32); b = np.ascontiguousarray(b, np.float32).reshape(-1)
    x = np.zeros(A.shape[1], np.float32)
    lib().oracle_solve_svd(_p(A), _p(b), A.shape[0], A.shape[1], _p(x))
    return x
