"""Host-side mirror of the reference's operator interface for the front-end hot path.

Same class names, argument meaning and error behaviour as the reference's three C++ classes, on top
of the C-ABI (include/b200slam.h); the C++ twin lives in include/b200slam_adapters.hpp.

  ORBextractor   <- ORB_SLAM2::ORBextractor     (reference include/ORBextractor.h:45-111)
  ORBmatcher     <- ORB_SLAM2::ORBmatcher       (reference include/ORBmatcher.h:38-104)
  MarkerDetector <- aruco::MarkerDetector       (reference Thirdparty/aruco/aruco/markerdetector.h:58-410)

Everything here is marshaling; all arithmetic happens in the CUDA kernels.  Batched entry points
(`extract_batch`, `detect_batch`, `SearchByBoW_batch`) are the additions that let one call feed a B200.
"""
import ctypes as C

import numpy as np

from . import _lib, kfgeom
from ._lib import KP_DTYPE, MARKER_DTYPE, B200Error, check, lib, ptr


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                 max_width=None, max_height=None, max_batch=1, device=0):
        """reference: ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
        (ORBextractor.h:51-52).  max_width/max_height/max_batch size the device scratch; when omitted the
        handle is (re)created lazily for the first image it sees."""
        self._args = (int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST))
        self._device = int(device)
        self._h = None
        self._dims = (0, 0, 0)
        if max_width and max_height:
            self._create(int(max_width), int(max_height), int(max_batch))
        else:
            # validate the parameters eagerly, like a constructor would
            if nlevels < 1 or not scaleFactor > 1.0 or nfeatures < 0:
                raise B200Error(_lib.EINVAL, "bad extractor parameters")

    # -- lifetime ------------------------------------------------------------------------------
    def _create(self, w, h, b):
        self.close()
        hnd = C.c_void_p()
        check(lib().b200_orb_create(C.byref(hnd), *self._args, w, h, b, self._device))
        self._h = hnd
        self._dims = (w, h, b)
        self.cap = check(lib().b200_orb_max_keypoints(self._h))

    def _ensure(self, w, h, b):
        W, H, B = self._dims
        if self._h is None or w > W or h > H or b > B:
            self._create(max(w, W, 1), max(h, H, 1), max(b, B, 1))

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().b200_orb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference getters (ORBextractor.h:63-83) -------------------------------------------------
    def _level_info(self):
        self._ensure(64, 64, 1)
        n = self._args[2]
        sf, isf, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        q = np.empty(n, np.int32)
        nl = C.c_int()
        check(lib().b200_orb_get_level_info(self._h, C.addressof(nl), ptr(sf), ptr(isf), ptr(s2), ptr(is2), ptr(q)))
        return nl.value, sf, isf, s2, is2, q

    def GetLevels(self):
        return self._args[2]

    def GetScaleFactor(self):
        return self._args[1]

    def GetScaleFactors(self):
        return self._level_info()[1]

    def GetInverseScaleFactors(self):
        return self._level_info()[2]

    def GetScaleSigmaSquares(self):
        return self._level_info()[3]

    def GetInverseScaleSigmaSquares(self):
        return self._level_info()[4]

    def GetFeaturesPerLevel(self):
        return self._level_info()[5]

    @property
    def mvImagePyramid(self):
        """reference public member (ORBextractor.h:85): level images of the last frame; each entry is the
        w_l x h_l view into a buffer that carries the 19-px REFLECT_101 border (ORBextractor.cc:1107-1132)."""
        out = []
        for level in range(self._args[2]):
            out.append(self.pyramid_level(0, level)[19:-19, 19:-19])
        return out

    def pyramid_level(self, frame, level):
        """bordered (h_l+38, w_l+38) level buffer of frame `frame` of the last call"""
        W, H, _ = self._dims
        buf = np.empty((H + 38) * (W + 38), np.uint8)
        wl, hl = C.c_int(), C.c_int()
        check(lib().b200_orb_get_pyramid(self._h, frame, level, ptr(buf), C.addressof(wl), C.addressof(hl)))
        return buf[:(hl.value + 38) * (wl.value + 38)].reshape(hl.value + 38, wl.value + 38)

    def set_profile(self, enable=True):
        self._ensure(64, 64, 1)
        check(lib().b200_orb_set_profile(self._h, int(enable)))

    def stage_ms(self):
        """(pyramid, fast, quadtree, describe) device milliseconds of the last profiled call"""
        ms = np.zeros(4, np.float32)
        check(lib().b200_orb_get_stage_ms(self._h, ptr(ms)))
        return ms

    def stage_frames(self):
        """frames the profiled launches of the last call processed (large batches run as two halves)"""
        return check(lib().b200_orb_get_stage_frames(self._h))

    def candidates(self, frame, level):
        """debug tap: FAST candidates (x, y, score) of one level of the last call, before the quadtree"""
        cap = 1 << 20
        out = np.empty((cap, 3), np.int32)
        n = check(lib().b200_orb_get_candidates(self._h, frame, level, ptr(out), cap))
        return out[:n].copy()

    # -- the operator (ORBextractor.h:59-61, ORBextractor.cc:1043-1105) -----------------------------
    def __call__(self, image, mask=None):
        """(keypoints, descriptors) of one gray frame.  `mask` is ignored, as in the reference
        (ORBextractor.h:58).  Empty image => no keypoints (ORBextractor.cc:1046); non-u8 => AssertionError
        (the reference asserts CV_8UC1, ORBextractor.cc:1050)."""
        image = np.asarray(image)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "image.type() == CV_8UC1"
        kps, desc, counts = self.extract_batch(image[None])
        n = int(counts[0])
        return kps[0, :n].copy(), desc[0, :n].copy()

    def extract_batch(self, images, out=None):
        """images: uint8 [n, h, w] host array (row-contiguous; row/frame strides are honoured).
        Returns host arrays kps [n, cap] (KP_DTYPE), desc [n, cap, 32], counts [n].  `out` may pass those three
        arrays preallocated (e.g. numpy views of pinned memory, which are then written without staging)."""
        images = np.asarray(images)
        assert images.dtype == np.uint8 and images.ndim == 3, "images must be uint8 [n,h,w]"
        if images.strides[2] != 1:
            images = np.ascontiguousarray(images)
        n, h, w = images.shape
        self._ensure(w, h, n)
        if out is not None:
            kps, desc, counts = out
            assert kps.nbytes >= n * self.cap * 28 and desc.nbytes >= n * self.cap * 32 and counts.nbytes >= n * 4
        else:
            kps = np.zeros((n, self.cap), KP_DTYPE)
            desc = np.zeros((n, self.cap, 32), np.uint8)
            counts = np.zeros(n, np.int32)
        if n:
            check(lib().b200_orb_extract_host(self._h, ptr(images), n, w, h, images.strides[1], images.strides[0],
                                              ptr(kps), ptr(desc), ptr(counts)))
        return kps, desc, counts

    def extract_batch_device(self, images, kps, desc, counts, stream=None):
        """device-resident variant: torch uint8 CUDA tensor [n,h,w] in, preallocated CUDA tensors out
        (kps: uint8/[n,cap,28] or float32 [n,cap,7] view, desc: uint8 [n,cap,32], counts: int32 [n]);
        only enqueues kernels on `stream` (a torch.cuda.Stream, an int handle, or None = handle stream)."""
        n, h, w = images.shape
        self._ensure(w, h, n)
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream
        check(lib().b200_orb_extract(self._h, ptr(images), n, w, h, images.stride(1), images.stride(0),
                                     ptr(kps), ptr(desc), ptr(counts), s))


class ORBmatcher:
    TH_LOW = 50          # ORBmatcher.cc:37-39
    TH_HIGH = 100
    HISTO_LENGTH = 30

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        """reference: ORBmatcher(float nnratio=0.6, bool checkOri=true) (ORBmatcher.h:41)"""
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self._device = int(device)

    @staticmethod
    def DescriptorDistance(a, b, device=0):
        """Hamming distance of two 32-byte descriptors (ORBmatcher.h:44, ORBmatcher.cc:1651-1667)"""
        a = np.ascontiguousarray(a, np.uint8).reshape(1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(1, 32)
        out = np.empty((1, 1), np.int32)
        check(lib().b200_hamming_matrix_host(ptr(a), 1, ptr(b), 1, ptr(out), device))
        return int(out[0, 0])

    def distance_matrix(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.empty((len(a), len(b)), np.int32)
        check(lib().b200_hamming_matrix_host(ptr(a), len(a), ptr(b), len(b), ptr(out), self._device))
        return out

    def SearchByBoW(self, kf_desc, kf_angles, f_desc, f_angles):
        """SearchByBoW(KeyFrame, Frame, vpMapPointMatches) (ORBmatcher.h:55, ORBmatcher.cc:159-292) for the
        case every keyframe feature owns a good MapPoint and all features share one vocabulary node (the
        brute-force configuration of SURVEY.md section 8a).  Returns (nmatches, matches) where matches[i] is
        the keyframe index matched to frame keypoint i or -1 (the reference's vpMapPointMatches)."""
        f_desc = np.ascontiguousarray(f_desc, np.uint8).reshape(-1, 32)
        f_angles = np.ascontiguousarray(f_angles, np.float32)
        n, m = self.SearchByBoW_batch(kf_desc, kf_angles, f_desc[None], f_angles[None], np.array([len(f_desc)], np.int32))
        return int(n[0]), m[0]

    def SearchByBoW_KF(self, desc1, angles1, desc2, angles2):
        """SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (ORBmatcher.h:56, ORBmatcher.cc:526-659) for one all-inclusive
        vocabulary node and good MapPoints everywhere: strict bestDist1 < TH_LOW and the upstream factor 1.0f / HISTO_LENGTH.
        Returns (nmatches, matches12) with matches12[idx1] = index in KF2 or -1.  Same greedy kernel as SearchByBoW (KF1 rows are
        the sequential side, KF2 features the ones that get taken); only the direction of the answer differs."""
        d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        a2 = np.ascontiguousarray(angles2, np.float32)
        d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32)
        a1 = np.ascontiguousarray(angles1, np.float32)
        m21 = np.full((1, max(len(d2), 1)), -1, np.int32)
        nm = np.zeros(1, np.int32)
        nf = np.array([len(d2)], np.int32)
        check(lib().b200_match_bf_host(ptr(d1), ptr(a1), len(d1), ptr(d2), ptr(a2), ptr(nf), 1, len(d2), self.mfNNratio, self.TH_LOW - 1,
                                       int(self.mbCheckOrientation), float(np.float32(1.0) / np.float32(self.HISTO_LENGTH)), ptr(m21), ptr(nm), self._device))
        m12 = np.full(len(d1), -1, np.int32)
        idx2 = np.nonzero(m21[0, :len(d2)] >= 0)[0]
        m12[m21[0, idx2]] = idx2
        return int(nm[0]), m12

    @staticmethod
    def common_node_groups(featvec_q, valid_q, featvec_c, valid_c=None):
        """The merge walk of SearchByBoW over two FeatureVectors ({node: [feature indices]}, ascending nodes like the std::map;
        ORBmatcher.cc:185-279): one group per common node.  Query features without a good MapPoint are dropped here (the reference
        `continue`s on them), and in the KeyFrame-KeyFrame form (valid_c given) so are such candidates.
        Returns (grp_q_ofs, q_idx, grp_c_ofs, c_idx) for b200_match_by_bow_host."""
        gq, gc, qi, ci = [0], [0], [], []
        for node in sorted(set(featvec_q) & set(featvec_c)):
            qs = [i for i in featvec_q[node] if valid_q is None or valid_q[i]]
            cs = [i for i in featvec_c[node] if valid_c is None or valid_c[i]]
            qi += qs; ci += cs
            gq.append(len(qi)); gc.append(len(ci))
        a = lambda v: np.asarray(v, np.int32)
        return a(gq), a(qi), a(gc), a(ci)

    def _by_bow(self, mode, d1, a1, d2, a2, groups):
        d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
        d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        a1 = np.ascontiguousarray(a1, np.float32)
        a2 = np.ascontiguousarray(a2, np.float32)
        gq, qi, gc, ci = (np.ascontiguousarray(g, np.int32) for g in groups)
        out = np.full(max(len(d2) if mode == 0 else len(d1), 1), -1, np.int32)
        n = check(lib().b200_match_by_bow_host(ptr(d1), ptr(a1), len(d1), ptr(d2), ptr(a2), len(d2), ptr(gq), ptr(qi), ptr(gc), ptr(ci), len(gq) - 1,
                                               mode, self.mfNNratio, self.TH_LOW, int(self.mbCheckOrientation), ptr(out), self._device))
        return n, out[:len(d2) if mode == 0 else len(d1)]

    def SearchByBoW_nodes(self, kf_desc, kf_angles, kf_valid, kf_featvec, f_desc, f_angles, f_featvec):
        """SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (ORBmatcher.h:55, ORBmatcher.cc:159-292) with the real FeatureVectors
        (ORBVocabulary.transform): per common node, every keyframe feature with a good MapPoint (kf_valid) takes its best free frame
        feature of the same node.  Returns (nmatches, matches) with matches[idxF] = keyframe index or -1."""
        return self._by_bow(0, kf_desc, kf_angles, f_desc, f_angles, self.common_node_groups(kf_featvec, kf_valid, f_featvec))

    def SearchByBoW_KF_nodes(self, desc1, angles1, valid1, featvec1, desc2, angles2, valid2, featvec2):
        """SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (ORBmatcher.h:56, ORBmatcher.cc:526-659) with the real FeatureVectors.
        Returns (nmatches, matches12) with matches12[idx1] = index in KF2 or -1."""
        return self._by_bow(1, desc1, angles1, desc2, angles2, self.common_node_groups(featvec1, valid1, featvec2, valid2))

    def ComputeDistinctiveDescriptors(self, observations):
        """MapPoint::ComputeDistinctiveDescriptors (MapPoint.h:75, MapPoint.cc:271-331) for a list of map points: observations[p] is the
        [N_p, 32] array of the descriptors of the keyframes observing point p (bad keyframes dropped).  Returns (best_idx [P] with -1 for
        a point without observations, descriptors [P, 32])."""
        obs = [np.ascontiguousarray(o, np.uint8).reshape(-1, 32) for o in observations]
        ofs = np.zeros(len(obs) + 1, np.int32)
        ofs[1:] = np.cumsum([len(o) for o in obs])
        desc = np.concatenate(obs) if obs and ofs[-1] else np.zeros((1, 32), np.uint8)
        best = np.full(max(len(obs), 1), -1, np.int32)
        out = np.zeros((max(len(obs), 1), 32), np.uint8)
        check(lib().b200_distinctive_descriptors_host(ptr(desc), ptr(ofs), len(obs), ptr(best), ptr(out), self._device))
        return best[:len(obs)], out[:len(obs)]

    def SearchByBoW_batch(self, kf_desc, kf_angles, f_desc, f_angles, n_frame, histo_factor=None):
        kf_desc = np.ascontiguousarray(kf_desc, np.uint8).reshape(-1, 32)
        kf_angles = np.ascontiguousarray(kf_angles, np.float32)
        f_desc = np.ascontiguousarray(f_desc, np.uint8)
        f_angles = np.ascontiguousarray(f_angles, np.float32)
        n_frame = np.ascontiguousarray(n_frame, np.int32)
        nb, cap = f_desc.shape[0], f_desc.shape[1]
        if histo_factor is None:
            histo_factor = np.float32(self.HISTO_LENGTH) / np.float32(360.0)     # ORBmatcher.cc:176
        matches = np.full((nb, cap), -1, np.int32)
        nm = np.zeros(nb, np.int32)
        check(lib().b200_match_bf_host(ptr(kf_desc), ptr(kf_angles), len(kf_desc), ptr(f_desc), ptr(f_angles), ptr(n_frame), nb, cap,
                                       self.mfNNratio, self.TH_LOW, int(self.mbCheckOrientation), float(histo_factor),
                                       ptr(matches), ptr(nm), self._device))
        return nm, matches

    def SearchByBoW_device(self, ref_desc, ref_kps, n_ref, f_desc, f_kps, n_frame, matches, n_matches, stream=None):
        """device-resident batch variant: all arguments are CUDA tensors laid out like the extractor's outputs
        (desc uint8 [.., cap, 32], kps [.., cap, 7 floats], n_frame int32 [n]); only enqueues kernels"""
        nb, cap = f_desc.shape[0], f_desc.shape[1]
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream
        hf = float(np.float32(self.HISTO_LENGTH) / np.float32(360.0))
        check(lib().b200_match_bf_kp(ptr(ref_desc), ptr(ref_kps), int(n_ref), ptr(f_desc), ptr(f_kps), ptr(n_frame), nb, cap,
                                     self.mfNNratio, self.TH_LOW, int(self.mbCheckOrientation), hf, ptr(matches), ptr(n_matches),
                                     self._device, s))

    def match_candidates(self, query_desc, train_desc, cand_ofs, cand):
        """distance core of SearchByProjection / SearchForInitialization: per query the best index, best and
        second-best distance over its candidate list (first minimum in list order wins)"""
        query_desc = np.ascontiguousarray(query_desc, np.uint8).reshape(-1, 32)
        train_desc = np.ascontiguousarray(train_desc, np.uint8).reshape(-1, 32)
        cand_ofs = np.ascontiguousarray(cand_ofs, np.int32)
        cand = np.ascontiguousarray(cand, np.int32)
        nq = len(query_desc)
        bi, bd, sd = (np.empty(nq, np.int32) for _ in range(3))
        check(lib().b200_match_candidates_host(ptr(query_desc), nq, ptr(train_desc), len(train_desc), ptr(cand_ofs), ptr(cand),
                                               ptr(bi), ptr(bd), ptr(sd), self._device))
        return bi, bd, sd


    # ---- KeyFrame-side members (local mapping / loop closing threads): per-call glue in kfgeom.py, projection and searches on the device ---------------------
    def project_points(self, pose, cam4, bounds4, pos, normal, minmax, th, sim3=None, scale_factor=1.2, nlevels=8):
        """the projection front of Fuse / Fuse(Scw) / SearchByProjection(Scw) / SearchBySim3 for all points at once (b200_kf_project_host).
        pose = (R, t, Ow) of the keyframe; sim3 = (sR, t) for SearchBySim3's second transform.  Returns (valid bool [N], q_xyr [N, 3], level [N])."""
        R, t, Ow = (np.ascontiguousarray(a, np.float32) for a in pose)
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3); mm = np.ascontiguousarray(minmax, np.float32).reshape(-1, 2)
        nrm = None if normal is None else np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
        sR, tt = (None, None) if sim3 is None else (np.ascontiguousarray(sim3[0], np.float32), np.ascontiguousarray(sim3[1], np.float32))
        cam = np.ascontiguousarray(cam4, np.float32); b = np.ascontiguousarray(bounds4, np.float32)
        sf = np.ascontiguousarray(kfgeom.pyramid(scale_factor, nlevels)[0], np.float32)
        thr = np.ascontiguousarray(kfgeom.level_thresholds(float(scale_factor), int(nlevels)), np.float32)
        n = len(pos)
        valid = np.zeros(max(n, 1), np.uint8); q3 = np.zeros((max(n, 1), 3), np.float32); level = np.zeros(max(n, 1), np.int32)
        check(lib().b200_kf_project_host(ptr(R), ptr(t), ptr(Ow), ptr(sR), ptr(tt), ptr(cam), ptr(b), ptr(pos), ptr(nrm), ptr(mm), n, float(th), ptr(sf), ptr(thr),
                                         int(nlevels), ptr(valid), ptr(q3), ptr(level), self._device))
        return valid[:n].astype(bool), q3[:n], level[:n]

    def search_points(self, kps_un, desc, bounds4, pose, cam4, pos, normal, minmax, q_desc, skip, th, chi2=0.0, sim3=None, scale_factor=1.2, nlevels=8):
        """projection + keyframe search of all points in one device call (b200_kf_search_points_host): what Fuse / Fuse(Scw) / one direction of
        SearchBySim3 do per point up to `if(bestDist<=TH_...)`.  skip[i]: the reference `continue`s on point i before projecting it.
        Returns (valid bool [N], best_idx [N], best_dist [N])."""
        k = np.ascontiguousarray(kps_un); assert k.dtype == KP_DTYPE
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        R, t, Ow = (np.ascontiguousarray(a, np.float32) for a in pose)
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3); mm = np.ascontiguousarray(minmax, np.float32).reshape(-1, 2)
        nrm = None if normal is None else np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
        sR, tt = (None, None) if sim3 is None else (np.ascontiguousarray(sim3[0], np.float32), np.ascontiguousarray(sim3[1], np.float32))
        qd = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32); sk = np.ascontiguousarray(skip, np.uint8)
        cam = np.ascontiguousarray(cam4, np.float32); b = np.ascontiguousarray(bounds4, np.float32)
        sf, _, inv, _ = kfgeom.pyramid(scale_factor, nlevels)
        sf = np.ascontiguousarray(sf, np.float32); inv = np.ascontiguousarray(inv, np.float32)
        thr = np.ascontiguousarray(kfgeom.level_thresholds(float(scale_factor), int(nlevels)), np.float32)
        n = len(pos)
        assert len(qd) == n and len(sk) == n and len(mm) == n
        valid = np.zeros(max(n, 1), np.uint8); bi = np.full(max(n, 1), -1, np.int32); bd = np.full(max(n, 1), 256, np.int32)
        check(lib().b200_kf_search_points_host(ptr(k), ptr(d), len(k), ptr(b), ptr(R), ptr(t), ptr(Ow), ptr(sR), ptr(tt), ptr(cam), ptr(pos), ptr(nrm), ptr(mm), ptr(qd),
                                               ptr(sk), n, float(th), ptr(sf), ptr(inv), ptr(thr), int(nlevels), C.c_double(chi2), ptr(valid), ptr(bi), ptr(bd),
                                               None, None, self._device))
        return valid[:n].astype(bool), bi[:n], bd[:n]

    def kf_radius_search(self, kps_un, desc, bounds4, q_xyr, q_level, q_desc, chi2=0.0, scale_factor=1.2, nlevels=8):
        """best keyframe feature per projected point (b200_match_kf_radius_host) -> (best_idx, best_dist)"""
        k = np.ascontiguousarray(kps_un); assert k.dtype == KP_DTYPE
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        q3 = np.ascontiguousarray(q_xyr, np.float32).reshape(-1, 3); ql = np.ascontiguousarray(q_level, np.int32)
        qd = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(bounds4, np.float32)
        inv = np.ascontiguousarray(kfgeom.pyramid(scale_factor, nlevels)[2], np.float32)
        bi = np.full(max(len(q3), 1), -1, np.int32); bd = np.full(max(len(q3), 1), 256, np.int32)
        check(lib().b200_match_kf_radius_host(ptr(k), ptr(d), len(k), ptr(b), ptr(q3), ptr(ql), ptr(qd), len(q3), ptr(inv), int(nlevels), C.c_double(chi2),
                                              ptr(bi), ptr(bd), self._device))
        return bi[:len(q3)], bd[:len(q3)]

    def SearchForTriangulation(self, kps1_un, desc1, has_mp1, featvec1, T1, kps2_un, desc2, has_mp2, featvec2, T2, cam4, F12, scale_factor=1.2, nlevels=8):
        """SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo=false) (ORBmatcher.h:66-67, ORBmatcher.cc:661-829), monocular keyframes.
        has_mp*: the feature already owns a MapPoint; featvec*: {node: [feature indices]}; T*: Tcw.  Returns (nmatches, matches12) - the pairs
        (i, matches12[i]) with matches12[i] >= 0, in ascending i, are vMatchedPairs."""
        k1, k2 = np.ascontiguousarray(kps1_un), np.ascontiguousarray(kps2_un)
        assert k1.dtype == KP_DTYPE and k2.dtype == KP_DTYPE
        d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        free1 = ~np.asarray(has_mp1, bool); free2 = ~np.asarray(has_mp2, bool)
        gq, qi, gc, ci = (np.ascontiguousarray(g, np.int32) for g in self.common_node_groups(featvec1, free1, featvec2, free2))
        sf, s2, _, _ = kfgeom.pyramid(scale_factor, nlevels)
        F = np.ascontiguousarray(F12, np.float32).reshape(9); e = kfgeom.epipole(T1, T2, cam4)
        m12 = np.full(max(len(k1), 1), -1, np.int32)
        n = check(lib().b200_match_for_triangulation_host(ptr(k1), ptr(d1), len(k1), ptr(k2), ptr(d2), len(k2), ptr(gq), ptr(qi), ptr(gc), ptr(ci), len(gq) - 1,
                                                          ptr(F), ptr(e), ptr(sf), ptr(s2), int(nlevels), int(self.mbCheckOrientation), self.TH_LOW, ptr(m12),
                                                          self._device))
        return n, m12[:len(k1)]

    def FuseReal(self, kps_un, desc, bounds4, cam4, Tcw, held_state, held_nobs, mp_state, mp_pos, mp_normal, mp_desc, mp_minmax, mp_nobs, th=3.0):
        """Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th) with the outcome statements (ORBmatcher.cc:957-976) applied the way the reference's
        REAL MapPoint / KeyFrame objects behave: MapPoint::Replace (MapPoint.cc:192-236) re-points the keyframe's slot at the survivor and hands it the
        observation, AddObservation / AddMapPoint (MapPoint.cc:110-121) fill an empty slot, so a later point of the same call meets whoever holds the slot by
        then.  Same scene arguments as Fuse (every point observes at most this keyframe).  Returns dict(n = nFused, slot [n_kf] = holder of each feature
        afterwards (-1 nobody, 1000000 + j the keyframe's own point j, m list point m), mp_bad, held_bad, mp_nobs = Observations() afterwards)."""
        st = np.asarray(mp_state, np.uint8); nobs = np.asarray(mp_nobs, np.int32).copy()
        hs = np.asarray(held_state, np.uint8); hn = np.asarray(held_nobs, np.int32).copy()
        valid, bi, bd = self.search_points(kps_un, desc, bounds4, kfgeom.pose_from_T(Tcw), cam4, mp_pos, mp_normal, mp_minmax, mp_desc, st != 1, th, chi2=5.99)
        n_kf = len(hs)
        slot = np.where(hs > 0, 1000000 + np.arange(n_kf), -1).astype(np.int64)
        held_bad = hs == 2; mp_bad = st == 2
        in_kf = {}                                                   # list point -> feature it observes in this keyframe
        nf = 0

        def is_bad(who):
            return held_bad[who - 1000000] if who >= 1000000 else mp_bad[who]

        def observations(who):
            return hn[who - 1000000] if who >= 1000000 else nobs[who]

        def replace(x, y):
            """x->Replace(y): x dies; its observation of this keyframe goes to y, or the slot is cleared when y already observes the keyframe"""
            xi = x - 1000000 if x >= 1000000 else in_kf.get(x)
            if x >= 1000000:
                held_bad[x - 1000000] = True
            else:
                mp_bad[x] = True
            if xi is None:
                return
            y_in = True if y >= 1000000 else (y in in_kf)
            if not y_in:
                slot[xi] = y; in_kf[y] = xi; nobs[y] += 1
            else:
                slot[xi] = -1
        for m in np.nonzero(valid)[0]:
            idx, dist = int(bi[m]), int(bd[m])
            if mp_bad[m] or m in in_kf or dist > self.TH_LOW:
                continue
            owner = int(slot[idx])
            if owner >= 0:
                if not is_bad(owner):
                    if observations(owner) > observations(m):
                        replace(m, owner)
                    else:
                        replace(owner, m)
            else:
                slot[idx] = m; in_kf[m] = idx; nobs[m] += 1
            nf += 1
        return dict(n=nf, slot=slot.astype(np.int32), mp_bad=mp_bad.astype(np.uint8), held_bad=held_bad.astype(np.uint8), mp_nobs=nobs)

    def Fuse(self, kps_un, desc, bounds4, cam4, Tcw, held_state, held_nobs, mp_state, mp_pos, mp_normal, mp_desc, mp_minmax, mp_nobs, th=3.0):
        """Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th = 3.0) (ORBmatcher.h:76, ORBmatcher.cc:831-981) on arrays.  Keyframe: features,
        pose, held_state [n] 0 none / 1 good / 2 bad point at that feature with held_nobs observations.  Map point m: mp_state 0 NULL / 1 good / 2 bad /
        3 already observed by the keyframe, position, normal, descriptor, (mfMinDistance, mfMaxDistance), Observations().
        Returns (nFused, fused_idx, action): action 1 = added as a new observation of feature fused_idx, 2 = pMP->Replace(pMPinKF), 3 =
        pMPinKF->Replace(pMP); fused_idx names the keyframe feature, or -3 - o when pMPinKF is point o of this call that was added earlier."""
        st = np.asarray(mp_state, np.uint8); nobs = np.asarray(mp_nobs, np.int32).copy()
        hs = np.asarray(held_state, np.uint8); hn = np.asarray(held_nobs, np.int32)
        valid, bi, bd = self.search_points(kps_un, desc, bounds4, kfgeom.pose_from_T(Tcw), cam4, mp_pos, mp_normal, mp_minmax, mp_desc, st != 1, th, chi2=5.99)
        qs = np.nonzero(valid)[0]; bi, bd = bi[qs], bd[qs]
        # the outcome, applied in list order exactly as the reference does on its objects (ORBmatcher.cc:957-976)
        holder = np.where(hs > 0, -2, -1).astype(np.int64); own_bad = hs == 2
        bad = st == 2
        fused = np.full(len(st), -1, np.int32); action = np.zeros(len(st), np.int32)
        nf = 0
        for m, idx, dist in zip(qs, bi, bd):
            if bad[m] or dist > self.TH_LOW:
                continue
            h = holder[idx]
            if h == -2:
                if not own_bad[idx]:
                    if hn[idx] > nobs[m]:
                        bad[m] = True; fused[m], action[m] = idx, 2
                    else:
                        own_bad[idx] = True; fused[m], action[m] = idx, 3
            elif h >= 0:
                if not bad[h]:
                    if nobs[h] > nobs[m]:
                        bad[m] = True; fused[m], action[m] = -3 - h, 2
                    else:
                        bad[h] = True; fused[m], action[m] = -3 - h, 3
            else:
                holder[idx] = m; nobs[m] += 1; fused[m], action[m] = idx, 1
            nf += 1
        return nf, fused, action

    def FuseSim3(self, kps_un, desc, bounds4, cam4, Scw, held_state, mp_state, mp_pos, mp_normal, mp_desc, mp_minmax, th=4.0):
        """Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint) (ORBmatcher.h:79,
        ORBmatcher.cc:983-1104).  mp_state 1 good / 2 bad / 3 one of the keyframe's own good points.  Returns (nFused, replace_idx, added_idx):
        replace_idx[m] = keyframe feature whose point vpReplacePoint[m] names (-3 - o: point o added earlier in this call), added_idx[m] = feature the
        point became an observation of."""
        st = np.asarray(mp_state, np.uint8); hs = np.asarray(held_state, np.uint8)
        valid, bi, bd = self.search_points(kps_un, desc, bounds4, kfgeom.pose_from_S(Scw), cam4, mp_pos, mp_normal, mp_minmax, mp_desc, (st == 2) | (st == 3), th)
        qs = np.nonzero(valid)[0]; bi, bd = bi[qs], bd[qs]
        holder = np.where(hs > 0, -2, -1).astype(np.int64)
        rep = np.full(len(st), -1, np.int32); add = np.full(len(st), -1, np.int32)
        nf = 0
        for m, idx, dist in zip(qs, bi, bd):
            if idx < 0 or dist > self.TH_LOW:
                continue
            if holder[idx] == -2:
                if hs[idx] != 2:
                    rep[m] = idx
            elif holder[idx] >= 0:
                rep[m] = -3 - holder[idx]
            else:
                holder[idx] = m; add[m] = idx
            nf += 1
        return nf, rep, add

    def SearchByProjectionLoop(self, kps_un, desc, bounds4, cam4, Scw, mp_state, mp_pos, mp_normal, mp_desc, mp_minmax, matched, th=10):
        """SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th) (ORBmatcher.h:52,
        ORBmatcher.cc:294-407).  matched [n] = vpMatched before the call (-1 none, >= 0 index into vpPoints, -2 another point).  A matched keyframe
        feature is hidden from every later point, so the queries are replayed in order by the projection resolve kernel
        (b200_match_by_projection_host, mode 2 = mode 1 over the keyframe's grid origin, without the rotation histogram, TH_LOW).  Returns (nmatches, vpMatched after the call)."""
        st = np.asarray(mp_state, np.uint8); matched = np.asarray(matched, np.int32).copy()
        valid, q3, level = self.project_points(kfgeom.pose_from_S(Scw), cam4, bounds4, mp_pos, mp_normal, mp_minmax, int(th))
        found = np.zeros(len(st), bool); found[matched[matched >= 0]] = True
        valid &= (st != 2) & ~found
        qs = np.nonzero(valid)[0]
        lv = np.stack([level[qs] - 1, level[qs]], 1).astype(np.int32)
        n, assign, _ = search_by_projection(kps_un, desc, bounds4, (matched != -1).astype(np.uint8), q3[qs], lv, np.asarray(mp_desc, np.uint8).reshape(-1, 32)[qs],
                                            np.zeros(len(qs), np.float32), np.ones(len(qs), np.uint8), 2, check_ori=False, th_high=self.TH_LOW, device=self._device)
        hit = assign >= 0
        matched[hit] = qs[assign[hit]]
        return n, matched

    def SearchBySim3(self, kps1_un, desc1, T1, mp1_state, mp1_pos, mp1_desc, mp1_minmax, kps2_un, desc2, T2, mp2_state, mp2_pos, mp2_desc, mp2_minmax,
                     bounds4, cam4, matches12, s12, R12, t12, th=7.5):
        """SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (ORBmatcher.h:70-71, ORBmatcher.cc:1106-1330).  Feature i of keyframe x owns a map point
        with state mpx_state[i] (0 none / 1 good / 2 bad).  matches12 [n1]: KF2 feature whose point is already matched to KF1 feature i (-1 none).
        Both directions are searched on the device, the agreement check (:1311-1327) is two gathers.  Returns (nFound, matches12 after the call)."""
        st1, st2 = np.asarray(mp1_state, np.uint8), np.asarray(mp2_state, np.uint8)
        m12 = np.asarray(matches12, np.int32).copy()
        n1, n2 = len(st1), len(st2)
        done1 = m12 >= 0
        done2 = np.zeros(n2, bool)
        j = m12[done1]; done2[j[st2[j] > 0]] = True                  # GetIndexInKeyFrame(pKF2) >= 0 only when that KF2 feature holds the point
        sR12, sR21, t21 = kfgeom.sim3_between(s12, R12, t12)
        v1, bi1, bd1 = self.search_points(kps2_un, desc2, bounds4, kfgeom.pose_from_T(T1), cam4, mp1_pos, None, mp1_minmax, mp1_desc, (st1 != 1) | done1, th,
                                          sim3=(sR21, t21))
        v2, bi2, bd2 = self.search_points(kps1_un, desc1, bounds4, kfgeom.pose_from_T(T2), cam4, mp2_pos, None, mp2_minmax, mp2_desc, (st2 != 1) | done2, th,
                                          sim3=(sR12, np.asarray(t12, np.float32).reshape(3)))
        a, b = np.nonzero(v1)[0], np.nonzero(v2)[0]
        bi1, bd1, bi2, bd2 = bi1[a], bd1[a], bi2[b], bd2[b]
        vn1 = np.full(n1, -1, np.int64); vn2 = np.full(n2, -1, np.int64)
        ok = (bi1 >= 0) & (bd1 <= self.TH_HIGH); vn1[a[ok]] = bi1[ok]
        ok = (bi2 >= 0) & (bd2 <= self.TH_HIGH); vn2[b[ok]] = bi2[ok]
        i1 = np.nonzero(vn1 >= 0)[0]
        agree = i1[vn2[vn1[i1]] == i1]
        m12[agree] = vn1[agree]
        return len(agree), m12


class CameraParameters:
    """aruco::CameraParameters essentials (Thirdparty/aruco/aruco/cameraparameters.h): 3x3 camera matrix + distortion
    (k1 k2 p1 p2 [k3]), both float like the reference's mK / mDistCoef (src/Frame.cc:132)."""

    def __init__(self, camera_matrix, distortion=None, cam_size=None):
        K = np.asarray(camera_matrix, np.float32).reshape(3, 3)
        d = np.zeros(5, np.float32)
        if distortion is not None:
            dd = np.asarray(distortion, np.float32).ravel()
            d[:min(5, len(dd))] = dd[:5]
        self.CameraMatrix, self.Distorsion, self.CamSize = K, d, cam_size

    def isValid(self):
        return self.CameraMatrix[0, 0] != 0 and self.CameraMatrix[1, 1] != 0

    def resized(self, width, height):
        """CameraParameters::resize(cv::Size) (cameraparameters.cpp:158-173) as a copy: float factors, fx cx scaled by the width ratio, fy cy by the height
        ratio.  MarkerDetector::detect applies it whenever CamSize differs from the image (markerdetector_impl.cpp, detect(input, markers, camParams, ...)) -
        always, in the reference, whose CamSize is the hard-coded 1280 x 720 of src/Frame.cc:132."""
        if self.CamSize is None or tuple(self.CamSize) == (width, height):
            return self
        ax = np.float32(width) / np.float32(self.CamSize[0]); ay = np.float32(height) / np.float32(self.CamSize[1])
        K = self.CameraMatrix.copy()
        K[0, 0] *= ax; K[0, 2] *= ax; K[1, 1] *= ay; K[1, 2] *= ay
        return CameraParameters(K, self.Distorsion, (width, height))

    def cam9(self):
        K = self.CameraMatrix
        return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], *self.Distorsion], np.float32)


def search_for_initialization(kps1_un, desc1, kps2_un, desc2, bounds4, prev_matched, window=100, nnratio=0.9, check_ori=True, device=0):
    """ORBmatcher(0.9, true).SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, 100) (src/Tracking.cc:606-607,
    src/ORBmatcher.cc:409-524) on arrays -> (nmatches, vnMatches12, updated vbPrevMatched)"""
    k1, k2 = np.ascontiguousarray(kps1_un), np.ascontiguousarray(kps2_un)
    assert k1.dtype == KP_DTYPE and k2.dtype == KP_DTYPE
    d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
    prev = np.ascontiguousarray(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.full(len(k1), -1, np.int32)
    b = np.ascontiguousarray(bounds4, np.float32)
    n = check(lib().b200_match_for_initialization_host(ptr(k1), ptr(d1), len(k1), ptr(k2), ptr(d2), len(k2), ptr(b), ptr(prev), int(window),
                                                       float(nnratio), int(bool(check_ori)), ptr(m12), int(device)))
    return n, m12, prev


def search_by_projection(kps_un, desc, bounds4, occupied, q_xyr, q_levels, q_desc, q_angle, q_observed, mode, nnratio=0.8, check_ori=True, th_high=100, device=0):
    """SearchByProjection(Frame, mapPoints, th) (mode 0, ORBmatcher.cc:45-129) / SearchByProjection(Current, Last, th, mono) (mode 1,
    ORBmatcher.cc:1332-1474; with th_high = ORBdist the relocalisation variant :1476-1603) on ready-made projections -> (nmatches, assign[n_frame], updated occupied)"""
    k = np.ascontiguousarray(kps_un); assert k.dtype == KP_DTYPE
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    occ = np.ascontiguousarray(occupied, np.uint8).copy()
    q3 = np.ascontiguousarray(q_xyr, np.float32).reshape(-1, 3); lv = np.ascontiguousarray(q_levels, np.int32).reshape(-1, 2)
    qd = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32); qa = np.ascontiguousarray(q_angle, np.float32); qo = np.ascontiguousarray(q_observed, np.uint8)
    assign = np.full(len(k), -1, np.int32)
    b = np.ascontiguousarray(bounds4, np.float32)
    n = check(lib().b200_match_by_projection_host(ptr(k), ptr(d), len(k), ptr(b), ptr(occ), ptr(q3), ptr(lv), ptr(qd), ptr(qa), ptr(qo), len(q3), int(mode),
                                                  float(nnratio), int(bool(check_ori)), int(th_high), ptr(assign), int(device)))
    assign[assign == -2] = -1              # -2 = matched, then cleared by the rotation histogram: no match either way (the C++ adapter needs the difference)
    return n, assign, occ


class Marker:
    """aruco::Marker essentials (Thirdparty/aruco/aruco/marker.h:47-59): id + 4 corners, ordered by id; Rvec / Tvec / ssize
    once extrinsics have been computed (marker.cpp:322-343), plus the second IPPE solution and both reprojection errors"""
    __slots__ = ("id", "corners", "Rvec", "Tvec", "ssize", "Rvec2", "Tvec2", "err1", "err2")

    def __init__(self, id, corners):
        self.id = int(id)
        self.corners = np.asarray(corners, np.float32).reshape(4, 2)
        self.Rvec = self.Tvec = self.Rvec2 = self.Tvec2 = None
        self.ssize, self.err1, self.err2 = -1.0, None, None

    def __lt__(self, other):
        return self.id < other.id

    def __repr__(self):
        return "Marker(id=%d, corners=%s)" % (self.id, self.corners.tolist())


class MarkerDetector:
    def __init__(self, dict_name="ALL_DICTS", max_width=None, max_height=None, max_batch=1, device=0):
        """reference: MarkerDetector(std::string dict_name) + setDictionary (markerdetector.h:229,337); the
        configuration is the one src/Frame.cc:129-139 applies (DM_NORMAL, CORNER_LINES)."""
        self._dict = dict_name
        self._device = int(device)
        self._h = None
        self._dims = (0, 0, 0)
        if max_width and max_height:
            self._create(int(max_width), int(max_height), int(max_batch))

    def setDictionary(self, dict_name, error_correction_rate=0.0):
        assert error_correction_rate == 0.0, "only error_correction_rate 0 (the reference's setting) is supported"
        self._dict = dict_name
        W, H, B = self._dims
        if self._h is not None:
            self._create(W, H, B)

    def _create(self, w, h, b):
        self.close()
        if self._dict == "ALL_DICTS":
            # the reference's default-constructed detector searches several dictionaries at once (dictionary_based.cpp setParams); its configured
            # path never does: src/Frame.cc:133 always calls setDictionary(System::mArucoDic) before the first detect
            raise B200Error(_lib.EINVAL, "multi-dictionary ALL_DICTS detection is not on the reference's configured path (src/Frame.cc:133): "
                                         "call setDictionary(name) or pass dict_name before detect")
        hnd = C.c_void_p()
        check(lib().b200_aruco_create(C.byref(hnd), self._dict.encode(), w, h, b, self._device))
        self._h = hnd
        self._dims = (w, h, b)
        self.cap = check(lib().b200_aruco_max_markers(self._h))

    def _ensure(self, w, h, b):
        W, H, B = self._dims
        if self._h is None or w > W or h > H or b > B:
            self._create(max(w, W, 1), max(h, H, 1), max(b, B, 1))

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().b200_aruco_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def detect(self, image, camera_params=None, marker_size=-1.0):
        """std::vector<aruco::Marker> detect(const cv::Mat&, const CameraParameters&, float markerSizeMeters)
        (markerdetector.h:276-278): markers sorted by id; with valid camera parameters and a positive marker size the
        extrinsics of every marker are filled in (markerdetector_impl.cpp:8772 -> marker.cpp:322-343)"""
        image = np.asarray(image)
        assert image.dtype == np.uint8 and image.ndim == 2
        m, c = self.detect_batch(image[None])
        out = [Marker(r["id"], r["xy"]) for r in m[0, :c[0]]]
        if camera_params is not None and camera_params.isValid() and marker_size > 0 and out:
            poses = self.estimate_poses(m[0, :c[0]], marker_size, camera_params.resized(image.shape[1], image.shape[0]))
            for mk, p in zip(out, poses):
                mk.Rvec, mk.Tvec, mk.Rvec2, mk.Tvec2 = p["rvec"].copy(), p["tvec"].copy(), p["rvec2"].copy(), p["tvec2"].copy()
                mk.err1, mk.err2, mk.ssize = float(p["err1"]), float(p["err2"]), float(marker_size)
        return out

    @staticmethod
    def rt_matrix(rvec, tvec):
        """getRTMatrix(Rvec, Tvec, CV_32F) (ippe.cpp:16-60): 4 x 4 [R | t] from a Rodrigues vector and a translation"""
        r = np.asarray(rvec, np.float64).reshape(3); th = float(np.sqrt((r * r).sum()))
        R = np.eye(3)
        if th > 1e-12:
            k = r / th
            K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            R = np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * K
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = R.astype(np.float32); T[:3, 3] = np.asarray(tvec, np.float32).reshape(3)
        return T

    def solvePnP(self, marker_size, img_points, camera_params):
        """aruco::solvePnP(objPoints, imgPoints, cameraMatrix, distCoeffs) (ippe.h:14-15, ippe.cpp:72-88) for the canonical marker square of side
        marker_size: [(T1, err1), (T2, err2)], the solution of smaller reprojection error first.  src/Frame.cc:155-177 calls it per marker with the
        ORIGINAL mK / mDistCoef (detect() used the camera resized to the image) and tests err1 / err2 < 0.7."""
        mk = np.zeros(1, MARKER_DTYPE)
        mk["xy"][0] = np.asarray(img_points, np.float32).reshape(8)
        p = self.estimate_poses(mk, marker_size, camera_params)[0]
        return [(self.rt_matrix(p["rvec"], p["tvec"]), float(p["err1"])), (self.rt_matrix(p["rvec2"], p["tvec2"]), float(p["err2"]))]

    def estimate_poses(self, markers, marker_size, camera_params):
        """both IPPE poses + reprojection errors of an array of markers (aruco::solvePnP, ippe.cpp:72-88; the ratio
        err1 / err2 < 0.7 is the reference's test for a good marker, src/Frame.cc:172-174)"""
        from ._lib import POSE_DTYPE
        markers = np.ascontiguousarray(markers)
        assert markers.dtype == MARKER_DTYPE
        poses = np.zeros(len(markers), POSE_DTYPE)
        cam = np.ascontiguousarray(camera_params.cam9(), np.float32)
        check(lib().b200_aruco_pose_host(ptr(markers), len(markers), float(marker_size), ptr(cam), ptr(poses), self._device))
        return poses

    def detect_batch(self, images):
        images = np.asarray(images)
        assert images.dtype == np.uint8 and images.ndim == 3
        if images.strides[2] != 1:
            images = np.ascontiguousarray(images)
        n, h, w = images.shape
        self._ensure(w, h, n)
        markers = np.zeros((n, self.cap), MARKER_DTYPE)
        counts = np.zeros(n, np.int32)
        if n:
            check(lib().b200_aruco_detect_host(self._h, ptr(images), n, w, h, images.strides[1], images.strides[0], ptr(markers), ptr(counts)))
        return markers, counts

    def debug(self, frame):
        """validation taps of the last call: (counts[4], kept corners (k,4,2), decoded ids (k,))"""
        out4 = np.zeros(4, np.int32)
        corners = np.zeros((256, 8), np.float32); ids = np.zeros(256, np.int32)
        check(lib().b200_aruco_debug(self._h, frame, ptr(out4), ptr(corners), ptr(ids), 256))
        k = int(out4[2])
        return out4, corners[:k].reshape(-1, 4, 2).copy(), ids[:k].copy()

    def detect_batch_device(self, images, markers, counts, stream=None):
        n, h, w = images.shape
        self._ensure(w, h, n)
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream
        check(lib().b200_aruco_detect(self._h, ptr(images), n, w, h, images.stride(1), images.stride(0), ptr(markers), ptr(counts), s))


class FrontEnd:
    """Extractor + marker detector + matcher on one upload: what Frame::Frame (reference src/Frame.cc:91,142) and
    TrackReferenceKeyFrame (src/Tracking.cc:917) do per image, for a batch of frames through b200_frontend_host."""

    def __init__(self, extractor, detector=None, matcher=None):
        self.extractor, self.detector, self.matcher = extractor, detector, matcher

    def alloc_outputs(self, n, pinned=False):
        """host output buffers for n frames (optionally pinned through torch) -> dict of numpy arrays"""
        if self.extractor._h is None:
            raise B200Error(_lib.EINVAL, "create the extractor with max_width/max_height/max_batch (or run one frame) first")
        cap = self.extractor.cap
        mcap = self.detector.cap if self.detector is not None and self.detector._h is not None else check(lib().b200_aruco_max_markers(None))

        def mk(shape, dtype):
            if pinned:
                import torch
                t = torch.zeros(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8).pin_memory()
                self._keep = getattr(self, "_keep", []) + [t]
                return t.numpy().view(dtype).reshape(shape)
            return np.zeros(shape, dtype)
        return dict(kps=mk((n, cap), KP_DTYPE), desc=mk((n, cap, 32), np.uint8), counts=mk((n,), np.int32),
                    markers=mk((n, mcap), MARKER_DTYPE), marker_counts=mk((n,), np.int32),
                    matches=mk((n, cap), np.int32), n_matches=mk((n,), np.int32))

    def process_batch(self, images, ref_desc=None, ref_kps=None, out=None):
        images = np.asarray(images)
        assert images.dtype == np.uint8 and images.ndim == 3
        if images.strides[2] != 1:
            images = np.ascontiguousarray(images)
        n, h, w = images.shape
        ex, det = self.extractor, self.detector
        # batches above 256 frames stream through two 128-frame scratch regions (b200_frontend_host): the handles never need more than 256 slots
        ex._ensure(w, h, min(n, 256))
        if det is not None:
            det._ensure(w, h, min(n, 256))
        if out is None:
            out = self.alloc_outputs(n)
        do_match = ref_desc is not None and self.matcher is not None
        if do_match:
            ref_desc = np.ascontiguousarray(ref_desc, np.uint8).reshape(-1, 32)
            ref_kps = np.ascontiguousarray(ref_kps)
            assert ref_kps.dtype == KP_DTYPE and len(ref_kps) == len(ref_desc)
        check(lib().b200_frontend_host(
            ex._h, det._h if det is not None else None, ptr(images), n, w, h, images.strides[1], images.strides[0],
            ptr(out["kps"]), ptr(out["desc"]), ptr(out["counts"]),
            ptr(out["markers"]) if det is not None else None, ptr(out["marker_counts"]) if det is not None else None,
            ptr(ref_desc) if do_match else None, ptr(ref_kps) if do_match else None, len(ref_desc) if do_match else 0,
            self.matcher.mfNNratio if do_match else 0.0, int(self.matcher.mbCheckOrientation) if do_match else 0,
            ptr(out["matches"]) if do_match else None, ptr(out["n_matches"]) if do_match else None))
        return out


    def collate_host(self, collator, root_out=None, root=0):
        """sharded front end: after process_batch on every rank, the device-resident results travel to `root` in one NCCL group
        (b200_frontend_collate_host) and land in root_out, host arrays laid out [world][n][...] (alloc_outputs(world * n) on the root)."""
        o = root_out
        check(lib().b200_frontend_collate_host(
            self.extractor._h, collator._h, int(root), ptr(o["kps"]) if o else None, ptr(o["desc"]) if o else None, ptr(o["counts"]) if o else None,
            ptr(o["markers"]) if o and self.detector is not None else None, ptr(o["marker_counts"]) if o and self.detector is not None else None,
            ptr(o["matches"]) if o and self.matcher is not None else None, ptr(o["n_matches"]) if o and self.matcher is not None else None))
        return o


class FrameGrid:
    """Frame::UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid / GetFeaturesInArea (reference src/Frame.cc:357-388,
    418-447, 183-198, 280-330) on device buffers (torch tensors or raw device addresses): the step between the extractor and
    every SearchByProjection, so that candidate lists for b200_match_candidates are built without a host round trip."""
    COLS, ROWS = 64, 48

    def __init__(self, width, height, camera_params, device=0):
        self._device = int(device)
        self.cam9 = np.ascontiguousarray(camera_params.cam9(), np.float32)
        self.bounds = np.zeros(4, np.float32)                 # mnMinX mnMaxX mnMinY mnMaxY
        check(lib().b200_frame_image_bounds(int(width), int(height), ptr(self.cam9), ptr(self.bounds), self._device))

    def undistort(self, kps, counts, kps_un, stream=None):
        """kps, kps_un: [B][cap] keypoint records on the device; counts [B] int32 on the device"""
        B, cap = int(kps.shape[0]), int(kps.shape[1])
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream
        check(lib().b200_frame_undistort(ptr(kps), ptr(counts), B, cap, ptr(self.cam9), ptr(kps_un), self._device, s))

    def undistort_aruco_corners(self, markers):
        """Frame::UndistortArucoCorners (src/Frame.cc:388-416): the 4 * NA corners of a frame's markers (MARKER_DTYPE records or Marker objects) through
        cv::undistortPoints(..., mK, mDistCoef, cv::Mat(), mK) on the device -> mvArucoUn as a [4 * NA, 2] float32 array.  With k1 == 0 the reference
        returns before touching mvArucoUn; here the corners come back unchanged."""
        if len(markers) and not isinstance(markers, np.ndarray):
            xy = np.array([m.corners for m in markers], np.float32).reshape(-1, 2)
        else:
            xy = np.ascontiguousarray(np.asarray(markers)["xy"], np.float32).reshape(-1, 2) if len(markers) else np.zeros((0, 2), np.float32)
        xy = np.ascontiguousarray(xy)
        out = np.empty_like(xy)
        check(lib().b200_frame_undistort_points_host(ptr(xy), len(xy), ptr(self.cam9), ptr(out), self._device))
        return out

    def assign(self, kps_un, counts, cell_start, cell_items, stream=None):
        """cell_start [B][64*48+1], cell_items [B][cap] int32 on the device"""
        B, cap = int(kps_un.shape[0]), int(kps_un.shape[1])
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream
        check(lib().b200_frame_assign_grid(ptr(kps_un), ptr(counts), B, cap, ptr(self.bounds), ptr(cell_start), ptr(cell_items), self._device, s))

    def features_in_area(self, kps_un_f, cell_start_f, cell_items_f, queries_xyr, query_levels, out_idx, out_count, stream=None):
        """one frame's rows; queries_xyr [n][3] float32, query_levels [n][2] int32, out_idx [n][row_cap], out_count [n] (device)"""
        n, row_cap = int(queries_xyr.shape[0]), int(out_idx.shape[1])
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream
        check(lib().b200_frame_features_in_area(ptr(kps_un_f), ptr(cell_start_f), ptr(cell_items_f), ptr(self.bounds), ptr(queries_xyr),
                                                ptr(query_levels), n, ptr(out_idx), ptr(out_count), row_cap, self._device, s))


class ORBVocabulary:
    """DBoW2 vocabulary tree (ORB_SLAM2::ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>, reference
    include/ORBVocabulary.h:31): loadFromTextFile (TemplatedVocabulary.h:1338-1425) and transform (:1127-1259) for the configuration
    of ORBvoc.txt (TF_IDF weighting, L1 scoring).  The per-descriptor descent runs on the device; BowVector / FeatureVector (host
    std::maps in the reference) are assembled from its three output arrays."""

    def __init__(self, k, L, parent, is_leaf, node_desc, node_weight, device=0):
        self.k, self.L = int(k), int(L)
        self.parent = np.ascontiguousarray(parent, np.int32)
        self.is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        self.node_desc = np.ascontiguousarray(node_desc, np.uint8).reshape(-1, 32)
        self.node_weight = np.ascontiguousarray(node_weight, np.float64)
        assert len(self.parent) == len(self.is_leaf) == len(self.node_desc) == len(self.node_weight)
        self._device = int(device)
        h = C.c_void_p()
        check(lib().b200_voc_create(C.byref(h), self.k, self.L, len(self.parent), ptr(self.parent), ptr(self.is_leaf), ptr(self.node_desc),
                                    ptr(self.node_weight), self._device))
        self._h = h

    @classmethod
    def loadFromTextFile(cls, path, device=0):
        """the text format of ORBvoc.txt: 'k L scoring weighting', then one node per line: parent isLeaf d0 .. d31 weight"""
        with open(path) as fh:
            k, L, n1, n2 = [int(v) for v in fh.readline().split()[:4]]
            if k < 0 or k > 20 or L < 1 or L > 10 or n1 < 0 or n1 > 5 or n2 < 0 or n2 > 3:
                raise ValueError("Vocabulary loading failure: This is not a correct text file!")
            rows = np.loadtxt(fh, dtype=np.float64, ndmin=2)
        parent = np.concatenate([[0], rows[:, 0]]).astype(np.int32)
        leaf = np.concatenate([[0], rows[:, 1] > 0]).astype(np.uint8)
        desc = np.concatenate([np.zeros((1, 32)), rows[:, 2:34]]).astype(np.uint8)
        weight = np.concatenate([[0.0], rows[:, 34]])
        return cls(k, L, parent, leaf, desc, weight, device)

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().b200_voc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return check(lib().b200_voc_num_words(self._h))

    def descend(self, descriptors, levelsup=4):
        """(word_id, weight, node_id) per descriptor: transform(feature, id, weight, &nid, levelsup)"""
        d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        w = np.zeros(len(d), np.int32); wt = np.zeros(len(d), np.float64); nid = np.zeros(len(d), np.int32)
        check(lib().b200_voc_transform_host(self._h, ptr(d), len(d), int(levelsup), ptr(w), ptr(wt), ptr(nid)))
        return w, wt, nid

    def transform(self, descriptors, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup) (Frame::ComputeBoW, src/Frame.cc:348-355) ->
        (BowVector as {word: value}, FeatureVector as {node: [feature indices]}), both in ascending key order like the std::maps"""
        return self.vectors(*self.descend(descriptors, levelsup))

    @staticmethod
    def vectors(w, wt, nid):
        """BowVector::addWeight / FeatureVector::addFeature / normalize(L1) over per-feature (word, weight, node) (TemplatedVocabulary.h:1159-1190)"""
        bow, fv = {}, {}
        for i in range(len(w)):
            if wt[i] > 0:                                   # not stopped
                bow[int(w[i])] = bow.get(int(w[i]), 0.0) + float(wt[i])
                fv.setdefault(int(nid[i]), []).append(i)
        norm = sum(abs(v) for _, v in sorted(bow.items()))  # BowVector::normalize(L1), summed in key order
        if norm > 0.0:
            bow = {k2: v / norm for k2, v in bow.items()}
        return dict(sorted(bow.items())), dict(sorted(fv.items()))
