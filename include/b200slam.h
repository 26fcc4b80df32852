/* b200slam.h -- C-ABI of the B200-native ORB-SLAM2/ArUco per-frame front end.
 *
 * The reference (CarminLiu/ORB_SLAM2_aruco) has no FFI layer: Frame/Tracking call three C++ classes
 * directly.  This header is the boundary a drop-in replacement exports underneath those classes
 * (SURVEY.md section 8(b)); include/b200slam_adapters.hpp holds the source-compatible C++ classes
 * (same names / signatures as the reference) that marshal to these entry points.
 *
 *   entry point                replaces (reference file:line)
 *   -------------------------  ---------------------------------------------------------------------
 *   b200_orb_create            ORB_SLAM2::ORBextractor::ORBextractor        include/ORBextractor.h:51-52, src/ORBextractor.cc:410-470
 *   b200_orb_extract[_host]    ORB_SLAM2::ORBextractor::operator()          include/ORBextractor.h:59-61, src/ORBextractor.cc:1043-1105
 *   b200_orb_get_level_info    GetLevels/GetScaleFactors/...                include/ORBextractor.h:63-83
 *   b200_orb_get_pyramid       public member mvImagePyramid                 include/ORBextractor.h:85,  src/ORBextractor.cc:1107-1132
 *   b200_match_bf[_host]       ORB_SLAM2::ORBmatcher::SearchByBoW(KF,F,..)  include/ORBmatcher.h:55,    src/ORBmatcher.cc:159-292
 *                              (degenerate single-node FeatureVector == brute force; DescriptorDistance 1651-1667)
 *   b200_match_candidates      candidate-list core of SearchByProjection/SearchForInitialization
 *                                                                           src/ORBmatcher.cc:45-129,409-524,1332-1474
 *   b200_aruco_create          aruco::MarkerDetector + setDictionary/setDetectionMode/setCornerRefinementMethod
 *                                                                           Thirdparty/aruco/aruco/markerdetector.h:258,337, src/Frame.cc:129-139
 *   b200_aruco_detect[_host]   aruco::MarkerDetector::detect                Thirdparty/aruco/aruco/markerdetector.h:276-278, markerdetector_impl.cpp:5826
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on success or a
 * negative B200_E* code and never throws; handles are opaque; output buffers are caller-allocated;
 * `stream` is a cudaStream_t passed as void* (NULL = the handle's own stream).  Functions without the
 * _host suffix take DEVICE pointers and only enqueue work on `stream`; the _host variants take HOST
 * pointers, stage through pinned memory and return after the results are in the caller's buffers.
 * There is no CPU fallback: without a usable CUDA device every call fails with B200_ENODEV.
 */
#ifndef B200SLAM_H
#define B200SLAM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK          0
#define B200_EINVAL     -1   /* bad argument (NULL pointer, non-positive size, unknown dictionary ...)          */
#define B200_ENODEV     -2   /* no CUDA device / wrong architecture                                            */
#define B200_ECUDA      -3   /* a CUDA runtime call failed; see b200_last_error()                               */
#define B200_ECAPACITY  -4   /* a batch / image larger than the handle was created for, or scratch overflow     */
#define B200_ENOMEM     -5

/* 28 bytes, identical to cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id). */
typedef struct b200_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} b200_keypoint;

/* One detected marker: id + 4 refined corners (x0,y0,...,x3,y3), clockwise from the marker's canonical top-left. */
typedef struct b200_marker {
    int32_t id;
    float xy[8];
} b200_marker;

/* Marker pose (aruco::Marker::Rvec / Tvec, Thirdparty/aruco/aruco/marker.h:47-59): both IPPE solutions, the one with the
 * smaller reprojection error first, and both errors (their ratio is the quality test of src/Frame.cc:155-177). */
typedef struct b200_marker_pose {
    float rvec[3], tvec[3], err1;
    float rvec2[3], tvec2[3], err2;
} b200_marker_pose;

typedef struct b200_orb_s*   b200_orb_t;
typedef struct b200_aruco_s* b200_aruco_t;
typedef struct b200_voc_s*   b200_voc_t;
typedef struct b200_collate_s* b200_collate_t;

const char* b200_last_error(void);
/* number of CUDA kernel launches issued by this library in the calling process so far */
int64_t b200_launch_count(void);

/* ---------------------------------------------------------------- extractor ---------------------- */
int b200_orb_create(b200_orb_t* out, int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                    int max_width, int max_height, int max_batch, int device);
int b200_orb_destroy(b200_orb_t h);
/* capacity (keypoints per frame) the output buffers must provide: sum over levels of quota + slack */
int b200_orb_max_keypoints(b200_orb_t h);
/* per-level tables as the reference getters return them; any pointer may be NULL. Arrays hold nlevels entries. */
int b200_orb_get_level_info(b200_orb_t h, int* nlevels, float* scale_factors, float* inv_scale_factors,
                            float* level_sigma2, float* inv_level_sigma2, int32_t* features_per_level);
/* Device-pointer call.  imgs: n gray u8 frames, frame f row y at imgs + f*frame_stride + y*row_stride.
 * kps [n][cap], desc [n][cap][32], counts [n] with cap = b200_orb_max_keypoints(). */
int b200_orb_extract(b200_orb_t h, const uint8_t* imgs, int n, int width, int height, int64_t row_stride, int64_t frame_stride,
                     b200_keypoint* kps, uint8_t* desc, int32_t* counts, void* stream);
/* Host-pointer call (the reference-facing one): same layout, host memory, synchronous. */
int b200_orb_extract_host(b200_orb_t h, const uint8_t* imgs, int n, int width, int height, int64_t row_stride, int64_t frame_stride,
                          b200_keypoint* kps, uint8_t* desc, int32_t* counts);
/* The whole per-frame front end on host buffers with ONE upload: extractor + (optional, aruco != NULL) marker
 * detector + (optional, ref_desc != NULL) brute-force SearchByBoW against a reference set, i.e. what
 * Frame::Frame (src/Frame.cc:91,142) and TrackReferenceKeyFrame (src/Tracking.cc:917) do on one image.
 * Frames are uploaded in chunks that overlap with compute; detector and extractor run on separate streams.
 * match_ref_idx [n][cap] / n_matches [n] as in b200_match_bf with th_low = 50 and the 30/360 histogram factor. */
int b200_frontend_host(b200_orb_t orb, b200_aruco_t aruco, const uint8_t* imgs, int n, int width, int height,
                       int64_t row_stride, int64_t frame_stride,
                       b200_keypoint* kps, uint8_t* desc, int32_t* counts,
                       b200_marker* markers, int32_t* marker_counts,
                       const uint8_t* ref_desc, const b200_keypoint* ref_kps, int n_ref, float ratio, int check_ori,
                       int32_t* match_ref_idx, int32_t* n_matches);
/* Per-stage device timing for roofline reports: when enabled, every extract call records CUDA events on its
 * stream around the four stages; b200_orb_get_stage_ms returns the last call's pyramid / fast / quadtree /
 * describe durations in milliseconds (ms4[4]). */
int b200_orb_set_profile(b200_orb_t h, int enable);
int b200_orb_get_stage_ms(b200_orb_t h, float* ms4);
/* number of frames the profiled launches of the last call processed */
int b200_orb_get_stage_frames(b200_orb_t h);
/* Pyramid of frame `frame` of the LAST extract call, level `level`, with the reference's 19-px REFLECT_101
 * border: out (host) receives (w_l+38) x (h_l+38) bytes, inner size returned in *w_l, *h_l. */
int b200_orb_get_pyramid(b200_orb_t h, int frame, int level, uint8_t* out, int* w_l, int* h_l);
/* Debug/validation taps of the LAST call (host buffers): FAST candidates of one level before the quadtree,
 * xys [cap][3] = x, y (relative to the 16-px border, as in the reference) and score.  Returns count or <0. */
int b200_orb_get_candidates(b200_orb_t h, int frame, int level, int32_t* xys, int cap);

/* ---------------------------------------------------------------- matcher ------------------------ */
/* Brute-force SearchByBoW(KeyFrame, Frame) semantics with one all-inclusive vocabulary node
 * (src/ORBmatcher.cc:159-292): the outer, order-dependent loop runs over the reference set in index order,
 * the inner search over the frame's descriptors that are not taken yet; accept when best <= th_low and
 * best < ratio*second; then keep the three dominant bins of the 30-bin rotation histogram if check_ori.
 *   ref_desc [n_ref][32], ref_angle [n_ref]           (replicated for every frame of the batch)
 *   frame_desc [n_batch][frame_cap][32], frame_angle [n_batch][frame_cap], n_frame [n_batch]
 *   match_ref_idx [n_batch][frame_cap]: reference index matched to each frame keypoint or -1
 *   n_matches [n_batch]
 * histo_factor: the reference uses 30/360 here (ORBmatcher.cc:176) and 1/30 elsewhere (417,545,1340). */
int b200_match_bf(const uint8_t* ref_desc, const float* ref_angle, int n_ref,
                  const uint8_t* frame_desc, const float* frame_angle, const int32_t* n_frame, int n_batch, int frame_cap,
                  float ratio, int th_low, int check_ori, float histo_factor,
                  int32_t* match_ref_idx, int32_t* n_matches, int device, void* stream);
/* Device-pointer variant that reads the angles out of b200_keypoint records, so the extractor's output buffers
 * (kps [n][cap], desc [n][cap][32], counts [n]) feed the matcher without a repack. */
int b200_match_bf_kp(const uint8_t* ref_desc, const b200_keypoint* ref_kps, int n_ref,
                     const uint8_t* frame_desc, const b200_keypoint* frame_kps, const int32_t* n_frame, int n_batch, int frame_cap,
                     float ratio, int th_low, int check_ori, float histo_factor,
                     int32_t* match_ref_idx, int32_t* n_matches, int device, void* stream);
int b200_match_bf_host(const uint8_t* ref_desc, const float* ref_angle, int n_ref,
                       const uint8_t* frame_desc, const float* frame_angle, const int32_t* n_frame, int n_batch, int frame_cap,
                       float ratio, int th_low, int check_ori, float histo_factor,
                       int32_t* match_ref_idx, int32_t* n_matches, int device);
/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:409-524) on arrays:
 * undistorted keypoints + descriptors of both frames (HOST), bounds4 = mnMinX mnMaxX mnMinY mnMaxY of the frames
 * (b200_frame_image_bounds), prev_matched [n1][2] in/out (vbPrevMatched), matches12 [n1] out (vnMatches12; -1 = none).
 * Candidates come from the device feature grid of F2 (GetFeaturesInArea at level 0), the sequential accept / steal / rotation
 * histogram logic is replayed in order.  Returns nmatches (>= 0) or a negative error. */
int b200_match_for_initialization_host(const b200_keypoint* kps1_un, const uint8_t* desc1, int n1,
                                       const b200_keypoint* kps2_un, const uint8_t* desc2, int n2, const float* bounds4,
                                       float* prev_matched, int window, float ratio, int check_ori, int32_t* matches12, int device);
/* ORBmatcher::SearchByBoW over the two FeatureVectors (b200_voc_transform), merge-walked on the host into n_groups common vocabulary nodes:
 *   mode 0  SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)        (include/ORBmatcher.h:55, src/ORBmatcher.cc:159-292)   out [n_c]: out[idxF] = idxKF
 *   mode 1  SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)     (include/ORBmatcher.h:56, src/ORBmatcher.cc:526-659)   out [n_q]: out[idx1] = idx2
 * Group g: query features q_idx[grp_q_ofs[g] .. grp_q_ofs[g+1]) (node order, only those with a good MapPoint) against candidate features
 * c_idx[grp_c_ofs[g] .. grp_c_ofs[g+1]) (mode 1: only those with a good MapPoint).  HOST pointers.  th_low <= 0: TH_LOW = 50.  Returns nmatches. */
int b200_match_by_bow_host(const uint8_t* q_desc, const float* q_angle, int n_q, const uint8_t* c_desc, const float* c_angle, int n_c,
                           const int32_t* grp_q_ofs, const int32_t* q_idx, const int32_t* grp_c_ofs, const int32_t* c_idx, int n_groups,
                           int mode, float ratio, int th_low, int check_ori, int32_t* out, int device);
/* MapPoint::ComputeDistinctiveDescriptors (include/MapPoint.h:75, src/MapPoint.cc:271-331) for n_points map points at once.  desc [ofs[n_points]][32]:
 * the observing keyframes' descriptor rows (bad keyframes already dropped), map point p owning rows ofs[p] .. ofs[p+1]).  best_idx [n_points] out: the
 * row (relative to ofs[p]) with the least median distance to the rest, first minimum; -1 for a point without observations (the reference returns
 * without touching mDescriptor).  out_desc [n_points][32] (may be NULL): the chosen descriptor (zeros where best_idx = -1).  HOST pointers. */
int b200_distinctive_descriptors_host(const uint8_t* desc, const int32_t* ofs, int n_points, int32_t* best_idx, uint8_t* out_desc, int device);
/* ORBmatcher::SearchByProjection on ready-made projections (the projection of the map points is host glue in the reference):
 *   mode 0  SearchByProjection(Frame&, const vector<MapPoint*>&, th)          (src/ORBmatcher.cc:45-129, Tracking::SearchLocalPoints)
 *   mode 1  SearchByProjection(Frame& Current, const Frame& Last, th, mono)    (src/ORBmatcher.cc:1332-1474, TrackWithMotionModel); with
 *           th_high = ORBdist also SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1476-1603, Relocalization)
 * Frame side (HOST): undistorted keypoints, descriptors, bounds4, occupied [n_frame] in/out = "mvpMapPoints[i] && Observations() > 0".
 * Query q (HOST): q_xyr [n][3] = projected x, y and the search radius (already scaled), q_levels [n][2] = minLevel maxLevel of
 * GetFeaturesInArea, q_desc [n][32], q_angle [n] (mode 1 histogram), q_observed [n] = the map point's Observations() > 0.
 * th_high: accept bestDist <= th_high (<= 0: TH_HIGH = 100).
 * The loop-closing SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (include/ORBmatcher.h:52, src/ORBmatcher.cc:294-407) is mode 1 with
 * check_ori = 0, th_high = TH_LOW (50), q_levels = (nPredictedLevel - 1, nPredictedLevel), occupied = "vpMatched[i] != NULL", q_observed = 1;
 * pass mode 2 for it: mode 1 with the keyframe's grid origin (b200_keyframe_features_in_area).
 * assign [n_frame] out = query index assigned to that frame keypoint (F.mvpMapPoints[bestIdx] = pMP), -1 = untouched, -2 = assigned during the call and
 * then cleared by the rotation histogram (the reference stores NULL there, src/ORBmatcher.cc:1459-1467).  Returns nmatches. */
int b200_match_by_projection_host(const b200_keypoint* kps_un, const uint8_t* desc, int n_frame, const float* bounds4, uint8_t* occupied,
                                  const float* q_xyr, const int32_t* q_levels, const uint8_t* q_desc, const float* q_angle, const uint8_t* q_observed,
                                  int n_queries, int mode, float ratio, int check_ori, int th_high, int32_t* assign, int device);
/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo = false) (include/ORBmatcher.h:66-67, src/ORBmatcher.cc:661-829)
 * for monocular keyframes.  The host merge-walks the two FeatureVectors into n_groups common vocabulary nodes exactly as for
 * b200_match_by_bow_host, but keeps only the features WITHOUT a MapPoint on both sides (q_idx: KF1, c_idx: KF2).  F12 row-major 3x3,
 * epipole2 = (ex, ey) of camera 1 in image 2 (:668-674, host glue), scale_factors / level_sigma2 [nlevels] of KF2 (mvScaleFactors, mvLevelSigma2).
 * matches12 [n1] out: the KF2 feature paired with KF1 feature i or -1; the pairs in ascending i are vMatchedPairs.  th_low <= 0: TH_LOW = 50.
 * HOST pointers.  Returns nmatches. */
int b200_match_for_triangulation_host(const b200_keypoint* kps1_un, const uint8_t* desc1, int n1, const b200_keypoint* kps2_un, const uint8_t* desc2, int n2,
                                      const int32_t* grp_q_ofs, const int32_t* q_idx, const int32_t* grp_c_ofs, const int32_t* c_idx, int n_groups,
                                      const float* F12, const float* epipole2, const float* scale_factors, const float* level_sigma2, int nlevels,
                                      int check_ori, int th_low, int32_t* matches12, int device);
/* The keyframe search shared by ORBmatcher::Fuse (include/ORBmatcher.h:76, src/ORBmatcher.cc:831-981), Fuse(pKF, Scw, ...) (include/ORBmatcher.h:79,
 * src/ORBmatcher.cc:983-1104) and SearchBySim3 (include/ORBmatcher.h:70-71, src/ORBmatcher.cc:1106-1330): for every projected map point the most
 * similar keyframe feature inside its window.  The search does not depend on what earlier points did to the keyframe, so all queries run at once;
 * projecting the points (pose, depth, viewing-angle and scale tests, PredictScale) and applying the outcome to MapPoint / KeyFrame objects stay
 * host glue.  Query q: q_xyr = (u, v, radius), q_level = nPredictedLevel (candidates are limited to levels [q_level - 1, q_level]), q_desc.
 * chi2 > 0: Fuse's reprojection gate, a candidate is skipped when ((u - x)^2 + (v - y)^2) * inv_level_sigma2[octave] > chi2 (5.99 monocular).
 * best_idx / best_dist [n_queries] out: feature index and Hamming distance of the first minimum in GetFeaturesInArea order (-1 / 256 when no
 * candidate qualifies); the caller applies bestDist <= TH_LOW (Fuse) or TH_HIGH (SearchBySim3).  HOST pointers. */
int b200_match_kf_radius_host(const b200_keypoint* kps_un, const uint8_t* desc, int n_kf, const float* bounds4, const float* q_xyr, const int32_t* q_level,
                              const uint8_t* q_desc, int n_queries, const float* inv_level_sigma2, int nlevels, double chi2, int32_t* best_idx,
                              int32_t* best_dist, int device);
/* The projection front of those members and of SearchByProjection(pKF, Scw, ...) for n map points at once (src/ORBmatcher.cc:318-366, 848-898,
 * 1006-1061; per direction of SearchBySim3 :1158-1195 / :1238-1275), in the reference's cv::Mat CV_32F arithmetic.  Rcw / tcw / Ow: rotation (row-major
 * 3x3), translation and camera centre of the keyframe (for the Scw overloads the decomposed Sim3, :302-307).  sR / tt non-NULL: SearchBySim3 - the point
 * goes world -> camera a (Rcw, tcw) -> camera b = sR * p + tt, its distance is |p_b| and there is no viewing-angle test (Ow, normal unused).
 * pos [n][3], normal [n][3] (NULL: no viewing-angle test), minmax [n][2] = mfMinDistance, mfMaxDistance.  scale_factors [nlevels];
 * level_thresholds [nlevels - 1]: the largest float ratio mfMaxDistance / dist for which MapPoint::PredictScale still answers level k, found by the
 * caller with its own libm (bisection over float bit patterns), so that the device reproduces the host's ceil(logf(ratio) / logf(scaleFactor)).
 * valid [n] out: 0 when one of the reference's tests discards the point; q_xyr [n][3] = (u, v, th * scale_factors[level]); level [n] = nPredictedLevel.
 * HOST pointers. */
int b200_kf_project_host(const float* Rcw, const float* tcw, const float* Ow, const float* sR, const float* tt, const float* cam4, const float* bounds4,
                         const float* pos, const float* normal, const float* minmax, int n, float th, const float* scale_factors,
                         const float* level_thresholds, int nlevels, uint8_t* valid, float* q_xyr, int32_t* level, int device);
/* Projection and search in ONE call (k_kf_project -> k_features_in_area -> k_radius_best without a host round trip): what Fuse, Fuse(Scw) and each direction
 * of SearchBySim3 do per map point up to "if(bestDist<=TH_...)".  Arguments as b200_kf_project_host and b200_match_kf_radius_host; q_desc [n][32];
 * skip [n] (may be NULL): 1 = the reference `continue`s on this point before projecting it (NULL, bad, already in the keyframe / already found).
 * valid [n] out: the point reached GetFeaturesInArea; best_idx / best_dist [n] out (-1 / 256 for discarded points and empty windows);
 * q_xyr [n][3] and level [n] out (may be NULL; discarded points report (0, 0, -1)).  HOST pointers. */
int b200_kf_search_points_host(const b200_keypoint* kps_un, const uint8_t* desc, int n_kf, const float* bounds4, const float* Rcw, const float* tcw, const float* Ow,
                               const float* sR, const float* tt, const float* cam4, const float* pos, const float* normal, const float* minmax, const uint8_t* q_desc,
                               const uint8_t* skip, int n, float th, const float* scale_factors, const float* inv_level_sigma2, const float* level_thresholds,
                               int nlevels, double chi2, uint8_t* valid, int32_t* best_idx, int32_t* best_dist, float* q_xyr, int32_t* level, int device);
/* Plain 256-bit Hamming distance matrix rows x cols (ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1651-1667). */
int b200_hamming_matrix_host(const uint8_t* a, int na, const uint8_t* b, int nb, int32_t* dist, int device);
/* Candidate-list matching core shared by SearchByProjection / SearchForInitialization:
 * query q (in index order) examines train descriptors cand[cand_ofs[q] .. cand_ofs[q+1]); best/second
 * Hamming distance and the index of the best.  Host builds the lists, the device does distances.
 *   out_best_idx [nq], out_best_dist [nq], out_second_dist [nq]  (idx -1 / dist 256 when the list is empty) */
int b200_match_candidates_host(const uint8_t* query_desc, int nq, const uint8_t* train_desc, int nt,
                               const int32_t* cand_ofs, const int32_t* cand, int32_t* out_best_idx,
                               int32_t* out_best_dist, int32_t* out_second_dist, int device);

/* ---------------------------------------------------------------- ArUco detector ----------------- */
/* dict_name: "ARUCO", "ARUCO_MIP_25h7", "ARUCO_MIP_36h12", ... (Thirdparty/aruco/aruco/dictionary.cpp:367-383).
 * Fixed configuration == what src/Frame.cc:129-139 sets: DM_NORMAL (adaptive threshold, ThresHold 7,
 * window = max(3, 15*w/1920) made odd), CORNER_LINES refinement, minSize 0, no error correction. */
int b200_aruco_create(b200_aruco_t* out, const char* dict_name, int max_width, int max_height, int max_batch, int device);
int b200_aruco_destroy(b200_aruco_t h);
/* marker capacity per frame of every marker buffer (a library constant; h may be NULL) */
int b200_aruco_max_markers(b200_aruco_t h);
/* markers [n][cap] sorted by id per frame (cap = b200_aruco_max_markers()), counts [n]. */
int b200_aruco_detect(b200_aruco_t h, const uint8_t* imgs, int n, int width, int height, int64_t row_stride, int64_t frame_stride,
                      b200_marker* markers, int32_t* counts, void* stream);
/* Waits for the handle's work on `stream` (NULL = own stream) and reports a scratch overflow of the device-pointer
 * calls since the last check as B200_ECAPACITY. */
int b200_aruco_check(b200_aruco_t h, void* stream);
/* Validation taps of the LAST call for one frame: out4 = {borders > 70 points, convex quads, candidates after
 * prefilterCandidates, decoded markers before de-duplication}; corners [cap][8] and ids [cap] (may be NULL) receive the
 * prefiltered candidates in order with their decoded id (-1 = not a marker). */
int b200_aruco_debug(b200_aruco_t h, int frame, int32_t* out4, float* corners, int32_t* ids, int cap);
int b200_aruco_detect_host(b200_aruco_t h, const uint8_t* imgs, int n, int width, int height, int64_t row_stride, int64_t frame_stride,
                           b200_marker* markers, int32_t* counts);
/* aruco::Marker::contourPoints (Thirdparty/aruco/aruco/marker.h:59; copied from the candidate at markerdetector_impl.cpp:6759-6772): the border of output
 * marker `index` of frame `frame` of the LAST detect call on the handle, in the reference's point order.  xy [cap][2] int32 (HOST); returns the number
 * of points of the border (the array receives min(cap, that) of them). */
int b200_aruco_get_contour(b200_aruco_t h, int frame, int index, int32_t* xy, int cap);
/* The borders of the first n_markers output markers of that frame in ONE round trip: ofs [n_markers + 1] (HOST), marker m owns points
 * xy [ofs[m] .. ofs[m + 1]) of xy [xy_cap][2]; returns the total number of points (pass xy_cap = 0 to ask for it first). */
int b200_aruco_get_contours(b200_aruco_t h, int frame, int n_markers, int32_t* ofs, int32_t* xy, int xy_cap);


/* ---------------------------------------------------------------- marker pose -------------------- */
/* aruco::Marker::calculateExtrinsics(markerSizeMeters, CameraMatrix, Distorsion) (Thirdparty/aruco/aruco/marker.cpp:322-343)
 * == aruco::solvePnP -> IPPE::PoseSolver::solveGeneric (Thirdparty/aruco/aruco/ippe.cpp:72-169) for every marker of a
 * batch: markers [n_batch][marker_cap] and counts [n_batch] as written by b200_aruco_detect (device pointers),
 * cam9 = fx fy cx cy k1 k2 p1 p2 k3 (HOST pointer: the reference's float camera matrix and distortion vector),
 * poses [n_batch][marker_cap] (device).  marker_size <= 0 or fx/fy == 0 -> B200_EINVAL (marker.cpp:309-333 throws). */
int b200_aruco_pose(const b200_marker* markers, const int32_t* counts, int n_batch, int marker_cap, float marker_size,
                    const float* cam9, b200_marker_pose* poses, int device, void* stream);
/* Same for a host array of n_markers markers (the adapter's detect(image, cameraParams, markerSize) and aruco::solvePnP). */
int b200_aruco_pose_host(const b200_marker* markers, int n_markers, float marker_size, const float* cam9,
                         b200_marker_pose* poses, int device);
/* One frame, everything aruco::MarkerDetector::detect(image, camParams, markerSizeMeters) returns (markerdetector.h:277-278, src/Frame.cc:142), with a
 * single synchronisation: markers [b200_aruco_max_markers()] sorted by id and their count; with cam9 != NULL the IPPE poses [b200_aruco_max_markers()]
 * of marker.cpp:322-343 computed on the device from the detector's device output; with contour_ofs != NULL ([count + 1] offsets) the markers' contour
 * points (aruco::Marker::contourPoints, markerdetector_impl.cpp:6759-6772) as x, y pairs in contour_xy [contour_cap pairs].  Returns the total number of
 * contour points (0 without contours; when it exceeds contour_cap only contour_cap points were copied: grow and call b200_aruco_get_contours), or a
 * negative error code.  The handle needs max_batch >= 1. */
int b200_aruco_detect_frame_host(b200_aruco_t h, const uint8_t* img, int width, int height, int64_t row_stride, b200_marker* markers, int32_t* count,
                                 float marker_size, const float* cam9, b200_marker_pose* poses,
                                 int32_t* contour_ofs, int32_t* contour_xy, int contour_cap);


/* ---------------------------------------------------------------- frame grid -------- */
/* Frame::UndistortKeyPoints (src/Frame.cc:357-388): kps / kps_un [n_batch][cap] device records, counts [n_batch] device,
 * cam9 = fx fy cx cy k1 k2 p1 p2 k3 (HOST).  k1 == 0 copies the keypoints like the reference. */
int b200_frame_undistort(const b200_keypoint* kps, const int32_t* counts, int n_batch, int cap, const float* cam9,
                         b200_keypoint* kps_un, int device, void* stream);
/* Frame::UndistortArucoCorners (src/Frame.cc:388-416): cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) over the 4 * NA marker corners of
 * a frame - the same kernel as the keypoints.  xy / xy_un [n][2] HOST (may alias).  k1 == 0: the points are copied and no device is needed (the
 * reference returns early and leaves mvArucoUn alone). */
int b200_frame_undistort_points_host(const float* xy, int n, const float* cam9, float* xy_un, int device);
/* Frame::ComputeImageBounds (src/Frame.cc:418-447): bounds4 = mnMinX mnMaxX mnMinY mnMaxY (HOST output). */
int b200_frame_image_bounds(int width, int height, const float* cam9, float* bounds4, int device);
/* Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:183-198, 332-343): the 64 x 48 grid of every frame as a CSR list,
 * cell = ix * 48 + iy (mGrid[ix][iy]); cell_start [n_batch][64*48 + 1], cell_items [n_batch][cap] (keypoint indices in push order). */
int b200_frame_assign_grid(const b200_keypoint* kps_un, const int32_t* counts, int n_batch, int cap, const float* bounds4,
                           int32_t* cell_start, int32_t* cell_items, int device, void* stream);
/* Frame::GetFeaturesInArea (src/Frame.cc:280-330) for n_queries windows of ONE frame (its kps_un / cell_start / cell_items rows):
 * queries_xyr [n][3] = x y r, query_levels [n][2] = minLevel maxLevel; out_idx [n][row_cap] in the reference's visit order,
 * out_count [n] (may exceed row_cap: then the row is truncated). */
int b200_frame_features_in_area(const b200_keypoint* kps_un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                                const float* queries_xyr, const int32_t* query_levels, int n_queries,
                                int32_t* out_idx, int32_t* out_count, int row_cap, int device, void* stream);
/* KeyFrame::GetFeaturesInArea (include/KeyFrame.h:106, src/KeyFrame.cc:672-718): the same search, but the window is laid over the grid from the
 * keyframe's INT mnMinX / mnMinY (truncated copies of the frame's float bounds, include/KeyFrame.h:211-214) with the frame's float cell size.
 * Identical to b200_frame_features_in_area when the bounds are integers (no distortion).  query_levels as there ((-1, -1) = no level filter). */
int b200_keyframe_features_in_area(const b200_keypoint* kps_un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                                   const float* queries_xyr, const int32_t* query_levels, int n_queries,
                                   int32_t* out_idx, int32_t* out_count, int row_cap, int device, void* stream);


/* ---------------------------------------------------------------- bag of words ------- */
/* A DBoW2 vocabulary tree in the node order of TemplatedVocabulary::loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1425):
 * node 0 is the root, nodes 1 .. n_nodes-1 follow in file order with parent [i] (< i), is_leaf [i], node_desc [i][32] and node_weight [i]
 * (entries 0 are ignored); children keep file order and word ids are assigned in order of leaf appearance, like the reference. */
int b200_voc_create(b200_voc_t* out, int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* node_desc,
                    const double* node_weight, int device);
int b200_voc_destroy(b200_voc_t h);
int b200_voc_num_words(b200_voc_t h);
/* TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup) (:1218-1259) for n descriptors (device, 32-byte aligned):
 * word_id [n], weight [n] (WordValue = double), node_id [n] = the node at level L - levelsup (0 when that level is <= 0), device outputs.
 * Frame::ComputeBoW (src/Frame.cc:348-355) builds mBowVec / mFeatVec from these (b200slam_adapters.hpp: transform). */
int b200_voc_transform(b200_voc_t h, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id, void* stream);
int b200_voc_transform_host(b200_voc_t h, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id);


/* ---------------------------------------------------------------- multi-GPU collation -- */
/* The reference is a single process; the B200 build shards a batch of frames over the GPUs of a node (contiguous blocks of frames per rank, one
 * process per GPU, no data-path collective) and has ONE exchange step: the fixed-size per-frame result slots travel to the consumer rank over
 * NCCL (NVLink 5 / NVSwitch).  NCCL is bound at run time (dlopen libnccl.so.2); without it these calls fail with B200_ENODEV and everything
 * else keeps working.  The communicator belongs to the library: rank 0 obtains an id, the host hands its 128 bytes to every rank (any
 * transport: a file, MPI, torch.distributed's store), every rank creates its handle. */
int b200_collate_unique_id(uint8_t* id128);
int b200_collate_create(b200_collate_t* out, const uint8_t* id128, int rank, int world, int device);
int b200_collate_destroy(b200_collate_t h);
int b200_collate_rank(b200_collate_t h);
int b200_collate_world(b200_collate_t h);
int b200_collate_nccl_version(void);                       /* e.g. 22809; 0 when NCCL could not be loaded */
/* Gather-to-root of n_buffers DEVICE buffers inside one NCCL group (grouped ncclSend / ncclRecv; the root's own block is a device copy).
 * send[b]: this rank's buffer of bytes[b] bytes (same sizes on every rank); recv[b] (root only): world * bytes[b] bytes, rank-major.
 * Only enqueues work on `stream`. */
int b200_collate_gather(b200_collate_t h, int n_buffers, const void* const* send, void* const* recv, const int64_t* bytes, int root, void* stream);
/* The all-gather form (every rank receives every block); recv on every rank. */
int b200_collate_allgather(b200_collate_t h, int n_buffers, const void* const* send, void* const* recv, const int64_t* bytes, void* stream);
/* payload bytes this rank has sent / received through the handle so far (the NVLink traffic of the SCALE report) */
int b200_collate_traffic(b200_collate_t h, int64_t* bytes_sent, int64_t* bytes_received);
/* Sharded front end: every rank has run b200_frontend_host on its own block of n frames (same n, geometry and options everywhere); the
 * device-resident copies of those results are gathered at `root` in ONE NCCL group and downloaded there into HOST buffers laid out rank-major,
 * [world][n][...] with the per-rank layout of b200_frontend_host's outputs.  Non-root ranks pass NULL outputs.  Synchronous. */
int b200_frontend_collate_host(b200_orb_t orb, b200_collate_t c, int root, b200_keypoint* kps_all, uint8_t* desc_all, int32_t* counts_all,
                               b200_marker* markers_all, int32_t* marker_counts_all, int32_t* match_all, int32_t* n_matches_all);

#ifdef __cplusplus
}
#endif
#endif /* B200SLAM_H */
