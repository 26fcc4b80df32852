// oracle/oracle.h -- TEST INFRASTRUCTURE (CPU checker), not product code.
// Plain-C interface of liboracle.so: the CPU restatement of the reference's per-frame front end.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// Same record as the product's b200_keypoint (include/b200slam.h) == cv::KeyPoint's 28 bytes.
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} oracle_keypoint;
// same record as b200_marker: id + 4 corners
typedef struct { int32_t id; float xy[8]; } oracle_marker;

// ---- primitives (pinned against cv2 golden vectors) -------------------------------------------
void oracle_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep);
void oracle_border_reflect101(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep, int border);
void oracle_gaussian_blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep);
// returns count; xys = [cap][3] (x, y, score)
int  oracle_fast_nms(const uint8_t* img, int w, int h, int step, int thr, int32_t* xys, int cap);
float oracle_fast_atan2(float y, float x);

// ---- extractor: restatement of ORB_SLAM2::ORBextractor (src/ORBextractor.cc:410-1132) ---------
// level geometry: writes nlevels entries of w,h,quota; returns 0
int oracle_orb_levels(int w, int h, int nfeatures, float scale, int nlevels, int32_t* lw, int32_t* lh, int32_t* quota,
                      float* scale_factor);
// pyramid level without border; out holds lw*lh bytes
int oracle_orb_pyramid_level(const uint8_t* img, int w, int h, int stride, float scale, int nlevels, int level, uint8_t* out);
// candidates of one level after the per-cell FAST stage (before the quadtree), level coords relative
// to minBorder (16): xys [cap][3]; returns count or -1
int oracle_orb_candidates(const uint8_t* img, int w, int h, int stride, int nfeatures, float scale, int nlevels,
                          int ini_th, int min_th, int level, int32_t* xys, int cap);
// full extractor on one frame; returns number of keypoints (or -1 if cap too small)
int oracle_orb_extract(const uint8_t* img, int w, int h, int stride, int nfeatures, float scale, int nlevels,
                       int ini_th, int min_th, oracle_keypoint* kps, uint8_t* desc, int cap);
// batch version, std::thread pool over frames with nthreads threads; counts[n]; kps/desc are [n][cap]
int oracle_orb_extract_batch(const uint8_t* imgs, int n, int w, int h, int row_stride, long frame_stride,
                             int nfeatures, float scale, int nlevels, int ini_th, int min_th,
                             oracle_keypoint* kps, uint8_t* desc, int32_t* counts, int cap, int nthreads);

// ---- matcher: restatement of ORB_SLAM2::ORBmatcher cores (src/ORBmatcher.cc) -------------------
int oracle_descriptor_distance(const uint8_t* a, const uint8_t* b);
int oracle_search_by_bow_bf(const uint8_t* kf_desc, const float* kf_angle, int n_kf,
                            const uint8_t* f_desc, const float* f_angle, int n_f,
                            float nnratio, int check_ori, float factor, int32_t* matches);
int oracle_search_by_bow_bf_batch(const uint8_t* kf_desc, const float* kf_angle, int n_kf,
                                  const uint8_t* f_desc, const uint8_t* f_kps28, const int32_t* n_f, int n, int cap,
                                  float nnratio, int check_ori, float factor, int32_t* matches, int32_t* n_matches, int nthreads);
void oracle_match_candidates(const uint8_t* qd, int nq, const uint8_t* td, const int32_t* ofs, const int32_t* cand,
                             int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);

// ---- ArUco detector: restatement of aruco::MarkerDetector::detect (Thirdparty/aruco/aruco) ----------
int oracle_aruco_detect(const uint8_t* img, int w, int h, int stride, const char* dict_name, oracle_marker* out, int cap);
int oracle_aruco_detect_batch(const uint8_t* imgs, int n, int w, int h, int row_stride, long frame_stride, const char* dict_name,
                              oracle_marker* out, int32_t* counts, int cap, int nthreads);
int oracle_aruco_stages(const uint8_t* img, int w, int h, int stride, const char* dict_name,
                        uint8_t* thres, int32_t* contour_sizes, int32_t* contour_pts, int max_contours, int max_points, int32_t* n_contours,
                        float* candidates, uint8_t* patches, int max_cand, int32_t* n_cand,
                        float* prerefine, oracle_marker* out, int cap);
int oracle_aruco_decode_patch(const uint8_t* patch, int size, const char* dict_name, int32_t* id, int32_t* nrot);
int oracle_dictionary_codes(const char* dict_name, uint64_t* codes, int cap, int32_t* nbits, int32_t* tau);
void oracle_adaptive_threshold(const uint8_t* src, int w, int h, uint8_t* dst, int bs, int C);
int oracle_find_contours(const uint8_t* img, int w, int h, int32_t* sizes, int32_t* pts, int max_contours, int max_points);
int oracle_approx_poly(const int32_t* pts, int n, double eps, int32_t* out, int* convex);
void oracle_resize_half(const uint8_t* src, int sw, int sh, uint8_t* dst);
void oracle_perspective_transform(const float* src, const float* dst, double* M);
void oracle_warp_perspective(const uint8_t* src, int sw, int sh, uint8_t* dst, int ds, const double* M);
int oracle_otsu(const uint8_t* img, int w, int h);
void oracle_solve_svd(const float* A, const float* b, int m, int n, float* x);

// ---- marker pose: restatement of aruco::Marker::calculateExtrinsics / IPPE (Thirdparty/aruco/aruco/ippe.cpp) ----------
// corners [4][2], cam9 = fx fy cx cy k1 k2 p1 p2 k3, out14 = rvec1[3] tvec1[3] err1 rvec2[3] tvec2[3] err2
void oracle_ippe_marker_pose(const float* corners, float msize, const double* cam9, double* out14);

// ---- frame grid: restatement of Frame::UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cc) ----
void oracle_undistort_keypoints(const oracle_keypoint* in, int n, const double* cam9, oracle_keypoint* out);
void oracle_image_bounds(int w, int h, const double* cam9, float* bounds4);
void oracle_assign_grid(const oracle_keypoint* un, int n, const float* bounds4, int32_t* cell_start, int32_t* cell_items);
int oracle_features_in_area(const oracle_keypoint* un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                            float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);
int oracle_keyframe_features_in_area(const oracle_keypoint* un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                            float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:271-331): BestIdx over desc [n][32], -1 when n == 0
int oracle_distinctive_descriptor(const uint8_t* desc, int n);
// SearchByBoW(KeyFrame*, Frame&) (src/ORBmatcher.cc:159-292) over real FeatureVectors given as sorted arrays (nodes, start, items); matches [n_f] out
int oracle_search_by_bow_nodes(const uint8_t* dkf, const float* akf, const uint8_t* kf_valid, const int32_t* kf_nodes, const int32_t* kf_start,
                               const int32_t* kf_items, int kf_nn, const uint8_t* df, const float* af, int n_f, const int32_t* f_nodes,
                               const int32_t* f_start, const int32_t* f_items, int f_nn, float nnratio, int check_ori, int32_t* matches);
// SearchByBoW(KeyFrame*, KeyFrame*) (src/ORBmatcher.cc:526-659) over real FeatureVectors; matches12 [n1] out
int oracle_search_by_bow_kfkf_nodes(const uint8_t* d1, const float* a1, const uint8_t* valid1, int n1, const int32_t* nodes1, const int32_t* start1,
                                    const int32_t* items1, int nn1, const uint8_t* d2, const float* a2, const uint8_t* valid2, int n2,
                                    const int32_t* nodes2, const int32_t* start2, const int32_t* items2, int nn2, float nnratio, int check_ori,
                                    int32_t* matches12);
// SearchByBoW(KeyFrame*, KeyFrame*) (src/ORBmatcher.cc:526-659), one all-inclusive vocabulary node; matches12 [n1] out; returns nmatches
int oracle_search_by_bow_kfkf_bf(const uint8_t* d1, const float* a1, int n1, const uint8_t* d2, const float* a2, int n2,
                                 float nnratio, int check_ori, int32_t* matches12);
// SearchByProjection(Frame, mapPoints, th) (mode 0, src/ORBmatcher.cc:45-129) / SearchByProjection(Current, Last, th, mono) (mode 1, :1332-1474)
// on ready-made projections; occupied [n2] in/out, assign [n2] out; returns nmatches
int oracle_search_by_projection(const oracle_keypoint* k2, const uint8_t* d2, int n2, const float* bounds4, uint8_t* occupied,
                                const float* q_xyr, const int32_t* q_lev, const uint8_t* q_desc, const float* q_angle, const uint8_t* q_observed, int nq,
                                int mode, float nnratio, int check_ori, int th_high, int32_t* assign);
// SearchForInitialization (src/ORBmatcher.cc:409-524) on arrays; prev_matched [n1][2] in/out, matches12 [n1] out; returns nmatches
int oracle_search_for_initialization(const oracle_keypoint* k1, const uint8_t* d1, int n1, const oracle_keypoint* k2, const uint8_t* d2, int n2,
                                     const float* bounds4, float* prev_matched, int window, float nnratio, int check_ori, int32_t* matches12);

// ---- bag of words: restatement of DBoW2 TemplatedVocabulary::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1259) ----------
void oracle_voc_transform(const int32_t* parent, const uint8_t* is_leaf, const uint8_t* node_desc, const double* node_weight, int n_nodes, int L,
                          const uint8_t* feat, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id);
void oracle_voc_vectors(const int32_t* word_id, const double* weight, const int32_t* node_id, int n,
                        int32_t* bow_words, double* bow_values, int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_items, int32_t* counts2);

// ---- KeyFrame-side ORBmatcher members (oracle/match2_oracle.cpp; argument lists = the ref_* wrappers of oracle/ref_match_wrap.cpp minus the trace) ----
void oracle_kf_radius_search(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* q_xyr, const int32_t* q_level,
                             const uint8_t* q_desc, int nq, float scale_factor, int nlevels, double chi2, int32_t* best_idx, int32_t* best_dist);
int oracle_search_for_triangulation(const oracle_keypoint* k1, const uint8_t* d1, const uint8_t* has_mp1, int n1, const int32_t* nodes1, const int32_t* start1,
                                    const int32_t* items1, int nn1, const float* T1, const oracle_keypoint* k2, const uint8_t* d2, const uint8_t* has_mp2, int n2,
                                    const int32_t* nodes2, const int32_t* start2, const int32_t* items2, int nn2, const float* T2, const float* bounds4,
                                    const float* cam4, const float* F12, int check_ori, int32_t* matches12);
int oracle_fuse(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* T, const uint8_t* held_state,
                const int32_t* held_nobs, int n_mp, const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc,
                const float* mp_minmax, const int32_t* mp_nobs, float th, int32_t* fused_idx, int32_t* action);
int oracle_fuse_sim3(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* S, const uint8_t* held_state,
                     int n_mp, const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc, const float* mp_minmax, float th,
                     int32_t* replace_idx, int32_t* added_idx);
int oracle_search_by_projection_loop(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* S, int n_mp,
                                     const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc, const float* mp_minmax, int th,
                                     int32_t* matched);
int oracle_search_by_sim3(const oracle_keypoint* k1, const uint8_t* d1, int n1, const float* T1, const uint8_t* mp1_state, const float* mp1_pos,
                          const uint8_t* mp1_desc, const float* mp1_minmax, const oracle_keypoint* k2, const uint8_t* d2, int n2, const float* T2,
                          const uint8_t* mp2_state, const float* mp2_pos, const uint8_t* mp2_desc, const float* mp2_minmax, const float* bounds4,
                          const float* cam4, float s12, const float* R12, const float* t12, float th, int32_t* matches12);

#ifdef __cplusplus
}
#endif
