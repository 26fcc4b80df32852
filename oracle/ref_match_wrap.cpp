// oracle/ref_match_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN src/ORBmatcher.cc, which oracle/Makefile compiles
// unmodified from /root/reference against oracle/matchshim (stand-ins for cv::Mat and for the Frame / KeyFrame / MapPoint data the matcher
// touches) into oracle/_ref/libref_match.so.  Flat arrays in, the matcher's answers out, plus the trace of every Frame::GetFeaturesInArea query the
// matcher made (so the product path can be given the very same projections).  Used by tests/test_oracle_vs_ref.py to pin oracle/match_oracle.cpp
// and by tests/golden/make_match_golden.py to write tests/golden/match_ref.npz, which travels to the GPU box.
// The same file builds oracle/_ref/libadapter_match.so with -DB200_ADAPTER_MATCHER: then the matcher behind these entry points is the PRODUCT's
// signature-exact ORB_SLAM2::ORBmatcher (include/b200slam_orbmatcher.hpp over libb200slam.so) - every call below is spelled the way the reference's call
// sites spell it, so the two libraries answer the same questions and tests/test_orbmatcher_exact_gpu.py compares them.
#ifdef B200_ADAPTER_MATCHER
#include "b200slam_orbmatcher.hpp"
#else
#include "ORBmatcher.h"          // the reference's own header (slam_types.h is force-included in front of it)
#endif
#include <atomic>
#include <cstring>
#include <thread>

RefTrace g_ref_trace;
namespace ORB_SLAM2 {
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
}
using namespace ORB_SLAM2;

namespace {

cv::Mat desc_mat(const uint8_t* d, int n) { return cv::Mat(n, 32, CV_8U, (void*)d); }
cv::Mat desc_row(const uint8_t* d, int i) { return cv::Mat(1, 32, CV_8U, (void*)(d + 32 * (size_t)i)).clone(); }

std::vector<cv::KeyPoint> keys_from(const oracle_keypoint* k, int n) {
    std::vector<cv::KeyPoint> v(n);
    if (n) memcpy((void*)v.data(), k, (size_t)n * sizeof(cv::KeyPoint));
    return v;
}
std::vector<cv::KeyPoint> keys_from_angles(const float* a, int n) {
    std::vector<cv::KeyPoint> v(n);
    for (int i = 0; i < n; i++) v[i].angle = a[i];
    return v;
}
DBoW2::FeatureVector featvec(const int32_t* nodes, const int32_t* start, const int32_t* items, int nn) {
    DBoW2::FeatureVector fv;
    for (int j = 0; j < nn; j++) fv[(DBoW2::NodeId)nodes[j]] = std::vector<unsigned int>(items + start[j], items + start[j + 1]);
    return fv;
}
void scale_pyramid(std::vector<float>& sf, std::vector<float>& sigma2, std::vector<float>& inv_sigma2, int levels, float f) {
    sf.assign(levels, 1.0f); sigma2.assign(levels, 1.0f); inv_sigma2.assign(levels, 1.0f);      // src/ORBextractor.cc:418-428
    for (int i = 1; i < levels; i++) { sf[i] = sf[i - 1] * f; sigma2[i] = sf[i] * sf[i]; }
    for (int i = 0; i < levels; i++) inv_sigma2[i] = 1.0f / sigma2[i];
}
void set_frame(Frame& F, const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4) {
    F.N = n;
    F.mvKeysUn = keys_from(k, n); F.mvKeys = F.mvKeysUn;
    F.mDescriptors = desc_mat(d, n);
    F.mvuRight.assign(n, -1.0f); F.mvDepth.assign(n, -1.0f);                                     // monocular (src/Frame.cc:213-215)
    F.mvpMapPoints.assign(n, (MapPoint*)NULL); F.mvbOutlier.assign(n, false);
    scale_pyramid(F.mvScaleFactors, F.mvLevelSigma2, F.mvInvLevelSigma2, 8, 1.2f);
    Frame::mnMinX = bounds4[0]; Frame::mnMaxX = bounds4[1]; Frame::mnMinY = bounds4[2]; Frame::mnMaxY = bounds4[3];
    F.grid.build(F.mvKeysUn, bounds4);
}
cv::Mat mat44(const float* t) { return cv::Mat(4, 4, CV_32F, (void*)t).clone(); }
cv::Mat vec3(const float* p) { return cv::Mat(3, 1, CV_32F, (void*)p).clone(); }

// copies the GetFeaturesInArea queries that found candidates (the matcher fetched the point's descriptor right after them)
int dump_trace(float* q_xyr, int32_t* q_lev, int32_t* q_mp, int cap) {
    int n = 0;
    for (size_t i = 0; i < g_ref_trace.mp.size(); i++) {
        if (g_ref_trace.mp[i] < 0) continue;
        if (n < cap) {
            for (int j = 0; j < 3; j++) q_xyr[3 * n + j] = g_ref_trace.xyr[3 * i + j];
            q_lev[2 * n] = g_ref_trace.levels[2 * i]; q_lev[2 * n + 1] = g_ref_trace.levels[2 * i + 1];
            q_mp[n] = g_ref_trace.mp[i];
        }
        n++;
    }
    return n;
}
void clear_trace() { g_ref_trace.xyr.clear(); g_ref_trace.levels.clear(); g_ref_trace.mp.clear(); }

// the frame's map points before the call: obs[i] < 0 none, otherwise a point with that many observations
void preset_points(Frame& F, const int32_t* obs, std::vector<MapPoint>& store) {
    store.assign(F.N, MapPoint());
    for (int i = 0; i < F.N; i++) if (obs[i] >= 0) { store[i].id = -2; store[i].nObs = obs[i]; F.mvpMapPoints[i] = &store[i]; }
}
void read_assign(const Frame& F, int32_t* assign) {
    for (int i = 0; i < F.N; i++) assign[i] = F.mvpMapPoints[i] ? F.mvpMapPoints[i]->id : -1;     // -2: still the preset point
}

}  // namespace

extern "C" {

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {                                 // src/ORBmatcher.cc:1651-1667
    return ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, (void*)a), cv::Mat(1, 32, CV_8U, (void*)b));
}

// SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:159-292); arguments as oracle_search_by_bow_nodes.  A keyframe feature with
// kf_valid = 0 has no MapPoint (even index) or a bad one (odd index) - both are skipped by the reference.
int ref_search_by_bow_nodes(const uint8_t* dkf, const float* akf, const uint8_t* kf_valid, int n_kf, const int32_t* kf_nodes, const int32_t* kf_start,
                            const int32_t* kf_items, int kf_nn, const uint8_t* df, const float* af, int n_f, const int32_t* f_nodes,
                            const int32_t* f_start, const int32_t* f_items, int f_nn, float nnratio, int check_ori, int32_t* matches) {
    KeyFrame kf; Frame F;
    kf.N = n_kf; kf.mvKeysUn = keys_from_angles(akf, n_kf); kf.mvKeys = kf.mvKeysUn; kf.mDescriptors = desc_mat(dkf, n_kf);
    kf.mFeatVec = featvec(kf_nodes, kf_start, kf_items, kf_nn);
    std::vector<MapPoint> pts(n_kf);
    kf.mvpMapPoints.assign(n_kf, (MapPoint*)NULL);
    for (int i = 0; i < n_kf; i++) {
        pts[i].id = i; pts[i].bad = !kf_valid[i];
        if (kf_valid[i] || (i & 1)) kf.mvpMapPoints[i] = &pts[i];
    }
    F.N = n_f; F.mvKeys = keys_from_angles(af, n_f); F.mvKeysUn = F.mvKeys; F.mDescriptors = desc_mat(df, n_f);
    F.mFeatVec = featvec(f_nodes, f_start, f_items, f_nn);
    std::vector<MapPoint*> vpMapPointMatches;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByBoW(&kf, F, vpMapPointMatches);
    for (int i = 0; i < n_f; i++) matches[i] = vpMapPointMatches[i] ? vpMapPointMatches[i]->id : -1;
    return n;
}

// SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (src/ORBmatcher.cc:526-659); arguments as oracle_search_by_bow_kfkf_nodes
int ref_search_by_bow_kfkf_nodes(const uint8_t* d1, const float* a1, const uint8_t* valid1, int n1, const int32_t* nodes1, const int32_t* start1,
                                 const int32_t* items1, int nn1, const uint8_t* d2, const float* a2, const uint8_t* valid2, int n2,
                                 const int32_t* nodes2, const int32_t* start2, const int32_t* items2, int nn2, float nnratio, int check_ori,
                                 int32_t* matches12) {
    KeyFrame k1, k2;
    std::vector<MapPoint> p1(n1), p2(n2);
    k1.N = n1; k1.mvKeysUn = keys_from_angles(a1, n1); k1.mDescriptors = desc_mat(d1, n1); k1.mFeatVec = featvec(nodes1, start1, items1, nn1);
    k2.N = n2; k2.mvKeysUn = keys_from_angles(a2, n2); k2.mDescriptors = desc_mat(d2, n2); k2.mFeatVec = featvec(nodes2, start2, items2, nn2);
    k1.mvpMapPoints.assign(n1, (MapPoint*)NULL); k2.mvpMapPoints.assign(n2, (MapPoint*)NULL);
    for (int i = 0; i < n1; i++) { p1[i].id = i; p1[i].bad = !valid1[i]; if (valid1[i] || (i & 1)) k1.mvpMapPoints[i] = &p1[i]; }
    for (int i = 0; i < n2; i++) { p2[i].id = i; p2[i].bad = !valid2[i]; if (valid2[i] || (i & 1)) k2.mvpMapPoints[i] = &p2[i]; }
    std::vector<MapPoint*> vpMatches12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByBoW(&k1, &k2, vpMatches12);
    for (int i = 0; i < n1; i++) matches12[i] = vpMatches12[i] ? vpMatches12[i]->id : -1;
    return n;
}

// SearchForInitialization (src/ORBmatcher.cc:409-524); arguments as oracle_search_for_initialization
int ref_search_for_initialization(const oracle_keypoint* k1, const uint8_t* d1, int n1, const oracle_keypoint* k2, const uint8_t* d2, int n2,
                                  const float* bounds4, float* prev_matched, int window, float nnratio, int check_ori, int32_t* matches12) {
    Frame F1, F2;
    set_frame(F1, k1, d1, n1, bounds4); set_frame(F2, k2, d2, n2, bounds4);
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    clear_trace();
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window);
    for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return n;
}

// SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129).  frame_obs [n2]: the frame's map points before the call.
// Map point m: in_view / bad flags, predicted level, viewing cosine, projection, descriptor, observations.  assign [n2] out: map point index now
// held by the keypoint (-1 none, -2 the preset one).  q_* out: the grid queries that found candidates, q_mp = map point of each.
int ref_search_by_projection_points(const oracle_keypoint* k2, const uint8_t* d2, int n2, const float* bounds4, const int32_t* frame_obs,
                                    int n_mp, const uint8_t* mp_in_view, const uint8_t* mp_bad, const int32_t* mp_level, const float* mp_viewcos,
                                    const float* mp_projxy, const uint8_t* mp_desc, const int32_t* mp_nobs, float th, float nnratio,
                                    int32_t* assign, float* q_xyr, int32_t* q_lev, int32_t* q_mp, int32_t* n_queries) {
    Frame F;
    set_frame(F, k2, d2, n2, bounds4);
    std::vector<MapPoint> held;
    preset_points(F, frame_obs, held);
    std::vector<MapPoint> pts(n_mp);
    std::vector<MapPoint*> vp(n_mp);
    for (int m = 0; m < n_mp; m++) {
        pts[m].id = m; pts[m].mbTrackInView = mp_in_view[m] != 0; pts[m].bad = mp_bad[m] != 0; pts[m].mnTrackScaleLevel = mp_level[m];
        pts[m].mTrackViewCos = mp_viewcos[m]; pts[m].mTrackProjX = mp_projxy[2 * m]; pts[m].mTrackProjY = mp_projxy[2 * m + 1];
        pts[m].descriptor = desc_row(mp_desc, m); pts[m].nObs = mp_nobs[m];
        vp[m] = &pts[m];
    }
    ORBmatcher matcher(nnratio, true);
    clear_trace();
    const int n = matcher.SearchByProjection(F, vp, th);
    read_assign(F, assign);
    *n_queries = dump_trace(q_xyr, q_lev, q_mp, n_mp);
    return n;
}

// SearchByProjection(Frame& Current, const Frame& Last, th, bMono = true) (src/ORBmatcher.cc:1332-1474).  cam4 = fx fy cx cy, tcw_* = row-major 4x4.
// Last-frame keypoint i: map point present?, outlier?, world position, descriptor, observations.  q_mp = last-frame index of each query.
int ref_search_by_projection_last(const oracle_keypoint* k2, const uint8_t* d2, int n2, const float* bounds4, const int32_t* frame_obs,
                                  const float* cam4, const float* tcw_cur, const float* tcw_last, const oracle_keypoint* k_last, int n_last,
                                  const uint8_t* mp_present, const uint8_t* mp_outlier, const float* mp_pos, const uint8_t* mp_desc,
                                  const int32_t* mp_nobs, float th, int check_ori, int32_t* assign, float* q_xyr, int32_t* q_lev, int32_t* q_mp,
                                  int32_t* n_queries) {
    Frame C, L;
    set_frame(C, k2, d2, n2, bounds4);
    std::vector<MapPoint> held;
    preset_points(C, frame_obs, held);
    Frame::fx = cam4[0]; Frame::fy = cam4[1]; Frame::cx = cam4[2]; Frame::cy = cam4[3];
    C.mTcw = mat44(tcw_cur);
    L.N = n_last; L.mvKeys = keys_from(k_last, n_last); L.mvKeysUn = L.mvKeys; L.mTcw = mat44(tcw_last);
    std::vector<MapPoint> pts(n_last);
    L.mvpMapPoints.assign(n_last, (MapPoint*)NULL); L.mvbOutlier.assign(n_last, false);
    for (int i = 0; i < n_last; i++) {
        pts[i].id = i; pts[i].worldPos = vec3(mp_pos + 3 * i); pts[i].descriptor = desc_row(mp_desc, i); pts[i].nObs = mp_nobs[i];
        if (mp_present[i]) L.mvpMapPoints[i] = &pts[i];
        L.mvbOutlier[i] = mp_outlier[i] != 0;
    }
    ORBmatcher matcher(0.9f, check_ori != 0);                                                     // Tracking.cc:1172 ORBmatcher matcher(0.9,true)
    clear_trace();
    const int n = matcher.SearchByProjection(C, L, th, true);
    read_assign(C, assign);
    *n_queries = dump_trace(q_xyr, q_lev, q_mp, n_last);
    return n;
}

// SearchByProjection(Frame& Current, KeyFrame*, const set<MapPoint*>& sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1476-1603, relocalisation).
// Keyframe feature i: map point state 0 none / 1 good / 2 bad / 3 already found; position, descriptor, mfMinDistance / mfMaxDistance.
int ref_search_by_projection_reloc(const oracle_keypoint* k2, const uint8_t* d2, int n2, const float* bounds4, const int32_t* frame_obs,
                                   const float* cam4, const float* tcw_cur, const oracle_keypoint* k_kf, int n_kf, const uint8_t* mp_state,
                                   const float* mp_pos, const uint8_t* mp_desc, const float* mp_minmax, float th, int orb_dist, int check_ori,
                                   int32_t* assign, float* q_xyr, int32_t* q_lev, int32_t* q_mp, int32_t* n_queries) {
    Frame C;
    set_frame(C, k2, d2, n2, bounds4);
    std::vector<MapPoint> held;
    preset_points(C, frame_obs, held);
    Frame::fx = cam4[0]; Frame::fy = cam4[1]; Frame::cx = cam4[2]; Frame::cy = cam4[3];
    C.mTcw = mat44(tcw_cur);
    KeyFrame kf;
    kf.N = n_kf; kf.mvKeysUn = keys_from(k_kf, n_kf); kf.mvKeys = kf.mvKeysUn;
    std::vector<MapPoint> pts(n_kf);
    kf.mvpMapPoints.assign(n_kf, (MapPoint*)NULL);
    std::set<MapPoint*> found;
    for (int i = 0; i < n_kf; i++) {
        pts[i].id = i; pts[i].worldPos = vec3(mp_pos + 3 * i); pts[i].descriptor = desc_row(mp_desc, i); pts[i].nObs = 1;
        pts[i].minDistance = mp_minmax[2 * i]; pts[i].maxDistance = mp_minmax[2 * i + 1];
        pts[i].bad = mp_state[i] == 2;
        if (mp_state[i]) kf.mvpMapPoints[i] = &pts[i];
        if (mp_state[i] == 3) found.insert(&pts[i]);
    }
    ORBmatcher matcher(0.9f, check_ori != 0);                                                     // Tracking.cc:1798 ORBmatcher matcher2(0.9,true)
    clear_trace();
    const int n = matcher.SearchByProjection(C, &kf, found, th, orb_dist);
    read_assign(C, assign);
    *n_queries = dump_trace(q_xyr, q_lev, q_mp, n_kf);
    return n;
}

}  // extern "C"

// ---- the KeyFrame-side searches (local mapping / loop closing threads) ------------------------------------------------------------------------
// Shared argument blocks.  Keyframe: undistorted keypoints, descriptors, bounds4 (integers in the reference's KeyFrame, include/KeyFrame.h:211-214),
// cam4 = fx fy cx cy, pose T (row-major 4x4, Tcw).  Map point list: state 0 = NULL / 1 = good / 2 = bad / 3 = good and already in the keyframe
// (an observation of it / a member of spAlreadyFound), world position, mean viewing direction, descriptor, mfMinDistance / mfMaxDistance, Observations().
namespace {
void set_keyframe(KeyFrame& kf, const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* T) {
    kf.N = n; kf.mvKeysUn = keys_from(k, n); kf.mvKeys = kf.mvKeysUn; kf.mDescriptors = desc_mat(d, n);
    kf.mvuRight.assign(n, -1.0f); kf.mvDepth.assign(n, -1.0f);
    kf.mvpMapPoints.assign(n, (MapPoint*)NULL);
    std::vector<float> inv;
    scale_pyramid(kf.mvScaleFactors, kf.mvLevelSigma2, kf.mvInvLevelSigma2, 8, 1.2f);
    kf.mnMinX = (int)bounds4[0]; kf.mnMaxX = (int)bounds4[1]; kf.mnMinY = (int)bounds4[2]; kf.mnMaxY = (int)bounds4[3];
    kf.fx = cam4[0]; kf.fy = cam4[1]; kf.cx = cam4[2]; kf.cy = cam4[3];
    if (T) kf.Tcw = mat44(T);
    // the float bounds every KeyFrame truncates (include/KeyFrame.h:211-214) are Frame's statics (include/Frame.h:191-194)
    Frame::mnMinX = bounds4[0]; Frame::mnMaxX = bounds4[1]; Frame::mnMinY = bounds4[2]; Frame::mnMaxY = bounds4[3];
    kf.grid.build(kf.mvKeysUn, bounds4);
}
void set_points(std::vector<MapPoint>& pts, int n_mp, const uint8_t* state, const float* pos, const float* normal, const uint8_t* desc, const float* minmax,
                const int32_t* nobs) {
    pts.assign(n_mp, MapPoint());
    for (int m = 0; m < n_mp; m++) {
        pts[m].id = m; pts[m].bad = state[m] == 2; pts[m].worldPos = vec3(pos + 3 * m); pts[m].normal = normal ? vec3(normal + 3 * m) : cv::Mat();
        pts[m].descriptor = desc_row(desc, m); pts[m].minDistance = minmax[2 * m]; pts[m].maxDistance = minmax[2 * m + 1]; pts[m].nObs = nobs ? nobs[m] : 1;
    }
}
// the points a keyframe holds before the call: held_state [n] 0 none / 1 good / 2 bad, held_nobs [n]; their ids are 1000000 + feature index
void hold_points(KeyFrame& kf, std::vector<MapPoint>& held, const uint8_t* held_state, const int32_t* held_nobs) {
    held.assign(kf.N, MapPoint());
    for (int i = 0; i < kf.N; i++) if (held_state[i]) {
        held[i].id = 1000000 + i; held[i].bad = held_state[i] == 2; held[i].nObs = held_nobs ? held_nobs[i] : 1; held[i].observations[&kf] = i;
        kf.mvpMapPoints[i] = &held[i];
    }
}
int dump_trace_all(float* q_xyr, int32_t* q_lev, int32_t* q_mp, int cap) {       // every query, also those that found nothing (q_mp = -1)
    int n = 0;
    for (size_t i = 0; i < g_ref_trace.mp.size(); i++, n++) if (n < cap) {
        for (int j = 0; j < 3; j++) q_xyr[3 * n + j] = g_ref_trace.xyr[3 * i + j];
        q_lev[2 * n] = g_ref_trace.levels[2 * i]; q_lev[2 * n + 1] = g_ref_trace.levels[2 * i + 1];
        q_mp[n] = g_ref_trace.mp[i];
    }
    return n;
}
}  // namespace

extern "C" {

// SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo = false) (src/ORBmatcher.cc:661-829), monocular keyframes.  has_mp* [n]: the
// feature already owns a MapPoint (skipped).  FeatureVectors as in ref_search_by_bow_nodes.  F12 row-major 3x3.  matches12 [n1] out (-1 none): the
// pairs (i, matches12[i]) in ascending i are vMatchedPairs.
int ref_search_for_triangulation(const oracle_keypoint* k1, const uint8_t* d1, const uint8_t* has_mp1, int n1, const int32_t* nodes1, const int32_t* start1,
                                 const int32_t* items1, int nn1, const float* T1, const oracle_keypoint* k2, const uint8_t* d2, const uint8_t* has_mp2, int n2,
                                 const int32_t* nodes2, const int32_t* start2, const int32_t* items2, int nn2, const float* T2, const float* bounds4,
                                 const float* cam4, const float* F12, int check_ori, int32_t* matches12) {
    KeyFrame a, b;
    set_keyframe(a, k1, d1, n1, bounds4, cam4, T1); set_keyframe(b, k2, d2, n2, bounds4, cam4, T2);
    a.mFeatVec = featvec(nodes1, start1, items1, nn1); b.mFeatVec = featvec(nodes2, start2, items2, nn2);
    MapPoint some;
    for (int i = 0; i < n1; i++) if (has_mp1[i]) a.mvpMapPoints[i] = &some;
    for (int i = 0; i < n2; i++) if (has_mp2[i]) b.mvpMapPoints[i] = &some;
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher matcher(0.6f, check_ori != 0);                                                     // LocalMapping.cc:230 ORBmatcher matcher(0.6,false)
    const int n = matcher.SearchForTriangulation(&a, &b, cv::Mat(3, 3, CV_32F, (void*)F12).clone(), pairs, false);
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    for (size_t j = 0; j < pairs.size(); j++) matches12[pairs[j].first] = (int32_t)pairs[j].second;
    return n;
}

// Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:831-981).  held_* = the keyframe's own points before the call.  Per map point m out:
// fused_idx [n_mp] = keyframe feature it ended up with (-1 none / not visible in the state) and action [n_mp]: 0 nothing visible, 1 added as a new observation,
// 2 pMP->Replace(pMPinKF) (the keyframe's point survives), 3 pMPinKF->Replace(pMP).  Returns nFused (which also counts matches to a bad keyframe point).
int ref_fuse(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* T, const uint8_t* held_state,
             const int32_t* held_nobs, int n_mp, const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc,
             const float* mp_minmax, const int32_t* mp_nobs, float th, int32_t* fused_idx, int32_t* action, float* q_xyr, int32_t* q_lev, int32_t* q_mp,
             int32_t* n_queries) {
    KeyFrame kf;
    set_keyframe(kf, k, d, n, bounds4, cam4, T);
    std::vector<MapPoint> held, pts;
    hold_points(kf, held, held_state, held_nobs);
    set_points(pts, n_mp, mp_state, mp_pos, mp_normal, mp_desc, mp_minmax, mp_nobs);
    KeyFrame other;
    std::vector<MapPoint*> vp(n_mp, (MapPoint*)NULL);
    for (int m = 0; m < n_mp; m++) {
        if (mp_state[m]) vp[m] = &pts[m];
        if (mp_state[m] == 3) pts[m].observations[&kf] = 0;                                        // IsInKeyFrame(pKF)
    }
    ORBmatcher matcher(0.6f, true);                                                               // LocalMapping.cc:851 ORBmatcher matcher;
    clear_trace();
    const int nf = matcher.Fuse(&kf, vp, th);
    for (int m = 0; m < n_mp; m++) {
        fused_idx[m] = -1; action[m] = 0;
        if (mp_state[m] == 1 && pts[m].observations.count(&kf)) { fused_idx[m] = (int32_t)pts[m].observations[&kf]; action[m] = 1; }
    }
    // Replace() calls: the keyframe's own point i is named by its feature index, a point o of this call that was added earlier by -3 - o
    for (int m = 0; m < n_mp; m++) if (pts[m].replaced) {
        const int other = pts[m].replaced->id;
        if (other >= 1000000) { fused_idx[m] = other - 1000000; action[m] = 2; }                  // pMP->Replace(pMPinKF), the keyframe's own point
        else if (other < m) { fused_idx[m] = -3 - other; action[m] = 2; }                         // pMP->Replace(pMPinKF), pMPinKF added earlier in this call
        else { fused_idx[other] = -3 - m; action[other] = 3; }                                    // pMPinKF->Replace(pMP) in the later point's turn
    }
    for (int i = 0; i < n; i++) if (held[i].replaced) { fused_idx[held[i].replaced->id] = i; action[held[i].replaced->id] = 3; }
    *n_queries = dump_trace_all(q_xyr, q_lev, q_mp, n_mp);
    return nf;
}

// Fuse(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, th, vpReplacePoint) (src/ORBmatcher.cc:983-1104, loop closing).  S = row-major 4x4 Sim3.
// replace_idx [n_mp] out = keyframe feature whose point vpReplacePoint[m] names (-1 none); added_idx [n_mp] = feature the point was added to (-1 none).
int ref_fuse_sim3(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* S, const uint8_t* held_state,
                  int n_mp, const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc, const float* mp_minmax, float th,
                  int32_t* replace_idx, int32_t* added_idx, float* q_xyr, int32_t* q_lev, int32_t* q_mp, int32_t* n_queries) {
    KeyFrame kf;
    set_keyframe(kf, k, d, n, bounds4, cam4, NULL);
    std::vector<MapPoint> held, pts;
    hold_points(kf, held, held_state, NULL);
    set_points(pts, n_mp, mp_state, mp_pos, mp_normal, mp_desc, mp_minmax, NULL);
    std::vector<MapPoint*> vp(n_mp), vrep(n_mp, (MapPoint*)NULL);
    // state 3: the point IS one of the keyframe's good points (member of pKF->GetMapPoints()): hand the keyframe's own pointer in
    int next_held = 0;
    for (int m = 0; m < n_mp; m++) {
        vp[m] = &pts[m];
        if (mp_state[m] == 3) {
            while (next_held < n && held_state[next_held] != 1) next_held++;
            if (next_held < n) { MapPoint& h = held[next_held++]; h.worldPos = pts[m].worldPos; h.normal = pts[m].normal; h.descriptor = pts[m].descriptor;
                                 h.minDistance = pts[m].minDistance; h.maxDistance = pts[m].maxDistance; vp[m] = &h; }
        }
    }
    ORBmatcher matcher(0.8f, true);                                                               // LoopClosing.cc:1076 ORBmatcher matcher(0.8)
    clear_trace();
    const int nf = matcher.Fuse(&kf, mat44(S), vp, th, vrep);
    for (int m = 0; m < n_mp; m++) {
        replace_idx[m] = !vrep[m] ? -1 : vrep[m]->id >= 1000000 ? vrep[m]->id - 1000000 : -3 - vrep[m]->id;   // -3 - o: point o, added earlier in this call
        added_idx[m] = (vp[m] == &pts[m] && pts[m].observations.count(&kf)) ? (int32_t)pts[m].observations[&kf] : -1;
    }
    *n_queries = dump_trace_all(q_xyr, q_lev, q_mp, n_mp);
    return nf;
}

// SearchByProjection(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, th) (src/ORBmatcher.cc:294-407, loop closing).
// matched [n] in/out: -1 none, >= 0 index into vpPoints, -2 some other point.
int ref_search_by_projection_loop(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* S, int n_mp,
                                  const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc, const float* mp_minmax, int th,
                                  int32_t* matched, float* q_xyr, int32_t* q_lev, int32_t* q_mp, int32_t* n_queries) {
    KeyFrame kf;
    set_keyframe(kf, k, d, n, bounds4, cam4, NULL);
    std::vector<MapPoint> pts;
    set_points(pts, n_mp, mp_state, mp_pos, mp_normal, mp_desc, mp_minmax, NULL);
    MapPoint other; other.id = -2;
    std::vector<MapPoint*> vp(n_mp), vm(n, (MapPoint*)NULL);
    for (int m = 0; m < n_mp; m++) vp[m] = &pts[m];
    for (int i = 0; i < n; i++) if (matched[i] >= 0) vm[i] = &pts[matched[i]]; else if (matched[i] == -2) vm[i] = &other;
    ORBmatcher matcher(0.75f, true);                                                              // LoopClosing.cc:493 ORBmatcher matcher(0.75,true)
    clear_trace();
    const int nm = matcher.SearchByProjection(&kf, mat44(S), vp, vm, th);
    for (int i = 0; i < n; i++) matched[i] = vm[i] ? vm[i]->id : -1;
    *n_queries = dump_trace_all(q_xyr, q_lev, q_mp, n_mp);
    return nm;
}

// SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (src/ORBmatcher.cc:1106-1330).  Feature i of keyframe x owns point state mpx_state[i]
// (0 none / 1 good / 2 bad) with position / descriptor / distances.  matches12 [n1] in/out: index of the KF2 feature whose point is matched, -1 none.
int ref_search_by_sim3(const oracle_keypoint* k1, const uint8_t* d1, int n1, const float* T1, const uint8_t* mp1_state, const float* mp1_pos,
                       const uint8_t* mp1_desc, const float* mp1_minmax, const oracle_keypoint* k2, const uint8_t* d2, int n2, const float* T2,
                       const uint8_t* mp2_state, const float* mp2_pos, const uint8_t* mp2_desc, const float* mp2_minmax, const float* bounds4,
                       const float* cam4, float s12, const float* R12, const float* t12, float th, int32_t* matches12, float* q_xyr, int32_t* q_lev,
                       int32_t* q_mp, int32_t* n_queries) {
    KeyFrame a, b;
    set_keyframe(a, k1, d1, n1, bounds4, cam4, T1); set_keyframe(b, k2, d2, n2, bounds4, cam4, T2);
    std::vector<MapPoint> p1, p2;
    set_points(p1, n1, mp1_state, mp1_pos, NULL, mp1_desc, mp1_minmax, NULL); set_points(p2, n2, mp2_state, mp2_pos, NULL, mp2_desc, mp2_minmax, NULL);
    for (int i = 0; i < n1; i++) if (mp1_state[i]) { a.mvpMapPoints[i] = &p1[i]; p1[i].observations[&a] = i; }
    for (int i = 0; i < n2; i++) { p2[i].id = 2000000 + i; if (mp2_state[i]) { b.mvpMapPoints[i] = &p2[i]; p2[i].observations[&b] = i; } }
    std::vector<MapPoint*> vm(n1, (MapPoint*)NULL);
    for (int i = 0; i < n1; i++) if (matches12[i] >= 0) vm[i] = &p2[matches12[i]];
    ORBmatcher matcher(0.75f, true);                                                              // LoopClosing.cc:398 ORBmatcher matcher(0.75,true)
    clear_trace();
    const int nf = matcher.SearchBySim3(&a, &b, vm, s12, cv::Mat(3, 3, CV_32F, (void*)R12).clone(), cv::Mat(3, 1, CV_32F, (void*)t12).clone(), th);
    for (int i = 0; i < n1; i++) matches12[i] = vm[i] ? vm[i]->id - 2000000 : -1;
    *n_queries = dump_trace_all(q_xyr, q_lev, q_mp, n1 + n2);
    return nf;
}

// batch driver for the CPU reference arm of bench.py: the reference's own SearchByBoW(KeyFrame*, Frame&, ...) with one all-inclusive vocabulary node
// (== brute force) for one reference set against n frames laid out like the extractor's output (desc [n][cap][32], angles inside
// 28-byte keypoint records), nthreads workers (the member touches no shared state).  Same argument list as oracle_search_by_bow_bf_batch.
int ref_search_by_bow_bf_batch(const uint8_t* kf_desc, const float* kf_angle, int n_kf, const uint8_t* f_desc, const uint8_t* f_kps28, const int32_t* n_f, int n,
                               int cap, float nnratio, int check_ori, int32_t* matches, int32_t* n_matches, int nthreads) {
    KeyFrame kf;
    kf.N = n_kf; kf.mvKeysUn = keys_from_angles(kf_angle, n_kf); kf.mvKeys = kf.mvKeysUn; kf.mDescriptors = desc_mat(kf_desc, n_kf);
    std::vector<unsigned int> all_kf(n_kf);
    for (int i = 0; i < n_kf; i++) all_kf[i] = i;
    kf.mFeatVec[7] = all_kf;
    std::vector<MapPoint> pts(n_kf);
    kf.mvpMapPoints.assign(n_kf, (MapPoint*)NULL);
    for (int i = 0; i < n_kf; i++) { pts[i].id = i; kf.mvpMapPoints[i] = &pts[i]; }
    std::atomic<int> next(0);
    auto work = [&]() {
        for (int f; (f = next.fetch_add(1)) < n;) {
            const int nf = n_f[f];
            Frame F;
            F.N = nf; F.mvKeys.resize(nf);
            for (int i = 0; i < nf; i++) memcpy(&F.mvKeys[i].angle, f_kps28 + ((size_t)f * cap + i) * 28 + 12, 4);
            F.mvKeysUn = F.mvKeys; F.mDescriptors = desc_mat(f_desc + (size_t)f * cap * 32, nf);
            std::vector<unsigned int> all_f(nf);
            for (int i = 0; i < nf; i++) all_f[i] = i;
            F.mFeatVec[7] = all_f;
            std::vector<MapPoint*> vp;
            ORBmatcher matcher(nnratio, check_ori != 0);
            n_matches[f] = matcher.SearchByBoW(&kf, F, vp);
            for (int i = 0; i < nf; i++) matches[(size_t)f * cap + i] = vp[i] ? vp[i]->id : -1;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return 0;
}

}  // extern "C"
