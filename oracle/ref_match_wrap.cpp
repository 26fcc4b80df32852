// oracle/ref_match_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN src/ORBmatcher.cc, which oracle/Makefile compiles
// unmodified from /root/reference against oracle/matchshim (stand-ins for cv::Mat and for the Frame / KeyFrame / MapPoint data the matcher
// touches) into oracle/_ref/libref_match.so.  Flat arrays in, the matcher's answers out, plus the trace of every Frame::GetFeaturesInArea query the
// matcher made (so the product path can be given the very same projections).  Used by tests/test_oracle_vs_ref.py to pin oracle/match_oracle.cpp
// and by tests/golden/make_match_golden.py to write tests/golden/match_ref.npz, which travels to the GPU box.
#include "ORBmatcher.h"          // the reference's own header (slam_types.h is force-included in front of it)
#include <cstring>

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
