// oracle/ref_mappoint_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN src/MapPoint.cc (with include/MapPoint.h) and
// src/ORBmatcher.cc, compiled unmodified against oracle/matchshim with REF_REAL_MAPPOINT (KeyFrame / Frame / Map stay data stand-ins) into
// oracle/_ref/libref_mappoint.so (oracle/Makefile).  Pins MapPoint::ComputeDistinctiveDescriptors (oracle_distinctive_descriptor),
// MapPoint::PredictScale and Get{Min,Max}DistanceInvariance (the level prediction behind the KeyFrame-side matcher members: kfgeom.level_thresholds,
// k_kf_project) - tests/test_oracle_mappoint_vs_ref.py.
#include "MapPoint.h"
#include "ORBmatcher.h"
#include <cstring>

RefTrace g_ref_trace;
namespace ORB_SLAM2 {
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
std::set<MapPoint*> KeyFrame::GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
    return s;
}
}  // namespace ORB_SLAM2
using namespace ORB_SLAM2;

namespace {
void pyramid(KeyFrame& kf, int levels, float f) {                  // src/ORBextractor.cc:418-428, src/Frame.cc:93-100
    kf.mnScaleLevels = levels; kf.mfScaleFactor = f; kf.mfLogScaleFactor = std::log(f);
    kf.mvScaleFactors.assign(levels, 1.0f);
    for (int i = 1; i < levels; i++) kf.mvScaleFactors[i] = kf.mvScaleFactors[i - 1] * f;
}
cv::Mat vec3(float x, float y, float z) { cv::Mat m(3, 1, CV_32F); m.at<float>(0) = x; m.at<float>(1) = y; m.at<float>(2) = z; return m; }
}  // namespace

extern "C" {

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:271-331) for a point observed by n keyframes, keyframe i contributing descriptor row i.
// bad [n] (may be NULL): keyframe i is bad and skipped.  out32: GetDescriptor() afterwards (32 bytes; untouched when there is no observation); returns
// the number of descriptors that took part.
int ref_distinctive_descriptor(const uint8_t* desc, const uint8_t* bad, int n, uint8_t* out32) {
    Map map;
    std::vector<KeyFrame> kfs(n > 0 ? n : 1);
    for (int i = 0; i < (n > 0 ? n : 1); i++) {
        KeyFrame& kf = kfs[i];
        kf.mnId = i; kf.N = 1; kf.mvKeysUn.resize(1); kf.mvuRight.assign(1, -1.0f); kf.mvpMapPoints.assign(1, (MapPoint*)NULL);
        kf.Tcw = cv::Mat(4, 4, CV_32F);
        for (int r = 0; r < 4; r++) kf.Tcw.at<float>(r, r) = 1.0f;
        pyramid(kf, 8, 1.2f);
        if (i < n) { kf.mDescriptors = cv::Mat(1, 32, CV_8U, (void*)(desc + 32 * (size_t)i)).clone(); kf.bad = bad && bad[i]; }
    }
    MapPoint mp(vec3(0, 0, 4), &kfs[0], &map);
    for (int i = 0; i < n; i++) mp.AddObservation(&kfs[i], 0);
    mp.ComputeDistinctiveDescriptors();
    cv::Mat d = mp.GetDescriptor();
    int used = 0;
    for (int i = 0; i < n; i++) used += !(bad && bad[i]);
    if (!d.empty()) memcpy(out32, d.ptr<uchar>(0), 32);
    return d.empty() ? 0 : used;
}

// MapPoint::PredictScale(currentDist, KeyFrame*) (src/MapPoint.cc:403-418) for a point whose mfMaxDistance is max_distance: the reference keyframe sits
// at the origin and observes the point at (0, 0, max_distance) in level 0, so UpdateNormalAndDepth (src/MapPoint.cc:333-372) stores exactly that value.
// levels [n] out; inv2 out = GetMinDistanceInvariance(), GetMaxDistanceInvariance().
int ref_predict_scale(float max_distance, const float* dists, int n, float scale_factor, int nlevels, int32_t* levels, float* inv2) {
    Map map;
    KeyFrame kf;
    kf.mnId = 0; kf.N = 1; kf.mvKeysUn.resize(1); kf.mvKeysUn[0].octave = 0; kf.mvuRight.assign(1, -1.0f); kf.mvpMapPoints.assign(1, (MapPoint*)NULL);
    kf.Tcw = cv::Mat(4, 4, CV_32F);
    for (int r = 0; r < 4; r++) kf.Tcw.at<float>(r, r) = 1.0f;
    pyramid(kf, nlevels, scale_factor);
    MapPoint mp(vec3(0, 0, max_distance), &kf, &map);
    mp.AddObservation(&kf, 0);
    mp.UpdateNormalAndDepth();
    for (int i = 0; i < n; i++) levels[i] = mp.PredictScale(dists[i], &kf);
    inv2[0] = mp.GetMinDistanceInvariance(); inv2[1] = mp.GetMaxDistanceInvariance();
    return 0;
}

// Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:831-981) on REAL MapPoint objects: Replace() re-points keyframe slots and moves
// observations (src/MapPoint.cc:192-236), AddObservation counts (:110-121).  Scenario arguments as ref_fuse (oracle/ref_match_wrap.cpp): the keyframe's own
// point at feature i (held_state 0 none / 1 good / 2 bad, held_nobs) observes the keyframe at i; list point m: mp_state 0 NULL / 1 good / 2 bad / 3 already
// observed by the keyframe.  Out: slot [n] = who holds feature i afterwards (-1 nobody, 1000000 + j = the keyframe's own point j, m = list point m),
// mp_bad [n_mp], held_bad [n], mp_nobs_out [n_mp] = Observations() afterwards.  Returns nFused.
namespace {
struct TestPoint : public MapPoint {                               // reaches the protected state of the real class to set up a scene
    TestPoint(const cv::Mat& pos, KeyFrame* kf, Map* map) : MapPoint(pos, kf, map) {}
    void setup(const cv::Mat& normal, const cv::Mat& desc, float mind, float maxd, int nobs, bool bad) {
        mNormalVector = normal.clone(); mDescriptor = desc.clone(); mfMinDistance = mind; mfMaxDistance = maxd; nObs = nobs; mbBad = bad;
    }
    void observe(KeyFrame* kf, size_t idx) { mObservations[kf] = idx; }
};
}  // namespace

int ref_fuse_real(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* T, const uint8_t* held_state,
                  const int32_t* held_nobs, int n_mp, const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc,
                  const float* mp_minmax, const int32_t* mp_nobs, float th, int32_t* slot, uint8_t* mp_bad, uint8_t* held_bad, int32_t* mp_nobs_out) {
    Map map;
    KeyFrame kf;
    kf.mnId = 1; kf.N = n; kf.mvKeysUn.resize(n);
    if (n) memcpy((void*)kf.mvKeysUn.data(), k, (size_t)n * sizeof(cv::KeyPoint));
    kf.mvKeys = kf.mvKeysUn; kf.mDescriptors = cv::Mat(n, 32, CV_8U, (void*)d);
    kf.mvuRight.assign(n, -1.0f); kf.mvDepth.assign(n, -1.0f); kf.mvpMapPoints.assign(n, (MapPoint*)NULL);
    pyramid(kf, 8, 1.2f);
    kf.mvLevelSigma2.assign(8, 1.0f); kf.mvInvLevelSigma2.assign(8, 1.0f);
    for (int i = 1; i < 8; i++) kf.mvLevelSigma2[i] = kf.mvScaleFactors[i] * kf.mvScaleFactors[i];
    for (int i = 0; i < 8; i++) kf.mvInvLevelSigma2[i] = 1.0f / kf.mvLevelSigma2[i];
    kf.mnMinX = (int)bounds4[0]; kf.mnMaxX = (int)bounds4[1]; kf.mnMinY = (int)bounds4[2]; kf.mnMaxY = (int)bounds4[3];
    kf.fx = cam4[0]; kf.fy = cam4[1]; kf.cx = cam4[2]; kf.cy = cam4[3];
    kf.Tcw = cv::Mat(4, 4, CV_32F, (void*)T).clone();
    kf.grid.build(kf.mvKeysUn, bounds4);
    KeyFrame other; other.mnId = 2; other.N = 1; other.mvKeysUn.resize(1); other.mvuRight.assign(1, -1.0f); other.mvpMapPoints.assign(1, (MapPoint*)NULL);
    other.mDescriptors = cv::Mat(1, 32, CV_8U); pyramid(other, 8, 1.2f);
    other.Tcw = cv::Mat(4, 4, CV_32F); for (int r = 0; r < 4; r++) other.Tcw.at<float>(r, r) = 1.0f;
    std::vector<TestPoint*> held(n, (TestPoint*)NULL), pts(n_mp, (TestPoint*)NULL);
    const cv::Mat zero3 = vec3(0, 0, 0), zdesc = cv::Mat(1, 32, CV_8U);
    for (int i = 0; i < n; i++) if (held_state[i]) {
        held[i] = new TestPoint(vec3(0, 0, 1), &kf, &map);
        held[i]->setup(zero3, kf.mDescriptors.row(i), 0.f, 1e30f, held_nobs[i], held_state[i] == 2);
        held[i]->observe(&kf, i);
        kf.mvpMapPoints[i] = held[i];
    }
    std::vector<MapPoint*> vp(n_mp, (MapPoint*)NULL);
    for (int m = 0; m < n_mp; m++) if (mp_state[m]) {
        pts[m] = new TestPoint(vec3(mp_pos[3 * m], mp_pos[3 * m + 1], mp_pos[3 * m + 2]), &other, &map);
        pts[m]->setup(vec3(mp_normal[3 * m], mp_normal[3 * m + 1], mp_normal[3 * m + 2]), cv::Mat(1, 32, CV_8U, (void*)(mp_desc + 32 * (size_t)m)),
                      mp_minmax[2 * m], mp_minmax[2 * m + 1], mp_nobs[m], mp_state[m] == 2);
        if (mp_state[m] == 3) pts[m]->observe(&kf, 0);                                             // IsInKeyFrame(pKF)
        vp[m] = pts[m];
    }
    ORBmatcher matcher(0.6f, true);
    const int nf = matcher.Fuse(&kf, vp, th);
    for (int i = 0; i < n; i++) {
        slot[i] = -1;
        MapPoint* p = kf.mvpMapPoints[i];
        if (p) {
            for (int j = 0; j < n && slot[i] < 0; j++) if (p == held[j]) slot[i] = 1000000 + j;
            for (int m = 0; m < n_mp && slot[i] < 0; m++) if (p == pts[m]) slot[i] = m;
        }
        held_bad[i] = held[i] ? (held[i]->isBad() ? 1 : 0) : 0;
    }
    for (int m = 0; m < n_mp; m++) { mp_bad[m] = pts[m] ? (pts[m]->isBad() ? 1 : 0) : 0; mp_nobs_out[m] = pts[m] ? pts[m]->Observations() : 0; }
    for (auto p : held) delete p;
    for (auto p : pts) delete p;
    return nf;
}

}  // extern "C"
