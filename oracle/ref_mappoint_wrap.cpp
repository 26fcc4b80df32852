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

}  // extern "C"
