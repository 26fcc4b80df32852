// oracle/matchshim/slam_types.h -- TEST INFRASTRUCTURE, force-included (-include) in front of the reference's src/ORBmatcher.cc.
// include/ORBmatcher.h:29-31 pulls MapPoint.h / KeyFrame.h / Frame.h, whose include closure reaches Eigen, g2o, Pangolin and full OpenCV -
// none of which exist in this image.  Defining their include guards here turns those three headers into no-ops, and the plain-data
// stand-ins below give ORBmatcher.cc exactly the members it reads and writes (names, types and default arguments as declared in
// include/MapPoint.h, include/KeyFrame.h, include/Frame.h).  The matcher's own logic is therefore the reference's, statement for statement;
// what is restated here is only the glue around it: Frame::GetFeaturesInArea (src/Frame.cc:280-330, through oracle/frame_oracle.cpp, itself
// checked against a brute-force definition) and the MapPoint getters.  Every grid query the matcher makes is recorded (g_ref_trace), so
// tests can hand the product path the very same queries.
#pragma once
// REF_REAL_MAPPOINT (oracle/_ref/libref_mappoint.so): the reference's own include/MapPoint.h + src/MapPoint.cc are compiled as well, so only KeyFrame,
// Frame and Map are stand-ins and the MapPoint below is left out
#ifndef REF_REAL_MAPPOINT
#define MAPPOINT_H
#else
#define MAP_H
#endif
#define KEYFRAME_H
#define FRAME_H
#include <algorithm>
#include <climits>
#include <cmath>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include <opencv2/core/core.hpp>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "../oracle.h"

using namespace std;                 // the real closure gets this from Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:36

struct RefTrace { std::vector<float> xyr; std::vector<int> levels; std::vector<int> mp; int lastPredicted = 0; };
extern RefTrace g_ref_trace;

namespace ORB_SLAM2 {

class KeyFrame;
class Frame;

#ifndef REF_REAL_MAPPOINT
class MapPoint {
public:
    // tracking variables (include/MapPoint.h:95-101)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = -1;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0;
    // stand-in state
    int id = -1; bool bad = false; int nObs = 0;
    cv::Mat worldPos, normal, descriptor;
    float minDistance = 0, maxDistance = 1e30f;
    std::map<KeyFrame*, size_t> observations;
    MapPoint* replaced = nullptr;

    bool isBad() { return bad; }
    int Observations() { return nObs; }
    cv::Mat GetWorldPos() { return worldPos.clone(); }
    cv::Mat GetNormal() { return normal.clone(); }
    cv::Mat GetDescriptor() {
        if (!g_ref_trace.mp.empty() && g_ref_trace.mp.back() == -1) g_ref_trace.mp.back() = id;      // the query just made belongs to this point
        return descriptor.clone();
    }
    float GetMinDistanceInvariance() { return 0.8f * minDistance; }                                   // src/MapPoint.cc:391-401
    float GetMaxDistanceInvariance() { return 1.2f * maxDistance; }
    template <class T> int PredictScale(const float& currentDist, T* p) {                             // src/MapPoint.cc:403-435
        const float ratio = maxDistance / currentDist;
        int nScale = (int)ceil(log(ratio) / p->mfLogScaleFactor);
        if (nScale < 0) nScale = 0; else if (nScale >= p->mnScaleLevels) nScale = p->mnScaleLevels - 1;
        g_ref_trace.lastPredicted = nScale;                                                           // the KeyFrame queries below record it
        return nScale;
    }
    bool IsInKeyFrame(KeyFrame* pKF) { return observations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return observations.count(pKF) ? (int)observations[pKF] : -1; }
    void AddObservation(KeyFrame* pKF, size_t idx) { if (!observations.count(pKF)) { observations[pKF] = idx; nObs++; } }
    void Replace(MapPoint* pMP) { replaced = pMP; bad = true; }
};

#else
class MapPoint;
class Map {
public:
    std::mutex mMutexPointCreation;
    std::vector<MapPoint*> erased;
    void EraseMapPoint(MapPoint* p) { erased.push_back(p); }
};
#endif

// shared by the Frame and KeyFrame stand-ins: keypoints, descriptors, the 64x48 grid behind GetFeaturesInArea
struct RefGrid {
    std::vector<cv::KeyPoint> keysUn;
    float bounds[4] = {0, 0, 0, 0};
    std::vector<int32_t> cellStart, cellItems;
    void build(const std::vector<cv::KeyPoint>& k, const float* b4) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(oracle_keypoint), "keypoint layout");
        keysUn = k;
        for (int i = 0; i < 4; i++) bounds[i] = b4[i];
        cellStart.assign(64 * 48 + 1, 0); cellItems.assign(k.size() + 1, 0);
        oracle_assign_grid((const oracle_keypoint*)keysUn.data(), (int)keysUn.size(), bounds, cellStart.data(), cellItems.data());
    }
    std::vector<size_t> query(float x, float y, float r, int minLevel, int maxLevel, bool keyframe = false) const {
        std::vector<int32_t> out(keysUn.size() + 1);
        const int n = (keyframe ? oracle_keyframe_features_in_area : oracle_features_in_area)((const oracle_keypoint*)keysUn.data(), cellStart.data(),
                                                                                             cellItems.data(), bounds, x, y, r, minLevel, maxLevel, out.data(),
                                                                                             (int)keysUn.size());
        return std::vector<size_t>(out.begin(), out.begin() + std::min(n, (int)keysUn.size()));
    }
};

class Frame {
public:
    // include/Frame.h: calibration and bounds are static members there; one frame geometry per test is enough here as well
    static float fx, fy, cx, cy, invfx, invfy;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    float mbf = 0, mb = 0;
    int N = 0;
    long unsigned int mnId = 0;
    cv::Mat mOw;
    cv::Mat GetCameraCenter() { return mOw.clone(); }
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    int mnScaleLevels = 8;
    float mfScaleFactor = 1.2f, mfLogScaleFactor = std::log(1.2f);
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    cv::Mat mTcw;
    RefGrid grid;

    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
        g_ref_trace.xyr.push_back(x); g_ref_trace.xyr.push_back(y); g_ref_trace.xyr.push_back(r);
        g_ref_trace.levels.push_back(minLevel); g_ref_trace.levels.push_back(maxLevel);
        g_ref_trace.mp.push_back(-1);
        return grid.query(x, y, r, minLevel, maxLevel);
    }
};

class KeyFrame {
public:
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    int mnScaleLevels = 8;
    float mfScaleFactor = 1.2f, mfLogScaleFactor = std::log(1.2f);
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Tcw;
    RefGrid grid;

    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    long unsigned int mnId = 0;
    long int mnFrameId = 0;
    bool bad = false;
    bool isBad() { return bad; }
    void EraseMapPointMatch(const size_t& idx) { mvpMapPoints[idx] = nullptr; }
    void EraseMapPointMatch(MapPoint* pMP) { for (auto& p : mvpMapPoints) if (p == pMP) p = nullptr; }
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }
#ifndef REF_REAL_MAPPOINT
    std::set<MapPoint*> GetMapPoints() {
        std::set<MapPoint*> s;
        for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
        return s;
    }
#else
    std::set<MapPoint*> GetMapPoints();                            // defined where MapPoint is complete (oracle/ref_mappoint_wrap.cpp)
#endif
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    cv::Mat GetCameraCenter() { return -GetRotation().t() * GetTranslation(); }
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    // src/KeyFrame.cc:416-458: no level window; every caller filters kpLevel to [predicted - 1, predicted] afterwards and calls
    // PredictScale right before, so the trace stores that window with the query
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const {
        g_ref_trace.xyr.push_back(x); g_ref_trace.xyr.push_back(y); g_ref_trace.xyr.push_back(r);
        g_ref_trace.levels.push_back(g_ref_trace.lastPredicted - 1); g_ref_trace.levels.push_back(g_ref_trace.lastPredicted);
        g_ref_trace.mp.push_back(-1);
        return grid.query(x, y, r, -1, -1, true);                  // from the int mnMinX / mnMinY (src/KeyFrame.cc:677-689)
    }
};

}  // namespace ORB_SLAM2
