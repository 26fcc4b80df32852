// oracle/mapshim/slam_types_map.h -- TEST INFRASTRUCTURE, force-included (-include) in front of the reference's src/Map.cc so that Map::Save / Map::Load
// (src/Map.cc:219-533) compile UNMODIFIED from /root/reference into oracle/_ref/libref_map.so.  include/Map.h pulls MapPoint.h, KeyFrame.h, MapAruco.h,
// Converter.h, SystemSetting.h and InitKeyFrame.h, whose closure reaches Eigen, g2o, DBoW2 and full OpenCV; their include guards are pre-defined here and plain
// data stand-ins supply exactly the members Map.cc touches.  What is the reference's own: every byte Map::Save writes and every statement Map::Load runs
// (the file layout).  What is restated here: the quaternion <-> rotation conversions of Converter (Eigen's algorithms, src/Converter.cc:92-103, 150-162) and
// the data holders.  Pins orb_slam2_aruco_b200/mapfile.py (tests/test_mapfile_vs_ref.py, tests/golden/map_ref.bin).
#pragma once
#define MAPPOINT_H
#define KEYFRAME_H
#define MAPARUCO_H
#define CONVERTER_H
#define SYSTEMSETTING_H
#define INITKEYFRAME_H
#include <climits>
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>
#include <opencv2/core/core.hpp>

using namespace std;

namespace ORB_SLAM2 {

class Map;
class KeyFrame;
class KeyFrameDatabase;

class SystemSetting {};

class InitKeyFrame {                                  // include/InitKeyFrame.h: the members Map::LoadKeyFrame fills
public:
    explicit InitKeyFrame(SystemSetting&) : nId(0), TimeStamp(0), N(0), undistorted(0), gridded(0) {}
    long unsigned int nId;
    double TimeStamp;
    int N;
    std::vector<cv::KeyPoint> vKps;
    cv::Mat Descriptors;
    std::vector<float> vRight, vDepth;
    int undistorted, gridded;
    void UndistortKeyPoints() { undistorted++; }      // recomputation half of Map::Load: the product's rebuild() (tests/test_mapfile.py)
    void AssignFeaturesToGrid() { gridded++; }
};

class MapPoint {
public:
    MapPoint(const cv::Mat& Pos, Map*) : mnId(0), refKF(nullptr), distinctive(0), normal_depth(0) { SetWorldPos(Pos); }
    long unsigned int mnId;
    cv::Mat mWorldPos;
    KeyFrame* refKF;
    std::vector<std::pair<KeyFrame*, size_t> > obs;
    int distinctive, normal_depth;
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    void SetWorldPos(const cv::Mat& Pos) { mWorldPos = Pos.clone(); }
    void AddObservation(KeyFrame* pKF, size_t idx) {                  // src/MapPoint.cc:110-124: a keyframe observes a point once
        for (size_t i = 0; i < obs.size(); i++) if (obs[i].first == pKF) return;
        obs.push_back(std::make_pair(pKF, idx));
    }
    KeyFrame* GetReferenceKeyFrame() { return refKF; }
    void SetReferenceKeyFrame(KeyFrame* kf) { refKF = kf; }
    void ComputeDistinctiveDescriptors() { distinctive++; }
    void UpdateNormalAndDepth() { normal_depth++; }
};

class KeyFrame {
public:
    KeyFrame() : mnId(0), mTimeStamp(0), N(0), parent(nullptr), bow(0) {}
    KeyFrame(InitKeyFrame& ikf, Map*, KeyFrameDatabase*, std::vector<MapPoint*>& vpMapPoints)
        : mnId(ikf.nId), mTimeStamp(ikf.TimeStamp), N(ikf.N), mvKeys(ikf.vKps), mDescriptors(ikf.Descriptors.clone()), mvpMapPoints(vpMapPoints),
          parent(nullptr), bow(0) {}
    long unsigned int mnId;
    double mTimeStamp;
    int N;
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Tcw;
    KeyFrame* parent;
    std::vector<std::pair<KeyFrame*, int> > connections;              // in GetConnectedKeyFrames() order
    int bow;
    cv::Mat GetPose() { return Tcw.clone(); }
    void SetPose(const cv::Mat& T) { Tcw = T.clone(); }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    KeyFrame* GetParent() { return parent; }
    void ChangeParent(KeyFrame* p) { parent = p; }
    std::set<KeyFrame*> GetConnectedKeyFrames() { std::set<KeyFrame*> s; for (size_t i = 0; i < connections.size(); i++) s.insert(connections[i].first); return s; }
    int GetWeight(KeyFrame* kf) { for (size_t i = 0; i < connections.size(); i++) if (connections[i].first == kf) return connections[i].second; return 0; }
    void AddConnection(KeyFrame* kf, const int& weight) { connections.push_back(std::make_pair(kf, weight)); }
    void ComputeBoW() { bow++; }
    bool isBad() { return false; }
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetCameraCenter() { return -GetRotation().t() * Tcw.rowRange(0, 3).col(3); }
};

class MapAruco {                                      // only named by the marker bookkeeping members of Map.cc, never reached by Save / Load
public:
    int GetMapArucoID() { return 0; }
    std::map<KeyFrame*, size_t> GetObservations() { return std::map<KeyFrame*, size_t>(); }
    void UpdateTcmByKF(KeyFrame*, size_t) {}
    void SetRtwmByKeyFrame(const cv::Mat&, const cv::Mat&) {}
};

// src/Converter.cc:150-162 / 92-103 go through Eigen::Quaterniond; the same two algorithms (Eigen/src/Geometry/Quaternion.h: quaternionbase_assign_impl
// for a 3 x 3 matrix, QuaternionBase::toRotationMatrix) in double, results narrowed to float like the reference does
class Converter {
public:
    static std::vector<float> toQuaternion(const cv::Mat& M) {
        double m[3][3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = M.at<float>(i, j);
        double q[4];                                   // x y z w
        double t = m[0][0] + m[1][1] + m[2][2];
        if (t > 0.0) {
            t = std::sqrt(t + 1.0);
            q[3] = 0.5 * t;
            t = 0.5 / t;
            q[0] = (m[2][1] - m[1][2]) * t; q[1] = (m[0][2] - m[2][0]) * t; q[2] = (m[1][0] - m[0][1]) * t;
        } else {
            int i = 0;
            if (m[1][1] > m[0][0]) i = 1;
            if (m[2][2] > m[i][i]) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
            q[i] = 0.5 * t;
            t = 0.5 / t;
            q[3] = (m[k][j] - m[j][k]) * t; q[j] = (m[j][i] + m[i][j]) * t; q[k] = (m[k][i] + m[i][k]) * t;
        }
        std::vector<float> v(4);
        for (int i = 0; i < 4; i++) v[i] = (float)q[i];
        return v;
    }
    static cv::Mat toCvMat(const std::vector<float>& v) {
        const double x = v[0], y = v[1], z = v[2], w = v[3];
        const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
                     tyz = tz * y, tzz = tz * z;
        cv::Mat M(3, 3, CV_32F);
        M.at<float>(0, 0) = (float)(1 - (tyy + tzz)); M.at<float>(0, 1) = (float)(txy - twz); M.at<float>(0, 2) = (float)(txz + twy);
        M.at<float>(1, 0) = (float)(txy + twz); M.at<float>(1, 1) = (float)(1 - (txx + tzz)); M.at<float>(1, 2) = (float)(tyz - twx);
        M.at<float>(2, 0) = (float)(txz - twy); M.at<float>(2, 1) = (float)(tyz + twx); M.at<float>(2, 2) = (float)(1 - (txx + tyy));
        return M;
    }
};

}  // namespace ORB_SLAM2
