// b200slam_adapters.hpp -- source-compatible C++ front-end classes on top of the C-ABI (include/b200slam.h).
//
// The reference has no FFI layer: src/Frame.cc and src/Tracking.cc call three C++ classes directly.
// These adapters keep the reference's class names, constructor arguments and call signatures for the hot path, so
// that Frame/Tracking compile against them unchanged, and marshal to libb200slam.so:
//
//   ORB_SLAM2::ORBextractor   reference include/ORBextractor.h:45-111  (operator() at :59-61, getters :63-83)
//   ORB_SLAM2::ORBmatcher     reference include/ORBmatcher.h:38-104    (DescriptorDistance :44, constants :87-89; the
//                             brute-force SearchByBoW core; the map-point glue of the other Search* stays host code)
//   aruco::MarkerDetector     reference Thirdparty/aruco/aruco/markerdetector.h:58-410 (detect :276-278,
//                             setDictionary :337, setDetectionMode :258, Params::setCornerRefinementMethod :129)
//
// OpenCV types: when <opencv2/core.hpp> is on the include path (a real integration) cv::Mat / cv::KeyPoint are used
// as is.  Define B200SLAM_NO_OPENCV to get the minimal stand-ins below (used by tests/cpp/adapter_smoke.cpp, which
// must build in an image without OpenCV).  Header only; link with -lb200slam.
#pragma once
#include <cassert>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "b200slam.h"

#ifndef B200SLAM_NO_OPENCV
#include <opencv2/core/core.hpp>
#else
namespace cv {
struct Point2f { float x, y; Point2f(float _x = 0, float _y = 0) : x(_x), y(_y) {} };
struct Point { int x, y; Point(int _x = 0, int _y = 0) : x(_x), y(_y) {} };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
static const int CV_8UC1_ = 0;
#define CV_8U 0
#define CV_8UC1 0
// 8-bit single channel matrix owning or borrowing its pixels (enough for marshaling)
class Mat {
public:
    int rows, cols; size_t step; unsigned char* data; std::vector<unsigned char> own;
    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int, void* p, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data((unsigned char*)p) {}
    void create(int r, int c, int) { rows = r; cols = c; step = (size_t)c; own.assign((size_t)r * c, 0); data = own.data(); }
    void release() { rows = cols = 0; step = 0; data = nullptr; own.clear(); }
    bool empty() const { return !data || rows * cols == 0; }
    int type() const { return CV_8UC1; }
    unsigned char* ptr(int y = 0) { return data + (size_t)y * step; }
    const unsigned char* ptr(int y = 0) const { return data + (size_t)y * step; }
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace cv
#endif

namespace b200slam_detail {
inline void check(int rc) { if (rc < 0) throw std::runtime_error(std::string("b200slam: ") + b200_last_error()); }
static_assert(sizeof(b200_keypoint) == 28, "b200_keypoint must mirror cv::KeyPoint");
}

// DBoW2 containers the reference fills in Frame::ComputeBoW (Thirdparty/DBoW2/DBoW2/BowVector.h:52, FeatureVector.h:21-22)
namespace DBoW2 {
typedef unsigned int WordId;
typedef unsigned int NodeId;
typedef double WordValue;
typedef std::map<WordId, WordValue> BowVector;
typedef std::map<NodeId, std::vector<unsigned int> > FeatureVector;
}  // namespace DBoW2

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // reference: ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
        : nfeatures_(nfeatures), scale_(scaleFactor), nlevels_(nlevels), ini_(iniThFAST), min_(minThFAST), device_(device), h_(nullptr), w_(0), ht_(0) {}
    ~ORBextractor() { if (h_) b200_orb_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image; mask is ignored, exactly like the reference (ORBextractor.h:58)
    void operator()(cv::InputArray image_, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors_) {
#ifndef B200SLAM_NO_OPENCV
        cv::Mat image = image_.getMat();
#else
        const cv::Mat& image = image_;
#endif
        if (image.empty()) return;                                    // ORBextractor.cc:1046
        assert(image.type() == CV_8UC1);                              // ORBextractor.cc:1050
        ensure(image.cols, image.rows);
        const int cap = b200_orb_max_keypoints(h_);
        kps_.resize(cap); desc_.resize((size_t)cap * 32);
        int32_t n = 0;
        b200slam_detail::check(b200_orb_extract_host(h_, image.data, 1, image.cols, image.rows, (int64_t)image.step, (int64_t)image.step * image.rows,
                                                     kps_.data(), desc_.data(), &n));
        keypoints.clear();
        keypoints.resize(n);
        static_assert(sizeof(cv::KeyPoint) == sizeof(b200_keypoint), "cv::KeyPoint layout");
        if (n) std::memcpy((void*)keypoints.data(), kps_.data(), (size_t)n * sizeof(b200_keypoint));
#ifndef B200SLAM_NO_OPENCV
        if (n == 0) { descriptors_.release(); return; }
        descriptors_.create(n, 32, CV_8U);
        cv::Mat d = descriptors_.getMat();
#else
        cv::Mat& d = descriptors_;
        if (n == 0) { d.release(); return; }
        d.create(n, 32, CV_8U);
#endif
        for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), &desc_[(size_t)i * 32], 32);
        if (fillImagePyramid) fillPyramid();
    }

    // public member of the reference (include/ORBextractor.h:85): the level images of the LAST call.  Only the stereo path reads it (src/Frame.cc:456), so it
    // is filled on request: set fillImagePyramid (one device-to-host copy per level and call) or call ComputePyramidMember() after operator().  Every entry is
    // the level's (w_l + 38) x (h_l + 38) buffer with the 19-px REFLECT_101 border, and mvImagePyramidOffset = 19: the reference's entries are ROIs at (19, 19)
    // of exactly such buffers (src/ORBextractor.cc:1107-1132); with real OpenCV use mvImagePyramid[l](cv::Rect(19, 19, w_l, h_l)).
    std::vector<cv::Mat> mvImagePyramid;
    bool fillImagePyramid = false;
    static const int mvImagePyramidOffset = 19;
    void ComputePyramidMember() { fillPyramid(); }

    int inline GetLevels() { return nlevels_; }
    float inline GetScaleFactor() { return scale_; }
    std::vector<float> inline GetScaleFactors() { return table(0); }
    std::vector<float> inline GetInverseScaleFactors() { return table(1); }
    std::vector<float> inline GetScaleSigmaSquares() { return table(2); }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return table(3); }

    // reference public member mvImagePyramid (ORBextractor.h:85): filled on demand, level images incl. the 19-px border
    std::vector<std::vector<unsigned char> > ImagePyramid(std::vector<int>* widths = nullptr, std::vector<int>* heights = nullptr) {
        std::vector<std::vector<unsigned char> > out(nlevels_);
        for (int l = 0; l < nlevels_; l++) {
            out[l].resize((size_t)(w_ + 38) * (ht_ + 38));
            int wl = 0, hl = 0;
            b200slam_detail::check(b200_orb_get_pyramid(h_, 0, l, out[l].data(), &wl, &hl));
            out[l].resize((size_t)(wl + 38) * (hl + 38));
            if (widths) widths->push_back(wl);
            if (heights) heights->push_back(hl);
        }
        return out;
    }

private:
    void fillPyramid() {
        mvImagePyramid.resize(nlevels_);
        std::vector<unsigned char> buf((size_t)(w_ + 38) * (ht_ + 38));
        for (int l = 0; l < nlevels_; l++) {
            int wl = 0, hl = 0;
            b200slam_detail::check(b200_orb_get_pyramid(h_, 0, l, buf.data(), &wl, &hl));
            mvImagePyramid[l].create(hl + 38, wl + 38, CV_8U);
            for (int y = 0; y < hl + 38; y++) std::memcpy(mvImagePyramid[l].ptr(y), &buf[(size_t)y * (wl + 38)], (size_t)wl + 38);
        }
    }
    void ensure(int w, int h) {
        if (h_ && w <= w_ && h <= ht_) return;
        if (h_) { b200_orb_destroy(h_); h_ = nullptr; }
        w_ = w > w_ ? w : w_; ht_ = h > ht_ ? h : ht_;
        b200slam_detail::check(b200_orb_create(&h_, nfeatures_, scale_, nlevels_, ini_, min_, w_, ht_, 1, device_));
    }
    std::vector<float> table(int which) {
        ensure(w_ ? w_ : 64, ht_ ? ht_ : 64);
        std::vector<float> t[4];
        for (auto& v : t) v.resize(nlevels_);
        b200slam_detail::check(b200_orb_get_level_info(h_, nullptr, t[0].data(), t[1].data(), t[2].data(), t[3].data(), nullptr));
        return t[which];
    }
    int nfeatures_; float scale_; int nlevels_, ini_, min_, device_;
    b200_orb_t h_; int w_, ht_;
    std::vector<b200_keypoint> kps_; std::vector<uint8_t> desc_;
};

class ORBmatcher {
public:
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;      // ORBmatcher.cc:37-39

    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), device_(device) {}

    // Computes the Hamming distance between two ORB descriptors (ORBmatcher.h:44)
    // One pair stays on the host (a popcount over four 64-bit words): the reference calls it O(N^2) times per map point (src/MapPoint.cc:310);
    // the batched forms are b200_hamming_matrix_host and ComputeDistinctiveDescriptors below.
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        const unsigned char* pa = a.ptr(0);
        const unsigned char* pb = b.ptr(0);
        int d = 0;
        for (int i = 0; i < 4; i++) {
            uint64_t x, y;
            std::memcpy(&x, pa + 8 * i, 8); std::memcpy(&y, pb + 8 * i, 8);
            d += __builtin_popcountll(x ^ y);
        }
        return d;
    }

    // Brute-force core of SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (ORBmatcher.h:55): every keyframe feature that
    // owns a good MapPoint, all features in one vocabulary node.  matches[i] = keyframe index matched to frame keypoint i or -1
    // (the caller maps indices back to MapPoint*); returns nmatches.
    int SearchByBoW(const cv::Mat& kfDescriptors, const std::vector<cv::KeyPoint>& kfKeysUn,
                    const cv::Mat& fDescriptors, const std::vector<cv::KeyPoint>& fKeys, std::vector<int>& matches) {
        const int nkf = kfDescriptors.rows, nf = fDescriptors.rows;
        std::vector<uint8_t> kd((size_t)nkf * 32), fd((size_t)nf * 32);
        std::vector<float> ka(nkf), fa(nf);
        for (int i = 0; i < nkf; i++) { std::memcpy(&kd[(size_t)i * 32], kfDescriptors.ptr(i), 32); ka[i] = kfKeysUn[i].angle; }
        for (int i = 0; i < nf; i++) { std::memcpy(&fd[(size_t)i * 32], fDescriptors.ptr(i), 32); fa[i] = fKeys[i].angle; }
        matches.assign(nf, -1);
        int32_t n_frame = nf, nm = 0;
        b200slam_detail::check(b200_match_bf_host(kd.data(), ka.data(), nkf, fd.data(), fa.data(), &n_frame, 1, nf, mfNNratio, TH_LOW,
                                                  mbCheckOrientation ? 1 : 0, HISTO_LENGTH / 360.0f, matches.data(), &nm, device_));
        return nm;
    }

    // One projected map point of SearchByProjection: where it lands in the frame, how far to look, in which pyramid levels, and what it
    // looks like.  The projection itself (pose, MapPoint::mTrackProjX / GetWorldPos, RadiusByViewingCos, th * mvScaleFactors[level]) is
    // the reference's host glue and stays with the caller.
    struct ProjectedPoint { float x, y, radius; int minLevel, maxLevel; const unsigned char* descriptor; float angle; bool observed; };

    // SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th) (mode 0, ORBmatcher.h:48, ORBmatcher.cc:45-129) and
    // SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (mode 1, ORBmatcher.h:52, ORBmatcher.cc:1332-1474) on
    // projected points; mode 1 with thHigh = ORBdist is SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1476-1603).
    // occupied[i] = "F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0" (in/out);
    // assign[i] = index into `points` now held by frame keypoint i, or -1; returns nmatches.
    int SearchByProjection(const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& descriptors, const float bounds[4], std::vector<unsigned char>& occupied,
                           const std::vector<ProjectedPoint>& points, int mode, std::vector<int>& assign, int thHigh = TH_HIGH) {
        const int n = (int)keysUn.size(), nq = (int)points.size();
        std::vector<uint8_t> d((size_t)n * 32), qd((size_t)nq * 32), qo(nq);
        for (int i = 0; i < n; i++) std::memcpy(&d[(size_t)i * 32], descriptors.ptr(i), 32);
        std::vector<float> q3((size_t)nq * 3), qa(nq);
        std::vector<int32_t> lv((size_t)nq * 2);
        for (int q = 0; q < nq; q++) {
            q3[3 * q] = points[q].x; q3[3 * q + 1] = points[q].y; q3[3 * q + 2] = points[q].radius;
            lv[2 * q] = points[q].minLevel; lv[2 * q + 1] = points[q].maxLevel;
            std::memcpy(&qd[(size_t)q * 32], points[q].descriptor, 32); qa[q] = points[q].angle; qo[q] = points[q].observed ? 1 : 0;
        }
        occupied.resize(n, 0);
        assign.assign(n, -1);
        static_assert(sizeof(cv::KeyPoint) == sizeof(b200_keypoint), "cv::KeyPoint layout");
        const int nm = b200_match_by_projection_host((const b200_keypoint*)keysUn.data(), d.data(), n, bounds, occupied.data(), q3.data(), lv.data(), qd.data(),
                                                     qa.data(), qo.data(), nq, mode, mfNNratio, mbCheckOrientation ? 1 : 0, thHigh, assign.data(), device_);
        b200slam_detail::check(nm);
        return nm;
    }

    // int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint*> &vpMapPointMatches) (ORBmatcher.h:55, ORBmatcher.cc:159-292) on what it
    // reads: pKF->mFeatVec / F.mFeatVec, the descriptors and undistorted keypoints of both, and kfGood[i] = "vpMapPointsKF[i] && !isBad()".
    // matches[idxF] = keyframe index whose MapPoint the frame keypoint receives, or -1; returns nmatches.
    int SearchByBoW(const cv::Mat& kfDescriptors, const std::vector<cv::KeyPoint>& kfKeysUn, const std::vector<bool>& kfGood,
                    const DBoW2::FeatureVector& kfFeatVec, const cv::Mat& fDescriptors, const std::vector<cv::KeyPoint>& fKeysUn,
                    const DBoW2::FeatureVector& fFeatVec, std::vector<int>& matches) {
        return byBoW(0, kfDescriptors, kfKeysUn, &kfGood, kfFeatVec, fDescriptors, fKeysUn, nullptr, fFeatVec, matches);
    }

    // int SearchByBoW(KeyFrame *pKF1, KeyFrame* pKF2, std::vector<MapPoint*> &vpMatches12) (ORBmatcher.h:56, ORBmatcher.cc:526-659) with the
    // two feature vectors; good1 / good2 = the keyframes' map points that exist and are not bad.  matches12[idx1] = index in KF2 or -1.
    int SearchByBoW_KF(const cv::Mat& desc1, const std::vector<cv::KeyPoint>& keysUn1, const std::vector<bool>& good1, const DBoW2::FeatureVector& featVec1,
                       const cv::Mat& desc2, const std::vector<cv::KeyPoint>& keysUn2, const std::vector<bool>& good2, const DBoW2::FeatureVector& featVec2,
                       std::vector<int>& matches12) {
        return byBoW(1, desc1, keysUn1, &good1, featVec1, desc2, keysUn2, &good2, featVec2, matches12);
    }

    // Brute-force core of SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (ORBmatcher.h:56,
    // ORBmatcher.cc:526-659): strict bestDist1 < TH_LOW, histogram factor 1.0f / HISTO_LENGTH; matches12[idx1] = index in KF2 or -1.
    int SearchByBoW_KF(const cv::Mat& desc1, const std::vector<cv::KeyPoint>& keysUn1, const cv::Mat& desc2, const std::vector<cv::KeyPoint>& keysUn2,
                       std::vector<int>& matches12) {
        const int n1 = desc1.rows, n2 = desc2.rows;
        std::vector<uint8_t> d1((size_t)n1 * 32), d2((size_t)n2 * 32);
        std::vector<float> a1(n1), a2(n2);
        for (int i = 0; i < n1; i++) { std::memcpy(&d1[(size_t)i * 32], desc1.ptr(i), 32); a1[i] = keysUn1[i].angle; }
        for (int i = 0; i < n2; i++) { std::memcpy(&d2[(size_t)i * 32], desc2.ptr(i), 32); a2[i] = keysUn2[i].angle; }
        std::vector<int> m21(n2 > 0 ? n2 : 1, -1);
        int32_t n_frame = n2, nm = 0;
        b200slam_detail::check(b200_match_bf_host(d1.data(), a1.data(), n1, d2.data(), a2.data(), &n_frame, 1, n2, mfNNratio, TH_LOW - 1,
                                                  mbCheckOrientation ? 1 : 0, 1.0f / HISTO_LENGTH, m21.data(), &nm, device_));
        matches12.assign(n1, -1);
        for (int i2 = 0; i2 < n2; i2++) if (m21[i2] >= 0) matches12[m21[i2]] = i2;
        return nm;
    }

    // int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12,
    //                             int windowSize = 10) (ORBmatcher.h:63-64, ORBmatcher.cc:409-524) on what it reads from the two frames:
    // mvKeysUn + mDescriptors of both and the image bounds mnMinX, mnMaxX, mnMinY, mnMaxY (Frame.h:191-194) behind F2's grid.
    int SearchForInitialization(const std::vector<cv::KeyPoint>& keysUn1, const cv::Mat& desc1,
                                const std::vector<cv::KeyPoint>& keysUn2, const cv::Mat& desc2, const float bounds[4],
                                std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10) {
        const int n1 = (int)keysUn1.size(), n2 = (int)keysUn2.size();
        std::vector<uint8_t> d1((size_t)n1 * 32), d2((size_t)n2 * 32);
        for (int i = 0; i < n1; i++) std::memcpy(&d1[(size_t)i * 32], desc1.ptr(i), 32);
        for (int i = 0; i < n2; i++) std::memcpy(&d2[(size_t)i * 32], desc2.ptr(i), 32);
        std::vector<float> prev((size_t)n1 * 2);
        for (int i = 0; i < n1; i++) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
        vnMatches12.assign(n1, -1);
        static_assert(sizeof(cv::KeyPoint) == sizeof(b200_keypoint), "cv::KeyPoint layout");
        const int nm = b200_match_for_initialization_host((const b200_keypoint*)keysUn1.data(), d1.data(), n1, (const b200_keypoint*)keysUn2.data(), d2.data(), n2,
                                                          bounds, prev.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, vnMatches12.data(), device_);
        b200slam_detail::check(nm);
        for (int i = 0; i < n1; i++) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
        return nm;
    }

    // best / second-best over explicit candidate lists: the distance core of SearchByProjection / SearchForInitialization
    void MatchCandidates(const cv::Mat& queryDesc, const cv::Mat& trainDesc, const std::vector<int32_t>& candOfs, const std::vector<int32_t>& cand,
                         std::vector<int32_t>& bestIdx, std::vector<int32_t>& bestDist, std::vector<int32_t>& secondDist) {
        const int nq = queryDesc.rows, nt = trainDesc.rows;
        std::vector<uint8_t> q((size_t)nq * 32), t((size_t)nt * 32);
        for (int i = 0; i < nq; i++) std::memcpy(&q[(size_t)i * 32], queryDesc.ptr(i), 32);
        for (int i = 0; i < nt; i++) std::memcpy(&t[(size_t)i * 32], trainDesc.ptr(i), 32);
        bestIdx.resize(nq); bestDist.resize(nq); secondDist.resize(nq);
        b200slam_detail::check(b200_match_candidates_host(q.data(), nq, t.data(), nt, candOfs.data(), cand.data(), bestIdx.data(), bestDist.data(),
                                                          secondDist.data(), device_));
    }

    // void MapPoint::ComputeDistinctiveDescriptors() (MapPoint.h:75, MapPoint.cc:271-331) for many map points at once: vDescriptors[p] = the
    // descriptor rows of the good keyframes observing point p.  best[p] = BestIdx (-1: no observation, mDescriptor stays as it is).
    void ComputeDistinctiveDescriptors(const std::vector<std::vector<cv::Mat> >& vDescriptors, std::vector<int>& best) {
        std::vector<int32_t> ofs(1, 0);
        std::vector<uint8_t> rows;
        for (const auto& v : vDescriptors) {
            for (const cv::Mat& d : v) rows.insert(rows.end(), d.ptr(0), d.ptr(0) + 32);
            ofs.push_back((int32_t)(rows.size() / 32));
        }
        std::vector<int32_t> b(vDescriptors.size() + 1, -1);
        b200slam_detail::check(b200_distinctive_descriptors_host(rows.data(), ofs.data(), (int)vDescriptors.size(), b.data(), nullptr, device_));
        best.assign(b.begin(), b.begin() + vDescriptors.size());
    }

    // int SearchForTriangulation(KeyFrame *pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo)
    // (ORBmatcher.h:66-67, ORBmatcher.cc:661-829) for monocular keyframes on what it reads: undistorted keypoints, descriptors, "the feature already
    // has a MapPoint" flags and FeatureVectors of both keyframes, F12 (row-major 3x3), the epipole of camera 1 in image 2 (:668-674) and
    // pKF2->mvScaleFactors / mvLevelSigma2.  Returns nmatches and fills vMatchedPairs like the reference (ascending first index).
    int SearchForTriangulation(const std::vector<cv::KeyPoint>& keysUn1, const cv::Mat& desc1, const std::vector<bool>& hasMapPoint1, const DBoW2::FeatureVector& featVec1,
                               const std::vector<cv::KeyPoint>& keysUn2, const cv::Mat& desc2, const std::vector<bool>& hasMapPoint2, const DBoW2::FeatureVector& featVec2,
                               const float F12[9], float ex, float ey, const std::vector<float>& scaleFactors2, const std::vector<float>& levelSigma2_2,
                               std::vector<std::pair<size_t, size_t> >& vMatchedPairs) {
        const int n1 = desc1.rows, n2 = desc2.rows;
        std::vector<uint8_t> d1((size_t)n1 * 32), d2((size_t)n2 * 32);
        for (int i = 0; i < n1; i++) std::memcpy(&d1[(size_t)i * 32], desc1.ptr(i), 32);
        for (int i = 0; i < n2; i++) std::memcpy(&d2[(size_t)i * 32], desc2.ptr(i), 32);
        std::vector<int32_t> gq(1, 0), gc(1, 0), qi, ci;
        DBoW2::FeatureVector::const_iterator it1 = featVec1.begin(), it2 = featVec2.begin();
        while (it1 != featVec1.end() && it2 != featVec2.end()) {
            if (it1->first == it2->first) {
                for (unsigned int i : it1->second) if (!hasMapPoint1[i]) qi.push_back((int32_t)i);
                for (unsigned int i : it2->second) if (!hasMapPoint2[i]) ci.push_back((int32_t)i);
                gq.push_back((int32_t)qi.size()); gc.push_back((int32_t)ci.size());
                ++it1; ++it2;
            } else if (it1->first < it2->first) it1 = featVec1.lower_bound(it2->first);
            else it2 = featVec2.lower_bound(it1->first);
        }
        const float e2[2] = {ex, ey};
        std::vector<int32_t> m12((size_t)n1 + 1, -1);
        static_assert(sizeof(cv::KeyPoint) == sizeof(b200_keypoint), "cv::KeyPoint layout");
        const int nm = b200_match_for_triangulation_host((const b200_keypoint*)keysUn1.data(), d1.data(), n1, (const b200_keypoint*)keysUn2.data(), d2.data(), n2,
                                                         gq.data(), qi.data(), gc.data(), ci.data(), (int)gq.size() - 1, F12, e2, scaleFactors2.data(),
                                                         levelSigma2_2.data(), (int)scaleFactors2.size(), mbCheckOrientation ? 1 : 0, TH_LOW, m12.data(), device_);
        b200slam_detail::check(nm);
        vMatchedPairs.clear();
        vMatchedPairs.reserve(nm);
        for (int i = 0; i < n1; i++) if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
        return nm;
    }

    // The keyframe search inside Fuse(KeyFrame*, vpMapPoints, th) (ORBmatcher.h:76, ORBmatcher.cc:831-981), Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)
    // (ORBmatcher.h:79, :983-1104) and SearchBySim3 (ORBmatcher.h:70-71, :1106-1330).  Those functions keep their own statements up to
    // "const float radius = th*pKF->mvScaleFactors[nPredictedLevel]" (they are cv::Mat algebra on one point), push a RadiusQuery instead of calling
    // GetFeaturesInArea, and after ONE SearchInRadius call run their "if(bestDist<=TH_LOW) ..." statements over the answers in list order: the search
    // itself never depends on what earlier points did to the keyframe.  chi2 > 0 adds Fuse's reprojection gate (5.99, monocular; invLevelSigma2 =
    // pKF->mvInvLevelSigma2).  bestIdx[q] = -1 / bestDist[q] = 256 when no feature qualifies.
    struct RadiusQuery { float u, v, radius; int predictedLevel; const unsigned char* descriptor; };
    void SearchInRadius(const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& descriptors, const float bounds[4], const std::vector<RadiusQuery>& queries,
                        const std::vector<float>& invLevelSigma2, double chi2, std::vector<int>& bestIdx, std::vector<int>& bestDist) {
        const int n = (int)keysUn.size(), nq = (int)queries.size();
        std::vector<uint8_t> d((size_t)n * 32), qd((size_t)nq * 32);
        for (int i = 0; i < n; i++) std::memcpy(&d[(size_t)i * 32], descriptors.ptr(i), 32);
        std::vector<float> q3((size_t)nq * 3);
        std::vector<int32_t> ql(nq), bi((size_t)nq + 1, -1), bd((size_t)nq + 1, 256);
        for (int q = 0; q < nq; q++) {
            q3[3 * q] = queries[q].u; q3[3 * q + 1] = queries[q].v; q3[3 * q + 2] = queries[q].radius; ql[q] = queries[q].predictedLevel;
            std::memcpy(&qd[(size_t)q * 32], queries[q].descriptor, 32);
        }
        static_assert(sizeof(cv::KeyPoint) == sizeof(b200_keypoint), "cv::KeyPoint layout");
        b200slam_detail::check(b200_match_kf_radius_host((const b200_keypoint*)keysUn.data(), d.data(), n, bounds, q3.data(), ql.data(), qd.data(), nq,
                                                         invLevelSigma2.data(), (int)invLevelSigma2.size(), chi2, bi.data(), bd.data(), device_));
        bestIdx.assign(bi.begin(), bi.begin() + nq);
        bestDist.assign(bd.begin(), bd.begin() + nq);
    }

    // MapPoint::PredictScale (MapPoint.cc:403-435) as a threshold table for the device projection: thresholds[n] = the largest float ratio
    // mfMaxDistance / dist that still yields a level <= n with THIS host's std::log(float), by bisection over float bit patterns.
    static std::vector<float> PredictScaleThresholds(float scaleFactor, int nlevels) {
        const float logSf = std::log(scaleFactor);
        std::vector<float> thr(nlevels > 1 ? nlevels - 1 : 0);
        for (int n = 0; n + 1 < nlevels; n++) {
            uint32_t lo = 0x00800000u, hi = 0x7f000000u;
            while (hi - lo > 1) {
                const uint32_t mid = lo + (hi - lo) / 2;
                float r; std::memcpy(&r, &mid, 4);
                if ((int)std::ceil(std::log(r) / logSf) <= n) lo = mid; else hi = mid;
            }
            std::memcpy(&thr[n], &lo, 4);
        }
        return thr;
    }

    // Projection + search of a whole map point list in one device call (b200_kf_search_points_host): the loop bodies of Fuse (ORBmatcher.cc:846-955),
    // Fuse(Scw) (:1004-1082) and of each direction of SearchBySim3 (:1151-1217, :1231-1297) up to "if(bestDist<=TH_...)", evaluated with the
    // reference's cv::Mat CV_32F roundings.  R / t / Ow: rotation (row-major), translation and centre of the keyframe camera (for Scw the
    // decomposed Sim3, :302-307); sR / tt: NULL, or SearchBySim3's second transform (then Ow and the normals are unused).  skip: the reference
    // `continue`s on the point before projecting it.  Afterwards the caller runs the reference's outcome statements over (valid, bestIdx, bestDist)
    // in list order.  mfMinDistance / mfMaxDistance are the MapPoint members behind Get{Min,Max}DistanceInvariance() (MapPoint.cc:391-401).
    struct MapPointView { float pos[3]; float normal[3]; float minDistance, maxDistance; const unsigned char* descriptor; bool skip; };
    void SearchPoints(const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& descriptors, const float bounds[4], const float R[9], const float t[3], const float Ow[3],
                      const float* sR, const float* tt, const float cam4[4], const std::vector<MapPointView>& points, bool useNormals, float th,
                      const std::vector<float>& scaleFactors, const std::vector<float>& invLevelSigma2, double chi2,
                      std::vector<bool>& valid, std::vector<int>& bestIdx, std::vector<int>& bestDist) {
        const int n = (int)keysUn.size(), np = (int)points.size(), nl = (int)scaleFactors.size();
        std::vector<uint8_t> d((size_t)n * 32), qd((size_t)np * 32), sk(np), v((size_t)np + 1, 0);
        for (int i = 0; i < n; i++) std::memcpy(&d[(size_t)i * 32], descriptors.ptr(i), 32);
        std::vector<float> pos((size_t)np * 3), nrm((size_t)np * 3), mm((size_t)np * 2);
        for (int i = 0; i < np; i++) {
            for (int j = 0; j < 3; j++) { pos[3 * i + j] = points[i].pos[j]; nrm[3 * i + j] = points[i].normal[j]; }
            mm[2 * i] = points[i].minDistance; mm[2 * i + 1] = points[i].maxDistance; sk[i] = points[i].skip ? 1 : 0;
            std::memcpy(&qd[(size_t)i * 32], points[i].descriptor, 32);
        }
        const std::vector<float> thr = PredictScaleThresholds(nl > 1 ? scaleFactors[1] : 1.2f, nl);
        std::vector<int32_t> bi((size_t)np + 1, -1), bd((size_t)np + 1, 256);
        static_assert(sizeof(cv::KeyPoint) == sizeof(b200_keypoint), "cv::KeyPoint layout");
        b200slam_detail::check(b200_kf_search_points_host((const b200_keypoint*)keysUn.data(), d.data(), n, bounds, R, t, Ow, sR, tt, cam4, pos.data(),
                                                          useNormals ? nrm.data() : nullptr, mm.data(), qd.data(), sk.data(), np, th, scaleFactors.data(),
                                                          invLevelSigma2.data(), thr.data(), nl, chi2, v.data(), bi.data(), bd.data(), nullptr, nullptr, device_));
        valid.assign(np, false);
        for (int i = 0; i < np; i++) valid[i] = v[i] != 0;
        bestIdx.assign(bi.begin(), bi.begin() + np);
        bestDist.assign(bd.begin(), bd.begin() + np);
    }

    // int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, std::vector<MapPoint*> &vpMatched, int th)
    // (ORBmatcher.h:52, ORBmatcher.cc:294-407) on projected points: here a matched feature hides itself from every later point, so the queries are
    // replayed in order by the projection resolve kernel.  matchedBefore[i] = "vpMatched[i] != NULL"; assign[i] = index into `queries` newly
    // matched to keyframe feature i or -1 (the caller stores vpMatched[i] = that point); returns nmatches.
    int SearchByProjectionLoop(const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& descriptors, const float bounds[4], const std::vector<bool>& matchedBefore,
                               const std::vector<RadiusQuery>& queries, std::vector<int>& assign) {
        std::vector<ProjectedPoint> pts(queries.size());
        for (size_t q = 0; q < queries.size(); q++)
            pts[q] = ProjectedPoint{queries[q].u, queries[q].v, queries[q].radius, queries[q].predictedLevel - 1, queries[q].predictedLevel, queries[q].descriptor, 0.f, true};
        std::vector<unsigned char> occ(matchedBefore.size());
        for (size_t i = 0; i < occ.size(); i++) occ[i] = matchedBefore[i] ? 1 : 0;
        const bool ori = mbCheckOrientation;
        mbCheckOrientation = false;                                 // no rotation histogram in this member (ORBmatcher.cc:294-407)
        int nm;
        try { nm = SearchByProjection(keysUn, descriptors, bounds, occ, pts, 2 /* mode 1 over KeyFrame::GetFeaturesInArea's grid origin */, assign, TH_LOW); }
        catch (...) { mbCheckOrientation = ori; throw; }
        mbCheckOrientation = ori;
        return nm;
    }

protected:
    // the merge walk over common vocabulary nodes (ORBmatcher.cc:185-279 / 547-632) builds the groups of b200_match_by_bow_host
    int byBoW(int mode, const cv::Mat& desc1, const std::vector<cv::KeyPoint>& keys1, const std::vector<bool>* good1, const DBoW2::FeatureVector& fv1,
              const cv::Mat& desc2, const std::vector<cv::KeyPoint>& keys2, const std::vector<bool>* good2, const DBoW2::FeatureVector& fv2,
              std::vector<int>& out) {
        const int n1 = desc1.rows, n2 = desc2.rows;
        std::vector<uint8_t> d1((size_t)n1 * 32), d2((size_t)n2 * 32);
        std::vector<float> a1(n1), a2(n2);
        for (int i = 0; i < n1; i++) { std::memcpy(&d1[(size_t)i * 32], desc1.ptr(i), 32); a1[i] = keys1[i].angle; }
        for (int i = 0; i < n2; i++) { std::memcpy(&d2[(size_t)i * 32], desc2.ptr(i), 32); a2[i] = keys2[i].angle; }
        std::vector<int32_t> gq(1, 0), gc(1, 0), qi, ci;
        DBoW2::FeatureVector::const_iterator it1 = fv1.begin(), it2 = fv2.begin();
        while (it1 != fv1.end() && it2 != fv2.end()) {
            if (it1->first == it2->first) {
                for (unsigned int i : it1->second) if (!good1 || (*good1)[i]) qi.push_back((int32_t)i);
                for (unsigned int i : it2->second) if (!good2 || (*good2)[i]) ci.push_back((int32_t)i);
                gq.push_back((int32_t)qi.size()); gc.push_back((int32_t)ci.size());
                ++it1; ++it2;
            } else if (it1->first < it2->first) it1 = fv1.lower_bound(it2->first);
            else it2 = fv2.lower_bound(it1->first);
        }
        out.assign(mode == 0 ? n2 : n1, -1);
        std::vector<int32_t> o(out.size() + 1, -1);
        const int nm = b200_match_by_bow_host(d1.data(), a1.data(), n1, d2.data(), a2.data(), n2, gq.data(), qi.data(), gc.data(), ci.data(), (int)gq.size() - 1,
                                              mode, mfNNratio, TH_LOW, mbCheckOrientation ? 1 : 0, o.data(), device_);
        b200slam_detail::check(nm);
        for (size_t i = 0; i < out.size(); i++) out[i] = o[i];
        return nm;
    }

    float mfNNratio;
    bool mbCheckOrientation;
    int device_;
};

}  // namespace ORB_SLAM2

namespace ORB_SLAM2 {

// Frame::UndistortKeyPoints (src/Frame.cc:357-388) and Frame::ComputeImageBounds (src/Frame.cc:418-447) on what they read and write; cam9 = fx fy cx cy
// k1 k2 p1 p2 k3 (mK and mDistCoef).  With k1 == 0 the keypoints are copied / the bounds are the image rectangle, like the reference.  (The grid itself,
// AssignFeaturesToGrid + GetFeaturesInArea, lives on the device behind the matcher entry points and b200_frame_assign_grid / b200_frame_features_in_area.)
inline void UndistortKeyPoints(const std::vector<cv::KeyPoint>& mvKeys, const float cam9[9], std::vector<cv::KeyPoint>& mvKeysUn, int device = 0) {
    const int n = (int)mvKeys.size();
    mvKeysUn = mvKeys;                                             // size, angle, response, octave are kept (Frame.cc:383-387)
    if (n == 0 || cam9[4] == 0.0f) return;
    std::vector<float> xy((size_t)2 * n), un((size_t)2 * n);
    for (int i = 0; i < n; i++) { xy[2 * i] = mvKeys[i].pt.x; xy[2 * i + 1] = mvKeys[i].pt.y; }
    b200slam_detail::check(b200_frame_undistort_points_host(xy.data(), n, cam9, un.data(), device));
    for (int i = 0; i < n; i++) { mvKeysUn[i].pt.x = un[2 * i]; mvKeysUn[i].pt.y = un[2 * i + 1]; }
}
inline void ComputeImageBounds(int width, int height, const float cam9[9], float& mnMinX, float& mnMaxX, float& mnMinY, float& mnMaxY, int device = 0) {
    float b[4];
    b200slam_detail::check(b200_frame_image_bounds(width, height, cam9, b, device));
    mnMinX = b[0]; mnMaxX = b[1]; mnMinY = b[2]; mnMaxY = b[3];
}

// ORB_SLAM2::ORBVocabulary (include/ORBVocabulary.h:31 = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>): the two calls the reference
// makes, loadFromTextFile (System.cc:80) and transform(features, BowVector, FeatureVector, 4) (Frame.cc:353, KeyFrame.cc ComputeBoW).
// The per-descriptor tree descent runs on the device (b200_voc_transform); the two std::maps are assembled from its output arrays
// exactly like TemplatedVocabulary::transform does for TF_IDF weighting + L1 scoring (TemplatedVocabulary.h:1145-1194).
class ORBVocabulary {
public:
    explicit ORBVocabulary(int device = 0) : h_(nullptr), device_(device) {}
    ~ORBVocabulary() { if (h_) b200_voc_destroy(h_); }
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;

    bool loadFromTextFile(const std::string& filename) {          // TemplatedVocabulary.h:1338-1425
        std::ifstream f(filename.c_str());
        if (!f.good()) return false;
        std::string s;
        std::getline(f, s);
        std::stringstream ss(s);
        int k = -1, L = -1, n1 = -1, n2 = -1;
        ss >> k >> L >> n1 >> n2;
        if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;
        std::vector<int32_t> parent(1, 0);
        std::vector<uint8_t> leaf(1, 0), desc(32, 0);
        std::vector<double> weight(1, 0.0);
        while (std::getline(f, s)) {
            if (s.empty()) continue;
            std::stringstream sn(s);
            int pid = 0, is_leaf = 0;
            sn >> pid >> is_leaf;
            parent.push_back(pid); leaf.push_back(is_leaf > 0 ? 1 : 0);
            for (int i = 0; i < 32; i++) { int v = 0; sn >> v; desc.push_back((uint8_t)v); }
            double w = 0; sn >> w; weight.push_back(w);
        }
        if (h_) { b200_voc_destroy(h_); h_ = nullptr; }
        return b200_voc_create(&h_, k, L, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), device_) == B200_OK;
    }
    bool empty() const { return h_ == nullptr; }
    unsigned int size() const { return h_ ? (unsigned int)b200_voc_num_words(h_) : 0u; }

    // void transform(const std::vector<TDescriptor>& features, BowVector& v, FeatureVector& fv, int levelsup) const (TemplatedVocabulary.h:1127)
    // with the descriptors as the rows of mDescriptors (Converter::toDescriptorVector only splits the matrix into rows)
    void transform(const cv::Mat& descriptors, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const {
        v.clear(); fv.clear();
        if (empty()) return;
        const int n = descriptors.rows;
        std::vector<uint8_t> d((size_t)n * 32);
        for (int i = 0; i < n; i++) std::memcpy(&d[(size_t)i * 32], descriptors.ptr(i), 32);
        std::vector<int32_t> word(n), node(n);
        std::vector<double> weight(n);
        b200slam_detail::check(b200_voc_transform_host(h_, d.data(), n, levelsup, word.data(), weight.data(), node.data()));
        for (int i = 0; i < n; i++)
            if (weight[i] > 0) {                                   // not stopped
                v[(DBoW2::WordId)word[i]] += weight[i];           // BowVector::addWeight
                fv[(DBoW2::NodeId)node[i]].push_back((unsigned int)i);
            }
        double norm = 0.0;                                         // BowVector::normalize(L1)
        for (DBoW2::BowVector::iterator it = v.begin(); it != v.end(); ++it) norm += std::fabs(it->second);
        if (norm > 0.0) for (DBoW2::BowVector::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
    }

private:
    b200_voc_t h_;
    int device_;
};

}  // namespace ORB_SLAM2

namespace aruco {

// aruco::CameraParameters essentials (cameraparameters.h:46-60): float camera matrix (fx fy cx cy) and distortion k1 k2 p1 p2 k3
class CameraParameters {
public:
    float fx, fy, cx, cy, dist[5];
    int width, height;                     // CamSize
    CameraParameters() : fx(0), fy(0), cx(0), cy(0), dist{0, 0, 0, 0, 0}, width(-1), height(-1) {}
    // setParams(cv::Mat cameraMatrix, cv::Mat distorsionCoeff, cv::Size size) (cameraparameters.h:83), on plain numbers
    void setParams(float fx_, float fy_, float cx_, float cy_, const float* distorsion, int ndist, int w, int h) {
        fx = fx_; fy = fy_; cx = cx_; cy = cy_;
        for (int i = 0; i < 5; i++) dist[i] = (distorsion && i < ndist) ? distorsion[i] : 0.f;
        width = w; height = h;
    }
#ifndef B200SLAM_NO_OPENCV
    void setParams(const cv::Mat& K, const cv::Mat& D, cv::Size size) {
        cv::Mat k32, d32;
        K.convertTo(k32, CV_32F); D.convertTo(d32, CV_32F);
        float d[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < (int)d32.total() && i < 5; i++) d[i] = d32.ptr<float>(0)[i];
        setParams(k32.at<float>(0, 0), k32.at<float>(1, 1), k32.at<float>(0, 2), k32.at<float>(1, 2), d, 5, size.width, size.height);
    }
#endif
    bool isValid() const { return fx != 0 && fy != 0; }
    // void CameraParameters::resize(cv::Size size) (cameraparameters.cpp:158-173), as a copy: float factors, fx cx by width, fy cy by height
    CameraParameters resized(int w, int h) const {
        CameraParameters c = *this;
        if (w == width && h == height) return c;
        const float AxFactor = float(w) / float(width), AyFactor = float(h) / float(height);
        c.fx *= AxFactor; c.cx *= AxFactor; c.fy *= AyFactor; c.cy *= AyFactor;
        c.width = w; c.height = h;
        return c;
    }
    void cam9(float* out) const { out[0] = fx; out[1] = fy; out[2] = cx; out[3] = cy; for (int i = 0; i < 5; i++) out[4 + i] = dist[i]; }
};

// aruco::Marker essentials (marker.h:47-59): std::vector<cv::Point2f> of 4 corners + id + ssize + Rvec/Tvec; ordered by id
class Marker : public std::vector<cv::Point2f> {
public:
    int id;
    float ssize;
#ifndef B200SLAM_NO_OPENCV
    cv::Mat Rvec, Tvec;                    // 3x1 CV_32F once calculateExtrinsics ran (marker.cpp:337-338)
#else
    float Rvec[3], Tvec[3];
#endif
    std::string dict_info;                 // additional info about the dictionary (marker.h:57): Dictionary::getName() of the table the id came from
    std::vector<cv::Point> contourPoints;  // points of the contour (marker.h:59), copied from the candidate's border (markerdetector_impl.cpp:6759-6772)
    float pose2_rvec[3], pose2_tvec[3], err1, err2;      // the second IPPE solution and both reprojection errors (Frame.cc:155-177)
    Marker() : id(-1), ssize(-1), err1(-1), err2(-1) {
#ifdef B200SLAM_NO_OPENCV
        for (int i = 0; i < 3; i++) Rvec[i] = Tvec[i] = 0;
#endif
        for (int i = 0; i < 3; i++) pose2_rvec[i] = pose2_tvec[i] = 0;
    }
    bool operator<(const Marker& m) const { return id < m.id; }
    bool isValid() const { return id != -1 && size() == 4; }

    // void calculateExtrinsics(float markerSize, const CameraParameters& CP, bool setYPerpendicular = false) (marker.h:98-99)
    void calculateExtrinsics(float markerSizeMeters, const CameraParameters& CP, bool setYPerpendicular = false, int device = 0) {
        if (!isValid()) throw std::runtime_error("!isValid(): invalid marker. It is not possible to calculate extrinsics");
        if (markerSizeMeters <= 0) throw std::runtime_error("markerSize<=0: invalid markerSize");
        if (!CP.isValid()) throw std::runtime_error("!CP.isValid(): invalid camera parameters. It is not possible to calculate extrinsics");
        if (setYPerpendicular) throw std::runtime_error("b200slam: setYPerpendicular is not on the reference's path (src/Frame.cc:142)");
        b200_marker m; m.id = id;
        for (int k = 0; k < 4; k++) { m.xy[2 * k] = (*this)[k].x; m.xy[2 * k + 1] = (*this)[k].y; }
        float cam[9]; CP.cam9(cam);
        b200_marker_pose p;
        b200slam_detail::check(b200_aruco_pose_host(&m, 1, markerSizeMeters, cam, &p, device));
        setPose(p, markerSizeMeters);
    }
    void setPose(const b200_marker_pose& p, float markerSizeMeters) {
#ifndef B200SLAM_NO_OPENCV
        Rvec.create(3, 1, CV_32F); Tvec.create(3, 1, CV_32F);
        for (int i = 0; i < 3; i++) { Rvec.at<float>(i) = p.rvec[i]; Tvec.at<float>(i) = p.tvec[i]; }
#else
        for (int i = 0; i < 3; i++) { Rvec[i] = p.rvec[i]; Tvec[i] = p.tvec[i]; }
#endif
        for (int i = 0; i < 3; i++) { pose2_rvec[i] = p.rvec2[i]; pose2_tvec[i] = p.tvec2[i]; }
        err1 = p.err1; err2 = p.err2; ssize = markerSizeMeters;
    }
};

// void Frame::UndistortArucoCorners() (src/Frame.cc:388-416) on what it reads and writes: the frame's markers and the camera -> mvArucoUn (4 * NA points,
// marker-major).  Like the reference it leaves mvArucoUn untouched when k1 == 0.
class CameraParameters;
class Marker;
inline void UndistortArucoCorners(const std::vector<Marker>& mvMarkers, const CameraParameters& cam, std::vector<cv::Point2f>& mvArucoUn, int device = 0);

// std::vector<std::pair<cv::Mat,double> > aruco::solvePnP(objPoints, imgPoints, cameraMatrix, distCoeffs) (ippe.h:14-15, ippe.cpp:72-88): both IPPE
// solutions of one marker, smaller reprojection error first, as (4 x 4 [R | t] float, error).  src/Frame.cc:155-177 calls it per marker with the ORIGINAL
// mK / mDistCoef (detect() itself used the camera resized to the image) and tests v2pose[0].second / v2pose[1].second < 0.7.
// solvePnPSquare is the array form (marker side length, 4 image points, fx fy cx cy k1 k2 p1 p2 k3); poses row-major 4 x 4.
struct PnPSolution { float T[16]; double error; };
inline std::vector<PnPSolution> solvePnPSquare(float markerSize, const cv::Point2f imgPoints[4], const float cam9[9], int device = 0) {
    b200_marker m; m.id = 0;
    for (int k = 0; k < 4; k++) { m.xy[2 * k] = imgPoints[k].x; m.xy[2 * k + 1] = imgPoints[k].y; }
    b200_marker_pose p;
    b200slam_detail::check(b200_aruco_pose_host(&m, 1, markerSize, cam9, &p, device));
    std::vector<PnPSolution> out(2);
    for (int s = 0; s < 2; s++) {
        const float* r = s ? p.rvec2 : p.rvec; const float* t = s ? p.tvec2 : p.tvec;
        const double th = std::sqrt((double)r[0] * r[0] + (double)r[1] * r[1] + (double)r[2] * r[2]);      // Rodrigues (getRTMatrix, ippe.cpp:16-60)
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (th > 1e-12) {
            const double x = r[0] / th, y = r[1] / th, z = r[2] / th, c = std::cos(th), sn = std::sin(th), c1 = 1 - c;
            const double Rr[9] = {c + c1 * x * x, c1 * x * y - sn * z, c1 * x * z + sn * y, c1 * x * y + sn * z, c + c1 * y * y, c1 * y * z - sn * x,
                                  c1 * x * z - sn * y, c1 * y * z + sn * x, c + c1 * z * z};
            for (int i = 0; i < 9; i++) R[i] = Rr[i];
        }
        float* T = out[s].T;
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T[4 * i + j] = (float)R[3 * i + j]; T[4 * i + 3] = t[i]; }
        T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
        out[s].error = s ? p.err2 : p.err1;
    }
    return out;
}
#ifndef B200SLAM_NO_OPENCV
inline std::vector<std::pair<cv::Mat, double> > solvePnP(const std::vector<cv::Point3f>& objPoints, const std::vector<cv::Point2f>& imgPoints,
                                                         const cv::Mat& cameraMatrix, const cv::Mat& distCoeffs) {
    if (objPoints.size() != 4 || imgPoints.size() != 4) throw std::runtime_error("b200slam: aruco::solvePnP expects the 4 corners of one marker");
    const float size = objPoints[1].x - objPoints[0].x;            // Marker::get3DPoints order (marker.cpp:358-366), as src/Frame.cc:157-160 builds it
    if (!(size > 0) || objPoints[0].y != objPoints[1].y || objPoints[2].x != objPoints[1].x || objPoints[3].x != objPoints[0].x)
        throw std::runtime_error("b200slam: aruco::solvePnP supports the canonical marker square only");
    cv::Mat k32, d32;
    cameraMatrix.convertTo(k32, CV_32F); distCoeffs.convertTo(d32, CV_32F);
    float cam9[9] = {k32.at<float>(0, 0), k32.at<float>(1, 1), k32.at<float>(0, 2), k32.at<float>(1, 2), 0, 0, 0, 0, 0};
    for (int i = 0; i < (int)d32.total() && i < 5; i++) cam9[4 + i] = d32.ptr<float>(0)[i];
    const std::vector<PnPSolution> sol = solvePnPSquare(size, imgPoints.data(), cam9);
    std::vector<std::pair<cv::Mat, double> > out;
    for (size_t s = 0; s < sol.size(); s++) out.push_back(std::make_pair(cv::Mat(4, 4, CV_32F, (void*)sol[s].T).clone(), sol[s].error));
    return out;
}
#endif

enum DetectionMode : int { DM_NORMAL = 0, DM_FAST = 1, DM_VIDEO_FAST = 2 };                      // markerdetector.h:67 (namespace scope, as src/Frame.cc:134 spells it)
enum CornerRefinementMethod : int { CORNER_SUBPIX = 0, CORNER_LINES = 1, CORNER_NONE = 2 };      // markerdetector.h:76

inline void UndistortArucoCorners(const std::vector<Marker>& mvMarkers, const CameraParameters& cam, std::vector<cv::Point2f>& mvArucoUn, int device) {
    if (cam.dist[0] == 0.0f) return;                               // src/Frame.cc:391-394
    const int n = 4 * (int)mvMarkers.size();
    std::vector<float> xy((size_t)2 * n + 2), un((size_t)2 * n + 2);
    for (size_t i = 0; i < mvMarkers.size(); i++)
        for (int j = 0; j < 4; j++) { xy[2 * (4 * i + j)] = mvMarkers[i][j].x; xy[2 * (4 * i + j) + 1] = mvMarkers[i][j].y; }
    float cam9[9]; cam.cam9(cam9);
    b200slam_detail::check(b200_frame_undistort_points_host(xy.data(), n, cam9, un.data(), device));
    mvArucoUn.resize(n);
    for (int i = 0; i < n; i++) mvArucoUn[i] = cv::Point2f(un[2 * i], un[2 * i + 1]);
}

class MarkerDetector {
public:
    // MarkerDetector::Params (markerdetector.h:85-200), the one member src/Frame.cc:135 touches
    struct Params {
        void setCornerRefinementMethod(CornerRefinementMethod m) {                                   // markerdetector.h:129
            if (m != CORNER_LINES) throw std::runtime_error("b200slam: only CORNER_LINES is supported");
        }
    };
    Params& getParameters() { return params_; }                                                      // markerdetector.h:320

    MarkerDetector() : dict_("ALL_DICTS"), h_(nullptr), w_(0), ht_(0), device_(0) {}
    explicit MarkerDetector(const std::string& dict_name, int device = 0) : dict_(dict_name), h_(nullptr), w_(0), ht_(0), device_(device) {}
    ~MarkerDetector() { if (h_) b200_aruco_destroy(h_); }
    MarkerDetector(const MarkerDetector&) = delete;
    MarkerDetector& operator=(const MarkerDetector&) = delete;

    // markerdetector.h:337.  error_correction_rate must be 0 (the reference's setting, src/Frame.cc:133)
    void setDictionary(const std::string& dict_type, float error_correction_rate = 0) {
        if (error_correction_rate != 0) throw std::runtime_error("b200slam: only error_correction_rate = 0 is supported");
        dict_ = dict_type;
        if (h_) { b200_aruco_destroy(h_); h_ = nullptr; }
    }
    // markerdetector.h:258: only the mode the reference uses (src/Frame.cc:134)
    void setDetectionMode(DetectionMode dm, float minMarkerSize = 0) {
        if (dm != DM_NORMAL || minMarkerSize != 0) throw std::runtime_error("b200slam: only DM_NORMAL with minMarkerSize 0 is supported");
    }
    void setCornerRefinementMethod(CornerRefinementMethod m) {
        if (m != CORNER_LINES) throw std::runtime_error("b200slam: only CORNER_LINES is supported");
    }

    // std::vector<aruco::Marker> detect(const cv::Mat& input) (markerdetector.h:276); markers sorted by id.
    std::vector<Marker> detect(const cv::Mat& input) { return detect(input, CameraParameters(), -1.f); }

    // std::vector<aruco::Marker> detect(const cv::Mat& input, const CameraParameters& camParams, float markerSizeMeters,
    //                                   bool setYPerpendicular = false) (markerdetector.h:277-278) -- the call of src/Frame.cc:142.
    // With valid camera parameters and a positive marker size every marker gets Rvec / Tvec / ssize
    // (markerdetector_impl.cpp:8772 -> marker.cpp:322-343 -> ippe.cpp:90-100), computed on the device for all markers at once.
    std::vector<Marker> detect(const cv::Mat& input, const CameraParameters& camParams, float markerSizeMeters, bool setYPerpendicular = false) {
        if (input.empty()) return std::vector<Marker>();
        if (input.type() != CV_8UC1) throw std::runtime_error("b200slam: detect expects a CV_8UC1 image");
        if (setYPerpendicular) throw std::runtime_error("b200slam: setYPerpendicular is not on the reference's path (src/Frame.cc:142)");
        ensure(input.cols, input.rows);
        const int cap = b200_aruco_max_markers(h_);
        std::vector<b200_marker> m(cap);
        std::vector<b200_marker_pose> p(cap);
        int32_t n = 0;
        // markerdetector_impl.cpp detect(input, markers, camParams, size, ...): "if (camParams.CamSize != input.size() && camParams.isValid() &&
        // markerSizeMeters > 0) { cp_aux = camParams; cp_aux.resize(input.size()); ... }" - always taken in the reference, whose CamSize is the
        // hard-coded 1280 x 720 of src/Frame.cc:132 whatever the camera delivers
        const bool want_pose = camParams.isValid() && markerSizeMeters > 0;
        float cam[9];
        if (want_pose) {
            const CameraParameters cp = (camParams.width > 0 && camParams.height > 0) ? camParams.resized(input.cols, input.rows) : camParams;
            cp.cam9(cam);
        }
        // markers, poses (marker.cpp:322-343 -> ippe.cpp, on the device for all markers at once) and contour points (markerdetector_impl.cpp:6759-6772)
        // in ONE library call with one synchronisation
        std::vector<int32_t> ofs((size_t)cap + 1, 0);
        if (fillContourPoints) contour_xy_.resize((size_t)2 * contour_cap_ + 2);
        int total = b200_aruco_detect_frame_host(h_, input.data, input.cols, input.rows, (int64_t)input.step, m.data(), &n, markerSizeMeters,
                                                 want_pose ? cam : nullptr, p.data(), fillContourPoints ? ofs.data() : nullptr,
                                                 fillContourPoints ? contour_xy_.data() : nullptr, fillContourPoints ? contour_cap_ : 0);
        b200slam_detail::check(total);
        std::vector<Marker> out(n);
        for (int i = 0; i < n; i++) {
            out[i].id = m[i].id;
            for (int k = 0; k < 4; k++) out[i].push_back(cv::Point2f(m[i].xy[2 * k], m[i].xy[2 * k + 1]));
            out[i].dict_info = dict_;                                                                // Dictionary::getName() (dictionary.cpp:118-231)
        }
        if (n > 0 && fillContourPoints) {
            if (total > contour_cap_) {                                                              // grow once, fetch again
                contour_cap_ = total;
                contour_xy_.resize((size_t)2 * contour_cap_ + 2);
                b200slam_detail::check(b200_aruco_get_contours(h_, 0, n, ofs.data(), contour_xy_.data(), contour_cap_));
            }
            for (int i = 0; i < n; i++) {
                const int len = ofs[i + 1] - ofs[i];
                out[i].contourPoints.resize(len);
                for (int j = 0; j < len; j++) out[i].contourPoints[j] = cv::Point(contour_xy_[2 * (ofs[i] + j)], contour_xy_[2 * (ofs[i] + j) + 1]);
            }
        }
        if (want_pose) for (int i = 0; i < n; i++) out[i].setPose(p[i], markerSizeMeters);
        return out;
    }

private:
    void ensure(int w, int h) {
        if (h_ && w <= w_ && h <= ht_) return;
        if (h_) { b200_aruco_destroy(h_); h_ = nullptr; }
        w_ = w > w_ ? w : w_; ht_ = h > ht_ ? h : ht_;
        // the reference's default-constructed detector searches several dictionaries at once; src/Frame.cc:133 always calls setDictionary first
        if (dict_ == "ALL_DICTS") throw std::runtime_error("b200slam: multi-dictionary ALL_DICTS detection is not on the reference's configured path "
                                                           "(src/Frame.cc:133): call setDictionary(name) before detect");
        b200slam_detail::check(b200_aruco_create(&h_, dict_.c_str(), w_, ht_, 1, device_));
    }
    std::string dict_;
    b200_aruco_t h_; int w_, ht_, device_;
    Params params_;
    std::vector<int32_t> contour_xy_; int contour_cap_ = 16384;
public:
    bool fillContourPoints = true;            // aruco::Marker::contourPoints (marker.h:59); nothing on the reference's path reads it (src/ has no use): switch off to save the copy
};

}  // namespace aruco
