// b200slam_orbmatcher.hpp -- ORB_SLAM2::ORBmatcher with the reference's EXACT member signatures (reference include/ORBmatcher.h:38-104), on top of
// the C-ABI (include/b200slam.h).  It replaces include/ORBmatcher.h + src/ORBmatcher.cc: every call site of the reference compiles unchanged -
//   src/Tracking.cc:531-532, 917-920, 1011-1017, 1515, 1778, 1858, 1875      src/LocalMapping.cc:283, 857, 882
//   src/LoopClosing.cc:400, 424, 458, 519, 577, 629, 1086                     src/MapPoint.cc:310 (DescriptorDistance)
// - because the arguments are the reference's own objects (Frame&, KeyFrame*, std::vector<MapPoint*>&, cv::Mat Scw ...).
//
// Split between host and device, per member: the per-call and per-point cv::Mat statements that come BEFORE a grid query (pose decomposition,
// projection, depth / viewing-angle / scale tests, MapPoint::PredictScale) are evaluated here with the includer's own cv::Mat arithmetic, exactly
// like the reference does on the host, and the outcome statements AFTER the search (F.mvpMapPoints[i] = pMP, Replace, AddObservation ...) are
// applied here on the real objects in the reference's order.  Everything in between - the 64 x 48 grid, GetFeaturesInArea, the 256-bit Hamming
// distances, best / second-best, the greedy "already matched" replay and the rotation histogram - runs on the GPU in one library call per member.
//
// Include AFTER the reference's Frame.h / KeyFrame.h / MapPoint.h (in the reference: from a one-line include/ORBmatcher.h).
// Header only, C++11; link with -lb200slam.  The CUDA device is ORBmatcher::Device() (default 0).
#ifndef B200SLAM_ORBMATCHER_HPP
#define B200SLAM_ORBMATCHER_HPP
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "b200slam.h"

namespace ORB_SLAM2 {

namespace b200_detail {
inline int checked(int rc) { if (rc < 0) throw std::runtime_error(std::string("b200slam: ") + b200_last_error()); return rc; }
// rows of an N x 32 CV_8U descriptor matrix as one contiguous block
inline void descriptor_rows(const cv::Mat& m, int n, std::vector<uint8_t>& out) {
    out.resize((size_t)(n > 0 ? n : 0) * 32 + 32);
    for (int i = 0; i < n; i++) std::memcpy(&out[(size_t)i * 32], m.ptr(i), 32);
}
inline void descriptor_row(const cv::Mat& d, uint8_t* out) { std::memcpy(out, d.ptr(0), 32); }
template <class T> struct Constants { static const int TH_LOW, TH_HIGH, HISTO_LENGTH; };
template <class T> const int Constants<T>::TH_HIGH = 100;            // src/ORBmatcher.cc:37-39
template <class T> const int Constants<T>::TH_LOW = 50;
template <class T> const int Constants<T>::HISTO_LENGTH = 30;
static_assert(sizeof(cv::KeyPoint) == sizeof(b200_keypoint), "cv::KeyPoint must be the 28-byte record the library reads");
}  // namespace b200_detail

class ORBmatcher : public b200_detail::Constants<void> {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

    // the CUDA device every matcher object of the process works on
    static int& Device() { static int d = 0; return d; }

    // Computes the Hamming distance between two ORB descriptors (src/ORBmatcher.cc:1651-1667).  Stays on the host: the reference calls it O(N^2) times
    // per map point (src/MapPoint.cc:310); the batched forms live on the device (b200_distinctive_descriptors_host, b200_hamming_matrix_host).
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        const unsigned char* pa = a.ptr(0);
        const unsigned char* pb = b.ptr(0);
        int dist = 0;
        for (int i = 0; i < 4; i++) {
            uint64_t x, y;
            std::memcpy(&x, pa + 8 * i, 8); std::memcpy(&y, pb + 8 * i, 8);
            dist += __builtin_popcountll(x ^ y);
        }
        return dist;
    }

    // Search matches between Frame keypoints and projected MapPoints (Tracking::SearchLocalPoints).  src/ORBmatcher.cc:45-129
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3) {
        const bool bFactor = th != 1.0;
        Queries q;
        std::vector<MapPoint*> owner;
        for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
            MapPoint* pMP = vpMapPoints[iMP];
            if (!pMP->mbTrackInView || pMP->isBad()) continue;
            const int& level = pMP->mnTrackScaleLevel;
            float r = RadiusByViewingCos(pMP->mTrackViewCos);
            if (bFactor) r *= th;
            q.push(pMP->mTrackProjX, pMP->mTrackProjY, r * F.mvScaleFactors[level], level - 1, level, pMP->GetDescriptor(), 0.f, pMP->Observations() > 0);
            owner.push_back(pMP);
        }
        return frameSearch(F, q, owner, 0, TH_HIGH);
    }

    // Project MapPoints tracked in last frame into the current frame and search matches (Tracking::TrackWithMotionModel).  src/ORBmatcher.cc:1332-1474
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
        if (!bMono) throw std::runtime_error("b200slam: stereo / RGB-D SearchByProjection is not built (the reference's configured path is monocular)");
        const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
        const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
        Queries q;
        std::vector<MapPoint*> owner;
        for (int i = 0; i < LastFrame.N; i++) {
            MapPoint* pMP = LastFrame.mvpMapPoints[i];
            if (!pMP || LastFrame.mvbOutlier[i]) continue;
            const cv::Mat x3Dw = pMP->GetWorldPos();
            const cv::Mat x3Dc = Rcw * x3Dw + tcw;
            const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
            const float invzc = 1.0 / x3Dc.at<float>(2);
            if (invzc < 0) continue;
            const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
            const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
            if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX || v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
            const int nLastOctave = LastFrame.mvKeys[i].octave;
            const float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];
            q.push(u, v, radius, nLastOctave - 1, nLastOctave + 1, pMP->GetDescriptor(), LastFrame.mvKeysUn[i].angle, pMP->Observations() > 0);
            owner.push_back(pMP);
        }
        return frameSearch(CurrentFrame, q, owner, 1, TH_HIGH);
    }

    // Project MapPoints seen in KeyFrame into the Frame and search matches (Tracking::Relocalization).  src/ORBmatcher.cc:1476-1603
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
        const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
        const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
        const cv::Mat Ow = -Rcw.t() * tcw;
        const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
        Queries q;
        std::vector<MapPoint*> owner;
        for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
            MapPoint* pMP = vpMPs[i];
            if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
            const cv::Mat x3Dw = pMP->GetWorldPos();
            const cv::Mat x3Dc = Rcw * x3Dw + tcw;
            const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
            const float invzc = 1.0 / x3Dc.at<float>(2);
            const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
            const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
            if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX || v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
            const cv::Mat PO = x3Dw - Ow;
            const float dist3D = cv::norm(PO);
            const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            const int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
            const float radius = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
            // here ANY map point hides its keypoint (`if(CurrentFrame.mvpMapPoints[i2]) continue;`, :1545), observed or not
            q.push(u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, pMP->GetDescriptor(), pKF->mvKeysUn[i].angle, true);
            owner.push_back(pMP);
        }
        return frameSearch(CurrentFrame, q, owner, 1, ORBdist, true);
    }

    // Project MapPoints using a Similarity Transformation and search matches (LoopClosing::ComputeSim3 / DetectLoop).  src/ORBmatcher.cc:294-407
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th) {
        cv::Mat Rcw, tcw, Ow;
        decomposeSim3(Scw, Rcw, tcw, Ow);
        std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
        spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
        Queries q;
        std::vector<MapPoint*> owner;
        for (int iMP = 0, iendMP = (int)vpPoints.size(); iMP < iendMP; iMP++) {
            MapPoint* pMP = vpPoints[iMP];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            float u, v; int level;
            if (!projectWorldPoint(pKF, pMP, Rcw, tcw, Ow, true, true, u, v, level)) continue;
            const float radius = th * pKF->mvScaleFactors[level];
            q.push(u, v, radius, level - 1, level, pMP->GetDescriptor(), 0.f, true);
            owner.push_back(pMP);
        }
        // a matched feature hides itself from every later point (vpMatched[idx], :374): replayed in order on the device, no rotation histogram, TH_LOW
        KeyFrameArrays a(pKF);
        std::vector<uint8_t> occupied(a.n + 1, 0);
        for (int i = 0; i < a.n && i < (int)vpMatched.size(); i++) occupied[i] = vpMatched[i] ? 1 : 0;
        std::vector<int32_t> assign(a.n + 1, -1);
        const int nmatches = b200_detail::checked(b200_match_by_projection_host(
            a.keys(), a.desc.data(), a.n, a.bounds, occupied.data(), q.xyr.data(), q.levels.data(), q.desc.data(), q.angle.data(), q.observed.data(), q.size(), 2,
            mfNNratio, 0, TH_LOW, assign.data(), Device()));
        for (int i = 0; i < a.n; i++) if (assign[i] >= 0) vpMatched[i] = owner[assign[i]];
        return nmatches;
    }

    // Search matches between MapPoints in a KeyFrame and ORB in a Frame, constrained to the same vocabulary node.  src/ORBmatcher.cc:159-292
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
        const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
        const int nKF = (int)vpMapPointsKF.size();
        std::vector<char> good(nKF, 0);
        for (int i = 0; i < nKF; i++) good[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
        std::vector<float> aKF(nKF), aF(F.N);
        for (int i = 0; i < nKF; i++) aKF[i] = pKF->mvKeysUn[i].angle;
        for (int i = 0; i < F.N; i++) aF[i] = F.mvKeys[i].angle;                                   // :242 reads F.mvKeys (the reference's choice)
        std::vector<int32_t> out;
        const int nmatches = byBoW(0, pKF->mDescriptors, aKF, &good, pKF->mFeatVec, F.mDescriptors, aF, NULL, F.mFeatVec, out);
        for (int iF = 0; iF < F.N; iF++) if (out[iF] >= 0) vpMapPointMatches[iF] = vpMapPointsKF[out[iF]];
        return nmatches;
    }

    // src/ORBmatcher.cc:526-659
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
        const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int n1 = (int)vpMapPoints1.size(), n2 = (int)vpMapPoints2.size();
        vpMatches12 = std::vector<MapPoint*>(n1, static_cast<MapPoint*>(NULL));
        std::vector<char> good1(n1, 0), good2(n2, 0);
        for (int i = 0; i < n1; i++) good1[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();
        for (int i = 0; i < n2; i++) good2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();
        std::vector<float> a1(n1), a2(n2);
        for (int i = 0; i < n1; i++) a1[i] = pKF1->mvKeysUn[i].angle;
        for (int i = 0; i < n2; i++) a2[i] = pKF2->mvKeysUn[i].angle;
        std::vector<int32_t> out;
        const int nmatches = byBoW(1, pKF1->mDescriptors, a1, &good1, pKF1->mFeatVec, pKF2->mDescriptors, a2, &good2, pKF2->mFeatVec, out);
        for (int i = 0; i < n1; i++) if (out[i] >= 0) vpMatches12[i] = vpMapPoints2[out[i]];
        return nmatches;
    }

    // Matching for the Map Initialization (only used in the monocular case).  src/ORBmatcher.cc:409-524
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10) {
        const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
        vnMatches12 = std::vector<int>(n1, -1);
        std::vector<uint8_t> d1, d2;
        b200_detail::descriptor_rows(F1.mDescriptors, n1, d1); b200_detail::descriptor_rows(F2.mDescriptors, n2, d2);
        std::vector<float> prev((size_t)n1 * 2 + 2);
        for (int i = 0; i < n1; i++) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
        const float bounds[4] = {(float)F2.mnMinX, (float)F2.mnMaxX, (float)F2.mnMinY, (float)F2.mnMaxY};
        std::vector<int32_t> m12((size_t)n1 + 1, -1);
        const int nmatches = b200_detail::checked(b200_match_for_initialization_host(
            (const b200_keypoint*)F1.mvKeysUn.data(), d1.data(), n1, (const b200_keypoint*)F2.mvKeysUn.data(), d2.data(), n2, bounds, prev.data(), windowSize,
            mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), Device()));
        for (int i = 0; i < n1; i++) { vnMatches12[i] = m12[i]; vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
        return nmatches;
    }

    // Matching to triangulate new MapPoints. Check Epipolar Constraint.  src/ORBmatcher.cc:661-829
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
        if (bOnlyStereo) throw std::runtime_error("b200slam: stereo SearchForTriangulation is not built (the reference's configured path is monocular)");
        // epipole of camera 1 in image 2 (:668-674)
        const cv::Mat Cw = pKF1->GetCameraCenter();
        const cv::Mat R2w = pKF2->GetRotation();
        const cv::Mat t2w = pKF2->GetTranslation();
        const cv::Mat C2 = R2w * Cw + t2w;
        const float invz = 1.0f / C2.at<float>(2);
        const float e2[2] = {pKF2->fx * C2.at<float>(0) * invz + pKF2->cx, pKF2->fy * C2.at<float>(1) * invz + pKF2->cy};
        const int n1 = pKF1->N, n2 = pKF2->N;
        std::vector<char> free1(n1, 0), free2(n2, 0);
        for (int i = 0; i < n1; i++) free1[i] = !pKF1->GetMapPoint(i);
        for (int i = 0; i < n2; i++) free2[i] = !pKF2->GetMapPoint(i);
        std::vector<int32_t> gq, qi, gc, ci;
        commonNodes(pKF1->mFeatVec, &free1, pKF2->mFeatVec, &free2, gq, qi, gc, ci);
        std::vector<uint8_t> d1, d2;
        b200_detail::descriptor_rows(pKF1->mDescriptors, n1, d1); b200_detail::descriptor_rows(pKF2->mDescriptors, n2, d2);
        float F[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F[3 * r + c] = F12.at<float>(r, c);
        std::vector<int32_t> m12((size_t)n1 + 1, -1);
        const int nmatches = b200_detail::checked(b200_match_for_triangulation_host(
            (const b200_keypoint*)pKF1->mvKeysUn.data(), d1.data(), n1, (const b200_keypoint*)pKF2->mvKeysUn.data(), d2.data(), n2, gq.data(), qi.data(), gc.data(),
            ci.data(), (int)gq.size() - 1, F, e2, pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(), (int)pKF2->mvScaleFactors.size(),
            mbCheckOrientation ? 1 : 0, TH_LOW, m12.data(), Device()));
        vMatchedPairs.clear();
        vMatchedPairs.reserve(nmatches);
        for (int i = 0; i < n1; i++) if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
        return nmatches;
    }

    // Search matches between MapPoints seen in KF1 and KF2 transforming by a Sim3 [s12*R12|t12].  src/ORBmatcher.cc:1106-1330
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12, const float th) {
        const cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
        const cv::Mat sR12 = s12 * R12;
        const cv::Mat sR21 = (1.0 / s12) * R12.t();
        const cv::Mat t21 = -sR21 * t12;
        const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
        std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
        for (int i = 0; i < N1; i++) {
            MapPoint* pMP = vpMatches12[i];
            if (!pMP) continue;
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
        std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
        sim3Direction(pKF2, vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, th, vnMatch1);       // KF1's points into KF2 (:1151-1217)
        sim3Direction(pKF1, vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, th, vnMatch2);       // KF2's points into KF1 (:1231-1297)
        int nFound = 0;
        for (int i1 = 0; i1 < N1; i1++) {
            const int idx2 = vnMatch1[i1];
            if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
        }
        return nFound;
    }

    // Project MapPoints into KeyFrame and search for duplicated MapPoints.  src/ORBmatcher.cc:831-981
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0) {
        const cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
        const int nMPs = (int)vpMapPoints.size();
        Queries q;
        std::vector<int> who;
        for (int i = 0; i < nMPs; i++) {
            MapPoint* pMP = vpMapPoints[i];
            if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
            float u, v; int level;
            if (!projectWorldPoint(pKF, pMP, Rcw, tcw, Ow, false, true, u, v, level)) continue;
            q.push(u, v, th * pKF->mvScaleFactors[level], level, level, pMP->GetDescriptor(), 0.f, true);
            who.push_back(i);
        }
        std::vector<int> bestIdx, bestDist;
        keyFrameSearch(pKF, q, 5.99, bestIdx, bestDist);
        // the outcome statements on the real objects, in list order; a point that an EARLIER iteration of this call made bad or put into the
        // keyframe is skipped like the reference's `continue` at :851 would
        int nFused = 0;
        for (size_t k = 0; k < who.size(); k++) {
            MapPoint* pMP = vpMapPoints[who[k]];
            if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
            if (bestIdx[k] < 0 || bestDist[k] > TH_LOW) continue;
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx[k]);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) {
                    if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                    else pMPinKF->Replace(pMP);
                }
            } else {
                pMP->AddObservation(pKF, bestIdx[k]);
                pKF->AddMapPoint(pMP, bestIdx[k]);
            }
            nFused++;
        }
        return nFused;
    }

    // Project MapPoints into KeyFrame using a given Sim3 and search for duplicated MapPoints.  src/ORBmatcher.cc:983-1104
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
        cv::Mat Rcw, tcw, Ow;
        decomposeSim3(Scw, Rcw, tcw, Ow);
        const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
        const int nPoints = (int)vpPoints.size();
        Queries q;
        std::vector<int> who;
        for (int iMP = 0; iMP < nPoints; iMP++) {
            MapPoint* pMP = vpPoints[iMP];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            float u, v; int level;
            if (!projectWorldPoint(pKF, pMP, Rcw, tcw, Ow, true, false, u, v, level)) continue;
            q.push(u, v, th * pKF->mvScaleFactors[level], level, level, pMP->GetDescriptor(), 0.f, true);
            who.push_back(iMP);
        }
        std::vector<int> bestIdx, bestDist;
        keyFrameSearch(pKF, q, 0.0, bestIdx, bestDist);
        int nFused = 0;
        for (size_t k = 0; k < who.size(); k++) {
            if (bestIdx[k] < 0 || bestDist[k] > TH_LOW) continue;
            MapPoint* pMP = vpPoints[who[k]];
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx[k]);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) vpReplacePoint[who[k]] = pMPinKF;
            } else {
                pMP->AddObservation(pKF, bestIdx[k]);
                pKF->AddMapPoint(pMP, bestIdx[k]);
            }
            nFused++;
        }
        return nFused;
    }

protected:
    // ---- the reference's protected members (kept for source compatibility of subclasses; the device owns the hot versions) ---------------------
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF2) {   // src/ORBmatcher.cc:139-157
        const float a = kp1.pt.x * F12.at<float>(0, 0) + kp1.pt.y * F12.at<float>(1, 0) + F12.at<float>(2, 0);
        const float b = kp1.pt.x * F12.at<float>(0, 1) + kp1.pt.y * F12.at<float>(1, 1) + F12.at<float>(2, 1);
        const float c = kp1.pt.x * F12.at<float>(0, 2) + kp1.pt.y * F12.at<float>(1, 2) + F12.at<float>(2, 2);
        const float num = a * kp2.pt.x + b * kp2.pt.y + c, den = a * a + b * b;
        if (den == 0) return false;
        return num * num / den < 3.84 * pKF2->mvLevelSigma2[kp2.octave];
    }
    float RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }                                   // src/ORBmatcher.cc:131-137
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {                           // src/ORBmatcher.cc:1605-1646
        int max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < L; i++) {
            const int s = (int)histo[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
    }

    float mfNNratio;
    bool mbCheckOrientation;

private:
    // projected map points of one call, in the order the reference visits them
    struct Queries {
        std::vector<float> xyr, angle;
        std::vector<int32_t> levels;
        std::vector<uint8_t> desc, observed;
        int size() const { return (int)angle.size(); }
        void push(float u, float v, float radius, int minLevel, int maxLevel, const cv::Mat& d, float a, bool obs) {
            xyr.push_back(u); xyr.push_back(v); xyr.push_back(radius);
            levels.push_back(minLevel); levels.push_back(maxLevel);
            desc.resize(desc.size() + 32);
            b200_detail::descriptor_row(d, &desc[desc.size() - 32]);
            angle.push_back(a); observed.push_back(obs ? 1 : 0);
        }
        void pad() { xyr.resize(xyr.size() + 3); levels.resize(levels.size() + 2); desc.resize(desc.size() + 32); angle.push_back(0.f); observed.push_back(0); }
    };

    // what the device reads of a keyframe: undistorted keypoints, descriptors, and the float image bounds behind its grid.  KeyFrame keeps truncated
    // int copies of the bounds (include/KeyFrame.h:211-214) of Frame's static floats (include/Frame.h:191-194, one calibration per process), which the
    // grid was built from (src/Frame.cc:332-343): the library lays the search window from the ints and the cells from the floats, like the reference.
    struct KeyFrameArrays {
        int n;
        std::vector<uint8_t> desc;
        float bounds[4];
        const std::vector<cv::KeyPoint>& k;
        explicit KeyFrameArrays(KeyFrame* pKF) : n((int)pKF->mvKeysUn.size()), k(pKF->mvKeysUn) {
            b200_detail::descriptor_rows(pKF->mDescriptors, n, desc);
            bounds[0] = Frame::mnMinX; bounds[1] = Frame::mnMaxX; bounds[2] = Frame::mnMinY; bounds[3] = Frame::mnMaxY;
        }
        const b200_keypoint* keys() const { return (const b200_keypoint*)k.data(); }
    };

    // Scw -> Rcw, tcw, Ow (src/ORBmatcher.cc:302-307, 992-997)
    static void decomposeSim3(const cv::Mat& Scw, cv::Mat& Rcw, cv::Mat& tcw, cv::Mat& Ow) {
        const cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
        const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
        Rcw = sRcw / scw;
        tcw = Scw.rowRange(0, 3).col(3) / scw;
        Ow = -Rcw.t() * tcw;
    }

    // One map point into a keyframe camera: the statements between GetWorldPos() and GetFeaturesInArea() of Fuse (:865-892), Fuse(Scw) (:1019-1056) and
    // SearchByProjection(pKF, Scw) (:325-360).  doubleInvz: `1.0 / z` (double division, then float) where the reference writes it so, else `1 / z`;
    // doubleZero: the depth test compares against 0.0 instead of 0.0f (same answer; kept for the record).  false = the reference `continue`s.
    template <class KF>
    static bool projectWorldPoint(KF* pKF, MapPoint* pMP, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, bool doubleInvz, bool /*doubleZero*/,
                                  float& u, float& v, int& level) {
        const cv::Mat p3Dw = pMP->GetWorldPos();
        const cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if (p3Dc.at<float>(2) < 0.0f) return false;
        float invz;
        if (doubleInvz) invz = 1.0 / p3Dc.at<float>(2); else invz = 1 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
        u = pKF->fx * x + pKF->cx;
        v = pKF->fy * y + pKF->cy;
        if (!pKF->IsInImage(u, v)) return false;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        const cv::Mat PO = p3Dw - Ow;
        const float dist3D = cv::norm(PO);
        if (dist3D < minDistance || dist3D > maxDistance) return false;
        const cv::Mat Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist3D) return false;
        level = pMP->PredictScale(dist3D, pKF);
        return true;
    }

    // one direction of SearchBySim3: the points of the OTHER keyframe through (Rw, tw) and then (sR, t) into pKF's camera (:1151-1217 / :1231-1297)
    void sim3Direction(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const std::vector<bool>& vbAlreadyMatched, const cv::Mat& Rw, const cv::Mat& tw,
                       const cv::Mat& sR, const cv::Mat& t, float th, std::vector<int>& vnMatch) {
        Queries q;
        std::vector<int> who;
        for (int i = 0, n = (int)vpMapPoints.size(); i < n; i++) {
            MapPoint* pMP = vpMapPoints[i];
            if (!pMP || vbAlreadyMatched[i] || pMP->isBad()) continue;
            const cv::Mat p3Dw = pMP->GetWorldPos();
            const cv::Mat p3Dc1 = Rw * p3Dw + tw;
            const cv::Mat p3Dc2 = sR * p3Dc1 + t;
            if (p3Dc2.at<float>(2) < 0.0) continue;
            const float invz = 1.0 / p3Dc2.at<float>(2);
            const float x = p3Dc2.at<float>(0) * invz, y = p3Dc2.at<float>(1) * invz;
            const float u = pKF->fx * x + pKF->cx, v = pKF->fy * y + pKF->cy;
            if (!pKF->IsInImage(u, v)) continue;
            const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
            const float dist3D = cv::norm(p3Dc2);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            const int level = pMP->PredictScale(dist3D, pKF);
            q.push(u, v, th * pKF->mvScaleFactors[level], level, level, pMP->GetDescriptor(), 0.f, true);
            who.push_back(i);
        }
        std::vector<int> bestIdx, bestDist;
        keyFrameSearch(pKF, q, 0.0, bestIdx, bestDist);
        for (size_t k = 0; k < who.size(); k++) if (bestIdx[k] >= 0 && bestDist[k] <= TH_HIGH) vnMatch[who[k]] = bestIdx[k];
    }

    // the search inside Fuse / Fuse(Scw) / SearchBySim3: per projected point the most similar keyframe feature in its window, levels [predicted - 1,
    // predicted] (the query carries `predicted` as minLevel), optional chi-square gate; independent of what earlier points did to the keyframe
    void keyFrameSearch(KeyFrame* pKF, Queries& q, double chi2, std::vector<int>& bestIdx, std::vector<int>& bestDist) {
        const int nq = q.size();
        bestIdx.assign(nq, -1); bestDist.assign(nq, 256);
        if (nq == 0) return;
        KeyFrameArrays a(pKF);
        std::vector<int32_t> ql(nq), bi((size_t)nq + 1, -1), bd((size_t)nq + 1, 256);
        for (int i = 0; i < nq; i++) ql[i] = q.levels[2 * i];
        b200_detail::checked(b200_match_kf_radius_host(a.keys(), a.desc.data(), a.n, a.bounds, q.xyr.data(), ql.data(), q.desc.data(), nq, pKF->mvInvLevelSigma2.data(),
                                                       (int)pKF->mvInvLevelSigma2.size(), chi2, bi.data(), bd.data(), Device()));
        for (int i = 0; i < nq; i++) { bestIdx[i] = bi[i]; bestDist[i] = bd[i]; }
    }

    // the three Tracking-thread SearchByProjection members after their projections: F's grid, greedy replay, outcome on F.mvpMapPoints.
    // mode 0: best / second with the same-level ratio test (:118-121); mode 1: rotation histogram with factor 1 / HISTO_LENGTH (:1340, :1487)
    // anyPoint: a keypoint is hidden by any map point it holds (relocalisation, :1545) instead of only by one with Observations() > 0 (:87-89, :1407-1409)
    int frameSearch(Frame& F, Queries& q, const std::vector<MapPoint*>& owner, int mode, int thHigh, bool anyPoint = false) {
        const int n = F.N;
        if (q.size() == 0 || n == 0) return 0;
        std::vector<uint8_t> desc;
        b200_detail::descriptor_rows(F.mDescriptors, n, desc);
        std::vector<uint8_t> occupied((size_t)n + 1, 0);
        for (int i = 0; i < n; i++) occupied[i] = (F.mvpMapPoints[i] && (anyPoint || F.mvpMapPoints[i]->Observations() > 0)) ? 1 : 0;
        const float bounds[4] = {(float)F.mnMinX, (float)F.mnMaxX, (float)F.mnMinY, (float)F.mnMaxY};
        std::vector<int32_t> assign((size_t)n + 1, -1);
        const int nmatches = b200_detail::checked(b200_match_by_projection_host(
            (const b200_keypoint*)F.mvKeysUn.data(), desc.data(), n, bounds, occupied.data(), q.xyr.data(), q.levels.data(), q.desc.data(), q.angle.data(),
            q.observed.data(), q.size(), mode, mfNNratio, mbCheckOrientation ? 1 : 0, thHigh, assign.data(), Device()));
        for (int i = 0; i < n; i++) {
            if (assign[i] >= 0) F.mvpMapPoints[i] = owner[assign[i]];
            else if (assign[i] == -2) F.mvpMapPoints[i] = static_cast<MapPoint*>(NULL);       // matched, then cleared by the rotation histogram (:1459-1467)
        }
        return nmatches;
    }

    // merge walk over the vocabulary nodes both FeatureVectors contain (:185-279, :547-632, :692-812): one group per common node
    static void commonNodes(const DBoW2::FeatureVector& fv1, const std::vector<char>* keep1, const DBoW2::FeatureVector& fv2, const std::vector<char>* keep2,
                            std::vector<int32_t>& gq, std::vector<int32_t>& qi, std::vector<int32_t>& gc, std::vector<int32_t>& ci) {
        gq.assign(1, 0); gc.assign(1, 0); qi.clear(); ci.clear();
        DBoW2::FeatureVector::const_iterator it1 = fv1.begin(), it2 = fv2.begin();
        while (it1 != fv1.end() && it2 != fv2.end()) {
            if (it1->first == it2->first) {
                for (size_t k = 0; k < it1->second.size(); k++) if (!keep1 || (*keep1)[it1->second[k]]) qi.push_back((int32_t)it1->second[k]);
                for (size_t k = 0; k < it2->second.size(); k++) if (!keep2 || (*keep2)[it2->second[k]]) ci.push_back((int32_t)it2->second[k]);
                gq.push_back((int32_t)qi.size()); gc.push_back((int32_t)ci.size());
                ++it1; ++it2;
            } else if (it1->first < it2->first) it1 = fv1.lower_bound(it2->first);
            else it2 = fv2.lower_bound(it1->first);
        }
        qi.push_back(0); ci.push_back(0);                       // never empty: the library wants valid pointers
    }

    int byBoW(int mode, const cv::Mat& desc1, const std::vector<float>& a1, const std::vector<char>* keep1, const DBoW2::FeatureVector& fv1,
              const cv::Mat& desc2, const std::vector<float>& a2, const std::vector<char>* keep2, const DBoW2::FeatureVector& fv2, std::vector<int32_t>& out) {
        const int n1 = (int)a1.size(), n2 = (int)a2.size();
        std::vector<uint8_t> d1, d2;
        b200_detail::descriptor_rows(desc1, n1, d1); b200_detail::descriptor_rows(desc2, n2, d2);
        std::vector<int32_t> gq, qi, gc, ci;
        commonNodes(fv1, keep1, fv2, keep2, gq, qi, gc, ci);
        out.assign((size_t)(mode == 0 ? n2 : n1) + 1, -1);
        std::vector<float> p1(a1), p2(a2);
        p1.push_back(0.f); p2.push_back(0.f);
        return b200_detail::checked(b200_match_by_bow_host(d1.data(), p1.data(), n1, d2.data(), p2.data(), n2, gq.data(), qi.data(), gc.data(), ci.data(),
                                                           (int)gq.size() - 1, mode, mfNNratio, TH_LOW, mbCheckOrientation ? 1 : 0, out.data(), Device()));
    }
};

}  // namespace ORB_SLAM2
#endif  // B200SLAM_ORBMATCHER_HPP
