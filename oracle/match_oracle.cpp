// oracle/match_oracle.cpp -- TEST INFRASTRUCTURE (CPU checker), not product code.
// Restatement of the ORBmatcher cores (reference src/ORBmatcher.cc) over plain arrays.  The reference ships no tests;
// this file is pinned by the reference ITSELF: src/ORBmatcher.cc compiles unmodified on oracle/matchshim into oracle/_ref/libref_match.so
// (oracle/ref_match_wrap.cpp), and tests/test_oracle_match_vs_ref.py demands equality - live where /root/reference exists, and everywhere by
// replaying tests/golden/match_ref.npz.  Known-answer checks (numpy popcount, self-match distance 0) stay in tests/test_oracle_match.py.
// Not reachable that way: ComputeDistinctiveDescriptors (src/MapPoint.cc needs the real classes) - pinned by a numpy definition.
#include "oracle.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <vector>
#include <cstring>
#include <thread>
#include <atomic>

namespace {
const int TH_LOW = 50, HISTO_LENGTH = 30;   // ORBmatcher.cc:37-39

// ORBmatcher::DescriptorDistance, ORBmatcher.cc:1651-1667 (SWAR popcount over 8 x 32 bit)
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4); memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// ORBmatcher::ComputeThreeMaxima, ORBmatcher.cc:1605-1646
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}
}  // namespace

extern "C" {

int oracle_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (ORBmatcher.cc:159-292) with one vocabulary node that holds
// every index on both sides and a good MapPoint behind every keyframe feature.
// matches[n_f]: keyframe index matched to frame keypoint i, or -1.  Returns nmatches.
int oracle_search_by_bow_bf(const uint8_t* kf_desc, const float* kf_angle, int n_kf,
                            const uint8_t* f_desc, const float* f_angle, int n_f,
                            float nnratio, int check_ori, float factor, int32_t* matches) {
    for (int i = 0; i < n_f; i++) matches[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int iKF = 0; iKF < n_kf; iKF++) {
        const uint8_t* dKF = kf_desc + (size_t)iKF * 32;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int iF = 0; iF < n_f; iF++) {
            if (matches[iF] >= 0) continue;
            const int dist = descriptor_distance(dKF, f_desc + (size_t)iF * 32);
            if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = iF; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= TH_LOW) {
            if ((float)bestDist1 < nnratio * (float)bestDist2) {
                matches[bestIdxF] = iKF;
                if (check_ori) {
                    float rot = kf_angle[iKF] - f_angle[bestIdxF];
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    if (bin < 0) bin = 0;
                    if (bin >= HISTO_LENGTH) bin = HISTO_LENGTH - 1;   // the reference asserts the range
                    rotHist[bin].push_back(bestIdxF);
                }
                nmatches++;
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { matches[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// batch driver for CPU-baseline timing: one reference set against n frames laid out like the extractor's output
// (desc [n][cap][32], angles inside 28-byte keypoint records), nthreads workers
int oracle_search_by_bow_bf_batch(const uint8_t* kf_desc, const float* kf_angle, int n_kf,
                                  const uint8_t* f_desc, const uint8_t* f_kps28, const int32_t* n_f, int n, int cap,
                                  float nnratio, int check_ori, float factor, int32_t* matches, int32_t* n_matches, int nthreads) {
    std::atomic<int> next(0);
    auto work = [&]() {
        std::vector<float> ang(cap);
        for (int f; (f = next.fetch_add(1)) < n;) {
            for (int i = 0; i < n_f[f]; i++) memcpy(&ang[i], f_kps28 + ((size_t)f * cap + i) * 28 + 12, 4);
            n_matches[f] = oracle_search_by_bow_bf(kf_desc, kf_angle, n_kf, f_desc + (size_t)f * cap * 32, ang.data(), n_f[f], nnratio, check_ori, factor,
                                                   matches + (size_t)f * cap);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return 0;
}

// best / second-best over an explicit candidate list, the inner loop shared by SearchByProjection
// (ORBmatcher.cc:84-116) and SearchForInitialization (437-461): strict '<' keeps the first minimum.
void oracle_match_candidates(const uint8_t* qd, int nq, const uint8_t* td, const int32_t* ofs, const int32_t* cand,
                             int32_t* best_idx, int32_t* best_dist, int32_t* second_dist) {
    for (int q = 0; q < nq; q++) {
        int b1 = 256, b2 = 256, bi = -1;
        for (int i = ofs[q]; i < ofs[q + 1]; i++) {
            const int d = descriptor_distance(qd + (size_t)q * 32, td + (size_t)cand[i] * 32);
            if (d < b1) { b2 = b1; b1 = d; bi = cand[i]; }
            else if (d < b2) b2 = d;
        }
        best_idx[q] = bi; best_dist[q] = b1; second_dist[q] = b2;
    }
}

}  // extern "C"

// ---- SearchForInitialization (src/ORBmatcher.cc:409-524) on plain arrays ------------------------------------------
// F1 / F2 = undistorted keypoints + descriptors; F2's grid comes from the frame oracle (AssignFeaturesToGrid / GetFeaturesInArea).
// prev_matched [n1][2] is read and updated like vbPrevMatched; matches12 [n1] receives vnMatches12; returns nmatches.
extern "C" void oracle_assign_grid(const oracle_keypoint* un, int n, const float* bounds4, int32_t* cell_start, int32_t* cell_items);
extern "C" int oracle_keyframe_features_in_area(const oracle_keypoint* un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                                                float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);
extern "C" int oracle_features_in_area(const oracle_keypoint* un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                                       float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);
extern "C" int oracle_search_for_initialization(const oracle_keypoint* k1, const uint8_t* d1, int n1, const oracle_keypoint* k2, const uint8_t* d2, int n2,
                                                const float* bounds4, float* prev_matched, int window, float nnratio, int check_ori, int32_t* matches12) {
    int nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<std::vector<int> > rotHist(HISTO_LENGTH);
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> matched_dist(n2, 0x7fffffff), matches21(n2, -1);
    std::vector<int32_t> cs(64 * 48 + 1), ci(n2 > 0 ? n2 : 1), cand(n2 > 0 ? n2 : 1);
    oracle_assign_grid(k2, n2, bounds4, cs.data(), ci.data());
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = k1[i1].octave;
        if (level1 > 0) continue;
        const int nc = oracle_features_in_area(k2, cs.data(), ci.data(), bounds4, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window, level1, level1,
                                               cand.data(), n2);
        if (nc == 0) continue;
        int bestDist = 0x7fffffff, bestDist2 = 0x7fffffff, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            const int i2 = cand[c];
            const int dist = descriptor_distance(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)i2);
            if (matched_dist[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                matches21[bestIdx2] = i1;
                matched_dist[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) {
                    float rot = k1[i1].angle - k2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist.data(), HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; i1++) if (matches12[i1] >= 0) { prev_matched[2 * i1] = k2[matches12[i1]].x; prev_matched[2 * i1 + 1] = k2[matches12[i1]].y; }
    return nmatches;
}

// ---- SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (src/ORBmatcher.cc:526-659) with one vocabulary node that holds every
// index on both sides and a good MapPoint behind every feature: strict bestDist1 < TH_LOW, factor 1.0f/HISTO_LENGTH (upstream),
// vbMatched2 greedy.  matches12[n1] = index in KF2 or -1; returns nmatches.
extern "C" int oracle_search_by_bow_kfkf_bf(const uint8_t* d1, const float* a1, int n1, const uint8_t* d2, const float* a2, int n2,
                                            float nnratio, int check_ori, int32_t* matches12) {
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<char> matched2(n2, 0);
    std::vector<std::vector<int> > rotHist(HISTO_LENGTH);
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    for (int idx1 = 0; idx1 < n1; idx1++) {
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int idx2 = 0; idx2 < n2; idx2++) {
            if (matched2[idx2]) continue;
            const int dist = descriptor_distance(d1 + 32 * (size_t)idx1, d2 + 32 * (size_t)idx2);
            if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 < TH_LOW) {
            if ((float)bestDist1 < nnratio * (float)bestDist2) {
                matches12[idx1] = bestIdx2;
                matched2[bestIdx2] = 1;
                if (check_ori) {
                    float rot = a1[idx1] - a2[bestIdx2];
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(idx1);
                }
                nmatches++;
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist.data(), HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) { matches12[idx1] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ---- SearchByProjection on arrays -----------------------------------------------------------------------------------------
// The geometry (projection of the map points, radius, level window) is host glue in the reference and is passed in ready-made:
// query q = (x, y, radius, minLevel, maxLevel, descriptor, angle, observed).  `occupied[i2]` is the reference's
// "F.mvpMapPoints[i2] && F.mvpMapPoints[i2]->Observations() > 0" and is updated when an observed point is assigned.
//   mode 0: SearchByProjection(Frame&, const vector<MapPoint*>&, th)          (src/ORBmatcher.cc:45-129): best / second with their
//           levels, bestDist <= TH_HIGH, rejected when both lie in the same level and bestDist > ratio * bestDist2
//   mode 1: SearchByProjection(Frame& Current, const Frame& Last, th, mono)   (src/ORBmatcher.cc:1332-1474): best only,
//           bestDist <= TH_HIGH, rotation histogram with factor 1.0f / HISTO_LENGTH over the assigned frame indices
// assign[n2] = query index assigned to frame keypoint i2 or -1 (F.mvpMapPoints[bestIdx] = pMP); returns nmatches.
extern "C" int oracle_search_by_projection(const oracle_keypoint* k2, const uint8_t* d2, int n2, const float* bounds4, uint8_t* occupied,
                                           const float* q_xyr, const int32_t* q_lev, const uint8_t* q_desc, const float* q_angle, const uint8_t* q_observed, int nq,
                                           int mode, float nnratio, int check_ori, int th_high, int32_t* assign) {
    const int TH_HIGH = th_high > 0 ? th_high : 100;          // ORBdist of the relocalisation variant (ORBmatcher.cc:1559), else TH_HIGH
    const bool keyframe = mode == 2;                          // mode 2: mode 1 over KeyFrame::GetFeaturesInArea (SearchByProjection(KeyFrame*, Scw, ...), :294-407)
    if (keyframe) mode = 1;
    for (int i = 0; i < n2; i++) assign[i] = -1;
    int nmatches = 0;
    std::vector<int32_t> cs(64 * 48 + 1), ci(n2 > 0 ? n2 : 1), cand(n2 > 0 ? n2 : 1);
    oracle_assign_grid(k2, n2, bounds4, cs.data(), ci.data());
    std::vector<std::vector<int> > rotHist(HISTO_LENGTH);
    const float factor = 1.0f / HISTO_LENGTH;
    for (int q = 0; q < nq; q++) {
        const int nc = (keyframe ? oracle_keyframe_features_in_area : oracle_features_in_area)(k2, cs.data(), ci.data(), bounds4, q_xyr[3 * q], q_xyr[3 * q + 1],
                                                                                              q_xyr[3 * q + 2], q_lev[2 * q], q_lev[2 * q + 1], cand.data(), n2);
        if (nc == 0) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = cand[c];
            if (occupied[idx]) continue;
            const int dist = descriptor_distance(q_desc + 32 * (size_t)q, d2 + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = k2[idx].octave; bestIdx = idx; }
            else if (mode == 0 && dist < bestDist2) { bestLevel2 = k2[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (mode == 0 && bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            assign[bestIdx] = q;
            if (q_observed[q]) occupied[bestIdx] = 1;
            nmatches++;
            if (mode == 1 && check_ori) {
                float rot = q_angle[q] - k2[bestIdx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx);
            }
        }
    }
    if (mode == 1 && check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist.data(), HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { assign[idx] = -1; nmatches--; }
    }
    return nmatches;
}

// ---- SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:159-292) over REAL feature vectors ------------------------------
// FeatureVectors as sorted arrays (nodes [nn], start [nn + 1], items): the merge walk over common vocabulary nodes, the greedy
// vpMapPointMatches and the rotation histogram (factor HISTO_LENGTH / 360) follow the reference statement by statement.
// kf_valid [n_kf] = "vpMapPointsKF[i] && !isBad()".  matches [n_f] = keyframe index or -1; returns nmatches.
extern "C" int oracle_search_by_bow_nodes(const uint8_t* dkf, const float* akf, const uint8_t* kf_valid, const int32_t* kf_nodes, const int32_t* kf_start,
                                          const int32_t* kf_items, int kf_nn, const uint8_t* df, const float* af, int n_f, const int32_t* f_nodes,
                                          const int32_t* f_start, const int32_t* f_items, int f_nn, float nnratio, int check_ori, int32_t* matches) {
    for (int i = 0; i < n_f; i++) matches[i] = -1;
    int nmatches = 0;
    std::vector<std::vector<int> > rotHist(HISTO_LENGTH);
    const float factor = HISTO_LENGTH / 360.0f;
    int a = 0, b = 0;
    while (a < kf_nn && b < f_nn) {
        if (kf_nodes[a] == f_nodes[b]) {
            for (int iKF = kf_start[a]; iKF < kf_start[a + 1]; iKF++) {
                const int realIdxKF = kf_items[iKF];
                if (!kf_valid[realIdxKF]) continue;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int iF = f_start[b]; iF < f_start[b + 1]; iF++) {
                    const int realIdxF = f_items[iF];
                    if (matches[realIdxF] >= 0) continue;
                    const int dist = descriptor_distance(dkf + 32 * (size_t)realIdxKF, df + 32 * (size_t)realIdxF);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        matches[bestIdxF] = realIdxKF;
                        if (check_ori) {
                            float rot = akf[realIdxKF] - af[bestIdxF];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)round(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(bestIdxF);
                        }
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (kf_nodes[a] < f_nodes[b]) { while (a < kf_nn && kf_nodes[a] < f_nodes[b]) a++; }      // lower_bound
        else { while (b < f_nn && f_nodes[b] < kf_nodes[a]) b++; }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist.data(), HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { matches[idx] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ---- SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (src/ORBmatcher.cc:526-659) over real feature vectors; matches12 [n1] out ---------
extern "C" int oracle_search_by_bow_kfkf_nodes(const uint8_t* d1, const float* a1, const uint8_t* valid1, int n1, const int32_t* nodes1, const int32_t* start1,
                                               const int32_t* items1, int nn1, const uint8_t* d2, const float* a2, const uint8_t* valid2, int n2,
                                               const int32_t* nodes2, const int32_t* start2, const int32_t* items2, int nn2, float nnratio, int check_ori,
                                               int32_t* matches12) {
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<bool> vbMatched2(n2, false);
    std::vector<std::vector<int> > rotHist(HISTO_LENGTH);
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int i1 = start1[a]; i1 < start1[a + 1]; i1++) {
                const int idx1 = items1[i1];
                if (!valid1[idx1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = start2[b]; i2 < start2[b + 1]; i2++) {
                    const int idx2 = items2[i2];
                    if (vbMatched2[idx2] || !valid2[idx2]) continue;
                    const int dist = descriptor_distance(d1 + 32 * (size_t)idx1, d2 + 32 * (size_t)idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        matches12[idx1] = bestIdx2;
                        vbMatched2[bestIdx2] = true;
                        if (check_ori) {
                            float rot = a1[idx1] - a2[bestIdx2];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)round(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(idx1);
                        }
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (nodes1[a] < nodes2[b]) { while (a < nn1 && nodes1[a] < nodes2[b]) a++; }
        else { while (b < nn2 && nodes2[b] < nodes1[a]) b++; }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist.data(), HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { matches12[idx] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:271-331): index of the observation with the least median distance -----------
// desc [n][32] = vDescriptors in observation order.  Returns BestIdx, or -1 for n == 0 (the reference returns early, mDescriptor untouched).
extern "C" int oracle_distinctive_descriptor(const uint8_t* desc, int n) {
    if (n <= 0) return -1;
    const size_t N = (size_t)n;
    std::vector<float> Distances(N * N);
    for (size_t i = 0; i < N; i++) {
        Distances[i * N + i] = 0;
        for (size_t j = i + 1; j < N; j++) {
            const int distij = descriptor_distance(desc + 32 * i, desc + 32 * j);
            Distances[i * N + j] = (float)distij;
            Distances[j * N + i] = (float)distij;
        }
    }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (size_t i = 0; i < N; i++) {
        std::vector<int> vDists(Distances.begin() + i * N, Distances.begin() + (i + 1) * N);
        std::sort(vDists.begin(), vDists.end());
        const int median = vDists[(size_t)(0.5 * (N - 1))];
        if (median < BestMedian) { BestMedian = median; BestIdx = (int)i; }
    }
    return BestIdx;
}
