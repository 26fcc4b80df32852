// oracle/match2_oracle.cpp -- TEST INFRASTRUCTURE (CPU restatement, never linked into the product).
// The KeyFrame-side members of ORBmatcher used by the local-mapping and loop-closing threads, restated on flat arrays:
//   SearchForTriangulation                       src/ORBmatcher.cc:661-829   (+ CheckDistEpipolarLine :139-157)
//   Fuse(KeyFrame*, vpMapPoints, th)             src/ORBmatcher.cc:831-981
//   Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)   src/ORBmatcher.cc:983-1104
//   SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)   src/ORBmatcher.cc:294-407
//   SearchBySim3                                 src/ORBmatcher.cc:1106-1330
// Pinned by tests/test_oracle_match2_vs_ref.py against the reference's own ORBmatcher.cc (oracle/_ref/libref_match.so, entry points
// ref_search_for_triangulation / ref_fuse / ref_fuse_sim3 / ref_search_by_projection_loop / ref_search_by_sim3 with the same argument lists)
// and by tests/golden/match_ref2.npz.  Small-matrix algebra follows cv::Mat on CV_32F: products accumulate in double and are stored as float,
// sums / differences are float (oracle/matchshim/opencv2/core/core.hpp).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "oracle.h"

namespace {

const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;                                      // src/ORBmatcher.cc:37-39

int hamming(const uint8_t* a, const uint8_t* b) {                                             // src/ORBmatcher.cc:1651-1667
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

void three_maxima(const std::vector<std::vector<int> >& histo, int& ind1, int& ind2, int& ind3) {      // src/ORBmatcher.cc:1605-1647
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < (int)histo.size(); i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

// cv::Mat (CV_32F) algebra on 3x3 / 3x1 blocks
void mul33_31(const float* A, const float* x, float* y) {
    for (int r = 0; r < 3; r++) { double s = 0; for (int k = 0; k < 3; k++) s += (double)A[3 * r + k] * (double)x[k]; y[r] = (float)s; }
}
void mul33_33(const float* A, const float* B, float* Cm) {
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { double s = 0; for (int k = 0; k < 3; k++) s += (double)A[3 * r + k] * (double)B[3 * k + c]; Cm[3 * r + c] = (float)s; }
}
void transpose33(const float* A, float* T) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) T[3 * c + r] = A[3 * r + c]; }
double dot3(const float* a, const float* b) { double s = 0; for (int k = 0; k < 3; k++) s += (double)a[k] * (double)b[k]; return s; }

struct Pyramid { float sf[16], sigma2[16], inv_sigma2[16], log_sf; int levels; };
Pyramid pyramid(float f, int levels) {                                                        // src/ORBextractor.cc:418-428, src/Frame.cc:93-100
    Pyramid p; p.levels = levels; p.sf[0] = 1.0f; p.sigma2[0] = 1.0f;
    for (int i = 1; i < levels; i++) { p.sf[i] = p.sf[i - 1] * f; p.sigma2[i] = p.sf[i] * p.sf[i]; }
    for (int i = 0; i < levels; i++) p.inv_sigma2[i] = 1.0f / p.sigma2[i];
    p.log_sf = std::log(f);
    return p;
}
int predict_scale(float max_distance, float current, const Pyramid& p) {                        // src/MapPoint.cc:403-435
    const float ratio = max_distance / current;
    int n = (int)std::ceil(std::log(ratio) / p.log_sf);
    if (n < 0) n = 0; else if (n >= p.levels) n = p.levels - 1;
    return n;
}

struct Pose { float R[9], t[3], Ow[3]; };
Pose pose_from_T(const float* T) {                                                            // GetRotation / GetTranslation / GetCameraCenter
    Pose p;
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) p.R[3 * r + c] = T[4 * r + c]; p.t[r] = T[4 * r + 3]; }
    float Rt[9], nRt[9]; transpose33(p.R, Rt);
    for (int i = 0; i < 9; i++) nRt[i] = -Rt[i];
    mul33_31(nRt, p.t, p.Ow);
    return p;
}
Pose pose_from_S(const float* S) {                                                            // src/ORBmatcher.cc:302-307 / 991-996
    Pose p; float sR[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) sR[3 * r + c] = S[4 * r + c];
    const float scw = (float)std::sqrt(dot3(sR, sR));
    for (int i = 0; i < 9; i++) p.R[i] = (float)((double)sR[i] / (double)scw);
    for (int r = 0; r < 3; r++) p.t[r] = (float)((double)S[4 * r + 3] / (double)scw);
    float Rt[9], nRt[9]; transpose33(p.R, Rt);
    for (int i = 0; i < 9; i++) nRt[i] = -Rt[i];
    mul33_31(nRt, p.t, p.Ow);
    return p;
}

struct KFGrid {
    const oracle_keypoint* k; const uint8_t* d; int n; const float* b4; std::vector<int32_t> cs, ci, cand;
    KFGrid(const oracle_keypoint* k_, const uint8_t* d_, int n_, const float* b) : k(k_), d(d_), n(n_), b4(b), cs(64 * 48 + 1), ci(n_ > 0 ? n_ : 1), cand(n_ > 0 ? n_ : 1) {
        oracle_assign_grid(k, n, b4, cs.data(), ci.data());
    }
    int query(float x, float y, float r) { return oracle_keyframe_features_in_area(k, cs.data(), ci.data(), b4, x, y, r, -1, -1, cand.data(), n); }      // KeyFrame.cc:672-718
    bool in_image(float x, float y) const { return x >= (float)(int)b4[0] && x < (float)(int)b4[1] && y >= (float)(int)b4[2] && y < (float)(int)b4[3]; }
};

// the common front of Fuse / Fuse(Scw) / SearchByProjection(Scw): world point -> (u, v, 1/z, predicted level), false when the point is discarded
bool project(const Pose& P, const float* cam4, const KFGrid& G, const Pyramid& py, const float* pos, const float* normal, const float* minmax,
             float& u, float& v, float& invz, int& level) {
    float pc[3]; mul33_31(P.R, pos, pc);
    for (int i = 0; i < 3; i++) pc[i] = pc[i] + P.t[i];
    if (pc[2] < 0.0f) return false;
    invz = 1.0f / pc[2];
    const float x = pc[0] * invz, y = pc[1] * invz;
    u = cam4[0] * x + cam4[2]; v = cam4[1] * y + cam4[3];
    if (!G.in_image(u, v)) return false;
    const float maxD = 1.2f * minmax[1], minD = 0.8f * minmax[0];                              // src/MapPoint.cc:391-401
    float PO[3]; for (int i = 0; i < 3; i++) PO[i] = pos[i] - P.Ow[i];
    const float dist = (float)std::sqrt(dot3(PO, PO));
    if (dist < minD || dist > maxD) return false;
    if (dot3(PO, normal) < 0.5 * dist) return false;
    level = predict_scale(minmax[1], dist, py);
    return true;
}

// best feature of a radius query: level window [level - 1, level], optional reprojection gate of Fuse (chi2 > 0), optional occupancy (skip[idx] != 0)
int best_in_radius(KFGrid& G, const Pyramid& py, float u, float v, float radius, int level, const uint8_t* qd, double chi2, const uint8_t* skip, int& best_dist) {
    const int nc = G.query(u, v, radius);
    int best = -1; best_dist = 256;
    for (int c = 0; c < nc; c++) {
        const int idx = G.cand[c];
        if (skip && skip[idx]) continue;
        const oracle_keypoint& kp = G.k[idx];
        if (kp.octave < level - 1 || kp.octave > level) continue;
        if (chi2 > 0) {
            const float ex = u - kp.x, ey = v - kp.y, e2 = ex * ex + ey * ey;
            if (e2 * py.inv_sigma2[kp.octave] > chi2) continue;
        }
        const int dist = hamming(qd, G.d + 32 * (size_t)idx);
        if (dist < best_dist) { best_dist = dist; best = idx; }
    }
    return best;
}

}  // namespace

extern "C" {

// The device primitive restated: per query (x, y, r, predicted level, descriptor) the best feature of the keyframe inside the window, level filter
// [level - 1, level], optional chi2 gate e2 * invSigma2[octave] > chi2 (Fuse, mono: 5.99), strict '<' (first minimum in grid visit order).
// best_idx / best_dist [nq] (-1 / 256 when nothing qualifies).
void oracle_kf_radius_search(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* q_xyr, const int32_t* q_level,
                             const uint8_t* q_desc, int nq, float scale_factor, int nlevels, double chi2, int32_t* best_idx, int32_t* best_dist) {
    KFGrid G(k, d, n, bounds4);
    const Pyramid py = pyramid(scale_factor, nlevels);
    for (int q = 0; q < nq; q++) {
        int bd;
        best_idx[q] = best_in_radius(G, py, q_xyr[3 * q], q_xyr[3 * q + 1], q_xyr[3 * q + 2], q_level[q], q_desc + 32 * (size_t)q, chi2, nullptr, bd);
        best_dist[q] = bd;
    }
}

// SearchForTriangulation, monocular (mvuRight < 0 everywhere), bOnlyStereo = false.  Arguments as ref_search_for_triangulation.
int oracle_search_for_triangulation(const oracle_keypoint* k1, const uint8_t* d1, const uint8_t* has_mp1, int n1, const int32_t* nodes1, const int32_t* start1,
                                    const int32_t* items1, int nn1, const float* T1, const oracle_keypoint* k2, const uint8_t* d2, const uint8_t* has_mp2, int n2,
                                    const int32_t* nodes2, const int32_t* start2, const int32_t* items2, int nn2, const float* T2, const float* bounds4,
                                    const float* cam4, const float* F12, int check_ori, int32_t* matches12) {
    const Pyramid py = pyramid(1.2f, 8);
    const Pose P1 = pose_from_T(T1), P2 = pose_from_T(T2);
    float C2[3]; mul33_31(P2.R, P1.Ow, C2);
    for (int i = 0; i < 3; i++) C2[i] = C2[i] + P2.t[i];
    const float invz = 1.0f / C2[2];
    const float ex = cam4[0] * C2[0] * invz + cam4[2], ey = cam4[1] * C2[1] * invz + cam4[3];
    int nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<std::vector<int> > rotHist(HISTO_LENGTH);
    const float factor = 1.0f / HISTO_LENGTH;
    int j1 = 0, j2 = 0;
    while (j1 < nn1 && j2 < nn2) {
        if (nodes1[j1] == nodes2[j2]) {
            for (int a = start1[j1]; a < start1[j1 + 1]; a++) {
                const int idx1 = items1[a];
                if (has_mp1[idx1]) continue;
                const oracle_keypoint& kp1 = k1[idx1];
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int b = start2[j2]; b < start2[j2 + 1]; b++) {
                    const int idx2 = items2[b];
                    if (has_mp2[idx2]) continue;                                              // vbMatched2 is never set by the reference (:661-829)
                    const int dist = hamming(d1 + 32 * (size_t)idx1, d2 + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const oracle_keypoint& kp2 = k2[idx2];
                    const float dex = ex - kp2.x, dey = ey - kp2.y;
                    if (dex * dex + dey * dey < 100 * py.sf[kp2.octave]) continue;
                    // CheckDistEpipolarLine (:139-157)
                    const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                    const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                    const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                    const float num = la * kp2.x + lb * kp2.y + lc, den = la * la + lb * lb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (dsqr < 3.84 * py.sigma2[kp2.octave]) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    matches12[idx1] = bestIdx2; nmatches++;
                    if (check_ori) {
                        float rot = kp1.angle - k2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            j1++; j2++;
        } else if (nodes1[j1] < nodes2[j2]) { while (j1 < nn1 && nodes1[j1] < nodes2[j2]) j1++; }      // lower_bound
        else { while (j2 < nn2 && nodes2[j2] < nodes1[j1]) j2++; }
    }
    if (check_ori) {
        int ind1, ind2, ind3;
        three_maxima(rotHist, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) { matches12[idx1] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// Fuse(KeyFrame*, vpMapPoints, th).  Arguments and outputs as ref_fuse (without the query trace).
int oracle_fuse(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* T, const uint8_t* held_state,
                const int32_t* held_nobs, int n_mp, const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc,
                const float* mp_minmax, const int32_t* mp_nobs, float th, int32_t* fused_idx, int32_t* action) {
    KFGrid G(k, d, n, bounds4);
    const Pyramid py = pyramid(1.2f, 8);
    const Pose P = pose_from_T(T);
    // keyframe side: holder[i] = -1 none, -2 - own point, >= 0 map point m added by this call; own points can go bad through Replace
    std::vector<int> holder(n, -1); std::vector<uint8_t> own_bad(n, 0);
    for (int i = 0; i < n; i++) if (held_state[i]) { holder[i] = -2; own_bad[i] = held_state[i] == 2; }
    std::vector<uint8_t> mp_bad(n_mp, 0); std::vector<int> nobs(n_mp, 0);
    for (int m = 0; m < n_mp; m++) { mp_bad[m] = mp_state[m] == 2; nobs[m] = mp_nobs[m]; fused_idx[m] = -1; action[m] = 0; }
    int nFused = 0;
    for (int m = 0; m < n_mp; m++) {
        if (!mp_state[m]) continue;
        if (mp_bad[m] || mp_state[m] == 3) continue;
        float u, v, invz; int level;
        if (!project(P, cam4, G, py, mp_pos + 3 * m, mp_normal + 3 * m, mp_minmax + 2 * m, u, v, invz, level)) continue;
        const float radius = th * py.sf[level];
        int bestDist;
        const int bestIdx = best_in_radius(G, py, u, v, radius, level, mp_desc + 32 * (size_t)m, 5.99, nullptr, bestDist);
        if (bestDist <= TH_LOW) {
            if (holder[bestIdx] == -2) {
                if (!own_bad[bestIdx]) {
                    if (held_nobs[bestIdx] > nobs[m]) { mp_bad[m] = 1; fused_idx[m] = bestIdx; action[m] = 2; }
                    else { own_bad[bestIdx] = 1; fused_idx[m] = bestIdx; action[m] = 3; }
                }
            } else if (holder[bestIdx] >= 0) {
                // a point added earlier in this call sits there now: it is a live pointer like any other (Observations() was bumped by AddObservation)
                const int o = holder[bestIdx];
                if (!mp_bad[o]) {
                    if (nobs[o] > nobs[m]) { mp_bad[m] = 1; fused_idx[m] = -3 - o; action[m] = 2; }
                    else { mp_bad[o] = 1; fused_idx[m] = -3 - o; action[m] = 3; }
                }
            } else { holder[bestIdx] = m; nobs[m]++; fused_idx[m] = bestIdx; action[m] = 1; }
            nFused++;
        }
    }
    return nFused;
}

// Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint).  Arguments and outputs as ref_fuse_sim3 (without the query trace).
int oracle_fuse_sim3(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* S, const uint8_t* held_state,
                     int n_mp, const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc, const float* mp_minmax, float th,
                     int32_t* replace_idx, int32_t* added_idx) {
    KFGrid G(k, d, n, bounds4);
    const Pyramid py = pyramid(1.2f, 8);
    const Pose P = pose_from_S(S);
    std::vector<int> holder(n, -1);
    for (int i = 0; i < n; i++) if (held_state[i]) holder[i] = -2;
    int nFused = 0;
    for (int m = 0; m < n_mp; m++) {
        replace_idx[m] = -1; added_idx[m] = -1;
        if (mp_state[m] == 2 || mp_state[m] == 3) continue;
        float u, v, invz; int level;
        if (!project(P, cam4, G, py, mp_pos + 3 * m, mp_normal + 3 * m, mp_minmax + 2 * m, u, v, invz, level)) continue;
        const float radius = th * py.sf[level];
        int bestDist;
        const int bestIdx = best_in_radius(G, py, u, v, radius, level, mp_desc + 32 * (size_t)m, 0, nullptr, bestDist);
        if (bestIdx >= 0 && bestDist <= TH_LOW) {
            if (holder[bestIdx] == -2) { if (held_state[bestIdx] != 2) replace_idx[m] = bestIdx; }
            else if (holder[bestIdx] >= 0) replace_idx[m] = -3 - holder[bestIdx];             // a point added earlier in this call (never bad here)
            else { holder[bestIdx] = m; added_idx[m] = bestIdx; }
            nFused++;
        }
    }
    return nFused;
}

// SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th).  Arguments as ref_search_by_projection_loop (without the query trace).
int oracle_search_by_projection_loop(const oracle_keypoint* k, const uint8_t* d, int n, const float* bounds4, const float* cam4, const float* S, int n_mp,
                                     const uint8_t* mp_state, const float* mp_pos, const float* mp_normal, const uint8_t* mp_desc, const float* mp_minmax, int th,
                                     int32_t* matched) {
    KFGrid G(k, d, n, bounds4);
    const Pyramid py = pyramid(1.2f, 8);
    const Pose P = pose_from_S(S);
    std::vector<uint8_t> found(n_mp, 0), occ(n, 0);
    for (int i = 0; i < n; i++) { if (matched[i] >= 0) found[matched[i]] = 1; occ[i] = matched[i] != -1; }
    int nmatches = 0;
    for (int m = 0; m < n_mp; m++) {
        if (mp_state[m] == 2 || found[m]) continue;
        float u, v, invz; int level;
        if (!project(P, cam4, G, py, mp_pos + 3 * m, mp_normal + 3 * m, mp_minmax + 2 * m, u, v, invz, level)) continue;
        const float radius = th * py.sf[level];
        int bestDist;
        const int bestIdx = best_in_radius(G, py, u, v, radius, level, mp_desc + 32 * (size_t)m, 0, occ.data(), bestDist);
        if (bestDist <= TH_LOW) { matched[bestIdx] = m; occ[bestIdx] = 1; nmatches++; }
    }
    return nmatches;
}

// SearchBySim3.  Arguments as ref_search_by_sim3 (without the query trace).
int oracle_search_by_sim3(const oracle_keypoint* k1, const uint8_t* d1, int n1, const float* T1, const uint8_t* mp1_state, const float* mp1_pos,
                          const uint8_t* mp1_desc, const float* mp1_minmax, const oracle_keypoint* k2, const uint8_t* d2, int n2, const float* T2,
                          const uint8_t* mp2_state, const float* mp2_pos, const uint8_t* mp2_desc, const float* mp2_minmax, const float* bounds4,
                          const float* cam4, float s12, const float* R12, const float* t12, float th, int32_t* matches12) {
    KFGrid G1(k1, d1, n1, bounds4), G2(k2, d2, n2, bounds4);
    const Pyramid py = pyramid(1.2f, 8);
    const Pose P1 = pose_from_T(T1), P2 = pose_from_T(T2);
    float sR12[9], sR21[9], R12t[9], t21[3], nsR21[9];
    for (int i = 0; i < 9; i++) sR12[i] = (float)((double)s12 * (double)R12[i]);
    transpose33(R12, R12t);
    const double inv_s = 1.0 / (double)s12;
    for (int i = 0; i < 9; i++) { sR21[i] = (float)(inv_s * (double)R12t[i]); nsR21[i] = -sR21[i]; }
    mul33_31(nsR21, t12, t21);
    std::vector<uint8_t> done1(n1, 0), done2(n2, 0);
    for (int i = 0; i < n1; i++) if (matches12[i] >= 0) { done1[i] = 1; if (mp2_state[matches12[i]] && matches12[i] < n2) done2[matches12[i]] = 1; }
    std::vector<int> m1(n1, -1), m2(n2, -1);
    for (int dir = 0; dir < 2; dir++) {
        const int n = dir ? n2 : n1;
        const uint8_t* st = dir ? mp2_state : mp1_state; const float* pos = dir ? mp2_pos : mp1_pos; const uint8_t* desc = dir ? mp2_desc : mp1_desc;
        const float* mm = dir ? mp2_minmax : mp1_minmax;
        const Pose& Pa = dir ? P2 : P1; const float* sR = dir ? sR12 : sR21; const float* tt = dir ? t12 : t21;
        KFGrid& G = dir ? G1 : G2; std::vector<uint8_t>& done = dir ? done2 : done1; std::vector<int>& out = dir ? m2 : m1;
        for (int i = 0; i < n; i++) {
            if (!st[i] || done[i] || st[i] == 2) continue;
            float pa[3], pb[3];
            mul33_31(Pa.R, pos + 3 * i, pa); for (int j = 0; j < 3; j++) pa[j] = pa[j] + Pa.t[j];
            mul33_31(sR, pa, pb); for (int j = 0; j < 3; j++) pb[j] = pb[j] + tt[j];
            if (pb[2] < 0.0f) continue;
            const float invz = 1.0f / pb[2], x = pb[0] * invz, y = pb[1] * invz;
            const float u = cam4[0] * x + cam4[2], v = cam4[1] * y + cam4[3];
            if (!G.in_image(u, v)) continue;
            const float maxD = 1.2f * mm[2 * i + 1], minD = 0.8f * mm[2 * i];
            const float dist3D = (float)std::sqrt(dot3(pb, pb));
            if (dist3D < minD || dist3D > maxD) continue;
            const int level = predict_scale(mm[2 * i + 1], dist3D, py);
            const float radius = th * py.sf[level];
            int bestDist;
            const int bestIdx = best_in_radius(G, py, u, v, radius, level, desc + 32 * (size_t)i, 0, nullptr, bestDist);
            if (bestIdx >= 0 && bestDist <= TH_HIGH) out[i] = bestIdx;
        }
    }
    int nFound = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        const int idx2 = m1[i1];
        if (idx2 >= 0 && m2[idx2] == i1) { matches12[i1] = idx2; nFound++; }
    }
    return nFound;
}

}  // extern "C"
