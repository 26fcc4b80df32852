// oracle/ref_map_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN src/Map.cc (Map::Save / Map::Load, :219-533), compiled unmodified from
// /root/reference on oracle/mapshim (data stand-ins for MapPoint / KeyFrame / InitKeyFrame / Converter) into oracle/_ref/libref_map.so (oracle/Makefile).
//   ref_map_save: builds a map from flat arrays, calls Map::Save(filename)                      -> the bytes orb_slam2_aruco_b200/mapfile.py must parse
//   ref_map_load: calls Map::Load(filename, &settings), dumps what the reference reconstructed  -> the bytes mapfile.py writes must load to the same map
#include "Map.h"
#include <algorithm>
#include <cstdint>
#include <cstring>

using namespace ORB_SLAM2;

// Canonicalisation (documented in DESIGN.md): Map::Load resolves every keypoint's map point index through GetAllMapPoints() - a std::set<MapPoint*>, i.e. in
// POINTER order (src/Map.cc:352, 486, 508) - and Map::Save writes its sets in pointer order too (:228, 240), so what the reference reads and writes
// depends on the addresses malloc happens to return.  Inside this shared object operator new is a monotonic arena (local symbols: version script), which
// makes pointer order == creation order == file order, the only order a file can pin.
#include <cstdio>
#include <cstdlib>
namespace {
struct NewArena { char* base = nullptr; size_t cap = 0, used = 0; };
NewArena g_new;
}
void* operator new(size_t n) {
    if (!g_new.base) { g_new.cap = (size_t)1 << 30; g_new.base = (char*)malloc(g_new.cap); g_new.used = 0; }     // virtual; only touched pages are committed
    const size_t need = (n + 15) & ~(size_t)15;
    if (!g_new.base || g_new.used + need > g_new.cap) { fprintf(stderr, "ref_map_wrap: arena exhausted\n"); abort(); }
    void* p = g_new.base + g_new.used;
    g_new.used += need;
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}

namespace {
// one arena per call: std::set<MapPoint*> / std::set<KeyFrame*> iterate in POINTER order (src/Map.cc:228, 240); objects carved from one ascending block keep
// creation order, the canonical choice (an allocator that hands out descending addresses would save the same map in reverse order)
template <class T> struct Arena {
    std::vector<unsigned char> buf; size_t used = 0;
    explicit Arena(size_t n) : buf(n * sizeof(T) + 64) {}
    void* take() { void* p = buf.data() + used; used += sizeof(T); return p; }
};
}

extern "C" {

// Keyframe k: id, timestamp, Tcw (row-major 4 x 4), n_kp[k] keypoints (x y size angle response as 5 floats + octave), descriptors, map point INDEX per keypoint
// (-1 none; index into the points as passed), parent (index of a keyframe or -1), connections (index, weight) as CSR.
int ref_map_save(const char* filename, int n_mp, const uint64_t* mp_id, const float* mp_pos, int n_kf, const uint64_t* kf_id, const double* kf_time,
                 const float* kf_T, const int32_t* n_kp, const float* kp5, const int32_t* kp_octave, const uint8_t* desc, const int64_t* kp_mp,
                 const int32_t* kf_parent, const int32_t* con_ofs, const int32_t* con_kf, const int32_t* con_w) {
    g_new.used = 0;                                                  // a fresh address sequence for every call
    Map map;
    Arena<MapPoint> amp(n_mp); Arena<KeyFrame> akf(n_kf);
    std::vector<MapPoint*> mps(n_mp);
    for (int i = 0; i < n_mp; i++) {
        cv::Mat pos(3, 1, CV_32F);
        for (int j = 0; j < 3; j++) pos.at<float>(j) = mp_pos[3 * i + j];
        mps[i] = new (amp.take()) MapPoint(pos, &map);
        mps[i]->mnId = mp_id[i];
        map.AddMapPoint(mps[i]);
    }
    std::vector<KeyFrame*> kfs(n_kf);
    size_t o = 0;
    for (int k = 0; k < n_kf; k++) {
        KeyFrame* kf = new (akf.take()) KeyFrame();
        kf->mnId = kf_id[k]; kf->mTimeStamp = kf_time[k]; kf->N = n_kp[k];
        kf->Tcw = cv::Mat(4, 4, CV_32F, (void*)(kf_T + 16 * k)).clone();
        kf->mvKeys.resize(n_kp[k]); kf->mvpMapPoints.assign(n_kp[k], (MapPoint*)NULL);
        kf->mDescriptors = cv::Mat(n_kp[k], 32, CV_8U);
        for (int i = 0; i < n_kp[k]; i++, o++) {
            cv::KeyPoint& kp = kf->mvKeys[i];
            kp.pt.x = kp5[5 * o]; kp.pt.y = kp5[5 * o + 1]; kp.size = kp5[5 * o + 2]; kp.angle = kp5[5 * o + 3]; kp.response = kp5[5 * o + 4]; kp.octave = kp_octave[o];
            memcpy(kf->mDescriptors.ptr(i), desc + 32 * o, 32);
            if (kp_mp[o] >= 0) kf->mvpMapPoints[i] = mps[kp_mp[o]];
        }
        kfs[k] = kf;
        map.AddKeyFrame(kf);
    }
    for (int k = 0; k < n_kf; k++) {
        if (kf_parent[k] >= 0) kfs[k]->ChangeParent(kfs[kf_parent[k]]);
        for (int c = con_ofs[k]; c < con_ofs[k + 1]; c++) kfs[k]->AddConnection(kfs[con_kf[c]], con_w[c]);
    }
    map.Save(filename);
    for (int i = 0; i < n_mp; i++) mps[i]->~MapPoint();
    for (int k = 0; k < n_kf; k++) kfs[k]->~KeyFrame();
    return 0;
}

// Loads `filename` with the reference's Map::Load and reports the reconstructed map (std::set order: callers match points and keyframes by id).  Outputs as in
// ref_map_save; kp_mp holds the mnId of the point each keypoint references (-1 none); obs_count [n_mp] = observations per point (a keyframe counts once); hooks [4] = how often Load called
// UndistortKeyPoints, AssignFeaturesToGrid, ComputeBoW and ComputeDistinctiveDescriptors in total.  Returns 0, or -1 when an array is too small.
int ref_map_load(const char* filename, int cap_mp, int32_t* n_mp, uint64_t* mp_id, float* mp_pos, int32_t* obs_count, int cap_kf, int32_t* n_kf, uint64_t* kf_id,
                 double* kf_time, float* kf_T, int32_t* n_kp, int cap_kp, float* kp5, int32_t* kp_octave, uint8_t* desc, int64_t* kp_mp, int64_t* kf_parent_id,
                 int32_t* con_ofs, int cap_con, uint64_t* con_id, int32_t* con_w, int32_t* hooks) {
    g_new.used = 0;                                                  // a fresh address sequence for every call
    Map map;
    SystemSetting s;
    map.Load(filename, &s);
    std::vector<MapPoint*> mps = map.GetAllMapPoints();
    std::vector<KeyFrame*> kfs = map.GetAllKeyFrames();
    *n_mp = (int)mps.size(); *n_kf = (int)kfs.size();
    if ((int)mps.size() > cap_mp || (int)kfs.size() > cap_kf) return -1;
    hooks[0] = hooks[1] = hooks[2] = hooks[3] = 0;
    for (size_t i = 0; i < mps.size(); i++) {
        mp_id[i] = mps[i]->mnId;
        const cv::Mat p = mps[i]->GetWorldPos();
        for (int j = 0; j < 3; j++) mp_pos[3 * i + j] = p.at<float>(j);
        obs_count[i] = (int32_t)mps[i]->obs.size();
        hooks[3] += mps[i]->distinctive;
    }
    size_t o = 0; int nc = 0;
    con_ofs[0] = 0;
    for (size_t k = 0; k < kfs.size(); k++) {
        KeyFrame* kf = kfs[k];
        kf_id[k] = kf->mnId; kf_time[k] = kf->mTimeStamp; n_kp[k] = kf->N;
        for (int i = 0; i < 16; i++) kf_T[16 * k + i] = kf->Tcw.at<float>(i / 4, i % 4);
        hooks[2] += kf->bow;
        if ((int)(o + kf->N) > cap_kp) return -1;
        for (int i = 0; i < kf->N; i++, o++) {
            const cv::KeyPoint& kp = kf->mvKeys[i];
            kp5[5 * o] = kp.pt.x; kp5[5 * o + 1] = kp.pt.y; kp5[5 * o + 2] = kp.size; kp5[5 * o + 3] = kp.angle; kp5[5 * o + 4] = kp.response; kp_octave[o] = kp.octave;
            memcpy(desc + 32 * o, kf->mDescriptors.ptr(i), 32);
            kp_mp[o] = kf->mvpMapPoints[i] ? (int64_t)kf->mvpMapPoints[i]->mnId : -1;
        }
        kf_parent_id[k] = kf->parent ? (int64_t)kf->parent->mnId : -1;
        for (size_t c = 0; c < kf->connections.size(); c++, nc++) {
            if (nc >= cap_con) return -1;
            con_id[nc] = kf->connections[c].first->mnId; con_w[nc] = kf->connections[c].second;
        }
        con_ofs[k + 1] = nc;
    }
    return 0;
}

}  // extern "C"
