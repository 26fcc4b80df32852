// oracle/orb_oracle.cpp -- TEST INFRASTRUCTURE (CPU checker), not product code.
//
// Restatement of ORB_SLAM2::ORBextractor (reference src/ORBextractor.cc) over plain arrays, written
// independently of the reference's containers so that it can travel to the GPU box (where
// /root/reference does not exist).  It is validated against the reference's own source compiled on
// the shim (oracle/_ref/libref_orb.so, tests/test_oracle_vs_ref.py) and against cv2 golden vectors.
//
// Canonical choices where the reference is build/allocator dependent (DESIGN.md "determinism"):
//  * ORBextractor.cc:684 sorts pair<int,ExtractorNode*>: ties by heap address.  Canonical: ties by
//    creation order, later-created node = greater.
//  * ORBextractor.cc:113 (float)cos/sin of a float resolve to libm cosf/sinf (not correctly rounded,
//    libm-version dependent).  Canonical: (float)cos((double)angle), (float)sin((double)angle).
#include "oracle.h"
#include "cvprim.h"
#include <cstdio>
#include <thread>
#include <atomic>

using namespace cvprim;

namespace {

const int kPatch = 31, kHalfPatch = 15, kEdge = 19;   // ORBextractor.cc:71-73

static const int8_t kPattern[256 * 4] = {
#include "../orb_slam2_aruco_b200/csrc/orb_pattern.inc"
};

struct Params {
    int nfeatures, nlevels, ini_th, min_th;
    std::vector<float> sf, inv_sf;
    std::vector<int> quota;
    int umax[16];
};

// ORBextractor.cc:410-470
void make_params(Params& p, int nfeatures, float scale, int nlevels, int ini_th, int min_th) {
    p.nfeatures = nfeatures; p.nlevels = nlevels; p.ini_th = ini_th; p.min_th = min_th;
    double scale_d = scale;                       // member is double (ORBextractor.h:98), set from a float
    p.sf.assign(nlevels, 1.f); p.inv_sf.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; i++) p.sf[i] = (float)(p.sf[i - 1] * scale_d);   // float*double -> double -> float
    for (int i = 0; i < nlevels; i++) p.inv_sf[i] = 1.0f / p.sf[i];
    p.quota.assign(nlevels, 0);
    float factor = (float)(1.0f / scale_d);
    float nd = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        p.quota[l] = round_half_even_f(nd);
        sum += p.quota[l];
        nd *= factor;
    }
    p.quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // circular patch row ends (ORBextractor.cc:454-469)
    int vmax = floor_i(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = ceil_i(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; v++) p.umax[v] = round_half_even(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (p.umax[v0] == p.umax[v0 + 1]) ++v0;
        p.umax[v] = v0;
        ++v0;
    }
}

struct Level { int w, h; std::vector<u8> img; };   // no border: nothing on the mono path reads it

// ORBextractor.cc:1107-1132 (border omitted; see oracle_border_reflect101 for the full buffer)
void build_pyramid(const Params& p, const u8* img, int w, int h, int stride, std::vector<Level>& pyr) {
    pyr.resize(p.nlevels);
    for (int l = 0; l < p.nlevels; l++) {
        float s = p.inv_sf[l];
        Level& L = pyr[l];
        L.w = round_half_even_f((float)w * s);
        L.h = round_half_even_f((float)h * s);
        L.img.resize((size_t)L.w * L.h);
        if (l == 0) for (int y = 0; y < h; y++) memcpy(&L.img[(size_t)y * w], img + (size_t)y * stride, w);
        else resize_linear_u8(pyr[l - 1].img.data(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].w, L.img.data(), L.w, L.h, L.w);
    }
}

struct Cand { float x, y; int score; };   // coordinates relative to (minBorderX, minBorderY)

// per-cell FAST with threshold fallback: ORBextractor.cc:765-829
void level_candidates(const Params& p, const Level& L, std::vector<Cand>& out) {
    out.clear();
    const float W = 30;
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<FastKp> cell;
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            const u8* roi = &L.img[(size_t)(int)iniY * L.w + (int)iniX];
            int rw = (int)maxX - (int)iniX, rh = (int)maxY - (int)iniY;
            fast9_16_nms(roi, rw, rh, L.w, p.ini_th, cell);
            if (cell.empty()) fast9_16_nms(roi, rw, rh, L.w, p.min_th, cell);
            for (size_t k = 0; k < cell.size(); k++)
                out.push_back(Cand{(float)cell[k].x + j * wCell, (float)cell[k].y + i * hCell, cell[k].score});
        }
    }
}

// ---------------------------------------------------------------------------------------------
// DistributeOctTree (ORBextractor.cc:539-763) + DivideNode (481-537) on index arrays.
// Nodes live in a pool; the std::list is a doubly linked list of pool indices; a node's keys are
// a contiguous range of `perm` (DivideNode's four stable push_back loops == a stable 4-way
// partition of that range in child order n1,n2,n3,n4).
// ---------------------------------------------------------------------------------------------
struct Node { int x0, x1, y0, y1, beg, end, prev, next; bool no_more; };

struct Tree {
    const std::vector<Cand>& keys;
    std::vector<int> perm, tmp;
    std::vector<Node> pool;
    int head, tail, size;
    explicit Tree(const std::vector<Cand>& k) : keys(k), head(-1), tail(-1), size(0) {}

    int push_front(const Node& n) {
        int id = (int)pool.size();
        pool.push_back(n);
        pool[id].prev = -1; pool[id].next = head;
        if (head >= 0) pool[head].prev = id; else tail = id;
        head = id; size++;
        return id;
    }
    int push_back(const Node& n) {
        int id = (int)pool.size();
        pool.push_back(n);
        pool[id].next = -1; pool[id].prev = tail;
        if (tail >= 0) pool[tail].next = id; else head = id;
        tail = id; size++;
        return id;
    }
    int erase(int id) {   // returns successor
        int p = pool[id].prev, n = pool[id].next;
        if (p >= 0) pool[p].next = n; else head = n;
        if (n >= 0) pool[n].prev = p; else tail = p;
        size--;
        return n;
    }
    // DivideNode: children boxes + stable partition; returns child descriptors c[0..3] (may be empty)
    void divide(int id, Node c[4]) {
        const Node nd = pool[id];
        const int halfX = (int)std::ceil((float)(nd.x1 - nd.x0) / 2);
        const int halfY = (int)std::ceil((float)(nd.y1 - nd.y0) / 2);
        const int mx = nd.x0 + halfX, my = nd.y0 + halfY;
        c[0].x0 = nd.x0; c[0].x1 = mx;    c[0].y0 = nd.y0; c[0].y1 = my;      // n1: upper-left
        c[1].x0 = mx;    c[1].x1 = nd.x1; c[1].y0 = nd.y0; c[1].y1 = my;      // n2: upper-right
        c[2].x0 = nd.x0; c[2].x1 = mx;    c[2].y0 = my;    c[2].y1 = nd.y1;   // n3: lower-left
        c[3].x0 = mx;    c[3].x1 = nd.x1; c[3].y0 = my;    c[3].y1 = nd.y1;   // n4: lower-right
        int cnt[4] = {0, 0, 0, 0};
        tmp.resize(perm.size());
        for (int i = nd.beg; i < nd.end; i++) {
            const Cand& k = keys[perm[i]];
            int q = (k.x < (float)mx ? 0 : 1) + (k.y < (float)my ? 0 : 2);
            cnt[q]++;
        }
        int ofs[4]; ofs[0] = nd.beg; for (int q = 1; q < 4; q++) ofs[q] = ofs[q - 1] + cnt[q - 1];
        for (int q = 0; q < 4; q++) { c[q].beg = ofs[q]; c[q].end = ofs[q] + cnt[q]; c[q].no_more = (cnt[q] == 1); }
        for (int i = nd.beg; i < nd.end; i++) {
            const Cand& k = keys[perm[i]];
            int q = (k.x < (float)mx ? 0 : 1) + (k.y < (float)my ? 0 : 2);
            tmp[ofs[q]++] = perm[i];
        }
        for (int i = nd.beg; i < nd.end; i++) perm[i] = tmp[i];
    }
};

void distribute(const std::vector<Cand>& keys, int minX, int maxX, int minY, int maxY, int N, std::vector<int>& result) {
    result.clear();
    Tree t(keys);
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return;   // portrait frames with h > 2w divide by zero in the reference; not supported
    const float hX = (float)(maxX - minX) / nIni;
    // initial nodes and key assignment (ORBextractor.cc:553-570); stable counting sort by node
    std::vector<int> cnt(nIni + 1, 0), which(keys.size());
    for (size_t i = 0; i < keys.size(); i++) { which[i] = (int)(keys[i].x / hX); cnt[which[i] + 1]++; }
    for (int i = 0; i < nIni; i++) cnt[i + 1] += cnt[i];
    t.perm.resize(keys.size());
    {
        std::vector<int> o(cnt.begin(), cnt.end() - 1);
        for (size_t i = 0; i < keys.size(); i++) t.perm[o[which[i]]++] = (int)i;
    }
    for (int i = 0; i < nIni; i++) {
        Node n;
        n.x0 = (int)(hX * (float)i); n.x1 = (int)(hX * (float)(i + 1)); n.y0 = 0; n.y1 = maxY - minY;
        n.beg = cnt[i]; n.end = cnt[i + 1]; n.no_more = false;
        t.push_back(n);
    }
    for (int it = t.head; it >= 0;) {          // ORBextractor.cc:572-585
        Node& n = t.pool[it];
        if (n.end - n.beg == 1) { n.no_more = true; it = n.next; }
        else if (n.end == n.beg) it = t.erase(it);
        else it = n.next;
    }
    bool finish = false;
    std::vector<std::pair<int, int> > big;    // (key count, node id); ids grow with creation order
    while (!finish) {
        int prev_size = t.size, n_expand = 0;
        big.clear();
        for (int it = t.head; it >= 0;) {
            if (t.pool[it].no_more) { it = t.pool[it].next; continue; }
            Node c[4];
            t.divide(it, c);
            for (int q = 0; q < 4; q++) {
                int k = c[q].end - c[q].beg;
                if (k > 0) {
                    int id = t.push_front(c[q]);
                    if (k > 1) { n_expand++; big.push_back(std::make_pair(k, id)); }
                }
            }
            it = t.erase(it);
        }
        if (t.size >= N || t.size == prev_size) finish = true;
        else if (t.size + n_expand * 3 > N) {
            while (!finish) {
                prev_size = t.size;
                std::vector<std::pair<int, int> > prev = big;
                big.clear();
                std::sort(prev.begin(), prev.end());   // (count, creation id) ascending == canonical tie-break
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node c[4];
                    t.divide(prev[j].second, c);
                    for (int q = 0; q < 4; q++) {
                        int k = c[q].end - c[q].beg;
                        if (k > 0) {
                            int id = t.push_front(c[q]);
                            if (k > 1) big.push_back(std::make_pair(k, id));
                        }
                    }
                    t.erase(prev[j].second);
                    if (t.size >= N) break;
                }
                if (t.size >= N || t.size == prev_size) finish = true;
            }
        }
    }
    // best response per node, first maximum wins (ORBextractor.cc:742-760)
    for (int it = t.head; it >= 0; it = t.pool[it].next) {
        const Node& n = t.pool[it];
        int best = t.perm[n.beg];
        for (int i = n.beg + 1; i < n.end; i++)
            if (keys[t.perm[i]].score > keys[best].score) best = t.perm[i];
        result.push_back(best);
    }
}

// IC_Angle: ORBextractor.cc:77-104 (image without border: every keypoint is >= 19 px inside)
float ic_angle(const Params& p, const Level& L, float px, float py) {
    int m01 = 0, m10 = 0;
    const int step = L.w;
    const u8* c = &L.img[(size_t)round_half_even_f(py) * step + round_half_even_f(px)];
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vs = 0, d = p.umax[v];
        for (int u = -d; u <= d; ++u) {
            int a = c[u + v * step], b = c[u - v * step];
            vs += a - b;
            m10 += u * (a + b);
        }
        m01 += v * vs;
    }
    return fast_atan2_deg((float)m01, (float)m10);
}

// computeOrbDescriptor: ORBextractor.cc:107-147, canonical cos/sin (see header)
void orb_descriptor(const u8* blurred, int step, float kx, float ky, float angle_deg, u8* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angle_deg * factorPI;
    float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    const u8* c = blurred + (size_t)round_half_even_f(ky) * step + round_half_even_f(kx);
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; i++) {
        int val = 0;
        for (int k = 0; k < 8; k++, pat += 4) {
            float x0 = pat[0], y0 = pat[1], x1 = pat[2], y1 = pat[3];
            int t0 = c[round_half_even_f(x0 * b + y0 * a) * step + round_half_even_f(x0 * a - y0 * b)];
            int t1 = c[round_half_even_f(x1 * b + y1 * a) * step + round_half_even_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (u8)val;
    }
}

int extract(const Params& p, const u8* img, int w, int h, int stride, oracle_keypoint* kps, u8* desc, int cap) {
    if (!img || w <= 0 || h <= 0) return 0;                // empty image: silent return (ORBextractor.cc:1046)
    std::vector<Level> pyr;
    build_pyramid(p, img, w, h, stride, pyr);
    int total = 0;
    std::vector<Cand> cand;
    std::vector<int> keep;
    std::vector<u8> blurred;
    for (int l = 0; l < p.nlevels; l++) {
        const Level& L = pyr[l];
        level_candidates(p, L, cand);
        const int minB = kEdge - 3;
        distribute(cand, minB, L.w - kEdge + 3, minB, L.h - kEdge + 3, p.quota[l], keep);
        if (keep.empty()) continue;
        if (total + (int)keep.size() > cap) return -1;
        const int scaled_patch = (int)(kPatch * p.sf[l]);
        blurred.resize((size_t)L.w * L.h);
        gaussian_blur7_s2(L.img.data(), L.w, L.h, L.w, blurred.data(), L.w);
        for (size_t i = 0; i < keep.size(); i++) {
            const Cand& c = cand[keep[i]];
            oracle_keypoint& k = kps[total];
            float x = c.x + minB, y = c.y + minB;
            k.size = (float)scaled_patch; k.response = (float)c.score; k.octave = l; k.class_id = -1;
            k.angle = ic_angle(p, L, x, y);
            orb_descriptor(blurred.data(), L.w, x, y, k.angle, desc + (size_t)total * 32);
            if (l != 0) { float s = p.sf[l]; x *= s; y *= s; }
            k.x = x; k.y = y;
            total++;
        }
    }
    return total;
}

}  // namespace

extern "C" {

void oracle_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep) {
    resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void oracle_border_reflect101(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep, int border) {
    copy_make_border_reflect101(src, w, h, sstep, dst, dstep, border, border, border, border);
}
void oracle_gaussian_blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep) {
    gaussian_blur7_s2(src, w, h, sstep, dst, dstep);
}
int oracle_fast_nms(const uint8_t* img, int w, int h, int step, int thr, int32_t* xys, int cap) {
    std::vector<FastKp> v;
    fast9_16_nms(img, w, h, step, thr, v);
    if ((int)v.size() > cap) return -1;
    for (size_t i = 0; i < v.size(); i++) { xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score; }
    return (int)v.size();
}
float oracle_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }

int oracle_orb_levels(int w, int h, int nfeatures, float scale, int nlevels, int32_t* lw, int32_t* lh, int32_t* quota,
                      float* scale_factor) {
    Params p; make_params(p, nfeatures, scale, nlevels, 20, 7);
    for (int l = 0; l < nlevels; l++) {
        lw[l] = round_half_even_f((float)w * p.inv_sf[l]);
        lh[l] = round_half_even_f((float)h * p.inv_sf[l]);
        quota[l] = p.quota[l];
        if (scale_factor) scale_factor[l] = p.sf[l];
    }
    return 0;
}
int oracle_orb_pyramid_level(const uint8_t* img, int w, int h, int stride, float scale, int nlevels, int level, uint8_t* out) {
    Params p; make_params(p, 1000, scale, nlevels, 20, 7);
    std::vector<Level> pyr;
    build_pyramid(p, img, w, h, stride, pyr);
    memcpy(out, pyr[level].img.data(), pyr[level].img.size());
    return 0;
}
int oracle_orb_candidates(const uint8_t* img, int w, int h, int stride, int nfeatures, float scale, int nlevels,
                          int ini_th, int min_th, int level, int32_t* xys, int cap) {
    Params p; make_params(p, nfeatures, scale, nlevels, ini_th, min_th);
    std::vector<Level> pyr;
    build_pyramid(p, img, w, h, stride, pyr);
    std::vector<Cand> c;
    level_candidates(p, pyr[level], c);
    if ((int)c.size() > cap) return -1;
    for (size_t i = 0; i < c.size(); i++) { xys[3 * i] = (int)c[i].x; xys[3 * i + 1] = (int)c[i].y; xys[3 * i + 2] = c[i].score; }
    return (int)c.size();
}
int oracle_orb_extract(const uint8_t* img, int w, int h, int stride, int nfeatures, float scale, int nlevels,
                       int ini_th, int min_th, oracle_keypoint* kps, uint8_t* desc, int cap) {
    Params p; make_params(p, nfeatures, scale, nlevels, ini_th, min_th);
    return extract(p, img, w, h, stride, kps, desc, cap);
}
int oracle_orb_extract_batch(const uint8_t* imgs, int n, int w, int h, int row_stride, long frame_stride,
                             int nfeatures, float scale, int nlevels, int ini_th, int min_th,
                             oracle_keypoint* kps, uint8_t* desc, int32_t* counts, int cap, int nthreads) {
    Params p; make_params(p, nfeatures, scale, nlevels, ini_th, min_th);
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (int f; (f = next.fetch_add(1)) < n;)
            counts[f] = extract(p, imgs + (size_t)f * frame_stride, w, h, row_stride, kps + (size_t)f * cap, desc + (size_t)f * cap * 32, cap);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
    work();
    for (size_t t = 0; t < pool.size(); t++) pool[t].join();
    return 0;
}

}  // extern "C"
