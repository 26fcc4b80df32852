// oracle/cvprim.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the OpenCV primitives the reference's hot path calls.  OpenCV is an
// un-vendored system dependency of the reference (CMakeLists.txt:31-37, 3.4.x in the author's
// build); it is absent from /root/reference and from this image's C++ toolchain, so the
// published algorithms are restated here as plain integer / float32 models and pinned
// bit-exactly against python cv2 4.13 golden vectors (tests/golden/, made by
// tests/golden/make_golden.py).  Call sites in the reference:
//   resize            src/ORBextractor.cc:1120
//   copyMakeBorder    src/ORBextractor.cc:1122,1127
//   FAST              src/ORBextractor.cc:809,814
//   GaussianBlur      src/ORBextractor.cc:1086
//   fastAtan2         src/ORBextractor.cc:103
//   cvRound           src/ORBextractor.cc:82,118-120,440,469,1112
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>

namespace cvprim {

typedef unsigned char u8;

// cvRound(float/double): round half to even (SSE cvtss2si / lrint under the default mode).
static inline int round_half_even(double v) { return (int)std::nearbyint(v); }
static inline int round_half_even_f(float v) { return (int)std::nearbyintf(v); }
static inline int floor_i(double v) { int i = (int)v; return i - (i > v); }
static inline int ceil_i(double v) { int i = (int)v; return i + (i < v); }

// BORDER_REFLECT_101 index map: -k -> k, (n-1)+k -> (n-1)-k   (n > 1; repeated for big k)
static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
    return p;
}

// ---------------------------------------------------------------------------------------------
// cv::resize(u8, INTER_LINEAR), single channel (reference call: ORBextractor.cc:1120).
// Fixed point: 11-bit coefficients; horizontal pass in int, vertical pass with the
// ((b*(H>>4))>>16) products and +2 >>2 rounding used by OpenCV's 8-bit linear path.
// ---------------------------------------------------------------------------------------------
struct ResizeTab { std::vector<int> ofs; std::vector<short> c0, c1; };
static inline void resize_tab(int ssize, int dsize, ResizeTab& t) {
    t.ofs.resize(dsize); t.c0.resize(dsize); t.c1.resize(dsize);
    double scale = (double)ssize / dsize;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = floor_i(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }   // second tap gets weight 0 and is clamped
        t.ofs[d] = s;
        t.c0[d] = (short)round_half_even((1.f - f) * 2048.f);
        t.c1[d] = (short)round_half_even(f * 2048.f);
    }
}
static inline void resize_linear_u8(const u8* src, int sw, int sh, size_t sstep,
                                    u8* dst, int dw, int dh, size_t dstep) {
    ResizeTab tx, ty;
    resize_tab(sw, dw, tx);
    resize_tab(sh, dh, ty);
    std::vector<int> row0(dw), row1(dw);
    int have0 = -1, have1 = -1;
    for (int y = 0; y < dh; y++) {
        int sy0 = ty.ofs[y], sy1 = std::min(sy0 + 1, sh - 1);
        // horizontal pass for the two source rows (cached between consecutive output rows)
        if (have1 == sy0) { row0.swap(row1); std::swap(have0, have1); }
        if (have0 != sy0) {
            const u8* s = src + (size_t)sy0 * sstep;
            for (int x = 0; x < dw; x++) {
                int sx = tx.ofs[x], sx1 = std::min(sx + 1, sw - 1);
                row0[x] = s[sx] * tx.c0[x] + s[sx1] * tx.c1[x];
            }
            have0 = sy0;
        }
        if (have1 != sy1) {
            if (sy1 == sy0) { row1 = row0; }
            else {
                const u8* s = src + (size_t)sy1 * sstep;
                for (int x = 0; x < dw; x++) {
                    int sx = tx.ofs[x], sx1 = std::min(sx + 1, sw - 1);
                    row1[x] = s[sx] * tx.c0[x] + s[sx1] * tx.c1[x];
                }
            }
            have1 = sy1;
        }
        int b0 = ty.c0[y], b1 = ty.c1[y];
        u8* d = dst + (size_t)y * dstep;
        for (int x = 0; x < dw; x++)
            d[x] = (u8)((((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------------------------------------
// cv::copyMakeBorder(..., BORDER_REFLECT_101) into a (w+l+r) x (h+t+b) buffer.  src may alias
// the interior of dst (the reference fills the border around an ROI in place, ORBextractor.cc:1122).
// ---------------------------------------------------------------------------------------------
static inline void copy_make_border_reflect101(const u8* src, int w, int h, size_t sstep,
                                               u8* dst, size_t dstep, int top, int bottom, int left, int right) {
    u8* inner = dst + (size_t)top * dstep + left;
    if (inner != src)
        for (int y = 0; y < h; y++) memmove(inner + (size_t)y * dstep, src + (size_t)y * sstep, w);
    for (int y = 0; y < h; y++) {
        u8* r = inner + (size_t)y * dstep;
        for (int k = 1; k <= left; k++) r[-k] = r[reflect101(-k, w)];
        for (int k = 0; k < right; k++) r[w + k] = r[reflect101(w + k, w)];
    }
    int W = w + left + right;
    for (int k = 1; k <= top; k++)
        memcpy(dst + (size_t)(top - k) * dstep, dst + (size_t)(top + reflect101(-k, h)) * dstep, W);
    for (int k = 0; k < bottom; k++)
        memcpy(dst + (size_t)(top + h + k) * dstep, dst + (size_t)(top + reflect101(h + k, h)) * dstep, W);
}

// ---------------------------------------------------------------------------------------------
// cv::GaussianBlur(u8, Size(7,7), 2, 2, BORDER_REFLECT_101)  (reference call ORBextractor.cc:1086).
// OpenCV's bit-exact 8-bit path: Q8 kernel {18,34,48,56,48,34,18} (sum 256) horizontally into
// 16 bit, the same kernel vertically, dst = (V + 32768) >> 16.  In place allowed (src == dst).
// ---------------------------------------------------------------------------------------------
static const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};
static inline void gaussian_blur7_s2(const u8* src, int w, int h, size_t sstep, u8* dst, size_t dstep) {
    std::vector<uint16_t> hp((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const u8* s = src + (size_t)y * sstep;
        uint16_t* o = &hp[(size_t)y * w];
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x + 3 < w) {
                for (int k = 0; k < 7; k++) acc += kGauss7[k] * s[x + k - 3];
            } else {
                for (int k = 0; k < 7; k++) acc += kGauss7[k] * s[reflect101(x + k - 3, w)];
            }
            o[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* r[7];
        for (int k = 0; k < 7; k++) r[k] = &hp[(size_t)reflect101(y + k - 3, h) * w];
        u8* d = dst + (size_t)y * dstep;
        for (int x = 0; x < w; x++) {
            unsigned acc = 0;
            for (int k = 0; k < 7; k++) acc += (unsigned)kGauss7[k] * r[k][x];
            d[x] = (u8)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cv::fastAtan2(y, x) in degrees, float32 polynomial, no FMA contraction (volatile-free: this
// file must be compiled with -ffp-contract=off).
// ---------------------------------------------------------------------------------------------
static inline float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16 (ORBextractor.cc:809,814).
// Bresenham circle r=3 starting at (0,3) going through (3,0), (0,-3), (-3,0).
// score(p) = largest t for which p is still a FAST-9 corner
//          = max over the 16 arcs of 9 of max(min_arc d, -max_arc d) - 1,   d_k = I(p) - I(circle_k)
// corner iff score >= threshold.  NMS: keep iff score > score of all 8 neighbours (non-corners
// count 0).  Only rows 3..h-4 / cols 3..w-4 are examined.  Output in raster order.
// ---------------------------------------------------------------------------------------------
static const int kCircleDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kCircleDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

struct FastKp { int x, y, score; };

// full score of one pixel given its 16 differences; returns 0 when it is not a corner for thr
static inline int fast_score_from_d(const int* d16, int thr) {
    int d[25];
    for (int k = 0; k < 25; k++) d[k] = d16[k & 15];
    int best = thr;               // running "a0"
    for (int k = 0; k < 16; k += 2) {
        int a = std::min(d[k + 1], std::min(d[k + 2], d[k + 3]));
        if (a <= best) continue;
        a = std::min(a, std::min(d[k + 4], std::min(d[k + 5], std::min(d[k + 6], std::min(d[k + 7], d[k + 8])))));
        best = std::max(best, std::min(a, d[k]));
        best = std::max(best, std::min(a, d[k + 9]));
    }
    int nb = -best;               // running "b0"
    for (int k = 0; k < 16; k += 2) {
        int b = std::max(d[k + 1], std::max(d[k + 2], d[k + 3]));
        if (b >= nb) continue;
        b = std::max(b, std::max(d[k + 4], std::max(d[k + 5], std::max(d[k + 6], std::max(d[k + 7], d[k + 8])))));
        nb = std::min(nb, std::max(b, d[k]));
        nb = std::min(nb, std::max(b, d[k + 9]));
    }
    return -nb - 1;
}

static inline void fast9_16_nms(const u8* img, int w, int h, size_t step, int thr, std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = kCircleDy[k] * (int)step + kCircleDx[k];
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++) {
        const u8* row = img + (size_t)y * step;
        for (int x = 3; x < w - 3; x++) {
            const u8* p = row + x;
            int v = p[0];
            int lo = v - thr, hi = v + thr;
            // quick reject: a 9-arc always contains one of each opposite pair (k, k+8)
            int a0 = p[off[0]], a8 = p[off[8]];
            if (!((a0 > hi) | (a8 > hi) | (a0 < lo) | (a8 < lo))) continue;
            int a4 = p[off[4]], a12 = p[off[12]];
            if (!((a4 > hi) | (a12 > hi) | (a4 < lo) | (a12 < lo))) continue;
            unsigned br = 0, dk = 0;
            int d[16];
            for (int k = 0; k < 16; k++) {
                int q = p[off[k]];
                d[k] = v - q;
                br |= (unsigned)(q > hi) << k;
                dk |= (unsigned)(q < lo) << k;
            }
            unsigned m = br | (br << 16), n = dk | (dk << 16);
            // 9 contiguous set bits
            unsigned r = m & (m >> 1); r &= r >> 2; r &= r >> 4; r &= m >> 8;
            unsigned s = n & (n >> 1); s &= s >> 2; s &= s >> 4; s &= n >> 8;
            if (!((r | s) & 0xffffu)) continue;
            sc[(size_t)y * w + x] = fast_score_from_d(d, thr);
        }
    }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = sc[(size_t)y * w + x];
            if (!s) continue;
            const int* c = &sc[(size_t)y * w + x];
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                s > c[w - 1] && s > c[w] && s > c[w + 1])
                out.push_back(FastKp{x, y, s});
        }
}

}  // namespace cvprim
