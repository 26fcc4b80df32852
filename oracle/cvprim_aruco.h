// oracle/cvprim_aruco.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the OpenCV primitives the reference's ArUco detector calls (OpenCV is an
// un-vendored dependency, see cvprim.h).  Pinned against python cv2 4.13 golden vectors
// (tests/golden/aruco_prims.npz, made by tests/golden/make_golden.py).  Call sites in the reference
// (Thirdparty/aruco/aruco/, anchor lines of the obfuscated sources as listed in SURVEY.md section 8a):
//   adaptiveThreshold        markerdetector_impl.cpp:2984
//   findContours             markerdetector_impl.cpp:3108      (RETR_LIST, CHAIN_APPROX_NONE)
//   approxPolyDP             markerdetector_impl.cpp:3253
//   isContourConvex          markerdetector_impl.cpp:3292
//   resize (pyramid, 1/2)    markerdetector_impl.cpp:1386-1466
//   getPerspectiveTransform  markerdetector_impl.cpp:11079
//   warpPerspective          markerdetector_impl.cpp:11092
//   threshold(OTSU)          dictionary_based.cpp:1127
//   solve(DECOMP_SVD)        markerdetector_impl.cpp:11669,11841,12049
#pragma once
#include "cvprim.h"
#include <climits>

namespace cvprim {

struct Pt { int x, y; };

// ---------------------------------------------------------------------------------------------
// cv::adaptiveThreshold(src, dst, 255, ADAPTIVE_THRESH_MEAN_C, THRESH_BINARY_INV, bs, C)
// mean = rint(S * (1/bs^2)) over the bs x bs box with BORDER_REPLICATE; dst = src - mean <= -ceil(C) ? 255 : 0
// ---------------------------------------------------------------------------------------------
static inline void adaptive_threshold_mean_inv(const u8* src, int w, int h, size_t sstep, u8* dst, size_t dstep, int bs, int C) {
    const int r = bs / 2;
    const double scale = 1.0 / ((double)bs * bs);
    std::vector<int> colsum(w + 2 * r);
    auto px = [&](int y, int x) { y = std::min(std::max(y, 0), h - 1); x = std::min(std::max(x, 0), w - 1); return (int)src[(size_t)y * sstep + x]; };
    std::vector<int> rowbuf((size_t)(h + 2 * r) * 0);
    // vertical running sums per (replicated) column
    for (int xx = 0; xx < w + 2 * r; xx++) {
        int s = 0;
        for (int k = -r; k <= r; k++) s += px(k, xx - r);
        colsum[xx] = s;
    }
    for (int y = 0; y < h; y++) {
        if (y > 0)
            for (int xx = 0; xx < w + 2 * r; xx++) colsum[xx] += px(y + r, xx - r) - px(y - r - 1, xx - r);
        int s = 0;
        for (int k = 0; k < bs; k++) s += colsum[k];
        for (int x = 0; x < w; x++) {
            if (x > 0) s += colsum[x + bs - 1] - colsum[x - 1];
            int mean = round_half_even(s * scale);
            if (mean > 255) mean = 255;
            dst[(size_t)y * dstep + x] = ((int)src[(size_t)y * sstep + x] - mean <= -C) ? 255 : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cv::findContours(img, RETR_LIST, CHAIN_APPROX_NONE): Suzuki-Abe border following on a zero-padded binary
// copy.  Outer borders start where 0 -> unvisited 1, hole borders where (positive) foreground -> 0.  Traced
// pixels are re-labelled (right-edge pixels negative) so that no border is started twice.  The C++ API
// returns the contours in reverse order of discovery.
// ---------------------------------------------------------------------------------------------
static inline void trace_border(signed char* i0, int step, Pt pt, bool is_hole, int nbd, std::vector<Pt>& out) {
    const int deltas[16] = {1, -step + 1, -step, -step - 1, -1, step - 1, step, step + 1,
                            1, -step + 1, -step, -step - 1, -1, step - 1, step, step + 1};
    static const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    signed char *i1, *i3, *i4 = 0;
    int s, s_end;
    s_end = s = is_hole ? 0 : 4;
    do {
        s = (s - 1) & 7;
        i1 = i0 + deltas[s];
    } while (*i1 == 0 && s != s_end);
    if (s == s_end) {                 // isolated pixel
        *i0 = (signed char)(nbd | -128);
        out.push_back(pt);
        return;
    }
    i3 = i0;
    for (;;) {
        s_end = s;
        for (;;) {
            i4 = i3 + deltas[++s];
            if (*i4 != 0) break;
        }
        s &= 7;
        if ((unsigned)(s - 1) < (unsigned)s_end) *i3 = (signed char)(nbd | -128);   // the pixel has a 0 to its right on this border
        else if (*i3 == 1) *i3 = (signed char)nbd;
        out.push_back(pt);
        pt.x += dx[s]; pt.y += dy[s];
        if (i4 == i0 && i3 == i1) break;
        i3 = i4;
        s = (s + 4) & 7;
    }
}

static inline void find_contours_list_none(const u8* img, int w, int h, size_t step, std::vector<std::vector<Pt> >& contours) {
    const int W = w + 2, H = h + 2;
    std::vector<signed char> buf((size_t)W * H, 0);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) buf[(size_t)(y + 1) * W + x + 1] = img[(size_t)y * step + x] ? 1 : 0;
    std::vector<std::vector<Pt> > found;
    int nbd = 2;
    for (int y = 1; y < H - 1; y++) {
        signed char* row = &buf[(size_t)y * W];
        int prev = 0;
        for (int x = 1; x < W - 1; x++) {
            int p = row[x];
            if (p == prev) continue;
            bool is_hole = false, start = true;
            if (!(prev == 0 && p == 1)) {
                if (p != 0 || prev < 1) start = false;
                else is_hole = true;
            }
            if (start) {
                found.push_back(std::vector<Pt>());
                Pt o; o.x = x - (is_hole ? 1 : 0) - 1; o.y = y - 1;          // back to unpadded coordinates
                trace_border(row + x - (is_hole ? 1 : 0), W, o, is_hole, nbd, found.back());
                nbd = (nbd + 1) & 127;
                if (nbd == 0) nbd = 3;
                p = row[x];
            }
            prev = p;
        }
    }
    contours.assign(found.rbegin(), found.rend());
}

// ---------------------------------------------------------------------------------------------
// cv::approxPolyDP(curve, eps, closed = true) on integer points (Douglas-Peucker with OpenCV's start-point
// search, explicit slice stack and final collinearity clean-up).
// Canonical = cv2 4.13: the split criterion is the squared distance to the chord SEGMENT (a point projecting
// beyond an end is measured to that end).  OpenCV 3.4 -- what the reference's author linked -- measured the
// distance to the infinite chord LINE; `classic_line_distance` reproduces that (differs on ~1% of marker-frame
// contours; reported, not used for parity).
// ---------------------------------------------------------------------------------------------
static inline void approx_poly_dp_closed(const std::vector<Pt>& src, double eps, std::vector<Pt>& dst, bool classic_line_distance = false) {
    dst.clear();
    const int count = (int)src.size();
    if (count == 0) return;
    struct Slice { int start, end; };
    std::vector<Slice> stack;
    std::vector<Pt> out;
    eps *= eps;
    Slice slice = {0, 0}, right = {0, 0};
    Pt start_pt = {-1000000, -1000000}, end_pt = {0, 0}, pt = {0, 0};
    int pos = 0;
    bool le_eps = false;
    auto read = [&](Pt& p, int& ps) { p = src[ps]; if (++ps >= count) ps = 0; };
    // 1. approximately the two farthest points: three "go to the farthest point" hops
    right.start = 0;
    for (int i = 0; i < 3; i++) {
        double max_dist = 0;
        pos = (pos + right.start) % count;
        read(start_pt, pos);
        for (int j = 1; j < count; j++) {
            read(pt, pos);
            double dx = pt.x - start_pt.x, dy = pt.y - start_pt.y;
            double dist = dx * dx + dy * dy;
            if (dist > max_dist) { max_dist = dist; right.start = j; }
        }
        le_eps = max_dist <= eps;
    }
    // 2. initial slices
    if (!le_eps) {
        right.end = slice.start = pos % count;
        slice.end = right.start = (right.start + slice.start) % count;
        stack.push_back(right);
        stack.push_back(slice);
    } else out.push_back(start_pt);
    // 3. recursive subdivision
    while (!stack.empty()) {
        slice = stack.back(); stack.pop_back();
        end_pt = src[slice.end];
        pos = slice.start;
        read(start_pt, pos);
        if (pos != slice.end) {
            double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y, max_dist = 0;
            const double seg2 = dx * dx + dy * dy;
            while (pos != slice.end) {
                read(pt, pos);
                const double px = pt.x - start_pt.x, py = pt.y - start_pt.y;
                double dist;
                if (classic_line_distance) dist = std::fabs(py * dx - px * dy);
                else {
                    const double proj = px * dx + py * dy;
                    if (proj < 0) dist = px * px + py * py;
                    else if (proj > seg2) { const double ex = pt.x - end_pt.x, ey = pt.y - end_pt.y; dist = ex * ex + ey * ey; }
                    else { const double cr = py * dx - px * dy; dist = cr * cr / seg2; }
                }
                if (dist > max_dist) { max_dist = dist; right.start = (pos + count - 1) % count; }
            }
            le_eps = classic_line_distance ? (max_dist * max_dist <= eps * seg2) : (max_dist <= eps);
        } else {
            le_eps = true;
            start_pt = src[slice.start];
        }
        if (le_eps) out.push_back(start_pt);
        else {
            right.end = slice.end;
            slice.end = right.start;
            stack.push_back(right);
            stack.push_back(slice);
        }
    }
    // 4. clean-up: drop vertices lying (almost) on the line through their neighbours
    int cnt = (int)out.size(), new_count = cnt;
    auto readd = [&](Pt& p, int& ps) { p = out[ps]; if (++ps >= cnt) ps = 0; };
    pos = cnt - 1;
    readd(start_pt, pos);
    int wpos = pos;
    readd(pt, pos);
    for (int i = 0; i < cnt && new_count > 2; i++) {
        readd(end_pt, pos);
        double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
        double dist = std::fabs((pt.x - start_pt.x) * dy - (pt.y - start_pt.y) * dx);
        double sip = (double)(pt.x - start_pt.x) * (end_pt.x - pt.x) + (double)(pt.y - start_pt.y) * (end_pt.y - pt.y);
        if (dist * dist <= 0.5 * eps * (dx * dx + dy * dy) && dx != 0 && dy != 0 && sip >= 0) {
            new_count--;
            out[wpos] = start_pt = end_pt;
            if (++wpos >= cnt) wpos = 0;
            readd(pt, pos);
            i++;
            continue;
        }
        out[wpos] = start_pt = pt;
        if (++wpos >= cnt) wpos = 0;
        pt = end_pt;
    }
    dst.assign(out.begin(), out.begin() + new_count);
}

// cv::isContourConvex on integer points: all turns have the same sign, no collinear triple
static inline bool is_contour_convex(const std::vector<Pt>& p) {
    const int n = (int)p.size();
    if (n < 3) return false;
    Pt prev = p[(n - 2 + n) % n], cur = p[n - 1];
    int dx0 = cur.x - prev.x, dy0 = cur.y - prev.y, orientation = 0;
    for (int i = 0; i < n; i++) {
        prev = cur; cur = p[i];
        const int dx = cur.x - prev.x, dy = cur.y - prev.y;
        const int dxdy0 = dx * dy0, dydx0 = dy * dx0;
        orientation |= (dydx0 > dxdy0) ? 1 : ((dydx0 < dxdy0) ? 2 : 3);
        if (orientation == 3) return false;
        dx0 = dx; dy0 = dy;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// cv::resize(src, dst, Size(cols/2, rows/2)) with the default INTER_LINEAR: OpenCV switches to the 2x2 area
// mean when both scale factors are exactly 2; otherwise the generic fixed-point bilinear of cvprim.h.
// ---------------------------------------------------------------------------------------------
static inline void resize_half(const u8* src, int sw, int sh, size_t sstep, u8* dst, int dw, int dh, size_t dstep) {
    if (dw * 2 == sw && dh * 2 == sh) {
        for (int y = 0; y < dh; y++) {
            const u8 *a = src + (size_t)(2 * y) * sstep, *b = a + sstep;
            for (int x = 0; x < dw; x++) dst[(size_t)y * dstep + x] = (u8)((a[2 * x] + a[2 * x + 1] + b[2 * x] + b[2 * x + 1] + 2) >> 2);
        }
    } else resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}

// ---------------------------------------------------------------------------------------------
// cv::getPerspectiveTransform(src[4], dst[4]) (float points): 8x8 system solved by OpenCV's LU with partial
// pivoting in double (cv2 4.x default; OpenCV 3.4 used SVD: last-bit differences, SURVEY A-8).
// ---------------------------------------------------------------------------------------------
static inline bool lu_solve(double* A, int m, double* b) {
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::fabs(A[j * m + i]) > std::fabs(A[k * m + i])) k = j;
        if (std::fabs(A[k * m + i]) < DBL_EPSILON * 100) return false;
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i * m + j], A[k * m + j]);
            std::swap(b[i], b[k]);
        }
        const double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            const double alpha = A[j * m + i] * d;
            for (int c = i + 1; c < m; c++) A[j * m + c] += alpha * A[i * m + c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < m; k++) s -= A[i * m + k] * b[k];
        b[i] = s / A[i * m + i];
    }
    return true;
}
static inline bool get_perspective_transform(const float src[8], const float dst[8], double M[9]) {
    double A[64], b[8];
    for (int i = 0; i < 4; i++) {
        const float sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        double* r0 = A + i * 8; double* r1 = A + (i + 4) * 8;
        r0[0] = r1[3] = sx; r0[1] = r1[4] = sy; r0[2] = r1[5] = 1;
        r0[3] = r0[4] = r0[5] = r1[0] = r1[1] = r1[2] = 0;
        r0[6] = (float)(-sx * dx); r0[7] = (float)(-sy * dx);      // OpenCV forms these products in float (Point2f operands)
        r1[6] = (float)(-sx * dy); r1[7] = (float)(-sy * dy);
        b[i] = dx; b[i + 4] = dy;
    }
    if (!lu_solve(A, 8, b)) { for (int i = 0; i < 9; i++) M[i] = 0; return false; }
    for (int i = 0; i < 8; i++) M[i] = b[i];
    M[8] = 1.;
    return true;
}

// cv::invert of a 3x3 double matrix (closed form through the determinant)
static inline bool invert3x3(const double* S, double* t) {
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d == 0.) { for (int i = 0; i < 9; i++) t[i] = 0; return false; }
    d = 1. / d;
    t[0] = (S[4] * S[8] - S[5] * S[7]) * d;
    t[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    t[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    t[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    t[6] = (S[3] * S[7] - S[4] * S[6]) * d;
    t[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
    return true;
}

// ---------------------------------------------------------------------------------------------
// cv::warpPerspective(src, dst, M, Size(dw,dh), INTER_LINEAR, BORDER_CONSTANT 0), u8.
// Coordinates: 5 fractional bits, rounded half-even from double; weights 32768*(1-a)(1-b) etc. (exact for
// 5-bit fractions); dst = (sum + 16384) >> 15; taps outside the source read 0.  Valid for dw <= 64
// (OpenCV evaluates X0 = M0*x + M1*y + M2 per block origin x; the marker patches are a single block wide).
// ---------------------------------------------------------------------------------------------
static inline void warp_perspective_linear(const u8* src, int sw, int sh, size_t sstep, u8* dst, int dw, int dh, size_t dstep, const double M[9]) {
    double Mi[9];
    invert3x3(M, Mi);
    for (int y = 0; y < dh; y++) {
        const double X0 = Mi[0] * 0 + Mi[1] * y + Mi[2];
        const double Y0 = Mi[3] * 0 + Mi[4] * y + Mi[5];
        const double W0 = Mi[6] * 0 + Mi[7] * y + Mi[8];
        for (int x = 0; x < dw; x++) {
            double W = W0 + Mi[6] * x;
            W = W ? 32. / W : 0;
            const double fX = std::max((double)INT_MIN, std::min((double)INT_MAX, (X0 + Mi[0] * x) * W));
            const double fY = std::max((double)INT_MIN, std::min((double)INT_MAX, (Y0 + Mi[3] * x) * W));
            const int X = round_half_even(fX), Y = round_half_even(fY);
            // the integer part is saturated to short like OpenCV's map
            int sx = X >> 5, sy = Y >> 5;
            sx = std::max(-32768, std::min(32767, sx)); sy = std::max(-32768, std::min(32767, sy));
            const int ax = X & 31, ay = Y & 31;
            const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
            auto at = [&](int yy, int xx) -> int { return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstep + xx] : 0; };
            const int v = at(sy, sx) * w00 + at(sy, sx + 1) * w01 + at(sy + 1, sx) * w10 + at(sy + 1, sx + 1) * w11;
            dst[(size_t)y * dstep + x] = (u8)((v + (1 << 14)) >> 15);
        }
    }
}

// cv::threshold(THRESH_BINARY | THRESH_OTSU): returns the Otsu level
static inline int otsu_level(const u8* img, int w, int h, size_t step) {
    int hist[256] = {0};
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) hist[img[(size_t)y * step + x]]++;
    double mu = 0, scale = 1. / ((double)w * h);
    for (int i = 0; i < 256; i++) mu += i * (double)hist[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
    for (int i = 0; i < 256; i++) {
        const double p_i = hist[i] * scale;
        mu1 *= q1;
        q1 += p_i;
        const double q2 = 1. - q1;
        if (std::min(q1, q2) < FLT_EPSILON || std::max(q1, q2) > 1. - FLT_EPSILON) continue;
        mu1 = (mu1 + i * p_i) / q1;
        const double mu2 = (mu - q1 * mu1) / q2;
        const double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return (int)max_val;
}

// ---------------------------------------------------------------------------------------------
// cv::solve(A (m x n float), b (m x 1 float), x, DECOMP_SVD): OpenCV's built-in one-sided Jacobi SVD on A^T
// (float data, double accumulators) followed by SVBkSb back-substitution.  n <= 2 here.
// (A cv2 build with LAPACK routes m >= 25 to sgesdd: build-dependent at the 1e-4 px level, SURVEY A-12; the
// canonical definition is this built-in path.)
// ---------------------------------------------------------------------------------------------
static inline void solve_svd_f32(const float* A, const float* b, int m, int n, float* x) {
    std::vector<float> At((size_t)n * m), Vt((size_t)n * n, 0.f);
    std::vector<double> W(n);
    for (int i = 0; i < n; i++) for (int k = 0; k < m; k++) At[(size_t)i * m + k] = A[(size_t)k * n + i];
    const float eps = FLT_EPSILON * 2;
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { float t = At[(size_t)i * m + k]; sd += (double)t * t; }
        W[i] = sd;
        Vt[(size_t)i * n + i] = 1.f;
    }
    const int max_iter = std::max(m, 30);
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                float *Ai = &At[(size_t)i * m], *Aj = &At[(size_t)j * m];
                double a = W[i], p = 0, bb = W[j];
                for (int k = 0; k < m; k++) p += (double)Ai[k] * Aj[k];
                if (std::fabs(p) <= eps * std::sqrt((double)a * bb)) continue;
                p *= 2;
                const double beta = a - bb, gamma = hypot((double)p, beta);
                float c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = (float)std::sqrt(delta / gamma);
                    c = (float)(p / (gamma * s * 2));
                } else {
                    c = (float)std::sqrt((gamma + beta) / (gamma * 2));
                    s = (float)(p / (gamma * c * 2));
                }
                a = bb = 0;
                for (int k = 0; k < m; k++) {
                    const float t0 = c * Ai[k] + s * Aj[k];
                    const float t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += (double)t0 * t0; bb += (double)t1 * t1;
                }
                W[i] = a; W[j] = bb;
                changed = true;
                float *Vi = &Vt[(size_t)i * n], *Vj = &Vt[(size_t)j * n];
                for (int k = 0; k < n; k++) {
                    const float t0 = c * Vi[k] + s * Vj[k];
                    const float t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { float t = At[(size_t)i * m + k]; sd += (double)t * t; }
        W[i] = std::sqrt(sd);
    }
    for (int i = 0; i < n - 1; i++) {
        int j = i;
        for (int k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            std::swap(W[i], W[j]);
            for (int k = 0; k < m; k++) std::swap(At[(size_t)i * m + k], At[(size_t)j * m + k]);
            for (int k = 0; k < n; k++) std::swap(Vt[(size_t)i * n + k], Vt[(size_t)j * n + k]);
        }
    }
    std::vector<float> w(n);
    for (int i = 0; i < n; i++) {
        w[i] = (float)W[i];
        const double sd = W[i];
        const float s = (float)(sd > FLT_MIN ? 1 / sd : 0.);        // left singular vectors: rows of At scaled by 1/w
        for (int k = 0; k < m; k++) At[(size_t)i * m + k] *= s;
    }
    // SVBkSb: x = V * diag(1/w) * U^T * b   (u = rows of At, v = rows of Vt)
    for (int i = 0; i < n; i++) x[i] = 0.f;
    double threshold = 0;
    for (int i = 0; i < n; i++) threshold += w[i];
    threshold *= eps;
    for (int i = 0; i < n; i++) {
        double wi = w[i];
        if (std::fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double s = 0;
        for (int j = 0; j < m; j++) s += At[(size_t)i * m + j] * b[j];      // float product, double sum
        s *= wi;
        for (int j = 0; j < n; j++) x[j] = (float)(x[j] + s * Vt[(size_t)i * n + j]);
    }
}

}  // namespace cvprim
