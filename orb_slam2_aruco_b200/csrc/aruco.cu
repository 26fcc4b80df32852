// aruco.cu -- B200-native ArUco marker detector (sm_100a).
//
// Behavioural contract: aruco::MarkerDetector::detect on the path the reference takes (src/Frame.cc:129-142:
// DM_NORMAL => adaptive threshold, CORNER_LINES refinement, defaults of Thirdparty/aruco/aruco/markerdetector.h:162-200),
// i.e. the de-obfuscated Thirdparty/aruco/aruco/markerdetector_impl.cpp + dictionary_based.cpp (anchor lines below
// as in SURVEY.md section 8a).  Marker ids, candidate corners, warped patches and contour points are bit-identical to
// the CPU oracle (oracle/aruco_oracle.cpp); refined corners agree within 1e-4 px (float SVD with double accumulators
// whose summation order differs between a serial loop and a warp reduction).
//
// Pipeline for a batch of n frames (per-tile / per-border / per-candidate kernels, no serial raster scan):
//   k_athresh2    adaptiveThreshold MEAN_C BINARY_INV (markerdetector_impl.cpp:2984): packed 16-bit running window sums in registers,
//                 fused with the 8-neighbour foreground mask of every pixel (the binary image itself is never stored); k_athresh is
//                 the round-1 shared-memory form
//   k_halfpyr     image pyramid by exact 1/2 (2x2 mean; odd sizes: fixed-point bilinear) (1300-1466)
//   k_probe_a     cv::findContours RETR_LIST/CHAIN_APPROX_NONE (3108) without a raster scan: every 0->1 / 1->0 transition
//                 pixel is listed (byte-parallel tests on mask words)
//   k_probe_b1    every transition looks BACK along its border (Suzuki's successor rule, inverted, on the masks) to the previous
//                 transition and survives only if that one comes later in raster order
//   k_probe_b     persistent lanes walk the survivors FORWARD until they are back at themselves (=> the border's raster-first
//                 transition, i.e. Suzuki's start; the border is recorded with its length) or meet a transition the raster scan
//                 would have seen earlier (=> abort)
//   k_emit        borders longer than 70 points are followed once more and written as point lists
//   (batches of up to 32 frames: k_seg / k_link / k_ring / k_emit2 instead of k_probe_b / k_emit - the ring form, whose chains of
//    dependent loads end at the next surviving transition)
//   k_quads       warp per border: cv::approxPolyDP(eps = 0.05*len, closed) + isContourConvex (3253-3292)
//   k_prefilter   CTA per frame: candidate order = reverse discovery order, corner orientation, too-near pairs,
//                 frame-border rejection (4349-5070)
//   k_decode<0>   warp per candidate: pyramid level, getPerspectiveTransform (LU, double), warpPerspective 5-bit fixed point,
//                 histogram (6482-6803)
//   k_otsu        thread per candidate: the serial double-precision Otsu sweep (dictionary_based.cpp:1127)
//   k_decode<1>   warp per candidate: threshold, cell vote, 4 rotations, dictionary lookup (dictionary_based.cpp:1062-2509)
//   k_finalize    CTA per frame: stable sort by id, duplicate removal (8159-8311), CORNER_LINES refinement with the
//                 float one-sided Jacobi SVD of cv::solve(DECOMP_SVD) (8979-12049)
#include "common.h"
#include <atomic>
#include <math.h>
#include <float.h>
#include <algorithm>
#include <string>

namespace b200 {

constexpr int kMaxCand = 256;        // convex quads per frame
constexpr int kMaxMarkers = 64;      // output capacity per frame
constexpr int kMaxPyr = 8;
constexpr int kMinContour = 70;      // int(3.5*lowResMarkerSize), markerdetector_impl.cpp:3053
constexpr int kMaxVerts = 128;
constexpr int kMaxWarp = 50;         // warped patch side: 5*(sqrt(nbits)+2) <= 50 (64-bit dictionaries)

struct ArucoGeom {
    int w, h;
    int bpitch;                 // pitch of the padded mask images ((w + 2*kMaskPad) rounded up to 16)
    long long bframe;           // bytes per frame of the padded binary image
    int win;                    // adaptive threshold window
    unsigned mean_mul;          // ceil(2^24 / win^2)
    int nlev;                   // pyramid levels incl. level 0
    int lw[kMaxPyr], lh[kMaxPyr], lpitch[kMaxPyr];
    long long loff[kMaxPyr];    // offset of level l (>=1) inside a frame's pyramid block
    long long pyr_frame;
    int max_contours;           // contour descriptors per frame
    int max_points;             // contour points per frame
    int nbits, nb, nsub, wsize; // dictionary geometry
    int ncodes;
};

struct ContourDesc { int start; int s0; int len; int key; int off; int ck; };   // start: index into the padded binary image; ck: first checkpoint of the border in the frame's checkpoint pool, or -1
constexpr int kCkStride = 256;       // a walker that follows a border records its state every kCkStride steps (k_probe_b) ...
constexpr int kCkPerBorder = 16;     // ... at most this many times: k_emit then writes the border's points with one thread per stretch
struct SegNode { int next, len, key, off; };       // a surviving transition as a node of its border's ring: successor, steps to it, raster key, first point slot (-1: not emitted)   // start: index into the padded binary image
struct Candidate { int cx[4], cy[4]; int key; int contour; };
struct Kept { float c[8]; int contour; };
struct Decoded { int id, nrot; };


// ------------------------------------------------------------------------------------------------
// A1: adaptive threshold fused with the 8-neighbour foreground mask.  mean = rint(S/bs^2) (ties impossible for odd
// bs^2) computed as ((S + bs^2/2) * M) >> 24 with M = ceil(2^24 / bs^2) (exact for S <= 255 * 225, checked by the host
// at handle creation); out = src - mean <= -7.
// One CTA -> 64x32 mask tile.  The replicate-clamped source patch is staged with aligned 32-bit loads (patch column 0 =
// image column 64*bx - 8), window sums are separable running sums (u16), the 66x34 binary block sits 3 bytes into its
// rows so that the four-pixel output groups are word aligned, and the masks of four pixels are built with byte-parallel
// integer arithmetic (values 0/1 never carry between bytes) and stored as one word.
// Mask layout in HBM: pixel (x, y) at byte (y + 1) * bpitch + x + kMaskPad; everything around the image stays zero.
// ------------------------------------------------------------------------------------------------
constexpr int kThrTW = 64, kThrTH = 32, kThrMaxR = 7;
constexpr int kThrBW = kThrTW + 2, kThrBH = kThrTH + 2;             // binary values are needed one pixel around the tile
constexpr int kThrPP = 80;                                          // patch pitch: columns 64*bx - 8 .. 64*bx + 71
constexpr int kThrHP = 68;                                          // u16 pitch of the horizontal sums
constexpr int kThrBP = 72;                                          // byte pitch of the binary block (column c at byte c + 3)
constexpr int kMaskPad = 4;

__global__ void __launch_bounds__(256)
k_athresh(const uint8_t* __restrict__ img, long long row_stride, long long frame_stride, const __grid_constant__ ArucoGeom g,
          uint8_t* __restrict__ mask) {
    __shared__ __align__(16) uint8_t patch[(kThrBH + 2 * kThrMaxR) * kThrPP];
    __shared__ __align__(16) uint16_t hs[(kThrBH + 2 * kThrMaxR) * kThrHP];
    __shared__ __align__(16) uint8_t bin[kThrBH * kThrBP];
    const int f = blockIdx.z, x0 = blockIdx.x * kThrTW - 1, y0 = blockIdx.y * kThrTH - 1;    // origin of the 66x34 binary block
    const int bs = g.win, r = bs >> 1, sh = kThrBH + 2 * r;
    const uint8_t* src = img + (long long)f * frame_stride;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int xa = blockIdx.x * kThrTW - 8;                          // image column of patch column 0
    // ---- patch rows y0 - r .. y0 + 33 + r, columns xa .. xa + 79 (BORDER_REPLICATE)
    {
        const bool words_ok = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)row_stride) & 3) == 0;
        const int w_lo = (7 - r) >> 2, w_hi = (72 + r) >> 2;        // words that hold columns 7 - r .. 72 + r
        const int nw = w_hi - w_lo + 1;
        for (int i = tid; i < sh * 32; i += 256) {                  // (row, word) with a power-of-two row length
            const int py = i >> 5, wx = w_lo + (i & 31);
            if ((i & 31) >= nw) continue;
            const int yy = min(max(y0 + py - r, 0), g.h - 1);
            const uint8_t* row = src + (long long)yy * row_stride;
            const int gx = xa + 4 * wx;
            uint32_t v;
            if (words_ok && gx >= 0 && gx + 3 < g.w) v = *reinterpret_cast<const uint32_t*>(row + gx);
            else {
                v = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) v |= (uint32_t)row[min(max(gx + k, 0), g.w - 1)] << (8 * k);
            }
            *reinterpret_cast<uint32_t*>(&patch[py * kThrPP + 4 * wx]) = v;
        }
    }
    __syncthreads();
    // ---- horizontal running sums: hs[row][c] = sum of patch[row][c + 7 - r .. c + 7 + r], c = 0..65 (6 segments of 11)
    for (int i = tid; i < sh * 6; i += 256) {
        const int py = i / 6, c0 = (i - py * 6) * 11;
        const uint8_t* p = &patch[py * kThrPP + c0 + 7 - r];
        uint16_t* o = &hs[py * kThrHP + c0];
        int s = 0;
        for (int k = 0; k < bs; k++) s += p[k];
        o[0] = (uint16_t)s;
#pragma unroll
        for (int c = 1; c < 11; c++) { s += p[c + bs - 1] - p[c - 1]; o[c] = (uint16_t)s; }
    }
    __syncthreads();
    // ---- vertical running sums, mean, threshold: bin[y][c], 3 row segments (12, 11, 11) per column
    {
        const int half = (bs * bs) >> 1;
        const unsigned M = g.mean_mul;
        for (int i = tid; i < kThrBW * 3; i += 256) {
            const int seg = i / kThrBW, c = i - seg * kThrBW;
            const int yb0 = seg == 0 ? 0 : 1 + 11 * seg, yb1 = seg == 2 ? kThrBH : 12 + 11 * seg;     // 0..12, 12..23, 23..34
            const uint16_t* hcol = &hs[yb0 * kThrHP + c];
            int s = 0;
            for (int k = 0; k < bs; k++) s += hcol[k * kThrHP];
            const int x = x0 + c;
            const bool xin = x >= 0 && x < g.w;
            for (int yb = yb0; yb < yb1; yb++) {
                const int mean = (int)(((unsigned)(s + half) * M) >> 24);
                const int y = y0 + yb;
                const int v = patch[(yb + r) * kThrPP + c + 7];
                bin[yb * kThrBP + c + 3] = (uint8_t)((xin && y >= 0 && y < g.h && v + 7 <= mean) ? 1 : 0);
                if (yb + 1 < yb1) s += hcol[(yb - yb0 + bs) * kThrHP] - hcol[(yb - yb0) * kThrHP];
            }
        }
    }
    __syncthreads();
    // ---- 8-neighbour foreground mask (bit d = neighbour in direction d: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE); 0 for background.
    // Output pixels (tx, ty) = binary block (tx + 1, ty + 1) = bytes tx + 4 of row ty + 1: group k = word k + 1.
    for (int i = tid; i < kThrTH * (kThrTW / 4); i += 256) {
        const int ty = i >> 4, k = i & 15;
        const int y = y0 + 1 + ty, x = x0 + 1 + 4 * k;
        if (y >= g.h || x >= g.w) continue;
        const uint32_t* rn = reinterpret_cast<const uint32_t*>(&bin[ty * kThrBP]) + k;          // row above: words k, k+1, k+2
        const uint32_t* rc = rn + kThrBP / 4;
        const uint32_t* rs = rc + kThrBP / 4;
        const uint32_t n0 = rn[0], n1 = rn[1], n2 = rn[2], c0 = rc[0], c1 = rc[1], c2 = rc[2], s0 = rs[0], s1 = rs[1], s2 = rs[2];
        const uint32_t W = __funnelshift_r(c0, c1, 24), E = __funnelshift_r(c1, c2, 8);
        const uint32_t NW = __funnelshift_r(n0, n1, 24), NE = __funnelshift_r(n1, n2, 8);
        const uint32_t SW = __funnelshift_r(s0, s1, 24), SE = __funnelshift_r(s1, s2, 8);
        uint32_t m = E + 2u * NE + 4u * n1 + 8u * NW + 16u * W + 32u * SW + 64u * s1 + 128u * SE;
        m &= c1 * 255u;                                              // bytes of c1 are 0/1: 0x01 -> 0xff
        *reinterpret_cast<uint32_t*>(mask + (long long)f * g.bframe + (long long)(y + 1) * g.bpitch + x + kMaskPad) = m;
    }
}

// ------------------------------------------------------------------------------------------------
// A1, register form (the default): the same threshold and the same 8-neighbour mask as k_athresh, without shared memory.
// A warp owns a band of 128 columns (lane = one aligned word = 4 columns) and walks down a strip of kAt2Rows rows:
//   * vertical window sums of the lane's four columns as running sums in two u16x2 registers (even / odd columns), one word load for the row that
//     enters and one for the row that leaves.  sm_100a has a packed 16-bit add (VIADD.16x2) but no packed subtract, so the leaving row is added as its
//     complement (~x = -x - 1): after k steps every sum is k too small, the same k in every column, and the threshold constant of the row absorbs it
//   * horizontal window sums of the four columns from the lanes one (windows <= 7) or two (<= 15) to the left and right: shuffles of the two
//     registers, then adds of pairs that are either a neighbour's register or one PRMT across two of them
//   * v + 7 <= rint(S / bs^2)  <=>  S >= bs^2 (v + 7) - (bs^2 - 1) / 2 (both sides integers; checked for every S and v, tests/test_aruco_oracle.py), with
//     all pixels biased by -128 so that both sides fit signed 16 bits: IMAD on the packed centre pixels, VIADD.16x2 of the row constant, and
//     min.u16x2((max.s16x2(S, thr) ^ S), 1) is 1 exactly where S < thr
//   * the binary words of the lanes left and right arrive with two more shuffles; three rows of them give the mask bytes as in k_athresh.
// Lanes outside the image replicate the border columns (BORDER_REPLICATE), rows likewise; the first / last H + 1 lanes of a band are halo.
// ------------------------------------------------------------------------------------------------
constexpr int kAt2Warps = 4, kAt2Rows = 64;

__device__ __forceinline__ unsigned add16x2(unsigned a, unsigned b) { unsigned d; asm("add.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned max_s16x2(unsigned a, unsigned b) { unsigned d; asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned min_u16x2(unsigned a, unsigned b) { unsigned d; asm("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned hi_lo16(unsigned lo, unsigned hi) { return __byte_perm(lo, hi, 0x5432); }       // (lo.hi16, hi.lo16)

// column pair (V(c0 + J), V(c0 + J + 2)) of the lane whose first column is c0; A[m + H] = (V0, V2), B[m + H] = (V1, V3) of lane + m
template <int J, int H>
__device__ __forceinline__ unsigned at2_pair(const unsigned (&A)[2 * H + 1], const unsigned (&B)[2 * H + 1]) {
    constexpr int m = J >= 0 ? J / 4 : -((-J + 3) / 4);
    constexpr int q = J - 4 * m;
    if constexpr (q == 0) return A[m + H];
    else if constexpr (q == 1) return B[m + H];
    else if constexpr (q == 2) return hi_lo16(A[m + H], A[m + H + 1]);
    else return hi_lo16(B[m + H], B[m + H + 1]);
}
template <int J, int JE, int H> struct At2Sum {
    static __device__ __forceinline__ unsigned run(const unsigned (&A)[2 * H + 1], const unsigned (&B)[2 * H + 1]) { return add16x2(at2_pair<J, H>(A, B), At2Sum<J + 1, JE, H>::run(A, B)); }
};
template <int JE, int H> struct At2Sum<JE, JE, H> {
    static __device__ __forceinline__ unsigned run(const unsigned (&A)[2 * H + 1], const unsigned (&B)[2 * H + 1]) { return at2_pair<JE, H>(A, B); }
};

// FASTW: every lane of the warp reads aligned words inside the image (all warps but those of the first / last band and of unaligned images)
template <int R, bool FASTW>
__device__ __forceinline__ void at2_strip(const uint8_t* __restrict__ src, int row_stride, const ArucoGeom& g, uint8_t* __restrict__ mrow,
                                          int ys, int ye, int c0, int lane, bool fast) {
    constexpr int BS = 2 * R + 1, AREA = BS * BS, H = R <= 3 ? 1 : 2;
    const int x0 = min(max(c0, 0), g.w - 1), x1 = min(max(c0 + 1, 0), g.w - 1), x2 = min(max(c0 + 2, 0), g.w - 1), x3 = min(max(c0 + 3, 0), g.w - 1);
    unsigned colmask = 0;                                             // 0x01 in every byte whose column lies in the image
#pragma unroll
    for (int b = 0; b < 4; b++) if (c0 + b >= 0 && c0 + b < g.w) colmask |= 1u << (8 * b);
    const int hm1 = g.h - 1;
    auto load = [&](int y) -> unsigned {
        const uint8_t* row = src + (unsigned)(min(max(y, 0), hm1) * row_stride);
        if (FASTW || fast) return *reinterpret_cast<const uint32_t*>(row + c0);
        return (unsigned)row[x0] | ((unsigned)row[x1] << 8) | ((unsigned)row[x2] << 16) | ((unsigned)row[x3] << 24);
    };
    // vertical sums of binary row ys - 1: rows ys - 1 - R .. ys - 1 + R, every pixel biased by -128
    unsigned VA = (unsigned)((65536 - 128 * BS) & 0xffff) * 0x10001u, VB = VA;
#pragma unroll
    for (int k = -R; k <= R; k++) {
        const unsigned w = load(ys - 1 + k);
        VA = add16x2(VA, __byte_perm(w, 0u, 0x4240)); VB = add16x2(VB, __byte_perm(w, 0u, 0x4341));
    }
    int ck = (7 * AREA - (AREA - 1) / 2 - 128 * AREA) & 0xffff;       // threshold constant of the row (16 bits), minus BS per vertical step
    unsigned al = 0, ac = 0, ar = 0, bl = 0, bc = 0, br = 0;          // binary words (left lane, own, right lane) of rows yb - 2 and yb - 1
    const bool out_lane = lane >= H + 1 && lane <= 30 - H && c0 < g.w;
    unsigned wc = load(ys - 1), wn = load(ys + R), wo = load(ys - 1 - R);     // centre row of this step; the rows that enter / leave after it
#pragma unroll 1
    for (int yb = ys - 1; yb <= ye; yb++) {
        const unsigned wc2 = load(yb + 1), wn2 = load(yb + R + 2), wo2 = load(yb + 1 - R);      // the next step's rows, in flight during this one
        // horizontal window sums
        unsigned A[2 * H + 1], B[2 * H + 1];
        A[H] = VA; B[H] = VB;
#pragma unroll
        for (int m = 1; m <= H; m++) {
            A[H - m] = __shfl_up_sync(0xffffffffu, VA, m); B[H - m] = __shfl_up_sync(0xffffffffu, VB, m);
            A[H + m] = __shfl_down_sync(0xffffffffu, VA, m); B[H + m] = __shfl_down_sync(0xffffffffu, VB, m);
        }
        const unsigned M = At2Sum<-R + 1, R, H>::run(A, B);
        const unsigned S02 = add16x2(M, at2_pair<-R, H>(A, B)), S13 = add16x2(M, at2_pair<R + 1, H>(A, B));
        // threshold
        const unsigned Ck = (unsigned)ck * 0x10001u;
        const unsigned t02 = add16x2(__byte_perm(wc, 0u, 0x4240) * (unsigned)AREA, Ck), t13 = add16x2(__byte_perm(wc, 0u, 0x4341) * (unsigned)AREA, Ck);
        const unsigned n02 = min_u16x2(max_s16x2(S02, t02) ^ S02, 0x10001u), n13 = min_u16x2(max_s16x2(S13, t13) ^ S13, 0x10001u);       // 1 = background
        const unsigned bin = (yb >= 0 && yb <= hm1) ? ((__byte_perm(n02, n13, 0x6240) ^ 0x01010101u) & colmask) : 0u;
        const unsigned cl = __shfl_up_sync(0xffffffffu, bin, 1), cr = __shfl_down_sync(0xffffffffu, bin, 1);
        // mask row yb - 1 from the binary rows yb - 2 (a), yb - 1 (b), yb (c)
        if (yb >= ys + 1 && out_lane) {
            const uint32_t W = __funnelshift_r(bl, bc, 24), E = __funnelshift_r(bc, br, 8);
            const uint32_t NW = __funnelshift_r(al, ac, 24), NE = __funnelshift_r(ac, ar, 8);
            const uint32_t SW = __funnelshift_r(cl, bin, 24), SE = __funnelshift_r(bin, cr, 8);
            uint32_t m = E + 2u * NE + 4u * ac + 8u * NW + 16u * W + 32u * SW + 64u * bin + 128u * SE;
            m &= bc * 255u;
            *reinterpret_cast<uint32_t*>(mrow + (unsigned)(yb * g.bpitch)) = m;     // row yb - 1 lives at (yb - 1 + 1) * bpitch
        }
        al = bl; ac = bc; ar = br; bl = cl; bc = bin; br = cr;
        // slide the vertical window down by one row (the last step's slide is unused)
        const unsigned nwo = ~wo;
        VA = add16x2(add16x2(VA, __byte_perm(wn, 0u, 0x4240)), __byte_perm(nwo, 0u, 0x4240) | 0xff00ff00u);
        VB = add16x2(add16x2(VB, __byte_perm(wn, 0u, 0x4341)), __byte_perm(nwo, 0u, 0x4341) | 0xff00ff00u);
        ck = (ck - BS) & 0xffff;
        wc = wc2; wn = wn2; wo = wo2;
    }
}

template <int R>
__global__ void __launch_bounds__(kAt2Warps * 32)
k_athresh2(const uint8_t* __restrict__ img, long long row_stride, long long frame_stride, const __grid_constant__ ArucoGeom g, uint8_t* __restrict__ mask, int strip_rows) {
    constexpr int H = R <= 3 ? 1 : 2, NOUT = 30 - 2 * H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.z;
    const int ys = (blockIdx.y * kAt2Warps + warp) * strip_rows;
    if (ys >= g.h) return;
    const int ye = min(ys + strip_rows, g.h);
    const int c0 = (blockIdx.x * NOUT + lane - (H + 1)) * 4;          // image column of the lane's first pixel
    const uint8_t* src = img + (long long)f * frame_stride;
    const bool fast = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)row_stride) & 3) == 0 && c0 >= 0 && c0 + 3 < g.w;
    uint8_t* mrow = mask + (long long)f * g.bframe + c0 + kMaskPad;
    if (__all_sync(0xffffffffu, fast)) at2_strip<R, true>(src, (int)row_stride, g, mrow, ys, ye, c0, lane, true);
    else at2_strip<R, false>(src, (int)row_stride, g, mrow, ys, ye, c0, lane, fast);
}

// ------------------------------------------------------------------------------------------------
// pyramid by 1/2 (markerdetector_impl.cpp:1300-1466): cv::resize(INTER_LINEAR) == 2x2 area mean when both
// factors are exactly 2, else the generic 11-bit fixed-point bilinear (coefficients computed in-kernel in double/float
// exactly like the host table of the ORB pyramid).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lin_coef(int d, int ssize, int dsize, int& ofs, int& c0, int& c1) {
    const double scale = (double)ssize / dsize;
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)s);
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; fx = 0.f; }
    ofs = s;
    c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f));
    c1 = __float2int_rn(__fmul_rn(fx, 2048.f));
}

__global__ void __launch_bounds__(256)
k_halfpyr(const uint8_t* __restrict__ src0, long long srs, long long sfs, int sw, int sh,
          uint8_t* __restrict__ dst0, int dpitch, long long dfs, int dw, int dh) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    if (x >= dw || y >= dh) return;
    const uint8_t* s = src0 + (long long)f * sfs;
    int v;
    if (dw * 2 == sw && dh * 2 == sh) {
        const uint8_t* a = s + (long long)(2 * y) * srs + 2 * x;
        v = (a[0] + a[1] + a[srs] + a[srs + 1] + 2) >> 2;
    } else {
        int ox, cx0, cx1, oy, cy0, cy1;
        lin_coef(x, sw, dw, ox, cx0, cx1);
        lin_coef(y, sh, dh, oy, cy0, cy1);
        const uint8_t* r0 = s + (long long)oy * srs;
        const uint8_t* r1 = s + (long long)min(oy + 1, sh - 1) * srs;
        const int x1 = min(ox + 1, sw - 1);
        const int h0 = r0[ox] * cx0 + r0[x1] * cx1, h1 = r1[ox] * cx0 + r1[x1] * cx1;
        v = (((cy0 * (h0 >> 4)) >> 16) + ((cy1 * (h1 >> 4)) >> 16) + 2) >> 2;
    }
    dst0[(long long)f * dfs + (long long)y * dpitch + x] = (uint8_t)v;
}

// ------------------------------------------------------------------------------------------------
// A2: contours.  Directions as in OpenCV: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE (y down).
// A border "step" is (pixel p, direction s back to the previous border pixel); its successor looks at directions
// s+1, s+2, ... (counter-clockwise) for the first foreground neighbour d, moves there, s' = d+4.  A step whose sweep
// passes the zero West (East) neighbour is where the raster scan would see a 0->1 (1->0) transition at raster
// position pos(p) (pos(p)+1).  Suzuki starts every border at its raster-first transition.
// ------------------------------------------------------------------------------------------------
struct Step { int d, k; };     // direction taken and number of directions swept before it (all zero)
__device__ __forceinline__ Step next_step(int m, int s) {
    const unsigned r = (((unsigned)m | ((unsigned)m << 8)) >> ((s + 1) & 7)) & 0xffu;
    Step st;
    st.k = __ffs(r) - 1;
    st.d = (s + 1 + st.k) & 7;
    return st;
}
// raster-scan position at which this step is seen as a transition, or INT_MAX
__device__ __forceinline__ int step_key(int pos, int s, int k) {
    const int tw = (4 - (s + 1)) & 7, te = (0 - (s + 1)) & 7;      // sweep index of W / E
    if (tw < k) return pos;
    if (te < k) return pos + 1;
    return 0x7fffffff;
}

// the step that precedes (p, s) on its border: it sits on q = p + delta[s], left q in direction d = s+4, and its own
// back direction is the first foreground direction clockwise from d-1 (d itself when q has a single neighbour)
__device__ __forceinline__ void prev_step(int mq, int d, int& sq, int& kq) {
    const unsigned r = (((unsigned)mq | ((unsigned)mq << 8)) >> d) & 0xfeu;     // bit j <-> direction d+j, j = 1..7
    const int j = r ? 31 - __clz(r) : 0;
    sq = (d + j) & 7;
    kq = 7 - j;                                   // zero directions swept between sq and d
}

__device__ __forceinline__ int dir_delta(int d, int pitch) {
    // dx+1 / dy+1 of the 8 directions packed two bits each
    const int dx = ((0x901A >> (2 * d)) & 3) - 1, dy = ((0xA901 >> (2 * d)) & 3) - 1;
    return dy * pitch + dx;
}

// (neighbour mask, back direction) -> direction taken | transition class << 3 | pixel delta << 8, for the whole CTA (ends with a barrier)
__device__ __forceinline__ void fill_step_table(int* s_step, int bpitch) {
    for (int i = threadIdx.x; i < 256 * 8; i += blockDim.x) {
        const int m = i >> 3, sb = i & 7;
        const Step st = next_step(m, sb);
        const int key = m ? step_key(0, sb, st.k) : 0x7fffffff;
        s_step[i] = (st.d & 7) | ((key == 0 ? 1 : key == 1 ? 2 : 0) << 3) | (dir_delta(st.d & 7, bpitch) << 8);
    }
    __syncthreads();
}

// Phase A (thread per aligned 4-pixel word of the mask image): list every transition pixel -- foreground with a zero West
// neighbour (0->1, possible outer-border start) or a zero East neighbour (1->0, possible hole-border start).  The tests
// run byte-parallel on the word; transitions are sparse, so the few hits are appended with a CTA-aggregated atomic.
// (A foreground pixel without any neighbour has mask 0 like the background: a one-point contour, never a candidate.)
__global__ void __launch_bounds__(256)
k_probe_a(const uint8_t* __restrict__ mask0, const __grid_constant__ ArucoGeom g, int* __restrict__ cand, int* __restrict__ ncand,
          int max_cand, int* __restrict__ err) {
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    const int P = (y + 1) * g.bpitch + x + kMaskPad;
    uint32_t m = 0;
    if (x < g.w && y < g.h) m = *reinterpret_cast<const uint32_t*>(mask0 + (long long)f * g.bframe + P);     // bytes past the width are zero
    const uint32_t nz = (((m & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m) & 0x80808080u;          // bit 7 of every non-zero byte
    const uint32_t co = nz & ~((m & 0x10101010u) << 3), ch = nz & ~((m & 0x01010101u) << 7);
    __shared__ int s_n, s_base;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int k = __popc(co) + __popc(ch);
    int slot = 0;
    if (k) slot = atomicAdd(&s_n, k);
    __syncthreads();
    if (tid == 0 && s_n) s_base = atomicAdd(ncand + f, s_n);
    __syncthreads();
    if (k) {
        int idx = s_base + slot;
        int* out = cand + (long long)f * max_cand;
        if (idx + k > max_cand) { atomicExch(err, 8); return; }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (co & (0x80u << (8 * b))) out[idx++] = P + b;
            if (ch & (0x80u << (8 * b))) out[idx++] = (P + b) | (1 << 30);
        }
    }
}

// Phase B1 (thread per transition, grid-stride inside a frame; every lane runs the same code): the probe's own step, then
// BACKWARDS along the border (Suzuki's successor rule, inverted, on the masks) to the previous transition.  If the raster
// scan sees that one earlier, this transition cannot be the border's first one (this removes every non-topmost pixel of a
// left edge after one step).  Survivors are appended to the second list with one atomic per warp.
__global__ void __launch_bounds__(256)
k_probe_b1(const uint8_t* __restrict__ mask0, const __grid_constant__ ArucoGeom g, const int* __restrict__ cand, const int* __restrict__ ncand,
           int max_cand, int* __restrict__ surv, int* __restrict__ nsurv, int* __restrict__ err, int* __restrict__ smap0 = nullptr, uint32_t* __restrict__ sbits0 = nullptr, int sbits_words = 0) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int ns = min(ncand[f], max_cand);
    const uint8_t* mask = mask0 + (long long)f * g.bframe;
    const int* list = cand + (long long)f * max_cand;
    int* out = surv + (long long)f * max_cand;
    const int limit = 4 * g.max_points;
    for (int i0 = blockIdx.x * blockDim.x; i0 < ns; i0 += gridDim.x * blockDim.x) {        // warp-uniform trip count
        const int i = i0 + threadIdx.x;
        bool keep = false;
        int e = 0;
        if (i < ns) {
            e = list[i];
            const bool hole = (e >> 30) & 1;
            const int P = e & 0x3fffffff;
            const int m0 = mask[P];
            // Suzuki's first neighbour search: clockwise from NW (outer) / SE (hole) over 7 directions = highest set bit of
            // the mask rotated so that the first direction examined sits at bit 7 (W resp. E is known to be 0)
            const int from = hole ? 7 : 3;
            const unsigned rot = (((unsigned)m0 | ((unsigned)m0 << 8)) >> (from + 1)) & 0xffu;
            if (rot != 0) {                                // else: isolated pixel, a one-point contour
                const int s0 = (from + 1 + (31 - __clz(rot))) & 7;
                const int mykey = P + (hole ? 1 : 0);
                const Step st = next_step(m0, s0);
                // own step: a hole probe whose sweep also passes West belongs to the outer probe of the same pixel
                if (!(step_key(P, s0, st.k) < mykey)) {
                    int p = P, s = s0, n = 0;
                    for (;;) {
                        const int q = p + dir_delta(s, g.bpitch), d = (s + 4) & 7;
                        int sq, kq;
                        prev_step(mask[q], d, sq, kq);
                        p = q; s = sq;
                        if (p == P && s == s0) { keep = true; break; }             // all the way round: the border's only transition
                        const int key = step_key(p, s, kq);
                        if (key < mykey) break;
                        if (key != 0x7fffffff) { keep = true; break; }
                        if (++n > limit) { atomicExch(err, 3); break; }
                    }
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(nsurv + f, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                const int idx = base + __popc(m & ((1u << lane) - 1));
                out[idx] = e;
                if (smap0) {                                   // ring form: the survivor is found by the raster key of its step
                    const int key = (e & 0x3fffffff) + ((e >> 30) & 1);
                    smap0[(long long)f * g.bframe + key] = idx + 1;
                    atomicOr(sbits0 + (long long)f * sbits_words + (key >> 5), 1u << (key & 31));
                }
            }
        }
    }
}

// Phase B2 (persistent warps over the survivors; lanes refill in batches so that the refill code runs rarely and every lane
// spends its iterations in the same loop body): FORWARDS until back home (=> it is Suzuki's start: the border is recorded
// with its length if > 70 points) or until a transition that the raster scan sees earlier shows up (=> abort).
constexpr int kWalkSteps = 4;

__global__ void __launch_bounds__(128)
k_probe_b(const uint8_t* __restrict__ mask0, const __grid_constant__ ArucoGeom g, const int* __restrict__ surv, const int* __restrict__ nsurv,
          int max_cand, int* __restrict__ nfetch, ContourDesc* __restrict__ desc, int* __restrict__ ncont, int* __restrict__ npts,
          int* __restrict__ err, int walk_steps, int* __restrict__ ckalloc, int* __restrict__ ckpool0, int ck_cap) {
    const int f = blockIdx.x, lane = threadIdx.x & 31;
    int* ckpool = ckpool0 + (long long)f * (ck_cap + ck_cap / kCkPerBorder);      // checkpoints, then one owner (border index) per block
    int ckb = -1;                                      // this walk's block of checkpoints in the frame's pool (-1: none yet, -2: pool exhausted)
    const int ns = min(nsurv[f], max_cand);
    const uint8_t* mask = mask0 + (long long)f * g.bframe;
    const int* list = surv + (long long)f * max_cand;
    const int limit = 4 * g.max_points;
    // the whole step as one table lookup: (neighbour mask, back direction) -> direction taken | transition class << 3 | pixel delta << 8
    // (class 1: the raster scan sees this step at its pixel, 2: at the next pixel; next_step + step_key + dir_delta are ~25 dependent instructions)
    __shared__ int s_step[256 * 8];
    fill_step_table(s_step, g.bpitch);
    int P = 0, s0 = 0, mykey = 0, p = 0, s = 0, n = 0;
    bool busy = false, exhausted = false;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        // refill when a quarter of the lanes is idle (or nobody is walking any more)
        if (!exhausted && (__popc(idle) >= 8 || idle == 0xffffffffu)) {
            int base = 0;
            if (lane == 0) base = atomicAdd(nfetch + f, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + __popc(idle) >= ns) exhausted = true;                // warp-uniform: the list is (about to be) empty
            if (!busy) {
                const int i = base + __popc(idle & ((1u << lane) - 1));
                if (i < ns) {
                    const int e = list[i];
                    const bool hole = (e >> 30) & 1;
                    P = e & 0x3fffffff;
                    const int m0 = mask[P];
                    const int from = hole ? 7 : 3;
                    const unsigned rot = (((unsigned)m0 | ((unsigned)m0 << 8)) >> (from + 1)) & 0xffu;
                    s0 = (from + 1 + (31 - __clz(rot))) & 7;                // rot != 0: phase B1 kept it
                    mykey = P + (hole ? 1 : 0);
                    p = P; s = s0; n = 0; busy = true; ckb = -1;
                }
            }
        }
        if (!__any_sync(0xffffffffu, busy)) { if (exhausted) break; else continue; }
        // four steps per refill check: the two ballots and the refill test are paid once per four dependent loads
#pragma unroll 1
        for (int rep = 0; rep < walk_steps; rep++)
        if (busy) {
            const int e = s_step[((int)mask[p] << 3) | s];
            const int kc = (e >> 3) & 3;
            if (n > 0 && kc && p + kc - 1 < mykey) busy = false;
            else {
                p += e >> 8;
                s = (e + 4) & 7;
                n++;
                if ((n & (kCkStride - 1)) == 0 && n <= kCkStride * kCkPerBorder && ckb != -2) {      // state after n steps, for k_emit's stretches
                    if (ckb < 0) { const int b = atomicAdd(ckalloc + f, kCkPerBorder); ckb = b + kCkPerBorder <= ck_cap ? b : -2; if (ckb >= 0) ckpool[ck_cap + b / kCkPerBorder] = -1; }
                    if (ckb >= 0) ckpool[ckb + n / kCkStride - 1] = p | (s << 27);
                }
                if (p == P && s == s0) {
                    busy = false;
                    if (n > kMinContour) {
                        const int idx = atomicAdd(ncont + f, 1);
                        if (idx >= g.max_contours) atomicExch(err, 4);
                        else {
                            const int off = atomicAdd(npts + f, n);
                            if (off + n > g.max_points) atomicExch(err, 5);
                            else { ContourDesc c; c.start = P; c.s0 = s0; c.len = n; c.key = mykey; c.off = off; c.ck = ckb >= 0 ? ckb : -1; desc[(long long)f * g.max_contours + idx] = c; if (ckb >= 0) ckpool[ck_cap + ckb / kCkPerBorder] = idx; }
                        }
                    }
                } else if (n > limit) { atomicExch(err, 3); busy = false; }
            }
        }
    }
}

constexpr int kEmitCtasA = 4;        // k_emit: CTAs per frame that take one border per thread; the CTAs behind them take the checkpointed stretches

__global__ void __launch_bounds__(128)
k_emit(const uint8_t* __restrict__ mask0, const __grid_constant__ ArucoGeom g, const ContourDesc* __restrict__ desc,
       const int* __restrict__ ncont, short2* __restrict__ pts, const int* __restrict__ ckalloc, const int* __restrict__ ckpool0, int ck_cap) {
    // Part A (blockIdx.x < kEmitCtasA), one thread per border: its first kCkStride points when the walker that found it left checkpoints, else all of
    // them.  Part B, one thread per checkpoint of the frame's pool: the stretch that starts at the state k_probe_b recorded there (the last stretch of a
    // border runs to its end).  Only walks of 256 steps and more allocate checkpoints, so part B is a few hundred threads per frame, and no thread of
    // either part follows more than kCkStride points of a border shorter than kCkStride * (kCkPerBorder + 1).
    __shared__ int s_step[256 * 8];
    fill_step_table(s_step, g.bpitch);
    const int f = blockIdx.y;
    const uint8_t* mask = mask0 + (long long)f * g.bframe;
    const int* ckpool = ckpool0 + (long long)f * (ck_cap + ck_cap / kCkPerBorder);
    int p, s, n0, n1;
    short2* out;
    if (blockIdx.x < kEmitCtasA) {
        const int nc = min(ncont[f], g.max_contours);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += kEmitCtasA * blockDim.x) {
            const ContourDesc c = desc[(long long)f * g.max_contours + i];
            p = c.start; s = c.s0; n0 = 0; n1 = (c.ck >= 0 && c.len > kCkStride) ? kCkStride : c.len;
            out = pts + (long long)f * g.max_points + c.off;
            int x = p % g.bpitch - kMaskPad, y = p / g.bpitch - 1;
            for (int n = n0; n < n1; n++) {
                out[n] = make_short2((short)x, (short)y);
                const int e = s_step[((int)mask[p] << 3) | s];
                const int d = e & 7, dx = ((0x901A >> (2 * d)) & 3) - 1, dy = ((0xA901 >> (2 * d)) & 3) - 1;
                p += e >> 8; x += dx; y += dy;
                s = (d + 4) & 7;
            }
        }
    } else {
        const int nck = min(ckalloc[f], ck_cap);                     // allocated checkpoint slots (blocks of kCkPerBorder)
        for (int t = (blockIdx.x - kEmitCtasA) * blockDim.x + threadIdx.x; t < nck; t += (gridDim.x - kEmitCtasA) * blockDim.x) {
            const int b = t / kCkPerBorder * kCkPerBorder, k = t - b;                  // checkpoint k of block b = the state after (k + 1) * kCkStride steps
            const int owner = ckpool[ck_cap + b / kCkPerBorder];                       // border the block belongs to (-1: its walker gave up)
            if (owner < 0) continue;
            const ContourDesc c = desc[(long long)f * g.max_contours + owner];
            n0 = (k + 1) * kCkStride;
            if (c.ck != b || n0 >= c.len) continue;
            n1 = (k == kCkPerBorder - 1) ? c.len : min(c.len, n0 + kCkStride);
            const int v = ckpool[b + k];
            p = v & 0x7ffffff; s = v >> 27;
            out = pts + (long long)f * g.max_points + c.off;
            int x = p % g.bpitch - kMaskPad, y = p / g.bpitch - 1;
            for (int n = n0; n < n1; n++) {
                out[n] = make_short2((short)x, (short)y);
                const int e = s_step[((int)mask[p] << 3) | s];
                const int d = e & 7, dx = ((0x901A >> (2 * d)) & 3) - 1, dy = ((0xA901 >> (2 * d)) & 3) - 1;
                p += e >> 8; x += dx; y += dy;
                s = (d + 4) & 7;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Ring form of phases B2 + emit: the default for batches of up to kRingFrames frames, where latency counts (single frame: detector 0.84 -> 0.73 ms;
// 32 frames 1.56 -> 1.46 ms); bit-exact; slower than the end-to-end walkers on large batches, see the numbers at the end of this comment.  k_probe_b lets one lane follow a border from its start all the way round, and k_emit follows it a
// second time: two chains of dependent loads as long as the longest border (~1000 steps of ~300 cycles on the 640 x 480 frames).  Here nobody
// walks further than to the next survivor of phase B1:
//   k_seg    thread per survivor: forwards until the step of another survivor (found through smap, indexed by the raster key of the step) -> the
//            border becomes a ring of nodes {successor, steps to it, key}.  Every border step is walked exactly once
//   k_ring   thread per node: hop round the ring summing lengths; meeting a smaller key means "not the raster-first transition" (most nodes stop
//            after one or two hops).  The node that comes back to itself is Suzuki's start: it books the contour and its point range exactly like
//            k_probe_b and goes round once more to hand every node its first point slot
//   k_emit2  thread per node: walks its own segment once more and writes the points; clears its smap entry for the next call.
// Measured on a B200, 256 frames of 640 x 480: k_probe_b1 0.32 + k_seg 0.79 + k_link 0.06 + k_ring 0.45 + k_emit2 0.66 ms against 0.20 + 0.88 + 0.36 ms for
// the end-to-end walkers.  Phase B1 keeps the transitions of the DESCENDING side of a border only, so the longest segment is still half of the longest
// border (the chain of dependent loads shrinks by 2, not by 100), and with segments that uneven a static split over lanes leaves 3/4 of the lane-steps idle.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void probe_start(const uint8_t* __restrict__ mask, int e, int& P, int& s0, int& mykey) {
    const bool hole = (e >> 30) & 1;
    P = e & 0x3fffffff;
    const int m0 = mask[P];
    const int from = hole ? 7 : 3;
    const unsigned rot = (((unsigned)m0 | ((unsigned)m0 << 8)) >> (from + 1)) & 0xffu;
    s0 = (from + 1 + (31 - __clz(rot))) & 7;                            // rot != 0: phase B1 kept it
    mykey = P + (hole ? 1 : 0);
}

// raster key of a survivor list entry
__device__ __forceinline__ int entry_key(int e) { return (e & 0x3fffffff) + ((e >> 30) & 1); }

// All the kernels that walk are flattened state machines: a lane owns the survivors tid, tid + T, ... of its frame, performs ONE dependent step per loop
// iteration and moves on to its next survivor in the iteration in which it finishes one.  A warp iteration costs the latencies of all the divergent paths
// some lane takes, so the loops are written so that the only load anybody WAITS for is the step's own (mask byte / node): the next item's list entry and
// start mask are fetched one item ahead, and the survivor test of a transition is a bit in a 41 KB per-frame bitmap loaded together with the mask byte;
// the index of the survivor that ends a segment is looked up afterwards by k_link, one independent load per node.
// (Measured on the way, 256 frames, k_seg + k_ring (+ k_emit2): thread-per-item loops 0.79 + 1.36 ms - a warp runs as long as its longest item;
// k_probe_b's batch refill through an atomic counter 0.63 + 0.65 + 0.40 ms - items are 3-4 steps long and a refill costs a round trip; flattened
// but with the int map lookup, list entry and start mask loaded in line 0.97 + 0.63 + 0.68 ms - four dependent loads per warp iteration.)
// node = next (or, between k_seg and k_link, the raster key that ended the segment) : 22 | steps to it : 20 | own raster key : 22
__device__ __forceinline__ unsigned long long node_pack(int next, int len, int key) { return ((unsigned long long)(unsigned)next << 42) | ((unsigned long long)(unsigned)len << 22) | (unsigned)key; }
__device__ __forceinline__ int node_next(unsigned long long v) { return (int)(v >> 42); }
__device__ __forceinline__ int node_len(unsigned long long v) { return (int)(v >> 22) & 0xfffff; }
__device__ __forceinline__ int node_key(unsigned long long v) { return (int)v & 0x3fffff; }
constexpr int kNodeBits = 22, kNodeMaxLen = (1 << 20) - 1;
constexpr int kRingFrames = 32;      // batches up to this size take the ring form (measured crossover against the end-to-end walkers: ~48 frames of 640 x 480)

__device__ __forceinline__ int start_dir(int m0, int e) {            // Suzuki's first neighbour search of the probe e on its own mask byte
    const int from = ((e >> 30) & 1) ? 7 : 3;
    const unsigned rot = (((unsigned)m0 | ((unsigned)m0 << 8)) >> (from + 1)) & 0xffu;
    return (from + 1 + (31 - __clz(rot))) & 7;
}

__global__ void __launch_bounds__(128)
k_seg(const uint8_t* __restrict__ mask0, const __grid_constant__ ArucoGeom g, const int* __restrict__ surv, const int* __restrict__ nsurv,
      int max_cand, const uint32_t* __restrict__ sbits0, int sbits_words, unsigned long long* __restrict__ nodes0, int* __restrict__ off0, int* __restrict__ err) {
    const int f = blockIdx.x, T = gridDim.y * blockDim.x;
    const int ns = min(nsurv[f], max_cand);
    const uint8_t* mask = mask0 + (long long)f * g.bframe;
    const uint32_t* sbits = sbits0 + (long long)f * sbits_words;
    const int* list = surv + (long long)f * max_cand;
    unsigned long long* nodes = nodes0 + (long long)f * max_cand;
    int* off = off0 + (long long)f * max_cand;
    int i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    int e = list[i], e_pf = i + T < ns ? list[i + T] : 0;
    int m = mask[e & 0x3fffffff], m_pf = mask[e_pf & 0x3fffffff];
    int p = e & 0x3fffffff, s = start_dir(m, e), n = 0;
    for (;;) {
        const unsigned long long bw = *reinterpret_cast<const unsigned long long*>(sbits + ((p >> 5) & ~1));      // bits of p and p + 1 (p + 1 may cross into the odd word)
        const unsigned bw2 = (p & 63) == 63 ? sbits[(p >> 5) + 1] : 0u;
        if (n > 0) m = mask[p];
        const Step st = next_step(m, s);
        const int key = step_key(p, s, st.k);
        bool hit = false;
        if (n > 0 && key != 0x7fffffff) hit = key == p ? (bw >> (p & 63)) & 1ull : ((p & 63) == 63 ? bw2 & 1u : (bw >> ((p & 63) + 1)) & 1ull);
        if (!hit) {
            p += dir_delta(st.d, g.bpitch);
            s = (st.d + 4) & 7;
            if (++n >= kNodeMaxLen) { atomicExch(err, 3); hit = true; }
        }
        if (hit) {
            nodes[i] = node_pack(key == 0x7fffffff ? entry_key(e) : key, n, entry_key(e));
            off[i] = -1;
            i += T;
            if (i >= ns) break;
            e = e_pf; m = m_pf;
            p = e & 0x3fffffff; s = start_dir(m, e); n = 0;
            e_pf = i + T < ns ? list[i + T] : 0;
            m_pf = mask[e_pf & 0x3fffffff];
        }
    }
}

// segment-ending raster key -> survivor index
__global__ void __launch_bounds__(256)
k_link(const int* __restrict__ nsurv, int max_cand, const int* __restrict__ smap0, long long bframe, unsigned long long* __restrict__ nodes0) {
    const int f = blockIdx.y;
    const int ns = min(nsurv[f], max_cand);
    unsigned long long* nodes = nodes0 + (long long)f * max_cand;
    const int* smap = smap0 + (long long)f * bframe;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
        const unsigned long long v = nodes[i];
        nodes[i] = node_pack(smap[node_next(v)] - 1, node_len(v), node_key(v));
    }
}

__global__ void __launch_bounds__(128)
k_ring(const uint8_t* __restrict__ mask0, const __grid_constant__ ArucoGeom g, const int* __restrict__ surv, const int* __restrict__ nsurv,
       int max_cand, const unsigned long long* __restrict__ nodes0, int* __restrict__ off0, ContourDesc* __restrict__ desc,
       int* __restrict__ ncont, int* __restrict__ npts, int* __restrict__ err) {
    const int f = blockIdx.x, T = gridDim.y * blockDim.x;
    const int ns = min(nsurv[f], max_cand);
    const unsigned long long* nodes = nodes0 + (long long)f * max_cand;
    int* off = off0 + (long long)f * max_cand;
    const int limit = 4 * g.max_points;
    int i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    unsigned long long me = nodes[i], me_pf = i + T < ns ? nodes[i + T] : 0ull;
    int j = node_next(me), total = node_len(me), hops = 0, o = -1;     // o >= 0: second time round, handing out point slots
    for (;;) {
        bool done = false;
        if (j != i || o >= 0) {
            const unsigned long long nj = nodes[j];
            if (o >= 0) { off[j] = o; o += node_len(nj); j = node_next(nj); done = j == i; }
            else if (node_key(nj) < node_key(me)) done = true;        // not the raster-first transition of this border
            else {
                total += node_len(nj); j = node_next(nj);
                if (++hops > ns || total > limit) { atomicExch(err, 3); done = true; }
            }
        } else {                                                      // back home: Suzuki's start
            done = true;
            if (total > kMinContour) {
                const int idx = atomicAdd(ncont + f, 1);
                if (idx >= g.max_contours) atomicExch(err, 4);
                else {
                    const int o0 = atomicAdd(npts + f, total);
                    if (o0 + total > g.max_points) atomicExch(err, 5);
                    else {
                        const int e = surv[(long long)f * max_cand + i], P = e & 0x3fffffff;
                        ContourDesc c; c.start = P; c.s0 = start_dir(mask0[(long long)f * g.bframe + P], e); c.len = total; c.key = node_key(me); c.off = o0; c.ck = -1;
                        desc[(long long)f * g.max_contours + idx] = c;
                        o = o0; done = false;                         // j == i: the first slot is the start's own
                    }
                }
            }
        }
        if (done) {
            i += T;
            if (i >= ns) break;
            me = me_pf; j = node_next(me); total = node_len(me); hops = 0; o = -1;
            me_pf = i + T < ns ? nodes[i + T] : 0ull;
        }
    }
}

__global__ void __launch_bounds__(128)
k_emit2(const uint8_t* __restrict__ mask0, const __grid_constant__ ArucoGeom g, const int* __restrict__ surv, const int* __restrict__ nsurv,
        int max_cand, int* __restrict__ smap0, uint32_t* __restrict__ sbits0, int sbits_words, const unsigned long long* __restrict__ nodes0,
        const int* __restrict__ off0, short2* __restrict__ pts) {
    const int f = blockIdx.x, T = gridDim.y * blockDim.x;
    const int ns = min(nsurv[f], max_cand);
    const uint8_t* mask = mask0 + (long long)f * g.bframe;
    int* smap = smap0 + (long long)f * g.bframe;
    uint32_t* sbits = sbits0 + (long long)f * sbits_words;
    const int* list = surv + (long long)f * max_cand;
    const unsigned long long* nodes = nodes0 + (long long)f * max_cand;
    const int* off = off0 + (long long)f * max_cand;
    int i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    // item state one ahead: list entry, first point slot, node, start mask byte
    int e_pf = list[i], o_pf = off[i], m_pf = mask[e_pf & 0x3fffffff];
    unsigned long long v_pf = nodes[i];
    int p = 0, s = 0, x = 0, y = 0, left = 0, m = 0;
    short2* out = nullptr;
    bool first = false;
    for (;;) {
        if (left == 0) {                                              // take the prefetched item, prefetch the one after
            if (i >= ns) break;
            const int e = e_pf, o = o_pf; const unsigned long long v = v_pf; m = m_pf;
            const int key = node_key(v);
            smap[key] = 0; sbits[key >> 5] = 0u;                      // leave the maps empty for the next call
            i += T;
            if (i < ns) { e_pf = list[i]; o_pf = off[i]; v_pf = nodes[i]; m_pf = mask[e_pf & 0x3fffffff]; }
            if (o < 0 || node_len(v) == 0) continue;
            p = e & 0x3fffffff; s = start_dir(m, e); left = node_len(v);
            x = p % g.bpitch - kMaskPad; y = p / g.bpitch - 1;
            out = pts + (long long)f * g.max_points + o;
            first = true;
        }
        if (!first) m = mask[p];
        first = false;
        *out++ = make_short2((short)x, (short)y);
        const Step st = next_step(m, s);
        const int dx = ((0x901A >> (2 * st.d)) & 3) - 1, dy = ((0xA901 >> (2 * st.d)) & 3) - 1;
        p += dy * g.bpitch + dx; x += dx; y += dy;
        s = (st.d + 4) & 7;
        --left;
    }
}

// ------------------------------------------------------------------------------------------------
// A2, shared-memory form (B200_CONTOURS_SHARED=1, when the frame's bit image fits): the walks above cost one dependent L2 / HBM load per border
// step, so their kernels last as long as the longest border times the memory latency.  The thresholded image is one BIT per pixel:
// 41 KB at 640 x 480, 124 KB at 1280 x 720.  One CTA per frame stages it in shared memory as OVERLAPPED words - word k of a row holds
// pixels 30k - 1 .. 30k + 30, so every pixel has a word in which it sits at bit 1 .. 30 and its 3 x 3 neighbourhood is three shifts of
// three words (rows above / at / below) - and then runs every phase against shared memory:
//   0. stage: mask bytes -> bits (non-zero test + multiply gather), transitions (foreground with a zero West / East neighbour) appended to
//      the frame's list, overlapped words by funnel shifts
//   1. backward check of every transition (phase B1 above), survivors to the second list
//   2. forward walkers over the survivors (phase B2 above), contours longer than 70 points recorded
//   3. the recorded borders are walked once more to write their points (k_emit above)
// The step logic (next_step / prev_step / step_key) is the one above; only the source of the 8-neighbour mask differs.
// ------------------------------------------------------------------------------------------------
constexpr int kCtThreads = 512;
constexpr int kCtWarps = kCtThreads / 32;

// 8-neighbour mask (bit d = neighbour in direction d) of the pixel at bit `pos` (1 .. 30) of overlapped word `a`; pw = words per row
__device__ __forceinline__ int ow_mask(const uint32_t* __restrict__ ow, int a, int pos, int pw) {
    const unsigned t = (ow[a - pw] >> (pos - 1)) & 7u, m = (ow[a] >> (pos - 1)) & 7u, b = (ow[a + pw] >> (pos - 1)) & 7u;
    return (int)(((m >> 2) & 1u) | ((__brev(t) >> 29) << 1) | ((m & 1u) << 4) | (b << 5));
}

struct OwPos { int p, a, pos; };      // p: index into the padded mask image (raster key); (a, pos): the same pixel in the overlapped bit image

__device__ __forceinline__ OwPos ow_from_p(int p, int bpitch, int pw) {
    OwPos o;
    const int row = p / bpitch, x = p - row * bpitch - kMaskPad;       // row = y + 1
    const int k = (x * 34953) >> 20;                                   // x / 30 for x < 2^15
    o.p = p; o.a = row * pw + k; o.pos = x - 30 * k + 1;
    return o;
}

__device__ __forceinline__ void ow_move(OwPos& o, int d, int bpitch, int pw) {
    const int dx = ((0x901A >> (2 * d)) & 3) - 1, dy = ((0xA901 >> (2 * d)) & 3) - 1;
    o.p += dy * bpitch + dx;
    o.a += dy * pw;
    o.pos += dx;
    if (o.pos == 0) { o.pos = 30; o.a -= 1; } else if (o.pos == 31) { o.pos = 1; o.a += 1; }
}

__global__ void __launch_bounds__(kCtThreads)
k_contours(const uint8_t* __restrict__ mask0, const __grid_constant__ ArucoGeom g, int pw, int npw,
           int* __restrict__ cand0, int* __restrict__ ncand, int max_cand, int* __restrict__ surv0, int* __restrict__ nsurv,
           ContourDesc* __restrict__ desc0, int* __restrict__ ncont, int* __restrict__ npts, short2* __restrict__ pts0, int* __restrict__ err) {
    extern __shared__ __align__(16) uint32_t ct_raw[];
    __shared__ int s_ncand, s_nsurv, s_fetch, s_ncont, s_npts;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* mask = mask0 + (long long)f * g.bframe;
    uint32_t* ow = ct_raw;                                   // [(h + 2)][pw]
    uint32_t* prow = ct_raw + (g.h + 2) * pw + warp * (npw + 4);           // per warp: one plain bit row with a zero word in front and behind
    int* cand = cand0 + (long long)f * max_cand;
    int* surv = surv0 + (long long)f * max_cand;
    ContourDesc* desc = desc0 + (long long)f * g.max_contours;
    if (tid == 0) { s_ncand = 0; s_nsurv = 0; s_fetch = 0; s_ncont = 0; s_npts = 0; }
    for (int i = tid; i < pw; i += kCtThreads) { ow[i] = 0u; ow[(g.h + 1) * pw + i] = 0u; }
    if (lane < 4) { prow[0] = 0u; prow[npw + 1 + (lane & 1)] = 0u; prow[npw + 3] = 0u; }
    __syncthreads();
    // ---- phase 0: bits, transitions, overlapped words; one row per warp and iteration
    for (int y = warp; y < g.h; y += kCtWarps) {
        const uint8_t* mrow = mask + (long long)(y + 1) * g.bpitch + kMaskPad;
        for (int j = lane; j < npw; j += 32) {
            unsigned bits = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int x = 32 * j + 4 * q;
                if (x < g.w) {                                   // bytes past the width inside the last word are zero (padding)
                    const uint32_t m = *reinterpret_cast<const uint32_t*>(mrow + x);
                    const uint32_t nz = ((((m & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m) & 0x80808080u) >> 7;
                    bits |= ((nz * 0x00204081u) >> 21 & 0xfu) << (4 * q);
                }
            }
            prow[1 + j] = bits;
        }
        __syncwarp();
        // transitions of this row (phase A above): foreground with a zero West (outer start) / East (hole start) neighbour
        int base_row = 0;
        for (int j0 = 0; j0 < npw; j0 += 32) {
            const int j = j0 + lane;
            unsigned w = 0, co = 0, ch = 0;
            if (j < npw) {
                w = prow[1 + j];
                co = w & ~((w << 1) | (prow[j] >> 31));
                ch = w & ~((w >> 1) | (prow[2 + j] << 31));
            }
            const int k = __popc(co) + __popc(ch);
            int sc = k;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
            const int tot = __shfl_sync(0xffffffffu, sc, 31);
            if (tot) {
                if (lane == 0) base_row = atomicAdd(&s_ncand, tot);
                base_row = __shfl_sync(0xffffffffu, base_row, 0);
                int idx = base_row + sc - k;
                if (idx + k > max_cand) { if (k) atomicExch(err, 8); }
                else {
                    const int P0 = (y + 1) * g.bpitch + kMaskPad + 32 * j;
                    unsigned both = co | ch;
                    while (both) {
                        const int b = __ffs(both) - 1;
                        both &= both - 1;
                        if (co & (1u << b)) cand[idx++] = P0 + b;
                        if (ch & (1u << b)) cand[idx++] = (P0 + b) | (1 << 30);
                    }
                }
            }
        }
        for (int k = lane; k < pw; k += 32) {
            const int o = 30 * k + 31, j = o >> 5;               // bit offset of pixel 30k - 1 in the row that starts with the zero word
            ow[(y + 1) * pw + k] = __funnelshift_r(prow[j], prow[j + 1], o & 31);
        }
        __syncwarp();
    }
    __syncthreads();
    const int ns = min(s_ncand, max_cand);
    const int limit = 4 * g.max_points;
    // ---- phase 1: backward check (k_probe_b1)
    for (int i0 = 0; i0 < ns; i0 += kCtThreads) {
        const int i = i0 + tid;
        bool keep = false;
        int e = 0;
        if (i < ns) {
            e = cand[i];
            const bool hole = (e >> 30) & 1;
            const int P = e & 0x3fffffff;
            OwPos o = ow_from_p(P, g.bpitch, pw);
            const int m0 = ow_mask(ow, o.a, o.pos, pw);
            const int from = hole ? 7 : 3;
            const unsigned rot = (((unsigned)m0 | ((unsigned)m0 << 8)) >> (from + 1)) & 0xffu;
            if (rot != 0) {
                const int s0 = (from + 1 + (31 - __clz(rot))) & 7;
                const int mykey = P + (hole ? 1 : 0);
                const Step st = next_step(m0, s0);
                if (!(step_key(P, s0, st.k) < mykey)) {
                    int s = s0, n = 0;
                    for (;;) {
                        const int d = (s + 4) & 7;
                        ow_move(o, s, g.bpitch, pw);
                        int sq, kq;
                        prev_step(ow_mask(ow, o.a, o.pos, pw), d, sq, kq);
                        s = sq;
                        if (o.p == P && s == s0) { keep = true; break; }
                        const int key = step_key(o.p, s, kq);
                        if (key < mykey) break;
                        if (key != 0x7fffffff) { keep = true; break; }
                        if (++n > limit) { atomicExch(err, 3); break; }
                    }
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_nsurv, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) surv[base + __popc(m & ((1u << lane) - 1))] = e;
        }
    }
    __syncthreads();
    const int nsv = min(s_nsurv, max_cand);
    // ---- phase 2: forward walkers (k_probe_b), lanes refill in batches
    {
        int P = 0, s0 = 0, mykey = 0, s = 0, n = 0;
        OwPos o; o.p = 0; o.a = pw; o.pos = 1;
        bool busy = false, exhausted = false;
        for (;;) {
            const unsigned idle = __ballot_sync(0xffffffffu, !busy);
            if (!exhausted && (__popc(idle) >= 8 || idle == 0xffffffffu)) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_fetch, __popc(idle));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + __popc(idle) >= nsv) exhausted = true;
                if (!busy) {
                    const int i = base + __popc(idle & ((1u << lane) - 1));
                    if (i < nsv) {
                        const int e = surv[i];
                        const bool hole = (e >> 30) & 1;
                        P = e & 0x3fffffff;
                        o = ow_from_p(P, g.bpitch, pw);
                        const int m0 = ow_mask(ow, o.a, o.pos, pw);
                        const int from = hole ? 7 : 3;
                        const unsigned rot = (((unsigned)m0 | ((unsigned)m0 << 8)) >> (from + 1)) & 0xffu;
                        s0 = (from + 1 + (31 - __clz(rot))) & 7;
                        mykey = P + (hole ? 1 : 0);
                        s = s0; n = 0; busy = true;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, busy)) { if (exhausted) break; else continue; }
            if (busy) {
                const Step st = next_step(ow_mask(ow, o.a, o.pos, pw), s);
                if (n > 0 && step_key(o.p, s, st.k) < mykey) busy = false;
                else {
                    ow_move(o, st.d, g.bpitch, pw);
                    s = (st.d + 4) & 7;
                    n++;
                    if (o.p == P && s == s0) {
                        busy = false;
                        if (n > kMinContour) {
                            const int idx = atomicAdd(&s_ncont, 1);
                            if (idx >= g.max_contours) atomicExch(err, 4);
                            else {
                                const int off = atomicAdd(&s_npts, n);
                                ContourDesc c; c.start = P; c.s0 = s0; c.len = n; c.key = mykey; c.off = off; c.ck = -1;
                                if (off + n > g.max_points) { atomicExch(err, 5); c.len = 0; c.off = 0; }      // the call fails; keep the slot harmless
                                desc[idx] = c;
                            }
                        }
                    } else if (n > limit) { atomicExch(err, 3); busy = false; }
                }
            }
        }
    }
    __syncthreads();
    // ---- phase 3: the points of the recorded borders (k_emit)
    const int nc = min(s_ncont, g.max_contours);
    if (tid == 0) { ncont[f] = s_ncont; npts[f] = s_npts; ncand[f] = s_ncand; nsurv[f] = s_nsurv; }
    short2* pts = pts0 + (long long)f * g.max_points;
    for (int i = tid; i < nc; i += kCtThreads) {
        const ContourDesc c = desc[i];
        if (c.len <= 0 || c.off + c.len > g.max_points) continue;
        short2* out = pts + c.off;
        OwPos o = ow_from_p(c.start, g.bpitch, pw);
        int s = c.s0;
        int x = c.start % g.bpitch - kMaskPad, y = c.start / g.bpitch - 1;
        for (int n = 0; n < c.len; n++) {
            out[n] = make_short2((short)x, (short)y);
            const Step st = next_step(ow_mask(ow, o.a, o.pos, pw), s);
            x += ((0x901A >> (2 * st.d)) & 3) - 1; y += ((0xA901 >> (2 * st.d)) & 3) - 1;
            ow_move(o, st.d, g.bpitch, pw);
            s = (st.d + 4) & 7;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// A3: approxPolyDP(closed) + isContourConvex, one warp per contour.  Semantics of cv2 4.13 (distance to the chord
// SEGMENT, squared, first maximum wins); all split decisions in double like OpenCV.
// ------------------------------------------------------------------------------------------------
struct DMax { double v; int i; };
__device__ __forceinline__ DMax warp_argmax_first(double v, int i) {
    // maximum value, smallest index among equals; lanes without data pass v < 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    DMax r; r.v = v; r.i = i; return r;
}

constexpr int kQuadWarps = 4;

__global__ void __launch_bounds__(kQuadWarps * 32)
k_quads(const __grid_constant__ ArucoGeom g, const ContourDesc* __restrict__ desc, const int* __restrict__ ncont,
        const short2* __restrict__ pts0, Candidate* __restrict__ cand, int* __restrict__ ncand, int* __restrict__ err) {
    __shared__ int s_vx[kQuadWarps][kMaxVerts], s_vy[kQuadWarps][kMaxVerts];
    __shared__ int s_stack[kQuadWarps][2 * 64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const int ncontours = min(ncont[f], g.max_contours);
    for (int ci = blockIdx.x * kQuadWarps + warp; ci < ncontours; ci += gridDim.x * kQuadWarps) {
    const ContourDesc c = desc[(long long)f * g.max_contours + ci];
    const short2* P = pts0 + (long long)f * g.max_points + c.off;
    const int count = c.len;
    double eps = (double)count * 0.05;
    eps *= eps;
    int* vx = s_vx[warp]; int* vy = s_vy[warp]; int* stack = s_stack[warp];
    int nv = 0, top = 0;
    bool overflow = false;

    // 1. three "go to the farthest point" hops
    int pos = 0, rstart = 0;
    bool le_eps = false;
    short2 sp = make_short2(0, 0);
    for (int hop = 0; hop < 3; hop++) {
        pos = (pos + rstart) % count;
        sp = P[pos];
        double best = -1.0; int bi = 0x7fffffff;
        for (int j = 1 + lane; j < count; j += 32) {
            int q = pos + j; if (q >= count) q -= count;
            const short2 pt = P[q];
            const double dx = pt.x - sp.x, dy = pt.y - sp.y;
            const double d = dx * dx + dy * dy;
            if (d > best) { best = d; bi = j; }
        }
        const DMax r = warp_argmax_first(best, bi);
        if (r.v > 0) rstart = r.i;
        le_eps = !(r.v > eps);            // max_dist (>= 0) <= eps
    }
    // 2. initial slices (start of the last hop, farthest point from it)
    if (!le_eps) {
        const int s_start = pos % count, s_end = (rstart + s_start) % count;
        if (lane == 0) {
            stack[0] = s_end; stack[1] = s_start;      // right slice  [s_end -> s_start]
            stack[2] = s_start; stack[3] = s_end;      // slice        [s_start -> s_end], processed first
        }
        top = 2;
    } else { if (lane == 0) { vx[0] = sp.x; vy[0] = sp.y; } nv = 1; }
    __syncwarp();
    // 3. subdivision
    while (top > 0) {
        top--;
        const int a = stack[2 * top], b = stack[2 * top + 1];
        __syncwarp();                                              // all lanes hold the slice before lane 0 reuses its stack slots
        const short2 start_pt = P[a], end_pt = P[b];
        int nint = b - a - 1; if (nint < 0) nint += count;         // interior points a+1 .. b-1 (cyclic)
        bool le = true;
        int split = 0;
        if (nint > 0) {
            const double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
            const double seg2 = dx * dx + dy * dy;
            double best = -1.0; int bi = 0x7fffffff;
            for (int j = lane; j < nint; j += 32) {
                int q = a + 1 + j; if (q >= count) q -= count;
                const short2 pt = P[q];
                const double px = pt.x - start_pt.x, py = pt.y - start_pt.y;
                const double proj = px * dx + py * dy;
                double d;
                if (proj < 0) d = px * px + py * py;
                else if (proj > seg2) { const double ex = pt.x - end_pt.x, ey = pt.y - end_pt.y; d = ex * ex + ey * ey; }
                else { const double cr = py * dx - px * dy; d = cr * cr / seg2; }
                if (d > best) { best = d; bi = j; }
            }
            const DMax r = warp_argmax_first(best, bi);
            // OpenCV: max_dist starts at 0 and only strictly larger distances move the split point
            if (r.v > 0) { split = a + 1 + r.i; if (split >= count) split -= count; le = r.v <= eps; }
            else le = true;
        }
        if (le) {
            if (nv < kMaxVerts) { if (lane == 0) { vx[nv] = start_pt.x; vy[nv] = start_pt.y; } nv++; }
            else overflow = true;
        } else {
            if (top + 2 > 64) { overflow = true; break; }
            if (lane == 0) { stack[2 * top] = split; stack[2 * top + 1] = b; stack[2 * top + 2] = a; stack[2 * top + 3] = split; }
            top += 2;
        }
        __syncwarp();
    }
    __syncwarp();
    if (overflow) continue;            // far more than 4 vertices: not a marker candidate
    // 4. clean-up pass (serial, a handful of vertices) and the quad / convexity test
    if (lane == 0) {
        const int cnt = nv;
        int new_count = cnt, wpos, p2 = cnt - 1;
        int sx = vx[p2], sy = vy[p2]; if (++p2 >= cnt) p2 = 0;
        wpos = p2;
        int px = vx[p2], py = vy[p2]; if (++p2 >= cnt) p2 = 0;
        for (int i = 0; i < cnt && new_count > 2; i++) {
            const int ex = vx[p2], ey = vy[p2]; if (++p2 >= cnt) p2 = 0;
            const double dx = ex - sx, dy = ey - sy;
            const double dist = fabs((px - sx) * dy - (py - sy) * dx);
            const double sip = (double)(px - sx) * (ex - px) + (double)(py - sy) * (ey - py);
            if (dist * dist <= 0.5 * eps * (dx * dx + dy * dy) && dx != 0 && dy != 0 && sip >= 0) {
                new_count--;
                vx[wpos] = sx = ex; vy[wpos] = sy = ey;
                if (++wpos >= cnt) wpos = 0;
                px = vx[p2]; py = vy[p2]; if (++p2 >= cnt) p2 = 0;
                i++;
                continue;
            }
            vx[wpos] = sx = px; vy[wpos] = sy = py;
            if (++wpos >= cnt) wpos = 0;
            px = ex; py = ey;
        }
        if (new_count == 4) {
            // isContourConvex on integer points
            int prx = vx[2], pry = vy[2], cx = vx[3], cy = vy[3];
            int dx0 = cx - prx, dy0 = cy - pry, orientation = 0;
            bool convex = true;
            for (int i = 0; i < 4; i++) {
                prx = cx; pry = cy; cx = vx[i]; cy = vy[i];
                const int dx = cx - prx, dy = cy - pry;
                const int dxdy0 = dx * dy0, dydx0 = dy * dx0;
                orientation |= (dydx0 > dxdy0) ? 1 : ((dydx0 < dxdy0) ? 2 : 3);
                if (orientation == 3) { convex = false; break; }
                dx0 = dx; dy0 = dy;
            }
            if (convex) {
                const int idx = atomicAdd(ncand + f, 1);
                if (idx < kMaxCand) {
                    Candidate cd;
                    for (int k = 0; k < 4; k++) { cd.cx[k] = vx[k]; cd.cy[k] = vy[k]; }
                    cd.key = c.key; cd.contour = ci;
                    cand[(long long)f * kMaxCand + idx] = cd;
                } else atomicExch(err, 6);
            }
        }
    }
    __syncwarp();
    }   // contours of this warp
}

// ------------------------------------------------------------------------------------------------
// prefilterCandidates (markerdetector_impl.cpp:4349-5070), one CTA per frame
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int perimeter_i(const float* c) {       // markerdetector_impl.cpp:11123
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int j = (i + 1) & 3;
        const float dx = __fsub_rn(c[2 * i], c[2 * j]), dy = __fsub_rn(c[2 * i + 1], c[2 * j + 1]);
        sum += (int)__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    }
    return sum;
}

__global__ void __launch_bounds__(256)
k_prefilter(const __grid_constant__ ArucoGeom g, const Candidate* __restrict__ cand0, const int* __restrict__ ncand,
            Kept* __restrict__ kept0, int* __restrict__ nkept) {
    __shared__ float s_c[kMaxCand][8];
    __shared__ int s_contour[kMaxCand], s_perim[kMaxCand];
    __shared__ unsigned char s_rm[kMaxCand];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(ncand[f], kMaxCand);
    const Candidate* cand = cand0 + (long long)f * kMaxCand;
    // candidate order = contour order = reverse discovery order (cv::findContours returns the newest first)
    for (int i = tid; i < n; i += blockDim.x) {
        const Candidate c = cand[i];
        int rank = 0;
        for (int j = 0; j < n; j++) rank += cand[j].key > c.key;
        float q[8];
        for (int k = 0; k < 4; k++) { q[2 * k] = (float)c.cx[k]; q[2 * k + 1] = (float)c.cy[k]; }
        // consistent corner orientation (4352)
        const double dx1 = q[2] - q[0], dy1 = q[3] - q[1], dx2 = q[4] - q[0], dy2 = q[5] - q[1];
        const double o = (dx1 * dy2) - (dy1 * dx2);
        if (o < 0.0) { float t = q[2]; q[2] = q[6]; q[6] = t; t = q[3]; q[3] = q[7]; q[7] = t; }
        for (int k = 0; k < 8; k++) s_c[rank][k] = q[k];
        s_contour[rank] = c.contour;
        s_perim[rank] = perimeter_i(q);
        s_rm[rank] = 0;
    }
    __syncthreads();
    const float too_near = (float)g.win;
    for (int pi = tid; pi < n * n; pi += blockDim.x) {
        const int i = pi / n, j = pi - i * n;
        if (j <= i) continue;
        bool all = true;
        for (int k = 0; k < 4 && all; k++) {
            const float dx = __fsub_rn(s_c[i][2 * k], s_c[j][2 * k]), dy = __fsub_rn(s_c[i][2 * k + 1], s_c[j][2 * k + 1]);
            const float d = (float)sqrt((double)dx * dx + (double)dy * dy);
            if (!(d < too_near)) all = false;
        }
        if (all) { if (s_perim[i] > s_perim[j]) s_rm[j] = 1; else s_rm[i] = 1; }     // flags are only ever set: order-free
    }
    __syncthreads();
    const int bx = (int)__fmul_rn(0.015f, (float)g.w), by = (int)__fmul_rn(0.015f, (float)g.h);
    for (int i = tid; i < n; i += blockDim.x)
        for (int k = 0; k < 4; k++) {
            const float x = s_c[i][2 * k], y = s_c[i][2 * k + 1];
            if (x < bx || y < by || x > g.w - bx || y > g.h - by) s_rm[i] = 1;
        }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        Kept* kept = kept0 + (long long)f * kMaxCand;
        for (int i = 0; i < n; i++)
            if (!s_rm[i]) {
                for (int k = 0; k < 8; k++) kept[m].c[k] = s_c[i][k];
                kept[m].contour = s_contour[i];
                m++;
            }
        nkept[f] = m;
    }
}

// ------------------------------------------------------------------------------------------------
// per-candidate decode (markerdetector_impl.cpp:6482-6803; dictionary_based.cpp:1062-2509), one CTA per candidate
// ------------------------------------------------------------------------------------------------
__device__ bool lu_solve8(double* A, double* b) {           // OpenCV hal::LU with partial pivoting, double
    const int m = 8;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++) if (fabs(A[j * m + i]) > fabs(A[k * m + i])) k = j;
        if (fabs(A[k * m + i]) < DBL_EPSILON * 100) return false;
        if (k != i) {
            for (int j = i; j < m; j++) { const double t = A[i * m + j]; A[i * m + j] = A[k * m + j]; A[k * m + j] = t; }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
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

// Otsu threshold of every warped patch (cv::threshold THRESH_OTSU, dictionary_based.cpp:1127; SURVEY A-9): the sweep is a serial
// double-precision recurrence, so it runs one thread per candidate on the histograms k_decode<0> left in global memory.
__global__ void __launch_bounds__(128)
k_otsu(const __grid_constant__ ArucoGeom g, const int* __restrict__ nkept, const uint16_t* __restrict__ g_hist, int* __restrict__ g_level) {
    const int f = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= min(nkept[f], kMaxCand)) return;
    const uint16_t* hist = g_hist + ((long long)f * kMaxCand + k) * 256;
    const int ws = g.wsize;
    int lo = 256, hi = -1;
    double mu = 0; const double scale = 1. / ((double)ws * ws);
    for (int i = 0; i < 256; i++) { const int c = hist[i]; if (c) { lo = min(lo, i); hi = i; } mu += i * (double)c; }
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0; int max_val = 0;
    // bins below the first / above the last occupied one cannot change the result (q1 = 0 resp. q2 < FLT_EPSILON => `continue`)
    for (int i = lo; i <= hi; i++) {
        const double p_i = hist[i] * scale;
        mu1 *= q1;
        q1 += p_i;
        const double q2 = 1. - q1;
        if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) continue;
        mu1 = (mu1 + i * p_i) / q1;
        const double mu2 = (mu - q1 * mu1) / q2;
        const double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    g_level[(long long)f * kMaxCand + k] = max_val;
}

constexpr int kDecodeWarps = 4;

// PHASE 0: pyramid level, homography, warped patch + its histogram -> global scratch;  (k_otsu: one THREAD per candidate runs the
// serial Otsu sweep, 32 candidates per warp instead of one lane of 32);  PHASE 1: threshold, cell votes, codes, dictionary lookup.
template <int PHASE>
__global__ void __launch_bounds__(128)
k_decode(const uint8_t* __restrict__ img0, long long row_stride, long long frame_stride, const uint8_t* __restrict__ pyr,
         const __grid_constant__ ArucoGeom g, const Kept* __restrict__ kept0, const int* __restrict__ nkept,
         const unsigned long long* __restrict__ codes, Decoded* __restrict__ dec0,
         uint8_t* __restrict__ g_patch, uint16_t* __restrict__ g_hist, int* __restrict__ g_level) {
    // one WARP per candidate (4 per CTA): the serial double-precision section (8x8 LU) of different candidates
    // overlaps on the SM's four schedulers instead of idling 127 threads each
    struct WarpState {
        double A[64], b[8], Mi[9];
        unsigned long long ids[4];
        int hist[256], nz[100], tot[100], level, lvl, found[4], ok, lo, hi;
        uint8_t patch[kMaxWarp * kMaxWarp + 4];
    };
    __shared__ WarpState s_ws[kDecodeWarps];
    WarpState& S = s_ws[threadIdx.x >> 5];
    double* s_A = S.A; double* s_b = S.b; double* s_Mi = S.Mi;
    uint8_t* s_patch = S.patch; int* s_hist = S.hist; int* s_nz = S.nz; int* s_tot = S.tot; int* s_found = S.found;
    unsigned long long* s_ids = S.ids;
    int& s_level = S.level; int& s_lvl = S.lvl; int& s_ok = S.ok;
    const int f = blockIdx.y, tid = threadIdx.x & 31;
    constexpr int kStride = 32;
    const int nk = min(nkept[f], kMaxCand);
    for (int k = blockIdx.x * kDecodeWarps + (threadIdx.x >> 5); k < nk; k += gridDim.x * kDecodeWarps) {
    __syncwarp();                                      // shared state of the previous candidate is dead
    const Kept kp = kept0[(long long)f * kMaxCand + k];
    const int ws = g.wsize;
    uint8_t* gp = g_patch + ((long long)f * kMaxCand + k) * (kMaxWarp * kMaxWarp);
    if (PHASE == 0) {
    if (tid == 0) {
        // Marker::getArea (marker.cpp:405-416) and the pyramid level (6556)
        const float* c = kp.c;
        const float v01x = __fsub_rn(c[2], c[0]), v01y = __fsub_rn(c[3], c[1]), v03x = __fsub_rn(c[6], c[0]), v03y = __fsub_rn(c[7], c[1]);
        const float a1 = fabsf(__fsub_rn(__fmul_rn(v01x, v03y), __fmul_rn(v01y, v03x)));
        const float v21x = __fsub_rn(c[2], c[4]), v21y = __fsub_rn(c[3], c[5]), v23x = __fsub_rn(c[6], c[4]), v23y = __fsub_rn(c[7], c[5]);
        const float a2 = fabsf(__fsub_rn(__fmul_rn(v21x, v23y), __fmul_rn(v21y, v23x)));
        const float area = __fdiv_rn(__fadd_rn(a2, a1), 2.f);
        const float ws2 = __fmul_rn((float)ws, (float)ws);        // std::pow(float(ws), 2.f)
        int lvl = 0;
        double p4 = 1.0;
        for (int p = 1; p < g.nlev; p++) {
            p4 *= 4.0;
            if ((double)area / p4 >= (double)ws2) lvl = p; else break;
        }
        s_lvl = lvl;
        const float scale = __fdiv_rn((float)g.lw[lvl], (float)g.w);
        float sc[8];
        for (int i = 0; i < 8; i++) sc[i] = __fmul_rn(c[i], scale);
        const float q = (float)(ws - 1);
        const float dst[8] = {0.f, 0.f, q, 0.f, q, q, 0.f, q};
        // getPerspectiveTransform: the products are formed in float (Point2f operands), the system is double
        for (int i = 0; i < 4; i++) {
            const float sx = sc[2 * i], sy = sc[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
            double* r0 = s_A + i * 8; double* r1 = s_A + (i + 4) * 8;
            r0[0] = r1[3] = sx; r0[1] = r1[4] = sy; r0[2] = r1[5] = 1;
            r0[3] = r0[4] = r0[5] = r1[0] = r1[1] = r1[2] = 0;
            r0[6] = (double)__fmul_rn(-sx, dx); r0[7] = (double)__fmul_rn(-sy, dx);
            r1[6] = (double)__fmul_rn(-sx, dy); r1[7] = (double)__fmul_rn(-sy, dy);
            s_b[i] = dx; s_b[i + 4] = dy;
        }
        double M[9];
        if (lu_solve8(s_A, s_b)) { for (int i = 0; i < 8; i++) M[i] = s_b[i]; M[8] = 1.; }
        else for (int i = 0; i < 9; i++) M[i] = 0;
        // cv::invert 3x3
        double d = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
        if (d != 0.) {
            d = 1. / d;
            s_Mi[0] = (M[4] * M[8] - M[5] * M[7]) * d; s_Mi[1] = (M[2] * M[7] - M[1] * M[8]) * d; s_Mi[2] = (M[1] * M[5] - M[2] * M[4]) * d;
            s_Mi[3] = (M[5] * M[6] - M[3] * M[8]) * d; s_Mi[4] = (M[0] * M[8] - M[2] * M[6]) * d; s_Mi[5] = (M[2] * M[3] - M[0] * M[5]) * d;
            s_Mi[6] = (M[3] * M[7] - M[4] * M[6]) * d; s_Mi[7] = (M[1] * M[6] - M[0] * M[7]) * d; s_Mi[8] = (M[0] * M[4] - M[1] * M[3]) * d;
        } else for (int i = 0; i < 9; i++) s_Mi[i] = 0;
    }
    for (int i = tid; i < 256; i += kStride) s_hist[i] = 0;
    for (int i = tid; i < 100; i += kStride) { s_nz[i] = 0; s_tot[i] = 0; }
    __syncwarp();
    // warpPerspective INTER_LINEAR, BORDER_CONSTANT 0 (SURVEY A-8)
    const int lvl = s_lvl;
    const uint8_t* src; long long spitch; const int sw = g.lw[lvl], sh = g.lh[lvl];
    if (lvl == 0) { src = img0 + (long long)f * frame_stride; spitch = row_stride; }
    else { src = pyr + (long long)f * g.pyr_frame + g.loff[lvl]; spitch = g.lpitch[lvl]; }
    for (int i = tid; i < ws * ws; i += kStride) {
        const int y = i / ws, x = i - y * ws;
        const double X0 = s_Mi[0] * 0 + s_Mi[1] * y + s_Mi[2];
        const double Y0 = s_Mi[3] * 0 + s_Mi[4] * y + s_Mi[5];
        const double W0 = s_Mi[6] * 0 + s_Mi[7] * y + s_Mi[8];
        double W = W0 + s_Mi[6] * x;
        W = W ? 32. / W : 0;
        const double fX = fmax(-2147483648.0, fmin(2147483647.0, (X0 + s_Mi[0] * x) * W));
        const double fY = fmax(-2147483648.0, fmin(2147483647.0, (Y0 + s_Mi[3] * x) * W));
        const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
        int sx = X >> 5, sy = Y >> 5;
        sx = max(-32768, min(32767, sx)); sy = max(-32768, min(32767, sy));
        const int ax = X & 31, ay = Y & 31;
        int v = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int xx = sx + (t & 1), yy = sy + (t >> 1);
            const int wgt = ((t & 1) ? ax : 32 - ax) * ((t >> 1) ? ay : 32 - ay) * 32;
            if (xx >= 0 && xx < sw && yy >= 0 && yy < sh) v += src[(long long)yy * spitch + xx] * wgt;
        }
        const int px = (v + (1 << 14)) >> 15;
        s_patch[i] = (uint8_t)px;
        atomicAdd(&s_hist[px], 1);
    }
    __syncwarp();
    uint16_t* gh = g_hist + ((long long)f * kMaxCand + k) * 256;
    for (int i = tid; i < 256; i += kStride) gh[i] = (uint16_t)s_hist[i];
    for (int i = tid; i < (ws * ws + 3) / 4; i += kStride) reinterpret_cast<uint32_t*>(gp)[i] = reinterpret_cast<const uint32_t*>(s_patch)[i];
    continue;
    }
    // ---- PHASE 1
    for (int i = tid; i < (ws * ws + 3) / 4; i += kStride) reinterpret_cast<uint32_t*>(s_patch)[i] = reinterpret_cast<const uint32_t*>(gp)[i];
    for (int i = tid; i < 100; i += kStride) { s_nz[i] = 0; s_tot[i] = 0; }
    if (tid == 0) s_level = g_level[(long long)f * kMaxCand + k];
    __syncwarp();
    const int nsub = g.nsub;
    for (int i = tid; i < ws * ws; i += kStride) {
        const int y = i / ws, x = i - y * ws;
        const int my = (int)__fdiv_rn(__fmul_rn((float)nsub, (float)y), (float)ws);
        const int mx = (int)__fdiv_rn(__fmul_rn((float)nsub, (float)x), (float)ws);
        if (s_patch[i] > s_level) atomicAdd(&s_nz[my * 10 + mx], 1);
        atomicAdd(&s_tot[my * 10 + mx], 1);
    }
    __syncwarp();
    if (tid == 0) {
        const int nb = g.nb;
        unsigned char bits[10][10];
        bool ok = true;
        for (int y = 0; y < nsub; y++) for (int x = 0; x < nsub; x++) bits[y][x] = s_nz[y * 10 + x] > s_tot[y * 10 + x] / 2;
        for (int y = 0; y < nsub && ok; y++) {
            const int inc = (y == 0 || y == nsub - 1) ? 1 : nsub - 1;
            for (int x = 0; x < nsub; x += inc) if (bits[y][x]) { ok = false; break; }
        }
        if (ok) {
            unsigned char in[8][8], tmp[8][8];
            for (int y = 0; y < nb; y++) for (int x = 0; x < nb; x++) in[y][x] = bits[y + 1][x + 1];
            for (int r = 0; r < 4; r++) {
                unsigned long long code = 0; int b = 0;
                for (int y = nb - 1; y >= 0; y--) for (int x = nb - 1; x >= 0; x--) code |= (unsigned long long)in[y][x] << b++;
                s_ids[r] = code;
                for (int i = 0; i < nb; i++) for (int j = 0; j < nb; j++) tmp[i][j] = in[nb - j - 1][i];
                for (int i = 0; i < nb; i++) for (int j = 0; j < nb; j++) in[i][j] = tmp[i][j];
            }
            if (s_ids[0] == 0) ok = false;
        }
        s_ok = ok;
        for (int r = 0; r < 4; r++) s_found[r] = 0x7fffffff;
    }
    __syncwarp();
    Decoded out; out.id = -1; out.nrot = 0;
    if (s_ok) {
        for (int i = tid; i < g.ncodes; i += kStride) {
            const unsigned long long c = codes[i];
#pragma unroll
            for (int r = 0; r < 4; r++) if (c == s_ids[r]) atomicMin(&s_found[r], i);      // first index wins for repeated codes
        }
    }
    __syncwarp();
    if (tid == 0) {
        if (s_ok) for (int r = 0; r < 4; r++) if (s_found[r] != 0x7fffffff) { out.id = s_found[r]; out.nrot = r; break; }
        dec0[(long long)f * kMaxCand + k] = out;
    }
    }   // candidates of this warp
}

// ------------------------------------------------------------------------------------------------
// sort / de-duplicate / CORNER_LINES refinement, one CTA per frame
// ------------------------------------------------------------------------------------------------
// cv::solve(A (m x 2), b, DECOMP_SVD) for A = [t 1]: one-sided Jacobi on the two columns + back-substitution.
// Columns live in global scratch (a0[i], a1[i]); rhs in bb[i].  Warp-cooperative; returns x0, x1 on all lanes.
__device__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ void svd_solve_m2(float* a0, float* a1, const float* bb, int m, int lane, float& x0, float& x1) {
    const float eps = FLT_EPSILON * 2;
    double W0 = 0, W1 = 0;
    for (int k = lane; k < m; k += 32) { W0 += (double)a0[k] * a0[k]; W1 += (double)a1[k] * a1[k]; }
    W0 = warp_sum(W0); W1 = warp_sum(W1);
    float v00 = 1.f, v01 = 0.f, v10 = 0.f, v11 = 1.f;        // Vt rows
    const int max_iter = max(m, 30);
    for (int iter = 0; iter < max_iter; iter++) {
        double p = 0;
        for (int k = lane; k < m; k += 32) p += (double)a0[k] * a1[k];
        p = warp_sum(p);
        if (fabs(p) <= eps * sqrt(W0 * W1)) break;
        p *= 2;
        const double beta = W0 - W1, gamma = hypot(p, beta);
        float c, s;
        if (beta < 0) {
            const double delta = (gamma - beta) * 0.5;
            s = (float)sqrt(delta / gamma);
            c = (float)(p / (gamma * s * 2));
        } else {
            c = (float)sqrt((gamma + beta) / (gamma * 2));
            s = (float)(p / (gamma * c * 2));
        }
        double a = 0, b = 0;
        for (int k = lane; k < m; k += 32) {
            const float t0 = __fadd_rn(__fmul_rn(c, a0[k]), __fmul_rn(s, a1[k]));
            const float t1 = __fadd_rn(__fmul_rn(-s, a0[k]), __fmul_rn(c, a1[k]));
            a0[k] = t0; a1[k] = t1;
            a += (double)t0 * t0; b += (double)t1 * t1;
        }
        W0 = warp_sum(a); W1 = warp_sum(b);
        {
            float t0 = __fadd_rn(__fmul_rn(c, v00), __fmul_rn(s, v10)), t1 = __fadd_rn(__fmul_rn(-s, v00), __fmul_rn(c, v10));
            v00 = t0; v10 = t1;
            t0 = __fadd_rn(__fmul_rn(c, v01), __fmul_rn(s, v11)); t1 = __fadd_rn(__fmul_rn(-s, v01), __fmul_rn(c, v11));
            v01 = t0; v11 = t1;
        }
        __syncwarp();
    }
    double s0 = 0, s1 = 0;
    for (int k = lane; k < m; k += 32) { s0 += (double)a0[k] * a0[k]; s1 += (double)a1[k] * a1[k]; }
    double w0 = sqrt(warp_sum(s0)), w1 = sqrt(warp_sum(s1));
    bool swapped = false;
    if (w0 < w1) { const double t = w0; w0 = w1; w1 = t; swapped = true; }
    float* u0 = swapped ? a1 : a0; float* u1 = swapped ? a0 : a1;
    const float vt0x = swapped ? v10 : v00, vt0y = swapped ? v11 : v01, vt1x = swapped ? v00 : v10, vt1y = swapped ? v01 : v11;
    const float wf0 = (float)w0, wf1 = (float)w1;
    const float sc0 = (float)(w0 > FLT_MIN ? 1 / w0 : 0.), sc1 = (float)(w1 > FLT_MIN ? 1 / w1 : 0.);
    double d0 = 0, d1 = 0;       // u_i . b with u = column / w (float), float products, double sums
    for (int k = lane; k < m; k += 32) {
        d0 += (double)__fmul_rn(__fmul_rn(u0[k], sc0), bb[k]);
        d1 += (double)__fmul_rn(__fmul_rn(u1[k], sc1), bb[k]);
    }
    d0 = warp_sum(d0); d1 = warp_sum(d1);
    const double threshold = ((double)wf0 + (double)wf1) * eps;
    float r0 = 0.f, r1 = 0.f;
    if (fabs((double)wf0) > threshold) { const double sv = d0 * (1 / (double)wf0); r0 = (float)(r0 + sv * vt0x); r1 = (float)(r1 + sv * vt0y); }
    if (fabs((double)wf1) > threshold) { const double sv = d1 * (1 / (double)wf1); r0 = (float)(r0 + sv * vt1x); r1 = (float)(r1 + sv * vt1y); }
    x0 = r0; x1 = r1;
}

// 2x2 system of getCrossPoint (11899-12049), same SVD path, single thread
__device__ void svd_solve_2x2(const float A[4], const float B[2], float X[2]) {
    float a0[2] = {A[0], A[2]}, a1[2] = {A[1], A[3]};      // columns
    const float eps = FLT_EPSILON * 2;
    double W0 = (double)a0[0] * a0[0] + (double)a0[1] * a0[1], W1 = (double)a1[0] * a1[0] + (double)a1[1] * a1[1];
    float v00 = 1.f, v01 = 0.f, v10 = 0.f, v11 = 1.f;
    for (int iter = 0; iter < 30; iter++) {
        double p = (double)a0[0] * a1[0] + (double)a0[1] * a1[1];
        if (fabs(p) <= eps * sqrt(W0 * W1)) break;
        p *= 2;
        const double beta = W0 - W1, gamma = hypot(p, beta);
        float c, s;
        if (beta < 0) { const double delta = (gamma - beta) * 0.5; s = (float)sqrt(delta / gamma); c = (float)(p / (gamma * s * 2)); }
        else { c = (float)sqrt((gamma + beta) / (gamma * 2)); s = (float)(p / (gamma * c * 2)); }
        double a = 0, b = 0;
        for (int k = 0; k < 2; k++) {
            const float t0 = __fadd_rn(__fmul_rn(c, a0[k]), __fmul_rn(s, a1[k])), t1 = __fadd_rn(__fmul_rn(-s, a0[k]), __fmul_rn(c, a1[k]));
            a0[k] = t0; a1[k] = t1; a += (double)t0 * t0; b += (double)t1 * t1;
        }
        W0 = a; W1 = b;
        float t0 = __fadd_rn(__fmul_rn(c, v00), __fmul_rn(s, v10)), t1 = __fadd_rn(__fmul_rn(-s, v00), __fmul_rn(c, v10));
        v00 = t0; v10 = t1;
        t0 = __fadd_rn(__fmul_rn(c, v01), __fmul_rn(s, v11)); t1 = __fadd_rn(__fmul_rn(-s, v01), __fmul_rn(c, v11));
        v01 = t0; v11 = t1;
    }
    double w0 = sqrt((double)a0[0] * a0[0] + (double)a0[1] * a0[1]), w1 = sqrt((double)a1[0] * a1[0] + (double)a1[1] * a1[1]);
    bool swapped = false;
    if (w0 < w1) { const double t = w0; w0 = w1; w1 = t; swapped = true; }
    const float* u0 = swapped ? a1 : a0; const float* u1 = swapped ? a0 : a1;
    const float vt0x = swapped ? v10 : v00, vt0y = swapped ? v11 : v01, vt1x = swapped ? v00 : v10, vt1y = swapped ? v01 : v11;
    const float wf0 = (float)w0, wf1 = (float)w1;
    const float sc0 = (float)(w0 > FLT_MIN ? 1 / w0 : 0.), sc1 = (float)(w1 > FLT_MIN ? 1 / w1 : 0.);
    const double d0 = (double)__fmul_rn(__fmul_rn(u0[0], sc0), B[0]) + (double)__fmul_rn(__fmul_rn(u0[1], sc0), B[1]);
    const double d1 = (double)__fmul_rn(__fmul_rn(u1[0], sc1), B[0]) + (double)__fmul_rn(__fmul_rn(u1[1], sc1), B[1]);
    const double threshold = ((double)wf0 + (double)wf1) * eps;
    float r0 = 0.f, r1 = 0.f;
    if (fabs((double)wf0) > threshold) { const double sv = d0 * (1 / (double)wf0); r0 = (float)(r0 + sv * vt0x); r1 = (float)(r1 + sv * vt0y); }
    if (fabs((double)wf1) > threshold) { const double sv = d1 * (1 / (double)wf1); r0 = (float)(r0 + sv * vt1x); r1 = (float)(r1 + sv * vt1y); }
    X[0] = r0; X[1] = r1;
}

struct FMin { float v; int i; };
__device__ __forceinline__ FMin warp_argmin_first(float v, int i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    FMin r; r.v = v; r.i = i; return r;
}

constexpr int kFinWarps = 20;       // at most; batches above 32 frames launch 8 warps per frame (latency of the refinement against SM slots for the other kernels)

__global__ void __launch_bounds__(kFinWarps * 32)
k_finalize(const __grid_constant__ ArucoGeom g, const Kept* __restrict__ kept0, const int* __restrict__ nkept,
           const Decoded* __restrict__ dec0, const ContourDesc* __restrict__ desc0, const short2* __restrict__ pts0,
           float* __restrict__ scratch0, b200_marker* __restrict__ out0, int* __restrict__ counts, int out_cap, int* __restrict__ err,
           int* __restrict__ mcontour0) {
    __shared__ int s_idx[kMaxCand], s_id[kMaxCand], s_perim[kMaxCand], s_n, s_m;
    __shared__ float s_c[kMaxCand][8];
    __shared__ unsigned char s_rm[kMaxCand];
    __shared__ int s_final[kMaxMarkers];
    __shared__ int s_src[kMaxCand];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nk = min(nkept[f], kMaxCand);
    const Kept* kept = kept0 + (long long)f * kMaxCand;
    const Decoded* dec = dec0 + (long long)f * kMaxCand;
    if (tid == 0) {
        // successful candidates in candidate order, corners rotated (std::rotate(begin, begin + 4 - nRot, end), 6779)
        int n = 0;
        for (int k = 0; k < nk; k++) if (dec[k].id >= 0) { s_idx[n] = k; n++; }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    for (int i = tid; i < n; i += blockDim.x) {
        const int k = s_idx[i];
        const Decoded d = dec[k];
        // stable sort by id (8159): rank = markers with a smaller id, or the same id earlier in candidate order
        int rank = 0;
        for (int j = 0; j < n; j++) { const int idj = dec[s_idx[j]].id; rank += (idj < d.id) || (idj == d.id && j < i); }
        float q[8];
        for (int c = 0; c < 4; c++) { const int s = (c + 4 - d.nrot) & 3; q[2 * c] = kept[k].c[2 * s]; q[2 * c + 1] = kept[k].c[2 * s + 1]; }
        for (int c = 0; c < 8; c++) s_c[rank][c] = q[c];
        s_id[rank] = d.id; s_perim[rank] = perimeter_i(q); s_rm[rank] = 0;
        s_src[rank] = k;                      // which kept candidate (=> contour) the sorted marker came from
    }
    __syncthreads();
    if (tid == 0) {
        // duplicate removal (8159-8311): serial flag semantics
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n && !s_rm[i]; j++)
                if (s_id[i] == s_id[j]) { if (s_perim[i] < s_perim[j]) s_rm[i] = 1; else s_rm[j] = 1; }
        int m = 0;
        for (int i = 0; i < n; i++) if (!s_rm[i]) { if (m < kMaxMarkers && m < out_cap) s_final[m++] = i; else atomicExch(err, 7); }
        s_m = m;
        counts[f] = m;
    }
    __syncthreads();
    const int m = s_m;
    b200_marker* out = out0 + (long long)f * out_cap;
    // CORNER_LINES refinement (8979-12049): one warp per marker
    for (int mi = warp; mi < m; mi += (int)(blockDim.x >> 5)) {
        const int i = s_final[mi];
        const ContourDesc cd = desc0[(long long)f * g.max_contours + kept[s_src[i]].contour];
        const short2* cp = pts0 + (long long)f * g.max_points + cd.off;
        const int nc = cd.len;
        float c[8];
        for (int k = 0; k < 8; k++) c[k] = s_c[i][k];
        int ci[4];
        for (int k = 0; k < 4; k++) {
            float best = FLT_MAX; int bi = 0x7fffffff;
            for (int j = lane; j < nc; j += 32) {
                const float dx = __fsub_rn((float)cp[j].x, c[2 * k]), dy = __fsub_rn((float)cp[j].y, c[2 * k + 1]);
                const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                if (d < best) { best = d; bi = j; }
            }
            const FMin r = warp_argmin_first(best, bi);
            ci[k] = r.v < FLT_MAX ? r.i : -1;
        }
        bool inverse;
        if ((ci[1] > ci[0]) && (ci[2] > ci[1] || ci[2] < ci[0])) inverse = false;
        else if (ci[2] > ci[1] && ci[2] < ci[0]) inverse = false;
        else inverse = true;
        // the four sides as index walks (8647-8686 incl. its wrap-around quirks): side l = {first index, count}
        float line[4][3];
        bool ok = true;
        float* sa0 = scratch0 + ((long long)f * g.max_points + cd.off) * 3;      // 3 floats per contour point: a0, a1, b
        for (int l = 0; l < 4 && ok; l++) {
            const int stop = ci[(l + 1) & 3];
            // the reference loop `for (j = ci[l]; j != stop; j += inc) { wrap; push(pts[j]); if (j == stop) break; }` in closed
            // form: first visited index and number of points.  Quirks kept: walking forwards the end corner is included
            // only when it is index 0 (reached through the wrap); walking backwards index 0 is replaced by n-1 before
            // use (pts[0] is never taken) and the end corner is included only when it is n-1.
            const int start = ci[l];
            int cnt = 0, first = start;
            if (start != stop) {
                if (!inverse) {
                    if (stop > start) cnt = stop - start;
                    else cnt = (nc - start) + stop + (stop == 0 ? 1 : 0);
                } else if (start == 0) {
                    first = nc - 1;
                    cnt = (stop == nc - 1) ? 1 : (nc - 1 - stop);
                } else {
                    if (stop < start) cnt = start - stop;
                    else cnt = start + ((stop == nc - 1) ? 1 : (nc - 1 - stop));
                }
            }
            if (cnt < 2) { ok = false; break; }
            float* a0 = sa0;                         // sides are processed one after the other: the block is reused
            float* a1 = a0 + nc; float* bb = a1 + nc;    // every side has at most nc points
            // bounding box
            float minx = FLT_MAX, maxx = -FLT_MAX, miny = FLT_MAX, maxy = -FLT_MAX;
            for (int t = lane; t < cnt; t += 32) {
                int q;
                if (!inverse) { q = first + t; if (q >= nc) q -= nc; }
                else { q = first - t; if (q <= 0) q += nc - 1; }
                const float x = (float)cp[q].x, y = (float)cp[q].y;
                minx = fminf(minx, x); maxx = fmaxf(maxx, x); miny = fminf(miny, y); maxy = fmaxf(maxy, y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                minx = fminf(minx, __shfl_xor_sync(0xffffffffu, minx, o)); maxx = fmaxf(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
                miny = fminf(miny, __shfl_xor_sync(0xffffffffu, miny, o)); maxy = fmaxf(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
            }
            const bool xmajor = __fsub_rn(maxx, minx) > __fsub_rn(maxy, miny);
            for (int t = lane; t < cnt; t += 32) {
                int q;
                if (!inverse) { q = first + t; if (q >= nc) q -= nc; }
                else { q = first - t; if (q <= 0) q += nc - 1; }
                const float x = (float)cp[q].x, y = (float)cp[q].y;
                a0[t] = xmajor ? x : y; a1[t] = 1.f; bb[t] = xmajor ? y : x;
            }
            __syncwarp();
            float x0, x1;
            svd_solve_m2(a0, a1, bb, cnt, lane, x0, x1);
            if (xmajor) { line[l][0] = x0; line[l][1] = -1.f; line[l][2] = x1; }
            else { line[l][0] = -1.f; line[l][1] = x0; line[l][2] = x1; }
            __syncwarp();
        }
        if (ok) {
            for (int k = 0; k < 4; k++) {
                const float* l1 = line[(k + 3) & 3]; const float* l2 = line[k];
                const float A[4] = {l1[0], l1[1], l2[0], l2[1]}, B[2] = {-l1[2], -l2[2]};
                float X[2];
                svd_solve_2x2(A, B, X);
                c[2 * k] = X[0]; c[2 * k + 1] = X[1];
            }
        }
        if (lane == 0) {
            out[mi].id = s_id[i];
            for (int k = 0; k < 8; k++) out[mi].xy[k] = c[k];
            mcontour0[(long long)f * kMaxMarkers + mi] = kept[s_src[i]].contour;        // aruco::Marker::contourPoints (b200_aruco_get_contour)
        }
    }
}

// packs the borders of a frame's output markers back to back: offsets by one warp scan, points copied by the whole CTA
__global__ void __launch_bounds__(256)
k_pack_contours(const __grid_constant__ ArucoGeom g, const int* __restrict__ mcontour, int n_markers_host, const int* __restrict__ n_markers_dev,
                const ContourDesc* __restrict__ desc, const short2* __restrict__ pts, int* __restrict__ ofs, short2* __restrict__ out, int out_cap) {
    __shared__ int s_ofs[kMaxMarkers + 1], s_src[kMaxMarkers];
    const int tid = threadIdx.x;
    const int n_markers = n_markers_dev ? min(*n_markers_dev, kMaxMarkers) : n_markers_host;       // the count may only exist on the device yet
    if (tid < 32) {
        int run = 0;
        for (int m0 = 0; m0 < n_markers; m0 += 32) {
            const int m = m0 + tid;
            int len = 0, src = 0;
            if (m < n_markers) {
                const int ci = mcontour[m];
                if (ci >= 0 && ci < g.max_contours) { const ContourDesc c = desc[ci]; if (c.len > 0 && c.off >= 0 && c.off + c.len <= g.max_points) { len = c.len; src = c.off; } }
            }
            int sc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (tid >= o) sc += t; }
            if (m < n_markers) { s_ofs[m] = run + sc - len; s_src[m] = src; }
            run += __shfl_sync(0xffffffffu, sc, 31);
        }
        if (tid == 0) s_ofs[n_markers] = run;
    }
    __syncthreads();
    for (int m = tid; m <= n_markers; m += blockDim.x) ofs[m] = s_ofs[m];
    if (s_ofs[n_markers] > g.max_points) return;
    for (int m = 0; m < n_markers; m++) {
        const int len = s_ofs[m + 1] - s_ofs[m];
        for (int i = tid; i < len; i += blockDim.x) if (s_ofs[m] + i < out_cap) out[s_ofs[m] + i] = pts[s_src[m] + i];
    }
}

}  // namespace b200

// =================================================================================================
// host side
// =================================================================================================
using namespace b200;

struct b200_aruco_s {
    int device;
    cudaStream_t stream;
    int max_w, max_h, max_batch;
    int nbits, ncodes;
    std::string dict;
    unsigned long long* d_codes;
    int cur_w, cur_h;
    ArucoGeom geom;
    uint8_t *d_mask, *d_pyr; int* d_surv; int max_surv;
    ContourDesc* d_desc; short2* d_pts; float* d_scratch;
    Candidate* d_cand; Kept* d_kept; Decoded* d_dec;
    int *d_ncont, *d_npts, *d_ncand, *d_nkept, *d_nsurv, *d_nfetch, *d_err;
    int* d_surv2; int* d_nsurv2; size_t cap_surv2;
    int* d_smap; unsigned long long* d_nodes; uint32_t* d_sbits; int sbits_words; size_t cap_smap, cap_nodes, cap_sbits; int ring_slots, ring_mode;      // ring form of the border walks: raster key -> survivor; {successor, steps to it} per survivor
    int *d_nfetch2, *d_nfetch3;      // d_nfetch2: per-frame allocator of the checkpoint pool
    int* d_ckpool; int ck_cap; size_t cap_ckpool;
    uint8_t* d_wpatch; uint16_t* d_whist; int* d_wlevel;    // warped patches, their histograms and Otsu levels: [B][256][...]      // transitions that survive the backward check (phase B1)
    size_t cap_mask, cap_pyr, cap_desc, cap_pts, cap_scratch, cap_surv;
    // staging for the host API
    uint8_t* d_in; size_t cap_in; b200_marker* d_out; int* d_counts; size_t cap_out;
    int* d_mcontour;                          // [max_batch][kMaxMarkers]: the border (ContourDesc index) every output marker came from
    void* d_pack;                             // staging of b200_aruco_get_contours: offsets + packed points of one frame
    cudaStream_t pyr_stream; cudaEvent_t ev_pyr_fork, ev_pyr_join;      // the half pyramid runs beside the contour kernels (only k_decode reads it)
    void* d_poses; void* h_pack; size_t h_pack_cap;      // b200_aruco_detect_frame_host: device poses, pinned staging of the packed points
};

namespace {

struct DictTable { const char* name; int nbits; int n; const unsigned long long* codes; };
#define DICT_BEGIN(NAME, NBITS, TAU, N) static const unsigned long long codes_##NAME[] = {
#define C(x) x##ULL,
#define DICT_END(NAME) };
#include "aruco_dicts.inc"
#undef DICT_BEGIN
#undef C
#undef DICT_END
#define DICT_BEGIN(NAME, NBITS, TAU, N) {#NAME, NBITS, N, codes_##NAME},
#define C(x)
#define DICT_END(NAME)
static const DictTable kDicts[] = {
#include "aruco_dicts.inc"
};
#undef DICT_BEGIN
#undef C
#undef DICT_END

template <typename T> int ensure_buf(T*& p, size_t& cap, size_t need) {
    if (need <= cap && p) return B200_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    B200_CUDA(cudaMalloc((void**)&p, need));
    cap = need;
    return B200_OK;
}

int aruco_geometry(b200_aruco_s* h, int w, int hh) {
    if (w == h->cur_w && hh == h->cur_h) return B200_OK;
    ArucoGeom& g = h->geom;
    memset(&g, 0, sizeof(g));
    g.w = w; g.h = hh;
    g.bpitch = (int)align_up(w + 2 * kMaskPad, 16);
    g.bframe = (long long)g.bpitch * (hh + 2);
    int win = std::max(3, (int)(15 * float(w) / 1920.));         // markerdetector_impl.cpp:3769-3867
    if (win % 2 == 0) win++;
    if (win > 2 * kThrMaxR + 1) return fail(B200_EINVAL, "image wider than %s", "the 15-px threshold window supports (1920)");
    g.win = win;
    {   // rint(S / win^2) as a multiply-shift: verified exhaustively for every reachable window sum
        const int a = win * win, half = a / 2;
        g.mean_mul = (unsigned)(((1ull << 24) + a - 1) / a);
        for (int S = 0; S <= 255 * a; S++)
            if ((int)(((unsigned long long)(unsigned)(S + half) * g.mean_mul) >> 24) != (S + half) / a) return fail(B200_EINVAL, "mean multiplier inexact for window %s", "size");
    }
    g.nbits = h->nbits; g.nb = (int)sqrt((double)h->nbits); g.nsub = g.nb + 2; g.wsize = 5 * g.nsub; g.ncodes = h->ncodes;
    // pyramid (1300-1466): halve while width > 2*warpSize
    g.lw[0] = w; g.lh[0] = hh; g.nlev = 1;
    {
        int cw = w, nl = 1;
        while (cw > 2 * g.wsize) { cw /= 2; nl++; }
        long long off = 0;
        for (int l = 1; l < nl && l < kMaxPyr; l++) {
            const int lw = g.lw[l - 1] / 2, lh = g.lh[l - 1] / 2;
            if (lw < 1 || lh < 1) break;
            g.lw[l] = lw; g.lh[l] = lh; g.lpitch[l] = (int)align_up(lw, 16); g.loff[l] = off;
            off += align_up((long long)g.lpitch[l] * lh, 256);
            g.nlev = l + 1;
        }
        g.pyr_frame = std::max<long long>(off, 256);
    }
    g.max_points = w * hh;
    g.max_contours = w * hh / (kMinContour + 1) + 1;       // borders longer than 70 points sharing w*h points
    const size_t B = (size_t)h->max_batch;
    int rc;
    if ((rc = ensure_buf(h->d_mask, h->cap_mask, (size_t)g.bframe * B))) return rc;
    B200_CUDA(cudaMemset(h->d_mask, 0, (size_t)g.bframe * B));       // the 1-px zero frame is never written afterwards
    h->max_surv = std::max(1024, w * hh);              // transition list: at most two entries per foreground pixel
    if ((rc = ensure_buf(h->d_surv, h->cap_surv, sizeof(int) * (size_t)h->max_surv * B))) return rc;
    if ((rc = ensure_buf(h->d_surv2, h->cap_surv2, sizeof(int) * (size_t)h->max_surv * B))) return rc;
    h->ck_cap = std::max(1024, w * hh / 16) / kCkPerBorder * kCkPerBorder;       // checkpoints of the borders that are followed for 256 steps and more
    if ((rc = ensure_buf(h->d_ckpool, h->cap_ckpool, sizeof(int) * (size_t)(h->ck_cap + h->ck_cap / kCkPerBorder) * B))) return rc;
    {   // buffers of the ring form of the border walks (12 bytes per pixel and frame slot): it serves batches of up to kRingFrames frames, so that is
        // how many slots it gets (B200_CONTOURS_RING=1 forces it for every batch size, =0 disables it)
        const char* e = getenv("B200_CONTOURS_RING");
        h->ring_mode = e ? atoi(e) : -1;
        h->ring_slots = h->ring_mode == 0 ? 0 : h->ring_mode == 1 ? (int)B : (int)std::min<size_t>(B, kRingFrames);
        if (h->ring_slots > 0 && g.bframe < (1ll << kNodeBits) && h->max_surv < (1 << kNodeBits)) {
            const size_t R = (size_t)h->ring_slots;
            if ((rc = ensure_buf(h->d_smap, h->cap_smap, sizeof(int) * (size_t)g.bframe * R))) return rc;
            B200_CUDA(cudaMemset(h->d_smap, 0, sizeof(int) * (size_t)g.bframe * R));      // k_emit2 leaves it zero again after every call
            if ((rc = ensure_buf(h->d_nodes, h->cap_nodes, sizeof(unsigned long long) * (size_t)h->max_surv * R))) return rc;
            h->sbits_words = (int)((g.bframe + 63) / 64 * 2 + 2);        // survivor bitmap: one bit per raster key, read as 64-bit words
            if ((rc = ensure_buf(h->d_sbits, h->cap_sbits, sizeof(uint32_t) * (size_t)h->sbits_words * R))) return rc;
            B200_CUDA(cudaMemset(h->d_sbits, 0, sizeof(uint32_t) * (size_t)h->sbits_words * R));
        } else h->ring_slots = 0;
    }
    if ((rc = ensure_buf(h->d_pyr, h->cap_pyr, (size_t)g.pyr_frame * B))) return rc;
    if ((rc = ensure_buf(h->d_desc, h->cap_desc, sizeof(ContourDesc) * (size_t)g.max_contours * B))) return rc;
    if ((rc = ensure_buf(h->d_pts, h->cap_pts, sizeof(short2) * (size_t)g.max_points * B))) return rc;
    if ((rc = ensure_buf(h->d_scratch, h->cap_scratch, sizeof(float) * 3 * ((size_t)g.max_points * B + 64)))) return rc;
    h->cur_w = w; h->cur_h = hh;
    return B200_OK;
}

}  // namespace

extern "C" {

int b200_aruco_create(b200_aruco_t* out, const char* dict_name, int max_w, int max_h, int max_batch, int device) {
    if (!out) return fail(B200_EINVAL, "null %s", "out");
    *out = nullptr;
    if (!dict_name || max_w < 1 || max_h < 1 || max_batch < 1) return fail(B200_EINVAL, "bad %s parameters", "detector");
    const DictTable* dt = nullptr;
    for (const auto& d : kDicts) if (std::string(d.name) == dict_name) dt = &d;
    if (!dt) return fail(B200_EINVAL, "unknown dictionary '%s'", dict_name);
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    b200_aruco_s* h = new (std::nothrow) b200_aruco_s();
    if (!h) return B200_ENOMEM;
    h->device = device; h->max_w = max_w; h->max_h = max_h; h->max_batch = max_batch; h->cur_w = h->cur_h = -1;
    h->dict = dict_name; h->nbits = dt->nbits; h->ncodes = dt->n;
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->pyr_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_pyr_fork, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&h->ev_pyr_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_codes, sizeof(unsigned long long) * dt->n) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_codes, dt->codes, sizeof(unsigned long long) * dt->n, cudaMemcpyHostToDevice) == cudaSuccess;
    const size_t B = (size_t)max_batch;
    ok = ok && cudaMalloc((void**)&h->d_cand, sizeof(Candidate) * kMaxCand * B) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_kept, sizeof(Kept) * kMaxCand * B) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_dec, sizeof(Decoded) * kMaxCand * B) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_wpatch, (size_t)kMaxWarp * kMaxWarp * kMaxCand * B) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_whist, sizeof(uint16_t) * 256 * kMaxCand * B) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_wlevel, sizeof(int) * kMaxCand * B) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_mcontour, sizeof(int) * kMaxMarkers * (size_t)B) == cudaSuccess;
    if (ok) cudaMemset(h->d_mcontour, 0xff, sizeof(int) * kMaxMarkers * (size_t)B);
    ok = ok && cudaMalloc((void**)&h->d_ncont, 9 * B * 4 + 4) == cudaSuccess;
    if (!ok) { b200_aruco_destroy(h); return fail(B200_ECUDA, "%s failed", "allocation"); }
    h->d_npts = h->d_ncont + B; h->d_ncand = h->d_npts + B; h->d_nkept = h->d_ncand + B; h->d_nsurv = h->d_nkept + B; h->d_nfetch = h->d_nsurv + B; h->d_nsurv2 = h->d_nfetch + B; h->d_nfetch2 = h->d_nsurv2 + B; h->d_nfetch3 = h->d_nfetch2 + B; h->d_err = h->d_nfetch3 + B;
    cudaMemset(h->d_ncont, 0, 9 * B * 4 + 4);
    if ((rc = aruco_geometry(h, max_w, max_h))) { b200_aruco_destroy(h); return rc; }
    *out = h;
    return B200_OK;
}

int b200_aruco_destroy(b200_aruco_t h) {
    if (!h) return B200_OK;
    DeviceScope _ds; cudaSetDevice(h->device);
    cudaFree(h->d_codes); cudaFree(h->d_surv); cudaFree(h->d_mask); cudaFree(h->d_pyr); cudaFree(h->d_desc); cudaFree(h->d_pts);
    cudaFree(h->d_scratch); cudaFree(h->d_cand); cudaFree(h->d_kept); cudaFree(h->d_dec); cudaFree(h->d_ncont); cudaFree(h->d_surv2); cudaFree(h->d_ckpool); cudaFree(h->d_smap); cudaFree(h->d_nodes); cudaFree(h->d_sbits); cudaFree(h->d_wpatch); cudaFree(h->d_whist); cudaFree(h->d_wlevel); cudaFree(h->d_mcontour); cudaFree(h->d_pack); cudaFree(h->d_poses); if (h->h_pack) cudaFreeHost(h->h_pack);
    cudaFree(h->d_in); cudaFree(h->d_out); cudaFree(h->d_counts);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->pyr_stream) cudaStreamDestroy(h->pyr_stream);
    if (h->ev_pyr_fork) cudaEventDestroy(h->ev_pyr_fork);
    if (h->ev_pyr_join) cudaEventDestroy(h->ev_pyr_join);
    delete h;
    return B200_OK;
}

int b200_aruco_batch_capacity(b200_aruco_t h) { return h ? h->max_batch : 0; }

int b200_aruco_max_markers(b200_aruco_t h) {
    (void)h;                       // a library-wide constant: NULL asks for it without a handle
    return kMaxMarkers;
}

// `base`: first scratch frame slot (calls that may overlap on different streams use disjoint slot ranges of the handle)
int b200_aruco_detect_range(b200_aruco_t h, const uint8_t* imgs, int n, int w, int hh, int64_t rs, int64_t fs,
                            b200_marker* markers, int32_t* counts, int base, void* stream) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    if (n < 0 || w < 0 || hh < 0) return fail(B200_EINVAL, "negative %s", "size");
    if (base < 0 || base + n > h->max_batch || w > h->max_w || hh > h->max_h) return fail(B200_ECAPACITY, "batch/image larger than the handle's %s", "capacity");
    if (n == 0) return B200_OK;
    if (!markers || !counts) return fail(B200_EINVAL, "null %s", "output pointer");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (w < 8 || hh < 8) { B200_CUDA(cudaMemsetAsync(counts, 0, (size_t)n * 4, st)); return B200_OK; }
    if (!imgs) return fail(B200_EINVAL, "null %s", "image pointer");
    if (rs < w || (n > 1 && fs < rs * (hh - 1) + w)) return fail(B200_EINVAL, "bad %s", "strides");
    if ((rc = aruco_geometry(h, w, hh))) return rc;
    const ArucoGeom& g = h->geom;
    // the seven per-frame counter arrays are one [7][max_batch] block: clear this call's columns
    B200_CUDA(cudaMemset2DAsync(h->d_ncont + base, (size_t)h->max_batch * 4, 0, (size_t)n * 4, 9, st));
    uint8_t* d_mask = h->d_mask + (size_t)base * g.bframe;
    uint8_t* d_pyr = h->d_pyr + (size_t)base * g.pyr_frame;
    int* d_surv = h->d_surv + (size_t)base * h->max_surv;
    int* d_surv2 = h->d_surv2 + (size_t)base * h->max_surv;
    ContourDesc* d_desc = h->d_desc + (size_t)base * g.max_contours;
    short2* d_pts = h->d_pts + (size_t)base * g.max_points;
    float* d_scratch = h->d_scratch + 3 * (size_t)base * g.max_points;
    Candidate* d_cand = h->d_cand + (size_t)base * kMaxCand;
    Kept* d_kept = h->d_kept + (size_t)base * kMaxCand;
    Decoded* d_dec = h->d_dec + (size_t)base * kMaxCand;
    int *d_ncont = h->d_ncont + base, *d_npts = h->d_npts + base, *d_ncand = h->d_ncand + base, *d_nkept = h->d_nkept + base,
        *d_nsurv = h->d_nsurv + base, *d_nfetch = h->d_nfetch + base, *d_nsurv2 = h->d_nsurv2 + base;
    dim3 blk(32, 8);
    dim3 gt((w + kThrTW - 1) / kThrTW, (hh + kThrTH - 1) / kThrTH, n);
    static const bool athresh_smem = getenv("B200_ATHRESH_SMEM") != nullptr;      // the round-1 kernel (shared-memory box sums), kept for comparison
    if (athresh_smem || rs * (long long)hh >= (1ll << 31)) B200_LAUNCH(k_athresh, gt, blk, 0, st, imgs, rs, fs, g, d_mask);      // (k_athresh2 keeps row offsets in 32 bits)
    else {
        const int R = g.win >> 1, nout = 4 * (30 - 2 * (R <= 3 ? 1 : 2));
        const int rows = n >= 16 ? kAt2Rows : 16;          // few frames: shorter strips, four times as many warps (the window warm-up is paid more often)
        dim3 g2((w + nout - 1) / nout, ((hh + rows - 1) / rows + kAt2Warps - 1) / kAt2Warps, n);
        switch (R) {
            case 1: B200_LAUNCH(k_athresh2<1>, g2, kAt2Warps * 32, 0, st, imgs, rs, fs, g, d_mask, rows); break;
            case 2: B200_LAUNCH(k_athresh2<2>, g2, kAt2Warps * 32, 0, st, imgs, rs, fs, g, d_mask, rows); break;
            case 3: B200_LAUNCH(k_athresh2<3>, g2, kAt2Warps * 32, 0, st, imgs, rs, fs, g, d_mask, rows); break;
            case 4: B200_LAUNCH(k_athresh2<4>, g2, kAt2Warps * 32, 0, st, imgs, rs, fs, g, d_mask, rows); break;
            case 5: B200_LAUNCH(k_athresh2<5>, g2, kAt2Warps * 32, 0, st, imgs, rs, fs, g, d_mask, rows); break;
            case 6: B200_LAUNCH(k_athresh2<6>, g2, kAt2Warps * 32, 0, st, imgs, rs, fs, g, d_mask, rows); break;
            default: B200_LAUNCH(k_athresh2<7>, g2, kAt2Warps * 32, 0, st, imgs, rs, fs, g, d_mask, rows); break;
        }
    }
    // the half pyramid is read by k_decode only: it runs on the handle's second stream beside the contour kernels and joins in front of k_decode
    B200_CUDA(cudaEventRecord(h->ev_pyr_fork, st));
    B200_CUDA(cudaStreamWaitEvent(h->pyr_stream, h->ev_pyr_fork, 0));
    for (int l = 1; l < g.nlev; l++) {
        const uint8_t* src = l == 1 ? imgs : d_pyr + g.loff[l - 1];
        const long long srs = l == 1 ? rs : g.lpitch[l - 1], sfs = l == 1 ? fs : g.pyr_frame;
        dim3 gp((g.lw[l] + 31) / 32, (g.lh[l] + 7) / 8, n);
        B200_LAUNCH(k_halfpyr, gp, blk, 0, h->pyr_stream, src, srs, sfs, g.lw[l - 1], g.lh[l - 1], d_pyr + g.loff[l], g.lpitch[l], g.pyr_frame, g.lw[l], g.lh[l]);
    }
    B200_CUDA(cudaEventRecord(h->ev_pyr_join, h->pyr_stream));
    // contour following: against the frame's bit image in shared memory when it fits (one CTA per frame), else the global-memory walkers
    const int ct_pw = (w + 1) / 30 + 1, ct_npw = (w + 31) / 32;
    const size_t ct_smem = ((size_t)(hh + 2) * ct_pw + (size_t)kCtWarps * (ct_npw + 4)) * 4;
    // Default: the global-memory walkers (k_probe_a / b1 / b + k_emit), many CTAs per frame.  B200_CONTOURS_SHARED=1: k_contours, one CTA per frame against the
    // bit image in shared memory.  Measured on a B200 (profiles/r2f_variants.txt, r2i): the same 256 x 640 x 480 step time (4.47 vs 4.51 ms: the step is bound by
    // pipe throughput, and building the 8-neighbour mask from three bit rows costs ~12 ALU instructions where the global form does one load), slower at
    // 1280 x 720 (124 KB of shared memory: one CTA per SM) and 3x slower for a single frame (one CTA does the whole frame: 2.3 vs 0.8 ms per detect call).
    static const bool ct_want_shared = getenv("B200_CONTOURS_SHARED") != nullptr;
    const bool ct_shared = ct_want_shared && ct_smem <= 200 * 1024 && w < 32768 / 2;
    if (ct_shared) {
        static std::atomic<size_t> ct_smem_set(0);
        if (ct_smem > 48 * 1024 && ct_smem > ct_smem_set.load()) { B200_CUDA(cudaFuncSetAttribute(k_contours, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct_smem)); ct_smem_set.store(ct_smem); }
        B200_LAUNCH(k_contours, n, kCtThreads, ct_smem, st, d_mask, g, ct_pw, ct_npw, d_surv, d_nsurv, h->max_surv, d_surv2, d_nsurv2, d_desc, d_ncont, d_npts,
                    d_pts, h->d_err);
    }
    dim3 gm((w + 127) / 128, (hh + 7) / 8, n);
    if (!ct_shared) B200_LAUNCH(k_probe_a, gm, blk, 0, st, d_mask, g, d_surv, d_nsurv, h->max_surv, h->d_err);
    if (!ct_shared) {
        dim3 g1(16, n);
        // small batches (latency): the ring form, whose chains of dependent loads are half as long; large batches (throughput): the end-to-end walkers
        const bool walk = !(h->ring_slots > 0 && base + n <= h->ring_slots && (h->ring_mode == 1 || n <= kRingFrames));
        // persistent CTAs; frames are the fast grid index.  Every step is a dependent load, so the kernels are latency bound; their CTAs hold their SM slots
        // for the whole kernel, so next to the extractor's dense kernels about 3 per frame is best at 256 frames (B200_PROBE_CTAS overrides)
        static const int env_gb = [] { const char* e = getenv("B200_PROBE_CTAS"); return e ? atoi(e) : 0; }();
        // Checkpointed emission (B200_EMIT_CKPT=1): k_emit 0.345 -> 0.173 ms and the detector alone 2.24 -> 2.09 ms per 256 frames, but the C3 step, where the
        // detector shares the SMs with the extractor, goes from 3.84 to 3.89-3.93 ms (measured twice, 20 steps each): off by default.
        static const bool no_ckpt = getenv("B200_EMIT_CKPT") == nullptr;
        static const int walk_steps = [] { const char* e = getenv("B200_WALK_STEPS"); const int v = e ? atoi(e) : kWalkSteps; return v > 0 ? v : kWalkSteps; }();
        dim3 gb(n, env_gb > 0 ? env_gb : std::max(1, std::min(64, (148 * 6) / n)));
        // the ring form packs a node into 64 bits (22-bit indices and keys): larger frames than 4 M mask bytes take the end-to-end walkers
        if (walk) {
            B200_LAUNCH(k_probe_b1, g1, 256, 0, st, d_mask, g, d_surv, d_nsurv, h->max_surv, d_surv2, d_nsurv2, h->d_err, (int*)nullptr, (uint32_t*)nullptr, 0);
            B200_LAUNCH(k_probe_b, gb, 128, 0, st, d_mask, g, d_surv2, d_nsurv2, h->max_surv, d_nfetch, d_desc, d_ncont, d_npts, h->d_err, walk_steps, h->d_nfetch2 + base, h->d_ckpool + (size_t)base * (h->ck_cap + h->ck_cap / kCkPerBorder), no_ckpt ? 0 : h->ck_cap);
            dim3 ge(kEmitCtasA + 4, n);
            B200_LAUNCH(k_emit, ge, 128, 0, st, d_mask, g, d_desc, d_ncont, d_pts, h->d_nfetch2 + base, h->d_ckpool + (size_t)base * (h->ck_cap + h->ck_cap / kCkPerBorder), no_ckpt ? 0 : h->ck_cap);
        } else {
            // d_surv (the transition list) is dead once phase B1 has run: it becomes the nodes' first-point-slot array
            int* d_smap = h->d_smap + (size_t)base * g.bframe;
            unsigned long long* d_nodes = h->d_nodes + (size_t)base * h->max_surv;
            uint32_t* d_sbits = h->d_sbits + (size_t)base * h->sbits_words;
            B200_LAUNCH(k_probe_b1, g1, 256, 0, st, d_mask, g, d_surv, d_nsurv, h->max_surv, d_surv2, d_nsurv2, h->d_err, d_smap, d_sbits, h->sbits_words);
            B200_LAUNCH(k_seg, gb, 128, 0, st, d_mask, g, d_surv2, d_nsurv2, h->max_surv, d_sbits, h->sbits_words, d_nodes, d_surv, h->d_err);
            B200_LAUNCH(k_link, g1, 256, 0, st, d_nsurv2, h->max_surv, d_smap, g.bframe, d_nodes);
            B200_LAUNCH(k_ring, gb, 128, 0, st, d_mask, g, d_surv2, d_nsurv2, h->max_surv, d_nodes, d_surv, d_desc, d_ncont, d_npts, h->d_err);
            B200_LAUNCH(k_emit2, gb, 128, 0, st, d_mask, g, d_surv2, d_nsurv2, h->max_surv, d_smap, d_sbits, h->sbits_words, d_nodes, d_surv, d_pts);
        }
    }
    // contour counts are only known on the device: size the per-contour grids for the capacity and let idle threads exit
    {
        dim3 gq(48, n);
        B200_LAUNCH(k_quads, gq, kQuadWarps * 32, 0, st, g, d_desc, d_ncont, d_pts, d_cand, d_ncand, h->d_err);
    }
    B200_LAUNCH(k_prefilter, n, 256, 0, st, g, d_cand, d_ncand, d_kept, d_nkept);
    dim3 gd(16, n);
    uint8_t* d_wpatch = h->d_wpatch + (size_t)base * kMaxCand * kMaxWarp * kMaxWarp;
    uint16_t* d_whist = h->d_whist + (size_t)base * kMaxCand * 256;
    int* d_wlevel = h->d_wlevel + (size_t)base * kMaxCand;
    B200_CUDA(cudaStreamWaitEvent(st, h->ev_pyr_join, 0));
    B200_LAUNCH(k_decode<0>, gd, 128, 0, st, imgs, rs, fs, d_pyr, g, d_kept, d_nkept, h->d_codes, d_dec, d_wpatch, d_whist, d_wlevel);
    B200_LAUNCH(k_otsu, dim3(kMaxCand / 128, n), 128, 0, st, g, d_nkept, d_whist, d_wlevel);
    B200_LAUNCH(k_decode<1>, gd, 128, 0, st, imgs, rs, fs, d_pyr, g, d_kept, d_nkept, h->d_codes, d_dec, d_wpatch, d_whist, d_wlevel);
    B200_LAUNCH(k_finalize, n, (n <= 32 ? kFinWarps : 8) * 32, 0, st, g, d_kept, d_nkept, d_dec, d_desc, d_pts, d_scratch,
                markers, counts, kMaxMarkers, h->d_err, h->d_mcontour + (size_t)base * kMaxMarkers);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

int b200_aruco_detect(b200_aruco_t h, const uint8_t* imgs, int n, int w, int hh, int64_t rs, int64_t fs,
                      b200_marker* markers, int32_t* counts, void* stream) {
    return b200_aruco_detect_range(h, imgs, n, w, hh, rs, fs, markers, counts, 0, stream);
}

// Blocks until the handle's work on `stream` (NULL = own stream) is done and reports scratch overflows of the
// device-pointer calls since the last check (B200_ECAPACITY), clearing the flag.
int b200_aruco_check(b200_aruco_t h, void* stream) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    int err = 0;
    B200_CUDA(cudaMemcpyAsync(&err, h->d_err, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if (err) {
        B200_CUDA(cudaMemsetAsync(h->d_err, 0, 4, st));
        static const char* what[] = {"", "", "", "border longer than the point budget", "too many contours", "contour point budget", "more than 256 quads", "more than 64 markers", "transition survivor list"};
        return fail(B200_ECAPACITY, "detector scratch overflow: %s", what[err < 9 ? err : 0]);
    }
    return B200_OK;
}

// validation taps of the LAST call: out4 = {borders longer than 70 points, convex quads, candidates after the prefilter, markers};
// corners [cap][8] / ids [cap] = the prefiltered candidates in order with their decoded id (-1: not a marker)
int b200_aruco_debug(b200_aruco_t h, int frame, int32_t* out4, float* corners, int32_t* ids, int cap) {
    if (!h || !out4) return fail(B200_EINVAL, "null %s", "argument");
    if (frame < 0 || frame >= h->max_batch) return fail(B200_EINVAL, "no such %s", "frame");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    B200_CUDA(cudaDeviceSynchronize());
    const int B = h->max_batch;
    B200_CUDA(cudaMemcpy(&out4[0], h->d_ncont + frame, 4, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(&out4[1], h->d_ncand + frame, 4, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(&out4[2], h->d_nkept + frame, 4, cudaMemcpyDeviceToHost));
    (void)B;
    const int nk = std::min(out4[2], kMaxCand);
    std::vector<Kept> k(std::max(nk, 1)); std::vector<Decoded> d(std::max(nk, 1));
    if (nk) {
        B200_CUDA(cudaMemcpy(k.data(), h->d_kept + (size_t)frame * kMaxCand, sizeof(Kept) * nk, cudaMemcpyDeviceToHost));
        B200_CUDA(cudaMemcpy(d.data(), h->d_dec + (size_t)frame * kMaxCand, sizeof(Decoded) * nk, cudaMemcpyDeviceToHost));
    }
    int nm = 0;
    for (int i = 0; i < nk; i++) {
        if (d[i].id >= 0) nm++;
        if (corners && ids && i < cap) { memcpy(corners + 8 * i, k[i].c, 32); ids[i] = d[i].id; }
    }
    out4[3] = nm;
    return B200_OK;
}

// aruco::Marker::contourPoints (Thirdparty/aruco/aruco/marker.h:59; the detector copies the candidate's border into the marker at
// markerdetector_impl.cpp:6759-6772): the border of output marker `index` of frame slot `frame` of the LAST detect call, in the reference's
// point order (Suzuki border following from the canonical start).  xy [cap][2] int32; returns the number of points of the border.
int b200_aruco_get_contour(b200_aruco_t h, int frame, int index, int32_t* xy, int cap) {
    if (!h || (!xy && cap > 0)) return fail(B200_EINVAL, "null %s", "argument");
    if (frame < 0 || frame >= h->max_batch || index < 0 || index >= kMaxMarkers) return fail(B200_EINVAL, "no such %s", "frame / marker");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    B200_CUDA(cudaStreamSynchronize(h->stream));
    int ci = -1;
    B200_CUDA(cudaMemcpy(&ci, h->d_mcontour + (size_t)frame * kMaxMarkers + index, 4, cudaMemcpyDeviceToHost));
    if (ci < 0 || ci >= h->geom.max_contours) return fail(B200_EINVAL, "no such %s", "marker in the last call");
    ContourDesc cd;
    B200_CUDA(cudaMemcpy(&cd, h->d_desc + (size_t)frame * h->geom.max_contours + ci, sizeof(cd), cudaMemcpyDeviceToHost));
    if (cd.len <= 0 || cd.off < 0 || (long long)cd.off + cd.len > h->geom.max_points) return fail(B200_EINVAL, "stale %s", "contour");
    const int m = std::min(cd.len, cap);
    if (m > 0) {
        std::vector<short2> p(m);
        B200_CUDA(cudaMemcpy(p.data(), h->d_pts + (size_t)frame * h->geom.max_points + cd.off, sizeof(short2) * m, cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; i++) { xy[2 * i] = p[i].x; xy[2 * i + 1] = p[i].y; }
    }
    return cd.len;
}

// All contours of a frame's markers in one round trip (the adapter's detect() fills every Marker::contourPoints): k_pack_contours gathers the
// borders of the n_markers output markers into one packed block, then two copies bring offsets and points home.
// ofs [n_markers + 1] (HOST): marker m owns xy [ofs[m] .. ofs[m + 1]) of xy [xy_cap][2] int32.  Returns the total number of points.
int b200_aruco_get_contours(b200_aruco_t h, int frame, int n_markers, int32_t* ofs, int32_t* xy, int xy_cap) {
    if (!h || !ofs || (!xy && xy_cap > 0)) return fail(B200_EINVAL, "null %s", "argument");
    if (frame < 0 || frame >= h->max_batch || n_markers < 0 || n_markers > kMaxMarkers) return fail(B200_EINVAL, "no such %s", "frame / marker count");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    ofs[0] = 0;
    if (n_markers == 0) return 0;
    cudaStream_t st = h->stream;
    if (!h->d_pack) {
        B200_CUDA(cudaMalloc((void**)&h->d_pack, sizeof(int) * (kMaxMarkers + 1) + sizeof(short2) * (size_t)h->geom.max_points));
    }
    int* d_ofs = reinterpret_cast<int*>(h->d_pack);
    short2* d_pp = reinterpret_cast<short2*>(d_ofs + kMaxMarkers + 1);
    B200_LAUNCH(k_pack_contours, 1, 256, 0, st, h->geom, h->d_mcontour + (size_t)frame * kMaxMarkers, n_markers, (const int*)nullptr,
                h->d_desc + (size_t)frame * h->geom.max_contours, h->d_pts + (size_t)frame * h->geom.max_points, d_ofs, d_pp, h->geom.max_points);
    int hofs[kMaxMarkers + 1];
    B200_CUDA(cudaMemcpyAsync(hofs, d_ofs, sizeof(int) * (n_markers + 1), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    const int total = hofs[n_markers];
    if (total < 0 || total > h->geom.max_points) return fail(B200_EINVAL, "stale %s", "contours");
    for (int m = 0; m <= n_markers; m++) ofs[m] = hofs[m];
    const int take = std::min(total, xy_cap);
    if (take > 0) {
        std::vector<short2> p(take);
        B200_CUDA(cudaMemcpyAsync(p.data(), d_pp, sizeof(short2) * take, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < take; i++) { xy[2 * i] = p[i].x; xy[2 * i + 1] = p[i].y; }
    }
    return total;
}

int b200_aruco_detect_host(b200_aruco_t h, const uint8_t* imgs, int n, int w, int hh, int64_t rs, int64_t fs,
                           b200_marker* markers, int32_t* counts) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    if (n < 0 || w < 0 || hh < 0) return fail(B200_EINVAL, "negative %s", "size");
    if (n > h->max_batch || w > h->max_w || hh > h->max_h) return fail(B200_ECAPACITY, "batch/image larger than the handle's %s", "capacity");
    if (n == 0) return B200_OK;
    if (!markers || !counts || (!imgs && w > 0 && hh > 0)) return fail(B200_EINVAL, "null %s", "pointer");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    if (w < 8 || hh < 8) { for (int i = 0; i < n; i++) counts[i] = 0; return B200_OK; }
    const size_t fb = (size_t)w * hh;
    if ((rc = ensure_buf(h->d_in, h->cap_in, fb * n))) return rc;
    if (h->cap_out < (size_t)n) {
        cudaFree(h->d_out); cudaFree(h->d_counts); h->d_out = nullptr; h->d_counts = nullptr; h->cap_out = 0;
        B200_CUDA(cudaMalloc((void**)&h->d_out, sizeof(b200_marker) * kMaxMarkers * (size_t)h->max_batch));
        B200_CUDA(cudaMalloc((void**)&h->d_counts, 4 * (size_t)h->max_batch));
        h->cap_out = h->max_batch;
    }
    cudaStream_t st = h->stream;
    if (fs == rs * hh) B200_CUDA(cudaMemcpy2DAsync(h->d_in, w, imgs, rs, w, (size_t)hh * n, cudaMemcpyHostToDevice, st));
    else for (int f = 0; f < n; f++) B200_CUDA(cudaMemcpy2DAsync(h->d_in + f * fb, w, imgs + (size_t)f * fs, rs, w, hh, cudaMemcpyHostToDevice, st));
    if ((rc = b200_aruco_detect(h, h->d_in, n, w, hh, w, (int64_t)fb, h->d_out, h->d_counts, st))) return rc;
    B200_CUDA(cudaMemcpyAsync(markers, h->d_out, sizeof(b200_marker) * kMaxMarkers * (size_t)n, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(counts, h->d_counts, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    int err = 0;
    B200_CUDA(cudaMemcpyAsync(&err, h->d_err, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if (err) {
        cudaMemset(h->d_err, 0, 4);
        static const char* what[] = {"", "", "", "border longer than the point budget", "too many contours", "contour point budget", "more than 256 quads", "more than 64 markers", "transition survivor list"};
        return fail(B200_ECAPACITY, "detector scratch overflow: %s", what[err < 9 ? err : 0]);
    }
    return B200_OK;
}


// One frame, everything aruco::MarkerDetector::detect(image, camParams, markerSize) returns, with ONE synchronisation: markers, the IPPE poses
// (cam9 != NULL: marker.cpp:322-343 on the device, straight from the detector's device output) and the markers' contour points
// (contour_ofs != NULL: markerdetector_impl.cpp:6759-6772).  Returns the total number of contour points (0 without contours), which may exceed
// contour_cap (then only contour_cap points were copied: fetch again with b200_aruco_get_contours), or a negative error code.
int b200_aruco_detect_frame_host(b200_aruco_t h, const uint8_t* img, int w, int hh, int64_t rs, b200_marker* markers, int32_t* count,
                                 float marker_size, const float* cam9, b200_marker_pose* poses,
                                 int32_t* contour_ofs, int32_t* contour_xy, int contour_cap) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    if (w < 0 || hh < 0 || contour_cap < 0) return fail(B200_EINVAL, "negative %s", "size");
    if (w > h->max_w || hh > h->max_h) return fail(B200_ECAPACITY, "image larger than the handle's %s", "capacity");
    if (!markers || !count || (!img && w > 0 && hh > 0) || (cam9 && !poses) || (contour_ofs && !contour_xy && contour_cap > 0)) return fail(B200_EINVAL, "null %s", "pointer");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    *count = 0;
    if (contour_ofs) contour_ofs[0] = 0;
    if (w < 8 || hh < 8) return 0;
    const size_t fb = (size_t)w * hh;
    if ((rc = ensure_buf(h->d_in, h->cap_in, fb))) return rc;
    if (h->cap_out < 1) {
        cudaFree(h->d_out); cudaFree(h->d_counts); h->d_out = nullptr; h->d_counts = nullptr; h->cap_out = 0;
        B200_CUDA(cudaMalloc((void**)&h->d_out, sizeof(b200_marker) * kMaxMarkers * (size_t)h->max_batch));
        B200_CUDA(cudaMalloc((void**)&h->d_counts, 4 * (size_t)h->max_batch));
        h->cap_out = h->max_batch;
    }
    if (!h->d_pack) B200_CUDA(cudaMalloc((void**)&h->d_pack, sizeof(int) * (kMaxMarkers + 1) + sizeof(short2) * (size_t)h->geom.max_points));
    if (!h->d_poses) B200_CUDA(cudaMalloc((void**)&h->d_poses, sizeof(b200_marker_pose) * kMaxMarkers));
    cudaStream_t st = h->stream;
    B200_CUDA(cudaMemcpy2DAsync(h->d_in, w, img, rs, w, (size_t)hh, cudaMemcpyHostToDevice, st));
    if ((rc = b200_aruco_detect(h, h->d_in, 1, w, hh, w, (int64_t)fb, h->d_out, h->d_counts, st))) return rc;
    B200_CUDA(cudaMemcpyAsync(markers, h->d_out, sizeof(b200_marker) * kMaxMarkers, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(count, h->d_counts, 4, cudaMemcpyDeviceToHost, st));
    int err = 0;
    B200_CUDA(cudaMemcpyAsync(&err, h->d_err, 4, cudaMemcpyDeviceToHost, st));
    if (cam9) {
        if ((rc = b200_aruco_pose(h->d_out, h->d_counts, 1, kMaxMarkers, marker_size, cam9, (b200_marker_pose*)h->d_poses, h->device, st))) return rc;
        B200_CUDA(cudaMemcpyAsync(poses, h->d_poses, sizeof(b200_marker_pose) * kMaxMarkers, cudaMemcpyDeviceToHost, st));
    }
    int hofs[kMaxMarkers + 1];
    const int take_cap = std::min(contour_cap, h->geom.max_points);
    if (contour_ofs) {
        int* d_ofs = reinterpret_cast<int*>(h->d_pack);
        short2* d_pp = reinterpret_cast<short2*>(d_ofs + kMaxMarkers + 1);
        B200_LAUNCH(k_pack_contours, 1, 256, 0, st, h->geom, h->d_mcontour, 0, (const int*)h->d_counts, h->d_desc, h->d_pts, d_ofs, d_pp, take_cap);
        B200_CUDA(cudaMemcpyAsync(hofs, d_ofs, sizeof(int) * (kMaxMarkers + 1), cudaMemcpyDeviceToHost, st));
        if (take_cap > 0) {
            if (h->h_pack_cap < (size_t)take_cap) {
                if (h->h_pack) cudaFreeHost(h->h_pack);
                h->h_pack = nullptr; h->h_pack_cap = 0;
                B200_CUDA(cudaMallocHost((void**)&h->h_pack, sizeof(short2) * (size_t)take_cap));
                h->h_pack_cap = take_cap;
            }
            // the number of points is only known on the device: a fixed, small upper bound travels (a marker's border has a few hundred points)
            B200_CUDA(cudaMemcpyAsync(h->h_pack, d_pp, sizeof(short2) * (size_t)std::min(take_cap, 8192), cudaMemcpyDeviceToHost, st));
        }
    }
    B200_CUDA(cudaStreamSynchronize(st));
    if (err) {
        cudaMemset(h->d_err, 0, 4);
        static const char* what[] = {"", "", "", "border longer than the point budget", "too many contours", "contour point budget", "more than 256 quads", "more than 64 markers", "transition survivor list"};
        return fail(B200_ECAPACITY, "detector scratch overflow: %s", what[err < 9 ? err : 0]);
    }
    int total = 0;
    if (contour_ofs) {
        const int nm = std::min(*count, kMaxMarkers);
        total = hofs[nm];
        if (total < 0 || total > h->geom.max_points) return fail(B200_EINVAL, "stale %s", "contours");
        for (int m = 0; m <= nm; m++) contour_ofs[m] = hofs[m];
        const int have = std::min(std::min(total, take_cap), 8192);
        const short2* p = reinterpret_cast<const short2*>(h->h_pack);
        for (int i = 0; i < have; i++) { contour_xy[2 * i] = p[i].x; contour_xy[2 * i + 1] = p[i].y; }
        if (total > have && have < std::min(total, contour_cap)) {            // more points than the fixed transfer carried: the rest with a second copy
            const int more = std::min(total, take_cap) - have;
            std::vector<short2> q(more);
            const short2* d_pp = reinterpret_cast<const short2*>(reinterpret_cast<int*>(h->d_pack) + kMaxMarkers + 1);
            B200_CUDA(cudaMemcpyAsync(q.data(), d_pp + have, sizeof(short2) * (size_t)more, cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            for (int i = 0; i < more; i++) { contour_xy[2 * (have + i)] = q[i].x; contour_xy[2 * (have + i) + 1] = q[i].y; }
        }
    }
    return total;
}

}  // extern "C"
