// orb.cu -- B200-native ORB extractor (sm_100a): image pyramid, per-cell FAST-9/16 with NMS and threshold
// fallback, quadtree keypoint distribution, intensity-centroid orientation, 7x7 Gaussian, 256-bit rBRIEF.
//
// Behavioural contract: ORB_SLAM2::ORBextractor (reference src/ORBextractor.cc:410-1132); results are
// bit-identical to the CPU oracle (oracle/orb_oracle.cpp == the reference source on the cv shim).
// All arithmetic is integer except three float32 spots that use explicitly rounded intrinsics
// (no FMA contraction): fastAtan2, the rBRIEF rotation and the final pt*scale.
//
// Pipeline for a batch of n frames (everything stays in HBM/L2, 10 launches):
//   k_pyramid_tma x (nlevels-1)   level l from level l-1, fixed-point bilinear, source tile by TMA (ORBextractor.cc:1107-1132)
//   k_fast                    one CTA per (cell, frame): u8 tile in smem by TMA, compass pretest, packed scoring of the survivors, 3x3 NMS,
//                             ini/min threshold fallback, raster-ordered compaction (ORBextractor.cc:765-829); two more forms of this stage
//                             (dense score map + per-cell NMS, one warp per cell) sit behind switches with their measured numbers
//   k_quadtree                one warp per (frame, level): DistributeOctTree (ORBextractor.cc:539-763)
//   k_describe                one warp per keypoint: IC_Angle, 7x7 Gaussian on a 43x43 patch, rBRIEF
//                             (ORBextractor.cc:77-147,1085-1101)
#include "common.h"
#include <chrono>
#include <cuda.h>          // CUtensorMap and its enums only; cuTensorMapEncodeTiled is fetched through the runtime (no -lcuda)
#include <math.h>
#include <algorithm>

namespace b200 {

constexpr int kMaxLevels = 16;
constexpr int kEdge = 19;          // EDGE_THRESHOLD, ORBextractor.cc:73
constexpr int kMinBorder = 16;     // EDGE_THRESHOLD-3, ORBextractor.cc:771
constexpr int kHalfPatch = 15;
constexpr int kFrontendChunk = 128;
constexpr int kSplitMinFrames = 64;  // b200_orb_extract: batches of at least this many frames run as two halves on two streams  // frames per upload/compute/download pipeline stage of b200_frontend_host
constexpr int kMaxRoi = 72;        // largest FAST cell ROI side handled (cells are 30..59 px + 6)

struct LevelGeom {
    int w, h;                 // level image size
    int pitch;                // row pitch in the pyramid buffer (level > 0)
    long long offset;         // byte offset of the level inside a frame's pyramid block (level > 0)
    int quota;                // mnFeaturesPerLevel
    int ncols, nrows, wcell, hcell;
    int cell_base, ncells;    // cells of this level inside the frame's cell list
    long long slot_base;      // first candidate slot of this level (entries) inside the frame's slot block
    int nini; float hx;       // DistributeOctTree initial nodes
    int xsplit;               // two initial nodes (16:9 frames): smallest x with (int)(x / hx) >= 1
    int kp_cap, kp_base;      // result capacity of this level, base inside the frame's level-result block
    int xtab, ytab;           // offsets into the resize tables
    int pbox_w, pbox_h;       // TMA box of k_pyramid_tma over the SOURCE level (l - 1): largest source region one 128 x 64 output tile reads
    int box_w, box_h;         // TMA box of k_fast: (largest cell ROI of the level + 15 columns of alignment slack) rounded up to 16 x largest ROI height
    float scale;              // mvScaleFactor[l]
    float size;               // keypoint size = (int)(31*scale)
    long long soff; int spitch;   // the level inside a frame's score map (k_fast_score -> k_fast_nms): byte offset and row pitch
};

struct OrbGeom {
    int nlevels;
    int total_cells;
    long long slots_per_frame;
    long long pyr_frame_stride;
    int res_per_frame;        // sum of kp_cap
    int ini_th, min_th;
    int store_th;             // scores below min(iniThFAST, minThFAST) are stored as 0 in the score map
    long long score_frame_stride;
    LevelGeom L[kMaxLevels];
};

struct CellDesc {             // one FAST call of the reference (ORBextractor.cc:789-829)
    short level, pad;
    short x0, y0;             // ROI origin in level coordinates
    short rw, rh;             // ROI size (interior = ROI minus 3 px on every side)
    short sx, sy;             // shift added to ROI-relative keypoints: j*wCell, i*hCell
    int slot;                 // first candidate slot (relative to the frame's slot block)
    int cap;                  // slot capacity
};

struct ResizeEntry { int ofs; short c0, c1; };

__device__ signed char g_pattern[256 * 4];       // rBRIEF pairs, transposed on upload: [k(8)][coord(4)][byte i(32)] (lane-indexed => global, not constant)
__constant__ int c_umax[16];
__device__ uint32_t g_disc[31 * 8];               // byte masks of the radius-15 disc: row v + 15, word j <-> u = -15 + 4j .. -12 + 4j (IC_Angle)

// ------------------------------------------------------------------------------------------------
// K1: pyramid level from the previous level.  cv::resize INTER_LINEAR u8 fixed point (SURVEY A-1).
// One thread -> 4 horizontally adjacent output pixels (one 32-bit store; pitch is a multiple of 16).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kPyrRows = 8;          // output rows per thread: the four column entries are fetched once and reused

__global__ void __launch_bounds__(256)
k_pyramid(const uint8_t* __restrict__ src0, long long src_row_stride, long long src_frame_stride,
          uint8_t* __restrict__ dst0, int dst_pitch, long long dst_frame_stride,
          int sw, int sh, int dw, int dh,
          const ResizeEntry* __restrict__ xt, const ResizeEntry* __restrict__ yt) {
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int f = blockIdx.z;
    if (x4 >= dw) return;
    int o0[4], o1[4], c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const ResizeEntry xe = xt[min(x4 + i, dw - 1)];          // columns past the level width land in the row padding
        o0[i] = xe.ofs; o1[i] = min(xe.ofs + 1, sw - 1); c0[i] = xe.c0; c1[i] = xe.c1;
    }
    const uint8_t* sf = src0 + (long long)f * src_frame_stride;
    uint8_t* df = dst0 + (long long)f * dst_frame_stride + x4;
#pragma unroll 2
    for (int j = 0; j < kPyrRows; j++) {
        const int y = (blockIdx.y * kPyrRows + j) * 8 + threadIdx.y;
        if (y >= dh) break;
        const ResizeEntry ye = yt[y];
        const uint8_t* r0 = sf + (long long)ye.ofs * src_row_stride;
        const uint8_t* r1 = sf + (long long)min(ye.ofs + 1, sh - 1) * src_row_stride;
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int h0 = r0[o0[i]] * c0[i] + r0[o1[i]] * c1[i];
            const int h1 = r1[o0[i]] * c0[i] + r1[o1[i]] * c1[i];
            const int v = (((ye.c0 * (h0 >> 4)) >> 16) + ((ye.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
            packed |= (uint32_t)v << (8 * i);
        }
        *reinterpret_cast<uint32_t*>(df + (long long)y * dst_pitch) = packed;     // pitch >= align16(dw): the tail bytes are padding
    }
}

// Same arithmetic, source tile staged by TMA: the 128 x 64 output tile of a CTA needs a source region of about 156 x 79 pixels, fetched
// as ONE box (cp.async.bulk.tensor.3d) whose x starts at the first source column rounded down to 16 (u8 tensor maps only take 16-byte
// aligned inner coordinates); the four taps of every output pixel then come from shared memory.
__global__ void __launch_bounds__(256)
k_pyramid_tma(const __grid_constant__ CUtensorMap smap, int zbase, int box_w, int box_h,
              uint8_t* __restrict__ dst0, int dst_pitch, long long dst_frame_stride, int sw, int sh, int dw, int dh,
              const ResizeEntry* __restrict__ xt, const ResizeEntry* __restrict__ yt) {
    extern __shared__ __align__(1024) unsigned char pt_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * (8 * kPyrRows);
    const int xs0 = xt[x0].ofs & ~15, ys0 = yt[y0].ofs;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&s_bar)), "r"(box_w * box_h) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(pt_raw)), "l"(reinterpret_cast<uint64_t>(&smap)), "r"(xs0), "r"(ys0), "r"(zbase + f), "r"(smem_u32(&s_bar)) : "memory");
    }
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    int o0[4], o1[4], c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const ResizeEntry xe = xt[min(x4 + i, dw - 1)];          // columns past the level width land in the row padding
        o0[i] = xe.ofs - xs0; o1[i] = min(xe.ofs + 1, sw - 1) - xs0; c0[i] = xe.c0; c1[i] = xe.c1;
    }
    __syncthreads();                                             // the barrier word is initialised before anybody polls it
    asm volatile("{\n.reg .pred p;\nPYR_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra PYR_DONE;\nbra PYR_WAIT;\nPYR_DONE:\n}"
                 :: "r"(smem_u32(&s_bar)) : "memory");
    if (x4 >= dw) return;
    uint8_t* df = dst0 + (long long)f * dst_frame_stride + x4;
#pragma unroll 2
    for (int j = 0; j < kPyrRows; j++) {
        const int y = y0 + j * 8 + threadIdx.y;
        if (y >= dh) break;
        const ResizeEntry ye = yt[y];
        const uint8_t* r0 = pt_raw + (ye.ofs - ys0) * box_w;
        const uint8_t* r1 = pt_raw + (min(ye.ofs + 1, sh - 1) - ys0) * box_w;
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int h0 = r0[o0[i]] * c0[i] + r0[o1[i]] * c1[i];
            const int h1 = r1[o0[i]] * c0[i] + r1[o1[i]] * c1[i];
            const int v = (((ye.c0 * (h0 >> 4)) >> 16) + ((ye.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
            packed |= (uint32_t)v << (8 * i);
        }
        *reinterpret_cast<uint32_t*>(df + (long long)y * dst_pitch) = packed;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: FAST-9/16 score, NMS, threshold fallback, ordered compaction.  One CTA per (cell, frame).
// score(p) = max over the 16 arcs of 9 contiguous circle pixels of max(min d, -max d) - 1, d = I(circle) - I(p)
// (the largest threshold for which p is still a corner; SURVEY A-3).
//
// The cell's ROI arrives in shared memory as ONE TMA box (cp.async.bulk.tensor.3d: x, y, frame; per-level tensor maps;
// out-of-image bytes read as zero and are never used).  The box starts at the ROI's x rounded down to 16: a u8 tensor map
// only accepts 16-byte aligned inner coordinates (anything else raises an illegal-instruction fault, tools/tma_probe.cu).  All arithmetic runs in packed s16x2 registers, the one SIMD width sm_100a executes natively (VIADD.16x2,
// VIMNMX[3].S16x2; the u8x4 video instructions are emulated with 5-8 LOP3/IADD each).
//
// Per threshold T (iniThFAST, then minThFAST only when the cell has no keypoint after NMS, ORBextractor.cc:809-816):
//   1. pretest, one thread per aligned 4-pixel group: the five words (centre, N, S and the two row neighbours) are split
//      into even / odd pixel pairs (PRMT).  A 9-arc always contains two ADJACENT compass points (circle positions 0,4,8,12),
//      one of N/S and one of E/W, so a corner needs min(max(N,S), max(E,W)) > v+T or max(min(N,S), min(E,W)) < v-T:
//      12 packed instructions per pixel pair.  Survivors go to the warp's private queue (no atomics, no CTA barrier)
//   2. every thread scores TWO queued pixels at once, one per s16x2 half: 2 x 40 min3/max3 give all 16 arcs of both
//      polarities; scores >= T are written to the byte score tile and the pixel goes to the warp's corner queue
//   3. 3x3 NMS (strict >, neighbours outside the cell interior count 0) over the corner queues -> per-row bit masks
// then a raster-ordered compaction of the surviving pixels into the cell's candidate slots.
// ------------------------------------------------------------------------------------------------
constexpr int kFastMinWarps = 2;           // k_fast runs with 2, 3 or 4 warps per cell; the queues are sized for 2

struct FastSmemGeom {                        // dynamic shared memory carve-up, sized by the host for the largest box in use
    int tile_bytes;                          // max over levels of box_w * box_h, rounded up to 128
    int qcap;                                // pixel-queue entries per warp
    int keepw;                               // keep words
};

struct FastMaps { CUtensorMap m[kMaxLevels]; };     // level 0: the caller's frames; level l > 0: that level inside the pyramid buffer

__device__ __forceinline__ unsigned pair_e(unsigned w) { return __byte_perm(w, 0, 0x4240); }      // pixels 0, 2 of a word as u16x2
__device__ __forceinline__ unsigned pair_o(unsigned w) { return __byte_perm(w, 0, 0x4341); }      // pixels 1, 3
__device__ __forceinline__ unsigned sh16(unsigned lo, unsigned hi) { return __byte_perm(lo, hi, 0x5432); }   // (lo.hi16, hi.lo16)

// compass pretest of one pixel pair: non-zero halves are pixels that can still be a corner at threshold T
__device__ __forceinline__ unsigned fast_pretest_pair(unsigned V, unsigned N, unsigned S, unsigned E, unsigned W, unsigned Tp, unsigned Tn) {
    const unsigned a = __vmins2(__vmaxs2(N, S), __vmaxs2(E, W)), b = __vmaxs2(__vmins2(N, S), __vmins2(E, W));
    const unsigned hi = __vadd2(V, Tp), lo = __vadd2(V, Tn);
    return (a ^ __vmins2(a, hi)) | (b ^ __vmaxs2(b, lo));
}

// all 16 arcs of 9 for two pixels: d[j] = circle_j - centre (s16x2), returns max(maxmin, -minmax) - 1 per half
__device__ __forceinline__ unsigned fast_score_pair(const unsigned (&d)[16]) {
    unsigned t1[16], u1[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        t1[j] = __vimin3_s16x2(d[j], d[(j + 1) & 15], d[(j + 2) & 15]);
        u1[j] = __vimax3_s16x2(d[j], d[(j + 1) & 15], d[(j + 2) & 15]);
    }
    unsigned t2[16], u2[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        t2[j] = __vimin3_s16x2(t1[j], t1[(j + 3) & 15], t1[(j + 6) & 15]);       // min of d over positions j..j+8
        u2[j] = __vimax3_s16x2(u1[j], u1[(j + 3) & 15], u1[(j + 6) & 15]);
    }
    unsigned a = __vimax3_s16x2(t2[0], t2[1], t2[2]), b = __vimax3_s16x2(t2[3], t2[4], t2[5]), c = __vimax3_s16x2(t2[6], t2[7], t2[8]);
    unsigned e = __vimax3_s16x2(t2[9], t2[10], t2[11]), f = __vimax3_s16x2(t2[12], t2[13], t2[14]);
    a = __vimax3_s16x2(a, b, c); e = __vimax3_s16x2(e, f, t2[15]);
    const unsigned maxmin = __vmaxs2(a, e);
    a = __vimin3_s16x2(u2[0], u2[1], u2[2]); b = __vimin3_s16x2(u2[3], u2[4], u2[5]); c = __vimin3_s16x2(u2[6], u2[7], u2[8]);
    e = __vimin3_s16x2(u2[9], u2[10], u2[11]); f = __vimin3_s16x2(u2[12], u2[13], u2[14]);
    a = __vimin3_s16x2(a, b, c); e = __vimin3_s16x2(e, f, u2[15]);
    const unsigned minmax = __vmins2(a, e);
    return __vadd2(__vmaxs2(maxmin, __vneg2(minmax)), 0xffffffffu);
}

// TP: tile pitch in bytes = TMA box width, a compile-time constant so that every circle offset is an immediate
template <int TP, int NT>
__global__ void __launch_bounds__(NT, 1280 / NT)
k_fast(const uint8_t* __restrict__ img0, long long img_row_stride, long long img_frame_stride,
       const uint8_t* __restrict__ pyr, const __grid_constant__ OrbGeom g, const __grid_constant__ FastMaps maps,
       const FastSmemGeom sg, int tma_level0, int scratch_base,
       const CellDesc* __restrict__ cells, uint32_t* __restrict__ slots, int* __restrict__ cellcnt) {
    extern __shared__ __align__(1024) unsigned char fs_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_any;
    __shared__ int s_nc[(NT / 32)];

    const CellDesc cd = cells[blockIdx.x];
    const int f = blockIdx.y;
    const LevelGeom& lg = g.L[cd.level];
    constexpr int tp = TP;
    uint8_t* tile = fs_raw;                                                             // [box_h][tp], column = ox + ROI column
    uint8_t* score = fs_raw + sg.tile_bytes;                                            // same geometry
    uint16_t* q_all = reinterpret_cast<uint16_t*>(fs_raw + 2 * sg.tile_bytes);          // per-warp pixel queues: y << 8 | tile column
    uint32_t* keep = reinterpret_cast<uint32_t*>(q_all + (NT / 32) * sg.qcap);

    const int rw = cd.rw, rh = cd.rh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool use_tma = cd.level > 0 || tma_level0;
    const int ox = cd.x0 & 15;                        // level rows start 16-byte aligned, so this is the ROI's offset inside its aligned box
    if (use_tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&s_bar)), "r"(TP * lg.box_h) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&maps.m[cd.level])), "r"((int)cd.x0 - ox), "r"((int)cd.y0),
                            "r"(cd.level > 0 ? scratch_base + f : f), "r"(smem_u32(&s_bar)) : "memory");
        }
    } else {
        // caller frames whose pointer / strides are not 16-byte multiples cannot be described by a tensor map: byte-wise staging
        const uint8_t* base = img0 + (long long)f * img_frame_stride + (long long)cd.y0 * img_row_stride + cd.x0;
        for (int y = warp; y < rh; y += (NT / 32))
            for (int x = lane; x < rw; x += 32) tile[y * tp + ox + x] = base[(long long)y * img_row_stride + x];
    }
    {   // meanwhile: clear the score tile
        uint4* sc4 = reinterpret_cast<uint4*>(score);
        for (int i = tid; i < (rh * tp) >> 4; i += NT) sc4[i] = make_uint4(0u, 0u, 0u, 0u);      // tp is a multiple of 16
    }
    const int iw = rw - 6, ih = rh - 6;               // interior
    const int tc0 = ox + 3, tc1 = tc0 + iw;           // interior tile columns [tc0, tc1)
    const int g_lo = tc0 >> 2, ngroups = ((tc1 - 1) >> 2) - g_lo + 1;      // aligned 4-pixel groups that hold interior pixels
    // pretest lane mapping: lane -> (row inside the warp's row block, group); fixed for the whole kernel
    const int rpi = 32 / min(ngroups, 32);            // rows per warp iteration
    const int my_sub = lane / ngroups, my_g = g_lo + lane - my_sub * ngroups;
    const bool my_on = my_sub < rpi;
    // validity of the four pixels of my group (tile columns 4g .. 4g+3 inside [tc0, tc1)) as half-word masks
    unsigned valid_e = 0, valid_o = 0;
    {
        const int c = 4 * my_g;
        if (c >= tc0 && c < tc1) valid_e |= 0x0000ffffu;
        if (c + 2 >= tc0 && c + 2 < tc1) valid_e |= 0xffff0000u;
        if (c + 1 >= tc0 && c + 1 < tc1) valid_o |= 0x0000ffffu;
        if (c + 3 >= tc0 && c + 3 < tc1) valid_o |= 0xffff0000u;
    }
    uint16_t* q = q_all + warp * sg.qcap;
    if (use_tma) {
        __syncthreads();                              // the barrier word is initialised before anybody polls it
        asm volatile("{\n.reg .pred p;\nFAST_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra FAST_DONE;\nbra FAST_WAIT;\nFAST_DONE:\n}"
                     :: "r"(smem_u32(&s_bar)) : "memory");
    } else __syncthreads();

    int T = g.ini_th;
    for (int pass = 0; pass < 2; pass++) {
        for (int i = tid; i < ih * 3; i += NT) keep[i] = 0u;
        if (tid == 0) s_any = 0;
        const unsigned Tp = (unsigned)T * 0x10001u, Tn = (unsigned)(0x10000 - T) * 0x10001u;
        // 1. pretest + enqueue (warp-private queue)
        int nq = 0;
        for (int y0 = warp * rpi; y0 < ih; y0 += (NT / 32) * rpi) {
            const int y = y0 + my_sub;
            unsigned me = 0, mo = 0;
            if (my_on && y < ih) {
                const uint32_t* r = reinterpret_cast<const uint32_t*>(tile + (y + 3) * tp) + my_g;
                const int rp = tp >> 2;
                const uint32_t wl = r[-1], wc = r[0], wr = r[1], wn = r[-3 * rp], ws = r[3 * rp];
                const unsigned ce = pair_e(wc), co = pair_o(wc), le = pair_e(wl), lo = pair_o(wl), re = pair_e(wr), ro = pair_o(wr);
                // even pixels (0, 2): west = odd pair of the left word, east = (pixel 3, pixel 5)
                me = fast_pretest_pair(ce, pair_e(wn), pair_e(ws), sh16(co, ro), lo, Tp, Tn) & valid_e;
                // odd pixels (1, 3): west = (pixel -2, pixel 0), east = even pair of the right word
                mo = fast_pretest_pair(co, pair_o(wn), pair_o(ws), re, sh16(le, ce), Tp, Tn) & valid_o;
            }
            const bool p0 = (me & 0xffffu) != 0, p1 = (mo & 0xffffu) != 0, p2 = (me >> 16) != 0, p3 = (mo >> 16) != 0;
            const unsigned b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
            const unsigned b2 = __ballot_sync(0xffffffffu, p2), b3 = __ballot_sync(0xffffffffu, p3);
            if (b0 | b1 | b2 | b3) {            // (a scan of per-lane counts instead of four ballots measured 5 % slower)
                const unsigned lt = (1u << lane) - 1;
                const int ent = (y << 8) | (4 * my_g);
                const int o1 = nq + __popc(b0), o2 = o1 + __popc(b1), o3 = o2 + __popc(b2);
                if (p0) q[nq + __popc(b0 & lt)] = (uint16_t)ent;
                if (p1) q[o1 + __popc(b1 & lt)] = (uint16_t)(ent + 1);
                if (p2) q[o2 + __popc(b2 & lt)] = (uint16_t)(ent + 2);
                if (p3) q[o3 + __popc(b3 & lt)] = (uint16_t)(ent + 3);
                nq = o3 + __popc(b3);
            }
        }
        __syncwarp();
        // 2. scores, two queued pixels per thread (the second one of an odd tail repeats the first)
        int nc = 0;
        for (int i0 = 0; i0 < nq; i0 += 64) {
            const int i = i0 + 2 * lane;
            const bool on = i < nq;
            int s0 = 0, s1 = 0, ea = 0, eb = 0;
            if (on) {
                ea = q[i]; eb = i + 1 < nq ? q[i + 1] : ea;
                const uint8_t* pa = tile + ((ea >> 8) + 3) * tp + (ea & 255);
                const uint8_t* pb = tile + ((eb >> 8) + 3) * tp + (eb & 255);
                // circle: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
                unsigned d[16];
#define B200_LD2(j, off) d[j] = (unsigned)pa[off] | ((unsigned)pb[off] << 16)
                B200_LD2(0, 3 * tp); B200_LD2(1, 3 * tp + 1); B200_LD2(2, 2 * tp + 2); B200_LD2(3, tp + 3);
                B200_LD2(4, 3); B200_LD2(5, -tp + 3); B200_LD2(6, -2 * tp + 2); B200_LD2(7, -3 * tp + 1);
                B200_LD2(8, -3 * tp); B200_LD2(9, -3 * tp - 1); B200_LD2(10, -2 * tp - 2); B200_LD2(11, -tp - 3);
                B200_LD2(12, -3); B200_LD2(13, tp - 3); B200_LD2(14, 2 * tp - 2); B200_LD2(15, 3 * tp - 1);
#undef B200_LD2
                const unsigned nV = __vneg2((unsigned)pa[0] | ((unsigned)pb[0] << 16));
#pragma unroll
                for (int j = 0; j < 16; j++) d[j] = __vadd2(d[j], nV);
                const unsigned sc2 = fast_score_pair(d);
                s0 = (int)(short)(sc2 & 0xffffu); s1 = (int)sc2 >> 16;
                if (i + 1 >= nq) s1 = 0;
            }
            const bool k0 = s0 >= T, k1 = s1 >= T;
            if (k0) score[((ea >> 8) + 3) * tp + (ea & 255)] = (uint8_t)s0;
            if (k1) score[((eb >> 8) + 3) * tp + (eb & 255)] = (uint8_t)s1;
            // corners overwrite the head of the queue (entries below i0 + 64 have been consumed; nc <= i0 + 64 always)
            const unsigned c0 = __ballot_sync(0xffffffffu, k0), c1 = __ballot_sync(0xffffffffu, k1);
            const unsigned lt = (1u << lane) - 1;
            const int o1 = nc + __popc(c0);
            __syncwarp();
            if (k0) q[nc + __popc(c0 & lt)] = (uint16_t)ea;
            if (k1) q[o1 + __popc(c1 & lt)] = (uint16_t)eb;
            nc = o1 + __popc(c1);
            __syncwarp();
        }
        if (lane == 0) s_nc[warp] = nc;
        __syncthreads();
        // 3. NMS of the corners -> keep bits (bit index = interior column); the warps share all four corner lists
        for (int w = 0; w < (NT / 32); w++) {
            const int ncw = s_nc[w];
            const uint16_t* qc = q_all + w * sg.qcap;
            for (int i = tid; i < ncw; i += NT) {
                const int e = qc[i], y = e >> 8, c = e & 255;
                const uint8_t* sp = score + (y + 3) * tp + c;
                const int sc = sp[0];
                if (sc > sp[-1] && sc > sp[1] && sc > sp[-tp - 1] && sc > sp[-tp] && sc > sp[-tp + 1] && sc > sp[tp - 1] && sc > sp[tp] && sc > sp[tp + 1]) {
                    const int x = c - tc0;
                    atomicOr(&keep[y * 3 + (x >> 5)], 1u << (x & 31));
                    s_any = 1;
                }
            }
        }
        __syncthreads();
        const int any = s_any;
        __syncthreads();                              // everybody has read the flag before thread 0 clears it for the next pass
        if (any || pass == 1 || g.min_th >= g.ini_th) break;
        // nothing at iniThFAST: the whole cell again at minThFAST (a superset of the pixels examined so far; the scores already
        // in the tile stay valid)
        T = g.min_th;
    }
    // raster-ordered emission: keep words in (row, chunk) order; warp 0 scans their popcounts
    if (warp == 0) {
        uint32_t* out = slots + (long long)f * g.slots_per_frame + cd.slot;
        const int nkw = ih * 3;
        int run = 0;
        for (int w0 = 0; w0 < nkw; w0 += 32) {
            const int wi = w0 + lane;
            uint32_t m = wi < nkw ? keep[wi] : 0u;
            const int c = __popc(m);
            int sc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
            int pos = run + sc - c;
            const int y = wi / 3, ch = wi - y * 3;
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int x = ch * 32 + bit;
                // key = x | y << 12 | score << 24, coordinates relative to the 16-px border like the reference's vToDistributeKeys
                if (pos < cd.cap) out[pos] = (uint32_t)(x + 3 + cd.sx) | ((uint32_t)(y + 3 + cd.sy) << 12) | ((uint32_t)score[(y + 3) * tp + tc0 + x] << 24);
                pos++;
            }
            run += __shfl_sync(0xffffffffu, sc, 31);
        }
        if (lane == 0) cellcnt[(long long)f * g.total_cells + blockIdx.x] = run;
    }
}

// ------------------------------------------------------------------------------------------------
// K2, warp-per-cell form (B200_FAST_WARP=1).  The same three stages as k_fast above with the same arithmetic, but ONE warp owns a cell from the TMA box to the
// candidate slots, so nothing in it is a CTA barrier, an atomic on a shared flag or a loop that four warps repeat: the per-line ncu profile of k_fast
// (profiles/r2n_fast_lines.txt) spends 4800 warp instructions per cell, about 1000 of them in prologues and clears every warp executes and as many in
// ballots and queue bookkeeping per pretest row.  Here
//   * the pretest keeps its verdicts in a 64-bit register mask per lane (4 bits per row the lane visits) and the survivors are compacted once per 16
//     rows-of-lanes: one warp scan of the popcounts, then every lane writes its own entries
//   * survivor queue and corner list have fixed sizes (kFwQ, kFwC): survivors beyond the queue are scored in further rounds of the same compaction; a
//     cell with more than kFwC corners falls back to an NMS sweep over the whole score tile
//   * NMS hits are an OR into the row's keep words, and the raster-ordered emission is the warp's own scan.
// A CTA is kFwWarps independent warps (cells blockIdx.x * kFwWarps ...), each with its own mbarrier and its own slice of dynamic shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kFwWarps = 4;
constexpr int kFwQ = 512;                    // survivor queue entries per warp
constexpr int kFwC = 256;                    // corner list entries per warp

template <int TP>
__global__ void __launch_bounds__(kFwWarps * 32)
k_fastw(const uint8_t* __restrict__ img0, long long img_row_stride, long long img_frame_stride,
        const uint8_t* __restrict__ pyr, const __grid_constant__ OrbGeom g, const __grid_constant__ FastMaps maps,
        int tile_bytes, int keepw, int warp_bytes, int tma_level0, int scratch_base,
        const CellDesc* __restrict__ cells, uint32_t* __restrict__ slots, int* __restrict__ cellcnt) {
    extern __shared__ __align__(1024) unsigned char fs_raw[];
    __shared__ __align__(8) unsigned long long s_bar[kFwWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ci = blockIdx.x * kFwWarps + warp;
    if (ci >= g.total_cells) return;
    const CellDesc cd = cells[ci];
    const int f = blockIdx.y;
    const LevelGeom& lg = g.L[cd.level];
    constexpr int tp = TP;
    uint8_t* tile = fs_raw + (size_t)warp * warp_bytes;                                 // [box_h][tp], column = ox + ROI column
    uint8_t* score = tile + tile_bytes;                                                 // same geometry
    uint16_t* q = reinterpret_cast<uint16_t*>(score + tile_bytes);                      // survivors: y << 8 | tile column
    uint16_t* cq = q + kFwQ;                                                            // corners
    uint32_t* keep = reinterpret_cast<uint32_t*>(cq + kFwC);                            // [ih][3] bit index = interior column

    const int rw = cd.rw, rh = cd.rh;
    const bool use_tma = cd.level > 0 || tma_level0;
    const int ox = cd.x0 & 15;
    const uint32_t bar = smem_u32(&s_bar[warp]);
    if (use_tma) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(TP * lg.box_h) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&maps.m[cd.level])), "r"((int)cd.x0 - ox), "r"((int)cd.y0),
                            "r"(cd.level > 0 ? scratch_base + f : f), "r"(bar) : "memory");
        }
    } else {
        const uint8_t* base = img0 + (long long)f * img_frame_stride + (long long)cd.y0 * img_row_stride + cd.x0;
        for (int y = 0; y < rh; y++)
            for (int x = lane; x < rw; x += 32) tile[y * tp + ox + x] = base[(long long)y * img_row_stride + x];
    }
    {   // meanwhile: clear the score tile and the keep words
        uint4* sc4 = reinterpret_cast<uint4*>(score);
        for (int i = lane; i < (rh * tp) >> 4; i += 32) sc4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const int iw = rw - 6, ih = rh - 6;               // interior
    const int tc0 = ox + 3, tc1 = tc0 + iw;           // interior tile columns [tc0, tc1)
    const int g_lo = tc0 >> 2, ngroups = ((tc1 - 1) >> 2) - g_lo + 1;      // aligned 4-pixel groups that hold interior pixels (<= 18)
    const int rpi = 32 / ngroups;                     // rows per warp iteration
    const int my_sub = lane / ngroups, my_g = g_lo + lane - my_sub * ngroups;
    const bool my_on = my_sub < rpi;
    unsigned valid_e = 0, valid_o = 0;
    {
        const int c = 4 * my_g;
        if (c >= tc0 && c < tc1) valid_e |= 0x0000ffffu;
        if (c + 2 >= tc0 && c + 2 < tc1) valid_e |= 0xffff0000u;
        if (c + 1 >= tc0 && c + 1 < tc1) valid_o |= 0x0000ffffu;
        if (c + 3 >= tc0 && c + 3 < tc1) valid_o |= 0xffff0000u;
    }
    const unsigned lt = (1u << lane) - 1;
    __syncwarp();
    if (use_tma)
        asm volatile("{\n.reg .pred p;\nFASTW_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra FASTW_DONE;\nbra FASTW_WAIT;\nFASTW_DONE:\n}"
                     :: "r"(bar) : "memory");

    int T = g.ini_th;
    for (int pass = 0; pass < 2; pass++) {
        for (int i = lane; i < ih * 3; i += 32) keep[i] = 0u;
        const unsigned Tp = (unsigned)T * 0x10001u, Tn = (unsigned)(0x10000 - T) * 0x10001u;
        int nc = 0;                                   // corners (score >= T) found so far; > kFwC = the list overflowed
        for (int yb = 0; yb < ih; yb += 16 * rpi) {   // block of up to 16 warp iterations = 64 verdict bits per lane
            // 1. pretest
            unsigned mlo = 0, mhi = 0;
            const uint32_t* r = reinterpret_cast<const uint32_t*>(tile + (yb + my_sub + 3) * tp) + my_g;
            constexpr int rp = tp >> 2;
#pragma unroll 2
            for (int it = 0; it < 16; it++) {
                const int y = yb + it * rpi + my_sub;
                if (yb + it * rpi >= ih) break;
                if (my_on && y < ih) {
                    const uint32_t* rr = r + it * rpi * rp;
                    const uint32_t wl = rr[-1], wc = rr[0], wr = rr[1], wn = rr[-3 * rp], ws = rr[3 * rp];
                    const unsigned ce = pair_e(wc), co = pair_o(wc), le = pair_e(wl), lo = pair_o(wl), re = pair_e(wr), ro = pair_o(wr);
                    const unsigned me = fast_pretest_pair(ce, pair_e(wn), pair_e(ws), sh16(co, ro), lo, Tp, Tn) & valid_e;
                    const unsigned mo = fast_pretest_pair(co, pair_o(wn), pair_o(ws), re, sh16(le, ce), Tp, Tn) & valid_o;
                    const unsigned bits = ((me & 0xffffu) ? 1u : 0u) | ((mo & 0xffffu) ? 2u : 0u) | ((me >> 16) ? 4u : 0u) | ((mo >> 16) ? 8u : 0u);
                    if (it < 8) mlo |= bits << (4 * it); else mhi |= bits << (4 * it - 32);
                }
            }
            // 2. compaction + scoring in rounds of kFwQ survivors
            const int cnt = __popc(mlo) + __popc(mhi);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            for (int r0 = 0; r0 < total; r0 += kFwQ) {
                {
                    int rank = incl - cnt - r0;
                    unsigned a = mlo, b = mhi;
                    const int ent0 = ((yb + my_sub) << 8) | (4 * my_g);
                    while (a | b) {
                        int bit;
                        if (a) { bit = __ffs(a) - 1; a &= a - 1; } else { bit = 32 + __ffs(b) - 1; b &= b - 1; }
                        if (rank >= 0 && rank < kFwQ) q[rank] = (uint16_t)(ent0 + (((bit >> 2) * rpi) << 8) + (bit & 3));
                        rank++;
                    }
                }
                __syncwarp();
                const int nq = min(kFwQ, total - r0);
                for (int i0 = 0; i0 < nq; i0 += 64) {
                    const int i = i0 + 2 * lane;
                    int s0 = 0, s1 = 0, ea = 0, eb = 0;
                    if (i < nq) {
                        const unsigned e2 = *reinterpret_cast<const unsigned*>(q + i);
                        ea = e2 & 0xffffu; eb = i + 1 < nq ? (int)(e2 >> 16) : ea;
                        const uint8_t* pa = tile + ((ea >> 8) + 3) * tp + (ea & 255);
                        const uint8_t* pb = tile + ((eb >> 8) + 3) * tp + (eb & 255);
                        unsigned d[16];
#define B200_LD2(j, off) d[j] = (unsigned)pa[off] | ((unsigned)pb[off] << 16)
                        B200_LD2(0, 3 * tp); B200_LD2(1, 3 * tp + 1); B200_LD2(2, 2 * tp + 2); B200_LD2(3, tp + 3);
                        B200_LD2(4, 3); B200_LD2(5, -tp + 3); B200_LD2(6, -2 * tp + 2); B200_LD2(7, -3 * tp + 1);
                        B200_LD2(8, -3 * tp); B200_LD2(9, -3 * tp - 1); B200_LD2(10, -2 * tp - 2); B200_LD2(11, -tp - 3);
                        B200_LD2(12, -3); B200_LD2(13, tp - 3); B200_LD2(14, 2 * tp - 2); B200_LD2(15, 3 * tp - 1);
#undef B200_LD2
                        const unsigned nV = __vneg2((unsigned)pa[0] | ((unsigned)pb[0] << 16));
#pragma unroll
                        for (int j = 0; j < 16; j++) d[j] = __vadd2(d[j], nV);
                        const unsigned sc2 = fast_score_pair(d);
                        s0 = (int)(short)(sc2 & 0xffffu); s1 = (int)sc2 >> 16;
                        if (i + 1 >= nq) s1 = 0;
                    }
                    const bool k0 = s0 >= T, k1 = s1 >= T;
                    if (k0) score[((ea >> 8) + 3) * tp + (ea & 255)] = (uint8_t)s0;
                    if (k1) score[((eb >> 8) + 3) * tp + (eb & 255)] = (uint8_t)s1;
                    const unsigned c0 = __ballot_sync(0xffffffffu, k0), c1 = __ballot_sync(0xffffffffu, k1);
                    if (c0 | c1) {
                        const int p0 = nc + __popc(c0 & lt), p1 = nc + __popc(c0) + __popc(c1 & lt);
                        if (k0 && p0 < kFwC) cq[p0] = (uint16_t)ea;
                        if (k1 && p1 < kFwC) cq[p1] = (uint16_t)eb;
                        nc += __popc(c0) + __popc(c1);
                    }
                }
                __syncwarp();
            }
        }
        // 3. NMS -> keep bits
        __syncwarp();
        bool any = false;
        if (nc <= kFwC) {
            for (int i = lane; i < nc; i += 32) {
                const int e = cq[i], y = e >> 8, c = e & 255;
                const uint8_t* sp = score + (y + 3) * tp + c;
                const int sc = sp[0];
                if (sc > sp[-1] && sc > sp[1] && sc > sp[-tp - 1] && sc > sp[-tp] && sc > sp[-tp + 1] && sc > sp[tp - 1] && sc > sp[tp] && sc > sp[tp + 1]) {
                    const int x = c - tc0;
                    atomicOr(&keep[y * 3 + (x >> 5)], 1u << (x & 31));
                    any = true;
                }
            }
        } else {
            for (int i = lane; i < iw * ih; i += 32) {
                const int y = i / iw, x = i - y * iw;
                const uint8_t* sp = score + (y + 3) * tp + tc0 + x;
                const int sc = sp[0];
                if (sc >= T && sc > sp[-1] && sc > sp[1] && sc > sp[-tp - 1] && sc > sp[-tp] && sc > sp[-tp + 1] && sc > sp[tp - 1] && sc > sp[tp] && sc > sp[tp + 1]) {
                    atomicOr(&keep[y * 3 + (x >> 5)], 1u << (x & 31));
                    any = true;
                }
            }
        }
        any = __any_sync(0xffffffffu, any);
        if (any || pass == 1 || g.min_th >= g.ini_th) break;
        T = g.min_th;                                 // nothing at iniThFAST: the whole cell again at minThFAST (scores already in the tile stay valid)
    }
    __syncwarp();
    {   // raster-ordered emission: keep words in (row, chunk) order
        uint32_t* out = slots + (long long)f * g.slots_per_frame + cd.slot;
        const int nkw = ih * 3;
        int run = 0;
        for (int w0 = 0; w0 < nkw; w0 += 32) {
            const int wi = w0 + lane;
            uint32_t m = wi < nkw ? keep[wi] : 0u;
            const int c = __popc(m);
            int sc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
            int pos = run + sc - c;
            const int y = wi / 3, ch = wi - y * 3;
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int x = ch * 32 + bit;
                if (pos < cd.cap) out[pos] = (uint32_t)(x + 3 + cd.sx) | ((uint32_t)(y + 3 + cd.sy) << 12) | ((uint32_t)score[(y + 3) * tp + tc0 + x] << 24);
                pos++;
            }
            run += __shfl_sync(0xffffffffu, sc, 31);
        }
        if (lane == 0) cellcnt[(long long)f * g.total_cells + ci] = run;
    }
}

// ------------------------------------------------------------------------------------------------
// K2 (dense form, B200_FAST_DENSE=1): the FAST score is threshold independent and the cells' interiors tile the level, so the work splits
// into a perfectly regular part and a tiny per-cell part.
//
// k_fast_score: the score of EVERY pixel of a 64 x 32 tile, no pretest, no queues, no divergence.  The (64 + 6) x (32 + 6) pixel
// region arrives as one TMA box (96 x 38 bytes: the box starts at the tile's x rounded down to 16, which puts the first pixel at byte
// 13 and every 4-pixel group on a word).  A thread owns an aligned group of four pixels = two pixel pairs in u16x2 registers; per
// row it reads 3 words from each of the 7 rows it touches and takes the 17 circle / centre values of both pairs out of them with one
// PRMT each (a zero register supplies the high bytes; the four straddling positions need one AND more).  Subtracting the centre
// commutes with min / max, so the arcs run on the raw pixel values: 2 x (16 + 16) min3 / max3 for "min over 9 contiguous" of both
// polarities, 2 x 8 for the outer max / min, then score = max(maxmin - v, v - minmax) - 1.  Scores below minThFAST become 0; the
// four bytes of a thread are one coalesced 32-bit store into the level's score map.  Lanes 16-31 work two rows below lanes 0-15:
// with the 96-byte tile pitch that shifts their words by 16 banks, so no shared-memory load has a conflict.
//
// k_fast_nms: one warp per cell (ORBextractor.cc:789-829).  The cell interior's scores are staged into shared memory with a zero ring
// (neighbours outside the interior count 0, as cv::FAST on the cell ROI treats them).  A pixel with score >= T can never lose against a
// neighbour below T, so NMS is "strictly greater than the 8 raw neighbours" for every threshold; every lane scans a contiguous raster run
// of words, tests only the bytes >= T (iniThFAST first, minThFAST only when the cell stays empty), and a warp prefix turns the per-lane
// keep masks into the raster-ordered candidate slots the quadtree reads.
// ------------------------------------------------------------------------------------------------
constexpr int kScTW = 64, kScTH = 32;                 // score tile (pixels)
constexpr int kScBoxW = 96, kScBoxH = kScTH + 6;      // TMA box: 13 bytes of alignment slack + 64 + 2 x 3 halo -> 83, rounded up to 96
constexpr int kScThreads = 128;

struct ScoreTile { short level, tx, ty, pad; };       // tile origin in level coordinates: x = 16 + 64 tx, y = 16 + 32 ty

// pixel pair at column offset DX from the thread's word: PAIR 0 = pixels 0, 1 of the group, PAIR 1 = pixels 2, 3; (wm, w0, wp) = the words
// left of / at / right of the group in that row.  Result: the two pixel values zero-extended to u16x2.
template <int PAIR, int DX>
__device__ __forceinline__ unsigned sc_take(unsigned wm, unsigned w0, unsigned wp) {
    constexpr int o = 2 * PAIR + DX;                  // byte offset of the pair's first pixel relative to w0's byte 0: -3 .. 5
    if (o == -3) return __byte_perm(wm, 0u, 0x4241);
    if (o == -2) return __byte_perm(wm, 0u, 0x4342);
    if (o == -1) return __byte_perm(wm, w0, 0x0403) & 0x00ff00ffu;
    if (o == 0) return __byte_perm(w0, 0u, 0x4140);
    if (o == 1) return __byte_perm(w0, 0u, 0x4241);
    if (o == 2) return __byte_perm(w0, 0u, 0x4342);
    if (o == 3) return __byte_perm(w0, wp, 0x0403) & 0x00ff00ffu;
    if (o == 4) return __byte_perm(wp, 0u, 0x4140);
    return __byte_perm(wp, 0u, 0x4241);               // o == 5
}

// 0xffff in every half word of x that is negative (prmt's sign-replicate mode on bytes 1 and 3), else 0
__device__ __forceinline__ unsigned sign_spread16x2(unsigned x) {
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0u), "r"(0xbb99u));
    return r;
}

// score of one pixel pair from its 16 circle values (raw u16x2) and the centre; s16x2, may be negative for flat neighbourhoods
__device__ __forceinline__ unsigned sc_score_pair(const unsigned (&c)[16], unsigned v) {
    unsigned t1[16], u1[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        t1[j] = __vimin3_s16x2(c[j], c[(j + 1) & 15], c[(j + 2) & 15]);
        u1[j] = __vimax3_s16x2(c[j], c[(j + 1) & 15], c[(j + 2) & 15]);
    }
    unsigned t2[16], u2[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        t2[j] = __vimin3_s16x2(t1[j], t1[(j + 3) & 15], t1[(j + 6) & 15]);       // min over positions j .. j + 8
        u2[j] = __vimax3_s16x2(u1[j], u1[(j + 3) & 15], u1[(j + 6) & 15]);
    }
    unsigned a = __vimax3_s16x2(t2[0], t2[1], t2[2]), b = __vimax3_s16x2(t2[3], t2[4], t2[5]), d = __vimax3_s16x2(t2[6], t2[7], t2[8]);
    unsigned e = __vimax3_s16x2(t2[9], t2[10], t2[11]), f = __vimax3_s16x2(t2[12], t2[13], t2[14]);
    a = __vimax3_s16x2(a, b, d); e = __vimax3_s16x2(e, f, t2[15]);
    const unsigned maxmin = __vmaxs2(a, e);
    a = __vimin3_s16x2(u2[0], u2[1], u2[2]); b = __vimin3_s16x2(u2[3], u2[4], u2[5]); d = __vimin3_s16x2(u2[6], u2[7], u2[8]);
    e = __vimin3_s16x2(u2[9], u2[10], u2[11]); f = __vimin3_s16x2(u2[12], u2[13], u2[14]);
    a = __vimin3_s16x2(a, b, d); e = __vimin3_s16x2(e, f, u2[15]);
    const unsigned minmax = __vmins2(a, e);
    // max(maxmin - v, v - minmax) - 1
    return __vadd2(__vmaxs2(__vsub2(maxmin, v), __vsub2(v, minmax)), 0xffffffffu);
}

template <int PAIR>
__device__ __forceinline__ unsigned sc_pair(const unsigned (&wm)[7], const unsigned (&w0)[7], const unsigned (&wp)[7]) {
    // rows: index 0 .. 6 = dy -3 .. +3.  circle: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
    unsigned c[16];
    c[0] = sc_take<PAIR, 0>(wm[6], w0[6], wp[6]);   c[1] = sc_take<PAIR, 1>(wm[6], w0[6], wp[6]);   c[2] = sc_take<PAIR, 2>(wm[5], w0[5], wp[5]);
    c[3] = sc_take<PAIR, 3>(wm[4], w0[4], wp[4]);   c[4] = sc_take<PAIR, 3>(wm[3], w0[3], wp[3]);   c[5] = sc_take<PAIR, 3>(wm[2], w0[2], wp[2]);
    c[6] = sc_take<PAIR, 2>(wm[1], w0[1], wp[1]);   c[7] = sc_take<PAIR, 1>(wm[0], w0[0], wp[0]);   c[8] = sc_take<PAIR, 0>(wm[0], w0[0], wp[0]);
    c[9] = sc_take<PAIR, -1>(wm[0], w0[0], wp[0]);  c[10] = sc_take<PAIR, -2>(wm[1], w0[1], wp[1]); c[11] = sc_take<PAIR, -3>(wm[2], w0[2], wp[2]);
    c[12] = sc_take<PAIR, -3>(wm[3], w0[3], wp[3]); c[13] = sc_take<PAIR, -3>(wm[4], w0[4], wp[4]); c[14] = sc_take<PAIR, -2>(wm[5], w0[5], wp[5]);
    c[15] = sc_take<PAIR, -1>(wm[6], w0[6], wp[6]);
    return sc_score_pair(c, sc_take<PAIR, 0>(wm[3], w0[3], wp[3]));
}

__global__ void __launch_bounds__(kScThreads)
k_fast_score(const uint8_t* __restrict__ img0, long long img_row_stride, long long img_frame_stride, const uint8_t* __restrict__ pyr,
             const __grid_constant__ OrbGeom g, const __grid_constant__ FastMaps maps, int tma_level0, int scratch_base,
             const ScoreTile* __restrict__ tiles, uint8_t* __restrict__ score0) {
    __shared__ __align__(128) unsigned char tile[kScBoxW * kScBoxH];
    __shared__ __align__(8) unsigned long long s_bar;
    const ScoreTile td = tiles[blockIdx.x];
    const int f = blockIdx.y;
    const LevelGeom& lg = g.L[td.level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx0 = kScTW * td.tx, by0 = 13 + kScTH * td.ty;          // box origin; pixel (16 + 64 tx + i, 16 + 32 ty + j) sits at tile (16 + i, 3 + j)
    const bool use_tma = td.level > 0 || tma_level0;
    if (use_tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&s_bar)), "r"(kScBoxW * kScBoxH) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&maps.m[td.level])), "r"(bx0), "r"(by0),
                            "r"(td.level > 0 ? scratch_base + f : f), "r"(smem_u32(&s_bar)) : "memory");
        }
        __syncthreads();                              // the barrier word is initialised before anybody polls it
        asm volatile("{\n.reg .pred p;\nSC_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra SC_DONE;\nbra SC_WAIT;\nSC_DONE:\n}"
                     :: "r"(smem_u32(&s_bar)) : "memory");
    } else {
        // caller frames whose pointer / strides are not 16-byte multiples cannot be described by a tensor map: byte-wise staging, zero outside
        const uint8_t* base = img0 + (long long)f * img_frame_stride;
        for (int i = tid; i < kScBoxW * kScBoxH; i += kScThreads) {
            const int ty = i / kScBoxW, tx = i - ty * kScBoxW, x = bx0 + tx, y = by0 + ty;
            tile[i] = (x < lg.w && y < lg.h) ? base[(long long)y * img_row_stride + x] : (unsigned char)0;
        }
        __syncthreads();
    }
    const int half = lane >> 4, q = lane & 15;
    const int X = 16 + kScTW * td.tx + 4 * q;                         // level x of the group's first pixel
    uint8_t* out = score0 + (long long)f * g.score_frame_stride + lg.soff;
    const unsigned minp = (unsigned)g.store_th * 0x10001u;
    // warp w: rows 8w + {0, 1, 4, 5} for lanes 0-15 and the rows two below ({2, 3, 6, 7}) for lanes 16-31
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
        const int r = 8 * warp + (it & 1) + 4 * (it >> 1) + 2 * half;
        const int Y = 16 + kScTH * td.ty + r;
        if (Y >= lg.h - 16 || X >= lg.w - 16) continue;               // warp-divergent only in edge tiles
        const uint32_t* row = reinterpret_cast<const uint32_t*>(tile + r * kScBoxW) + 4 + q;       // row dy = -3 of the group's word
        unsigned wm[7], w0[7], wp[7];
#pragma unroll
        for (int k = 0; k < 7; k++) { wm[k] = row[k * (kScBoxW / 4) - 1]; w0[k] = row[k * (kScBoxW / 4)]; wp[k] = row[k * (kScBoxW / 4) + 1]; }
        unsigned sa = sc_pair<0>(wm, w0, wp), sb = sc_pair<1>(wm, w0, wp);
        // scores below the stored threshold (and negative ones) become 0: the sign of (score - th) spread over the half word
        const unsigned da = __vsub2(sa, minp), db = __vsub2(sb, minp);
        sa &= ~sign_spread16x2(da); sb &= ~sign_spread16x2(db);
        *reinterpret_cast<uint32_t*>(out + (long long)Y * lg.spitch + X) = __byte_perm(sa, sb, 0x6420);
    }
}

constexpr int kNmsWarps = 8;

// One warp per cell, one LANE per interior row (lanes 0 and 31 hold the rows above / below a block of 30), the row's scores in registers as aligned words split
// into even / odd pixel pairs (u16x2).  "Strictly greater than the eight neighbours" is separable: h3 = max(left, centre, right) and h2 = max(left, right) per
// row inside the lane, the rows above and below arrive with two shuffles per register, and the verdict of a pixel is the sign of
// (max(h3 above, h3 below, h2) - s) & (T - 1 - s).  Scores outside the interior are masked to zero when the words are loaded (cv::FAST on the cell ROI never
// sees them), so no ring has to be staged anywhere.  A pixel >= T never loses against a neighbour below T, which makes the test independent of the threshold:
// iniThFAST first, minThFAST only when the cell stays empty (ORBextractor.cc:809-816).  The keep bits of a row are one nibble per word; a warp prefix over
// their popcounts gives every row its first candidate slot, and the rows write their keys in raster order.
template <int NW>
__device__ __forceinline__ int nms_cell(const uint8_t* __restrict__ src, int spitch, int xi, int yi, int iw, int ih, int ini_th, int min_th, int lane,
                                        uint32_t* __restrict__ out, int cap) {
    const int xa = xi & ~3;                                           // first word
    const int lo_clear = xi - xa;                                     // bytes of word 0 left of the interior
    const int xe = xi + iw;                                           // one past the interior
    int total = 0;
    int T = ini_th;
    for (int pass = 0; pass < 2; pass++) {
        total = 0;
        const unsigned Tm1 = (unsigned)(T - 1) * 0x10001u;
        for (int rb = 0; rb < ih; rb += 30) {
            __syncwarp();                                             // (tells the compiler the lanes are converged: plain SHFL instead of a collective per shuffle)
            const int r = rb + lane - 1;                              // interior row of this lane
            const bool row_on = r >= 0 && r < ih;
            unsigned E[NW + 1], O[NW + 1];
            const uint8_t* rowp = src + (long long)(yi + r) * spitch + xa;
#pragma unroll
            for (int j = 0; j < NW; j++) {
                unsigned w = 0;
                if (row_on && xa + 4 * j < xe) {
                    w = *reinterpret_cast<const uint32_t*>(rowp + 4 * j);
                    if (j == 0 && lo_clear) w &= 0xffffffffu << (8 * lo_clear);
                    const int hi = xa + 4 * j + 4 - xe;               // bytes past the interior
                    if (hi > 0) w &= 0xffffffffu >> (8 * hi);
                }
                E[j] = __byte_perm(w, 0u, 0x4240); O[j] = __byte_perm(w, 0u, 0x4341);
            }
            E[NW] = 0u; O[NW] = 0u;
            unsigned km[NW];                                          // keep masks: bit 8 b + 7 = pixel 4 j + b of the row
            unsigned cnt = 0;
            const bool emit = lane >= 1 && lane <= 30 && row_on;
#pragma unroll
            for (int j = 0; j < NW; j++) {
                const unsigned lfE = sh16(j ? O[j - 1] : 0u, O[j]);   // pixels (4j - 1, 4j + 1): left neighbours of the even pair
                const unsigned rtO = sh16(E[j], E[j + 1]);            // pixels (4j + 2, 4j + 4): right neighbours of the odd pair
                const unsigned h3E = __vimax3_s16x2(E[j], lfE, O[j]), h3O = __vimax3_s16x2(O[j], E[j], rtO);
                const unsigned h2E = __vmaxs2(lfE, O[j]), h2O = __vmaxs2(E[j], rtO);
                const unsigned upE = __shfl_up_sync(0xffffffffu, h3E, 1), dnE = __shfl_down_sync(0xffffffffu, h3E, 1);
                const unsigned upO = __shfl_up_sync(0xffffffffu, h3O, 1), dnO = __shfl_down_sync(0xffffffffu, h3O, 1);
                // m = max(eight neighbours, T - 1); keep <=> s > m <=> s + ~m >= 0 per half word (s + ~m = s - m - 1)
                const unsigned mE = __vmaxs2(__vimax3_s16x2(upE, dnE, h2E), Tm1), mO = __vmaxs2(__vimax3_s16x2(upO, dnO, h2O), Tm1);
                unsigned tE, tO;
                asm("add.s16x2 %0, %1, %2;" : "=r"(tE) : "r"(E[j]), "r"(~mE));
                asm("add.s16x2 %0, %1, %2;" : "=r"(tO) : "r"(O[j]), "r"(~mO));
                // sign bytes in pixel order (E.lo, O.lo, E.hi, O.hi); a clear sign = keep
                km[j] = emit ? (~__byte_perm(tE, tO, 0x7351) & 0x80808080u) : 0u;
                cnt = __dp4a(km[j] >> 7, 0x01010101u, cnt);
            }
            int sc = (int)cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
            int pos = total + sc - (int)cnt;
            total += __shfl_sync(0xffffffffu, sc, 31);
            if (cnt) {
#pragma unroll
                for (int j = 0; j < NW; j++) {
                    unsigned m = km[j];
                    while (m) {
                        const int b = (__ffs(m) - 1) >> 3;
                        m &= m - 1;
                        const int x = xa + 4 * j + b, y = yi + r;    // level coordinates
                        // key = x | y << 12 | score << 24, coordinates relative to the 16-px border like the reference's vToDistributeKeys
                        if (pos < cap) out[pos] = (uint32_t)(x - 16) | ((uint32_t)(y - 16) << 12) | ((uint32_t)rowp[4 * j + b] << 24);
                        pos++;
                    }
                }
            }
        }
        if (total > 0 || pass == 1 || min_th >= ini_th) break;
        T = min_th;
    }
    return total;
}

__global__ void __launch_bounds__(kNmsWarps * 32)
k_fast_nms(const __grid_constant__ OrbGeom g, const CellDesc* __restrict__ cells, const uint8_t* __restrict__ score0,
           uint32_t* __restrict__ slots, int* __restrict__ cellcnt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cell = blockIdx.x * kNmsWarps + warp, f = blockIdx.y;
    if (cell >= g.total_cells) return;
    const CellDesc cd = cells[cell];
    const LevelGeom& lg = g.L[cd.level];
    const int iw = cd.rw - 6, ih = cd.rh - 6;
    const int xi = cd.x0 + 3, yi = cd.y0 + 3;                        // first interior pixel (level coordinates)
    const uint8_t* src = score0 + (long long)f * g.score_frame_stride + lg.soff;
    uint32_t* out = slots + (long long)f * g.slots_per_frame + cd.slot;
    const int nww = ((xi + iw + 3) >> 2) - (xi >> 2);                // aligned words that hold interior pixels
    int total;
    if (nww <= 9) total = nms_cell<9>(src, lg.spitch, xi, yi, iw, ih, g.ini_th, g.min_th, lane, out, cd.cap);
    else if (nww <= 10) total = nms_cell<10>(src, lg.spitch, xi, yi, iw, ih, g.ini_th, g.min_th, lane, out, cd.cap);
    else total = nms_cell<18>(src, lg.spitch, xi, yi, iw, ih, g.ini_th, g.min_th, lane, out, cd.cap);
    if (lane == 0) cellcnt[(long long)f * g.total_cells + cell] = total;
}

// ------------------------------------------------------------------------------------------------
// K3: DistributeOctTree.  One warp per (frame, level).  The list logic runs redundantly (warp-uniform) on all
// lanes; stable 4-way partitions, the initial gather and the final per-node arg-max are lane-parallel.
// Keys are packed u32 (x | y<<12 | score<<24) living in two global scratch buffers; a node's keys are a
// contiguous range in one of them (children are written to the other one, so no copy-back).
// Canonical tie-break of the size-sorted phase (ORBextractor.cc:684): (count, creation sequence).
// ------------------------------------------------------------------------------------------------
struct QtNode {
    short x0, x1, y0, y1;
    int beg, end;
    short prev, next;
    int seq;
    unsigned char no_more, buf, pad0, pad1;
};

constexpr int kQtWarps = 4;
constexpr int kQtSmemKeys = 4096;       // per-warp capacity of the shared-memory key buffers of k_quadtree<true> (2 x 16 KB)

struct QtWarp {          // per-warp view of its shared memory
    QtNode* pool; short* freelist; int nfree;
    int* big_cnt; int* big_seq; short* big_id;      // children with > 1 key created in the current round
    int* prv_cnt; int* prv_seq; short* prv_id;      // the previous round's list (the one being sorted)
    short* order;
    int head, tail, size, seq;
};

__device__ __forceinline__ int qt_alloc(QtWarp& w) { return w.freelist[--w.nfree]; }

__device__ __forceinline__ void qt_push_back(QtWarp& w, int id, int lane) {
    if (lane == 0) {
        w.pool[id].next = -1; w.pool[id].prev = (short)w.tail;
        if (w.tail >= 0) w.pool[w.tail].next = (short)id;
    }
    if (w.tail < 0) w.head = id;
    w.tail = id; w.size++;
    __syncwarp();
}

// DivideNode (ORBextractor.cc:481-537): stable 4-way partition of the node's key range into the other buffer.
// Returns the four child counts (warp-uniform) and fills child boxes.
// `pre` (valid when has_pre): this node's keys, one per lane, fetched while the previous node was being divided (nodes of at most 32 keys)
__device__ __forceinline__ void qt_divide(const QtNode nd, uint32_t* bufA, uint32_t* bufB, int lane,
                                          int cnt[4], int& mx, int& my, const uint32_t (&pre)[4], bool has_pre) {
    const int halfX = (nd.x1 - nd.x0 + 1) >> 1;      // ceil((float)(UR.x-UL.x)/2) for non-negative ints
    const int halfY = (nd.y1 - nd.y0 + 1) >> 1;
    mx = nd.x0 + halfX; my = nd.y0 + halfY;
    const uint32_t* src = nd.buf ? bufB : bufA;
    uint32_t* dst = nd.buf ? bufA : bufB;
    const int n = nd.end - nd.beg;
    cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
    if (n <= 128) {                                   // common case: the node's keys sit in four registers per lane, counted and scattered from there
        uint32_t key[4]; int q[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool valid = 32 * u + lane < n;
            key[u] = has_pre ? pre[u] : (valid ? src[nd.beg + 32 * u + lane] : 0u);
            q[u] = valid ? (((int)(key[u] & 0xfff) >= mx) + 2 * ((int)((key[u] >> 12) & 0xfff) >= my)) : 4;
        }
        unsigned m[4][4];                             // [u][k]
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (32 * u < n) {
#pragma unroll
                for (int k = 0; k < 4; k++) { m[u][k] = __ballot_sync(0xffffffffu, q[u] == k); cnt[k] += __popc(m[u][k]); }
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) m[u][k] = 0u;
            }
        }
        int run[4];
        run[0] = nd.beg; run[1] = run[0] + cnt[0]; run[2] = run[1] + cnt[1]; run[3] = run[2] + cnt[2];
        const unsigned lt = (1u << lane) - 1;
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (q[u] == k) dst[run[k] + __popc(m[u][k] & lt)] = key[u];
                run[k] += __popc(m[u][k]);
            }
        }
    } else {
        // both passes fetch 4 x 32 keys per round trip (the keys live in global scratch: latency, not bandwidth, is the cost)
        for (int i0 = nd.beg; i0 < nd.end; i0 += 128) {          // pass 1: counts
            uint32_t key[4]; bool valid[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const int i = i0 + 32 * u + lane; valid[u] = i < nd.end; key[u] = valid[u] ? src[i] : 0u; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int q = valid[u] ? (((int)(key[u] & 0xfff) >= mx) + 2 * ((int)((key[u] >> 12) & 0xfff) >= my)) : 4;
#pragma unroll
                for (int k = 0; k < 4; k++) cnt[k] += __popc(__ballot_sync(0xffffffffu, q == k));
            }
        }
        int run[4];
        run[0] = nd.beg; run[1] = run[0] + cnt[0]; run[2] = run[1] + cnt[1]; run[3] = run[2] + cnt[2];
        for (int i0 = nd.beg; i0 < nd.end; i0 += 128) {          // pass 2: stable scatter
            uint32_t key[4]; bool valid[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const int i = i0 + 32 * u + lane; valid[u] = i < nd.end; key[u] = valid[u] ? src[i] : 0u; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int q = valid[u] ? (((int)(key[u] & 0xfff) >= mx) + 2 * ((int)((key[u] >> 12) & 0xfff) >= my)) : 4;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned m = __ballot_sync(0xffffffffu, q == k);
                    if (q == k) dst[run[k] + __popc(m & ((1u << lane) - 1))] = key[u];
                    run[k] += __popc(m);
                }
            }
        }
    }
    __syncwarp();
}

// start fetching the keys of node `id` (if it is small enough for the register path): the load is in flight while the current node is divided
__device__ __forceinline__ bool qt_prefetch(const QtWarp& w, int id, const uint32_t* bufA, const uint32_t* bufB, int lane, uint32_t (&key)[4]) {
#pragma unroll
    for (int u = 0; u < 4; u++) key[u] = 0u;
    if (id < 0) return false;
    const QtNode nd = w.pool[id];
    const int n = nd.end - nd.beg;
    if (n > 128) return false;
    const uint32_t* src = (nd.buf ? bufB : bufA) + nd.beg;
#pragma unroll
    for (int u = 0; u < 4; u++) if (32 * u + lane < n) key[u] = src[32 * u + lane];
    return true;
}


// The non-empty children of `nd` go to the list front in the order n1..n4 (ORBextractor.cc:606-665; children with more than one key are appended to
// the `big` list) and the parent leaves the list, in one step with two warp barriers: lanes 0..3 write one child each (node record, links
// among the new children, entry of the `big` list), lane 4 re-links the old list head; then the parent is unlinked and its slot goes back to the
// free list.  Same list order as four push_fronts in the order n1..n4: [n4, n3, n2, n1, old head, ...]; same creation sequence numbers.
__device__ __forceinline__ int qt_split(QtWarp& w, int it, const QtNode nd, const int cnt[4], int mx, int my, int lane, int& nbig) {
    int re[4], rb[4], ne = 0, nb = 0;                 // rank among the non-empty / the big children
#pragma unroll
    for (int q = 0; q < 4; q++) { re[q] = ne; rb[q] = nb; ne += cnt[q] > 0; nb += cnt[q] > 1; }
    int ids[4];
#pragma unroll
    for (int q = 0; q < 4; q++) ids[q] = cnt[q] > 0 ? (int)w.freelist[w.nfree - 1 - re[q]] : -1;
    const int old_head = w.head;
    int lo = -1, hi = -1;                              // lowest / highest non-empty child
#pragma unroll
    for (int q = 3; q >= 0; q--) if (cnt[q] > 0) lo = q;
#pragma unroll
    for (int q = 0; q < 4; q++) if (cnt[q] > 0) hi = q;
    if (lane < 4) {
        const int q = lane;
        const int k = q == 0 ? cnt[0] : q == 1 ? cnt[1] : q == 2 ? cnt[2] : cnt[3];
        if (k > 0) {
            const int beg = nd.beg + (q > 0 ? cnt[0] : 0) + (q > 1 ? cnt[1] : 0) + (q > 2 ? cnt[2] : 0);
            const int myrank = q == 0 ? re[0] : q == 1 ? re[1] : q == 2 ? re[2] : re[3];
            const int id = q == 0 ? ids[0] : q == 1 ? ids[1] : q == 2 ? ids[2] : ids[3];
            int nxt = old_head, prv = -1;              // next = the non-empty child below me (or the old head), prev = the one above me
#pragma unroll
            for (int j = 0; j < 4; j++) { if (j < q && cnt[j] > 0) nxt = ids[j]; }
#pragma unroll
            for (int j = 3; j >= 0; j--) { if (j > q && cnt[j] > 0) prv = ids[j]; }
            QtNode c;
            c.x0 = (q & 1) ? (short)mx : nd.x0; c.x1 = (q & 1) ? nd.x1 : (short)mx;
            c.y0 = (q & 2) ? (short)my : nd.y0; c.y1 = (q & 2) ? nd.y1 : (short)my;
            c.beg = beg; c.end = beg + k; c.prev = (short)prv; c.next = (short)nxt; c.seq = w.seq + myrank;
            c.no_more = (k == 1); c.buf = nd.buf ^ 1; c.pad0 = c.pad1 = 0;
            w.pool[id] = c;
            if (k > 1) {
                const int bi = nbig + (q == 0 ? rb[0] : q == 1 ? rb[1] : q == 2 ? rb[2] : rb[3]);
                w.big_cnt[bi] = k; w.big_seq[bi] = c.seq; w.big_id[bi] = (short)id;
            }
        }
    } else if (lane == 4 && ne > 0 && old_head >= 0) w.pool[old_head].prev = (short)ids[lo < 0 ? 0 : lo];
    if (ne > 0) { w.head = ids[hi < 0 ? 0 : hi]; w.size += ne; w.seq += ne; w.nfree -= ne; nbig += nb; }
    __syncwarp();
    // unlink the parent
    const int p = w.pool[it].prev, n = w.pool[it].next;
    if (lane == 0) {
        if (p >= 0) w.pool[p].next = (short)n;
        if (n >= 0) w.pool[n].prev = (short)p;
        w.freelist[w.nfree] = (short)it;
    }
    if (p < 0) w.head = n;
    if (n < 0) w.tail = p;
    w.size--; w.nfree++;
    __syncwarp();
    return nb;
}

// SMEMKEYS (small batches, where the L2 round trips of the key traffic are the whole run time): a level whose keys fit `key_cap` keeps both key
// buffers in shared memory behind the node pool; larger levels use the global scratch as always.
template <bool SMEMKEYS>
__global__ void __launch_bounds__(kQtWarps * 32)
k_quadtree(const __grid_constant__ OrbGeom g, const CellDesc* __restrict__ cells,
           const uint32_t* __restrict__ slots, const int* __restrict__ cellcnt,
           uint32_t* __restrict__ keysA, uint32_t* __restrict__ keysB,
           uint32_t* __restrict__ lvlres, int* __restrict__ lvlcnt, int nframes, int pool_cap, int* __restrict__ errflag, int key_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int prob = blockIdx.x * kQtWarps + warp;
    if (prob >= nframes * g.nlevels) return;
    // heavy levels first: problem index -> (level, frame) with the level as the slow index
    const int level = prob / nframes, f = prob % nframes;
    const LevelGeom& lg = g.L[level];

    // carve this warp's shared memory
    const size_t per_warp = (size_t)pool_cap * (sizeof(QtNode) + 2 + 2 * (4 + 4 + 2) + 2) + 64 + (SMEMKEYS ? (size_t)key_cap * 8 : 0);
    unsigned char* sm = smem_raw + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    uint32_t* skeys = reinterpret_cast<uint32_t*>(sm);       // SMEMKEYS: [2][key_cap] in front of the pool
    if (SMEMKEYS) sm += (size_t)key_cap * 8;
    QtWarp w;
    w.pool = reinterpret_cast<QtNode*>(sm); sm += (size_t)pool_cap * sizeof(QtNode);
    w.big_cnt = reinterpret_cast<int*>(sm); sm += (size_t)pool_cap * 4;
    w.big_seq = reinterpret_cast<int*>(sm); sm += (size_t)pool_cap * 4;
    w.prv_cnt = reinterpret_cast<int*>(sm); sm += (size_t)pool_cap * 4;
    w.prv_seq = reinterpret_cast<int*>(sm); sm += (size_t)pool_cap * 4;
    w.freelist = reinterpret_cast<short*>(sm); sm += (size_t)pool_cap * 2;
    w.big_id = reinterpret_cast<short*>(sm); sm += (size_t)pool_cap * 2;
    w.prv_id = reinterpret_cast<short*>(sm); sm += (size_t)pool_cap * 2;
    w.order = reinterpret_cast<short*>(sm);
    w.head = w.tail = -1; w.size = 0; w.seq = 0; w.nfree = pool_cap;
    for (int i = lane; i < pool_cap; i += 32) w.freelist[i] = (short)(pool_cap - 1 - i);
    __syncwarp();

    uint32_t* A = keysA + (long long)f * g.slots_per_frame + lg.slot_base;
    uint32_t* B = keysB + (long long)f * g.slots_per_frame + lg.slot_base;
    const uint32_t* S = slots + (long long)f * g.slots_per_frame;
    const int* cc = cellcnt + (long long)f * g.total_cells + lg.cell_base;
    const CellDesc* cl = cells + lg.cell_base;
    if (SMEMKEYS) {                           // how many keys does this level hold?  (the gather below must know its destination first)
        int tk = 0;
        for (int c = lane; c < lg.ncells; c += 32) tk += min(cc[c], cl[c].cap);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tk += __shfl_xor_sync(0xffffffffu, tk, o);
        if (tk <= key_cap) { A = skeys; B = skeys + key_cap; }
    }
    const int N = lg.quota;
    int* out_cnt = lvlcnt + (long long)f * g.nlevels + level;
    uint32_t* out = lvlres + (long long)f * g.res_per_frame + lg.kp_base;

    // ---- initial nodes (ORBextractor.cc:543-570): key -> node (int)(x / hX); stable gather from the cell slots
    const int nini = lg.nini;
    int total = 0;
    auto make_node = [&](int ni, int beg, int end) {
        const int k = end - beg;
        if (k <= 0) return;                   // empty initial nodes are erased right away (ORBextractor.cc:581-582)
        const int id = qt_alloc(w);
        if (lane == 0) {
            QtNode c;
            c.x0 = (short)(int)__fmul_rn(lg.hx, (float)ni); c.x1 = (short)(int)__fmul_rn(lg.hx, (float)(ni + 1));
            c.y0 = 0; c.y1 = (short)(lg.h - 2 * kMinBorder);
            c.beg = beg; c.end = end; c.prev = c.next = -1; c.seq = w.seq; c.no_more = (k == 1); c.buf = 0; c.pad0 = c.pad1 = 0;
            w.pool[id] = c;
        }
        __syncwarp();
        qt_push_back(w, id, lane);
        w.seq++;
    };
    if (nini == 2) {
        // two initial nodes (16:9 frames): (int)(x / hX) is monotone in x, so node 1 starts at the host-computed column xsplit and no key needs a
        // division.  Pass 1 counts node 0 (node 1 begins behind it), pass 2 places both nodes' keys in cell order; the two per-cell counts travel
        // through one scan as a packed pair.  (The per-node form below read every key four times and divided each time: 38 % of this kernel's
        // stall samples at 1920 x 1080.)
        const unsigned xb = (unsigned)lg.xsplit;
        int tot0 = 0;
        for (int c0 = 0; c0 < lg.ncells; c0 += 32) {
            const int c = c0 + lane;
            int n = 0, slot = 0;
            if (c < lg.ncells) { n = min(cc[c], cl[c].cap); slot = cl[c].slot; }
            const uint32_t* src = S + slot;
            int mine0 = 0;
            for (int k0 = 0; k0 < n; k0 += 4) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = k0 + u < n ? src[k0 + u] : 0xfffu;
#pragma unroll
                for (int u = 0; u < 4; u++) mine0 += (v[u] & 0xfff) < xb;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine0 += __shfl_xor_sync(0xffffffffu, mine0, o);
            tot0 += mine0;
        }
        int run0 = 0, run1 = tot0;
        for (int c0 = 0; c0 < lg.ncells; c0 += 32) {
            const int c = c0 + lane;
            int n = 0, slot = 0;
            if (c < lg.ncells) { n = min(cc[c], cl[c].cap); slot = cl[c].slot; }
            const uint32_t* src = S + slot;
            int mine0 = 0;
            for (int k0 = 0; k0 < n; k0 += 4) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = k0 + u < n ? src[k0 + u] : 0xfffu;
#pragma unroll
                for (int u = 0; u < 4; u++) mine0 += (v[u] & 0xfff) < xb;
            }
            const unsigned mine = (unsigned)mine0 | ((unsigned)(n - mine0) << 16);      // a cell holds at most 324 keys: 32 cells fit 16 bits
            unsigned incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            uint32_t* d0 = A + run0 + (int)((incl - mine) & 0xffffu);
            uint32_t* d1 = A + run1 + (int)((incl - mine) >> 16);
            for (int k0 = 0; k0 < n; k0 += 4) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = k0 + u < n ? src[k0 + u] : 0u;
#pragma unroll
                for (int u = 0; u < 4; u++) if (k0 + u < n) { if ((v[u] & 0xfff) < xb) *d0++ = v[u]; else *d1++ = v[u]; }
            }
            const unsigned all = __shfl_sync(0xffffffffu, incl, 31);
            run0 += (int)(all & 0xffffu); run1 += (int)(all >> 16);
        }
        make_node(0, 0, tot0);
        make_node(1, tot0, run1);
        total = run1;
    } else if (nini >= 1) {
        int run = 0;
        for (int ni = 0; ni < nini; ni++) {
            const int beg = run;
            if (nini == 1) {
                // one initial node (4:3 frames): all keys in cell order.  32 cells per round: every lane owns one cell (count and slot
                // fetched in one round trip), a warp scan places the cells, then each lane copies its cell's few keys.
                for (int c0 = 0; c0 < lg.ncells; c0 += 32) {
                    const int c = c0 + lane;
                    int n = 0, slot = 0;
                    if (c < lg.ncells) { n = min(cc[c], cl[c].cap); slot = cl[c].slot; }
                    int incl = n;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                    const uint32_t* src = S + slot;
                    uint32_t* dstp = A + run + incl - n;
                    for (int k0 = 0; k0 < n; k0 += 4) {
                        uint32_t v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) v[u] = k0 + u < n ? src[k0 + u] : 0u;
#pragma unroll
                        for (int u = 0; u < 4; u++) if (k0 + u < n) dstp[k0 + u] = v[u];
                    }
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
            } else {
                // three or more initial nodes (wider than 5:2): node ni takes the keys with (int)(x / hX) == ni in cell order.  Same
                // lane-per-cell scheme, run once per node: a counting sweep is not needed because the nodes are filled one after the
                // other (run continues where the previous node ended)
                for (int c0 = 0; c0 < lg.ncells; c0 += 32) {
                    const int c = c0 + lane;
                    int n = 0, slot = 0;
                    if (c < lg.ncells) { n = min(cc[c], cl[c].cap); slot = cl[c].slot; }
                    const uint32_t* src = S + slot;
                    int mine = 0;
                    for (int k0 = 0; k0 < n; k0 += 4) {
                        uint32_t v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) v[u] = k0 + u < n ? src[k0 + u] : 0u;
#pragma unroll
                        for (int u = 0; u < 4; u++) mine += (k0 + u < n) && (int)__fdiv_rn((float)(v[u] & 0xfff), lg.hx) == ni;
                    }
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                    uint32_t* dstp = A + run + incl - mine;
                    for (int k0 = 0; k0 < n; k0 += 4) {
                        uint32_t v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) v[u] = k0 + u < n ? src[k0 + u] : 0u;
#pragma unroll
                        for (int u = 0; u < 4; u++) if ((k0 + u < n) && (int)__fdiv_rn((float)(v[u] & 0xfff), lg.hx) == ni) *dstp++ = v[u];
                    }
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            make_node(ni, beg, run);
        }
        total = run;
    }
    __syncwarp();

    bool finish = (total == 0);
    while (!finish) {
        const int prev_size = w.size;
        int nbig = 0, n_expand = 0;
        int it = w.head;
        while (it >= 0 && w.pool[it].no_more) it = w.pool[it].next;
        uint32_t pre[4]; bool has_pre = qt_prefetch(w, it, A, B, lane, pre);
        while (it >= 0) {
            const QtNode nd = w.pool[it];
            __syncwarp();
            if (w.nfree < 4) { if (lane == 0) atomicExch(errflag, 1); finish = true; break; }
            int nx = nd.next;                                   // the next node that can still be divided: its keys are requested now
            while (nx >= 0 && w.pool[nx].no_more) nx = w.pool[nx].next;
            uint32_t pre_n[4]; const bool has_n = qt_prefetch(w, nx, A, B, lane, pre_n);
            int cnt[4], mx, my;
            qt_divide(nd, A, B, lane, cnt, mx, my, pre, has_pre);
            n_expand += qt_split(w, it, nd, cnt, mx, my, lane, nbig);
            it = nx; has_pre = has_n;
#pragma unroll
            for (int u = 0; u < 4; u++) pre[u] = pre_n[u];
        }
        if (finish) break;
        if (w.size >= N || w.size == prev_size) finish = true;
        else if (w.size + n_expand * 3 > N) {
            while (!finish) {
                const int prev2 = w.size;
                const int m = nbig;
                // previous round's list -> prv_*, then order it by (count, seq) descending
                for (int i = lane; i < m; i += 32) { w.prv_cnt[i] = w.big_cnt[i]; w.prv_seq[i] = w.big_seq[i]; w.prv_id[i] = w.big_id[i]; }
                __syncwarp();
                for (int i = lane; i < m; i += 32) {
                    const int ci = w.prv_cnt[i], si = w.prv_seq[i];
                    int rank = 0;
                    for (int j = 0; j < m; j++) {
                        const int cj = w.prv_cnt[j], sj = w.prv_seq[j];
                        rank += (cj > ci) || (cj == ci && sj > si);
                    }
                    w.order[rank] = (short)i;
                }
                __syncwarp();
                nbig = 0;
                uint32_t pre[4]; bool has_pre = qt_prefetch(w, m > 0 ? w.prv_id[w.order[0]] : -1, A, B, lane, pre);
                for (int j = 0; j < m; j++) {
                    const int id = w.prv_id[w.order[j]];
                    const QtNode nd = w.pool[id];
                    __syncwarp();
                    if (w.nfree < 4) { if (lane == 0) atomicExch(errflag, 1); finish = true; break; }
                    uint32_t pre_n[4]; const bool has_n = qt_prefetch(w, j + 1 < m ? w.prv_id[w.order[j + 1]] : -1, A, B, lane, pre_n);
                    int cnt[4], mx, my;
                    qt_divide(nd, A, B, lane, cnt, mx, my, pre, has_pre);
                    qt_split(w, id, nd, cnt, mx, my, lane, nbig);
                    has_pre = has_n;
#pragma unroll
                    for (int u = 0; u < 4; u++) pre[u] = pre_n[u];
                    if (w.size >= N) break;
                }
                if (w.size >= N || w.size == prev2) finish = true;
            }
        }
    }

    // ---- best response per node in list order, first maximum wins (ORBextractor.cc:742-760)
    int nres = 0;
    for (int it = w.head; it >= 0; it = w.pool[it].next) {
        if (lane == 0) w.order[nres] = (short)it;
        nres++;
    }
    __syncwarp();
    if (nres > lg.kp_cap) { if (lane == 0) atomicExch(errflag, 2); nres = lg.kp_cap; }
    for (int i = lane; i < nres; i += 32) {
        const QtNode nd = w.pool[w.order[i]];
        const uint32_t* src = nd.buf ? B : A;
        uint32_t best = src[nd.beg];
        for (int k = nd.beg + 1; k < nd.end; k++) {
            const uint32_t key = src[k];
            if ((key >> 24) > (best >> 24)) best = key;
        }
        // back to level coordinates (ORBextractor.cc:842-843)
        out[i] = (((best & 0xfff) + kMinBorder)) | ((((best >> 12) & 0xfff) + kMinBorder) << 12) | (best & 0xff000000u);
    }
    if (lane == 0) *out_cnt = nres;
}

// ------------------------------------------------------------------------------------------------
// K4+K5+K6: one warp per keypoint.  The 43x43 source patch arrives as one TMA box (64 x 43 bytes starting at the
// patch's x rounded down to 16; patches that touch the image border go byte by byte through the REFLECT_101 map).
//   IC_Angle over the radius-15 disc: per (row, 4-pixel word) two dp4a (u-weighted sum and plain row sum) on the word
//     masked to the disc; fastAtan2 in explicitly rounded float ops
//   7x7 Gaussian (Q8 kernel 18,34,48,56,48,34,18; (V+32768)>>16; exact integer arithmetic, so any evaluation order
//     gives OpenCV's result) on the 37x37 region the pattern can reach: horizontal pass with dp4a (4 outputs from 3
//     words), stored TRANSPOSED as u16 so that the vertical pass is dp2a over row pairs (2 outputs from 4 words)
//   256 steered comparisons, lane i produces descriptor byte i; a, b = float(cos/sin(double(angle_rad))) from an
//     fdlibm-style double kernel (quadrant reduction with a two-part pi/2, degree-13/14 polynomials)
// ------------------------------------------------------------------------------------------------
constexpr int kDescWarps = 4;
constexpr int kPatchR = 21;                  // 18 (max rotated pattern reach) + 3 (blur taps)
constexpr int kPatchW = 2 * kPatchR + 1;     // 43
constexpr int kBlurW = 37;
constexpr int kPatchPitch = 64;              // = TMA box width: patch column 0 sits at byte ox = x0 & 15
constexpr int kHTPitch = 46;                 // u16 pitch of the transposed horizontally filtered patch: HT[c][r]; 23 words: odd, so a column per lane is conflict-free
constexpr int kHTCols = 40;

struct DescMaps { CUtensorMap m[kMaxLevels]; };

__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {     // cv::fastAtan2, SURVEY A-4
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// sin and cos of x in [0, 2 pi + eps] in double, error below one ulp (fdlibm kernels after a quadrant reduction)
__device__ __forceinline__ void sincos_small(double x, double& sn, double& cs) {
    const double k = rint(x * 6.36619772367581382433e-01);
    double r = __fma_rn(-k, 1.57079632679489655800e+00, x);
    r = __fma_rn(-k, 6.12323399573676603587e-17, r);
    const double z = r * r;
    double ps = __fma_rn(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = __fma_rn(z, ps, 2.75573137070700676789e-06);
    ps = __fma_rn(z, ps, -1.98412698298579493134e-04);
    ps = __fma_rn(z, ps, 8.33333333332248946124e-03);
    ps = __fma_rn(z, ps, -1.66666666666666324348e-01);
    const double sr = __fma_rn(r * z, ps, r);
    double pc = __fma_rn(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = __fma_rn(z, pc, -2.75573143513906633035e-07);
    pc = __fma_rn(z, pc, 2.48015872894767294178e-05);
    pc = __fma_rn(z, pc, -1.38888888888741095749e-03);
    pc = __fma_rn(z, pc, 4.16666666666666019037e-02);
    const double cr = __fma_rn(z * z, pc, __fma_rn(z, -0.5, 1.0));
    const int q = (int)k & 3;
    sn = (q & 1) ? cr : sr; cs = (q & 1) ? sr : cr;
    if (q & 2) sn = -sn;
    if (q == 1 || q == 2) cs = -cs;
}

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {          // unsigned pixels x signed coefficients
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__global__ void __launch_bounds__(kDescWarps * 32)
k_describe(const uint8_t* __restrict__ img0, long long img_row_stride, long long img_frame_stride,
           const uint8_t* __restrict__ pyr, const __grid_constant__ OrbGeom g, const __grid_constant__ DescMaps maps,
           int tma_level0, int scratch_base,
           const uint32_t* __restrict__ lvlres, const int* __restrict__ lvlcnt,
           b200_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int32_t* __restrict__ counts, int out_cap) {
    __shared__ __align__(128) uint8_t s_patch[kDescWarps][(kPatchW * kPatchPitch + 127) / 128 * 128];       // TMA destinations: 128-byte aligned
    __shared__ __align__(16) uint16_t s_ht[kDescWarps][kHTCols * kHTPitch];
    __shared__ __align__(8) unsigned long long s_bar[kDescWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const int k = blockIdx.x * kDescWarps + warp;            // keypoint index inside the frame (levels concatenated)
    const int* lc = lvlcnt + (long long)f * g.nlevels;
    int level = -1, base = 0, total = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const int c = lc[l];
        if (level < 0 && k < total + c) { level = l; base = total; }
        total += c;
    }
    if (k == 0 && lane == 0) counts[f] = min(total, out_cap);
    if (level < 0 || k >= out_cap) return;
    const LevelGeom& lg = g.L[level];
    const uint32_t key = lvlres[(long long)f * g.res_per_frame + lg.kp_base + (k - base)];
    const int cx = key & 0xfff, cy = (key >> 12) & 0xfff, resp = key >> 24;

    uint8_t* P = s_patch[warp];
    const int x0 = cx - kPatchR, y0 = cy - kPatchR;
    const int ox = x0 & 15;
    if ((level > 0 || tma_level0) && x0 >= 0 && y0 >= 0 && cx + kPatchR < lg.w && cy + kPatchR < lg.h) {
        const uint32_t bar = smem_u32(&s_bar[warp]);
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kPatchPitch * kPatchW) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(smem_u32(P)), "l"(reinterpret_cast<uint64_t>(&maps.m[level])), "r"(x0 - ox), "r"(y0),
                            "r"(level > 0 ? scratch_base + f : f), "r"(bar) : "memory");
        }
        __syncwarp();
        asm volatile("{\n.reg .pred p;\nDESC_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DESC_DONE;\nbra DESC_WAIT;\nDESC_DONE:\n}"
                     :: "r"(bar) : "memory");
    } else {
        const uint8_t* im; long long pitch;
        if (level == 0) { im = img0 + (long long)f * img_frame_stride; pitch = img_row_stride; }
        else { im = pyr + (long long)f * g.pyr_frame_stride + lg.offset; pitch = lg.pitch; }
        for (int idx = lane; idx < kPatchW * kPatchW; idx += 32) {
            const int r = idx / kPatchW, c = idx - r * kPatchW;
            const int yy = reflect101(y0 + r, lg.h), xx = reflect101(x0 + c, lg.w);
            P[r * kPatchPitch + ox + c] = im[(long long)yy * pitch + xx];
        }
        __syncwarp();
    }
    const uint32_t* P32 = reinterpret_cast<const uint32_t*>(P);
    const int sh8 = 8 * (ox & 3);

    // IC_Angle (ORBextractor.cc:77-104): task = (row v, word j); the word holds u = -15 + 4j .. -12 + 4j of that row
    int m10 = 0, m01 = 0;
    {
        const int wb = (ox + kPatchR - kHalfPatch) >> 2;                 // word of patch column 6 (u = -15) ...
        const int sh = 8 * ((ox + kPatchR - kHalfPatch) & 3);            // ... and its byte inside that word
        for (int t = lane; t < 31 * 8; t += 32) {
            const int v = (t >> 3) - kHalfPatch, j = t & 7;
            const uint32_t* w = P32 + (v + kPatchR) * (kPatchPitch / 4) + wb + j;
            const uint32_t px = __funnelshift_r(w[0], w[1], sh) & __ldg(&g_disc[t]);
            const int u0 = -kHalfPatch + 4 * j;
            const int ucoef = (u0 & 0xff) | (((u0 + 1) & 0xff) << 8) | (((u0 + 2) & 0xff) << 16) | (((u0 + 3) & 0xff) << 24);
            m10 = dp4a_us(px, ucoef, m10);
            m01 += v * (int)__dp4a(px, 0x01010101u, 0u);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // Gaussian, horizontal pass over 43 rows x 40 columns (37 used): lane = (row pair, group of 4 outputs), fixed for the whole pass; the two rows'
    // sums of a column go out as ONE word of the transposed tile, HT[c][2 rp .. 2 rp + 1] (row 43 does not exist: it repeats row 42 and is never used)
    uint16_t* HT = s_ht[warp];
    {
        const unsigned K0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), K1 = 48u | (34u << 8) | (18u << 16);
        const int gq = lane % 10, rsub = lane / 10;
        uint32_t* HT32 = reinterpret_cast<uint32_t*>(HT) + 4 * gq * (kHTPitch / 2);
        const uint32_t* wbase = P32 + (ox >> 2) + gq;
        if (rsub < 3) {
#pragma unroll 2
            for (int rp = rsub; rp < 22; rp += 3) {
                const uint32_t* wa = wbase + (2 * rp) * (kPatchPitch / 4);
                const uint32_t* wb2 = wbase + min(2 * rp + 1, kPatchW - 1) * (kPatchPitch / 4);
                const uint32_t a0 = wa[0], a1 = wa[1], a2 = wa[2], a3 = wa[3], b0 = wb2[0], b1 = wb2[1], b2 = wb2[2], b3 = wb2[3];
                const uint32_t u0 = __funnelshift_r(a0, a1, sh8), u1 = __funnelshift_r(a1, a2, sh8), u2 = __funnelshift_r(a2, a3, sh8);
                const uint32_t v0 = __funnelshift_r(b0, b1, sh8), v1 = __funnelshift_r(b1, b2, sh8), v2 = __funnelshift_r(b2, b3, sh8);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t ha = __dp4a(__funnelshift_r(u0, u1, 8 * i), K0, __dp4a(__funnelshift_r(u1, u2, 8 * i), K1, 0u));
                    const uint32_t hb = __dp4a(__funnelshift_r(v0, v1, 8 * i), K0, __dp4a(__funnelshift_r(v1, v2, 8 * i), K1, 0u));
                    HT32[i * (kHTPitch / 2) + rp] = ha + (hb << 16);
                }
            }
        }
    }
    __syncwarp();
    // vertical pass: lane = column, walking down the row pairs with a sliding window of four words (rows 2p .. 2p + 7):
    // Bl[2p][c], Bl[2p + 1][c] from 8 dp2a; the odd pitch of HT keeps the 32 columns on 32 banks
    uint8_t* Bl = P;                     // the source patch is dead once HT exists: reuse its storage
    {
        const unsigned K01 = 18u | (34u << 8), K23 = 48u | (56u << 8), K45 = 48u | (34u << 8), K6 = 18u, K6h = 18u << 24;
        const uint32_t* HT32 = reinterpret_cast<const uint32_t*>(HT);
        {
            const uint32_t* w = HT32 + lane * (kHTPitch / 2);
            uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            uint32_t s01 = sh16(w0, w1), s12 = sh16(w1, w2);
#pragma unroll
            for (int pp = 0; pp < 19; pp++) {
                const uint32_t w3 = w[pp + 3];
                const uint32_t s23 = sh16(w2, w3);
                unsigned e = __dp2a_lo(w0, K01, 32768u); e = __dp2a_lo(w1, K23, e); e = __dp2a_lo(w2, K45, e); e = __dp2a_lo(w3, K6, e);
                unsigned o = __dp2a_lo(s01, K01, 32768u); o = __dp2a_lo(s12, K23, o); o = __dp2a_lo(s23, K45, o); o = __dp2a_hi(w3, K6h, o);
                Bl[(2 * pp) * kBlurW + lane] = (uint8_t)(e >> 16);
                if (2 * pp + 1 < kBlurW) Bl[(2 * pp + 1) * kBlurW + lane] = (uint8_t)(o >> 16);
                w0 = w1; w1 = w2; w2 = w3; s01 = s12; s12 = s23;
            }
        }
        // columns 32 .. 36: lane = (column, every sixth row pair)
        if (lane < 30) {
            const int c = 32 + lane % 5;
            for (int pp = lane / 5; pp < 19; pp += 6) {
                const uint32_t* w = HT32 + c * (kHTPitch / 2) + pp;
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
                unsigned e = __dp2a_lo(w0, K01, 32768u); e = __dp2a_lo(w1, K23, e); e = __dp2a_lo(w2, K45, e); e = __dp2a_lo(w3, K6, e);
                unsigned o = __dp2a_lo(sh16(w0, w1), K01, 32768u); o = __dp2a_lo(sh16(w1, w2), K23, o); o = __dp2a_lo(sh16(w2, w3), K45, o); o = __dp2a_hi(w3, K6h, o);
                Bl[(2 * pp) * kBlurW + c] = (uint8_t)(e >> 16);
                if (2 * pp + 1 < kBlurW) Bl[(2 * pp + 1) * kBlurW + c] = (uint8_t)(o >> 16);
            }
        }
    }
    __syncwarp();

    // steered BRIEF (ORBextractor.cc:107-147); canonical a,b = float(cos/sin(double(angle_rad)))
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    const float ang = __fmul_rn(angle, factorPI);
    double sd, cd;
    sincos_small((double)ang, sd, cd);
    const float a = __double2float_rn(cd), b = __double2float_rn(sd);
    int val = 0;
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        const float px0 = (float)__ldg(&g_pattern[(kk * 4 + 0) * 32 + lane]), py0 = (float)__ldg(&g_pattern[(kk * 4 + 1) * 32 + lane]);
        const float px1 = (float)__ldg(&g_pattern[(kk * 4 + 2) * 32 + lane]), py1 = (float)__ldg(&g_pattern[(kk * 4 + 3) * 32 + lane]);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(px0, b), __fmul_rn(py0, a))), c0 = __float2int_rn(__fsub_rn(__fmul_rn(px0, a), __fmul_rn(py0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(px1, b), __fmul_rn(py1, a))), c1 = __float2int_rn(__fsub_rn(__fmul_rn(px1, a), __fmul_rn(py1, b)));
        const int t0 = Bl[(r0 + 18) * kBlurW + c0 + 18], t1 = Bl[(r1 + 18) * kBlurW + c1 + 18];
        val |= (t0 < t1) << kk;
    }
    desc[((long long)f * out_cap + k) * 32 + lane] = (uint8_t)val;

    // keypoint record (ORBextractor.cc:837-847,1095-1101)
    float kx = (float)cx, ky = (float)cy;
    if (level != 0) { kx = __fmul_rn(kx, lg.scale); ky = __fmul_rn(ky, lg.scale); }
    if (lane < 7) {
        uint32_t wv;
        switch (lane) {
            case 0: wv = __float_as_uint(kx); break;
            case 1: wv = __float_as_uint(ky); break;
            case 2: wv = __float_as_uint(lg.size); break;
            case 3: wv = __float_as_uint(angle); break;
            case 4: wv = __float_as_uint((float)resp); break;
            case 5: wv = (uint32_t)level; break;
            default: wv = 0xffffffffu; break;   // class_id = -1
        }
        reinterpret_cast<uint32_t*>(kps + (long long)f * out_cap + k)[lane] = wv;
    }
}

}  // namespace b200

// =================================================================================================
// Host side: handle, geometry, C-ABI
// =================================================================================================
using namespace b200;

struct b200_orb_s {
    int device;
    cudaStream_t stream;
    int nfeatures, nlevels, ini_th, min_th;
    float scale;
    int max_w, max_h, max_batch;
    std::vector<float> sf, inv_sf, sigma2, inv_sigma2;
    std::vector<int> quota;
    // geometry of the image size in use
    int cur_w, cur_h;
    OrbGeom geom;
    std::vector<CellDesc> cells;
    int pool_cap;
    FastMaps smaps; std::vector<ScoreTile> tiles; ScoreTile* d_tiles; size_t cap_tiles; uint8_t* d_score; size_t cap_score; int nms_warp_words;   // dense FAST: score tiles, score map
    FastSmemGeom fsg; size_t fast_smem; FastMaps maps; DescMaps dmaps; FastMaps pmaps;      // pmaps.m[l]: source level l - 1 with level l's pyramid box      // maps.m[l > 0]: level l of the pyramid buffer; m[0] is encoded per call
    // device buffers
    uint8_t* d_pyr; ResizeEntry* d_tab; CellDesc* d_cells; uint32_t* d_slots; int* d_cellcnt;
    uint32_t *d_keysA, *d_keysB, *d_lvlres; int* d_lvlcnt; int* d_err;
    size_t cap_pyr, cap_tab, cap_cells, cap_slots, cap_keysA, cap_keysB, cap_cellcnt, cap_lvlres;
    // staging for the host-pointer API
    uint8_t* d_in; size_t cap_in; uint8_t* h_in; size_t cap_hin;
    b200_keypoint* d_kps; uint8_t* d_desc; int32_t* d_counts; size_t cap_kps, cap_desc, cap_counts;
    uint8_t* h_out; size_t cap_hout;
    // extra device outputs of b200_frontend_host (detector + matcher results, reference set)
    b200_marker* d_markers; int32_t* d_mcounts; int32_t* d_match; int32_t* d_nmatch; uint8_t* d_refdesc; b200_keypoint* d_refkps;
    size_t cap_markers, cap_mcounts, cap_match, cap_nmatch, cap_refdesc, cap_refkps;
    cudaStream_t aux_stream; cudaEvent_t ev_aux;
    // optional per-stage timing (b200_orb_set_profile): events around pyramid / fast / quadtree / describe
    int profile; cudaEvent_t ev[5]; float stage_ms[4]; int stage_valid, stage_frames;
    cudaStream_t copy_stream; cudaEvent_t ev_copy[2];
    cudaStream_t down_stream; cudaEvent_t ev_done[2];
    cudaStream_t stream2, aux_stream2; cudaEvent_t ev_aux2, ev_ref;       // second stream set: odd chunks of b200_frontend_host
    // last b200_frontend_host call (b200_frontend_collate_host gathers its device-resident results) and the root's receive buffers
    int fe_n, fe_mcap, fe_aruco, fe_match;
    uint8_t* d_coll[7]; size_t cap_coll[7];
    // last call (debug taps)
    const uint8_t* last_imgs; long long last_row_stride, last_frame_stride; int last_n, last_base;
};

namespace {

template <typename T> int ensure(T*& p, size_t& cap, size_t need_bytes) {
    if (need_bytes <= cap && p) return B200_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    B200_CUDA(cudaMalloc((void**)&p, need_bytes));
    cap = need_bytes;
    return B200_OK;
}

// 3-D u8 tensor map (x, y, frame) with a box of box_w x box_h x 1 for the TMA tile loads of k_fast
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_tile_map(CUtensorMap* m, const void* base, int w, int hgt, int nframes, long long row_stride, long long frame_stride, int box_w, int box_h) {
    static EncodeTiledFn encode = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    if (!encode) return fail(B200_ECUDA, "driver entry point %s not found", "cuTensorMapEncodeTiled");
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)hgt, (cuuint64_t)std::max(nframes, 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)row_stride, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u}, estr[3] = {1u, 1u, 1u};
    const CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200_ECUDA, "cuTensorMapEncodeTiled %s", "failed");
    return B200_OK;
}

// ORBextractor.cc:410-446
void make_tables(b200_orb_s* h) {
    const int n = h->nlevels;
    const double scale_d = h->scale;
    h->sf.assign(n, 1.f); h->inv_sf.assign(n, 1.f); h->sigma2.assign(n, 1.f); h->inv_sigma2.assign(n, 1.f);
    for (int i = 1; i < n; i++) { h->sf[i] = (float)(h->sf[i - 1] * scale_d); h->sigma2[i] = h->sf[i] * h->sf[i]; }
    for (int i = 0; i < n; i++) { h->inv_sf[i] = 1.0f / h->sf[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    h->quota.assign(n, 0);
    const float factor = (float)(1.0f / scale_d);
    float nd = h->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n));
    int sum = 0;
    for (int l = 0; l < n - 1; l++) { h->quota[l] = host_round(nd); sum += h->quota[l]; nd *= factor; }
    h->quota[n - 1] = std::max(h->nfeatures - sum, 0);
}

// cv::resize INTER_LINEAR coefficient table (SURVEY A-1)
void resize_table(int ssize, int dsize, ResizeEntry* t) {
    const double scale = (double)ssize / dsize;
    for (int d = 0; d < dsize; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(fx);
        fx -= s;
        if (s < 0) { s = 0; fx = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; fx = 0.f; }
        t[d].ofs = s;
        t[d].c0 = (short)host_round((1.f - fx) * 2048.f);
        t[d].c1 = (short)host_round(fx * 2048.f);
    }
}

int set_geometry(b200_orb_s* h, int w, int h_img) {
    if (w == h->cur_w && h_img == h->cur_h) return B200_OK;
    OrbGeom& g = h->geom;
    memset(&g, 0, sizeof(g));
    g.nlevels = h->nlevels; g.ini_th = h->ini_th; g.min_th = h->min_th;
    h->cells.clear();
    std::vector<ResizeEntry> tab;
    long long pyr_ofs = 0, slot = 0;
    int res = 0, maxcap = 0;
    for (int l = 0; l < h->nlevels; l++) {
        LevelGeom& L = g.L[l];
        L.w = host_round((float)w * h->inv_sf[l]);
        L.h = host_round((float)h_img * h->inv_sf[l]);
        if (L.w < 1 || L.h < 1) return fail(B200_EINVAL, "image too small for %s pyramid levels", "this many");
        L.quota = h->quota[l]; L.scale = h->sf[l]; L.size = (float)(int)(31 * h->sf[l]);
        if (l > 0) {
            L.pitch = (int)align_up(L.w, 16); L.offset = pyr_ofs;
            pyr_ofs += align_up((long long)L.pitch * L.h, 256);
            L.xtab = (int)tab.size(); tab.resize(tab.size() + L.w); resize_table(g.L[l - 1].w, L.w, &tab[L.xtab]);
            L.ytab = (int)tab.size(); tab.resize(tab.size() + L.h); resize_table(g.L[l - 1].h, L.h, &tab[L.ytab]);
            {   // source box of k_pyramid_tma: the largest source region a 128 x 64 output tile touches
                const int sw = g.L[l - 1].w, sh = g.L[l - 1].h;
                int bw = 16, bh = 1;
                for (int x0 = 0; x0 < L.w; x0 += 128) {
                    const int xs0 = tab[L.xtab + x0].ofs & ~15, xs1 = std::min(tab[L.xtab + std::min(x0 + 127, L.w - 1)].ofs + 1, sw - 1);
                    bw = std::max(bw, xs1 - xs0 + 1);
                }
                for (int y0 = 0; y0 < L.h; y0 += 8 * kPyrRows) {
                    const int ys0 = tab[L.ytab + y0].ofs, ys1 = std::min(tab[L.ytab + std::min(y0 + 8 * kPyrRows - 1, L.h - 1)].ofs + 1, sh - 1);
                    bh = std::max(bh, ys1 - ys0 + 1);
                }
                L.pbox_w = (int)align_up(bw, 16); L.pbox_h = bh;
            }
        }
        // cells (ORBextractor.cc:771-806)
        const int minB = kMinBorder, maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        L.ncols = (int)(width / 30.f); L.nrows = (int)(height / 30.f);
        L.cell_base = (int)h->cells.size(); L.slot_base = slot;
        L.nini = 0; L.hx = 1.f;
        if (L.ncols >= 1 && L.nrows >= 1) {
            L.wcell = (int)ceilf(width / L.ncols); L.hcell = (int)ceilf(height / L.nrows);
            for (int i = 0; i < L.nrows; i++) {
                const float iniY = (float)(minB + i * L.hcell);
                float maxY = iniY + L.hcell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < L.ncols; j++) {
                    const float iniX = (float)(minB + j * L.wcell);
                    float maxX = iniX + L.wcell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    CellDesc c;
                    c.level = (short)l; c.pad = 0; c.x0 = (short)iniX; c.y0 = (short)iniY;
                    c.rw = (short)((int)maxX - (int)iniX); c.rh = (short)((int)maxY - (int)iniY);
                    c.sx = (short)(j * L.wcell); c.sy = (short)(i * L.hcell);
                    if (c.rw < 7 || c.rh < 7) continue;      // cv::FAST finds nothing in ROIs without an interior
                    if (c.rw > kMaxRoi || c.rh > kMaxRoi) return fail(B200_EINVAL, "cell ROI larger than %s", "kMaxRoi");
                    c.slot = (int)slot;
                    c.cap = ((c.rw - 6 + 1) / 2) * ((c.rh - 6 + 1) / 2);   // strict 3x3 maxima: at most one per 2x2 block
                    slot += c.cap;
                    h->cells.push_back(c);
                }
            }
            const int nini = (int)roundf((float)(maxBX - minB) / (float)(maxBY - minB));
            if (nini >= 1) {
                L.nini = nini; L.hx = (float)(maxBX - minB) / nini;
                int xs = 0;
                while (xs < 4096 && (int)((float)xs / L.hx) < 1) xs++;      // float division like the kernel's __fdiv_rn: monotone in x
                L.xsplit = xs;
            }
        }
        L.ncells = (int)h->cells.size() - L.cell_base;
        L.kp_cap = L.quota + 3 + 4 * std::max(L.nini, 1);
        L.kp_base = res; res += L.kp_cap;
        maxcap = std::max(maxcap, L.kp_cap);
        if (L.w > 4095 || L.h > 4095) return fail(B200_EINVAL, "level larger than %s", "4095 px (12-bit packed coordinates)");
    }
    g.total_cells = (int)h->cells.size();
    {   // dense FAST: score map layout, score tiles over [16, w - 16) x [16, h - 16) of every level, and the per-warp tile of k_fast_nms
        long long sofs = 0;
        h->tiles.clear();
        for (int l = 0; l < h->nlevels; l++) {
            LevelGeom& L = g.L[l];
            L.spitch = (int)align_up(L.w, 16); L.soff = sofs;
            sofs += align_up((long long)L.spitch * L.h, 256);
            if (L.ncells == 0) continue;
            for (int ty = 0; 16 + kScTH * ty < L.h - 16; ty++)
                for (int tx = 0; 16 + kScTW * tx < L.w - 16; tx++) { ScoreTile t; t.level = (short)l; t.tx = (short)tx; t.ty = (short)ty; t.pad = 0; h->tiles.push_back(t); }
        }
        g.score_frame_stride = std::max<long long>(sofs, 256);
        g.store_th = std::min(h->ini_th, h->min_th);
        int words = 16;
        for (const CellDesc& c : h->cells) {
            const int iw = c.rw - 6, ih = c.rh - 6, xi = c.x0 + 3, xa = (xi - 1) & ~3, wpr = (xi + iw + 1 - xa + 3) >> 2;
            words = std::max(words, (ih + 2) * wpr);
        }
        h->nms_warp_words = (words + 3) & ~3;
    }
    {   // TMA boxes and the shared-memory carve-up of k_fast
        FastSmemGeom& sgm = h->fsg;
        sgm.tile_bytes = 128; sgm.qcap = 8; sgm.keepw = 3;
        for (int l = 0; l < h->nlevels; l++) {
            LevelGeom& L = g.L[l];
            int mrw = 7, mrh = 7;
            for (int c = 0; c < L.ncells; c++) {
                const CellDesc& cd = h->cells[L.cell_base + c];
                mrw = std::max(mrw, (int)cd.rw); mrh = std::max(mrh, (int)cd.rh);
                const int iw = cd.rw - 6, ih = cd.rh - 6, tc0 = (cd.x0 & 15) + 3, ng = ((tc0 + iw - 1) >> 2) - (tc0 >> 2) + 1, rpi = 32 / std::min(ng, 32);
                const int rows_w = (ih + kFastMinWarps * rpi - 1) / (kFastMinWarps * rpi) * rpi;          // rows one warp pretests
                sgm.qcap = std::max(sgm.qcap, (rows_w * 4 * ng + 7) & ~7);
                sgm.keepw = std::max(sgm.keepw, ih * 3);
            }
            L.box_w = (int)align_up(mrw + 15, 16); L.box_h = mrh;
        }
        // one tile pitch for all levels (k_fast is instantiated for 64, 80 and 96 bytes)
        int tp = 64;
        for (int l = 0; l < h->nlevels; l++) tp = std::max(tp, g.L[l].box_w);
        { static const int env_tp = [] { const char* e = getenv("B200_FAST_TP"); return e ? atoi(e) : 0; }(); if (env_tp > tp) tp = env_tp; }
        tp = tp <= 64 ? 64 : tp <= 80 ? 80 : 96;
        for (int l = 0; l < h->nlevels; l++) {
            g.L[l].box_w = tp;
            sgm.tile_bytes = std::max(sgm.tile_bytes, (int)align_up((long long)tp * g.L[l].box_h, 128));
        }
        h->fast_smem = (size_t)2 * sgm.tile_bytes + (size_t)sgm.keepw * 4;       // + warps * qcap * 2 at launch
    }
    g.slots_per_frame = slot;
    g.pyr_frame_stride = std::max<long long>(pyr_ofs, 256);
    g.res_per_frame = res;
    h->pool_cap = maxcap + 8;
    if (h->pool_cap > 32000) return fail(B200_EINVAL, "nfeatures per level too large for %s", "16-bit node ids");

    const size_t B = (size_t)h->max_batch;
    int rc;
    if ((rc = ensure(h->d_pyr, h->cap_pyr, (size_t)g.pyr_frame_stride * B))) return rc;
    if ((rc = ensure(h->d_tab, h->cap_tab, std::max<size_t>(tab.size(), 1) * sizeof(ResizeEntry)))) return rc;
    if ((rc = ensure(h->d_cells, h->cap_cells, std::max<size_t>(h->cells.size(), 1) * sizeof(CellDesc)))) return rc;
    size_t slots_bytes = std::max<size_t>((size_t)g.slots_per_frame * B * 4, 4);
    if ((rc = ensure(h->d_slots, h->cap_slots, slots_bytes))) return rc;
    if ((rc = ensure(h->d_keysA, h->cap_keysA, slots_bytes))) return rc;
    if ((rc = ensure(h->d_keysB, h->cap_keysB, slots_bytes))) return rc;
    if ((rc = ensure(h->d_cellcnt, h->cap_cellcnt, std::max<size_t>((size_t)g.total_cells * B * 4, 4)))) return rc;
    if ((rc = ensure(h->d_score, h->cap_score, (size_t)g.score_frame_stride * B))) return rc;
    if ((rc = ensure(h->d_tiles, h->cap_tiles, std::max<size_t>(h->tiles.size(), 1) * sizeof(ScoreTile)))) return rc;
    if (!h->tiles.empty()) B200_CUDA(cudaMemcpy(h->d_tiles, h->tiles.data(), h->tiles.size() * sizeof(ScoreTile), cudaMemcpyHostToDevice));
    if ((rc = ensure(h->d_lvlres, h->cap_lvlres, (size_t)res * B * 4))) return rc;
    if (!h->d_lvlcnt) B200_CUDA(cudaMalloc((void**)&h->d_lvlcnt, (size_t)kMaxLevels * B * 4));
    if (!h->d_err) { B200_CUDA(cudaMalloc((void**)&h->d_err, 4)); B200_CUDA(cudaMemset(h->d_err, 0, 4)); }
    if (!tab.empty()) B200_CUDA(cudaMemcpy(h->d_tab, tab.data(), tab.size() * sizeof(ResizeEntry), cudaMemcpyHostToDevice));
    if (!h->cells.empty()) B200_CUDA(cudaMemcpy(h->d_cells, h->cells.data(), h->cells.size() * sizeof(CellDesc), cudaMemcpyHostToDevice));
    for (int l = 1; l < h->nlevels; l++) {
        const LevelGeom& L = g.L[l];
        if ((rc = make_tile_map(&h->maps.m[l], h->d_pyr + L.offset, L.w, L.h, h->max_batch, L.pitch, g.pyr_frame_stride, L.box_w, L.box_h))) return rc;
        if ((rc = make_tile_map(&h->smaps.m[l], h->d_pyr + L.offset, L.w, L.h, h->max_batch, L.pitch, g.pyr_frame_stride, kScBoxW, kScBoxH))) return rc;
        if ((rc = make_tile_map(&h->dmaps.m[l], h->d_pyr + L.offset, L.w, L.h, h->max_batch, L.pitch, g.pyr_frame_stride, kPatchPitch, kPatchW))) return rc;
        if (l + 1 < h->nlevels && g.L[l + 1].pbox_w <= 256 && g.L[l + 1].pbox_h <= 256 &&
            (rc = make_tile_map(&h->pmaps.m[l + 1], h->d_pyr + L.offset, L.w, L.h, h->max_batch, L.pitch, g.pyr_frame_stride, g.L[l + 1].pbox_w, g.L[l + 1].pbox_h)))
            return rc;
    }
    h->cur_w = w; h->cur_h = h_img;
    return B200_OK;
}

int upload_constants() {
    static const signed char pat[256 * 4] = {
#include "orb_pattern.inc"
    };
    signed char t[256 * 4];
    // descriptor byte i uses pairs 8i..8i+7; lane i reads [k][coord][i]
    for (int i = 0; i < 32; i++)
        for (int k = 0; k < 8; k++)
            for (int c = 0; c < 4; c++) t[(k * 4 + c) * 32 + i] = pat[(8 * i + k) * 4 + c];
    B200_CUDA(cudaMemcpyToSymbol(g_pattern, t, sizeof(t)));
    // umax (ORBextractor.cc:454-469)
    int umax[16];
    const int vmax = (int)floor(kHalfPatch * sqrt(2.f) / 2 + 1), vmin = (int)ceil(kHalfPatch * sqrt(2.f) / 2);
    for (int v = 0; v <= vmax; v++) umax[v] = host_round_d(sqrt((double)kHalfPatch * kHalfPatch - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0; ++v0;
    }
    B200_CUDA(cudaMemcpyToSymbol(c_umax, umax, sizeof(umax)));
    uint32_t disc[31 * 8];
    for (int i = 0; i < 31 * 8; i++) {
        const int um = umax[abs((i >> 3) - 15)];
        uint32_t m = 0;
        for (int k = 0; k < 4; k++) { const int u = -15 + 4 * (i & 7) + k; if (abs(u) <= um) m |= 0xffu << (8 * k); }
        disc[i] = m;
    }
    B200_CUDA(cudaMemcpyToSymbol(g_disc, disc, sizeof(disc)));
    return B200_OK;
}

int enqueue(b200_orb_s* h, const uint8_t* imgs, int n, int w, int hh, long long rs, long long fs,
            b200_keypoint* kps, uint8_t* desc, int32_t* counts, int out_cap, cudaStream_t st, int base = 0) {
    // `base`: first scratch frame slot used by this call (calls that may run concurrently on different streams use disjoint slots)
    const OrbGeom& g = h->geom;
    uint8_t* d_pyr = h->d_pyr + (size_t)base * g.pyr_frame_stride;
    uint32_t* d_slots = h->d_slots + (size_t)base * g.slots_per_frame;
    uint32_t* d_keysA = h->d_keysA + (size_t)base * g.slots_per_frame;
    uint32_t* d_keysB = h->d_keysB + (size_t)base * g.slots_per_frame;
    int* d_cellcnt = h->d_cellcnt + (size_t)base * g.total_cells;
    uint32_t* d_lvlres = h->d_lvlres + (size_t)base * g.res_per_frame;
    int* d_lvlcnt = h->d_lvlcnt + (size_t)base * g.nlevels;
    int tma0_used = 0;
    static const bool pyr_no_tma = getenv("B200_PYRAMID_NO_TMA") != nullptr;
    if (h->profile) B200_CUDA(cudaEventRecord(h->ev[0], st));
    for (int l = 1; l < g.nlevels; l++) {
        const LevelGeom& L = g.L[l];
        const LevelGeom& Lp = g.L[l - 1];
        const uint8_t* src = l == 1 ? imgs : d_pyr + Lp.offset;
        const long long srs = l == 1 ? rs : Lp.pitch, sfs = l == 1 ? fs : g.pyr_frame_stride;
        dim3 grid((L.w + 127) / 128, (L.h + 8 * kPyrRows - 1) / (8 * kPyrRows), n), block(32, 8);
        // the source tile comes through TMA when it can be described by a tensor map (pyramid levels always; the caller's frames when
        // pointer and strides are 16-byte multiples), else the taps are read from global memory
        bool tma = L.pbox_w <= 256 && L.pbox_h <= 256 && !pyr_no_tma;
        if (tma && l == 1) {
            const long long fs0 = n > 1 ? fs : align_up(rs * (long long)g.L[0].h, 16);
            tma = ((reinterpret_cast<uintptr_t>(imgs) | (uintptr_t)rs | (uintptr_t)fs0) & 15) == 0;
            if (tma) { int rc = make_tile_map(&h->pmaps.m[1], imgs, g.L[0].w, g.L[0].h, n, rs, fs0, L.pbox_w, L.pbox_h); if (rc) return rc; }
        }
        if (tma)
            B200_LAUNCH(k_pyramid_tma, grid, block, (size_t)L.pbox_w * L.pbox_h, st, h->pmaps.m[l], l == 1 ? 0 : base, L.pbox_w, L.pbox_h,
                        d_pyr + L.offset, L.pitch, g.pyr_frame_stride, Lp.w, Lp.h, L.w, L.h, h->d_tab + L.xtab, h->d_tab + L.ytab);
        else
            B200_LAUNCH(k_pyramid, grid, block, 0, st, src, srs, sfs, d_pyr + L.offset, L.pitch, g.pyr_frame_stride,
                        Lp.w, Lp.h, L.w, L.h, h->d_tab + L.xtab, h->d_tab + L.ytab);
    }
    if (h->profile) B200_CUDA(cudaEventRecord(h->ev[1], st));
    if (g.total_cells > 0) {
        dim3 grid(g.total_cells, n);
        // level 0 = the caller's frames: a tensor map needs a 16-byte aligned pointer and strides (else k_fast stages level 0 byte-wise)
        const long long fs0 = n > 1 ? fs : align_up(rs * (long long)g.L[0].h, 16);
        const int tma0 = ((reinterpret_cast<uintptr_t>(imgs) | (uintptr_t)rs | (uintptr_t)fs0) & 15) == 0;
        if (tma0) {
            int rc = make_tile_map(&h->maps.m[0], imgs, g.L[0].w, g.L[0].h, n, rs, fs0, g.L[0].box_w, g.L[0].box_h);
            if (!rc) rc = make_tile_map(&h->dmaps.m[0], imgs, g.L[0].w, g.L[0].h, n, rs, fs0, kPatchPitch, kPatchW);
            if (!rc) rc = make_tile_map(&h->smaps.m[0], imgs, g.L[0].w, g.L[0].h, n, rs, fs0, kScBoxW, kScBoxH);
            if (rc) return rc;
        }
        tma0_used = tma0;
        const int tp = g.L[0].box_w;
        // Three forms of K2, all bit-exact (the same tests run against each).  Default: k_fast, one CTA per cell (pretest + queues + paired scoring).
        // B200_FAST_DENSE=1: dense score map + per-cell NMS (k_fast_score + k_fast_nms).  B200_FAST_WARP=1: k_fastw, one warp per cell.
        // Measured on a B200, 256 frames of 640 x 480 (profiles/r2_fast_variants.md): 1.14 ms / 1006 M warp instructions for k_fast,
        // 0.76 + 0.37 ms / 503 M + 220 M for the dense pair (78 % of the score kernel's instructions are VIMNMX3 / PRMT on the ALU pipe, which issues
        // at 0.6 of the FFMA rate, tools/pipe_probe2.cu), and 802 M instructions but a longer run time for k_fastw (one warp per cell leaves too
        // little parallelism to hide its shared-memory round trips).  On the synthetic frames 32 % of all pixels pass the compass pretest at
        // iniThFAST and 8 % are corners, so none of the forms can skip much work.
        static const bool fast_dense = getenv("B200_FAST_DENSE") != nullptr;
        static const bool fast_warp = getenv("B200_FAST_WARP") != nullptr;
        if (fast_dense) {
            uint8_t* d_score = h->d_score + (size_t)base * g.score_frame_stride;
            dim3 gs((unsigned)h->tiles.size(), n);
            B200_LAUNCH(k_fast_score, gs, kScThreads, 0, st, imgs, rs, fs, d_pyr, g, h->smaps, tma0, base, h->d_tiles, d_score);
            dim3 gn((g.total_cells + kNmsWarps - 1) / kNmsWarps, n);
            B200_LAUNCH(k_fast_nms, gn, kNmsWarps * 32, 0, st, g, h->d_cells, d_score, d_slots, d_cellcnt);
        } else if (fast_warp) {
            const int keepw = h->fsg.keepw, tb = h->fsg.tile_bytes;
            const int warp_bytes = (int)align_up(2 * tb + (kFwQ + kFwC) * 2 + keepw * 4, 128);
            const size_t smem = (size_t)warp_bytes * kFwWarps;
            static std::atomic<size_t> fw_smem_set(0);
            if (smem > 48 * 1024 && smem > fw_smem_set.load()) {
                B200_CUDA(cudaFuncSetAttribute(k_fastw<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                B200_CUDA(cudaFuncSetAttribute(k_fastw<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                B200_CUDA(cudaFuncSetAttribute(k_fastw<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                fw_smem_set.store(smem);
            }
            dim3 gw((g.total_cells + kFwWarps - 1) / kFwWarps, n);
            if (tp == 64) B200_LAUNCH(k_fastw<64>, gw, kFwWarps * 32, smem, st, imgs, rs, fs, d_pyr, g, h->maps, tb, keepw, warp_bytes, tma0, base, h->d_cells, d_slots, d_cellcnt);
            else if (tp == 80) B200_LAUNCH(k_fastw<80>, gw, kFwWarps * 32, smem, st, imgs, rs, fs, d_pyr, g, h->maps, tb, keepw, warp_bytes, tma0, base, h->d_cells, d_slots, d_cellcnt);
            else B200_LAUNCH(k_fastw<96>, gw, kFwWarps * 32, smem, st, imgs, rs, fs, d_pyr, g, h->maps, tb, keepw, warp_bytes, tma0, base, h->d_cells, d_slots, d_cellcnt);
        } else {
            static const int nt = [] { const char* e = getenv("B200_FAST_THREADS"); const int v = e ? atoi(e) : 128; return v == 64 || v == 96 ? v : 128; }();
            static const int pad = [] { const char* e = getenv("B200_FAST_SMEM_PAD"); return e ? atoi(e) : 0; }();
            const size_t sm = h->fast_smem + (size_t)(nt / 32) * h->fsg.qcap * 2 + (size_t)pad;
#define B200_FAST_GO(TPV, NTV) B200_LAUNCH((k_fast<TPV, NTV>), grid, NTV, sm, st, imgs, rs, fs, d_pyr, g, h->maps, h->fsg, tma0, base, h->d_cells, d_slots, d_cellcnt)
#define B200_FAST_TP(NTV) do { if (tp == 64) B200_FAST_GO(64, NTV); else if (tp == 80) B200_FAST_GO(80, NTV); else B200_FAST_GO(96, NTV); } while (0)
            if (nt == 64) B200_FAST_TP(64); else if (nt == 96) B200_FAST_TP(96); else B200_FAST_TP(128);
#undef B200_FAST_TP
#undef B200_FAST_GO
        }
    }
    if (h->profile) B200_CUDA(cudaEventRecord(h->ev[2], st));
    {
        const int nprob = n * g.nlevels;
        // up to 64 frames (one wave of CTAs): key buffers of up to kQtSmemKeys keys per level live in shared memory
        const int key_cap = n <= 64 ? kQtSmemKeys : 0;
        const size_t per_warp = ((size_t)h->pool_cap * (sizeof(QtNode) + 2 + 2 * (4 + 4 + 2) + 2) + 64 + (size_t)key_cap * 8 + 15) & ~(size_t)15;
        const size_t smem = per_warp * kQtWarps;
        if (key_cap && smem <= 200 * 1024) {
            static std::atomic<size_t> qts_smem_set(0);
            if (smem > qts_smem_set.load()) { B200_CUDA(cudaFuncSetAttribute(k_quadtree<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); qts_smem_set.store(smem); }
            B200_LAUNCH(k_quadtree<true>, (nprob + kQtWarps - 1) / kQtWarps, kQtWarps * 32, smem, st, g, h->d_cells, d_slots, d_cellcnt,
                        d_keysA, d_keysB, d_lvlres, d_lvlcnt, n, h->pool_cap, h->d_err, key_cap);
        } else {
            const size_t per_warp0 = ((size_t)h->pool_cap * (sizeof(QtNode) + 2 + 2 * (4 + 4 + 2) + 2) + 64 + 15) & ~(size_t)15;
            const size_t smem0 = per_warp0 * kQtWarps;
            static std::atomic<size_t> qt_smem_set(0);
            if (smem0 > qt_smem_set.load()) { B200_CUDA(cudaFuncSetAttribute(k_quadtree<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0)); qt_smem_set.store(smem0); }
            B200_LAUNCH(k_quadtree<false>, (nprob + kQtWarps - 1) / kQtWarps, kQtWarps * 32, smem0, st, g, h->d_cells, d_slots, d_cellcnt,
                        d_keysA, d_keysB, d_lvlres, d_lvlcnt, n, h->pool_cap, h->d_err, 0);
        }
    }
    if (h->profile) B200_CUDA(cudaEventRecord(h->ev[3], st));
    {
        dim3 grid((g.res_per_frame + kDescWarps - 1) / kDescWarps, n);
        B200_LAUNCH(k_describe, grid, kDescWarps * 32, 0, st, imgs, rs, fs, d_pyr, g, h->dmaps, tma0_used, base, d_lvlres, d_lvlcnt, kps, desc, counts, out_cap);
    }
    if (h->profile) { B200_CUDA(cudaEventRecord(h->ev[4], st)); h->stage_valid = 1; h->stage_frames = n; }
    B200_CUDA(cudaGetLastError());
    h->last_imgs = imgs; h->last_row_stride = rs; h->last_frame_stride = fs; h->last_n = n; h->last_base = base;
    (void)w; (void)hh;
    return B200_OK;
}

// frames per pipeline stage of b200_frontend_host.  Small chunks multiply the latency-bound kernels (quadtree, contour walk, greedy resolve take
// as long for 32 frames as for 256), large ones expose the first upload: half the batch, at most kFrontendChunk
int frontend_chunk(int n) {
    static const int env_chunk = [] { const char* e = getenv("B200_FRONTEND_CHUNK"); return e ? atoi(e) : 0; }();
    return env_chunk > 0 ? env_chunk : (n <= 16 ? n : std::min(kFrontendChunk, std::max(32, (n + 1) / 2)));
}

// scratch_frames: how many frame slots of scratch the call needs (the whole batch for the device-pointer calls; two pipeline chunks for
// b200_frontend_host, whose even and odd chunks alternate between two slot regions)
int check_args(b200_orb_s* h, const void* imgs, int n, int w, int hh, long long rs, long long fs, int scratch_frames = -1) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    if (n < 0 || w < 0 || hh < 0) return fail(B200_EINVAL, "negative %s", "size");
    if ((scratch_frames < 0 ? n : scratch_frames) > h->max_batch || w > h->max_w || hh > h->max_h)
        return fail(B200_ECAPACITY, "batch/image larger than the handle's %s", "capacity");
    if (n > 0 && w > 0 && hh > 0) {
        if (!imgs) return fail(B200_EINVAL, "null %s", "image pointer");
        if (rs < w || (n > 1 && fs < rs * (hh - 1) + w)) return fail(B200_EINVAL, "bad %s", "strides");
    }
    return B200_OK;
}

}  // namespace

extern "C" {

int b200_orb_create(b200_orb_t* out, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                    int max_w, int max_h, int max_batch, int device) {
    if (!out) return fail(B200_EINVAL, "null %s", "out");
    *out = nullptr;
    if (nfeatures < 0 || nlevels < 1 || nlevels > kMaxLevels || !(scale_factor > 1.f) || ini_th < 1 || min_th < 1 || ini_th > 254 || min_th > ini_th ||
        max_w < 1 || max_h < 1 || max_batch < 1)
        return fail(B200_EINVAL, "bad %s parameters", "extractor");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    b200_orb_s* h = new (std::nothrow) b200_orb_s();
    if (!h) return B200_ENOMEM;
    h->device = device; h->nfeatures = nfeatures; h->scale = scale_factor; h->nlevels = nlevels; h->ini_th = ini_th; h->min_th = min_th;
    h->max_w = max_w; h->max_h = max_h; h->max_batch = max_batch; h->cur_w = h->cur_h = -1;
    make_tables(h);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(B200_ECUDA, "%s failed", "cudaStreamCreate"); }
    for (int i = 0; i < 5; i++) cudaEventCreate(&h->ev[i]);
    cudaEventCreateWithFlags(&h->ev_copy[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_copy[1], cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_aux2, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_ref, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_done[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_done[1], cudaEventDisableTiming);
    {   // the detector's kernels are latency bound (contour walks, per-candidate serial sections): they get the higher priority so that
        // they start early and the extractor's dense kernels fill the remaining SM slots (measured: 5.09 -> 4.95 ms per 256-frame step)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, hi);
        cudaStreamCreateWithPriority(&h->aux_stream2, cudaStreamNonBlocking, hi);
    }
    cudaEventCreateWithFlags(&h->ev_aux, cudaEventDisableTiming);
    if ((rc = upload_constants()) || (rc = set_geometry(h, max_w, max_h))) { b200_orb_destroy(h); return rc; }
    *out = h;
    return B200_OK;
}

int b200_orb_destroy(b200_orb_t h) {
    if (!h) return B200_OK;
    DeviceScope _ds; cudaSetDevice(h->device);
    cudaFree(h->d_pyr); cudaFree(h->d_tab); cudaFree(h->d_cells); cudaFree(h->d_slots); cudaFree(h->d_cellcnt);
    cudaFree(h->d_keysA); cudaFree(h->d_keysB); cudaFree(h->d_lvlres); cudaFree(h->d_lvlcnt); cudaFree(h->d_err);
    cudaFree(h->d_in); cudaFree(h->d_kps); cudaFree(h->d_desc); cudaFree(h->d_counts);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_out) cudaFreeHost(h->h_out);
    for (int i = 0; i < 5; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 2; i++) if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->down_stream) cudaStreamDestroy(h->down_stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->aux_stream2) cudaStreamDestroy(h->aux_stream2);
    if (h->ev_aux2) cudaEventDestroy(h->ev_aux2);
    if (h->ev_ref) cudaEventDestroy(h->ev_ref);
    for (int i = 0; i < 2; i++) if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->ev_aux) cudaEventDestroy(h->ev_aux);
    cudaFree(h->d_markers); cudaFree(h->d_mcounts); cudaFree(h->d_match); cudaFree(h->d_nmatch); cudaFree(h->d_refdesc); cudaFree(h->d_refkps);
    cudaFree(h->d_score); cudaFree(h->d_tiles);
    for (int b = 0; b < 7; b++) cudaFree(h->d_coll[b]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B200_OK;
}

int b200_orb_max_keypoints(b200_orb_t h) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    // independent of the image size: quota + 3 + 4*nIni per level with nIni <= 4 covers aspect ratios up to 4.5:1
    int cap = 0;
    for (int l = 0; l < h->nlevels; l++) cap += h->quota[l] + 3 + 16;
    return cap;
}

int b200_orb_get_level_info(b200_orb_t h, int* nlevels, float* sf, float* inv_sf, float* s2, float* inv_s2, int32_t* fpl) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    if (nlevels) *nlevels = h->nlevels;
    for (int l = 0; l < h->nlevels; l++) {
        if (sf) sf[l] = h->sf[l];
        if (inv_sf) inv_sf[l] = h->inv_sf[l];
        if (s2) s2[l] = h->sigma2[l];
        if (inv_s2) inv_s2[l] = h->inv_sigma2[l];
        if (fpl) fpl[l] = h->quota[l];
    }
    return B200_OK;
}

int b200_orb_extract(b200_orb_t h, const uint8_t* imgs, int n, int w, int hh, int64_t rs, int64_t fs,
                     b200_keypoint* kps, uint8_t* desc, int32_t* counts, void* stream) {
    int rc = check_args(h, imgs, n, w, hh, rs, fs);
    if (rc) return rc;
    if (!counts || !kps || !desc) return fail(B200_EINVAL, "null %s", "output pointer");
    DeviceScope _ds; if ((rc = use_device(h->device))) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (n == 0) return B200_OK;
    if (w == 0 || hh == 0) { B200_CUDA(cudaMemsetAsync(counts, 0, (size_t)n * 4, st)); return B200_OK; }   // empty image: no keypoints (ORBextractor.cc:1046)
    if ((rc = set_geometry(h, w, hh))) return rc;
    // the caller's buffers use cap = b200_orb_max_keypoints(); internally res_per_frame <= cap
    const int cap = b200_orb_max_keypoints(h);
    if (h->geom.res_per_frame > cap) return fail(B200_ECAPACITY, "aspect ratio beyond %s", "4.5:1");
    // output rows use the caller-visible capacity as pitch; the internal level-result block is indexed by res_per_frame
    // (running large batches as two halves on two streams, so that one half's quadtree overlaps the other's dense kernels, was measured twice:
    // 4.536 -> 4.505 ms per 256-frame C3 step, and no change at 2000 / 4000 features (C4 9.82 -> 9.85 ms, C5 121.5 -> 121.5 ms per step): not kept.
    // What does pay at C5 is overlapping whole sub-batches on two sets of handles, as b200_frontend_host does with its chunks: +32 %)
    return enqueue(h, imgs, n, w, hh, rs, fs, kps, desc, counts, cap, st);
}

namespace {
bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
}  // namespace

// Host-pointer calls: frames go up in chunks on a copy stream while the previous chunk is being processed on the
// compute stream (per-chunk events order the two); pinned caller buffers are used in place, pageable ones are
// staged through the handle's pinned buffers.  b200_frontend_host additionally runs the marker detector (on a
// third stream, concurrently with the extractor) and the brute-force matcher on every chunk that has just landed:
// one upload feeds all three stages, like Frame::Frame + Tracking do on one image in the reference
// (src/Frame.cc:91,142; src/Tracking.cc:917).
int b200_frontend_host(b200_orb_t h, b200_aruco_t aruco, const uint8_t* imgs, int n, int w, int hh, int64_t rs, int64_t fs,
                       b200_keypoint* kps, uint8_t* desc, int32_t* counts,
                       b200_marker* markers, int32_t* marker_counts,
                       const uint8_t* ref_desc, const b200_keypoint* ref_kps, int n_ref, float ratio, int check_ori,
                       int32_t* match_ref_idx, int32_t* n_matches) {
    // a batch larger than the handle's capacity streams through two alternating scratch regions of one pipeline chunk each
    int rc = check_args(h, imgs, n, w, hh, rs, fs, n <= (h ? h->max_batch : 0) ? n : 2 * frontend_chunk(n));
    if (rc) return rc;
    if (!counts || !kps || !desc) return fail(B200_EINVAL, "null %s", "output pointer");
    if (aruco && (!markers || !marker_counts)) return fail(B200_EINVAL, "null %s", "marker output pointer");
    const bool do_match = ref_desc != nullptr;
    if (do_match && (!ref_kps || !match_ref_idx || !n_matches || n_ref < 0)) return fail(B200_EINVAL, "bad %s", "matcher arguments");
    DeviceScope _ds; if ((rc = use_device(h->device))) return rc;
    if (n == 0) return B200_OK;
    if (w == 0 || hh == 0) {
        for (int i = 0; i < n; i++) { counts[i] = 0; if (aruco) marker_counts[i] = 0; if (do_match) n_matches[i] = 0; }
        return B200_OK;
    }
    const int cap = b200_orb_max_keypoints(h);
    const int mcap = aruco ? b200_aruco_max_markers(aruco) : 0;
    const size_t frame_bytes = (size_t)w * hh, in_bytes = frame_bytes * n;
    const size_t kp_bytes = (size_t)n * cap * sizeof(b200_keypoint), de_bytes = (size_t)n * cap * 32, ct_bytes = (size_t)n * 4;
    const size_t mk_bytes = (size_t)n * mcap * sizeof(b200_marker), ma_bytes = do_match ? (size_t)n * cap * 4 : 0;
    if ((rc = ensure(h->d_in, h->cap_in, in_bytes))) return rc;
    if ((rc = ensure(h->d_kps, h->cap_kps, kp_bytes))) return rc;
    if ((rc = ensure(h->d_desc, h->cap_desc, de_bytes))) return rc;
    if ((rc = ensure(h->d_counts, h->cap_counts, ct_bytes))) return rc;
    if (aruco) {
        if ((rc = ensure(h->d_markers, h->cap_markers, mk_bytes))) return rc;
        if ((rc = ensure(h->d_mcounts, h->cap_mcounts, ct_bytes))) return rc;
    }
    if (do_match) {
        if ((rc = ensure(h->d_match, h->cap_match, ma_bytes))) return rc;
        if ((rc = ensure(h->d_nmatch, h->cap_nmatch, ct_bytes))) return rc;
        if ((rc = ensure(h->d_refdesc, h->cap_refdesc, std::max<size_t>((size_t)n_ref * 32, 32)))) return rc;
        if ((rc = ensure(h->d_refkps, h->cap_refkps, std::max<size_t>((size_t)n_ref * sizeof(b200_keypoint), 32)))) return rc;
    }
    const bool in_pinned = is_pinned(imgs);
    const bool out_pinned = is_pinned(kps) && is_pinned(desc) && is_pinned(counts) && (!aruco || (is_pinned(markers) && is_pinned(marker_counts))) &&
                            (!do_match || (is_pinned(match_ref_idx) && is_pinned(n_matches)));
    if (!in_pinned && h->cap_hin < in_bytes) {
        if (h->h_in) cudaFreeHost(h->h_in);
        h->h_in = nullptr; h->cap_hin = 0;
        B200_CUDA(cudaMallocHost((void**)&h->h_in, in_bytes));
        h->cap_hin = in_bytes;
    }
    const size_t out_bytes = kp_bytes + de_bytes + ct_bytes + mk_bytes + ct_bytes + ma_bytes + ct_bytes;
    if (!out_pinned && h->cap_hout < out_bytes) {
        if (h->h_out) cudaFreeHost(h->h_out);
        h->h_out = nullptr; h->cap_hout = 0;
        B200_CUDA(cudaMallocHost((void**)&h->h_out, out_bytes));
        h->cap_hout = out_bytes;
    }
    if ((rc = set_geometry(h, w, hh))) return rc;
    if (h->geom.res_per_frame > cap) return fail(B200_ECAPACITY, "aspect ratio beyond %s", "4.5:1");
    cudaStream_t cs = h->copy_stream;
    cudaStream_t sts[2] = {h->stream, h->stream2}, ass[2] = {h->aux_stream, h->aux_stream2};
    cudaEvent_t ev_auxs[2] = {h->ev_aux, h->ev_aux2};
    if (do_match) {
        B200_CUDA(cudaMemcpyAsync(h->d_refdesc, ref_desc, (size_t)n_ref * 32, cudaMemcpyHostToDevice, sts[0]));
        B200_CUDA(cudaMemcpyAsync(h->d_refkps, ref_kps, (size_t)n_ref * sizeof(b200_keypoint), cudaMemcpyHostToDevice, sts[0]));
        B200_CUDA(cudaEventRecord(h->ev_ref, sts[0]));
        B200_CUDA(cudaStreamWaitEvent(sts[1], h->ev_ref, 0));
    }
    // host staging layout (pageable outputs): the seven result arrays back to back
    const size_t o_kps = 0, o_desc = o_kps + kp_bytes, o_cnt = o_desc + de_bytes, o_mk = o_cnt + ct_bytes, o_mc = o_mk + mk_bytes,
                 o_ma = o_mc + ct_bytes, o_nm = o_ma + ma_bytes;
    uint8_t* stage_out = h->h_out;
    cudaStream_t ds = h->down_stream;
    auto down = [&](void* user, size_t stage_ofs, const void* dev, size_t first_byte, size_t bytes) -> int {
        if (!bytes) return B200_OK;
        void* dst = out_pinned ? (void*)((uint8_t*)user + first_byte) : (void*)(stage_out + stage_ofs + first_byte);
        B200_CUDA(cudaMemcpyAsync(dst, (const uint8_t*)dev + first_byte, bytes, cudaMemcpyDeviceToHost, ds));
        return B200_OK;
    };
    static const int env_chunk = [] { const char* e = getenv("B200_FRONTEND_CHUNK"); return e ? atoi(e) : 0; }();
    const int chunk = frontend_chunk(n);
    const bool streamed = n > h->max_batch;           // scratch slots: the frame index itself, or two alternating chunk-sized regions
    // detector scratch slots: the whole batch when the handle is large enough, else two alternating chunk-sized regions, else one region
    // (then every detector call is serialised on the first auxiliary stream)
    const int acap = aruco ? b200_aruco_batch_capacity(aruco) : 0;
    const int amode = !aruco ? 0 : (acap >= n && !streamed) ? 2 : acap >= 2 * chunk ? 1 : 0;
    if (aruco && acap < std::min(chunk, n)) return fail(B200_ECAPACITY, "detector handle smaller than one pipeline chunk (%s frames)", "128");
    static const bool trace = getenv("B200_FRONTEND_TRACE") != nullptr;       // host time spent enqueueing, printed per call
    const auto t_enq0 = std::chrono::steady_clock::now();
    int ci = 0;
    // the first chunk is small so that compute starts after a short upload; the second one completes the regular grid
    static const int env_first = [] { const char* e = getenv("B200_FRONTEND_FIRST"); return e ? atoi(e) : 0; }();
    const int first = (env_first > 0 && env_first < chunk && n > env_first) ? env_first : (env_chunk <= 0 && chunk >= 64 && n > chunk) ? chunk / 4 : chunk;
    for (int f0 = 0, nf = 0; f0 < n; f0 += nf, ci++) {
        nf = std::min(ci == 0 ? first : (ci == 1 && first != chunk) ? chunk - first : chunk, n - f0);
        uint8_t* dst = h->d_in + (size_t)f0 * frame_bytes;
        if (in_pinned) {
            if (fs == rs * hh && rs == w)     // densely packed frames: one flat copy for the whole chunk
                B200_CUDA(cudaMemcpyAsync(dst, imgs + (size_t)f0 * fs, frame_bytes * nf, cudaMemcpyHostToDevice, cs));
            else if (fs == rs * hh)           // frames are back to back: one 2-D copy for the whole chunk
                B200_CUDA(cudaMemcpy2DAsync(dst, w, imgs + (size_t)f0 * fs, rs, w, (size_t)hh * nf, cudaMemcpyHostToDevice, cs));
            else
                for (int f = 0; f < nf; f++)
                    B200_CUDA(cudaMemcpy2DAsync(dst + (size_t)f * frame_bytes, w, imgs + (size_t)(f0 + f) * fs, rs, w, hh, cudaMemcpyHostToDevice, cs));
        } else {
            uint8_t* stage = h->h_in + (size_t)f0 * frame_bytes;
            for (int f = 0; f < nf; f++) {
                const uint8_t* src = imgs + (size_t)(f0 + f) * fs;
                if (rs == w) memcpy(stage + (size_t)f * frame_bytes, src, frame_bytes);
                else for (int y = 0; y < hh; y++) memcpy(stage + (size_t)f * frame_bytes + (size_t)y * w, src + (size_t)y * rs, w);
            }
            B200_CUDA(cudaMemcpyAsync(dst, stage, frame_bytes * nf, cudaMemcpyHostToDevice, cs));
        }
        // even and odd chunks use different stream sets and disjoint scratch slots (base = f0): the latency-bound tail of one
        // chunk (quadtree, contour following, greedy resolve) overlaps the dense kernels of the next
        cudaStream_t st = sts[ci & 1], as = ass[amode ? (ci & 1) : 0];
        const int abase = amode == 2 ? f0 : amode == 1 ? (ci & 1) * chunk : 0;
        cudaEvent_t ev = h->ev_copy[ci & 1];
        B200_CUDA(cudaEventRecord(ev, cs));
        B200_CUDA(cudaStreamWaitEvent(st, ev, 0));
        if (aruco) {
            B200_CUDA(cudaStreamWaitEvent(as, ev, 0));
            if ((rc = b200_aruco_detect_range(aruco, dst, nf, w, hh, w, (int64_t)frame_bytes, h->d_markers + (size_t)f0 * mcap, h->d_mcounts + f0, abase, as))) return rc;
            B200_CUDA(cudaEventRecord(ev_auxs[ci & 1], as));
            B200_CUDA(cudaStreamWaitEvent(ds, ev_auxs[ci & 1], 0));
        }
        const int sbase = streamed ? (ci & 1) * chunk : f0;                 // chunks ci and ci + 2 share a region and a stream set: stream order keeps them apart
        if ((rc = enqueue(h, dst, nf, w, hh, w, (long long)frame_bytes, h->d_kps + (size_t)f0 * cap, h->d_desc + (size_t)f0 * cap * 32,
                          h->d_counts + f0, cap, st, sbase)))
            return rc;
        if (do_match &&
            (rc = b200_match_bf_kp_range(h->d_refdesc, h->d_refkps, n_ref, h->d_desc + (size_t)f0 * cap * 32, h->d_kps + (size_t)f0 * cap, h->d_counts + f0, nf, cap,
                                         ratio, 50, check_ori, 30.0f / 360.0f, h->d_match + (size_t)f0 * cap, h->d_nmatch + f0, h->device, st, sbase,
                                         streamed ? 2 * chunk : n)))
            return rc;
        // this chunk's result slots go home on the download stream while the next chunk is being processed
        B200_CUDA(cudaEventRecord(h->ev_done[ci & 1], st));
        B200_CUDA(cudaStreamWaitEvent(ds, h->ev_done[ci & 1], 0));
        if ((rc = down(kps, o_kps, h->d_kps, (size_t)f0 * cap * sizeof(b200_keypoint), (size_t)nf * cap * sizeof(b200_keypoint))) ||
            (rc = down(desc, o_desc, h->d_desc, (size_t)f0 * cap * 32, (size_t)nf * cap * 32)) ||
            (rc = down(counts, o_cnt, h->d_counts, (size_t)f0 * 4, (size_t)nf * 4))) return rc;
        if (aruco && ((rc = down(markers, o_mk, h->d_markers, (size_t)f0 * mcap * sizeof(b200_marker), (size_t)nf * mcap * sizeof(b200_marker))) ||
                      (rc = down(marker_counts, o_mc, h->d_mcounts, (size_t)f0 * 4, (size_t)nf * 4)))) return rc;
        if (do_match && ((rc = down(match_ref_idx, o_ma, h->d_match, (size_t)f0 * cap * 4, (size_t)nf * cap * 4)) ||
                         (rc = down(n_matches, o_nm, h->d_nmatch, (size_t)f0 * 4, (size_t)nf * 4)))) return rc;
    }
    // (the debug taps b200_orb_get_pyramid / _get_candidates now refer to the LAST chunk)
    const auto t_enq1 = std::chrono::steady_clock::now();
    // the download stream has waited for every chunk of both stream sets: the error flags are final when it drains
    int err = 0;
    B200_CUDA(cudaMemcpyAsync(&err, h->d_err, 4, cudaMemcpyDeviceToHost, ds));
    B200_CUDA(cudaStreamSynchronize(ds));
    if (trace) {
        const auto t_enq2 = std::chrono::steady_clock::now();
        fprintf(stderr, "b200_frontend_host: n=%d chunks=%d enqueue %.3f ms, wait %.3f ms\n", n, ci, std::chrono::duration<double, std::milli>(t_enq1 - t_enq0).count(),
                std::chrono::duration<double, std::milli>(t_enq2 - t_enq1).count());
    }
    if (err) { cudaMemset(h->d_err, 0, 4); return fail(B200_ECAPACITY, "quadtree scratch overflow (%s)", err == 1 ? "node pool" : "result slots"); }
    if (aruco && (rc = b200_aruco_check(aruco, ds))) return rc;
    if (!out_pinned) {
        auto back = [&](void* user, size_t ofs, size_t bytes) { if (bytes) memcpy(user, stage_out + ofs, bytes); };
        back(kps, o_kps, kp_bytes); back(desc, o_desc, de_bytes); back(counts, o_cnt, ct_bytes);
        if (aruco) { back(markers, o_mk, mk_bytes); back(marker_counts, o_mc, ct_bytes); }
        if (do_match) { back(match_ref_idx, o_ma, ma_bytes); back(n_matches, o_nm, ct_bytes); }
    }
    h->fe_n = n; h->fe_mcap = mcap; h->fe_aruco = aruco ? 1 : 0; h->fe_match = do_match ? 1 : 0;
    return B200_OK;
}

// Sharded front end: every rank has run b200_frontend_host on its own contiguous block of n frames (same n, geometry and
// options on every rank).  The device-resident copies of those results are gathered at `root` in ONE NCCL group (b200_collate_gather: rank 0
// is the only consumer, nobody else receives a byte) and the root downloads them rank-major into host buffers [world][n][...] laid out like
// b200_frontend_host's outputs.  Non-root ranks pass NULL outputs.  Returns when the transfer has completed on the calling rank.
int b200_frontend_collate_host(b200_orb_t h, b200_collate_t c, int root, b200_keypoint* kps_all, uint8_t* desc_all, int32_t* counts_all,
                               b200_marker* markers_all, int32_t* marker_counts_all, int32_t* match_all, int32_t* n_matches_all) {
    if (!h || !c) return fail(B200_EINVAL, "null %s", "handle");
    if (h->fe_n <= 0) return fail(B200_EINVAL, "no b200_frontend_host call %s", "to collate");
    const int world = b200_collate_world(c), rank = b200_collate_rank(c);
    if (root < 0 || root >= world) return fail(B200_EINVAL, "bad %s", "root");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    const int n = h->fe_n, cap = b200_orb_max_keypoints(h), mcap = h->fe_mcap;
    const void* send[7] = {h->d_kps, h->d_desc, h->d_counts, h->d_markers, h->d_mcounts, h->d_match, h->d_nmatch};
    void* host[7] = {kps_all, desc_all, counts_all, markers_all, marker_counts_all, match_all, n_matches_all};
    int64_t bytes[7] = {(int64_t)n * cap * (int64_t)sizeof(b200_keypoint), (int64_t)n * cap * 32, (int64_t)n * 4,
                        h->fe_aruco ? (int64_t)n * mcap * (int64_t)sizeof(b200_marker) : 0, h->fe_aruco ? (int64_t)n * 4 : 0,
                        h->fe_match ? (int64_t)n * cap * 4 : 0, h->fe_match ? (int64_t)n * 4 : 0};
    void* recv[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (rank == root)
        for (int b = 0; b < 7; b++) {
            if (!bytes[b]) continue;
            if (!host[b]) return fail(B200_EINVAL, "the root needs every %s", "output buffer");
            if ((rc = ensure(h->d_coll[b], h->cap_coll[b], (size_t)bytes[b] * world))) return rc;
            recv[b] = h->d_coll[b];
        }
    if ((rc = b200_collate_gather(c, 7, send, recv, bytes, root, h->stream))) return rc;
    if (rank == root)
        for (int b = 0; b < 7; b++)
            if (bytes[b]) B200_CUDA(cudaMemcpyAsync(host[b], h->d_coll[b], (size_t)bytes[b] * world, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    return B200_OK;
}

int b200_orb_extract_host(b200_orb_t h, const uint8_t* imgs, int n, int w, int hh, int64_t rs, int64_t fs,
                          b200_keypoint* kps, uint8_t* desc, int32_t* counts) {
    return b200_frontend_host(h, nullptr, imgs, n, w, hh, rs, fs, kps, desc, counts, nullptr, nullptr, nullptr, nullptr, 0, 0.f, 0, nullptr, nullptr);
}

int b200_orb_set_profile(b200_orb_t h, int enable) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    h->profile = enable ? 1 : 0; h->stage_valid = 0;
    return B200_OK;
}

int b200_orb_get_stage_frames(b200_orb_t h) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    return h->stage_valid ? h->stage_frames : 0;
}

int b200_orb_get_stage_ms(b200_orb_t h, float* ms4) {
    if (!h || !ms4) return fail(B200_EINVAL, "null %s", "argument");
    if (!h->stage_valid) return fail(B200_EINVAL, "no profiled call %s", "yet");
    B200_CUDA(cudaEventSynchronize(h->ev[4]));
    for (int i = 0; i < 4; i++) B200_CUDA(cudaEventElapsedTime(&ms4[i], h->ev[i], h->ev[i + 1]));
    return B200_OK;
}

int b200_orb_get_pyramid(b200_orb_t h, int frame, int level, uint8_t* out, int* w_l, int* h_l) {
    if (!h || !out) return fail(B200_EINVAL, "null %s", "argument");
    if (!h->last_imgs || frame < 0 || frame >= h->last_n || level < 0 || level >= h->nlevels) return fail(B200_EINVAL, "no such %s", "frame/level");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    const LevelGeom& L = h->geom.L[level];
    std::vector<uint8_t> img((size_t)L.w * L.h);
    B200_CUDA(cudaStreamSynchronize(h->stream));
    if (level == 0)
        B200_CUDA(cudaMemcpy2D(img.data(), L.w, h->last_imgs + (size_t)frame * h->last_frame_stride, h->last_row_stride, L.w, L.h, cudaMemcpyDeviceToHost));
    else
        B200_CUDA(cudaMemcpy2D(img.data(), L.w, h->d_pyr + (size_t)(h->last_base + frame) * h->geom.pyr_frame_stride + L.offset, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    // the 19-px REFLECT_101 frame of mvImagePyramid (ORBextractor.cc:1113-1128) is never read on the mono path;
    // it is synthesised here on demand instead of being stored in HBM
    const int W = L.w + 2 * kEdge;
    auto refl = [](int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p; return p; };
    for (int y = 0; y < L.h + 2 * kEdge; y++) {
        const uint8_t* s = &img[(size_t)refl(y - kEdge, L.h) * L.w];
        for (int x = 0; x < W; x++) out[(size_t)y * W + x] = s[refl(x - kEdge, L.w)];
    }
    if (w_l) *w_l = L.w;
    if (h_l) *h_l = L.h;
    return B200_OK;
}

int b200_orb_get_candidates(b200_orb_t h, int frame, int level, int32_t* xys, int cap) {
    if (!h || !xys) return fail(B200_EINVAL, "null %s", "argument");
    if (!h->last_imgs || frame < 0 || frame >= h->last_n || level < 0 || level >= h->nlevels) return fail(B200_EINVAL, "no such %s", "frame/level");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    const OrbGeom& g = h->geom;
    const LevelGeom& L = g.L[level];
    B200_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<int> cnt(std::max(L.ncells, 1));
    if (L.ncells) B200_CUDA(cudaMemcpy(cnt.data(), h->d_cellcnt + (size_t)(h->last_base + frame) * g.total_cells + L.cell_base, (size_t)L.ncells * 4, cudaMemcpyDeviceToHost));
    int n = 0;
    std::vector<uint32_t> buf;
    for (int c = 0; c < L.ncells; c++) {
        const CellDesc& cd = h->cells[L.cell_base + c];
        const int k = std::min(cnt[c], cd.cap);
        if (!k) continue;
        buf.resize(k);
        B200_CUDA(cudaMemcpy(buf.data(), h->d_slots + (size_t)(h->last_base + frame) * g.slots_per_frame + cd.slot, (size_t)k * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < k; i++) {
            if (n >= cap) return fail(B200_ECAPACITY, "candidate buffer too %s", "small");
            xys[3 * n] = buf[i] & 0xfff; xys[3 * n + 1] = (buf[i] >> 12) & 0xfff; xys[3 * n + 2] = buf[i] >> 24;
            n++;
        }
    }
    return n;
}

}  // extern "C"
