// frame.cu -- B200-native keypoint undistortion and the 64 x 48 feature grid (sm_100a).
//
// Behavioural contract (reference src/Frame.cc): Frame::UndistortKeyPoints :357-388 (cv::undistortPoints(mat, mat, mK, mDistCoef,
// cv::Mat(), mK): 5 fixed-point iterations in double, re-projection through P = K, float result), Frame::ComputeImageBounds
// :418-447, Frame::AssignFeaturesToGrid :183-198 with Frame::PosInGrid :332-343 (round((x - mnMinX) * mfGridElementWidthInv)),
// Frame::GetFeaturesInArea :280-330.  Results are bit-identical to the CPU oracle (oracle/frame_oracle.cpp), which is pinned to
// cv2.undistortPoints golden vectors.  The grid of a frame is a CSR list: cell = ix * 48 + iy like mGrid[ix][iy], items in keypoint
// index order (the reference pushes them in that order), so GetFeaturesInArea visits candidates in the reference's order and its
// output feeds b200_match_candidates_host unchanged.
#include "common.h"
#include <math.h>

namespace b200 {

constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;      // include/Frame.h:40-41

struct FrameCam { double fx, fy, cx, cy, k[5]; int distorted; };

__device__ __forceinline__ void undistort_px(const FrameCam& c, float u, float v, float& xo, float& yo) {
    double x = ((double)u - c.cx) * (1.0 / c.fx), y = ((double)v - c.cy) * (1.0 / c.fy);
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dx = 2 * c.k[2] * x * y + c.k[3] * (r2 + 2 * x * x);
        const double dy = c.k[2] * (r2 + 2 * y * y) + 2 * c.k[3] * x * y;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    const double xx = c.fx * x + 0.0 * y + c.cx, yy = 0.0 * x + c.fy * y + c.cy, ww = 1. / (0.0 * x + 0.0 * y + 1.0);
    xo = (float)(xx * ww); yo = (float)(yy * ww);
}

// thread per keypoint slot; also works for the four image corners of ComputeImageBounds (n_batch = 1, cap = 4)
__global__ void __launch_bounds__(256)
k_undistort(const b200_keypoint* __restrict__ in, const int* __restrict__ counts, int n_batch, int cap, FrameCam c, b200_keypoint* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_batch * cap) return;
    const int f = i / cap;
    if (i - f * cap >= min(counts[f], cap)) return;
    b200_keypoint kp = in[i];
    if (c.distorted) undistort_px(c, kp.x, kp.y, kp.x, kp.y);
    out[i] = kp;
}

// One CTA per frame: cell of every keypoint, per-cell counts (smem atomics), exclusive scan, then a STABLE fill: the keypoints are
// walked in index order by one warp at a time per 32-slot chunk, ranks inside a chunk come from match-any ballots.
__global__ void __launch_bounds__(256)
k_grid(const b200_keypoint* __restrict__ un0, const int* __restrict__ counts, int cap, float min_x, float min_y, float inv_w, float inv_h,
       int* __restrict__ cell_start0, int* __restrict__ cell_items0) {
    __shared__ int s_cnt[kGridCells + 1];
    __shared__ int s_part[256];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = min(counts[f], cap);
    const b200_keypoint* un = un0 + (long long)f * cap;
    int* cell_start = cell_start0 + (long long)f * (kGridCells + 1);
    int* items = cell_items0 + (long long)f * cap;
    for (int c = tid; c <= kGridCells; c += 256) s_cnt[c] = 0;
    __syncthreads();
    auto cell_of = [&](int i) -> int {
        const b200_keypoint kp = un[i];
        const int px = (int)round((double)__fmul_rn(__fsub_rn(kp.x, min_x), inv_w)), py = (int)round((double)__fmul_rn(__fsub_rn(kp.y, min_y), inv_h));
        return (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) ? -1 : px * kGridRows + py;
    };
    for (int i = tid; i < n; i += 256) { const int c = cell_of(i); if (c >= 0) atomicAdd(&s_cnt[c], 1); }
    __syncthreads();
    // exclusive scan of 3072 counts: 12 per thread
    {
        int loc[12], sum = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) { loc[k] = s_cnt[tid * 12 + k]; sum += loc[k]; }
        s_part[tid] = sum;
        __syncthreads();
        if (tid < 32) {
            int run = 0;
            for (int b = 0; b < 256; b += 32) {
                const int v = s_part[b + lane];
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                s_part[b + lane] = run + incl - v;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) s_cnt[kGridCells] = run;
        }
        __syncthreads();
        int run = s_part[tid];
#pragma unroll
        for (int k = 0; k < 12; k++) { s_cnt[tid * 12 + k] = run; run += loc[k]; }
    }
    __syncthreads();
    for (int c = tid; c <= kGridCells; c += 256) cell_start[c] = s_cnt[c];
    __syncthreads();
    // stable fill by warp 0: chunks of 32 keypoints in index order; s_cnt[c] is the next free slot of cell c
    if (tid < 32) {
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const int c = i < n ? cell_of(i) : -1;
            const unsigned same = __match_any_sync(0xffffffffu, c);
            if (c >= 0) {
                const int rank = __popc(same & ((1u << lane) - 1));
                items[s_cnt[c] + rank] = i;
            }
            __syncwarp();
            if (c >= 0 && (same >> lane) == 1u) s_cnt[c] += __popc(same);         // the highest lane of each group advances the cursor
            __syncwarp();
        }
    }
}

// GetFeaturesInArea for a batch of queries of ONE frame: one warp per query, output lists in the reference's visit order
// (ix, iy, position in cell); out_ofs is filled by a host-side or device-side prefix over out_cnt when lists are packed, here every
// query owns a fixed-capacity row.
__global__ void __launch_bounds__(128)
k_features_in_area(const b200_keypoint* __restrict__ un, const int* __restrict__ cell_start, const int* __restrict__ items,
                   float min_x, float min_y, float inv_w, float inv_h, const float* __restrict__ q4, const int* __restrict__ qlev, int nq,
                   int* __restrict__ out, int* __restrict__ out_cnt, int row_cap) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float x = q4[3 * q], y = q4[3 * q + 1], r = q4[3 * q + 2];
    const int min_level = qlev[2 * q], max_level = qlev[2 * q + 1];
    int cnt = 0;
    const int c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, min_x), r), inv_w)));
    const int c1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, min_x), r), inv_w)));
    const int r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, min_y), r), inv_h)));
    const int r1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, min_y), r), inv_h)));
    if (c0 < kGridCols && c1 >= 0 && r0 < kGridRows && r1 >= 0) {
        const bool check = (min_level > 0) || (max_level >= 0);
        int* row = out + (long long)q * row_cap;
        for (int ix = c0; ix <= c1; ix++) {
            // the cells ix*48 + r0 .. ix*48 + r1 are contiguous in the CSR list: one run per column
            const int beg = cell_start[ix * kGridRows + r0], end = cell_start[ix * kGridRows + r1 + 1];
            for (int j0 = beg; j0 < end; j0 += 32) {
                const int j = j0 + lane;
                bool ok = false; int idx = 0;
                if (j < end) {
                    idx = items[j];
                    const b200_keypoint kp = un[idx];
                    ok = true;
                    if (check) { if (kp.octave < min_level) ok = false; if (max_level >= 0 && kp.octave > max_level) ok = false; }
                    ok = ok && fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r;
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) { const int pos = cnt + __popc(m & ((1u << lane) - 1)); if (pos < row_cap) row[pos] = idx; }
                cnt += __popc(m);
            }
        }
    }
    if (lane == 0) out_cnt[q] = cnt;
}

}  // namespace b200

using namespace b200;

namespace {
int make_cam(const float* cam9, FrameCam& c) {
    if (!cam9 || !(cam9[0] != 0) || !(cam9[1] != 0)) return fail(B200_EINVAL, "invalid camera %s", "parameters");
    c.fx = cam9[0]; c.fy = cam9[1]; c.cx = cam9[2]; c.cy = cam9[3];
    for (int i = 0; i < 5; i++) c.k[i] = cam9[4 + i];
    c.distorted = cam9[4] != 0.0f;                           // the reference only looks at k1 (Frame.cc:359)
    return B200_OK;
}
}  // namespace

extern "C" {

int b200_frame_undistort(const b200_keypoint* kps, const int32_t* counts, int n_batch, int cap, const float* cam9,
                         b200_keypoint* kps_un, int device, void* stream) {
    if (n_batch < 0 || cap < 0) return fail(B200_EINVAL, "negative %s", "size");
    FrameCam c;
    int rc = make_cam(cam9, c);
    if (rc) return rc;
    DeviceScope _ds; if ((rc = use_device(device))) return rc;
    if (n_batch == 0 || cap == 0) return B200_OK;
    if (!kps || !counts || !kps_un) return fail(B200_EINVAL, "null %s", "pointer");
    const long long total = (long long)n_batch * cap;
    B200_LAUNCH(k_undistort, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, kps, counts, n_batch, cap, c, kps_un);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

int b200_frame_undistort_points_host(const float* xy, int n, const float* cam9, float* xy_un, int device) {
    if (n < 0) return fail(B200_EINVAL, "negative %s", "size");
    FrameCam c;
    int rc = make_cam(cam9, c);
    if (rc) return rc;
    if (n == 0) return B200_OK;
    if (!xy || !xy_un) return fail(B200_EINVAL, "null %s", "pointer");
    if (!c.distorted) { if (xy_un != xy) memcpy(xy_un, xy, (size_t)n * 8); return B200_OK; }      // mDistCoef.at<float>(0) == 0: the points as they are
    DeviceScope _ds; if ((rc = use_device(device))) return rc;
    std::vector<b200_keypoint> h((size_t)n);
    memset(h.data(), 0, (size_t)n * sizeof(b200_keypoint));
    for (int i = 0; i < n; i++) { h[i].x = xy[2 * i]; h[i].y = xy[2 * i + 1]; }
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    DevBuf d, dc;
    if ((rc = d.upload(h.data(), (size_t)n * sizeof(b200_keypoint))) || (rc = dc.upload(&n, 4))) return rc;
    B200_LAUNCH(k_undistort, (unsigned)((n + 255) / 256), 256, 0, ts, (b200_keypoint*)d.p, (int*)dc.p, 1, n, c, (b200_keypoint*)d.p);
    B200_D2H(h.data(), d.p, (size_t)n * sizeof(b200_keypoint));
    for (int i = 0; i < n; i++) { xy_un[2 * i] = h[i].x; xy_un[2 * i + 1] = h[i].y; }
    return B200_OK;
}

int b200_frame_image_bounds(int width, int height, const float* cam9, float* bounds4, int device) {
    if (!bounds4 || width < 1 || height < 1) return fail(B200_EINVAL, "bad %s", "arguments");
    FrameCam c;
    int rc = make_cam(cam9, c);
    if (rc) return rc;
    if (!c.distorted) { bounds4[0] = 0.f; bounds4[1] = (float)width; bounds4[2] = 0.f; bounds4[3] = (float)height; return B200_OK; }
    DeviceScope _ds; if ((rc = use_device(device))) return rc;
    b200_keypoint h[4] = {};
    h[1].x = (float)width; h[2].y = (float)height; h[3].x = (float)width; h[3].y = (float)height;
    const int four = 4;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    DevBuf d, dc;
    if ((rc = d.upload(h, sizeof(h))) || (rc = dc.upload(&four, 4))) return rc;
    B200_LAUNCH(k_undistort, 1, 256, 0, ts, (b200_keypoint*)d.p, (int*)dc.p, 1, 4, c, (b200_keypoint*)d.p);
    B200_D2H(h, d.p, sizeof(h));
    bounds4[0] = fminf(h[0].x, h[2].x); bounds4[1] = fmaxf(h[1].x, h[3].x); bounds4[2] = fminf(h[0].y, h[1].y); bounds4[3] = fmaxf(h[2].y, h[3].y);
    return B200_OK;
}

int b200_frame_assign_grid(const b200_keypoint* kps_un, const int32_t* counts, int n_batch, int cap, const float* bounds4,
                           int32_t* cell_start, int32_t* cell_items, int device, void* stream) {
    if (n_batch < 0 || cap < 0) return fail(B200_EINVAL, "negative %s", "size");
    if (!bounds4 || !(bounds4[1] > bounds4[0]) || !(bounds4[3] > bounds4[2])) return fail(B200_EINVAL, "bad image %s", "bounds");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    if (n_batch == 0) return B200_OK;
    if (!kps_un || !counts || !cell_start || !cell_items) return fail(B200_EINVAL, "null %s", "pointer");
    const float inv_w = (float)kGridCols / (bounds4[1] - bounds4[0]), inv_h = (float)kGridRows / (bounds4[3] - bounds4[2]);       // Frame.cc:112-113
    B200_LAUNCH(k_grid, n_batch, 256, 0, (cudaStream_t)stream, kps_un, counts, cap, bounds4[0], bounds4[2], inv_w, inv_h, cell_start, cell_items);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

static int features_in_area_impl(const b200_keypoint* kps_un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                                 const float* queries_xyr, const int32_t* query_levels, int n_queries,
                                 int32_t* out_idx, int32_t* out_count, int row_cap, int device, void* stream, bool keyframe_origin) {
    if (n_queries < 0 || row_cap < 0) return fail(B200_EINVAL, "negative %s", "size");
    if (!bounds4 || !(bounds4[1] > bounds4[0]) || !(bounds4[3] > bounds4[2])) return fail(B200_EINVAL, "bad image %s", "bounds");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    if (n_queries == 0) return B200_OK;
    if (!kps_un || !cell_start || !cell_items || !queries_xyr || !query_levels || !out_idx || !out_count) return fail(B200_EINVAL, "null %s", "pointer");
    const float inv_w = (float)kGridCols / (bounds4[1] - bounds4[0]), inv_h = (float)kGridRows / (bounds4[3] - bounds4[2]);
    // KeyFrame keeps mnMinX / mnMinY as int (include/KeyFrame.h:211-214, truncated from the frame's float bounds) while the cell size stays the frame's
    // float mfGridElementWidthInv: its queries are laid over the grid from the truncated origin (src/KeyFrame.cc:677-689)
    const float min_x = keyframe_origin ? (float)(int)bounds4[0] : bounds4[0], min_y = keyframe_origin ? (float)(int)bounds4[2] : bounds4[2];
    B200_LAUNCH(k_features_in_area, (n_queries * 32 + 127) / 128, 128, 0, (cudaStream_t)stream, kps_un, cell_start, cell_items, min_x, min_y,
                inv_w, inv_h, queries_xyr, query_levels, n_queries, out_idx, out_count, row_cap);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

int b200_frame_features_in_area(const b200_keypoint* kps_un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                                const float* queries_xyr, const int32_t* query_levels, int n_queries,
                                int32_t* out_idx, int32_t* out_count, int row_cap, int device, void* stream) {
    return features_in_area_impl(kps_un, cell_start, cell_items, bounds4, queries_xyr, query_levels, n_queries, out_idx, out_count, row_cap, device, stream, false);
}

int b200_keyframe_features_in_area(const b200_keypoint* kps_un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                                   const float* queries_xyr, const int32_t* query_levels, int n_queries,
                                   int32_t* out_idx, int32_t* out_count, int row_cap, int device, void* stream) {
    return features_in_area_impl(kps_un, cell_start, cell_items, bounds4, queries_xyr, query_levels, n_queries, out_idx, out_count, row_cap, device, stream, true);
}

}  // extern "C"
