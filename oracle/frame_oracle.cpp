// oracle/frame_oracle.cpp -- TEST INFRASTRUCTURE (CPU checker), not product code.
//
// Restatement of the step between the extractor and every SearchByProjection of the reference:
//   Frame::UndistortKeyPoints      src/Frame.cc:357-388   (cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK), float in / float out)
//   Frame::ComputeImageBounds      src/Frame.cc:418-447
//   Frame::AssignFeaturesToGrid    src/Frame.cc:183-198,  Frame::PosInGrid src/Frame.cc:332-343  (64 x 48 grid, include/Frame.h:40-41)
//   Frame::GetFeaturesInArea       src/Frame.cc:280-330
// cv::undistortPoints (OpenCV, un-vendored; 4.13 semantics): 5 fixed-point iterations in double, then x' = fx*x + cx through the
// 3x3 product P*R, stored as float.  Pinned against cv2.undistortPoints golden vectors (tests/golden/frame.npz).
#include <math.h>
#include <string.h>
#include <vector>
#include "oracle.h"

namespace {
void undistort_px(const double* cam, float u, float v, float& xo, float& yo) {
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    const double* k = cam + 4;                  // k1 k2 p1 p2 k3
    double x = ((double)u - cx) * (1.0 / fx), y = ((double)v - cy) * (1.0 / fy);
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
        const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    // RR = P * I: xx = fx*x + 0*y + cx, yy = 0*x + fy*y + cy, ww = 1 / (0*x + 0*y + 1)
    const double xx = fx * x + 0.0 * y + cx, yy = 0.0 * x + fy * y + cy, ww = 1. / (0.0 * x + 0.0 * y + 1.0);
    xo = (float)(xx * ww); yo = (float)(yy * ww);
}
}  // namespace

extern "C" {

// Frame::UndistortKeyPoints: cam9 = fx fy cx cy k1 k2 p1 p2 k3 (float values widened); k1 == 0 -> plain copy (Frame.cc:359-363)
void oracle_undistort_keypoints(const oracle_keypoint* in, int n, const double* cam9, oracle_keypoint* out) {
    for (int i = 0; i < n; i++) {
        out[i] = in[i];
        if ((float)cam9[4] != 0.0f) undistort_px(cam9, in[i].x, in[i].y, out[i].x, out[i].y);
    }
}

// Frame::ComputeImageBounds: bounds4 = mnMinX mnMaxX mnMinY mnMaxY
void oracle_image_bounds(int w, int h, const double* cam9, float* bounds4) {
    if ((float)cam9[4] != 0.0f) {
        float x[4], y[4];
        const float px[4] = {0.f, (float)w, 0.f, (float)w}, py[4] = {0.f, 0.f, (float)h, (float)h};
        for (int i = 0; i < 4; i++) undistort_px(cam9, px[i], py[i], x[i], y[i]);
        bounds4[0] = fminf(x[0], x[2]); bounds4[1] = fmaxf(x[1], x[3]); bounds4[2] = fminf(y[0], y[1]); bounds4[3] = fmaxf(y[2], y[3]);
    } else { bounds4[0] = 0.f; bounds4[1] = (float)w; bounds4[2] = 0.f; bounds4[3] = (float)h; }
}

// AssignFeaturesToGrid + PosInGrid: cell = ix * 48 + iy (mGrid[ix][iy]); cell_start [64*48 + 1], cell_items [n] (indices in push order)
void oracle_assign_grid(const oracle_keypoint* un, int n, const float* bounds4, int32_t* cell_start, int32_t* cell_items) {
    const float inv_w = 64.f / (bounds4[1] - bounds4[0]), inv_h = 48.f / (bounds4[3] - bounds4[2]);
    std::vector<std::vector<int> > grid(64 * 48);
    for (int i = 0; i < n; i++) {
        const int px = (int)round((un[i].x - bounds4[0]) * inv_w), py = (int)round((un[i].y - bounds4[2]) * inv_h);       // int = round(float): double round of the float product
        if (px < 0 || px >= 64 || py < 0 || py >= 48) continue;
        grid[px * 48 + py].push_back(i);
    }
    int run = 0;
    for (int c = 0; c < 64 * 48; c++) { cell_start[c] = run; for (int v : grid[c]) cell_items[run++] = v; }
    cell_start[64 * 48] = run;
}

// GetFeaturesInArea: returns the number of indices written to out (visit order of the reference: ix, iy, position in cell)
static int features_in_area(const oracle_keypoint* un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4, float min_x, float min_y,
                            float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);

int oracle_features_in_area(const oracle_keypoint* un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                            float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    return features_in_area(un, cell_start, cell_items, bounds4, bounds4[0], bounds4[2], x, y, r, min_level, max_level, out, cap);
}

// KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:672-718): the window is placed from the keyframe's INT mnMinX / mnMinY (include/KeyFrame.h:211-214, truncated
// copies of the frame's float bounds) while the cell size is the frame's float mfGridElementWidthInv / HeightInv; no level arguments (pass -1, -1)
int oracle_keyframe_features_in_area(const oracle_keypoint* un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4,
                                     float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    return features_in_area(un, cell_start, cell_items, bounds4, (float)(int)bounds4[0], (float)(int)bounds4[2], x, y, r, min_level, max_level, out, cap);
}

}  // extern "C"

static int features_in_area(const oracle_keypoint* un, const int32_t* cell_start, const int32_t* cell_items, const float* bounds4, float min_x, float min_y,
                            float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    const float inv_w = 64.f / (bounds4[1] - bounds4[0]), inv_h = 48.f / (bounds4[3] - bounds4[2]);
    const int c0 = (int)floor((x - min_x - r) * inv_w) > 0 ? (int)floor((x - min_x - r) * inv_w) : 0;
    if (c0 >= 64) return 0;
    int c1 = (int)ceil((x - min_x + r) * inv_w); if (c1 > 63) c1 = 63;
    if (c1 < 0) return 0;
    const int r0 = (int)floor((y - min_y - r) * inv_h) > 0 ? (int)floor((y - min_y - r) * inv_h) : 0;
    if (r0 >= 48) return 0;
    int r1 = (int)ceil((y - min_y + r) * inv_h); if (r1 > 47) r1 = 47;
    if (r1 < 0) return 0;
    const bool check = (min_level > 0) || (max_level >= 0);
    int n = 0;
    for (int ix = c0; ix <= c1; ix++)
        for (int iy = r0; iy <= r1; iy++)
            for (int j = cell_start[ix * 48 + iy]; j < cell_start[ix * 48 + iy + 1]; j++) {
                const oracle_keypoint& kp = un[cell_items[j]];
                if (check) { if (kp.octave < min_level) continue; if (max_level >= 0 && kp.octave > max_level) continue; }
                const float dx = kp.x - x, dy = kp.y - y;
                if (fabs(dx) < r && fabs(dy) < r) { if (n < cap) out[n] = cell_items[j]; n++; }
            }
    return n;
}
