// oracle/ref_aruco_wrap.cpp -- TEST INFRASTRUCTURE.  C entry point around the reference's OWN marker detector: Thirdparty/aruco/aruco/markerdetector.cpp,
// markerdetector_impl.cpp (the whole MarkerDetector_Impl::detect: pyramid, adaptive threshold, contours, quad test, candidate filter, warp, identification,
// de-duplication, CORNER_LINES refinement), marker.cpp, markerlabeler.cpp, dictionary.cpp, dictionary_based.cpp and debug.cpp, compiled unmodified from
// /root/reference against oracle/arucoshim into oracle/_ref/libref_aruco.so (oracle/Makefile).  The control flow is the reference's, statement for
// statement; the OpenCV primitives it calls (adaptiveThreshold, findContours, approxPolyDP, isContourConvex, resize, getPerspectiveTransform,
// warpPerspective, threshold, solve) are the cv2-pinned restatements of oracle/cvprim_aruco.h.  Configured exactly as src/Frame.cc:133-139 does.
// Pins oracle/aruco_oracle.cpp (tests/test_oracle_aruco_vs_ref.py, tests/golden/aruco_ref.npz).
// cameraparameters.cpp is compiled too (CameraParameters::setParams / resize are the reference's).  ippe.cpp (pose algebra: Rodrigues, inv, SVD) is not:
// detect() runs without camera parameters here, the pose is a separate row of the scope table (oracle/ippe_oracle.cpp, pinned to cv2's IPPE); the few
// members the linked sources name are defined below.
#include <atomic>
#include <cstring>
#include <thread>
#include "cameraparameters.h"
#include "ippe.h"
#include "markerdetector.h"
#include "markermap.h"
#include "../oracle.h"

namespace aruco {
void solvePnP(const std::vector<cv::Point3f>&, const std::vector<cv::Point2f>&, cv::InputArray, cv::InputArray, cv::Mat&, cv::Mat&) {
    throw std::runtime_error("aruco::solvePnP (ippe.cpp) is not part of the stand-in");
}
// Dictionary::createMarkerMap (dictionary.cpp) names these; it is never called
MarkerMap::MarkerMap() {}
Marker3DInfo::Marker3DInfo() {}
Marker3DInfo::Marker3DInfo(int _id) : id(_id) {}
}  // namespace aruco

namespace {
void configure(aruco::MarkerDetector& det, const char* dict_name) {     // src/Frame.cc:133-139
    det.setDictionary(std::string(dict_name));
    det.setDetectionMode(aruco::DetectionMode::DM_NORMAL);
    det.getParameters().setCornerRefinementMethod(aruco::CornerRefinementMethod::CORNER_LINES);
}
int run(aruco::MarkerDetector& det, const uint8_t* img, int w, int h, int stride, oracle_marker* out, int cap) {
    cv::Mat m(h, w, CV_8UC1, (void*)img, (size_t)stride);
    std::vector<aruco::Marker> ms = det.detect(m);
    int n = 0;
    for (size_t i = 0; i < ms.size(); i++, n++) {
        if (n >= cap) continue;
        out[n].id = ms[i].id;
        for (int k = 0; k < 4; k++) { out[n].xy[2 * k] = ms[i][k].x; out[n].xy[2 * k + 1] = ms[i][k].y; }
    }
    return n;
}
}  // namespace

extern "C" {

// aruco::MarkerDetector::detect(image) with the reference's settings (src/Frame.cc:133-139: setDictionary(name), DM_NORMAL, CORNER_LINES).
// out [cap]: id + 4 corners per marker in the order the reference returns them; returns the number of markers.
int ref_aruco_detect(const uint8_t* img, int w, int h, int stride, const char* dict_name, oracle_marker* out, int cap) {
    aruco::MarkerDetector det;
    configure(det, dict_name);
    return run(det, img, w, h, stride, out, cap);
}

// the same call, returning also what else the reference stores in every aruco::Marker (marker.h:57-59): dict_info (names [cap][32]) and contourPoints
// (contour_ofs [cap + 1] offsets into contour_xy [xy_cap][2]).  Pins the adapter's Marker::contourPoints / dict_info (tests/test_adapters_gpu.py).
int ref_aruco_detect_contours(const uint8_t* img, int w, int h, int stride, const char* dict_name, oracle_marker* out, int cap, char* names,
                              int32_t* contour_ofs, int32_t* contour_xy, int xy_cap) {
    aruco::MarkerDetector det;
    configure(det, dict_name);
    cv::Mat m(h, w, CV_8UC1, (void*)img, (size_t)stride);
    std::vector<aruco::Marker> ms = det.detect(m);
    int n = 0, pts = 0;
    contour_ofs[0] = 0;
    for (size_t i = 0; i < ms.size() && n < cap; i++, n++) {
        out[n].id = ms[i].id;
        for (int k = 0; k < 4; k++) { out[n].xy[2 * k] = ms[i][k].x; out[n].xy[2 * k + 1] = ms[i][k].y; }
        memset(names + 32 * n, 0, 32);
        strncpy(names + 32 * n, ms[i].dict_info.c_str(), 31);
        for (size_t j = 0; j < ms[i].contourPoints.size(); j++, pts++)
            if (pts < xy_cap) { contour_xy[2 * pts] = ms[i].contourPoints[j].x; contour_xy[2 * pts + 1] = ms[i].contourPoints[j].y; }
        contour_ofs[n + 1] = pts;
    }
    return n;
}

// batch driver for the CPU reference arm of bench.py: one detector object per worker thread (the reference keeps one static instance, src/Frame.cc:41),
// frames handed out dynamically.  Same argument list as oracle_aruco_detect_batch.
int ref_aruco_detect_batch(const uint8_t* imgs, int n, int w, int h, int row_stride, long frame_stride, const char* dict_name, oracle_marker* out,
                           int32_t* counts, int cap, int nthreads) {
    std::atomic<int> next(0);
    auto work = [&]() {
        aruco::MarkerDetector det;
        configure(det, dict_name);
        for (int f; (f = next.fetch_add(1)) < n;) counts[f] = run(det, imgs + (size_t)f * frame_stride, w, h, row_stride, out + (size_t)f * cap, cap);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return 0;
}

// CameraParameters(K, D, Size(cam_w, cam_h)) then resize(Size(w, h)) (cameraparameters.cpp:50-54, 75-95, 158-173) - what MarkerDetector::detect does to the
// camera whenever CamSize differs from the image.  cam4 = fx fy cx cy in / out.
int ref_camera_resize(const float* cam4, const float* dist5, int cam_w, int cam_h, int w, int h, float* cam4_out) {
    cv::Mat K = cv::Mat::zeros(3, 3, CV_32FC1), D(1, 5, CV_32FC1);
    K.at<float>(0, 0) = cam4[0]; K.at<float>(1, 1) = cam4[1]; K.at<float>(0, 2) = cam4[2]; K.at<float>(1, 2) = cam4[3]; K.at<float>(2, 2) = 1.f;
    for (int i = 0; i < 5; i++) D.at<float>(0, i) = dist5[i];
    aruco::CameraParameters cp(K, D, cv::Size(cam_w, cam_h));
    aruco::CameraParameters aux = cp;                              // as detect(): CameraParameters cp_aux = camParams; cp_aux.resize(input.size());
    aux.resize(cv::Size(w, h));
    cam4_out[0] = aux.CameraMatrix.at<float>(0, 0); cam4_out[1] = aux.CameraMatrix.at<float>(1, 1);
    cam4_out[2] = aux.CameraMatrix.at<float>(0, 2); cam4_out[3] = aux.CameraMatrix.at<float>(1, 2);
    return aux.CamSize.width == w && aux.CamSize.height == h && cp.CamSize.width == cam_w ? 0 : -1;
}

}  // extern "C"
