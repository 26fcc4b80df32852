// oracle/ref_ippe_wrap.cpp -- TEST INFRASTRUCTURE.  C entry point around the reference's OWN pose solver: Thirdparty/aruco/aruco/ippe.cpp (IPPE::PoseSolver and
// aruco::solvePnP, :72-1169), compiled unmodified from /root/reference on oracle/ippeshim into oracle/_ref/libref_ippe.so (oracle/Makefile).  The entry point
// does what aruco::Marker::calculateExtrinsics does (marker.cpp:322-343: get3DPoints, solvePnP -> IPPE::PoseSolver::solveGeneric, rvec / tvec narrowed to
// CV_32F) and what src/Frame.cc:155-177 does with the second solution (both reprojection errors).  Pins oracle/ippe_oracle.cpp and, through it, k_pose.
#include <cstring>
#include "ippe.h"

extern "C" {

// corners [8]: the marker's four image corners; cam9 = fx fy cx cy k1 k2 p1 p2 k3 (floats, as the reference's CV_32F camera matrix / distortion hold them).
// out14 = rvec1[3] tvec1[3] err1 rvec2[3] tvec2[3] err2 (doubles: the solver's CV_64F vectors; Marker::Rvec / Tvec are their float roundings).
int ref_ippe_marker_pose(const float* corners, float msize, const float* cam9, double* out14) {
    const float halfSize = msize / 2.f;                                // Marker::get3DPoints (marker.cpp:358-366)
    std::vector<cv::Point3f> objpoints = {cv::Point3f(-halfSize, halfSize, 0), cv::Point3f(halfSize, halfSize, 0), cv::Point3f(halfSize, -halfSize, 0),
                                          cv::Point3f(-halfSize, -halfSize, 0)};
    std::vector<cv::Point2f> imgpoints(4);
    for (int i = 0; i < 4; i++) imgpoints[i] = cv::Point2f(corners[2 * i], corners[2 * i + 1]);
    cv::Mat K = cv::Mat::zeros(3, 3, CV_32FC1), D(1, 5, CV_32FC1);
    K.at<float>(0, 0) = cam9[0]; K.at<float>(1, 1) = cam9[1]; K.at<float>(0, 2) = cam9[2]; K.at<float>(1, 2) = cam9[3]; K.at<float>(2, 2) = 1.f;
    for (int i = 0; i < 5; i++) D.at<float>(0, i) = cam9[4 + i];
    cv::Mat rvec1, tvec1, rvec2, tvec2;
    float err1 = 0, err2 = 0;
    IPPE::PoseSolver solver;                                           // aruco::solvePnP (ippe.cpp:89-100) with the second solution kept
    solver.solveGeneric(objpoints, imgpoints, K, D, rvec1, tvec1, err1, rvec2, tvec2, err2);
    for (int i = 0; i < 3; i++) {
        out14[i] = rvec1.at<double>(i); out14[3 + i] = tvec1.at<double>(i);
        out14[7 + i] = rvec2.at<double>(i); out14[10 + i] = tvec2.at<double>(i);
    }
    out14[6] = err1; out14[13] = err2;
    return 0;
}

// aruco::solvePnP(objPoints, imgPoints, K, D) -> the two 4 x 4 float [R | t] matrices with their errors (ippe.cpp:72-88), what src/Frame.cc:170 calls.
// T2 [2][16] row-major, errs [2].
int ref_ippe_solvepnp(const float* corners, float msize, const float* cam9, float* T2, double* errs) {
    const float halfSize = msize / 2.f;
    std::vector<cv::Point3f> objpoints = {cv::Point3f(-halfSize, halfSize, 0), cv::Point3f(halfSize, halfSize, 0), cv::Point3f(halfSize, -halfSize, 0),
                                          cv::Point3f(-halfSize, -halfSize, 0)};
    std::vector<cv::Point2f> imgpoints(4);
    for (int i = 0; i < 4; i++) imgpoints[i] = cv::Point2f(corners[2 * i], corners[2 * i + 1]);
    cv::Mat K = cv::Mat::zeros(3, 3, CV_32FC1), D(1, 5, CV_32FC1);
    K.at<float>(0, 0) = cam9[0]; K.at<float>(1, 1) = cam9[1]; K.at<float>(0, 2) = cam9[2]; K.at<float>(1, 2) = cam9[3]; K.at<float>(2, 2) = 1.f;
    for (int i = 0; i < 5; i++) D.at<float>(0, i) = cam9[4 + i];
    std::vector<std::pair<cv::Mat, double> > v = aruco::solvePnP(objpoints, imgpoints, K, D);
    for (int s = 0; s < 2; s++) { for (int i = 0; i < 16; i++) T2[16 * s + i] = v[s].first.at<float>(i / 4, i % 4); errs[s] = v[s].second; }
    return 0;
}

}  // extern "C"
