// C1 of BASELINE.json: ONE 640x480 gray frame per call through the source-compatible C++ adapters, exactly the two calls Frame::Frame makes
// (src/Frame.cc:91,142 on the mono_cvcam path, Examples/Monocular/mono_cvcam.cc:124-145): ORB_SLAM2::ORBextractor::operator() and
// aruco::MarkerDetector::detect(image, camParams, markerSize).  Prints one JSON object with the latency distribution of `iters` calls after
// `warm` untimed ones.  Built by orb_slam2_aruco_b200/build.py (g++, header-only adapters over libb200slam.so), run by bench.py.
#define B200SLAM_NO_OPENCV
#include "b200slam_adapters.hpp"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

static double pct(std::vector<double> v, double p) {
    std::sort(v.begin(), v.end());
    return v[(size_t)std::min<double>(v.size() - 1, p * (v.size() - 1) + 0.5)];
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s frames.raw w h nframes dict nfeatures [iters] [warm]\n", argv[0]); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]), nframes = atoi(argv[4]), nfeatures = atoi(argv[6]);
    const int iters = argc > 7 ? atoi(argv[7]) : 200, warm = argc > 8 ? atoi(argv[8]) : 20;
    std::vector<unsigned char> px((size_t)w * h * nframes);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(px.data(), 1, px.size(), f) != px.size()) { fprintf(stderr, "cannot read frames\n"); return 2; }
    fclose(f);
    try {
        ORB_SLAM2::ORBextractor extractor(nfeatures, 1.2f, 8, 20, 7);                                  // Tracking.cc:124
        aruco::MarkerDetector detector;
        detector.setDictionary(argv[5], 0.f);                                                          // Frame.cc:133
        detector.setDetectionMode(aruco::DetectionMode::DM_NORMAL);                                    // Frame.cc:134
        detector.getParameters().setCornerRefinementMethod(aruco::CornerRefinementMethod::CORNER_LINES);   // Frame.cc:135
        aruco::CameraParameters cam;
        const float distortion[5] = {0.2624f, -0.9531f, -0.0054f, 0.0026f, 1.1633f};
        cam.setParams(517.3f, 516.5f, 318.6f, 255.3f, distortion, 5, 1280, 720);                       // Frame.cc:132: CamSize hard-coded 1280 x 720
        std::vector<double> t_all, t_ext, t_det;
        long nk = 0, nm = 0;
        for (int it = 0; it < warm + iters; it++) {
            cv::Mat im(h, w, CV_8UC1, px.data() + (size_t)(it % nframes) * w * h);
            std::vector<cv::KeyPoint> keys;
            cv::Mat desc;
            const auto t0 = std::chrono::steady_clock::now();
            extractor(im, cv::Mat(), keys, desc);                                                      // Frame.cc:203
            const auto t1 = std::chrono::steady_clock::now();
            std::vector<aruco::Marker> markers = detector.detect(im, cam, 0.187f);                     // Frame.cc:142
            const auto t2 = std::chrono::steady_clock::now();
            if (it >= warm) {
                t_ext.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
                t_det.push_back(std::chrono::duration<double, std::milli>(t2 - t1).count());
                t_all.push_back(std::chrono::duration<double, std::milli>(t2 - t0).count());
                nk += (long)keys.size(); nm += (long)markers.size();
            }
        }
        // The same frame through ONE library call (b200_frontend_host, n = 1: extractor and detector run concurrently on two streams, one upload) and
        // the pose step: what a maintainer gets who replaces the two calls of Frame::Frame (src/Frame.cc:91,142) by one.
        std::vector<double> t_one;
        {
            b200_orb_t orb; b200_aruco_t det;
            if (b200_orb_create(&orb, nfeatures, 1.2f, 8, 20, 7, w, h, 1, 0) || b200_aruco_create(&det, argv[5], w, h, 1, 0)) throw std::runtime_error(b200_last_error());
            const int cap = b200_orb_max_keypoints(orb), mcap = b200_aruco_max_markers(det);
            std::vector<b200_keypoint> kps(cap); std::vector<uint8_t> de((size_t)cap * 32); std::vector<b200_marker> mk(mcap); std::vector<b200_marker_pose> poses(mcap);
            int32_t cnt = 0, mcnt = 0;
            const float cam9[9] = {517.3f * w / 1280.f, 516.5f * h / 720.f, 318.6f * w / 1280.f, 255.3f * h / 720.f, 0.2624f, -0.9531f, -0.0054f, 0.0026f, 1.1633f};
            for (int it = 0; it < warm + iters; it++) {
                const unsigned char* im = px.data() + (size_t)(it % nframes) * w * h;
                const auto t0 = std::chrono::steady_clock::now();
                if (b200_frontend_host(orb, det, im, 1, w, h, w, (int64_t)w * h, kps.data(), de.data(), &cnt, mk.data(), &mcnt, nullptr, nullptr, 0, 0.f, 0, nullptr, nullptr))
                    throw std::runtime_error(b200_last_error());
                if (mcnt > 0 && b200_aruco_pose_host(mk.data(), mcnt, 0.187f, cam9, poses.data(), 0)) throw std::runtime_error(b200_last_error());
                const auto t1 = std::chrono::steady_clock::now();
                if (it >= warm) t_one.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
            }
            b200_orb_destroy(orb); b200_aruco_destroy(det);
        }
        printf("{\"calls\": %d, \"warmup_calls\": %d, \"latency_ms_median\": %.4f, \"latency_ms_p10\": %.4f, \"latency_ms_p90\": %.4f, "
               "\"extract_ms_median\": %.4f, \"detect_ms_median\": %.4f, \"keypoints_per_frame\": %.1f, \"markers_per_frame\": %.2f, "
               "\"one_call_ms_median\": %.4f, \"gpu_launches\": %lld}\n",
               iters, warm, pct(t_all, 0.5), pct(t_all, 0.1), pct(t_all, 0.9), pct(t_ext, 0.5), pct(t_det, 0.5), (double)nk / iters, (double)nm / iters,
               pct(t_one, 0.5), (long long)b200_launch_count());
    } catch (const std::exception& e) {
        fprintf(stderr, "c1_latency: %s\n", e.what());
        return 1;
    }
    return 0;
}
