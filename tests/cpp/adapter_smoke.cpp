// Drives the source-compatible C++ adapters (include/b200slam_adapters.hpp) the way Frame::Frame / Tracking use the
// reference classes: extractor operator(), detector detect(), matcher SearchByBoW.  Reads a raw gray frame, writes a
// flat binary result that tests/test_adapters_gpu.py compares with the oracle.
#define B200SLAM_NO_OPENCV
#include "b200slam_adapters.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

// level 0 of the pyramid member: the input image sits at offset (19, 19) of its bordered buffer (src/ORBextractor.cc:1127)
static int l0_check(const cv::Mat& L0, const cv::Mat& im) {
    for (int y = 0; y < im.rows; y++) if (memcmp(L0.ptr(y + 19) + 19, im.ptr(y), im.cols)) return 1;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s in.raw w h dict out.bin\n", argv[0]); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    std::vector<unsigned char> px((size_t)w * h);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(px.data(), 1, px.size(), f) != px.size()) { fprintf(stderr, "cannot read frame\n"); return 2; }
    fclose(f);
    try {
        cv::Mat im(h, w, CV_8UC1, px.data());
        ORB_SLAM2::ORBextractor extractor(1000, 1.2f, 8, 20, 7);          // as in Tracking.cc:124
        std::vector<cv::KeyPoint> keys;
        cv::Mat desc;
        extractor(im, cv::Mat(), keys, desc);                             // Frame.cc:203
        aruco::MarkerDetector detector;
        detector.setDictionary(argv[4], 0.f);                             // Frame.cc:133
        detector.setDetectionMode(aruco::DetectionMode::DM_NORMAL);       // Frame.cc:134, spelled as there
        detector.getParameters().setCornerRefinementMethod(aruco::CornerRefinementMethod::CORNER_LINES);      // Frame.cc:135
        aruco::CameraParameters cam;                                      // Frame.cc:132: setParams(mK, mDistCoef, Size(1280,720))
        const float distortion[5] = {0.2624f, -0.9531f, -0.0054f, 0.0026f, 1.1633f};
        // CamSize: the image size unless given (argv[6], argv[7]); src/Frame.cc:132 hard-codes 1280 x 720, which makes detect() resize the camera
        cam.setParams(517.3f, 516.5f, 318.6f, 255.3f, distortion, 5, argc > 7 ? atoi(argv[6]) : w, argc > 7 ? atoi(argv[7]) : h);
        std::vector<aruco::Marker> markers = detector.detect(im, cam, 0.187f);   // Frame.cc:142 (mMarkerSize = 0.187, Frame.cc:131)
        ORB_SLAM2::ORBmatcher matcher(0.7f, true);                        // Tracking.cc:917
        std::vector<int> matches;
        const int nm = matcher.SearchByBoW(desc, keys, desc, keys, matches);
        ORB_SLAM2::ORBVocabulary voc;                                     // System.cc:80; no vocabulary file here: empty() like an unloaded one
        DBoW2::BowVector bowv; DBoW2::FeatureVector featv;
        voc.transform(desc, bowv, featv, 4);                             // Frame.cc:353 (no-op while empty)
        // SearchByBoW through FeatureVectors (ORBmatcher.cc:159-292): one node holding every feature must reproduce the brute-force answer
        DBoW2::FeatureVector one;
        for (int i = 0; i < desc.rows; i++) one[5].push_back((unsigned)i);
        std::vector<int> matches_fv, matches12;
        const int nm_fv = matcher.SearchByBoW(desc, keys, std::vector<bool>(keys.size(), true), one, desc, keys, one, matches_fv);
        if (nm_fv != nm || matches_fv != matches) { fprintf(stderr, "node-wise SearchByBoW differs from brute force\n"); return 1; }
        matcher.SearchByBoW_KF(desc, keys, std::vector<bool>(keys.size(), true), one, desc, keys, std::vector<bool>(keys.size(), true), one, matches12);
        if (matches12.size() != keys.size()) { fprintf(stderr, "SearchByBoW_KF size\n"); return 1; }
        // MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:271-331): {a, a, b} -> the duplicated descriptor (index 0) has median 0
        if (desc.rows >= 2) {
            std::vector<std::vector<cv::Mat> > obs(2);
            obs[0].push_back(cv::Mat(1, 32, CV_8U, desc.ptr(0))); obs[0].push_back(cv::Mat(1, 32, CV_8U, desc.ptr(0))); obs[0].push_back(cv::Mat(1, 32, CV_8U, desc.ptr(1)));
            std::vector<int> best;
            matcher.ComputeDistinctiveDescriptors(obs, best);
            if (best.size() != 2 || best[0] != 0 || best[1] != -1) { fprintf(stderr, "ComputeDistinctiveDescriptors: %d %d\n", best[0], best[1]); return 1; }
        }
        // the KeyFrame-side members on a keyframe matched against itself: every feature must find itself
        {
            const float bounds[4] = {0.f, (float)w, 0.f, (float)h};
            const std::vector<float> sf = extractor.GetScaleFactors(), s2 = extractor.GetScaleSigmaSquares(), is2 = extractor.GetInverseScaleSigmaSquares();
            std::vector<ORB_SLAM2::ORBmatcher::RadiusQuery> qs(keys.size());
            for (size_t i = 0; i < keys.size(); i++)
                qs[i] = ORB_SLAM2::ORBmatcher::RadiusQuery{keys[i].pt.x, keys[i].pt.y, 3.0f * sf[keys[i].octave], keys[i].octave, desc.ptr((int)i)};
            std::vector<int> bi, bd;
            matcher.SearchInRadius(keys, desc, bounds, qs, is2, 5.99, bi, bd);                  // Fuse (ORBmatcher.cc:906-955)
            for (size_t i = 0; i < keys.size(); i++)
                if (bi[i] != (int)i || bd[i] != 0) { fprintf(stderr, "SearchInRadius: feature %zu -> %d (dist %d)\n", i, bi[i], bd[i]); return 1; }
            std::vector<int> assign;
            const int nl = matcher.SearchByProjectionLoop(keys, desc, bounds, std::vector<bool>(keys.size(), false), qs, assign);      // ORBmatcher.cc:294-407
            if (nl != (int)keys.size()) { fprintf(stderr, "SearchByProjectionLoop: %d of %zu\n", nl, keys.size()); return 1; }
            for (size_t i = 0; i < keys.size(); i++) if (assign[i] != (int)i) { fprintf(stderr, "SearchByProjectionLoop: feature %zu -> %d\n", i, assign[i]); return 1; }
            // Fuse's loop body in one call (ORBmatcher.cc:846-955): every keypoint back-projected to depth 4 under the identity pose, with
            // mfMaxDistance chosen so that PredictScale answers the keypoint's own level, must come back to its own feature
            {
                const float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0}, Ow[3] = {0, 0, 0}, cam4[4] = {517.3f, 516.5f, 318.6f, 255.3f};
                std::vector<ORB_SLAM2::ORBmatcher::MapPointView> pts(keys.size());
                for (size_t i = 0; i < keys.size(); i++) {
                    ORB_SLAM2::ORBmatcher::MapPointView& p = pts[i];
                    p.pos[0] = (keys[i].pt.x - cam4[2]) / cam4[0] * 4.f; p.pos[1] = (keys[i].pt.y - cam4[3]) / cam4[1] * 4.f; p.pos[2] = 4.f;
                    const float dist = std::sqrt(p.pos[0] * p.pos[0] + p.pos[1] * p.pos[1] + p.pos[2] * p.pos[2]);
                    p.normal[0] = 0.f; p.normal[1] = 0.f; p.normal[2] = 1.f;
                    p.maxDistance = dist * std::pow(1.2f, (float)keys[i].octave - 0.5f); p.minDistance = 0.01f;
                    p.descriptor = desc.ptr((int)i); p.skip = (i % 7) == 3;
                }
                std::vector<bool> valid; std::vector<int> bi2, bd2;
                matcher.SearchPoints(keys, desc, bounds, R, t, Ow, nullptr, nullptr, cam4, pts, true, 3.0f, sf, is2, 5.99, valid, bi2, bd2);
                for (size_t i = 0; i < keys.size(); i++) {
                    const bool skipped = (i % 7) == 3;
                    if (valid[i] == skipped || (!skipped && (bi2[i] != (int)i || bd2[i] != 0)) || (skipped && bi2[i] != -1)) {
                        fprintf(stderr, "SearchPoints: point %zu valid %d -> %d (dist %d)\n", i, (int)valid[i], bi2[i], bd2[i]); return 1;
                    }
                }
                const std::vector<float> thr = ORB_SLAM2::ORBmatcher::PredictScaleThresholds(1.2f, 8);
                if (thr.size() != 7 || thr[0] != 1.0f || !(thr[1] > 1.19f && thr[1] < 1.21f)) { fprintf(stderr, "PredictScaleThresholds\n"); return 1; }
            }
            // F12 of a pure image shift s = (5, 3): the epipolar line of x1 is the line through x1 along s, which contains x2 = x1
            const float F12[9] = {0.f, 0.f, 0.03f, 0.f, 0.f, -0.05f, -0.03f, 0.05f, 0.f};
            std::vector<std::pair<size_t, size_t> > pairs;
            const int nt = matcher.SearchForTriangulation(keys, desc, std::vector<bool>(keys.size(), false), one, keys, desc, std::vector<bool>(keys.size(), false), one,
                                                          F12, -5000.f, -5000.f, sf, s2, pairs);                                         // ORBmatcher.cc:661-829
            if (nt != (int)keys.size() || pairs.size() != keys.size()) { fprintf(stderr, "SearchForTriangulation: %d of %zu\n", nt, keys.size()); return 1; }
            for (size_t i = 0; i < pairs.size(); i++) if (pairs[i].first != i || pairs[i].second != i) { fprintf(stderr, "SearchForTriangulation: pair %zu\n", i); return 1; }
        }
        const int dist = ORB_SLAM2::ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, desc.ptr(0)), cv::Mat(1, 32, CV_8U, desc.ptr(1)));
        FILE* o = fopen(argv[5], "wb");
        int hdr[4] = {(int)keys.size(), (int)markers.size(), nm, dist};
        fwrite(hdr, 4, 4, o);
        fwrite(keys.data(), sizeof(cv::KeyPoint), keys.size(), o);
        for (int i = 0; i < desc.rows; i++) fwrite(desc.ptr(i), 1, 32, o);
        for (auto& m : markers) { fwrite(&m.id, 4, 1, o); for (auto& p : m) { fwrite(&p.x, 4, 1, o); fwrite(&p.y, 4, 1, o); } }
        fwrite(matches.data(), 4, matches.size(), o);
        for (auto& m : markers) { fwrite(m.Rvec, 4, 3, o); fwrite(m.Tvec, 4, 3, o); fwrite(&m.err1, 4, 1, o); fwrite(&m.err2, 4, 1, o); fwrite(&m.ssize, 4, 1, o); }
        // src/Frame.cc:155-177: aruco::solvePnP per marker with the ORIGINAL camera, the ratio of the two reprojection errors is the quality test
        float cam_orig[9]; cam.cam9(cam_orig);
        std::vector<float> errs;
        for (auto& m : markers) {
            const cv::Point2f c4[4] = {m[0], m[1], m[2], m[3]};
            try {                                                        // judged by tests/test_zz_reference_replay_gpu.py, not here
                const std::vector<aruco::PnPSolution> v2pose = aruco::solvePnPSquare(0.187f, c4, cam_orig);
                errs.push_back((float)v2pose[0].error); errs.push_back((float)v2pose[1].error);
            } catch (const std::exception&) { errs.push_back(-1.f); errs.push_back(-1.f); }
        }
        {   // Frame::ComputeImageBounds / UndistortKeyPoints through the adapters: without distortion the image rectangle and the keypoints themselves
            const float pinhole[9] = {517.3f, 516.5f, 318.6f, 255.3f, 0, 0, 0, 0, 0};
            float x0, x1, y0, y1;
            ORB_SLAM2::ComputeImageBounds(w, h, pinhole, x0, x1, y0, y1);
            std::vector<cv::KeyPoint> un;
            ORB_SLAM2::UndistortKeyPoints(keys, pinhole, un);
            if (x0 != 0.f || x1 != (float)w || y0 != 0.f || y1 != (float)h || un.size() != keys.size() || (keys.size() && un[0].pt.x != keys[0].pt.x)) {
                fprintf(stderr, "ComputeImageBounds / UndistortKeyPoints without distortion\n"); return 1;
            }
        }
        std::vector<cv::Point2f> arucoUn;                                // Frame.cc:149 UndistortArucoCorners(); checked in tests/test_zz_reference_replay_gpu.py
        try { aruco::UndistortArucoCorners(markers, cam, arucoUn); } catch (const std::exception&) { arucoUn.clear(); }
        float cam_used[9]; cam.resized(w, h).cam9(cam_used);             // what detect() handed to the pose step (CameraParameters::resize, cameraparameters.cpp:158-173)
        fwrite(cam_used, 4, 9, o);
        fwrite(errs.data(), 4, errs.size(), o);
        for (auto& p : arucoUn) { fwrite(&p.x, 4, 1, o); fwrite(&p.y, 4, 1, o); }
        fclose(o);
        {   // aruco::Marker::contourPoints / dict_info (marker.h:57-59) -> <out>.contours: per marker a 32-byte name, the point count and the points
            FILE* oc = fopen((std::string(argv[5]) + ".contours").c_str(), "wb");
            for (auto& m : markers) {
                char name[32] = {0};
                strncpy(name, m.dict_info.c_str(), 31);
                const int len = (int)m.contourPoints.size();
                fwrite(name, 1, 32, oc); fwrite(&len, 4, 1, oc);
                for (auto& p : m.contourPoints) { fwrite(&p.x, 4, 1, oc); fwrite(&p.y, 4, 1, oc); }
            }
            fclose(oc);
        }
        {   // the public member mvImagePyramid (ORBextractor.h:85) against the accessor the pyramid tests pin to the oracle
            extractor.ComputePyramidMember();
            std::vector<int> ws, hs;
            const std::vector<std::vector<unsigned char> > pyr = extractor.ImagePyramid(&ws, &hs);
            if ((int)extractor.mvImagePyramid.size() != extractor.GetLevels()) { fprintf(stderr, "mvImagePyramid size\n"); return 1; }
            for (int l = 0; l < extractor.GetLevels(); l++) {
                const cv::Mat& L = extractor.mvImagePyramid[l];
                if (L.cols != ws[l] + 38 || L.rows != hs[l] + 38) { fprintf(stderr, "mvImagePyramid[%d] is %d x %d\n", l, L.cols, L.rows); return 1; }
                for (int y = 0; y < L.rows; y++)
                    if (memcmp(L.ptr(y), &pyr[l][(size_t)y * L.cols], L.cols)) { fprintf(stderr, "mvImagePyramid[%d] row %d\n", l, y); return 1; }
            }
            if (l0_check(extractor.mvImagePyramid[0], im)) { fprintf(stderr, "mvImagePyramid[0] interior != input image\n"); return 1; }
        }
        printf("levels=%d scale0=%g keys=%zu markers=%zu matches=%d\n", extractor.GetLevels(), extractor.GetScaleFactors()[1], keys.size(), markers.size(), nm);
    } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
