// oracle/aruco_oracle.cpp -- TEST INFRASTRUCTURE (CPU checker), not product code.
//
// Restatement of aruco::MarkerDetector::detect on the path the reference takes (src/Frame.cc:129-142:
// DM_NORMAL => THRES_ADAPTIVE, CORNER_LINES => minSize 0, defaults of markerdetector.h:162-200), following
// the de-obfuscated Thirdparty/aruco/aruco/markerdetector_impl.cpp + dictionary_based.cpp statement by
// statement (anchor lines as in SURVEY.md section 8a / Appendix B).  The reference has no tests of its own; this
// file is pinned by (0) the reference's OWN detector sources, compiled unmodified on oracle/arucoshim into
// oracle/_ref/libref_aruco.so (oracle/ref_aruco_wrap.cpp): same markers, ids and bit-identical refined corners on
// every frame tried (tests/test_oracle_aruco_vs_ref.py, tests/golden/aruco_ref.npz), and its decode stage and
// dictionary tables by oracle/_ref/libref_dict.so (tests/test_oracle_dict_vs_ref.py); (1) cv2 golden vectors for
// every OpenCV primitive (cvprim_aruco.h, which the compiled reference runs on as well), (2) a stage-by-stage
// cross-check against a python script that drives the real cv2 primitives in the same order
// (tests/golden/aruco_pipeline.npz) and (3) known answers: planted marker ids/corners.
//
// Canonical choices where the reference is not deterministic: std::sort of markers with equal ids
// (markerdetector_impl.cpp:8159) is taken as a stable sort in detection order.
#include "oracle.h"
#include "cvprim_aruco.h"
#include <map>
#include <string>
#include <thread>
#include <atomic>

using namespace cvprim;

namespace {

struct Dict { const char* name; int nbits, tau; std::map<uint64_t, int> code_id; };

std::vector<Dict>& dicts() {
    static std::vector<Dict> d;
    if (d.empty()) {
#define DICT_BEGIN(NAME, NBITS, TAU, N) { Dict cur; cur.name = #NAME; cur.nbits = NBITS; cur.tau = TAU; int id = 0; const uint64_t codes[] = {
#define C(x) x##ULL,
#define DICT_END(NAME) }; for (size_t i = 0; i < sizeof(codes) / sizeof(codes[0]); i++) cur.code_id.insert(std::make_pair(codes[i], id++)); d.push_back(cur); }
#include "../orb_slam2_aruco_b200/csrc/aruco_dicts.inc"
#undef DICT_BEGIN
#undef C
#undef DICT_END
    }
    return d;
}
const Dict* find_dict(const char* name) {
    for (auto& d : dicts()) if (std::string(d.name) == name) return &d;
    return nullptr;
}

struct Candidate { float c[8]; std::vector<Pt> contour; int id; };

// Marker::getArea, marker.cpp:405-416
float get_area(const float* c) {
    const float v01x = c[2] - c[0], v01y = c[3] - c[1], v03x = c[6] - c[0], v03y = c[7] - c[1];
    const float area1 = std::fabs(v01x * v03y - v01y * v03x);
    const float v21x = c[2] - c[4], v21y = c[3] - c[5], v23x = c[6] - c[4], v23y = c[7] - c[5];
    const float area2 = std::fabs(v21x * v23y - v21y * v23x);
    return (area2 + area1) / 2.f;
}
// perimeter(), markerdetector_impl.cpp:11123
int perimeter(const float* c) {
    int sum = 0;
    for (int i = 0; i < 4; i++) {
        const int j = (i + 1) % 4;
        const float dx = c[2 * i] - c[2 * j], dy = c[2 * i + 1] - c[2 * j + 1];
        sum += (int)std::sqrt(dx * dx + dy * dy);      // float sqrt, truncated
    }
    return sum;
}

// DictionaryBased::detect, dictionary_based.cpp:1062-2509 (error_correction_rate 0)
bool decode_patch(std::vector<u8>& patch, int psize, const Dict& dict, int& id, int& nrot) {
    const int level = otsu_level(patch.data(), psize, psize, psize);
    for (auto& v : patch) v = v > level ? 255 : 0;
    const int nb = (int)std::sqrt((double)dict.nbits), nsub = nb + 2;
    int nz[10][10] = {{0}}, tot[10][10] = {{0}};
    for (int y = 0; y < psize; y++) {
        const int my = (int)(float(nsub) * float(y) / float(psize));
        for (int x = 0; x < psize; x++) {
            const int mx = (int)(float(nsub) * float(x) / float(psize));
            if (patch[(size_t)y * psize + x] > 125) nz[my][mx]++;
            tot[my][mx]++;
        }
    }
    u8 bits[10][10];
    for (int y = 0; y < nsub; y++) for (int x = 0; x < nsub; x++) bits[y][x] = nz[y][x] > tot[y][x] / 2 ? 1 : 0;
    for (int y = 0; y < nsub; y++) {
        const int inc = (y == 0 || y == nsub - 1) ? 1 : nsub - 1;
        for (int x = 0; x < nsub; x += inc) if (bits[y][x] != 0) return false;     // the border must be black
    }
    u8 inner[8][8], tmp[8][8];
    for (int y = 0; y < nb; y++) for (int x = 0; x < nb; x++) inner[y][x] = bits[y + 1][x + 1];
    uint64_t ids[4];
    for (int r = 0; r < 4; r++) {
        uint64_t code = 0; int b = 0;
        for (int y = nb - 1; y >= 0; y--) for (int x = nb - 1; x >= 0; x--) code |= (uint64_t)inner[y][x] << b++;
        ids[r] = code;
        for (int i = 0; i < nb; i++) for (int j = 0; j < nb; j++) tmp[i][j] = inner[nb - j - 1][i];
        memcpy(inner, tmp, sizeof(inner));
    }
    if (ids[0] == 0) return false;                       // dictionary_based.cpp: an all-zero first code is "nothing"
    for (int r = 0; r < 4; r++) {
        auto it = dict.code_id.find(ids[r]);
        if (it != dict.code_id.end()) { nrot = r; id = it->second; return true; }
    }
    return false;
}

struct Taps {   // optional stage outputs for cross-checks
    std::vector<u8>* thres; std::vector<std::vector<Pt> >* contours; std::vector<float>* candidates; std::vector<u8>* patches;
    std::vector<float>* prerefine;
};

int detect(const u8* img, int w, int h, int stride, const Dict& dict, oracle_marker* out, int cap, Taps* taps) {
    // ---- threshold window (markerdetector_impl.cpp:3741-3867) and adaptive threshold (2984)
    int win = std::max(3, (int)(15 * float(w) / 1920.));
    if (win % 2 == 0) win++;
    const float too_near = (float)win;
    std::vector<u8> thres((size_t)w * h);
    adaptive_threshold_mean_inv(img, w, h, stride, thres.data(), w, win, 7);
    if (taps && taps->thres) *taps->thres = thres;
    // ---- contours -> convex quads (2707-3558)
    std::vector<std::vector<Pt> > contours;
    find_contours_list_none(thres.data(), w, h, w, contours);
    if (taps && taps->contours) *taps->contours = contours;
    const int min_size = (int)(3.5 * float(20));
    std::vector<Candidate> cand;
    std::vector<Pt> approx;
    for (size_t i = 0; i < contours.size(); i++) {
        if (min_size < (int)contours[i].size()) {
            approx_poly_dp_closed(contours[i], (double)contours[i].size() * 0.05, approx);
            if (approx.size() == 4 && is_contour_convex(approx)) {
                Candidate c;
                for (int k = 0; k < 4; k++) { c.c[2 * k] = (float)approx[k].x; c.c[2 * k + 1] = (float)approx[k].y; }
                c.contour = contours[i]; c.id = -1;
                cand.push_back(c);
            }
        }
    }
    // ---- prefilterCandidates (4349-5070)
    for (auto& c : cand) {
        const double dx1 = c.c[2] - c.c[0], dy1 = c.c[3] - c.c[1], dx2 = c.c[4] - c.c[0], dy2 = c.c[5] - c.c[1];
        const double o = (dx1 * dy2) - (dy1 * dx2);
        if (o < 0.0) { std::swap(c.c[2], c.c[6]); std::swap(c.c[3], c.c[7]); }
    }
    std::vector<std::pair<int, int> > near_pairs;
    for (size_t i = 0; i < cand.size(); i++)
        for (size_t j = i + 1; j < cand.size(); j++) {
            bool all = true;
            for (int k = 0; k < 4 && all; k++) {
                const float dx = cand[i].c[2 * k] - cand[j].c[2 * k], dy = cand[i].c[2 * k + 1] - cand[j].c[2 * k + 1];
                const float d = (float)std::sqrt((double)dx * dx + (double)dy * dy);     // cv::norm(Point2f) is double
                if (!(d < too_near)) all = false;
            }
            if (all) near_pairs.push_back(std::make_pair((int)i, (int)j));
        }
    std::vector<char> removed(cand.size(), 0);
    for (auto& pr : near_pairs) {
        if (perimeter(cand[pr.first].c) > perimeter(cand[pr.second].c)) removed[pr.second] = 1;
        else removed[pr.first] = 1;
    }
    const int bx = (int)(0.015f * float(w)), by = (int)(0.015f * float(h));
    for (size_t i = 0; i < cand.size(); i++)
        for (int k = 0; k < 4; k++) {
            const float x = cand[i].c[2 * k], y = cand[i].c[2 * k + 1];
            if (x < bx || y < by || x > w - bx || y > h - by) removed[i] = 1;
        }
    std::vector<Candidate> kept;
    for (size_t i = 0; i < cand.size(); i++) if (!removed[i]) kept.push_back(cand[i]);
    if (taps && taps->candidates) for (auto& c : kept) taps->candidates->insert(taps->candidates->end(), c.c, c.c + 8);
    // ---- image pyramid (1300-1466)
    const int nb = (int)std::sqrt((double)dict.nbits), nsub = nb + 2;        // setParams: sqrt(nbits)+2
    const int wsize = 5 * nsub;                                              // markerWarpPixSize * nSubdivisions
    struct Lvl { int w, h; std::vector<u8> img; };
    std::vector<Lvl> pyr(1);
    pyr[0].w = w; pyr[0].h = h; pyr[0].img.resize((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(&pyr[0].img[(size_t)y * w], img + (size_t)y * stride, w);
    {
        int nl = 1, cw = w, ch = h;
        while (cw > 2 * wsize) { cw /= 2; ch /= 2; nl++; }
        for (int l = 1; l < nl; l++) {
            Lvl L; L.w = pyr[l - 1].w / 2; L.h = pyr[l - 1].h / 2;
            if (L.w < 1 || L.h < 1) break;
            L.img.resize((size_t)L.w * L.h);
            resize_half(pyr[l - 1].img.data(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].w, L.img.data(), L.w, L.h, L.w);
            pyr.push_back(L);
        }
    }
    // ---- per-candidate decode (6482-6803)
    const float wsize2 = std::pow((float)wsize, 2.f);
    std::vector<Candidate> markers;
    std::vector<u8> patch((size_t)wsize * wsize);
    for (auto& c : kept) {
        size_t lvl = 0;
        for (size_t p = 1; p < pyr.size(); p++) {
            if (get_area(c.c) / std::pow(4, p) >= wsize2) lvl = p; else break;
        }
        float sc[8];
        const float scale = float(pyr[lvl].w) / float(w);
        for (int k = 0; k < 8; k++) sc[k] = c.c[k] * scale;
        const float dstq[8] = {0, 0, (float)(wsize - 1), 0, (float)(wsize - 1), (float)(wsize - 1), 0, (float)(wsize - 1)};
        double M[9];
        get_perspective_transform(sc, dstq, M);
        warp_perspective_linear(pyr[lvl].img.data(), pyr[lvl].w, pyr[lvl].h, pyr[lvl].w, patch.data(), wsize, wsize, wsize, M);
        if (taps && taps->patches) taps->patches->insert(taps->patches->end(), patch.begin(), patch.end());
        int id, nrot;
        if (decode_patch(patch, wsize, dict, id, nrot)) {
            Candidate m = c; m.id = id;
            float r[8];                                    // std::rotate(begin, begin + 4 - nRot, end)
            for (int k = 0; k < 4; k++) { const int s = (k + 4 - nrot) % 4; r[2 * k] = c.c[2 * s]; r[2 * k + 1] = c.c[2 * s + 1]; }
            memcpy(m.c, r, sizeof(r));
            markers.push_back(m);
        }
    }
    // ---- sort by id, remove duplicates keeping the larger perimeter (8159-8311)
    std::stable_sort(markers.begin(), markers.end(), [](const Candidate& a, const Candidate& b) { return a.id < b.id; });
    std::vector<char> rm(markers.size(), 0);
    for (int i = 0; i < (int)markers.size() - 1; i++)
        for (int j = i + 1; j < (int)markers.size() && !rm[i]; j++)
            if (markers[i].id == markers[j].id) {
                if (perimeter(markers[i].c) < perimeter(markers[j].c)) rm[i] = 1; else rm[j] = 1;
            }
    std::vector<Candidate> fin;
    for (size_t i = 0; i < markers.size(); i++) if (!rm[i]) fin.push_back(markers[i]);
    if (taps && taps->prerefine) for (auto& m : fin) taps->prerefine->insert(taps->prerefine->end(), m.c, m.c + 8);
    // ---- CORNER_LINES refinement (8647-8686, 8979-12049)
    for (auto& m : fin) {
        const std::vector<Pt>& cp = m.contour;
        const int n = (int)cp.size();
        int ci[4] = {-1, -1, -1, -1};
        float md[4] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};
        for (int j = 0; j < n; j++)
            for (int k = 0; k < 4; k++) {
                const float dx = cp[j].x - m.c[2 * k], dy = cp[j].y - m.c[2 * k + 1];
                const float d = dx * dx + dy * dy;
                if (d < md[k]) { ci[k] = j; md[k] = d; }
            }
        bool inverse;
        if ((ci[1] > ci[0]) && (ci[2] > ci[1] || ci[2] < ci[0])) inverse = false;
        else if (ci[2] > ci[1] && ci[2] < ci[0]) inverse = false;
        else inverse = true;
        const int inc = inverse ? -1 : 1;
        std::vector<float> side[4];
        bool ok = true;
        for (int l = 0; l < 4; l++) {
            const int stop = ci[(l + 1) % 4];
            long guard = 0;
            for (int j = ci[l]; j != stop; j += inc) {
                if (j == n && !inverse) j = 0;
                else if (j == 0 && inverse) j = n - 1;
                side[l].push_back((float)cp[j].x); side[l].push_back((float)cp[j].y);
                if (j == stop) break;
                if (++guard > 4L * n) { ok = false; break; }
            }
            if (side[l].size() < 4) ok = false;        // the reference would index an empty vector here (undefined)
        }
        if (!ok) continue;
        float line[4][3];
        for (int l = 0; l < 4; l++) {                     // interpolate2Dline, 11302-11841
            const int np = (int)side[l].size() / 2;
            float minx = side[l][0], maxx = minx, miny = side[l][1], maxy = miny;
            for (int i = 1; i < np; i++) {
                minx = std::min(minx, side[l][2 * i]); maxx = std::max(maxx, side[l][2 * i]);
                miny = std::min(miny, side[l][2 * i + 1]); maxy = std::max(maxy, side[l][2 * i + 1]);
            }
            std::vector<float> A((size_t)np * 2), B(np);
            float X[2];
            if (maxx - minx > maxy - miny) {
                for (int i = 0; i < np; i++) { A[2 * i] = side[l][2 * i]; A[2 * i + 1] = 1.f; B[i] = side[l][2 * i + 1]; }
                solve_svd_f32(A.data(), B.data(), np, 2, X);
                line[l][0] = X[0]; line[l][1] = -1.f; line[l][2] = X[1];
            } else {
                for (int i = 0; i < np; i++) { A[2 * i] = side[l][2 * i + 1]; A[2 * i + 1] = 1.f; B[i] = side[l][2 * i]; }
                solve_svd_f32(A.data(), B.data(), np, 2, X);
                line[l][0] = -1.f; line[l][1] = X[0]; line[l][2] = X[1];
            }
        }
        for (int i = 0; i < 4; i++) {                     // getCrossPoint(line[(i-1)%4 (unsigned)], line[i]), 11899-12049
            const float* l1 = line[(i + 3) % 4]; const float* l2 = line[i];
            const float A[4] = {l1[0], l1[1], l2[0], l2[1]}, B[2] = {-l1[2], -l2[2]};
            float X[2];
            solve_svd_f32(A, B, 2, 2, X);
            m.c[2 * i] = X[0]; m.c[2 * i + 1] = X[1];
        }
    }
    int nout = 0;
    for (auto& m : fin) {
        if (nout >= cap) return -1;
        out[nout].id = m.id;
        memcpy(out[nout].xy, m.c, sizeof(m.c));
        nout++;
    }
    return nout;
}

}  // namespace

extern "C" {

int oracle_aruco_detect(const uint8_t* img, int w, int h, int stride, const char* dict_name, oracle_marker* out, int cap) {
    const Dict* d = find_dict(dict_name);
    if (!d) return -2;
    return detect(img, w, h, stride, *d, out, cap, nullptr);
}

int oracle_aruco_detect_batch(const uint8_t* imgs, int n, int w, int h, int row_stride, long frame_stride, const char* dict_name,
                              oracle_marker* out, int32_t* counts, int cap, int nthreads) {
    const Dict* d = find_dict(dict_name);
    if (!d) return -2;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (int f; (f = next.fetch_add(1)) < n;)
            counts[f] = detect(imgs + (size_t)f * frame_stride, w, h, row_stride, *d, out + (size_t)f * cap, cap, nullptr);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return 0;
}

// stage taps for the cross-check against the cv2-driven pipeline: any output may be NULL.
//  thres [w*h]; contour_sizes [max_contours] (+ points [max_points][2]); candidates [max_cand][8]; patches [max_cand][ws*ws];
//  prerefine [cap][8] (corners of the final markers before CORNER_LINES)
int oracle_aruco_stages(const uint8_t* img, int w, int h, int stride, const char* dict_name,
                        uint8_t* thres, int32_t* contour_sizes, int32_t* contour_pts, int max_contours, int max_points, int32_t* n_contours,
                        float* candidates, uint8_t* patches, int max_cand, int32_t* n_cand,
                        float* prerefine, oracle_marker* out, int cap) {
    const Dict* d = find_dict(dict_name);
    if (!d) return -2;
    std::vector<u8> t; std::vector<std::vector<Pt> > c; std::vector<float> cd; std::vector<u8> pa; std::vector<float> pr;
    Taps taps = {&t, &c, &cd, &pa, &pr};
    const int n = detect(img, w, h, stride, *d, out, cap, &taps);
    if (thres) memcpy(thres, t.data(), t.size());
    if (n_contours) *n_contours = (int)c.size();
    if (contour_sizes) {
        size_t pts = 0;
        for (size_t i = 0; i < c.size() && (int)i < max_contours; i++) {
            contour_sizes[i] = (int)c[i].size();
            for (size_t k = 0; k < c[i].size(); k++, pts++)
                if (contour_pts && (int)pts < max_points) { contour_pts[2 * pts] = c[i][k].x; contour_pts[2 * pts + 1] = c[i][k].y; }
        }
    }
    const int nb = (int)std::sqrt((double)d->nbits), ws = 5 * (nb + 2);
    const int nc = (int)cd.size() / 8;
    if (n_cand) *n_cand = nc;
    if (candidates) memcpy(candidates, cd.data(), sizeof(float) * 8 * std::min(nc, max_cand));
    if (patches) memcpy(patches, pa.data(), (size_t)ws * ws * std::min(nc, max_cand));
    if (prerefine) memcpy(prerefine, pr.data(), sizeof(float) * pr.size());
    return n;
}

// the decode stage alone (DictionaryBased::detect, dictionary_based.cpp:1062-2509): one size x size canonical patch -> 1 and (id, nRotations), or 0.
// Pinned against the reference's own dictionary_based.cpp / dictionary.cpp (oracle/_ref/libref_dict.so, tests/test_oracle_dict_vs_ref.py).
int oracle_aruco_decode_patch(const uint8_t* patch, int size, const char* dict_name, int32_t* id, int32_t* nrot) {
    const Dict* d = find_dict(dict_name);
    if (!d) return -2;
    std::vector<u8> p(patch, patch + (size_t)size * size);
    int i = -1, r = -1;
    const bool ok = decode_patch(p, size, *d, i, r);
    *id = ok ? i : -1; *nrot = ok ? r : -1;
    return ok ? 1 : 0;
}
// the code table of a dictionary as the oracle and the product hold it (both include csrc/aruco_dicts.inc): codes [cap] by id; returns the highest id + 1
int oracle_dictionary_codes(const char* dict_name, uint64_t* codes, int cap, int32_t* nbits, int32_t* tau) {
    const Dict* d = find_dict(dict_name);
    if (!d) return -2;
    *nbits = d->nbits; *tau = d->tau;
    int n = 0;                                                    // ids run to the length of the list; an id whose code repeats an earlier one stays 0
    for (auto& kv : d->code_id) { if (kv.second < cap) codes[kv.second] = kv.first; if (kv.second + 1 > n) n = kv.second + 1; }
    return n;
}

// ---- primitive taps for the golden tests -------------------------------------------------------
void oracle_adaptive_threshold(const uint8_t* src, int w, int h, uint8_t* dst, int bs, int C) { adaptive_threshold_mean_inv(src, w, h, w, dst, w, bs, C); }
int oracle_find_contours(const uint8_t* img, int w, int h, int32_t* sizes, int32_t* pts, int max_contours, int max_points) {
    std::vector<std::vector<Pt> > c;
    find_contours_list_none(img, w, h, w, c);
    size_t np = 0;
    for (size_t i = 0; i < c.size(); i++) {
        if ((int)i < max_contours) sizes[i] = (int)c[i].size();
        for (size_t k = 0; k < c[i].size(); k++, np++) if ((int)np < max_points) { pts[2 * np] = c[i][k].x; pts[2 * np + 1] = c[i][k].y; }
    }
    return (int)c.size();
}
int oracle_approx_poly(const int32_t* pts, int n, double eps, int32_t* out, int* convex) {
    std::vector<Pt> s(n), d;
    for (int i = 0; i < n; i++) { s[i].x = pts[2 * i]; s[i].y = pts[2 * i + 1]; }
    approx_poly_dp_closed(s, eps, d);
    for (size_t i = 0; i < d.size(); i++) { out[2 * i] = d[i].x; out[2 * i + 1] = d[i].y; }
    if (convex) *convex = is_contour_convex(d) ? 1 : 0;
    return (int)d.size();
}
void oracle_resize_half(const uint8_t* src, int sw, int sh, uint8_t* dst) { resize_half(src, sw, sh, sw, dst, sw / 2, sh / 2, sw / 2); }
void oracle_perspective_transform(const float* src, const float* dst, double* M) { get_perspective_transform(src, dst, M); }
void oracle_warp_perspective(const uint8_t* src, int sw, int sh, uint8_t* dst, int ds, const double* M) { warp_perspective_linear(src, sw, sh, sw, dst, ds, ds, ds, M); }
int oracle_otsu(const uint8_t* img, int w, int h) { return otsu_level(img, w, h, w); }
void oracle_solve_svd(const float* A, const float* b, int m, int n, float* x) { solve_svd_f32(A, b, m, n, x); }

}  // extern "C"
