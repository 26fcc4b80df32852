// oracle/ref_orb_wrap.cpp -- TEST INFRASTRUCTURE.
// C-ABI wrapper around the reference's OWN extractor (/root/reference/src/ORBextractor.cc, compiled
// unmodified where it lies against oracle/cvshim by oracle/Makefile -> oracle/_ref/libref_orb.so).
// It validates the restatement in orb_oracle.cpp and serves as the "reference" CPU baseline.
//
// Determinism note (hard part 1): ORBextractor.cc:684 sorts
// vector<pair<int,ExtractorNode*>>, so ties between nodes holding equally many keys are broken by
// heap address -- allocator dependent in the reference.  This wrapper replaces operator new inside
// the shared object by a per-thread monotonic arena, which turns "address order" into "creation
// order" (later node = greater address).  That is the canonical tie-break used by the restatement
// and by the CUDA path.  The reference source itself is untouched.
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>
#include "ORBextractor.h"

namespace {
struct Arena {
    char* base; size_t cap, used;
    Arena() : base(0), cap(0), used(0) {}
};
thread_local Arena g_arena;
const size_t kArenaBytes = (size_t)1 << 30;   // virtual; only touched pages are committed

inline bool in_arena(void* p) { return g_arena.base && (char*)p >= g_arena.base && (char*)p < g_arena.base + g_arena.cap; }
}

void* operator new(size_t n) {
    Arena& a = g_arena;
    if (!a.base) { a.base = (char*)malloc(kArenaBytes); a.cap = a.base ? kArenaBytes : 0; a.used = 0; }
    size_t need = (n + 15) & ~(size_t)15;
    if (a.used + need > a.cap) { fprintf(stderr, "ref_orb_wrap: arena exhausted\n"); abort(); }
    void* p = a.base + a.used;
    a.used += need;
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}

#ifndef REF_NATIVE_LIBM
// Canonicalisation 2 (hard part 2): ORBextractor.cc:113 calls cos/sin on a float, i.e.
// libm cosf/sinf, which are not correctly rounded and differ between glibc versions / FMA ifunc variants.
// Inside this shared object they are bound (version script => local symbols) to the float rounding of the
// double routine, the definition the restatement and the CUDA path use.  -DREF_NATIVE_LIBM builds the
// variant with the box's own libm (used only to REPORT the delta).
#include <math.h>
extern "C" float cosf(float x) { volatile double d = cos((double)x); return (float)d; }
extern "C" float sinf(float x) { volatile double d = sin((double)x); return (float)d; }
extern "C" void sincosf(float x, float* s, float* c) { volatile double a = sin((double)x), b = cos((double)x); *s = (float)a; *c = (float)b; }
#endif

extern "C" {

// One frame through ORB_SLAM2::ORBextractor::operator() (ORBextractor.cc:1043-1105).
// kps: [cap][7] floats = x, y, size, angle, response, octave, class_id.   desc: [cap][32].
// Returns the number of keypoints, or -1 if cap is too small.
int ref_orb_extract(const unsigned char* img, int w, int h, int stride,
                    int nfeatures, float scale, int nlevels, int ini_th, int min_th,
                    float* kps, unsigned char* desc, int cap) {
    g_arena.used = 0;
    int n;
    {
        ORB_SLAM2::ORBextractor ex(nfeatures, scale, nlevels, ini_th, min_th);
        cv::Mat im(h, w, CV_8UC1, (void*)img, (size_t)stride);
        std::vector<cv::KeyPoint> k;
        cv::Mat d;
        ex(im, cv::Mat(), k, d);
        n = (int)k.size();
        if (n > cap) n = -1;
        for (int i = 0; i < n; i++) {
            float* o = kps + (size_t)i * 7;
            o[0] = k[i].pt.x; o[1] = k[i].pt.y; o[2] = k[i].size; o[3] = k[i].angle; o[4] = k[i].response;
            o[5] = (float)k[i].octave; o[6] = (float)k[i].class_id;
            memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
        }
    }
    g_arena.used = 0;
    return n;
}

// Pyramid level l (with its 19-px REFLECT_101 border) as the reference builds it (ORBextractor.cc:1107-1132).
// out must hold (w_l+38)*(h_l+38) bytes; dims returned through wl/hl (inner size).
struct PyrProbe : ORB_SLAM2::ORBextractor {
    PyrProbe(int a, float b, int c, int d, int e) : ORB_SLAM2::ORBextractor(a, b, c, d, e) {}
    void run(cv::Mat im) { ComputePyramid(im); }
};
int ref_orb_pyramid_level(const unsigned char* img, int w, int h, int stride, float scale, int nlevels, int level,
                          unsigned char* out, int* wl, int* hl) {
    g_arena.used = 0;
    {
        PyrProbe ex(1000, scale, nlevels, 20, 7);
        cv::Mat im(h, w, CV_8UC1, (void*)img, (size_t)stride);
        ex.run(im);
        const cv::Mat& m = ex.mvImagePyramid[level];
        *wl = m.cols; *hl = m.rows;
        const unsigned char* base = m.data - 19 * m.step - 19;
        for (int y = 0; y < m.rows + 38; y++) memcpy(out + (size_t)y * (m.cols + 38), base + (size_t)y * m.step, m.cols + 38);
    }
    g_arena.used = 0;
    return 0;
}

}  // extern "C"

// Threaded batch driver for CPU-baseline timing: nthreads workers pull frames from a shared counter; each
// worker runs the reference extractor with its own arena.  kps [n][cap][7] floats, desc [n][cap][32].
// Plain pthreads + malloc here: anything allocated through this object's operator new would live in the
// calling thread's arena, which every ref_orb_extract call rewinds.
#include <pthread.h>
namespace {
struct BatchJob {
    const unsigned char* imgs; int n, w, h, row_stride; long frame_stride;
    int nfeatures; float scale; int nlevels, ini_th, min_th;
    float* kps; unsigned char* desc; int* counts; int cap;
    int next;
};
void* batch_worker(void* arg) {
    BatchJob* j = (BatchJob*)arg;
    for (;;) {
        int f = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (f >= j->n) break;
        j->counts[f] = ref_orb_extract(j->imgs + (size_t)f * j->frame_stride, j->w, j->h, j->row_stride, j->nfeatures, j->scale,
                                       j->nlevels, j->ini_th, j->min_th, j->kps + (size_t)f * j->cap * 7, j->desc + (size_t)f * j->cap * 32, j->cap);
    }
    if (g_arena.base) { free(g_arena.base); g_arena.base = 0; g_arena.cap = 0; }   // worker threads die after the call
    return 0;
}
}
extern "C" int ref_orb_extract_batch(const unsigned char* imgs, int n, int w, int h, int row_stride, long frame_stride,
                                     int nfeatures, float scale, int nlevels, int ini_th, int min_th,
                                     float* kps, unsigned char* desc, int* counts, int cap, int nthreads) {
    BatchJob job = {imgs, n, w, h, row_stride, frame_stride, nfeatures, scale, nlevels, ini_th, min_th, kps, desc, counts, cap, 0};
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], 0, batch_worker, &job);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
    free(th);
    return 0;
}
