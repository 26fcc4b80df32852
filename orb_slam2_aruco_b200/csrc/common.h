// common.h -- shared host/device helpers for libb200slam.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <new>
#include <vector>
#include "../../include/b200slam.h"

namespace b200 {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define B200_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            snprintf(b200::g_err, sizeof(b200::g_err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return B200_ECUDA;                                                                   \
        }                                                                                        \
    } while (0)

// count every kernel launch this library makes (bench.py reports it as gpu_launches)
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                                      \
    do {                                                                                         \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                              \
        b200::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        cudaError_t _le = cudaGetLastError();        /* an invalid configuration never reaches the device: report it here */ \
        if (_le != cudaSuccess) {                                                                \
            snprintf(b200::g_err, sizeof(b200::g_err), "%s:%d launch of %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(_le)); \
            return B200_ECUDA;                                                                   \
        }                                                                                        \
    } while (0)

// Select the device and make sure it is a Blackwell B200-class part; no CPU fallback exists.
int use_device(int device);

// Every C-ABI entry point selects the handle's device for its own duration only: the calling thread's current device is put back on
// return (a multi-GPU host, e.g. a torch process, keeps allocating where it was).
struct DeviceScope {
    int prev;
    DeviceScope() : prev(-1) { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceScope() { int cur = -1; if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); }
};

// The _host entry points are called from the reference's three threads (Tracking, LocalMapping, LoopClosing): each calling thread
// works on its own non-blocking stream per device (created on first use) and waits for that stream only, so a LocalMapping Fuse never
// stalls the tracking thread's extractor, and nothing runs on the legacy default stream.
cudaStream_t thread_stream(int device);

// stream-ordered scratch of the _host entry points: everything a call allocates, copies and launches goes to the calling thread's own stream
// (thread_stream), so concurrent calls from the reference's three threads neither serialise on the legacy default stream nor wait for each other
extern thread_local cudaStream_t t_ts;
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, t_ts); }
    int alloc(size_t n) { B200_CUDA(cudaMallocAsync(&p, n > 32 ? n : 32, t_ts)); return B200_OK; }
    int upload(const void* h, size_t n) { int rc = alloc(n); if (rc) return rc; if (n) B200_CUDA(cudaMemcpyAsync(p, h, n, cudaMemcpyHostToDevice, t_ts)); return B200_OK; }
};
// selects the calling thread's stream for `device` (and keeps the device's default memory pool from giving blocks back at every synchronisation)
int host_call_stream(int device, cudaStream_t* ts);
// device -> host copy that is complete when it returns, whatever kind of host memory the caller passed
#define B200_D2H(dst, src, n) do { B200_CUDA(cudaMemcpyAsync((dst), (src), (n), cudaMemcpyDeviceToHost, ts)); B200_CUDA(cudaStreamSynchronize(ts)); } while (0)

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// round-half-even of a float (cvRound semantics) on the host side
int host_round(float v);
int host_round_d(double v);

}  // namespace b200

// library-internal entry points (not part of include/b200slam.h): the scratch-slot aware forms used by b200_frontend_host,
// whose chunks run on alternating stream sets
extern "C" {
int b200_aruco_batch_capacity(b200_aruco_t h);
int b200_aruco_detect_range(b200_aruco_t h, const uint8_t* imgs, int n, int w, int hh, int64_t rs, int64_t fs,
                            b200_marker* markers, int32_t* counts, int base, void* stream);
int b200_match_bf_kp_range(const uint8_t* ref_desc, const b200_keypoint* ref_kps, int n_ref,
                           const uint8_t* frame_desc, const b200_keypoint* frame_kps, const int32_t* n_frame, int n_batch, int frame_cap,
                           float ratio, int th_low, int check_ori, float histo_factor,
                           int32_t* match_ref_idx, int32_t* n_matches, int device, void* stream, int base, int total);
}
