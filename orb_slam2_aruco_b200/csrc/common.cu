// common.cu -- process-wide helpers of libb200slam.so: error string, launch counter, device selection.
#include "common.h"
#include <math.h>

namespace b200 {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches(0);

int host_round(float v) { return (int)lrintf(v); }        // round-half-even under the default FP environment (cvRound)
int host_round_d(double v) { return (int)lrint(v); }

int use_device(int device) {
    // validated once per device (cudaGetDeviceProperties costs milliseconds); afterwards only cudaSetDevice
    static std::atomic<int> ok_mask[64];
    if (device >= 0 && device < 64 && ok_mask[device].load(std::memory_order_acquire)) {
        B200_CUDA(cudaSetDevice(device));
        return B200_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(B200_ENODEV, "no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= n) return fail(B200_EINVAL, "device index out of %s", "range");
    cudaDeviceProp p;
    B200_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return fail(B200_ENODEV, "device is not sm_100 (%s); kernels are built for sm_100a only", p.name);
    B200_CUDA(cudaSetDevice(device));
    if (device < 64) ok_mask[device].store(1, std::memory_order_release);
    return B200_OK;
}

namespace {
struct ThreadStreams {
    cudaStream_t s[64];
    ThreadStreams() { for (auto& x : s) x = nullptr; }
    ~ThreadStreams() { for (auto& x : s) if (x) cudaStreamDestroy(x); }       // at thread exit; errors (context already gone) are ignored
};
thread_local ThreadStreams t_streams;
}

cudaStream_t thread_stream(int device) {
    if (device < 0 || device >= 64) return cudaStreamPerThread;
    if (!t_streams.s[device]) {
        if (cudaStreamCreateWithFlags(&t_streams.s[device], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return cudaStreamPerThread; }
    }
    return t_streams.s[device];
}

thread_local cudaStream_t t_ts = nullptr;

int host_call_stream(int device, cudaStream_t* ts) {
    static std::atomic<int> pool_set[64];
    if (device >= 0 && device < 64 && !pool_set[device].load(std::memory_order_acquire)) {
        cudaMemPool_t pool;
        B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pool_set[device].store(1, std::memory_order_release);
    }
    *ts = t_ts = thread_stream(device);
    return B200_OK;
}

}  // namespace b200

extern "C" {
const char* b200_last_error(void) { return b200::g_err; }
int64_t b200_launch_count(void) { return (int64_t)b200::g_launches.load(); }
}
