// collate.cu -- multi-GPU collation of the per-frame result slots: frames shard over the ranks with no
// data-path collective; the one exchange step brings every rank's fixed-size result buffers to the consumer rank.  Rank 0 is the only
// consumer, so this is a GATHER-TO-ROOT (grouped ncclSend / ncclRecv: non-root ranks send and receive nothing else), all buffers of a call
// inside ONE NCCL group (one fused transfer per call instead of one collective per buffer).
//
// NCCL is bound at run time (dlopen "libnccl.so.2": inside a torch process that is the library torch already loaded, in a plain C++ host
// the system one), so libb200slam.so keeps linking nothing but cudart and single-GPU hosts never need NCCL.  The communicator is the
// library's own: rank 0 asks b200_collate_unique_id for an id, the host distributes those 128 bytes however it likes (the reference has no
// transport of its own; bench.py and the tests use torch.distributed's store), every rank calls b200_collate_create.
#include "common.h"
#include <dlfcn.h>
#include <nccl.h>          // types and enums only; every function is looked up with dlsym

namespace b200 {
namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) return a;
#define B200_SYM(name) *(void**)(&a.name) = dlsym(a.lib, "nccl" #name)
        B200_SYM(GetUniqueId); B200_SYM(CommInitRank); B200_SYM(CommDestroy); B200_SYM(GroupStart); B200_SYM(GroupEnd);
        B200_SYM(Send); B200_SYM(Recv); B200_SYM(AllGather); B200_SYM(GetErrorString); B200_SYM(GetVersion);
#undef B200_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv && a.AllGather && a.GetErrorString;
        return a;
    }();
    return api;
}

int need_nccl() {
    if (!nccl().ok) return fail(B200_ENODEV, "NCCL is not loadable (%s): multi-GPU collation needs libnccl.so.2", dlerror() ? "dlopen failed" : "symbols missing");
    return B200_OK;
}

#define B200_NCCL(expr)                                                                          \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess) {                                                                 \
            snprintf(b200::g_err, sizeof(b200::g_err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, nccl().GetErrorString(_r)); \
            return B200_ECUDA;                                                                   \
        }                                                                                        \
    } while (0)

}  // namespace
}  // namespace b200

struct b200_collate_s {
    ncclComm_t comm;
    int rank, world, device;
    long long bytes_sent, bytes_received;        // NVLink payload of this rank so far (for the SCALE report)
};

using namespace b200;

extern "C" {

int b200_collate_unique_id(uint8_t* id128) {
    if (!id128) return fail(B200_EINVAL, "null %s", "pointer");
    int rc = need_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    ncclUniqueId id;
    B200_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return B200_OK;
}

int b200_collate_create(b200_collate_t* out, const uint8_t* id128, int rank, int world, int device) {
    if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return fail(B200_EINVAL, "bad %s", "communicator arguments");
    *out = nullptr;
    int rc = need_nccl();
    if (rc) return rc;
    DeviceScope _ds; rc = use_device(device);
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    B200_NCCL(nccl().CommInitRank(&comm, world, id, rank));
    b200_collate_s* h = new (std::nothrow) b200_collate_s{comm, rank, world, device, 0, 0};
    if (!h) { nccl().CommDestroy(comm); return fail(B200_ENOMEM, "out of %s", "memory"); }
    *out = h;
    return B200_OK;
}

int b200_collate_destroy(b200_collate_t h) {
    if (!h) return B200_OK;
    DeviceScope _ds; cudaSetDevice(h->device);
    if (nccl().ok) nccl().CommDestroy(h->comm);
    delete h;
    return B200_OK;
}

int b200_collate_rank(b200_collate_t h) { return h ? h->rank : -1; }
int b200_collate_world(b200_collate_t h) { return h ? h->world : 0; }

int b200_collate_nccl_version(void) {
    int v = 0;
    if (!nccl().ok || !nccl().GetVersion || nccl().GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

// Gather-to-root of n_buffers buffers in one NCCL group.  Every rank passes its own buffers send[b] (bytes[b] bytes each, the same sizes on
// every rank); the root additionally passes recv[b], rank-major (world x bytes[b]); its own block is a device-to-device copy on the same
// stream.  Only enqueues work on `stream`.
int b200_collate_gather(b200_collate_t h, int n_buffers, const void* const* send, void* const* recv, const int64_t* bytes, int root, void* stream) {
    if (!h || n_buffers < 0 || root < 0 || root >= h->world || (n_buffers > 0 && (!send || !bytes))) return fail(B200_EINVAL, "bad %s", "gather arguments");
    if (h->rank == root && n_buffers > 0 && !recv) return fail(B200_EINVAL, "the root needs %s", "receive buffers");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (h->world > 1) {
        B200_NCCL(nccl().GroupStart());
        for (int b = 0; b < n_buffers; b++) {
            if (bytes[b] <= 0) continue;
            if (h->rank == root) {
                for (int r = 0; r < h->world; r++)
                    if (r != root) {
                        B200_NCCL(nccl().Recv((uint8_t*)recv[b] + (size_t)r * bytes[b], (size_t)bytes[b], ncclUint8, r, h->comm, st));
                        h->bytes_received += bytes[b];
                    }
            } else {
                B200_NCCL(nccl().Send(send[b], (size_t)bytes[b], ncclUint8, root, h->comm, st));
                h->bytes_sent += bytes[b];
            }
        }
        B200_NCCL(nccl().GroupEnd());
    }
    if (h->rank == root)
        for (int b = 0; b < n_buffers; b++)
            if (bytes[b] > 0 && (uint8_t*)recv[b] + (size_t)root * bytes[b] != (const uint8_t*)send[b])
                B200_CUDA(cudaMemcpyAsync((uint8_t*)recv[b] + (size_t)root * bytes[b], send[b], (size_t)bytes[b], cudaMemcpyDeviceToDevice, st));
    return B200_OK;
}

// The all-gather form BASELINE.json's north_star names (every rank ends up with every slot), same buffer conventions, recv on every rank.
int b200_collate_allgather(b200_collate_t h, int n_buffers, const void* const* send, void* const* recv, const int64_t* bytes, void* stream) {
    if (!h || n_buffers < 0 || (n_buffers > 0 && (!send || !recv || !bytes))) return fail(B200_EINVAL, "bad %s", "all-gather arguments");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    B200_NCCL(nccl().GroupStart());
    for (int b = 0; b < n_buffers; b++)
        if (bytes[b] > 0) {
            B200_NCCL(nccl().AllGather(send[b], recv[b], (size_t)bytes[b], ncclUint8, h->comm, st));
            h->bytes_sent += bytes[b] * (h->world - 1); h->bytes_received += bytes[b] * (h->world - 1);
        }
    B200_NCCL(nccl().GroupEnd());
    return B200_OK;
}

int b200_collate_traffic(b200_collate_t h, int64_t* bytes_sent, int64_t* bytes_received) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    if (bytes_sent) *bytes_sent = h->bytes_sent;
    if (bytes_received) *bytes_received = h->bytes_received;
    return B200_OK;
}

}  // extern "C"
