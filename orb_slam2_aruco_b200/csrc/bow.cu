// bow.cu -- B200-native DBoW2 descriptor -> word / node descent (sm_100a).
//
// Behavioural contract: DBoW2::TemplatedVocabulary<FORB>::transform(feature, word_id, weight, nid, levelsup)
// (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1259) as called for every descriptor of a frame by
// Frame::ComputeBoW (src/Frame.cc:348-355, levelsup = 4), with the node / word numbering of loadFromTextFile (:1338-1425).
// One warp per descriptor: the lanes take the (<= 32) children of the current node, compute their 256-bit Hamming
// distances and agree on the first minimum (`d < best_d` in child order), level by level until a leaf.
// Word ids, node ids and weights are bit-identical to the CPU oracle (oracle/bow_oracle.cpp).
#include "common.h"

namespace b200 {

struct VocNode { int child_beg, child_end, word_id, pad; };        // children = child_ids[child_beg .. child_end); leaf: child_beg == child_end

__global__ void __launch_bounds__(256)
k_voc_transform(const VocNode* __restrict__ nodes, const int* __restrict__ child_ids, const ulonglong4* __restrict__ node_desc,
                const double* __restrict__ node_weight, int nid_level, const ulonglong4* __restrict__ feat, int n,
                int* __restrict__ word_id, double* __restrict__ weight, int* __restrict__ node_id) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= n) return;
    const ulonglong4 q = feat[f];
    int cur = 0, nid = 0, level = 0;
    VocNode nd = nodes[0];
    while (nd.child_beg != nd.child_end) {
        ++level;
        unsigned best = 0xffffffffu;                              // dist << 16 | position in the child list: first minimum wins
        for (int c0 = nd.child_beg; c0 < nd.child_end; c0 += 32) {
            const int c = c0 + lane;
            if (c < nd.child_end) {
                const ulonglong4 d = node_desc[child_ids[c]];
                const unsigned dist = __popcll(q.x ^ d.x) + __popcll(q.y ^ d.y) + __popcll(q.z ^ d.z) + __popcll(q.w ^ d.w);
                best = min(best, (dist << 16) | (unsigned)(c - nd.child_beg));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        cur = child_ids[nd.child_beg + (int)(best & 0xffffu)];
        if (level == nid_level) nid = cur;
        nd = nodes[cur];
    }
    if (lane == 0) { word_id[f] = nd.word_id; weight[f] = node_weight[cur]; node_id[f] = nid; }
}

}  // namespace b200

using namespace b200;

struct b200_voc_s {
    int device, k, L, n_nodes, n_words;
    VocNode* d_nodes; int* d_children; ulonglong4* d_desc; double* d_weight;
};

extern "C" {

int b200_voc_create(b200_voc_t* out, int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* node_desc,
                    const double* node_weight, int device) {
    if (!out) return fail(B200_EINVAL, "null %s", "out");
    *out = nullptr;
    if (k < 1 || k > 20 || L < 1 || L > 10 || n_nodes < 2 || !parent || !is_leaf || !node_desc || !node_weight)      // loadFromTextFile's sanity check (:1359)
        return fail(B200_EINVAL, "bad %s", "vocabulary");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    // children in file order (m_nodes[pid].children.push_back(nid)), word ids in order of leaf appearance
    std::vector<int> cnt(n_nodes, 0), beg(n_nodes + 1, 0), fill(n_nodes, 0), children(n_nodes - 1);
    for (int i = 1; i < n_nodes; i++) {
        if (parent[i] < 0 || parent[i] >= i) return fail(B200_EINVAL, "node %s: parent must precede it", "order");
        cnt[parent[i]]++;
    }
    for (int i = 0; i < n_nodes; i++) beg[i + 1] = beg[i] + cnt[i];
    std::vector<VocNode> nodes(n_nodes);
    int nwords = 0;
    for (int i = 1; i < n_nodes; i++) children[beg[parent[i]] + fill[parent[i]]++] = i;
    for (int i = 0; i < n_nodes; i++) {
        nodes[i].child_beg = beg[i]; nodes[i].child_end = beg[i + 1]; nodes[i].word_id = 0; nodes[i].pad = 0;
        if (i > 0 && is_leaf[i]) {
            if (cnt[i]) return fail(B200_EINVAL, "a leaf with %s", "children");
            nodes[i].word_id = nwords++;
        } else if (cnt[i] == 0) return fail(B200_EINVAL, "an inner node without %s", "children");
    }
    b200_voc_s* h = new (std::nothrow) b200_voc_s();
    if (!h) return B200_ENOMEM;
    h->device = device; h->k = k; h->L = L; h->n_nodes = n_nodes; h->n_words = nwords;
    cudaError_t e = cudaMalloc((void**)&h->d_nodes, sizeof(VocNode) * n_nodes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_children, sizeof(int) * (n_nodes - 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_desc, 32 * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_weight, sizeof(double) * n_nodes);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_nodes, nodes.data(), sizeof(VocNode) * n_nodes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_children, children.data(), sizeof(int) * (n_nodes - 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_desc, node_desc, 32 * (size_t)n_nodes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_weight, node_weight, sizeof(double) * n_nodes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { b200_voc_destroy(h); return fail(B200_ECUDA, "vocabulary upload: %s", cudaGetErrorString(e)); }
    *out = h;
    return B200_OK;
}

int b200_voc_destroy(b200_voc_t h) {
    if (!h) return B200_OK;
    DeviceScope _ds; cudaSetDevice(h->device);
    cudaFree(h->d_nodes); cudaFree(h->d_children); cudaFree(h->d_desc); cudaFree(h->d_weight);
    delete h;
    return B200_OK;
}

int b200_voc_num_words(b200_voc_t h) { return h ? h->n_words : fail(B200_EINVAL, "null %s", "handle"); }

int b200_voc_transform(b200_voc_t h, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id, void* stream) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    if (n < 0) return fail(B200_EINVAL, "negative %s", "size");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    if (n == 0) return B200_OK;
    if (!desc || !word_id || !weight || !node_id) return fail(B200_EINVAL, "null %s", "pointer");
    if ((uintptr_t)desc & 31) return fail(B200_EINVAL, "descriptor array must be %s", "32-byte aligned");
    B200_LAUNCH(k_voc_transform, (n * 32 + 255) / 256, 256, 0, (cudaStream_t)stream, h->d_nodes, h->d_children, h->d_desc, h->d_weight, h->L - levelsup,
                (const ulonglong4*)desc, n, word_id, weight, node_id);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

int b200_voc_transform_host(b200_voc_t h, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id) {
    if (!h) return fail(B200_EINVAL, "null %s", "handle");
    if (n < 0) return fail(B200_EINVAL, "negative %s", "size");
    DeviceScope _ds; int rc = use_device(h->device);
    if (rc) return rc;
    if (n == 0) return B200_OK;
    if (!desc || !word_id || !weight || !node_id) return fail(B200_EINVAL, "null %s", "pointer");
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(h->device, &ts))) return rc;
    DevBuf dd, dw, dwt, dn;
    if ((rc = dd.upload(desc, 32 * (size_t)n)) || (rc = dw.alloc(4 * (size_t)n)) || (rc = dwt.alloc(8 * (size_t)n)) || (rc = dn.alloc(4 * (size_t)n))) return rc;
    if ((rc = b200_voc_transform(h, (const uint8_t*)dd.p, n, levelsup, (int32_t*)dw.p, (double*)dwt.p, (int32_t*)dn.p, ts))) return rc;
    B200_D2H(word_id, dw.p, 4 * (size_t)n);
    B200_D2H(weight, dwt.p, 8 * (size_t)n);
    B200_D2H(node_id, dn.p, 4 * (size_t)n);
    return B200_OK;
}

}  // extern "C"
