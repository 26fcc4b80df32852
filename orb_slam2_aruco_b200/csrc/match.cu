// match.cu -- B200-native ORB descriptor matching (sm_100a): 256-bit Hamming as an int8 GEMM on tcgen05 / TMEM (k_match_mma; 4 x __popcll in
// k_match_topk and the small kernels), best / second-best ratio test, 30-bin rotation-consistency histogram.
//
// Behavioural contract: ORB_SLAM2::ORBmatcher (reference src/ORBmatcher.cc): DescriptorDistance
// (1651-1667), SearchByBoW(KeyFrame,Frame) (159-292) with one all-inclusive vocabulary node (== brute
// force), ComputeThreeMaxima (1605-1646).  Match indices are bit-identical to the CPU oracle.
//
// The reference's loop is greedy: reference descriptor r only looks at frame descriptors that no earlier
// r has taken.  Split:
//   k_match_mma      CTA per (frame, 128 reference rows): all distances through tcgen05.mma.kind::i8, the K smallest (dist<<16|idx) keys per row
//   k_match_topk     the same lists by popcount, warp per (frame, r) (B200_MATCH_POPC=1; frame descriptors staged once per CTA as 4 u64 planes)
//   k_match_resolve  warp per frame: replays r = 0..n_ref-1 in order against a "taken" bitmap using the
//                    top-K lists (exact; a list that runs dry while a match is still possible triggers a
//                    warp-wide rescan of the whole row), then the rotation histogram filter.
#include "common.h"
#include <atomic>

namespace b200 {

constexpr int kTopK = 8;             // list length per reference row: long enough that the exact rescan stays rare even when most rows match
constexpr int kMatchWarps = 8;
constexpr int kRefPerCta = 64;

__device__ __forceinline__ int hamming256(const ulonglong4 a, const ulonglong4 b) {
    return __popcll(a.x ^ b.x) + __popcll(a.y ^ b.y) + __popcll(a.z ^ b.z) + __popcll(a.w ^ b.w);
}

// insert key into an ascending sorted K-list
__device__ __forceinline__ void topk_insert(uint32_t (&t)[kTopK], uint32_t key) {
    if (key < t[kTopK - 1]) {
        t[kTopK - 1] = key;
#pragma unroll
        for (int k = kTopK - 1; k > 0; k--)
            if (t[k] < t[k - 1]) { const uint32_t s = t[k - 1]; t[k - 1] = t[k]; t[k] = s; }
    }
}

__global__ void __launch_bounds__(kMatchWarps * 32)
k_match_topk(const ulonglong4* __restrict__ ref_desc, int n_ref,
             const ulonglong4* __restrict__ frame_desc, const int* __restrict__ n_frame, int frame_cap,
             uint32_t* __restrict__ topk, int dstar) {
    extern __shared__ unsigned long long s_planes[];      // [4][nf_pad]
    const int f = blockIdx.y;
    const int nf = min(n_frame[f], frame_cap);
    const int nf_pad = (frame_cap + 31) & ~31;
    const ulonglong4* fd = frame_desc + (long long)f * frame_cap;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        const ulonglong4 d = fd[i];
        s_planes[i] = d.x; s_planes[nf_pad + i] = d.y; s_planes[2 * nf_pad + i] = d.z; s_planes[3 * nf_pad + i] = d.w;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r_end = min(n_ref, (int)(blockIdx.x + 1) * kRefPerCta);
    for (int r = blockIdx.x * kRefPerCta + warp; r < r_end; r += kMatchWarps) {
        const ulonglong4 q = ref_desc[r];
        uint32_t t[kTopK];
#pragma unroll
        for (int k = 0; k < kTopK; k++) t[k] = 0xffffffffu;
        for (int i = lane; i < nf; i += 32) {
            // (ncu: XU pipe, where POPC runs, 90 % busy, ALU 50 %.  Carry-save trees that trade POPC for LOP3 were measured: 8 -> 4 POPC is
            // 15 % slower (ALU pipe saturates), 8 -> 5 POPC is a wash; skipping the last 64 bits when the first 192 already exceed dstar
            // (6 POPC for almost every pair) does not change the time either)
            const int d = __popcll(q.x ^ s_planes[i]) + __popcll(q.y ^ s_planes[nf_pad + i]) +
                          __popcll(q.z ^ s_planes[2 * nf_pad + i]) + __popcll(q.w ^ s_planes[3 * nf_pad + i]);
            // distances >= dstar are all equivalent for the accept rule (see match_bf_impl): only the few closer ones are listed
            if (d < dstar) topk_insert(t, ((uint32_t)d << 16) | (uint32_t)i);
        }
        // K rounds of warp arg-min over the lanes' list heads
        uint32_t* out = topk + ((long long)f * n_ref + r) * kTopK;
#pragma unroll
        for (int k = 0; k < kTopK; k++) {
            uint32_t m = t[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (t[0] == m && m != 0xffffffffu) {                      // keys are unique: exactly one lane pops its head
#pragma unroll
                for (int j = 0; j + 1 < kTopK; j++) t[j] = t[j + 1];
                t[kTopK - 1] = 0xffffffffu;
            }
            if (lane == 0) out[k] = m;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same top-K lists from the 5th-generation tensor cores.  Hamming(a, b) = |a| + |b| - 2 a.b over 256-bit descriptors is a dense binary
// contraction: with the bits expanded to 0 / 1 bytes, a.b is an int8 GEMM with exact int32 accumulators (tcgen05.mma kind::i8, M = 128
// reference rows x N = 128 frame descriptors x K = 256 per tile, 8 instructions of K = 32), and the POPC-bound inner loop above (8 POPC per
// pair on the 16-lane XU pipe) disappears.  One CTA = one frame x 128 reference rows, 256 threads:
//   * operands live in shared memory in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices; 128 bytes to the
//     next K chunk, 2048 bytes to the next 8 rows); the expansion writes them directly (nibble * 0x00204081 & 0x01010101 = 4 bytes)
//   * two TMEM accumulator buffers of 128 columns and two B tiles: the MMAs of tile t + 1 are issued (one elected thread, tcgen05.commit
//     on an mbarrier) before the epilogue of tile t starts
//   * epilogue: thread (row, half) reads its 64 int32 dot products with tcgen05.ld.32x32b, forms |b| - 2 a.b, compares with dstar - |a|
//     and inserts the few survivors into its sorted K-list; the two halves of a row are merged through shared memory
// The lists are bit-identical to k_match_topk's (the K smallest keys dist << 16 | index below dstar), so k_match_resolve is unchanged.
// ------------------------------------------------------------------------------------------------
constexpr int kMmaThreads = 256, kMmaM = 128, kMmaN = 128;
constexpr int kMmaLbo = 128, kMmaSbo = 2048, kMmaTile = 32768;            // bytes: next K chunk, next 8-row group, one expanded 128 x 256 tile
constexpr int kMmaSmem = 3 * kMmaTile + 4096 + 2 * kMmaN * 4 + kMmaM * 4 + kMmaM * kTopK * 4 + 64;

__device__ __forceinline__ uint32_t mma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 128 raw descriptors (32 bytes each, shared memory) -> 128 x 256 bytes of 0 / 1 in the UMMA K-major no-swizzle layout
__device__ __forceinline__ void mma_expand(const unsigned char* __restrict__ raw, unsigned char* __restrict__ dst, int tid) {
    for (int idx = tid; idx < kMmaM * 16; idx += kMmaThreads) {
        const int r = idx & (kMmaM - 1), c = idx >> 7;                  // descriptor, 16-bit chunk
        const unsigned bits = reinterpret_cast<const unsigned short*>(raw)[r * 16 + c];
        uint4 v;
        v.x = ((bits & 15u) * 0x00204081u) & 0x01010101u;
        v.y = (((bits >> 4) & 15u) * 0x00204081u) & 0x01010101u;
        v.z = (((bits >> 8) & 15u) * 0x00204081u) & 0x01010101u;
        v.w = ((bits >> 12) * 0x00204081u) & 0x01010101u;
        *reinterpret_cast<uint4*>(dst + (r >> 3) * kMmaSbo + c * kMmaLbo + (r & 7) * 16) = v;
    }
}

__device__ __forceinline__ unsigned long long mma_desc(uint32_t saddr) {
    // UMMA shared-memory descriptor: start address, leading (K) and stride (M / N) byte offsets in 16-byte units, version 1 (sm_100), no swizzle
    const uint32_t lo = ((saddr >> 4) & 0x3fffu) | ((uint32_t)(kMmaLbo >> 4) << 16);
    const uint32_t hi = (uint32_t)(kMmaSbo >> 4) | (1u << 14);
    return ((unsigned long long)hi << 32) | lo;
}

__global__ void __launch_bounds__(kMmaThreads)
k_match_mma(const ulonglong4* __restrict__ ref_desc, int n_ref, const ulonglong4* __restrict__ frame_desc, const int* __restrict__ n_frame,
            int frame_cap, uint32_t* __restrict__ topk, int dstar) {
    extern __shared__ __align__(1024) unsigned char mm_raw[];
    unsigned char* sA = mm_raw;
    unsigned char* sB = mm_raw + kMmaTile;                               // two tiles
    unsigned char* sRaw = mm_raw + 3 * kMmaTile;
    int* s_nb = reinterpret_cast<int*>(sRaw + 4096);                      // [2][128]
    int* s_na = s_nb + 2 * kMmaN;
    uint32_t* s_merge = reinterpret_cast<uint32_t*>(s_na + kMmaM);       // [128][kTopK]
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_merge + kMmaM * kTopK);     // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.y, m0 = blockIdx.x * kMmaM;
    const int nf = min(n_frame[f], frame_cap);
    const int ntiles = (nf + kMmaN - 1) / kMmaN;
    const uint4* fd = reinterpret_cast<const uint4*>(frame_desc + (long long)f * frame_cap);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(mma_smem_u32(s_tmem)), "r"(2 * kMmaN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mma_smem_u32(&s_bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mma_smem_u32(&s_bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- A: this CTA's 128 reference rows (rows past n_ref are zero), expanded once
    {
        const uint4* rd = reinterpret_cast<const uint4*>(ref_desc + m0);
        const int row = tid >> 1;
        reinterpret_cast<uint4*>(sRaw)[tid] = (m0 + row < n_ref) ? rd[tid] : make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    if (tid < kMmaM) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(sRaw) + tid * 8;
        int c = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) c += __popc(w[k]);
        s_na[tid] = c;
    }
    mma_expand(sRaw, sA, tid);
    __syncthreads();                                                    // sRaw is free again; s_tmem is visible
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    const uint32_t idesc = (2u << 4) | ((uint32_t)(kMmaN >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);     // D = S32, A = B = u8 K-major, N, M
    const unsigned long long descA = mma_desc(mma_smem_u32(sA));

    // stage + expand frame tile t into B buffer t & 1 (descriptors past nf: zero rows, |b| = 10000 so that they never qualify)
    auto load_b = [&](int t) {
        const int i0 = t * kMmaN, row = tid >> 1;
        reinterpret_cast<uint4*>(sRaw)[tid] = (i0 + row < nf) ? fd[(long long)i0 * 2 + tid] : make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (tid < kMmaN) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(sRaw) + tid * 8;
            int c = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) c += __popc(w[k]);
            s_nb[(t & 1) * kMmaN + tid] = (i0 + tid < nf) ? c : 10000;
        }
        mma_expand(sRaw, sB + (t & 1) * kMmaTile, tid);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's async proxy
    };
    auto issue = [&](int t) {                                           // one thread: 8 x (128 x 128 x 32) into TMEM buffer t & 1
        const unsigned long long descB = mma_desc(mma_smem_u32(sB + (t & 1) * kMmaTile));
        const uint32_t d_tmem = tmem + (uint32_t)((t & 1) * kMmaN);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned long long a = descA + (unsigned long long)(k * 2 * kMmaLbo >> 4), b = descB + (unsigned long long)(k * 2 * kMmaLbo >> 4);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                         :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(k), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mma_smem_u32(&s_bar[t & 1])) : "memory");
    };

    uint32_t tk[kTopK];
#pragma unroll
    for (int k = 0; k < kTopK; k++) tk[k] = 0xffffffffu;
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;
    const int thr = dstar - s_na[row], na = s_na[row];
    if (ntiles > 0) {
        load_b(0);
        __syncthreads();
        if (tid == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); issue(0); }
    }
    for (int t = 0; t < ntiles; t++) {
        if (t + 1 < ntiles) load_b(t + 1);                              // its buffers were last used by tile t - 1, whose MMAs and epilogue are done
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t + 1 < ntiles && tid == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); issue(t + 1); }
        {   // wait for tile t's accumulators
            const uint32_t bar = mma_smem_u32(&s_bar[t & 1]), parity = (uint32_t)((t >> 1) & 1);
            asm volatile("{\n.reg .pred p;\nMM_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra MM_DONE;\nbra MM_WAIT;\nMM_DONE:\n}"
                         :: "r"(bar), "r"(parity) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int* nb = s_nb + (t & 1) * kMmaN + half * 64;
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((t & 1) * kMmaN + half * 64);
#pragma unroll
        for (int part = 0; part < 2; part++) {
            uint32_t r[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                         "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                           "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                           "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                           "=r"(r[31])
                         : "r"(taddr + (uint32_t)(32 * part)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int ibase = t * kMmaN + half * 64 + 32 * part;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int4 n4 = *reinterpret_cast<const int4*>(nb + 32 * part + j);
                const int v0 = n4.x - 2 * (int)r[j], v1 = n4.y - 2 * (int)r[j + 1], v2 = n4.z - 2 * (int)r[j + 2], v3 = n4.w - 2 * (int)r[j + 3];
                if (v0 < thr) topk_insert(tk, ((uint32_t)(v0 + na) << 16) | (uint32_t)(ibase + j));
                if (v1 < thr) topk_insert(tk, ((uint32_t)(v1 + na) << 16) | (uint32_t)(ibase + j + 1));
                if (v2 < thr) topk_insert(tk, ((uint32_t)(v2 + na) << 16) | (uint32_t)(ibase + j + 2));
                if (v3 < thr) topk_insert(tk, ((uint32_t)(v3 + na) << 16) | (uint32_t)(ibase + j + 3));
            }
        }
    }
    // ---- merge the two column halves of every row, write the lists
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (half == 1) {
#pragma unroll
        for (int k = 0; k < kTopK; k++) s_merge[row * kTopK + k] = tk[k];
    }
    __syncthreads();
    if (half == 0 && m0 + row < n_ref) {
#pragma unroll
        for (int k = 0; k < kTopK; k++) topk_insert(tk, s_merge[row * kTopK + k]);
        uint32_t* out = topk + ((long long)f * n_ref + m0 + row) * kTopK;
#pragma unroll
        for (int k = 0; k < kTopK; k += 4) *reinterpret_cast<uint4*>(out + k) = make_uint4(tk[k], tk[k + 1], tk[k + 2], tk[k + 3]);
    }
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(2 * kMmaN) : "memory");
    }
}

// ComputeThreeMaxima (ORBmatcher.cc:1605-1646)
__device__ void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

constexpr int kHistoLen = 30;      // HISTO_LENGTH, ORBmatcher.cc:39

constexpr int kResolveThreads = 256;

__global__ void __launch_bounds__(kResolveThreads)
k_match_resolve(const ulonglong4* __restrict__ ref_desc, const float* __restrict__ ref_angle, int ref_astride, int n_ref,
                const ulonglong4* __restrict__ frame_desc, const float* __restrict__ frame_angle, int frame_astride,
                const int* __restrict__ n_frame, int frame_cap, const uint32_t* __restrict__ topk,
                float ratio, int th_low, int check_ori, float histo_factor,
                int* __restrict__ match_ref_idx, int* __restrict__ n_matches, int stage_topk) {
    extern __shared__ __align__(16) unsigned int s_mem[];
    // kResolveThreads threads stage the frame's lists and angles and clear the outputs (a single warp spent most of this kernel waiting for its own
    // serial global loads); the replay itself is warp 0's
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
    const int nf = min(n_frame[f], frame_cap);
    unsigned int* taken = s_mem;                               // bitmap over frame keypoints
    const int words = (frame_cap + 31) / 32;
    unsigned char* bin_of = reinterpret_cast<unsigned char*>(s_mem + ((words + 3) & ~3));    // 16-byte aligned sections   // rotation bin of each matched frame keypoint
    __shared__ int histo[kHistoLen];
    for (int i = tid; i < words; i += nthr) taken[i] = 0;
    if (tid < kHistoLen) histo[tid] = 0;
    int* mout = match_ref_idx + (long long)f * frame_cap;
    for (int i = tid; i < frame_cap; i += nthr) mout[i] = -1;
    const ulonglong4* fd = frame_desc + (long long)f * frame_cap;
    const float* fa = frame_angle + (long long)f * frame_cap * frame_astride;
    // this frame's top-K lists: staged in shared memory when they fit (the loop below is a serial dependency chain)
    const uint32_t* tk_base = topk + (long long)f * n_ref * kTopK;
    if (stage_topk) {
        // (uint4 copies: the lists are 16-byte records; the angles of both sides follow so that a commit never waits on HBM)
        uint4* s_tk = reinterpret_cast<uint4*>(bin_of + ((frame_cap + 15) & ~15));
        const uint4* g_tk = reinterpret_cast<const uint4*>(tk_base);
        for (int i = tid; i < n_ref * (kTopK / 4); i += nthr) s_tk[i] = g_tk[i];
        tk_base = reinterpret_cast<const uint32_t*>(s_tk);
        if (check_ori) {
            float* s_ra = reinterpret_cast<float*>(s_tk + n_ref * (kTopK / 4));
            float* s_fa = s_ra + n_ref;
            for (int i = tid; i < n_ref; i += nthr) s_ra[i] = ref_angle[(long long)i * ref_astride];
            for (int i = tid; i < nf; i += nthr) s_fa[i] = fa[(long long)i * frame_astride];
            ref_angle = s_ra; ref_astride = 1; fa = s_fa; frame_astride = 1;
        }
    }
    __syncthreads();
    if (tid >= 32) return;
    // The reference loop is serial in r, but the only state it carries is the "taken" set, which changes on accepted
    // matches only.  So 32 consecutive r are evaluated speculatively, one per lane, against the current set.  The lanes are
    // then committed in order: the first lane that accepts a match (or needs the exact rescan) is resolved, and only the
    // later lanes whose top-K list contains the frame keypoint just taken are re-evaluated (from registers); everybody
    // else's verdict is still what the serial loop would have computed.
    int nm = 0;
    for (int r0 = 0; r0 < n_ref; r0 += 32) {
        const int r = r0 + lane;
        uint32_t key[kTopK];
#pragma unroll
        for (int k = 0; k < kTopK; k++) key[k] = 0xffffffffu;
        if (r < n_ref) {
#pragma unroll
            for (int k4 = 0; k4 < kTopK / 4; k4++) {
                const uint4 v = reinterpret_cast<const uint4*>(tk_base + (long long)r * kTopK)[k4];
                key[4 * k4] = v.x; key[4 * k4 + 1] = v.y; key[4 * k4 + 2] = v.z; key[4 * k4 + 3] = v.w;
            }
        }
        uint32_t b1, b2;
        bool exact, accept;
        auto evaluate = [&]() {
            b1 = 0xffffffffu; b2 = 0xffffffffu; exact = true; accept = false;
            if (r >= n_ref) return;
            uint32_t last = 0xffffffffu;
            int found = 0;
#pragma unroll
            for (int k = 0; k < kTopK; k++) {
                if (key[k] == 0xffffffffu) continue;
                last = key[k];
                const int idx = key[k] & 0xffff;
                if (!((taken[idx >> 5] >> (idx & 31)) & 1u)) {
                    if (found == 0) b1 = key[k]; else if (found == 1) b2 = key[k];
                    found++;
                }
            }
            if (key[kTopK - 1] != 0xffffffffu && nf > kTopK) {
                // a list that is not full holds EVERY candidate closer than dstar; a full one may hide untaken candidates beyond its end
                if (found == 0) exact = ((int)(last >> 16) > th_low);          // everything hidden is >= last > th_low: no match possible
                else if (found == 1) exact = ((int)(b1 >> 16) > th_low);       // best is known; second only matters if best can match
            }
            if (exact) {
                const int d1 = b1 == 0xffffffffu ? 256 : (int)(b1 >> 16), d2 = b2 == 0xffffffffu ? 256 : (int)(b2 >> 16);
                accept = d1 <= th_low && (float)d1 < __fmul_rn(ratio, (float)d2);
            }
        };
        evaluate();
        unsigned pending = 0xffffffffu;                                         // lanes not yet passed by the serial order
        for (;;) {
            const unsigned hot = __ballot_sync(0xffffffffu, accept || !exact) & pending;
            if (!hot) break;
            const int first = __ffs(hot) - 1;
            const int rc = r0 + first;                                          // the reference index to commit
            pending = first == 31 ? 0u : (0xffffffffu << (first + 1));
            const bool need_scan = __shfl_sync(0xffffffffu, (int)!exact, first) != 0;
            uint32_t c1 = __shfl_sync(0xffffffffu, b1, first), c2 = __shfl_sync(0xffffffffu, b2, first);
            if (need_scan) {                                                    // rare: rescan the whole row, all lanes
                const ulonglong4 q = ref_desc[rc];
                uint32_t l1 = 0xffffffffu, l2 = 0xffffffffu;
                for (int i = lane; i < nf; i += 32) {
                    if ((taken[i >> 5] >> (i & 31)) & 1u) continue;
                    const uint32_t k2 = ((uint32_t)hamming256(q, fd[i]) << 16) | (uint32_t)i;
                    if (k2 < l1) { l2 = l1; l1 = k2; } else if (k2 < l2) l2 = k2;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint32_t o1 = __shfl_xor_sync(0xffffffffu, l1, o), o2 = __shfl_xor_sync(0xffffffffu, l2, o);
                    const uint32_t lo = min(l1, o1), hi = max(l1, o1);
                    l2 = min(hi, min(l2, o2));
                    l1 = lo;
                }
                c1 = l1; c2 = l2;
            }
            const int d1 = c1 == 0xffffffffu ? 256 : (int)(c1 >> 16);
            const int d2 = c2 == 0xffffffffu ? 256 : (int)(c2 >> 16);
            if (d1 <= th_low && (float)d1 < __fmul_rn(ratio, (float)d2)) {
                const int idx = c1 & 0xffff;
                if (lane == 0) {
                    taken[idx >> 5] |= 1u << (idx & 31);
                    mout[idx] = rc;
                    if (check_ori) {
                        float rot = __fsub_rn(ref_angle[(long long)rc * ref_astride], fa[(long long)idx * frame_astride]);
                        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                        int bin = (int)roundf(__fmul_rn(rot, histo_factor));
                        if (bin == kHistoLen) bin = 0;
                        bin = max(0, min(bin, kHistoLen - 1));   // the reference asserts this range
                        bin_of[idx] = (unsigned char)bin;
                        histo[bin]++;
                    }
                }
                nm++;
                __syncwarp();
                // later rows that listed this keypoint see a different candidate set now
                bool touched = false;
#pragma unroll
                for (int k = 0; k < kTopK; k++) touched |= (int)(key[k] & 0xffff) == idx;
                touched = touched && ((pending >> lane) & 1u);
                if (touched) evaluate();
            }
            if (!pending) break;
        }
        __syncwarp();
    }
    __syncwarp();
    if (check_ori) {
        int i1, i2, i3;
        three_maxima(histo, kHistoLen, i1, i2, i3);
        int removed = 0;
        for (int i = lane; i < nf; i += 32) {
            if (mout[i] >= 0) {
                const int b = bin_of[i];
                if (b != i1 && b != i2 && b != i3) { mout[i] = -1; removed++; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nm -= removed;
    }
    if (lane == 0) n_matches[f] = nm;
}

__global__ void k_hamming_matrix(const ulonglong4* __restrict__ a, int na, const ulonglong4* __restrict__ b, int nb, int* __restrict__ dist) {
    // rows on grid.x (2^31 - 1 blocks), 256-column blocks on grid.y: a map-sized `na` must not hit the 65535 limit of grid.y / grid.z
    const int i = blockIdx.x;
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < nb; j += gridDim.y * blockDim.x) dist[(long long)i * nb + j] = hamming256(a[i], b[j]);
}

// candidate-list core: warp per query
__global__ void __launch_bounds__(256)
k_match_candidates(const ulonglong4* __restrict__ qd, int nq, const ulonglong4* __restrict__ td,
                   const int* __restrict__ ofs, const int* __restrict__ cand,
                   int* __restrict__ best_idx, int* __restrict__ best_dist, int* __restrict__ second_dist) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const ulonglong4 d = qd[q];
    const int beg = ofs[q], end = ofs[q + 1];
    // key = dist << 20 | position in the list: first minimum in list order wins, like the reference's strict '<'
    uint32_t l1 = 0xffffffffu, l2 = 0xffffffffu;
    for (int i = beg + lane; i < end; i += 32) {
        const uint32_t key = ((uint32_t)hamming256(d, td[cand[i]]) << 20) | (uint32_t)(i - beg);
        if (key < l1) { l2 = l1; l1 = key; } else if (key < l2) l2 = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t o1 = __shfl_xor_sync(0xffffffffu, l1, o), o2 = __shfl_xor_sync(0xffffffffu, l2, o);
        const uint32_t lo = min(l1, o1), hi = max(l1, o1);
        l2 = min(hi, min(l2, o2));
        l1 = lo;
    }
    if (lane == 0) {
        best_idx[q] = l1 == 0xffffffffu ? -1 : cand[beg + (l1 & 0xfffff)];
        best_dist[q] = l1 == 0xffffffffu ? 256 : (int)(l1 >> 20);
        second_dist[q] = l2 == 0xffffffffu ? 256 : (int)(l2 >> 20);
    }
}


// ------------------------------------------------------------------------------------------------
// SearchForInitialization (ORBmatcher.cc:409-524): level-0 keypoints of F1 look for their match among the level-0 keypoints of
// F2 inside a window around vbPrevMatched.  The candidate rows come from the frame grid (GetFeaturesInArea, frame.cu) in the
// reference's visit order; distances of all rows are computed in parallel; the accept / steal / histogram logic, which is
// strictly sequential in i1 (vMatchedDistance and vnMatches21 carry state), is replayed by one warp.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_init_queries(const b200_keypoint* __restrict__ k1, int n1, const float* __restrict__ prev, float window, float* __restrict__ q3, int* __restrict__ lv2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const bool on = k1[i].octave <= 0;                    // level1 > 0: skipped by the reference (ORBmatcher.cc:425-427)
    q3[3 * i] = prev[2 * i]; q3[3 * i + 1] = prev[2 * i + 1]; q3[3 * i + 2] = on ? window : -1.f;      // r < 0: no feature passes |d| < r
    lv2[2 * i] = 0; lv2[2 * i + 1] = 0;
}

__global__ void __launch_bounds__(256)
k_init_dist(const ulonglong4* __restrict__ d1, int n1, const ulonglong4* __restrict__ d2, const int* __restrict__ cand, const int* __restrict__ cnt,
            int row_cap, int* __restrict__ dist) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= n1) return;
    const int n = min(cnt[q], row_cap);
    if (n == 0) return;
    const ulonglong4 a = d1[q];
    for (int j = lane; j < n; j += 32) dist[(long long)q * row_cap + j] = hamming256(a, d2[cand[(long long)q * row_cap + j]]);
}

__global__ void __launch_bounds__(32)
k_init_resolve(const b200_keypoint* __restrict__ k1, int n1, const b200_keypoint* __restrict__ k2, int n2,
               const int* __restrict__ cand, const int* __restrict__ cnt, const int* __restrict__ dist, int row_cap,
               float ratio, int th_low, int check_ori, float* __restrict__ prev, int* __restrict__ m12, int* __restrict__ m21, int* __restrict__ mdist,
               unsigned char* __restrict__ rotbin, int* __restrict__ result /* nmatches, overflow flag */) {
    __shared__ int histo[kHistoLen];
    const int lane = threadIdx.x;
    if (lane < kHistoLen) histo[lane] = 0;
    for (int i = lane; i < n1; i += 32) { m12[i] = -1; rotbin[i] = 255; }
    for (int i = lane; i < n2; i += 32) { m21[i] = -1; mdist[i] = 0x7fffffff; }
    __syncwarp();
    int nmatches = 0, overflow = 0;
    const float factor = __fdiv_rn(1.0f, (float)kHistoLen);                 // the reference's 1.0f / HISTO_LENGTH (ORBmatcher.cc:417)
    for (int i1 = 0; i1 < n1; i1++) {
        if (k1[i1].octave > 0) continue;
        const int n = cnt[i1];
        if (n > row_cap) overflow = 1;
        const int nn = min(n, row_cap);
        if (nn == 0) continue;
        // key = dist << 20 | position in the row: strict '<' of the reference == first minimum in visit order
        unsigned long long l1 = ~0ull, l2 = ~0ull;
        for (int j = lane; j < nn; j += 32) {
            const int d = dist[(long long)i1 * row_cap + j], i2 = cand[(long long)i1 * row_cap + j];
            if (mdist[i2] <= d) continue;
            const unsigned long long key = ((unsigned long long)d << 20) | (unsigned)j;
            if (key < l1) { l2 = l1; l1 = key; } else if (key < l2) l2 = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, l1, o), o2 = __shfl_xor_sync(0xffffffffu, l2, o);
            const unsigned long long lo = min(l1, o1), hi = max(l1, o1);
            l2 = min(hi, min(l2, o2));
            l1 = lo;
        }
        if (l1 == ~0ull) continue;
        const int best = (int)(l1 >> 20);
        const bool has2 = l2 != ~0ull;
        const int best2 = has2 ? (int)(l2 >> 20) : 0x7fffffff;
        if (best <= th_low && (float)best < __fmul_rn((float)best2, ratio)) {
            const int i2 = cand[(long long)i1 * row_cap + (int)(l1 & 0xfffff)];
            const int old = m21[i2];
            __syncwarp();                                                  // everybody has read the old owner before lane 0 replaces it
            if (old >= 0) nmatches--;
            nmatches++;
            if (lane == 0) {
                if (old >= 0) m12[old] = -1;                               // steal (ORBmatcher.cc:467-471)
                m12[i1] = i2; m21[i2] = i1; mdist[i2] = best;
                if (check_ori) {
                    float rot = __fsub_rn(k1[i1].angle, k2[i2].angle);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, factor));
                    if (bin == kHistoLen) bin = 0;
                    bin = max(0, min(bin, kHistoLen - 1));
                    rotbin[i1] = (unsigned char)bin;
                    histo[bin]++;
                }
            }
        }
        __syncwarp();
    }
    __syncwarp();
    if (check_ori) {
        int a, b, c;
        three_maxima(histo, kHistoLen, a, b, c);
        int removed = 0;
        for (int i = lane; i < n1; i += 32) {
            const int bin = rotbin[i];
            if (bin != 255 && bin != a && bin != b && bin != c && m12[i] >= 0) { m12[i] = -1; removed++; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
    }
    __syncwarp();
    for (int i = lane; i < n1; i += 32) if (m12[i] >= 0) { prev[2 * i] = k2[m12[i]].x; prev[2 * i + 1] = k2[m12[i]].y; }
    if (lane == 0) { result[0] = nmatches; result[1] = overflow; }
}


// ------------------------------------------------------------------------------------------------
// SearchByProjection on ready-made projections (the geometry is host glue in the reference):
//   mode 0  SearchByProjection(Frame&, const vector<MapPoint*>&, th)        (ORBmatcher.cc:45-129)
//   mode 1  SearchByProjection(Frame& Current, const Frame& Last, th, mono)  (ORBmatcher.cc:1332-1474)
// Candidate rows (frame grid, frame.cu) and all distances (k_init_dist) are computed in parallel; one warp replays the queries in
// order because an assigned, observed map point hides its frame keypoint from every later query.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_proj_resolve(const b200_keypoint* __restrict__ k2, int n2, const int* __restrict__ cand, const int* __restrict__ cnt, const int* __restrict__ dist, int row_cap,
               const float* __restrict__ q_angle, const unsigned char* __restrict__ q_observed, int nq, int mode, float ratio, int th_high, int check_ori,
               unsigned char* __restrict__ occupied, int* __restrict__ assign, int* __restrict__ ent_bin, int* __restrict__ ent_idx,
               int* __restrict__ result /* nmatches, overflow flag */) {
    __shared__ int histo[kHistoLen];
    const int lane = threadIdx.x;
    if (lane < kHistoLen) histo[lane] = 0;
    for (int i = lane; i < n2; i += 32) assign[i] = -1;
    __syncwarp();
    int nmatches = 0, overflow = 0, nent = 0;
    const float factor = __fdiv_rn(1.0f, (float)kHistoLen);
    for (int q = 0; q < nq; q++) {
        const int n = cnt[q];
        if (n > row_cap) overflow = 1;
        const int nn = min(n, row_cap);
        if (nn == 0) continue;
        unsigned long long l1 = ~0ull, l2 = ~0ull;
        for (int j = lane; j < nn; j += 32) {
            const int i2 = cand[(long long)q * row_cap + j];
            if (occupied[i2]) continue;
            const unsigned long long key = ((unsigned long long)dist[(long long)q * row_cap + j] << 20) | (unsigned)j;
            if (key < l1) { l2 = l1; l1 = key; } else if (key < l2) l2 = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, l1, o), o2 = __shfl_xor_sync(0xffffffffu, l2, o);
            const unsigned long long lo = min(l1, o1), hi = max(l1, o1);
            l2 = min(hi, min(l2, o2));
            l1 = lo;
        }
        if (l1 == ~0ull) continue;
        const int best = (int)(l1 >> 20);
        if (best > th_high) continue;
        const int i2 = cand[(long long)q * row_cap + (int)(l1 & 0xfffff)];
        if (mode == 0 && l2 != ~0ull) {
            // ratio only when best and second lie in the same pyramid level (ORBmatcher.cc:118-121)
            const int best2 = (int)(l2 >> 20), j2 = cand[(long long)q * row_cap + (int)(l2 & 0xfffff)];
            if (k2[i2].octave == k2[j2].octave && (float)best > __fmul_rn(ratio, (float)best2)) continue;
        }
        nmatches++;
        if (lane == 0) {
            assign[i2] = q;
            if (q_observed[q]) occupied[i2] = 1;
            if (mode == 1 && check_ori) {
                float rot = __fsub_rn(q_angle[q], k2[i2].angle);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, factor));
                if (bin == kHistoLen) bin = 0;
                bin = max(0, min(bin, kHistoLen - 1));
                ent_bin[nent] = bin; ent_idx[nent] = i2;
                histo[bin]++;
            }
        }
        nent++;
        __syncwarp();
    }
    __syncwarp();
    if (mode == 1 && check_ori) {
        int a, b, c;
        three_maxima(histo, kHistoLen, a, b, c);
        // every histogram entry outside the three strongest bins clears its frame keypoint and is subtracted, even when the same keypoint
        // was pushed twice (ORBmatcher.cc:1459-1467); entries are independent, so the order does not matter
        int removed = 0;
        for (int e = lane; e < nent; e += 32) {
            const int bin = ent_bin[e];
            if (bin != a && bin != b && bin != c) { assign[ent_idx[e]] = -2; removed++; }     // -2: assigned, then cleared (the reference stores NULL there)
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
    }
    if (lane == 0) { result[0] = nmatches; result[1] = overflow; }
}

// ------------------------------------------------------------------------------------------------
// SearchByBoW over real FeatureVectors: the host merge-walks the two vectors (ORBmatcher.cc:185-279 / 547-632) into GROUPS, one per common
// vocabulary node: a run of query features (already filtered to good map points, in node order) and a run of candidate features.
//   mode 0  SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)   '<= TH_LOW', factor 30/360, out[idxF] = idxKF    (ORBmatcher.cc:159-292)
//   mode 1  SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)      '<  TH_LOW', factor 1/30,   out[idx1]  = idx2     (ORBmatcher.cc:526-659)
// All (query, candidate) distances of a group are independent -> k_bow_dist; the taken flags make the accept loop sequential -> one warp.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bow_dist(const ulonglong4* __restrict__ dq, const ulonglong4* __restrict__ dc, const int* __restrict__ q_idx, const int* __restrict__ q_grp,
           const int* __restrict__ grp_c_ofs, const int* __restrict__ c_idx, const int* __restrict__ q_dofs, int nq, int* __restrict__ dist) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const int g = q_grp[q], c0 = grp_c_ofs[g], n = grp_c_ofs[g + 1] - c0, o = q_dofs[q];
    if (n == 0) return;
    const ulonglong4 a = dq[q_idx[q]];
    for (int j = lane; j < n; j += 32) dist[o + j] = hamming256(a, dc[c_idx[c0 + j]]);
}

__global__ void __launch_bounds__(32)
k_bow_resolve(const float* __restrict__ q_angle, const float* __restrict__ c_angle, const int* __restrict__ q_idx, const int* __restrict__ q_grp,
              const int* __restrict__ grp_c_ofs, const int* __restrict__ c_idx, const int* __restrict__ q_dofs, const int* __restrict__ dist, int nq,
              int mode, float ratio, int th_low, int check_ori, unsigned char* __restrict__ taken, int n_out, int* __restrict__ out,
              int* __restrict__ ent_bin, int* __restrict__ ent_idx, int* __restrict__ result) {
    __shared__ int histo[kHistoLen];
    const int lane = threadIdx.x;
    if (lane < kHistoLen) histo[lane] = 0;
    for (int i = lane; i < n_out; i += 32) out[i] = -1;
    __syncwarp();
    int nmatches = 0, nent = 0;
    // HISTO_LENGTH / 360.0f (ORBmatcher.cc:176) against 1.0f / HISTO_LENGTH (ORBmatcher.cc:541)
    const float factor = mode == 0 ? __fdiv_rn((float)kHistoLen, 360.0f) : __fdiv_rn(1.0f, (float)kHistoLen);
    const int limit = mode == 0 ? th_low : th_low - 1;
    for (int q = 0; q < nq; q++) {
        const int g = q_grp[q], c0 = grp_c_ofs[g], n = grp_c_ofs[g + 1] - c0, o = q_dofs[q];
        unsigned long long l1 = ~0ull, l2 = ~0ull;
        for (int j = lane; j < n; j += 32) {
            if (taken[c_idx[c0 + j]]) continue;
            const unsigned long long key = ((unsigned long long)dist[o + j] << 32) | (unsigned)j;
            if (key < l1) { l2 = l1; l1 = key; } else if (key < l2) l2 = key;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, l1, s), o2 = __shfl_xor_sync(0xffffffffu, l2, s);
            const unsigned long long lo = min(l1, o1), hi = max(l1, o1);
            l2 = min(hi, min(l2, o2));
            l1 = lo;
        }
        if (l1 == ~0ull) continue;
        const int best = (int)(l1 >> 32), best2 = l2 != ~0ull ? (int)(l2 >> 32) : 256;
        if (best > limit || !((float)best < __fmul_rn(ratio, (float)best2))) continue;
        const int ic = c_idx[c0 + (int)(l1 & 0xffffffffu)], iq = q_idx[q];
        nmatches++;
        if (lane == 0) {
            taken[ic] = 1;
            if (mode == 0) out[ic] = iq; else out[iq] = ic;
            if (check_ori) {
                float rot = __fsub_rn(q_angle[iq], c_angle[ic]);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, factor));
                if (bin == kHistoLen) bin = 0;
                bin = max(0, min(bin, kHistoLen - 1));
                ent_bin[nent] = bin; ent_idx[nent] = mode == 0 ? ic : iq;
                histo[bin]++;
            }
        }
        nent++;
        __syncwarp();
    }
    __syncwarp();
    if (check_ori) {
        int a, b, c;
        three_maxima(histo, kHistoLen, a, b, c);
        int removed = 0;
        for (int e = lane; e < nent; e += 32) {
            const int bin = ent_bin[e];
            if (bin != a && bin != b && bin != c) { out[ent_idx[e]] = -1; removed++; }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, s);
        nmatches -= removed;
    }
    if (lane == 0) result[0] = nmatches;
}

// ------------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:271-331), batched over map points: the observation whose MEDIAN Hamming
// distance to all observations of the point (itself included, vDists[0.5 * (N - 1)] of the sorted row) is smallest; first minimum wins.
// One warp per map point.  Row i: lane j holds d(i, j) (N <= 32: in a register, descriptors loaded once; larger N: u16 row buffer in shared
// memory) and the k-th smallest value is found by bisection on the value range 0..256 with ballot / popc counts - no sort.
// ------------------------------------------------------------------------------------------------
constexpr int kMedoidWarps = 4;

__global__ void __launch_bounds__(kMedoidWarps * 32)
k_distinctive(const ulonglong4* __restrict__ desc, const int* __restrict__ ofs, int n_points, int row_cap, int* __restrict__ best_idx,
              ulonglong4* __restrict__ out_desc) {
    extern __shared__ unsigned short s_row[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int p = blockIdx.x * kMedoidWarps + wid;
    if (p >= n_points) return;
    const int o = ofs[p], n = ofs[p + 1] - o;
    if (n <= 0) { if (lane == 0) best_idx[p] = -1; return; }
    const int k = (n - 1) >> 1;                                         // (int)(0.5 * (N - 1))
    unsigned short* row = s_row + (size_t)wid * row_cap;
    int best_med = 0x7fffffff, best = 0;
    ulonglong4 mine = make_ulonglong4(0, 0, 0, 0);
    if (n <= 32 && lane < n) mine = desc[o + lane];
    for (int i = 0; i < n; i++) {
        const ulonglong4 a = desc[o + i];                               // uniform address: one broadcast load
        int d = 0x7fff;
        if (n <= 32) { if (lane < n) d = hamming256(a, mine); }
        else {
            for (int j = lane; j < n; j += 32) row[j] = (unsigned short)hamming256(a, desc[o + j]);
            __syncwarp();
        }
        int lo = 0, hi = 256;                                           // smallest v with #(d <= v) >= k + 1
        while (lo < hi) {
            const int v = (lo + hi) >> 1;
            int cnt;
            if (n <= 32) cnt = __popc(__ballot_sync(0xffffffffu, d <= v));
            else {
                cnt = 0;
                for (int j = lane; j < n; j += 32) cnt += row[j] <= v;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
            }
            if (cnt >= k + 1) hi = v; else lo = v + 1;
        }
        if (lo < best_med) { best_med = lo; best = i; }
        __syncwarp();
    }
    if (lane == 0) { best_idx[p] = best; if (out_desc) out_desc[p] = desc[o + best]; }
}

// ------------------------------------------------------------------------------------------------
// SearchForTriangulation (ORBmatcher.cc:661-829), monocular.  The reference never sets vbMatched2, so every KF1 feature is independent: its
// answer is the candidate of the same vocabulary node with the least distance among those that pass (dist <= TH_LOW, far enough from the
// epipole, close enough to the epipolar line), the LAST one in node order on ties ('dist > bestDist -> continue' lets an equal distance
// replace).  Warp per KF1 feature; groups as in SearchByBoW (features that already own a MapPoint dropped on both sides).
// ------------------------------------------------------------------------------------------------
struct TriangGeom { float F[9]; float ex, ey; float sf[16]; float sigma2[16]; };

__global__ void __launch_bounds__(256)
k_triang(const b200_keypoint* __restrict__ k1, const ulonglong4* __restrict__ d1, const b200_keypoint* __restrict__ k2, const ulonglong4* __restrict__ d2,
         const int* __restrict__ q_idx, const int* __restrict__ q_grp, const int* __restrict__ grp_c_ofs, const int* __restrict__ c_idx, int nq,
         const TriangGeom g, int th_low, int* __restrict__ m12, unsigned char* __restrict__ rotbin) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const int i1 = q_idx[q], grp = q_grp[q], c0 = grp_c_ofs[grp], n = grp_c_ofs[grp + 1] - c0;
    const b200_keypoint kp1 = k1[i1];
    const ulonglong4 a = d1[i1];
    // epipolar line in the second image l = x1' F12 (CheckDistEpipolarLine, ORBmatcher.cc:139-157)
    const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g.F[0]), __fmul_rn(kp1.y, g.F[3])), g.F[6]);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g.F[1]), __fmul_rn(kp1.y, g.F[4])), g.F[7]);
    const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g.F[2]), __fmul_rn(kp1.y, g.F[5])), g.F[8]);
    const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    uint32_t best = 0xffffffffu;                                  // dist << 20 | (0xfffff - position): least distance, last position
    if (den != 0.0f)
        for (int j = lane; j < n; j += 32) {
            const int i2 = c_idx[c0 + j];
            const int dist = hamming256(a, d2[i2]);
            if (dist > th_low) continue;
            const b200_keypoint kp2 = k2[i2];
            const int oct = max(0, min(kp2.octave, 15));
            const float dex = __fsub_rn(g.ex, kp2.x), dey = __fsub_rn(g.ey, kp2.y);
            if (__fadd_rn(__fmul_rn(dex, dex), __fmul_rn(dey, dey)) < __fmul_rn(100.0f, g.sf[oct])) continue;
            const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
            const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
            if (!((double)dsqr < __dmul_rn(3.84, (double)g.sigma2[oct]))) continue;
            best = min(best, ((uint32_t)dist << 20) | (uint32_t)(0xfffff - j));
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
        int i2 = -1, bin = 255;
        if (best != 0xffffffffu) {
            i2 = c_idx[c0 + (0xfffff - (int)(best & 0xfffff))];
            float rot = __fsub_rn(kp1.angle, k2[i2].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            bin = (int)roundf(__fmul_rn(rot, __fdiv_rn(1.0f, (float)kHistoLen)));
            if (bin == kHistoLen) bin = 0;
            bin = max(0, min(bin, kHistoLen - 1));
        }
        m12[i1] = i2; rotbin[i1] = (unsigned char)bin;
    }
}

// rotation histogram over the accepted pairs (ORBmatcher.cc:740-750, 776-794) and the count; one CTA
__global__ void __launch_bounds__(256)
k_triang_finish(int* __restrict__ m12, const unsigned char* __restrict__ rotbin, int n1, int check_ori, int* __restrict__ result) {
    __shared__ int histo[kHistoLen];
    __shared__ int keep[3];
    __shared__ int total;
    if (threadIdx.x < kHistoLen) histo[threadIdx.x] = 0;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (check_ori) {
        for (int i = threadIdx.x; i < n1; i += blockDim.x) if (m12[i] >= 0) atomicAdd(&histo[rotbin[i]], 1);
        __syncthreads();
        if (threadIdx.x == 0) three_maxima(histo, kHistoLen, keep[0], keep[1], keep[2]);
        __syncthreads();
    }
    int cnt = 0;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) {
        if (m12[i] < 0) continue;
        const int bin = rotbin[i];
        if (check_ori && bin != keep[0] && bin != keep[1] && bin != keep[2]) m12[i] = -1; else cnt++;
    }
    atomicAdd(&total, cnt);
    __syncthreads();
    if (threadIdx.x == 0) result[0] = total;
}

// ------------------------------------------------------------------------------------------------
// Best keyframe feature of a radius query: the inner loop shared by Fuse (ORBmatcher.cc:906-955), Fuse(Scw) (:1063-1082) and SearchBySim3
// (:1196-1217, :1276-1297).  These searches do not depend on what earlier points did to the keyframe, so all queries run at once; the
// candidates come from the keyframe's feature grid with the level window [predicted - 1, predicted] (k_features_in_area), the optional gate is
// Fuse's reprojection test e2 * invSigma2[level] > chi2.  Strict '<' of the reference == first minimum in grid visit order.  Warp per query.
// ------------------------------------------------------------------------------------------------
struct LevelTable { float inv_sigma2[16]; };

__global__ void __launch_bounds__(256)
k_radius_best(const b200_keypoint* __restrict__ k, const ulonglong4* __restrict__ d, const float* __restrict__ q3, const ulonglong4* __restrict__ qd,
              const int* __restrict__ cand, const int* __restrict__ cnt, int row_cap, int nq, const LevelTable lt, double chi2,
              int* __restrict__ best_idx, int* __restrict__ best_dist, int* __restrict__ overflow) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const int n = cnt[q];
    if (n > row_cap && lane == 0) *overflow = 1;
    const int nn = min(n, row_cap);
    const float u = q3[3 * q], v = q3[3 * q + 1];
    const ulonglong4 a = qd[q];
    uint32_t best = 0xffffffffu;                                  // dist << 20 | position in the row
    for (int j = lane; j < nn; j += 32) {
        const int idx = cand[(long long)q * row_cap + j];
        if (chi2 > 0.0) {
            const b200_keypoint kp = k[idx];
            const float ex = __fsub_rn(u, kp.x), ey = __fsub_rn(v, kp.y);
            const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            if ((double)__fmul_rn(e2, lt.inv_sigma2[max(0, min(kp.octave, 15))]) > chi2) continue;
        }
        best = min(best, ((uint32_t)hamming256(a, d[idx]) << 20) | (uint32_t)j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
        best_idx[q] = best == 0xffffffffu ? -1 : cand[(long long)q * row_cap + (int)(best & 0xfffff)];
        best_dist[q] = best == 0xffffffffu ? 256 : (int)(best >> 20);
    }
}

// ------------------------------------------------------------------------------------------------
// The projection front of Fuse (ORBmatcher.cc:848-898), Fuse(Scw) (:1006-1061), SearchByProjection(pKF, Scw, ...) (:318-366) and of each direction of
// SearchBySim3 (:1158-1195 / :1238-1275) for all map points at once: thread per point, the reference's statements in its cv::Mat CV_32F arithmetic
// (3-term products accumulate in double and are stored as float; sums, differences and the pinhole model are float; norm / dot in double).
// MapPoint::PredictScale (MapPoint.cc:403-435) is ceil(logf(ratio) / logf(scaleFactor)), monotone in ratio: the host finds, with ITS libm, the
// largest float ratio that still yields level n (b200_kf_project_host's level_thresholds), and the level is the number of thresholds below the
// ratio - the same integers as the host's logf without a device logf.
// ------------------------------------------------------------------------------------------------
struct ProjGeom {
    float R[9], t[3], Ow[3], sR[9], tt[3];
    float fx, fy, cx, cy, min_x, max_x, min_y, max_y, th;
    float sf[16], thr[16];
    int nlevels, sim3, has_normal;
};

__device__ __forceinline__ float dot3_f(const float* a, float b0, float b1, float b2) {
    return __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn((double)a[0], (double)b0), __dmul_rn((double)a[1], (double)b1)), __dmul_rn((double)a[2], (double)b2)));
}
__device__ __forceinline__ double dot3_d(float a0, float a1, float a2, float b0, float b1, float b2) {
    return __dadd_rn(__dadd_rn(__dmul_rn((double)a0, (double)b0), __dmul_rn((double)a1, (double)b1)), __dmul_rn((double)a2, (double)b2));
}

__global__ void __launch_bounds__(256)
k_kf_project(const float* __restrict__ pos, const float* __restrict__ normal, const float* __restrict__ minmax, int n, const ProjGeom g,
             unsigned char* __restrict__ valid, float* __restrict__ q3, int* __restrict__ level,
             const unsigned char* __restrict__ skip = nullptr, int* __restrict__ lv2 = nullptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float X = pos[3 * i], Y = pos[3 * i + 1], Z = pos[3 * i + 2];
    float p0 = __fadd_rn(dot3_f(g.R, X, Y, Z), g.t[0]), p1 = __fadd_rn(dot3_f(g.R + 3, X, Y, Z), g.t[1]), p2 = __fadd_rn(dot3_f(g.R + 6, X, Y, Z), g.t[2]);
    if (g.sim3) {
        const float a0 = p0, a1 = p1, a2 = p2;
        p0 = __fadd_rn(dot3_f(g.sR, a0, a1, a2), g.tt[0]); p1 = __fadd_rn(dot3_f(g.sR + 3, a0, a1, a2), g.tt[1]); p2 = __fadd_rn(dot3_f(g.sR + 6, a0, a1, a2), g.tt[2]);
    }
    bool ok = !(p2 < 0.0f);                                       // depth must be positive
    const float invz = __fdiv_rn(1.0f, p2);
    const float u = __fadd_rn(__fmul_rn(g.fx, __fmul_rn(p0, invz)), g.cx), v = __fadd_rn(__fmul_rn(g.fy, __fmul_rn(p1, invz)), g.cy);
    ok = ok && u >= g.min_x && u < g.max_x && v >= g.min_y && v < g.max_y;      // KeyFrame::IsInImage
    const float max_d = __fmul_rn(1.2f, minmax[2 * i + 1]), min_d = __fmul_rn(0.8f, minmax[2 * i]);  // GetMax/MinDistanceInvariance
    float dist;
    if (g.sim3) dist = __double2float_rn(__dsqrt_rn(dot3_d(p0, p1, p2, p0, p1, p2)));
    else {
        const float o0 = __fsub_rn(X, g.Ow[0]), o1 = __fsub_rn(Y, g.Ow[1]), o2 = __fsub_rn(Z, g.Ow[2]);
        dist = __double2float_rn(__dsqrt_rn(dot3_d(o0, o1, o2, o0, o1, o2)));
        if (g.has_normal && dot3_d(o0, o1, o2, normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]) < __dmul_rn(0.5, (double)dist)) ok = false;   // 60 degrees
    }
    if (dist < min_d || dist > max_d) ok = false;
    const float ratio = __fdiv_rn(minmax[2 * i + 1], dist);
    int lv = 0;
    for (int k = 0; k < g.nlevels - 1; k++) lv += ratio > g.thr[k];
    if (skip && skip[i]) ok = false;                              // the reference `continue`s on this point before projecting (NULL, bad, already found ...)
    valid[i] = ok ? 1 : 0;
    level[i] = lv;
    if (lv2) {
        // feeding k_features_in_area directly: the level window of the callers' filter, and an empty window (radius < 0) for discarded points
        lv2[2 * i] = lv - 1; lv2[2 * i + 1] = lv;
        q3[3 * i] = ok ? u : 0.0f; q3[3 * i + 1] = ok ? v : 0.0f; q3[3 * i + 2] = ok ? __fmul_rn(g.th, g.sf[lv]) : -1.0f;
    } else { q3[3 * i] = u; q3[3 * i + 1] = v; q3[3 * i + 2] = __fmul_rn(g.th, g.sf[lv]); }
}

// top-K scratch of the device-pointer matcher: one buffer per (device, stream) of the calling thread, so that two asynchronous calls a thread
// enqueues on different streams never share top-K lists; the chunked front end (base / total slots) owns the entry keyed by its first stream.
struct MatchScratch { uint32_t* topk; size_t cap; int device; cudaStream_t stream; };
struct MatchScratchSet {
    std::vector<MatchScratch> v;
    ~MatchScratchSet() { for (auto& m : v) if (m.topk) cudaFree(m.topk); }       // thread exit; errors after context teardown are ignored
    MatchScratch& get(int device, cudaStream_t st) {
        for (auto& m : v) if (m.device == device && m.stream == st) return m;
        v.push_back(MatchScratch{nullptr, 0, device, st});
        return v.back();
    }
};
static thread_local MatchScratchSet g_ms_set;

}  // namespace b200

using namespace b200;

extern "C" {

static int match_bf_impl(const uint8_t* ref_desc, const float* ref_angle, int ref_astride, int n_ref,
                         const uint8_t* frame_desc, const float* frame_angle, int frame_astride, const int32_t* n_frame, int n_batch, int frame_cap,
                         float ratio, int th_low, int check_ori, float histo_factor,
                         int32_t* match_ref_idx, int32_t* n_matches, int device, void* stream, int base = 0, int total = 0) {
    // base / total: this call uses the scratch slots [base, base + n_batch) of a scratch sized for `total` frames, so that calls on
    // different streams (the chunks of b200_frontend_host) do not share top-K lists
    if (n_batch < 0 || n_ref < 0 || frame_cap < 0) return fail(B200_EINVAL, "negative %s", "size");
    if (frame_cap > 65535) return fail(B200_ECAPACITY, "frame_cap above %s", "65535");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    if (n_batch == 0) return B200_OK;
    if (!n_frame || !match_ref_idx || !n_matches || (n_ref > 0 && (!ref_desc || !ref_angle)) || (frame_cap > 0 && (!frame_desc || !frame_angle)))
        return fail(B200_EINVAL, "null %s", "pointer");
    if (((uintptr_t)ref_desc | (uintptr_t)frame_desc) & 31) return fail(B200_EINVAL, "descriptor arrays must be %s", "32-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t need = std::max<size_t>((size_t)std::max(total, base + n_batch) * std::max(n_ref, 1) * kTopK * 4, 16);
    MatchScratch& g_ms = g_ms_set.get(device, total > 0 ? (cudaStream_t)(uintptr_t)1 : st);     // ranged calls (b200_frontend_host) share one slotted buffer
    if (g_ms.cap < need) {
        if (g_ms.topk) cudaFree(g_ms.topk);                     // cudaFree waits for the device: no launch still reads the old block
        g_ms.topk = nullptr; g_ms.cap = 0;
        B200_CUDA(cudaMalloc((void**)&g_ms.topk, need));
        g_ms.cap = need;
    }
    uint32_t* topk = g_ms.topk + (size_t)base * std::max(n_ref, 1) * kTopK;
    const int nf_pad = (frame_cap + 31) & ~31;
    // dstar: the smallest distance D with th_low < ratio * D in the kernel's float arithmetic.  A best distance d1 <= th_low passes the
    // ratio test against ANY second-best >= dstar, and a best distance > th_low never matches, so the exact value of a distance >= dstar
    // can never change a decision: k_match_topk lists only the candidates below it (for ORB descriptors that is a handful per row).
    int dstar = 257;
    for (int D = 0; D <= 256; D++) if ((float)th_low < ratio * (float)D) { dstar = D; break; }
    dstar = std::max(dstar, th_low + 1);
    if (n_ref > 0 && frame_cap > 0) {
        const size_t smem = (size_t)4 * nf_pad * 8;
        if (smem > 200 * 1024) return fail(B200_ECAPACITY, "frame_cap too large for the %s", "shared-memory descriptor planes");
        B200_CUDA(cudaFuncSetAttribute(k_match_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // tensor-core distance stage (k_match_mma) unless B200_MATCH_POPC asks for the popcount kernel; both write the same lists
        static const bool match_popc = getenv("B200_MATCH_POPC") != nullptr;
        if (!match_popc && frame_cap < 65536) {
            static std::atomic<int> mma_attr(0);
            if (!mma_attr.load()) { B200_CUDA(cudaFuncSetAttribute(k_match_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmem)); mma_attr.store(1); }
            dim3 grid((n_ref + kMmaM - 1) / kMmaM, n_batch);
            B200_LAUNCH(k_match_mma, grid, kMmaThreads, kMmaSmem, st, (const ulonglong4*)ref_desc, n_ref, (const ulonglong4*)frame_desc, n_frame, frame_cap, topk, dstar);
        } else {
            dim3 grid((n_ref + kRefPerCta - 1) / kRefPerCta, n_batch);
            B200_LAUNCH(k_match_topk, grid, kMatchWarps * 32, smem, st, (const ulonglong4*)ref_desc, n_ref, (const ulonglong4*)frame_desc,
                        n_frame, frame_cap, topk, dstar);
        }
    }
    size_t smem2 = (size_t)((((frame_cap + 31) / 32) + 3) & ~3) * 4 + (size_t)((frame_cap + 15) & ~15) + 16;
    const size_t tk_bytes = (size_t)n_ref * kTopK * 4 + ((size_t)n_ref + frame_cap) * 4;     // top-K lists + both angle arrays
    const int stage_topk = smem2 + tk_bytes <= 160 * 1024;
    if (stage_topk) smem2 += tk_bytes;
    B200_CUDA(cudaFuncSetAttribute(k_match_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem2, 1024)));
    B200_LAUNCH(k_match_resolve, n_batch, kResolveThreads, smem2, st, (const ulonglong4*)ref_desc, ref_angle, ref_astride, n_ref, (const ulonglong4*)frame_desc,
                frame_angle, frame_astride, n_frame, frame_cap, topk, ratio, th_low, check_ori, histo_factor, match_ref_idx, n_matches, stage_topk);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

int b200_match_bf(const uint8_t* ref_desc, const float* ref_angle, int n_ref,
                  const uint8_t* frame_desc, const float* frame_angle, const int32_t* n_frame, int n_batch, int frame_cap,
                  float ratio, int th_low, int check_ori, float histo_factor,
                  int32_t* match_ref_idx, int32_t* n_matches, int device, void* stream) {
    return match_bf_impl(ref_desc, ref_angle, 1, n_ref, frame_desc, frame_angle, 1, n_frame, n_batch, frame_cap, ratio, th_low, check_ori,
                         histo_factor, match_ref_idx, n_matches, device, stream);
}

// same, with the angles read straight out of keypoint records (the extractor's output feeds the matcher unchanged)
int b200_match_bf_kp(const uint8_t* ref_desc, const b200_keypoint* ref_kps, int n_ref,
                     const uint8_t* frame_desc, const b200_keypoint* frame_kps, const int32_t* n_frame, int n_batch, int frame_cap,
                     float ratio, int th_low, int check_ori, float histo_factor,
                     int32_t* match_ref_idx, int32_t* n_matches, int device, void* stream) {
    const int st = (int)(sizeof(b200_keypoint) / sizeof(float));
    return match_bf_impl(ref_desc, ref_kps ? &ref_kps->angle : nullptr, st, n_ref, frame_desc, frame_kps ? &frame_kps->angle : nullptr, st,
                         n_frame, n_batch, frame_cap, ratio, th_low, check_ori, histo_factor, match_ref_idx, n_matches, device, stream);
}

int b200_match_bf_kp_range(const uint8_t* ref_desc, const b200_keypoint* ref_kps, int n_ref,
                           const uint8_t* frame_desc, const b200_keypoint* frame_kps, const int32_t* n_frame, int n_batch, int frame_cap,
                           float ratio, int th_low, int check_ori, float histo_factor,
                           int32_t* match_ref_idx, int32_t* n_matches, int device, void* stream, int base, int total) {
    const int st = (int)(sizeof(b200_keypoint) / sizeof(float));
    return match_bf_impl(ref_desc, ref_kps ? &ref_kps->angle : nullptr, st, n_ref, frame_desc, frame_kps ? &frame_kps->angle : nullptr, st,
                         n_frame, n_batch, frame_cap, ratio, th_low, check_ori, histo_factor, match_ref_idx, n_matches, device, stream, base, total);
}

int b200_match_bf_host(const uint8_t* ref_desc, const float* ref_angle, int n_ref,
                       const uint8_t* frame_desc, const float* frame_angle, const int32_t* n_frame, int n_batch, int frame_cap,
                       float ratio, int th_low, int check_ori, float histo_factor,
                       int32_t* match_ref_idx, int32_t* n_matches, int device) {
    if (n_batch < 0 || n_ref < 0 || frame_cap < 0) return fail(B200_EINVAL, "negative %s", "size");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (n_batch == 0) return B200_OK;
    DevBuf rd, ra, fd, fa, nf, mi, nm;
    if ((rc = rd.upload(ref_desc, (size_t)n_ref * 32)) || (rc = ra.upload(ref_angle, (size_t)n_ref * 4)) ||
        (rc = fd.upload(frame_desc, (size_t)n_batch * frame_cap * 32)) || (rc = fa.upload(frame_angle, (size_t)n_batch * frame_cap * 4)) ||
        (rc = nf.upload(n_frame, (size_t)n_batch * 4)) || (rc = mi.alloc((size_t)n_batch * frame_cap * 4)) || (rc = nm.alloc((size_t)n_batch * 4)))
        return rc;
    if ((rc = b200_match_bf((const uint8_t*)rd.p, (const float*)ra.p, n_ref, (const uint8_t*)fd.p, (const float*)fa.p, (const int32_t*)nf.p,
                            n_batch, frame_cap, ratio, th_low, check_ori, histo_factor, (int32_t*)mi.p, (int32_t*)nm.p, device, ts)))
        return rc;
    B200_CUDA(cudaStreamSynchronize(ts));
    if (frame_cap) B200_D2H(match_ref_idx, mi.p, (size_t)n_batch * frame_cap * 4);
    B200_D2H(n_matches, nm.p, (size_t)n_batch * 4);
    return B200_OK;
}

int b200_match_for_initialization_host(const b200_keypoint* kps1_un, const uint8_t* desc1, int n1,
                                       const b200_keypoint* kps2_un, const uint8_t* desc2, int n2, const float* bounds4,
                                       float* prev_matched, int window, float ratio, int check_ori, int32_t* matches12, int device) {
    if (n1 < 0 || n2 < 0 || window < 0) return fail(B200_EINVAL, "negative %s", "size");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (n1 == 0) return 0;
    if (!kps1_un || !desc1 || !prev_matched || !matches12 || !bounds4 || (n2 > 0 && (!kps2_un || !desc2))) return fail(B200_EINVAL, "null %s", "pointer");
    if (n2 == 0) { for (int i = 0; i < n1; i++) matches12[i] = -1; return 0; }
    const int row_cap = std::min(n2, 4096);
    DevBuf k1, d1, k2, d2, un_cnt, cs, ci, prev, q3, lv2, cand, cnt, dist, m12, m21, mdist, rotbin, res;
    if ((rc = k1.upload(kps1_un, (size_t)n1 * sizeof(b200_keypoint))) || (rc = d1.upload(desc1, (size_t)n1 * 32)) ||
        (rc = k2.upload(kps2_un, (size_t)n2 * sizeof(b200_keypoint))) || (rc = d2.upload(desc2, (size_t)n2 * 32)) ||
        (rc = un_cnt.upload(&n2, 4)) || (rc = cs.alloc((size_t)(64 * 48 + 1) * 4)) || (rc = ci.alloc((size_t)n2 * 4)) ||
        (rc = prev.upload(prev_matched, (size_t)n1 * 8)) || (rc = q3.alloc((size_t)n1 * 12)) || (rc = lv2.alloc((size_t)n1 * 8)) ||
        (rc = cand.alloc((size_t)n1 * row_cap * 4)) || (rc = cnt.alloc((size_t)n1 * 4)) || (rc = dist.alloc((size_t)n1 * row_cap * 4)) ||
        (rc = m12.alloc((size_t)n1 * 4)) || (rc = m21.alloc((size_t)n2 * 4)) || (rc = mdist.alloc((size_t)n2 * 4)) || (rc = rotbin.alloc((size_t)n1)) ||
        (rc = res.alloc(8)))
        return rc;
    if ((rc = b200_frame_assign_grid((const b200_keypoint*)k2.p, (const int32_t*)un_cnt.p, 1, n2, bounds4, (int32_t*)cs.p, (int32_t*)ci.p, device, ts))) return rc;
    B200_LAUNCH(k_init_queries, (n1 + 255) / 256, 256, 0, ts, (const b200_keypoint*)k1.p, n1, (const float*)prev.p, (float)window, (float*)q3.p, (int*)lv2.p);
    if ((rc = b200_frame_features_in_area((const b200_keypoint*)k2.p, (const int32_t*)cs.p, (const int32_t*)ci.p, bounds4, (const float*)q3.p,
                                          (const int32_t*)lv2.p, n1, (int32_t*)cand.p, (int32_t*)cnt.p, row_cap, device, ts)))
        return rc;
    B200_LAUNCH(k_init_dist, (n1 * 32 + 255) / 256, 256, 0, ts, (const ulonglong4*)d1.p, n1, (const ulonglong4*)d2.p, (const int*)cand.p, (const int*)cnt.p,
                row_cap, (int*)dist.p);
    B200_LAUNCH(k_init_resolve, 1, 32, 0, ts, (const b200_keypoint*)k1.p, n1, (const b200_keypoint*)k2.p, n2, (const int*)cand.p, (const int*)cnt.p,
                (const int*)dist.p, row_cap, ratio, 50, check_ori, (float*)prev.p, (int*)m12.p, (int*)m21.p, (int*)mdist.p, (unsigned char*)rotbin.p, (int*)res.p);
    B200_CUDA(cudaStreamSynchronize(ts));
    int r2[2] = {0, 0};
    B200_D2H(r2, res.p, 8);
    if (r2[1]) return fail(B200_ECAPACITY, "more than %s candidates in one search window", "4096");
    B200_D2H(matches12, m12.p, (size_t)n1 * 4);
    B200_D2H(prev_matched, prev.p, (size_t)n1 * 8);
    return r2[0];
}

int b200_match_by_projection_host(const b200_keypoint* kps_un, const uint8_t* desc, int n_frame, const float* bounds4, uint8_t* occupied,
                                  const float* q_xyr, const int32_t* q_levels, const uint8_t* q_desc, const float* q_angle, const uint8_t* q_observed,
                                  int n_queries, int mode, float ratio, int check_ori, int th_high, int32_t* assign, int device) {
    if (th_high <= 0) th_high = 100;                            // TH_HIGH, ORBmatcher.cc:38
    if (n_frame < 0 || n_queries < 0 || mode < 0 || mode > 2) return fail(B200_EINVAL, "bad %s", "sizes or mode");
    const bool keyframe = mode == 2;                            // SearchByProjection(KeyFrame*, Scw, ...): mode 1 over KeyFrame::GetFeaturesInArea
    if (keyframe) mode = 1;
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (n_frame > 0 && (!kps_un || !desc || !occupied || !assign || !bounds4)) return fail(B200_EINVAL, "null %s", "frame pointer");
    for (int i = 0; i < n_frame; i++) assign[i] = -1;
    if (n_queries == 0 || n_frame == 0) return 0;
    if (!q_xyr || !q_levels || !q_desc || !q_angle || !q_observed) return fail(B200_EINVAL, "null %s", "query pointer");
    const int row_cap = std::min(n_frame, 4096);
    DevBuf k2, d2, ncnt, cs, ci, occ, q3, lv2, qd, qa, qo, cand, cnt, dist, asg, eb, ei, res;
    if ((rc = k2.upload(kps_un, (size_t)n_frame * sizeof(b200_keypoint))) || (rc = d2.upload(desc, (size_t)n_frame * 32)) || (rc = ncnt.upload(&n_frame, 4)) ||
        (rc = cs.alloc((size_t)(64 * 48 + 1) * 4)) || (rc = ci.alloc((size_t)n_frame * 4)) || (rc = occ.upload(occupied, (size_t)n_frame)) ||
        (rc = q3.upload(q_xyr, (size_t)n_queries * 12)) || (rc = lv2.upload(q_levels, (size_t)n_queries * 8)) || (rc = qd.upload(q_desc, (size_t)n_queries * 32)) ||
        (rc = qa.upload(q_angle, (size_t)n_queries * 4)) || (rc = qo.upload(q_observed, (size_t)n_queries)) ||
        (rc = cand.alloc((size_t)n_queries * row_cap * 4)) || (rc = cnt.alloc((size_t)n_queries * 4)) || (rc = dist.alloc((size_t)n_queries * row_cap * 4)) ||
        (rc = asg.alloc((size_t)n_frame * 4)) || (rc = eb.alloc((size_t)n_queries * 4)) || (rc = ei.alloc((size_t)n_queries * 4)) || (rc = res.alloc(8)))
        return rc;
    if ((rc = b200_frame_assign_grid((const b200_keypoint*)k2.p, (const int32_t*)ncnt.p, 1, n_frame, bounds4, (int32_t*)cs.p, (int32_t*)ci.p, device, ts))) return rc;
    if ((rc = (keyframe ? b200_keyframe_features_in_area : b200_frame_features_in_area)(
             (const b200_keypoint*)k2.p, (const int32_t*)cs.p, (const int32_t*)ci.p, bounds4, (const float*)q3.p, (const int32_t*)lv2.p, n_queries,
             (int32_t*)cand.p, (int32_t*)cnt.p, row_cap, device, ts)))
        return rc;
    B200_LAUNCH(k_init_dist, (n_queries * 32 + 255) / 256, 256, 0, ts, (const ulonglong4*)qd.p, n_queries, (const ulonglong4*)d2.p, (const int*)cand.p,
                (const int*)cnt.p, row_cap, (int*)dist.p);
    B200_LAUNCH(k_proj_resolve, 1, 32, 0, ts, (const b200_keypoint*)k2.p, n_frame, (const int*)cand.p, (const int*)cnt.p, (const int*)dist.p, row_cap,
                (const float*)qa.p, (const unsigned char*)qo.p, n_queries, mode, ratio, th_high, check_ori, (unsigned char*)occ.p, (int*)asg.p, (int*)eb.p, (int*)ei.p,
                (int*)res.p);
    B200_CUDA(cudaStreamSynchronize(ts));
    int r2[2] = {0, 0};
    B200_D2H(r2, res.p, 8);
    if (r2[1]) return fail(B200_ECAPACITY, "more than %s candidates in one search window", "4096");
    B200_D2H(assign, asg.p, (size_t)n_frame * 4);
    B200_D2H(occupied, occ.p, (size_t)n_frame);
    return r2[0];
}

int b200_match_by_bow_host(const uint8_t* q_desc, const float* q_angle, int n_q, const uint8_t* c_desc, const float* c_angle, int n_c,
                           const int32_t* grp_q_ofs, const int32_t* q_idx, const int32_t* grp_c_ofs, const int32_t* c_idx, int n_groups,
                           int mode, float ratio, int th_low, int check_ori, int32_t* out, int device) {
    if (th_low <= 0) th_low = 50;                               // TH_LOW, ORBmatcher.cc:39
    if (n_q < 0 || n_c < 0 || n_groups < 0 || (mode != 0 && mode != 1)) return fail(B200_EINVAL, "bad %s", "sizes or mode");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    const int n_out = mode == 0 ? n_c : n_q;
    if (n_out > 0 && !out) return fail(B200_EINVAL, "null %s", "output pointer");
    for (int i = 0; i < n_out; i++) out[i] = -1;
    if (n_groups == 0 || n_q == 0 || n_c == 0) return 0;
    if (!q_desc || !q_angle || !c_desc || !c_angle || !grp_q_ofs || !q_idx || !grp_c_ofs || !c_idx) return fail(B200_EINVAL, "null %s", "pointer");
    if (grp_q_ofs[0] != 0 || grp_c_ofs[0] != 0) return fail(B200_EINVAL, "group offsets must start at %s", "0");
    const int nq = grp_q_ofs[n_groups], nc = grp_c_ofs[n_groups];
    std::vector<int32_t> q_grp((size_t)std::max(nq, 0)), q_dofs((size_t)std::max(nq, 0) + 1);
    long long total = 0;
    for (int g = 0; g < n_groups; g++) {
        if (grp_q_ofs[g + 1] < grp_q_ofs[g] || grp_c_ofs[g + 1] < grp_c_ofs[g]) return fail(B200_EINVAL, "group offsets must be %s", "non-decreasing");
        const int cn = grp_c_ofs[g + 1] - grp_c_ofs[g];
        for (int q = grp_q_ofs[g]; q < grp_q_ofs[g + 1]; q++) { q_grp[q] = g; q_dofs[q] = (int32_t)total; total += cn; }
        if (total >= (1ll << 31)) return fail(B200_ECAPACITY, "more than %s (query, candidate) pairs", "2^31");
    }
    if (nq == 0 || nc == 0) return 0;
    q_dofs[nq] = (int32_t)total;
    for (int i = 0; i < nq; i++) if (q_idx[i] < 0 || q_idx[i] >= n_q) return fail(B200_EINVAL, "query index out of %s", "range");
    for (int i = 0; i < nc; i++) if (c_idx[i] < 0 || c_idx[i] >= n_c) return fail(B200_EINVAL, "candidate index out of %s", "range");
    DevBuf dq, aq, dc, ac, gco, qi, ci, qg, qo, dist, taken, o, eb, ei, res;
    if ((rc = dq.upload(q_desc, (size_t)n_q * 32)) || (rc = aq.upload(q_angle, (size_t)n_q * 4)) || (rc = dc.upload(c_desc, (size_t)n_c * 32)) ||
        (rc = ac.upload(c_angle, (size_t)n_c * 4)) || (rc = gco.upload(grp_c_ofs, (size_t)(n_groups + 1) * 4)) || (rc = qi.upload(q_idx, (size_t)nq * 4)) ||
        (rc = ci.upload(c_idx, (size_t)nc * 4)) || (rc = qg.upload(q_grp.data(), (size_t)nq * 4)) || (rc = qo.upload(q_dofs.data(), (size_t)(nq + 1) * 4)) ||
        (rc = dist.alloc((size_t)std::max(total, 1ll) * 4)) || (rc = taken.alloc((size_t)n_c)) || (rc = o.alloc((size_t)n_out * 4)) ||
        (rc = eb.alloc((size_t)nq * 4)) || (rc = ei.alloc((size_t)nq * 4)) || (rc = res.alloc(8)))
        return rc;
    B200_CUDA(cudaMemsetAsync(taken.p, 0, (size_t)n_c, ts));
    B200_LAUNCH(k_bow_dist, (nq * 32 + 255) / 256, 256, 0, ts, (const ulonglong4*)dq.p, (const ulonglong4*)dc.p, (const int*)qi.p, (const int*)qg.p,
                (const int*)gco.p, (const int*)ci.p, (const int*)qo.p, nq, (int*)dist.p);
    B200_LAUNCH(k_bow_resolve, 1, 32, 0, ts, (const float*)aq.p, (const float*)ac.p, (const int*)qi.p, (const int*)qg.p, (const int*)gco.p, (const int*)ci.p,
                (const int*)qo.p, (const int*)dist.p, nq, mode, ratio, th_low, check_ori, (unsigned char*)taken.p, n_out, (int*)o.p, (int*)eb.p, (int*)ei.p,
                (int*)res.p);
    B200_CUDA(cudaStreamSynchronize(ts));
    int r = 0;
    B200_D2H(&r, res.p, 4);
    B200_D2H(out, o.p, (size_t)n_out * 4);
    return r;
}

int b200_distinctive_descriptors_host(const uint8_t* desc, const int32_t* ofs, int n_points, int32_t* best_idx, uint8_t* out_desc, int device) {
    if (n_points < 0) return fail(B200_EINVAL, "negative %s", "size");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (n_points == 0) return B200_OK;
    if (!ofs || !best_idx) return fail(B200_EINVAL, "null %s", "pointer");
    int max_n = 0;
    for (int p = 0; p < n_points; p++) {
        if (ofs[p + 1] < ofs[p] || ofs[0] != 0) return fail(B200_EINVAL, "bad %s", "observation offsets");
        max_n = std::max(max_n, ofs[p + 1] - ofs[p]);
    }
    const int total = ofs[n_points];
    if (total > 0 && !desc) return fail(B200_EINVAL, "null %s", "descriptor pointer");
    const int row_cap = max_n > 32 ? (max_n + 7) / 8 * 8 : 0;
    const size_t smem = (size_t)kMedoidWarps * row_cap * 2;
    if (smem > 200 * 1024) return fail(B200_ECAPACITY, "more than %s observations of one map point", "25600");
    DevBuf dd, dofs, db, dout;
    if ((rc = dd.upload(desc, (size_t)std::max(total, 1) * 32)) || (rc = dofs.upload(ofs, (size_t)(n_points + 1) * 4)) || (rc = db.alloc((size_t)n_points * 4)) ||
        (out_desc && (rc = dout.alloc((size_t)n_points * 32))))
        return rc;
    if (out_desc) B200_CUDA(cudaMemsetAsync(dout.p, 0, (size_t)n_points * 32, ts));
    if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(k_distinctive, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_LAUNCH(k_distinctive, (n_points + kMedoidWarps - 1) / kMedoidWarps, kMedoidWarps * 32, smem, ts, (const ulonglong4*)dd.p, (const int*)dofs.p, n_points,
                row_cap, (int*)db.p, out_desc ? (ulonglong4*)dout.p : nullptr);
    B200_CUDA(cudaStreamSynchronize(ts));
    B200_D2H(best_idx, db.p, (size_t)n_points * 4);
    if (out_desc) B200_D2H(out_desc, dout.p, (size_t)n_points * 32);
    return B200_OK;
}

int b200_hamming_matrix_host(const uint8_t* a, int na, const uint8_t* b, int nb, int32_t* dist, int device) {
    if (na < 0 || nb < 0) return fail(B200_EINVAL, "negative %s", "size");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (na == 0 || nb == 0) return B200_OK;
    if (!a || !b || !dist) return fail(B200_EINVAL, "null %s", "pointer");
    DevBuf da, db, dd;
    if ((rc = da.upload(a, (size_t)na * 32)) || (rc = db.upload(b, (size_t)nb * 32)) || (rc = dd.alloc((size_t)na * nb * 4))) return rc;
    dim3 grid(na, std::min((nb + 255) / 256, 65535));
    B200_LAUNCH(k_hamming_matrix, grid, 256, 0, ts, (const ulonglong4*)da.p, na, (const ulonglong4*)db.p, nb, (int*)dd.p);
    B200_CUDA(cudaStreamSynchronize(ts));
    B200_D2H(dist, dd.p, (size_t)na * nb * 4);
    return B200_OK;
}

int b200_match_candidates_host(const uint8_t* query_desc, int nq, const uint8_t* train_desc, int nt,
                               const int32_t* cand_ofs, const int32_t* cand, int32_t* out_best_idx,
                               int32_t* out_best_dist, int32_t* out_second_dist, int device) {
    if (nq < 0 || nt < 0) return fail(B200_EINVAL, "negative %s", "size");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (nq == 0) return B200_OK;
    if (!query_desc || !cand_ofs || !out_best_idx || !out_best_dist || !out_second_dist) return fail(B200_EINVAL, "null %s", "pointer");
    const int total = cand_ofs[nq];
    for (int q = 0; q < nq; q++) if (cand_ofs[q + 1] < cand_ofs[q] || cand_ofs[q + 1] - cand_ofs[q] >= (1 << 20)) return fail(B200_EINVAL, "bad %s", "candidate offsets");
    for (int i = 0; i < total; i++) if (cand[i] < 0 || cand[i] >= nt) return fail(B200_EINVAL, "candidate index out of %s", "range");
    DevBuf dq, dt, dofs, dc, o1, o2, o3;
    if ((rc = dq.upload(query_desc, (size_t)nq * 32)) || (rc = dt.upload(train_desc, (size_t)nt * 32)) || (rc = dofs.upload(cand_ofs, (size_t)(nq + 1) * 4)) ||
        (rc = dc.upload(cand, (size_t)total * 4)) || (rc = o1.alloc((size_t)nq * 4)) || (rc = o2.alloc((size_t)nq * 4)) || (rc = o3.alloc((size_t)nq * 4)))
        return rc;
    B200_LAUNCH(k_match_candidates, (nq * 32 + 255) / 256, 256, 0, ts, (const ulonglong4*)dq.p, nq, (const ulonglong4*)dt.p, (const int*)dofs.p, (const int*)dc.p,
                (int*)o1.p, (int*)o2.p, (int*)o3.p);
    B200_CUDA(cudaStreamSynchronize(ts));
    B200_D2H(out_best_idx, o1.p, (size_t)nq * 4);
    B200_D2H(out_best_dist, o2.p, (size_t)nq * 4);
    B200_D2H(out_second_dist, o3.p, (size_t)nq * 4);
    return B200_OK;
}

int b200_match_for_triangulation_host(const b200_keypoint* kps1_un, const uint8_t* desc1, int n1, const b200_keypoint* kps2_un, const uint8_t* desc2, int n2,
                                      const int32_t* grp_q_ofs, const int32_t* q_idx, const int32_t* grp_c_ofs, const int32_t* c_idx, int n_groups,
                                      const float* F12, const float* epipole2, const float* scale_factors, const float* level_sigma2, int nlevels,
                                      int check_ori, int th_low, int32_t* matches12, int device) {
    if (th_low <= 0) th_low = 50;                               // TH_LOW, ORBmatcher.cc:39
    if (n1 < 0 || n2 < 0 || n_groups < 0 || nlevels < 1 || nlevels > 16) return fail(B200_EINVAL, "bad %s", "sizes");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (n1 > 0 && !matches12) return fail(B200_EINVAL, "null %s", "output pointer");
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    if (n_groups == 0 || n1 == 0 || n2 == 0) return 0;
    if (!kps1_un || !desc1 || !kps2_un || !desc2 || !grp_q_ofs || !q_idx || !grp_c_ofs || !c_idx || !F12 || !epipole2 || !scale_factors || !level_sigma2)
        return fail(B200_EINVAL, "null %s", "pointer");
    if (grp_q_ofs[0] != 0 || grp_c_ofs[0] != 0) return fail(B200_EINVAL, "group offsets must start at %s", "0");
    const int nq = grp_q_ofs[n_groups], nc = grp_c_ofs[n_groups];
    std::vector<int32_t> q_grp((size_t)std::max(nq, 0));
    for (int g = 0; g < n_groups; g++) {
        if (grp_q_ofs[g + 1] < grp_q_ofs[g] || grp_c_ofs[g + 1] < grp_c_ofs[g]) return fail(B200_EINVAL, "group offsets must be %s", "non-decreasing");
        if (grp_c_ofs[g + 1] - grp_c_ofs[g] >= (1 << 20)) return fail(B200_ECAPACITY, "more than %s candidates in one vocabulary node", "2^20");
        for (int q = grp_q_ofs[g]; q < grp_q_ofs[g + 1]; q++) q_grp[q] = g;
    }
    if (nq == 0 || nc == 0) return 0;
    for (int i = 0; i < nq; i++) if (q_idx[i] < 0 || q_idx[i] >= n1) return fail(B200_EINVAL, "query index out of %s", "range");
    for (int i = 0; i < nc; i++) if (c_idx[i] < 0 || c_idx[i] >= n2) return fail(B200_EINVAL, "candidate index out of %s", "range");
    TriangGeom g;
    for (int i = 0; i < 9; i++) g.F[i] = F12[i];
    g.ex = epipole2[0]; g.ey = epipole2[1];
    for (int i = 0; i < 16; i++) { g.sf[i] = scale_factors[std::min(i, nlevels - 1)]; g.sigma2[i] = level_sigma2[std::min(i, nlevels - 1)]; }
    DevBuf k1, d1, k2, d2, qi, qg, gco, ci, m12, rb, res;
    if ((rc = k1.upload(kps1_un, (size_t)n1 * sizeof(b200_keypoint))) || (rc = d1.upload(desc1, (size_t)n1 * 32)) ||
        (rc = k2.upload(kps2_un, (size_t)n2 * sizeof(b200_keypoint))) || (rc = d2.upload(desc2, (size_t)n2 * 32)) || (rc = qi.upload(q_idx, (size_t)nq * 4)) ||
        (rc = qg.upload(q_grp.data(), (size_t)nq * 4)) || (rc = gco.upload(grp_c_ofs, (size_t)(n_groups + 1) * 4)) || (rc = ci.upload(c_idx, (size_t)nc * 4)) ||
        (rc = m12.alloc((size_t)n1 * 4)) || (rc = rb.alloc((size_t)n1)) || (rc = res.alloc(4)))
        return rc;
    B200_CUDA(cudaMemsetAsync(m12.p, 0xff, (size_t)n1 * 4, ts));
    B200_CUDA(cudaMemsetAsync(rb.p, 0xff, (size_t)n1, ts));
    B200_LAUNCH(k_triang, (nq * 32 + 255) / 256, 256, 0, ts, (const b200_keypoint*)k1.p, (const ulonglong4*)d1.p, (const b200_keypoint*)k2.p, (const ulonglong4*)d2.p,
                (const int*)qi.p, (const int*)qg.p, (const int*)gco.p, (const int*)ci.p, nq, g, th_low, (int*)m12.p, (unsigned char*)rb.p);
    B200_LAUNCH(k_triang_finish, 1, 256, 0, ts, (int*)m12.p, (const unsigned char*)rb.p, n1, check_ori, (int*)res.p);
    B200_CUDA(cudaStreamSynchronize(ts));
    int r = 0;
    B200_D2H(&r, res.p, 4);
    B200_D2H(matches12, m12.p, (size_t)n1 * 4);
    return r;
}

int b200_match_kf_radius_host(const b200_keypoint* kps_un, const uint8_t* desc, int n_kf, const float* bounds4, const float* q_xyr, const int32_t* q_level,
                              const uint8_t* q_desc, int n_queries, const float* inv_level_sigma2, int nlevels, double chi2, int32_t* best_idx,
                              int32_t* best_dist, int device) {
    if (n_kf < 0 || n_queries < 0 || nlevels < 1 || nlevels > 16) return fail(B200_EINVAL, "bad %s", "sizes");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (n_queries == 0) return B200_OK;
    if (!best_idx || !best_dist) return fail(B200_EINVAL, "null %s", "output pointer");
    for (int q = 0; q < n_queries; q++) { best_idx[q] = -1; best_dist[q] = 256; }
    if (n_kf == 0) return B200_OK;
    if (!kps_un || !desc || !bounds4 || !q_xyr || !q_level || !q_desc || (chi2 > 0 && !inv_level_sigma2)) return fail(B200_EINVAL, "null %s", "pointer");
    std::vector<int32_t> lv((size_t)n_queries * 2);
    for (int q = 0; q < n_queries; q++) {
        if (q_level[q] < 0 || q_level[q] >= nlevels) return fail(B200_EINVAL, "predicted level out of %s", "range");
        lv[2 * q] = q_level[q] - 1; lv[2 * q + 1] = q_level[q];         // kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel -> continue
    }
    LevelTable lt;
    for (int i = 0; i < 16; i++) lt.inv_sigma2[i] = inv_level_sigma2 ? inv_level_sigma2[std::min(i, nlevels - 1)] : 1.0f;
    const int row_cap = std::min(n_kf, 4096);
    DevBuf k2, d2, ncnt, cs, ci, q3, lv2, qd, cand, cnt, bi, bd, ovf;
    if ((rc = k2.upload(kps_un, (size_t)n_kf * sizeof(b200_keypoint))) || (rc = d2.upload(desc, (size_t)n_kf * 32)) || (rc = ncnt.upload(&n_kf, 4)) ||
        (rc = cs.alloc((size_t)(64 * 48 + 1) * 4)) || (rc = ci.alloc((size_t)n_kf * 4)) || (rc = q3.upload(q_xyr, (size_t)n_queries * 12)) ||
        (rc = lv2.upload(lv.data(), (size_t)n_queries * 8)) || (rc = qd.upload(q_desc, (size_t)n_queries * 32)) ||
        (rc = cand.alloc((size_t)n_queries * row_cap * 4)) || (rc = cnt.alloc((size_t)n_queries * 4)) || (rc = bi.alloc((size_t)n_queries * 4)) ||
        (rc = bd.alloc((size_t)n_queries * 4)) || (rc = ovf.alloc(4)))
        return rc;
    B200_CUDA(cudaMemsetAsync(ovf.p, 0, 4, ts));
    if ((rc = b200_frame_assign_grid((const b200_keypoint*)k2.p, (const int32_t*)ncnt.p, 1, n_kf, bounds4, (int32_t*)cs.p, (int32_t*)ci.p, device, ts))) return rc;
    if ((rc = b200_keyframe_features_in_area((const b200_keypoint*)k2.p, (const int32_t*)cs.p, (const int32_t*)ci.p, bounds4, (const float*)q3.p,
                                             (const int32_t*)lv2.p, n_queries, (int32_t*)cand.p, (int32_t*)cnt.p, row_cap, device, ts)))
        return rc;
    B200_LAUNCH(k_radius_best, (n_queries * 32 + 255) / 256, 256, 0, ts, (const b200_keypoint*)k2.p, (const ulonglong4*)d2.p, (const float*)q3.p,
                (const ulonglong4*)qd.p, (const int*)cand.p, (const int*)cnt.p, row_cap, n_queries, lt, chi2, (int*)bi.p, (int*)bd.p, (int*)ovf.p);
    B200_CUDA(cudaStreamSynchronize(ts));
    int o = 0;
    B200_D2H(&o, ovf.p, 4);
    if (o) return fail(B200_ECAPACITY, "more than %s candidates in one search window", "4096");
    B200_D2H(best_idx, bi.p, (size_t)n_queries * 4);
    B200_D2H(best_dist, bd.p, (size_t)n_queries * 4);
    return B200_OK;
}

static int fill_proj_geom(ProjGeom& g, const float* Rcw, const float* tcw, const float* Ow, const float* sR, const float* tt, const float* cam4, const float* bounds4,
                          const float* normal, float th, const float* scale_factors, const float* level_thresholds, int nlevels) {
    if (!Rcw || !tcw || !cam4 || !bounds4 || !scale_factors || (nlevels > 1 && !level_thresholds) || (!sR && !Ow) || (sR && !tt)) return fail(B200_EINVAL, "null %s", "pointer");
    memset(&g, 0, sizeof(g));
    for (int i = 0; i < 9; i++) { g.R[i] = Rcw[i]; if (sR) g.sR[i] = sR[i]; }
    for (int i = 0; i < 3; i++) { g.t[i] = tcw[i]; if (Ow) g.Ow[i] = Ow[i]; if (sR) g.tt[i] = tt[i]; }
    g.fx = cam4[0]; g.fy = cam4[1]; g.cx = cam4[2]; g.cy = cam4[3];
    g.min_x = (float)(int)bounds4[0]; g.max_x = (float)(int)bounds4[1]; g.min_y = (float)(int)bounds4[2]; g.max_y = (float)(int)bounds4[3];   // int members of KeyFrame
    g.th = th; g.nlevels = nlevels; g.sim3 = sR ? 1 : 0; g.has_normal = normal ? 1 : 0;
    for (int i = 0; i < nlevels; i++) g.sf[i] = scale_factors[i];
    for (int i = 0; i + 1 < nlevels; i++) g.thr[i] = level_thresholds[i];
    return B200_OK;
}

int b200_kf_search_points_host(const b200_keypoint* kps_un, const uint8_t* desc, int n_kf, const float* bounds4, const float* Rcw, const float* tcw, const float* Ow,
                               const float* sR, const float* tt, const float* cam4, const float* pos, const float* normal, const float* minmax, const uint8_t* q_desc,
                               const uint8_t* skip, int n, float th, const float* scale_factors, const float* inv_level_sigma2, const float* level_thresholds,
                               int nlevels, double chi2, uint8_t* valid, int32_t* best_idx, int32_t* best_dist, float* q_xyr, int32_t* level, int device) {
    if (n < 0 || n_kf < 0 || nlevels < 1 || nlevels > 16) return fail(B200_EINVAL, "bad %s", "sizes");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (n == 0) return B200_OK;
    if (!pos || !minmax || !q_desc || !valid || !best_idx || !best_dist || (n_kf > 0 && (!kps_un || !desc)) || (chi2 > 0 && !inv_level_sigma2))
        return fail(B200_EINVAL, "null %s", "pointer");
    ProjGeom g;
    if ((rc = fill_proj_geom(g, Rcw, tcw, Ow, sR, tt, cam4, bounds4, normal, th, scale_factors, level_thresholds, nlevels))) return rc;
    LevelTable lt;
    for (int i = 0; i < 16; i++) lt.inv_sigma2[i] = inv_level_sigma2 ? inv_level_sigma2[std::min(i, nlevels - 1)] : 1.0f;
    const int row_cap = std::max(1, std::min(n_kf, 4096));
    DevBuf dp, dn, dm, dsk, dv, dq, dl, dl2, k2, d2, ncnt, cs, ci, qd, cand, cnt, bi, bd, ovf;
    if ((rc = dp.upload(pos, (size_t)n * 12)) || (normal && (rc = dn.upload(normal, (size_t)n * 12))) || (rc = dm.upload(minmax, (size_t)n * 8)) ||
        (skip && (rc = dsk.upload(skip, (size_t)n))) || (rc = dv.alloc((size_t)n)) || (rc = dq.alloc((size_t)n * 12)) || (rc = dl.alloc((size_t)n * 4)) ||
        (rc = dl2.alloc((size_t)n * 8)) || (rc = qd.upload(q_desc, (size_t)n * 32)) || (rc = bi.alloc((size_t)n * 4)) || (rc = bd.alloc((size_t)n * 4)) || (rc = ovf.alloc(4)))
        return rc;
    B200_LAUNCH(k_kf_project, (n + 255) / 256, 256, 0, ts, (const float*)dp.p, (const float*)dn.p, (const float*)dm.p, n, g, (unsigned char*)dv.p, (float*)dq.p, (int*)dl.p,
                (const unsigned char*)dsk.p, (int*)dl2.p);
    if (n_kf > 0) {
        if ((rc = k2.upload(kps_un, (size_t)n_kf * sizeof(b200_keypoint))) || (rc = d2.upload(desc, (size_t)n_kf * 32)) || (rc = ncnt.upload(&n_kf, 4)) ||
            (rc = cs.alloc((size_t)(64 * 48 + 1) * 4)) || (rc = ci.alloc((size_t)n_kf * 4)) || (rc = cand.alloc((size_t)n * row_cap * 4)) || (rc = cnt.alloc((size_t)n * 4)))
            return rc;
        B200_CUDA(cudaMemsetAsync(ovf.p, 0, 4, ts));
        if ((rc = b200_frame_assign_grid((const b200_keypoint*)k2.p, (const int32_t*)ncnt.p, 1, n_kf, bounds4, (int32_t*)cs.p, (int32_t*)ci.p, device, ts))) return rc;
        if ((rc = b200_keyframe_features_in_area((const b200_keypoint*)k2.p, (const int32_t*)cs.p, (const int32_t*)ci.p, bounds4, (const float*)dq.p,
                                                 (const int32_t*)dl2.p, n, (int32_t*)cand.p, (int32_t*)cnt.p, row_cap, device, ts)))
            return rc;
        B200_LAUNCH(k_radius_best, (n * 32 + 255) / 256, 256, 0, ts, (const b200_keypoint*)k2.p, (const ulonglong4*)d2.p, (const float*)dq.p, (const ulonglong4*)qd.p,
                    (const int*)cand.p, (const int*)cnt.p, row_cap, n, lt, chi2, (int*)bi.p, (int*)bd.p, (int*)ovf.p);
    }
    B200_CUDA(cudaStreamSynchronize(ts));
    if (n_kf > 0) {
        int o = 0;
        B200_D2H(&o, ovf.p, 4);
        if (o) return fail(B200_ECAPACITY, "more than %s candidates in one search window", "4096");
        B200_D2H(best_idx, bi.p, (size_t)n * 4);
        B200_D2H(best_dist, bd.p, (size_t)n * 4);
    } else for (int i = 0; i < n; i++) { best_idx[i] = -1; best_dist[i] = 256; }
    B200_D2H(valid, dv.p, (size_t)n);
    if (q_xyr) B200_D2H(q_xyr, dq.p, (size_t)n * 12);
    if (level) B200_D2H(level, dl.p, (size_t)n * 4);
    return B200_OK;
}

int b200_kf_project_host(const float* Rcw, const float* tcw, const float* Ow, const float* sR, const float* tt, const float* cam4, const float* bounds4,
                         const float* pos, const float* normal, const float* minmax, int n, float th, const float* scale_factors,
                         const float* level_thresholds, int nlevels, uint8_t* valid, float* q_xyr, int32_t* level, int device) {
    if (n < 0 || nlevels < 1 || nlevels > 16) return fail(B200_EINVAL, "bad %s", "sizes");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    if (n == 0) return B200_OK;
    if (!pos || !minmax || !valid || !q_xyr || !level) return fail(B200_EINVAL, "null %s", "pointer");
    ProjGeom g;
    if ((rc = fill_proj_geom(g, Rcw, tcw, Ow, sR, tt, cam4, bounds4, normal, th, scale_factors, level_thresholds, nlevels))) return rc;
    DevBuf dp, dn, dm, dv, dq, dl;
    if ((rc = dp.upload(pos, (size_t)n * 12)) || (normal && (rc = dn.upload(normal, (size_t)n * 12))) || (rc = dm.upload(minmax, (size_t)n * 8)) ||
        (rc = dv.alloc((size_t)n)) || (rc = dq.alloc((size_t)n * 12)) || (rc = dl.alloc((size_t)n * 4)))
        return rc;
    B200_LAUNCH(k_kf_project, (n + 255) / 256, 256, 0, ts, (const float*)dp.p, (const float*)dn.p, (const float*)dm.p, n, g, (unsigned char*)dv.p, (float*)dq.p, (int*)dl.p);
    B200_CUDA(cudaStreamSynchronize(ts));
    B200_D2H(valid, dv.p, (size_t)n);
    B200_D2H(q_xyr, dq.p, (size_t)n * 12);
    B200_D2H(level, dl.p, (size_t)n * 4);
    return B200_OK;
}

}  // extern "C"
