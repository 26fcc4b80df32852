#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
struct Maps { CUtensorMap m[16]; };
struct Big { int pad[434]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ Big big, const __grid_constant__ Maps maps, const CUtensorMap* gmaps, const int* lv, int bw, int bh, uint8_t* out, int xoff, int xstep) {
    extern __shared__ __align__(1024) unsigned char tile[];
    __shared__ __align__(8) unsigned long long bar;
    const int level = lv[blockIdx.x];
    const CUtensorMap* mp = MODE == 0 ? &maps.m[0] : MODE == 1 ? &maps.m[level] : &gmaps[level];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bw * bh) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(mp)), "r"(xoff + xstep * (int)blockIdx.x), "r"(16), "r"(0), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" :: "r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[blockIdx.x * bw * bh + i] = tile[i] + (big.pad[0] & 0);
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1; const int xoff = argc > 2 ? atoi(argv[2]) : 15; const int xstep = argc > 3 ? atoi(argv[3]) : 32; const int nblk = argc > 4 ? atoi(argv[4]) : 8;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Fn enc = (Fn)p;
    const int w = 320, h = 240, n = 2, bw = 48, bh = 41;
    uint8_t* img; cudaMalloc(&img, w * h * n);
    std::vector<uint8_t> himg(w * h * n); for (size_t i = 0; i < himg.size(); i++) himg[i] = (uint8_t)(i * 7 + i / 320);
    cudaMemcpy(img, himg.data(), himg.size(), cudaMemcpyHostToDevice);
    Maps* mapsp = new Maps(); Maps& maps = *mapsp;
    cuuint64_t dims[3] = {w, h, n}, strides[2] = {w, (cuuint64_t)w * h}; cuuint32_t box[3] = {bw, bh, 1}, es[3] = {1, 1, 1};
    for (int l = 0; l < 2; l++) {
        CUresult r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d -> %d  (maps align %zu, addr%%64 = %zu)\n", l, (int)r, alignof(Maps), (size_t)((uintptr_t)&maps.m[l] % 64));
    }
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(Maps)); cudaMemcpy(gm, &maps, sizeof(Maps), cudaMemcpyHostToDevice);
    int hl[8] = {0, 0, 1, 1, 0, 1, 0, 1}; int* lv; cudaMalloc(&lv, 32); cudaMemcpy(lv, hl, 32, cudaMemcpyHostToDevice);
    uint8_t* out; cudaMalloc(&out, 8 * bw * bh);
    Big big; big.pad[0] = 0;
    for (int mode = 0; mode < 3; mode++) {
        if (only >= 0 && mode != only) continue;
        cudaMemset(out, 0, 8 * bw * bh);
        if (mode == 0) k<0><<<nblk, 128, bw * bh>>>(big, maps, gm, lv, bw, bh, out, xoff, xstep);
        if (mode == 1) k<1><<<nblk, 128, bw * bh>>>(big, maps, gm, lv, bw, bh, out, xoff, xstep);
        if (mode == 2) k<2><<<nblk, 128, bw * bh>>>(big, maps, gm, lv, bw, bh, out, xoff, xstep);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint8_t> ho(8 * bw * bh); cudaMemcpy(ho.data(), out, ho.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int b = 0; b < nblk; b++) for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++) {
            const int gx = xoff + xstep * b + x, gy = 16 + y;
            const uint8_t want = gx < w ? himg[gy * w + gx] : 0;
            bad += ho[(b * bh + y) * bw + x] != want;
        }
        printf("mode %d: %s, mismatches %d\n", mode, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
