// tools/pipe_probe2.cu -- issue rates of the instructions the FAST / threshold / blur kernels are built from, alone and in pairs, on sm_100a.
// Every operation is a volatile inline-PTX statement on its own register chain (8 chains per thread, 8 warps per sub-partition), so nothing can be folded;
// rates are normalised by FFMA measured the same way (one warp instruction per cycle per sub-partition), which removes any doubt about the clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe2.bin pipe_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define OPS(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
enum { VHMNMX = 100, FFMA = 0, VMIN2, VMIN3, HMNMX2, HFMA2, HFMA2RELU, HADD2, IMAD, PRMT, LOP3, VADD2, SHF, DP4A, POPC, IADD3, FMNMX, VMIN3_HMNMX2, VMIN3_HFMA2, VMIN3_IMAD, VMIN3_FFMA, VMIN2_HMNMX2, PRMT_IMAD, VMIN3_DP4A, PRMT_HFMA2, VMIN3_VMIN2, NMODES };
const char* kNames[] = {"FFMA", "VIMNMX.S16x2 (2 in)", "VIMNMX3.S16x2", "HMNMX2", "HFMA2", "HFMA2.RELU", "HADD2", "IMAD", "PRMT", "LOP3", "VIADD.16x2", "SHF", "IDP4A", "POPC", "IADD3", "FMNMX",
                        "VIMNMX3 + HMNMX2", "VIMNMX3 + HFMA2", "VIMNMX3 + IMAD", "VIMNMX3 + FFMA", "VIMNMX2 + HMNMX2", "PRMT + IMAD", "VIMNMX3 + IDP4A", "PRMT + HFMA2", "VIMNMX3 + VIMNMX2"};

template <int OP, int P> __device__ __forceinline__ void one(unsigned& x, unsigned b, unsigned c) {
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
    if (OP == VMIN2 && P == 0) asm volatile("min.s16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
    if (OP == VMIN2 && P == 1) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(x) : "r"(c));
    if (OP == VMIN3) asm volatile("{.reg .b32 t; min.s16x2 t, %0, %1; min.s16x2 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
    if (OP == HMNMX2 && P == 0) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
    if (OP == HMNMX2 && P == 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(c));
    if (OP == VHMNMX) asm volatile("{.reg .b32 t; min.f16x2 t, %0, %1; min.f16x2 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
    if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
    if (OP == HFMA2RELU) asm volatile("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
    if (OP == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
    if (OP == VADD2) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
    if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
    if (OP == DP4A) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
    if (OP == POPC) asm volatile("{.reg .b32 t; popc.b32 t, %0; add.u32 %0, t, %1;}" : "+r"(x) : "r"(b));
    if (OP == IADD3) asm volatile("{.reg .b32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
    if (OP == FMNMX && P == 0) asm volatile("min.f32 %0, %0, %1;" : "+r"(x) : "r"(b));
    if (OP == FMNMX && P == 1) asm volatile("max.f32 %0, %0, %1;" : "+r"(x) : "r"(c));
}

template <int A, int B>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned seed) {
    unsigned a[8], h[8];
    for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + 1) + i * 0x10203; h[i] = (a[i] >> 3) & 0x03ff03ffu; }
    unsigned b = seed ^ 0x005a005a, c = seed + 77;
#pragma unroll 1
    for (int r = 0; r < 512; r++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#define X(i) if (u & 1) one<A, 1>(a[i], b, c); else one<A, 0>(a[i], b, c); if (B >= 0) { if (u & 1) one<(B >= 0 ? B : 0), 1>(h[i], c, b); else one<(B >= 0 ? B : 0), 0>(h[i], c, b); }
            OPS(X)
#undef X
        }
        b += 3; c ^= b;
    }
    unsigned s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float g_ffma_ms = 0;
template <int A, int B> void run(const char* name) {
    unsigned* out;
    cudaMalloc(&out, 148 * 4 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<A, B><<<148 * 4, 256>>>(out, 12345);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) k<A, B><<<148 * 4, 256>>>(out, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    if (A == FFMA && B < 0) g_ffma_ms = ms;
    const double n = (B >= 0 ? 2.0 : 1.0);
    // FFMA alone issues 1 / clk / SMSP: the same instruction count taking t ms runs at g_ffma_ms / t; a pair counts both instructions
    printf("%-24s %8.3f ms   %.3f instr / clk / SMSP  (%s)\n", name, ms, n * g_ffma_ms / ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<FFMA, -1>(kNames[FFMA]); run<VMIN2, -1>(kNames[VMIN2]); run<VMIN3, -1>(kNames[VMIN3]); run<HMNMX2, -1>(kNames[HMNMX2]); run<HFMA2, -1>(kNames[HFMA2]);
    run<HFMA2RELU, -1>(kNames[HFMA2RELU]); run<HADD2, -1>(kNames[HADD2]); run<IMAD, -1>(kNames[IMAD]); run<PRMT, -1>(kNames[PRMT]); run<LOP3, -1>(kNames[LOP3]);
    run<VADD2, -1>(kNames[VADD2]); run<SHF, -1>(kNames[SHF]); run<DP4A, -1>(kNames[DP4A]); run<POPC, -1>("POPC + IADD"); run<IADD3, -1>(kNames[IADD3]); run<FMNMX, -1>(kNames[FMNMX]);
    run<VMIN3, HMNMX2>(kNames[VMIN3_HMNMX2]); run<VMIN3, HFMA2>(kNames[VMIN3_HFMA2]); run<VMIN3, IMAD>(kNames[VMIN3_IMAD]); run<VMIN3, FFMA>(kNames[VMIN3_FFMA]);
    run<VMIN2, HMNMX2>(kNames[VMIN2_HMNMX2]); run<PRMT, IMAD>(kNames[PRMT_IMAD]); run<VMIN3, DP4A>(kNames[VMIN3_DP4A]); run<PRMT, HFMA2>(kNames[PRMT_HFMA2]); run<VMIN3, VMIN2>(kNames[VMIN3_VMIN2]);
    run<VHMNMX, -1>("VHMNMX (3 in, f16x2)"); run<VMIN3, VHMNMX>("VIMNMX3 + VHMNMX"); run<VHMNMX, IMAD>("VHMNMX + IMAD"); run<VHMNMX, HFMA2>("VHMNMX + HFMA2");
    run<HMNMX2, HFMA2>("HMNMX2 + HFMA2"); run<LOP3, IMAD>("LOP3 + IMAD"); run<VADD2, HFMA2>("VIADD.16x2 + HFMA2"); run<HMNMX2, IMAD>("HMNMX2 + IMAD"); run<FMNMX, IMAD>("FMNMX + IMAD");
    return 0;
}
