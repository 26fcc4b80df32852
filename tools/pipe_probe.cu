// tools/pipe_probe.cu -- which pipe executes what on sm_100a, and how fast: warp-instructions per cycle per SM sub-partition for the packed 16-bit
// min / max forms FAST scoring can be written in, alone and mixed.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>

#define REP 64
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned seed, long long* cyc) {
    unsigned a[8];
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x10203;
    unsigned b = seed ^ 0x5a5a5a5a, c = seed + 77;
    __half2 h[8];
    for (int i = 0; i < 8; i++) h[i] = __halves2half2(__int2half_rn((threadIdx.x + i) & 255), __int2half_rn((threadIdx.x * 3 + i) & 255));
    const __half2 hb = __halves2half2(__int2half_rn(seed & 255), __int2half_rn((seed >> 8) & 255));
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 256; r++) {
#pragma unroll
        for (int u = 0; u < REP / 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) a[i] = __vimin3_s16x2(a[i], b, c);                       // VIMNMX3.S16x2
                if (MODE == 1) h[i] = __hmin2(h[i], hb);                                 // HMNMX2
                if (MODE == 2) a[i] = __vmins2(a[i], b);                                 // VIMNMX.S16x2 (2 inputs)
                if (MODE == 3) a[i] = a[i] * b + c;                                      // IMAD
                if (MODE == 4) a[i] = __byte_perm(a[i], b, c);                           // PRMT
                if (MODE == 5) { if (i & 1) a[i] = __vimin3_s16x2(a[i], b, c); else h[i] = __hmin2(h[i], hb); }          // VIMNMX3 + HMNMX2
                if (MODE == 6) { if (i & 1) a[i] = __vimin3_s16x2(a[i], b, c); else a[i] = a[i] * b + c; }                // VIMNMX3 + IMAD
                if (MODE == 7) h[i] = __hfma2(h[i], hb, hb);                             // HFMA2
                if (MODE == 8) { if (i & 1) a[i] = __vimin3_s16x2(a[i], b, c); else h[i] = __hfma2(h[i], hb, hb); }       // VIMNMX3 + HFMA2
                if (MODE == 9) a[i] = __vimax3_u16x2(a[i], b, c);                        // VIMNMX3.U16x2
                if (MODE == 10) a[i] = max(min((int)a[i], (int)b), (int)c);              // IMNMX x2 (or VIMNMX3 32-bit)
            }
        b += 3; c ^= b;
    }
    const long long t1 = clock64();
    unsigned s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + __half2float(__low2half(h[i])) + __half2float(__high2half(h[i]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name) {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&cyc, 148 * 4 * 8);
    k<MODE><<<148 * 4, 256>>>(out, 12345, cyc);      // 4 CTAs x 8 warps per SM: 8 warps per sub-partition
    cudaDeviceSynchronize();
    k<MODE><<<148 * 4, 256>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148 * 4; i++) avg += h[i]; avg /= 148 * 4;
    // per sub-partition: 8 warps x 256 x REP instructions in `avg` cycles (the four CTAs of an SM run concurrently)
    printf("%-28s %8.0f cycles  -> %.3f warp-instr / clk / SMSP (%s)\n", name, avg, 8.0 * 256 * REP / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("VIMNMX3.S16x2"); run<9>("VIMNMX3.U16x2"); run<2>("VIMNMX.S16x2 (2-input)"); run<10>("IMNMX min+max 32-bit");
    run<1>("HMNMX2"); run<7>("HFMA2"); run<3>("IMAD"); run<4>("PRMT");
    run<5>("VIMNMX3 + HMNMX2 (1:1)"); run<6>("VIMNMX3 + IMAD (1:1)"); run<8>("VIMNMX3 + HFMA2 (1:1)");
    return 0;
}
