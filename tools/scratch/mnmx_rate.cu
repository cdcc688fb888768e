// Throughput of the packed min/max instructions the FAST score network uses (sm_100a), 8 independent chains per thread:
// VIMNMX3.S16x2 (3-input), VIMNMX.S16x2 (2-input), HMNMX2 (half2, 2-input) and mixes of them (do they share a pipe?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mnmx_rate mnmx_rate.cu && ./mnmx_rate
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<unsigned*>(&r);
}
template <int MODE>
__global__ void k(unsigned* out, unsigned seed, int iters) {
    unsigned a[8], b = (seed ^ threadIdx.x) & 0x03ff03ffu | 0x64006400u, c = (seed * 3 + blockIdx.x) & 0x03ff03ffu | 0x64006400u;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = ((seed + j * 0x01010101u + threadIdx.x) & 0x03ff03ffu) | 0x64006400u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = __vimax3_s16x2(a[j], b, c);
            else if (MODE == 1) a[j] = __vmaxs2(a[j], b);
            else if (MODE == 2) a[j] = hmax2u(a[j], b);
            else if (MODE == 3) a[j] = (j & 1) ? hmax2u(a[j], b) : __vimax3_s16x2(a[j], b, c);       // 4 + 4
            else if (MODE == 4) a[j] = (j & 1) ? hmax2u(a[j], b) : __vmaxs2(a[j], c);                 // 4 + 4
            else a[j] = a[j] * 3u + b;                                                                 // IMAD
        }
        b ^= 0x00010001u; c ^= (i & 1) * 0x00020002u;
    }
    unsigned r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r ^= a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> float run(unsigned* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 12345u, 100);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 12345u, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    const char* names[6] = {"VIMNMX3.S16x2", "VIMNMX.S16x2 (2-input)", "HMNMX2 (half2, 2-input)", "4 x VIMNMX3 + 4 x HMNMX2", "4 x VIMNMX + 4 x HMNMX2", "IMAD"};
    float ms[6] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters), run<5>(d, iters)};
    const double warp_inst = 148.0 * 8 * 8 /*warps*/ * iters * 8.0;
    for (int m = 0; m < 6; ++m)
        printf("%-28s %.3f ms  %.1f G warp-inst/s  (%.2f per SM sub-partition per clock at 1.965 GHz)\n", names[m], ms[m],
               warp_inst / ms[m] / 1e6, warp_inst / ms[m] / 1e6 / (148 * 4 * 1.965));
    return 0;
}
