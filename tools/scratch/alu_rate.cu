// Issue rates (per clock per SM sub-partition) of the integer instructions the FAST kernel is made of, sm_100a.
// 8 independent chains per thread, inline PTX so each mode is one SASS opcode.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rate alu_rate.cu && ./alu_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, unsigned seed, int iters) {
    __shared__ unsigned char sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (unsigned char)(i * 7 + seed);
    __syncthreads();
    unsigned a[8], b = seed ^ threadIdx.x, c = seed * 3 + blockIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + j * 0x01010101u + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c));
            else if (MODE == 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
            else if (MODE == 2) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
            else if (MODE == 3) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
            else if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(b));
            else if (MODE == 5) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(sm) + (a[j] & 4095u))); a[j] += v; }
            else if (MODE == 6) asm volatile("popc.b32 %0, %0;" : "+r"(a[j]));
            else { unsigned p; asm volatile("{.reg .pred q; setp.gt.u32 q, %1, %2; selp.u32 %0, %1, %2, q;}" : "=r"(p) : "r"(a[j]), "r"(b)); a[j] = p + 1; }
        }
        b += 0x00010001u; c ^= i;
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
    const char* names[8] = {"LOP3", "SHF.R.W", "PRMT", "VABSDIFF4 (+acc)", "IADD (add.u32)", "LDS.U8 + IADD", "POPC", "ISETP + SEL + IADD"};
    float ms[8] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters), run<5>(d, iters), run<6>(d, iters), run<7>(d, iters)};
    const double warp_inst = 148.0 * 8 * 8 * iters * 8.0;
    for (int m = 0; m < 8; ++m)
        printf("%-22s %.3f ms  -> %.2f statements per SM sub-partition per clock at 1.965 GHz\n", names[m], ms[m], warp_inst / ms[m] / 1e6 / (148 * 4 * 1.965));
    return 0;
}
