// CUDA programming guide TMA example (2D int tile), to check the environment
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cstdio>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#ifndef ELT
#define ELT int
#define DT CU_TENSOR_MAP_DATA_TYPE_INT32
#endif
#ifndef SW
#define SW 32
#define SH 8
#endif
constexpr int SMEM_W = SW, SMEM_H = SH;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
  __shared__ alignas(128) ELT smem_buffer[SMEM_H][SMEM_W];
  #pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else { token = bar.arrive(); }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < SMEM_H * SMEM_W; i += blockDim.x) out[i] = (int)smem_buffer[i / SMEM_W][i % SMEM_W];
}
int main() {
  const int GW = GWV, GH = GHV;
  ELT* d; cudaMalloc(&d, GW * GH * sizeof(ELT));
  ELT* h = new ELT[GW * GH]; for (int i = 0; i < GW * GH; ++i) h[i] = (ELT)(i * 7 + (i >> 10));
  cudaMemcpy(d, h, GW * GH * sizeof(ELT), cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  CUtensorMap tm{};
  cuuint64_t size[2] = {GW, GH}; cuuint64_t stride[1] = {GW * sizeof(ELT)};
  cuuint32_t box[2] = {SMEM_W, SMEM_H}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, DT, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  int* out; cudaMalloc(&out, SMEM_H * SMEM_W * 4);
  kernel<<<1, 128>>>(tm, XV, YV, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) { int ho[SMEM_H * SMEM_W]; cudaMemcpy(ho, out, sizeof ho, cudaMemcpyDeviceToHost); printf("first %d expect %d\n", ho[0], (int)h[YV * GW + XV]); }
  return 0;
}
