// scratch: isolate the TMA tile load used by fast_cells_kernel
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cstdlib>
struct Maps { CUtensorMap m[16]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ Maps maps, const __grid_constant__ CUtensorMap single, const CUtensorMap* gmaps, int variant, int idx, uint8_t* out, int BW, int BH, int x, int y, int z) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    uint8_t* tile = raw + ((128 - (smem_u32(raw) & 127)) & 127);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (variant == 6) {           // warp-uniform control flow + elect.sync, CUTLASS style
        if (threadIdx.x < 32) {
            uint32_t pred = 0;
            asm volatile("{\n\t.reg .pred P1;\n\t.reg .b32 R;\n\telect.sync R|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
            if (pred) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(tile)), "l"(&single), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
            }
        }
    } else if (threadIdx.x == 0) {
        const CUtensorMap* mp = MODE == 0 ? &maps.m[0] : &maps.m[idx];
        if (variant == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
        } else if (variant == 1) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(mp), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
        } else if (variant == 2) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(&maps.m[4]), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
        } else if (variant == 4) {   // tensor map in global memory
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(gmaps + idx), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
        } else if (variant == 5) {   // single grid-constant map
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(&single), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
        } else if (variant == 3) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(tile)), "l"(out + 4096), "r"(BW * BH), "r"(smem_u32(&bar)) : "memory");
        }
    }
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 1;
    const int pitch = 704, rows = 518, frames = 2; const size_t slab = 2 * 1024 * 1024;
    uint8_t* d; cudaMalloc(&d, slab * frames);
    std::vector<uint8_t> h(slab * frames);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    Maps maps; memset(&maps, 0, sizeof maps);
    printf("driver fn %p\n", fn);
    const int BW = argc > 2 ? atoi(argv[2]) : 48, BH = argc > 3 ? atoi(argv[3]) : 38;
    const int dtype = argc > 4 ? atoi(argv[4]) : 0; const int l2p = argc > 5 ? atoi(argv[5]) : 0;
    for (int l = 0; l < 3; ++l) {
        cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)frames};
        cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)slab};
        cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d + 256 * l, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d -> %d\n", l, (int)r);
    }
    {   // 2D map in slot 4
        cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&maps.m[4], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 2d -> %d\n", (int)r);
    }
    for (int i = 0; i < 16; ++i) printf("%016llx ", (unsigned long long)((cuuint64_t*)&maps.m[0])[i]);
    printf("\n");
    uint8_t* out; cudaMalloc(&out, BW * BH + 8192);
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(Maps)); cudaMemcpy(gm, &maps, sizeof(Maps), cudaMemcpyHostToDevice);
    std::vector<uint8_t> ho(BW * BH);
    for (int mode = 0; mode < 2; ++mode) {
        const int x = 47 + 31 * 3, y = 35 + 32, z = 1, idx = mode ? 2 : 0;
        if (mode == 0) k<0><<<1, 64, BW * BH + 256>>>(maps, maps.m[idx], gm, variant, idx, out, BW, BH, x, y, z);
        else k<1><<<1, 64, BW * BH + 256>>>(maps, maps.m[idx], gm, variant, idx, out, BW, BH, x, y, z);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(ho.data(), out, BW * BH, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < BH; ++r) for (int c = 0; c < BW; ++c) {
            uint8_t want = h[slab * z + 256 * idx + (size_t)(y + r) * pitch + x + c];
            bad += ho[r * BW + c] != want;
        }
        printf("mode %d mismatches %d\n", mode, bad);
    }
    return 0;
}
