// Probe: TMA tile::gather4 (UTMALDG.2D.GATHER4) -- which tensor-map box shape does it take and where do the 4 gathered rows land?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_gather4 scripts/probes/tma_gather4.cu -lcuda   (run on the GPU box)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, int col, uint32_t bytes, uint16_t *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem), mb = (uint32_t)__cvta_generic_to_shared(&bar);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint16_t *>(smem)[i] = 0xFFFF;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(mb) : "memory");
    }
    // bounded wait so that a wrong byte count cannot hang the box
    uint32_t done = 0;
    for (int it = 0; it < 200000 && !done; ++it)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(smem)[i];
    if (threadIdx.x == 0) out[4096] = (uint16_t)done;
}

int main() {
    const int R = 64, C = 256;
    std::vector<__nv_bfloat16> h(R * C);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = __float2bfloat16((float)(r * 256 + c));   // exact up to 256? use small ints
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = __float2bfloat16((float)(r + (c % 64) * 0.0f + (c / 64) * 64));   // value = row + 64 * colblock (exact)
    __nv_bfloat16 *d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    uint16_t *out; cudaMalloc(&out, 4097 * 2);
    typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Enc enc = (Enc)p;
    struct Cfg { unsigned bx, by; CUtensorMapSwizzle sw; const char *name; } cfgs[] = {
        {64, 1, CU_TENSOR_MAP_SWIZZLE_128B, "box 64x1 SW128"}, {64, 4, CU_TENSOR_MAP_SWIZZLE_128B, "box 64x4 SW128"},
        {256, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "box 256x1 none"}, {64, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "box 64x1 none"}};
    for (auto &cf : cfgs) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {cf.bx, cf.by}, es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, cf.sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", cf.name, (int)r); continue; }
        cudaMemset(out, 0, 4097 * 2);
        k<<<1, 128, 8192>>>(map, 5, 17, 3, 40, 64, 4 * cf.bx * 2, out);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint16_t> ho(4097); cudaMemcpy(ho.data(), out, 4097 * 2, cudaMemcpyDeviceToHost);
        printf("%s: sync=%s done=%d\n", cf.name, cudaGetErrorString(e), (int)ho[4096]);
        if (e != cudaSuccess) return 1;
        // print, per 128-byte line of smem (64 bf16), the first element of every 16-byte chunk as float
        for (int line = 0; line < 8; ++line) {
            printf("  line %d:", line);
            for (int ch = 0; ch < 8; ++ch) {
                uint16_t v = ho[line * 64 + ch * 8];
                if (v == 0xFFFF) printf("   --"); else { uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4); printf(" %4.0f", f); }
            }
            printf("\n");
        }
    }
    return 0;
}
