// Shared device helpers for the ALIGNN hot-path kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/alignn_b200.h"

#define ALIGNN_CUDA_TRY(expr)                                         \
    do {                                                              \
        cudaError_t _e = (expr);                                      \
        if (_e != cudaSuccess) return ALIGNN_ERR_CUDA_BASE + (int)_e; \
    } while (0)

#define ALIGNN_LAUNCH_CHECK()                                         \
    do {                                                              \
        cudaError_t _e = cudaGetLastError();                          \
        if (_e != cudaSuccess) return ALIGNN_ERR_CUDA_BASE + (int)_e; \
    } while (0)

namespace alignn {

constexpr unsigned FULL = 0xffffffffu;
constexpr float LOG2E = 1.4426950408889634f;

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- 8-wide row fragments ---------------------------------------------------------------------------
// Every lane of a row group owns 8 consecutive channels: one 16-byte load for bf16, two for fp32.
struct F8 {
    float v[8];
};

__device__ __forceinline__ F8 ld8(const float *p) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    F8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

__device__ __forceinline__ F8 ld8(const __nv_bfloat16 *p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
    F8 r;
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __uint_as_float(w[i] << 16);
        r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
}

// streaming variant: read-once rows (edge projections) should not displace the gathered node rows in L1
__device__ __forceinline__ F8 ld8_stream(const float *p) {
    float4 a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
    F8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

__device__ __forceinline__ F8 ld8_stream(const __nv_bfloat16 *p) {
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    F8 r;
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __uint_as_float(w[i] << 16);
        r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
}

__device__ __forceinline__ void st8(float *p, const F8 &r) {
    *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *(reinterpret_cast<float4 *>(p) + 1) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

__device__ __forceinline__ void st8(__nv_bfloat16 *p, const F8 &r) {
    uint4 u;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t *>(&h);
    }
    u.x = w[0]; u.y = w[1]; u.z = w[2]; u.w = w[3];
    *reinterpret_cast<uint4 *>(p) = u;
}

__device__ __forceinline__ float ldf(const float *p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16 *p) {
    return __bfloat162float(__ldg(p));
}
__device__ __forceinline__ void stf(float *p, float x) { *p = x; }
__device__ __forceinline__ void stf(__nv_bfloat16 *p, float x) { *p = __float2bfloat16_rn(x); }

// butterfly sum over aligned groups of `width` lanes (width = power of two <= 32)
template <int WIDTH>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
    for (int off = WIDTH / 2; off > 0; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
    return x;
}
__device__ __forceinline__ float group_sum_rt(float x, int width) {
    for (int off = width >> 1; off > 0; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
    return x;
}
__device__ __forceinline__ float warp_sum(float x) { return group_sum<32>(x); }

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- Philox4x32-10: counter-based RNG so backward regenerates the forward's dropout masks -----------
struct Philox4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint64_t seed, uint64_t offset, uint64_t index) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32);
    uint32_t c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// keep-scale of element `index` of a dropout mask: 0 with probability p, else 1/(1-p).
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t offset, uint64_t index, float p,
                                               float inv_keep) {
    const Philox4 r = philox4x32_10(seed, offset, index >> 2);
    const uint32_t lane = (uint32_t)(index & 3);
    const uint32_t bits = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
    const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);  // [0,1)
    return u < p ? 0.0f : inv_keep;
}

}  // namespace alignn
