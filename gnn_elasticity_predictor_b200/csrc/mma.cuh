// Warp-level tensor-core helpers for the edge-attention kernels (sm_100a, bf16 operands, fp32 accumulate).
//
// The per-edge contractions of the streaming path (<qt_i,t , f_ij>, sum_j a_ij f_ij, ...) have a different
// right-hand operand for every TARGET ROW, so they are 16-edge x 256-channel x 8-column problems: far too
// small for tcgen05 (M = 128 rows sharing one B), exactly the shape of mma.sync.m16n8k16.  The kernels stay
// bound by the gathers; these helpers only take the FMA work off the CUDA cores.
//
// Fragment conventions (PTX ISA, m16n8k16 .bf16):  g = lane / 4, q = lane % 4
//   A (16 x 16, row):  a0 = (row g,   k 2q..2q+1)   a1 = (row g+8, k 2q..2q+1)
//                      a2 = (row g,   k 2q+8..+9)   a3 = (row g+8, k 2q+8..+9)
//   B (16 x 8,  col):  b0 = (k 2q..2q+1, n g)       b1 = (k 2q+8..+9, n g)
//   C (16 x 8):        c0,c1 = (row g, n 2q..2q+1)  c2,c3 = (row g+8, n 2q..2q+1)
#pragma once

#include "stream.cuh"

namespace alignn {

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// transpose of an 8x8 b16 matrix held in fragment layout (lane holds row g, cols 2q..2q+1)
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

// four transposed 8x8 b16 matrices from shared memory; lane l supplies the address of row (l & 7) of matrix (l >> 3)
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t smem_addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_addr));
}

// four 8x8 b16 matrices, not transposed: with lane l pointing at row (l & 7) + 8 * ((l >> 3) & 1), column 8 * (l >> 4) of a
// row-major 16 x 16 tile the four registers are the A fragment (a0..a3) of mma.m16n8k16
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t smem_addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_addr));
}

__device__ __forceinline__ uint4 lds128(uint32_t smem_addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t smem_addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr));
    return v;
}
__device__ __forceinline__ void sts32f(uint32_t smem_addr, float x) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_addr), "f"(x) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t smem_addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// 16-byte asynchronous global -> shared copy (L1 bypass); completion is tracked per thread in commit groups
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// generic-proxy writes/reads of shared memory ordered against later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// max / sum over the 8 row groups g (lanes with equal q)
__device__ __forceinline__ float colmax8(float x) {
    x = fmaxf(x, __shfl_xor_sync(FULL, x, 4));
    x = fmaxf(x, __shfl_xor_sync(FULL, x, 8));
    x = fmaxf(x, __shfl_xor_sync(FULL, x, 16));
    return x;
}
__device__ __forceinline__ float colsum8(float x) {
    x += __shfl_xor_sync(FULL, x, 4);
    x += __shfl_xor_sync(FULL, x, 8);
    x += __shfl_xor_sync(FULL, x, 16);
    return x;
}

// Philox keep-scales of the 4 heads of edge position `pos` (one counter block per edge)
__device__ __forceinline__ void dropout_scale4(uint64_t seed, uint64_t offset, uint64_t pos, float p, float inv_keep,
                                               float (&out)[4]) {
    const Philox4 r = philox4x32_10(seed, offset, pos);
    const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) out[t] = ((float)(bits[t] >> 8) * (1.0f / 16777216.0f)) < p ? 0.0f : inv_keep;
}

// Keep-scales of heads (hsel, hsel + 1), hsel = 2 * (q & 1), for the two edges a lane owns in the m16n8 fragment layout
// (positions pos + g and pos + g + 8).  One Philox block yields the 4 heads of one edge, and the four lanes of a quad need
// the same two edges: lanes q < 2 evaluate edge g, lanes q >= 2 edge g + 8, and partners (q ^ 2: same head pair) swap.
// Half the generator work of calling dropout_scale4 twice per lane; identical masks.  Warp-uniform call sites only.
__device__ __forceinline__ void dropout_scale_quad(uint64_t seed, uint64_t offset, uint64_t pos, int g, int q, float p,
                                                   float inv_keep, float &lo0, float &lo1, float &hi0, float &hi1) {
    const bool second = q >= 2;
    float d[4];
    dropout_scale4(seed, offset, pos + (uint64_t)(g + (second ? 8 : 0)), p, inv_keep, d);
    const bool upper = (q & 1) != 0;
    const float m0 = upper ? d[2] : d[0], m1 = upper ? d[3] : d[1];
    const float o0 = __shfl_xor_sync(FULL, m0, 2), o1 = __shfl_xor_sync(FULL, m1, 2);
    lo0 = second ? o0 : m0;
    lo1 = second ? o1 : m1;
    hi0 = second ? m0 : o0;
    hi1 = second ? m1 : o1;
}

}  // namespace alignn
