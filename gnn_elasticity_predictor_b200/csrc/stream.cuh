// Warp-private TMA (cp.async.bulk) pipelines for the streaming edge-attention kernels (sm_100a).
//
// Every warp owns a contiguous range of target rows = a contiguous range of the CSR edge list and streams it
// through a ring of shared-memory stages.  Rows of the gathered operands (512 B in bf16, 1 KB in fp32) are
// fetched with one bulk-copy instruction per row (SASS: UBLKCP) that completes on an mbarrier, so the
// gathers cost neither registers nor LSU issue slots and run STAGES-1 stages ahead of the math.
#pragma once

#include "common.cuh"

namespace alignn {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(addr), "r"(parity)
        : "memory");
}

// global -> shared bulk copy of `bytes` (multiple of 16, both sides 16-byte aligned); completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- 8-wide fragments from shared memory -----------------------------------------------------------
__device__ __forceinline__ F8 lds8(const __nv_bfloat16 *p) {
    const uint4 u = *reinterpret_cast<const uint4 *>(p);
    F8 r;
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __uint_as_float(w[i] << 16);
        r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
}
__device__ __forceinline__ F8 lds8(const float *p) {
    const float4 a = *reinterpret_cast<const float4 *>(p);
    const float4 b = *(reinterpret_cast<const float4 *>(p) + 1);
    F8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

// ---- row-range partition --------------------------------------------------------------------------------
// cost(r) = rowptr[r] + KAPPA * r is strictly increasing; warp w of W owns rows [bound(w), bound(w+1)).
constexpr int ROW_KAPPA = 3;

__device__ __forceinline__ int64_t row_bound(const int32_t *__restrict__ rowptr, int64_t n_rows, int64_t target) {
    int64_t lo = 0, hi = n_rows;  // smallest r in [0, n_rows] with cost(r) >= target
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowptr + mid) + ROW_KAPPA * mid < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Walks the (row, position) sequence of one warp's edge range in stages of at most EPS edges that never
// cross a row boundary.  Producer and consumer each keep one cursor; both see the same sequence.
template <int EPS>
struct StageCursor {
    int64_t row, row_hi;   // current row, end of the warp's row range
    int pos, end;          // current edge position, end of the current row
    bool first;            // the next stage is the first stage of `row`

    __device__ __forceinline__ void init(const int32_t *__restrict__ rowptr, int64_t r0, int64_t r1) {
        row = r0;
        row_hi = r1;
        pos = r0 < r1 ? __ldg(rowptr + r0) : 0;
        end = pos;
        first = true;
        seek(rowptr);
    }
    // move `row` to the next row that has edges (or to row_hi)
    __device__ __forceinline__ void seek(const int32_t *__restrict__ rowptr) {
        while (row < row_hi) {
            end = __ldg(rowptr + row + 1);
            if (end > pos) break;
            ++row;
        }
    }
    __device__ __forceinline__ bool done() const { return row >= row_hi; }
    __device__ __forceinline__ int count() const { return min(EPS, end - pos); }
    __device__ __forceinline__ bool last() const { return end - pos <= EPS; }
    __device__ __forceinline__ void advance(const int32_t *__restrict__ rowptr) {
        pos += count();
        if (pos == end) {
            ++row;
            first = true;
            seek(rowptr);
        } else {
            first = false;
        }
    }
};

// Register window over an int32 array: lane l holds a[base + l] (cur) and a[base + 32 + l] (nxt).
struct IndexWindow {
    int cur, nxt;
    __device__ __forceinline__ void init(const int32_t *__restrict__ a, int base, int limit, int lane) {
        cur = base + lane < limit ? __ldg(a + base + lane) : 0;
        nxt = base + 32 + lane < limit ? __ldg(a + base + 32 + lane) : 0;
    }
    __device__ __forceinline__ void shift(const int32_t *__restrict__ a, int new_base, int limit, int lane) {
        cur = nxt;
        nxt = new_base + 32 + lane < limit ? __ldg(a + new_base + 32 + lane) : 0;
    }
    // value at window offset o in [0, 64); all lanes must call
    __device__ __forceinline__ int get(int o) const {
        const int a = __shfl_sync(FULL, cur, o & 31);
        const int b = __shfl_sync(FULL, nxt, o & 31);
        return o < 32 ? a : b;
    }
};

// ---- all-lane reduction of HEADS per-lane partials; lane l ends with the total of head l / (32/HEADS) -----
template <int HEADS>
__device__ __forceinline__ float reduce_heads(const float (&p)[HEADS], int lane);

template <>
__device__ __forceinline__ float reduce_heads<1>(const float (&p)[1], int) {
    return warp_sum(p[0]);
}
template <>
__device__ __forceinline__ float reduce_heads<2>(const float (&p)[2], int lane) {
    const bool hi = lane & 16;
    float a = (hi ? p[1] : p[0]) + __shfl_xor_sync(FULL, hi ? p[0] : p[1], 16);
    a += __shfl_xor_sync(FULL, a, 8);
    a += __shfl_xor_sync(FULL, a, 4);
    a += __shfl_xor_sync(FULL, a, 2);
    a += __shfl_xor_sync(FULL, a, 1);
    return a;
}
template <>
__device__ __forceinline__ float reduce_heads<4>(const float (&p)[4], int lane) {
    const bool hi = lane & 16, b3 = lane & 8;
    const float a0 = (hi ? p[2] : p[0]) + __shfl_xor_sync(FULL, hi ? p[0] : p[2], 16);
    const float a1 = (hi ? p[3] : p[1]) + __shfl_xor_sync(FULL, hi ? p[1] : p[3], 16);
    float b = (b3 ? a1 : a0) + __shfl_xor_sync(FULL, b3 ? a0 : a1, 8);
    b += __shfl_xor_sync(FULL, b, 4);
    b += __shfl_xor_sync(FULL, b, 2);
    b += __shfl_xor_sync(FULL, b, 1);
    return b;
}

}  // namespace alignn
