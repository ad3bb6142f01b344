// Fused node epilogue: beta-gated skip + LayerNorm + ReLU + dropout + residual (forward, backward).
//
// Replaces the tail of PyG TransformerConv.forward (beta gate against the skip projection) and
// `x + Dropout(ReLU(LayerNorm(out)))` of EdgeUpdateBlock / NodeUpdateBlock
// (reference scripts/train.py:316-317, 335-336).  The [rows, 3*hidden] concat that PyG feeds to
// lin_beta is never materialised: the gate logit is three fused dot products.
//
// HBM-bound elementwise/row-reduction work: each operand row is read once with 128-bit loads, row
// statistics live in registers (xor-shuffle reductions over the row's lanes), parameter gradients
// are accumulated in registers by a persistent grid and reduced in a fixed order (no atomics).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace alignn {

constexpr int EPI_THREADS = 256;
constexpr int EPI_WARPS = EPI_THREADS / 32;
constexpr int EPI_PARTIAL_BLOCKS = 296;  // 2 x 148 SMs: persistent grid of the backward
constexpr int EPI_GEN_WARPS = 4;

// dropout keep-scales of 8 consecutive elements starting at a multiple of 8: ONE Philox4x32-10 block per 8 elements, 16
// random bits each (drop iff u16 < p * 65536: the drop probability is p rounded down to a multiple of 2^-16).  Forward
// and backward of the fast kernels call this with the same (seed, offset, first), so they see the same mask.
__device__ __forceinline__ void dropout8(uint64_t seed, uint64_t offset, uint64_t first, float p,
                                         float inv_keep, float *out) {
    const Philox4 r = philox4x32_10(seed, offset, first >> 3);
    const uint32_t thr = (uint32_t)(p * 65536.0f);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t u = (w[c >> 1] >> (16 * (c & 1))) & 0xffffu;
        out[c] = u < thr ? 0.0f : inv_keep;
    }
}

// optional operands of the streaming (edge-GEMM-free) path; all-null / ld = hidden reproduces the plain op
struct GateExtra {
    const void *agge;      // [heads, rows, C] storage dtype: per-head Wc[t] abar_t (may be null)
    const float *cvec;     // [hidden] f32: c = W_e b (may be null)
    const float *stat_s;   // [rows, heads] f32: S = sum_j a~
    float *agg_out;        // [rows, hidden] f32: full aggregate, saved for backward (may be null)
    void *dagg_lp;         // backward: storage-dtype copy of dagg (may be null)
    const void *dy2;       // backward: second upstream-gradient addend in storage dtype (may be null)
    int64_t lddy2;
    int64_t agg_rows;      // rows >= agg_rows have no aggregate (agg = 0, nothing read or written for it); < 0: all rows
    bool dcvec;            // backward: also reduce dc = sum_rows dagg * S (needs stat_s, heads) -> 6 parameter vectors
    int heads;
    int64_t ldxr, lddxr;   // row strides (elements) of xr and dxr
    const uint64_t *rng_step;   // optional device counter added to the dropout offset (CUDA-graph replays)
    const void *x_lp;           // forward: residual input in the STORAGE dtype, read when the fp32 `x` is null
};

// ---- per-lane cp.async ring --------------------------------------------------------------------------
// The fast forward kernel is persistent (2 CTAs per SM) and streams its operand rows through shared memory: every lane copies
// the 16-byte pieces of ITS OWN 8 channels of a row EPI_STAGES - 1 iterations ahead with cp.async and reads back only
// what it copied itself, so the ring needs no barrier -- cp.async.wait_group is the only synchronisation -- and the
// bytes in flight per SM (2 CTAs x 8 warps x 3 rows ahead) no longer depend on the register budget.
template <typename T> struct EpiRing {
    static constexpr int CT = (int)sizeof(T) * 8 / 16;          // 16-byte chunks per lane for 8 elements of T
    static constexpr int STAGES = sizeof(T) == 2 ? 4 : 3;
    static constexpr int NCH = 4 + 2 * CT;                      // f32 row (2) | T row (CT) | f32-or-T row (2) | T row (CT)
    static constexpr int BYTES = EPI_WARPS * STAGES * NCH * 512;
};

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// this lane's 8 elements starting at g -> chunk slots `chunk`, `chunk + 1` (512 B apart: conflict-free across the warp)
__device__ __forceinline__ void cp_row8(uint32_t lane_base, int chunk, const float *g) {
    cp_async16(lane_base + chunk * 512, g);
    cp_async16(lane_base + (chunk + 1) * 512, g + 4);
}
__device__ __forceinline__ void cp_row8(uint32_t lane_base, int chunk, const __nv_bfloat16 *g) {
    cp_async16(lane_base + chunk * 512, g);
}
__device__ __forceinline__ F8 lds8(const uint4 *lane_ptr, int chunk, float) {
    const uint4 a = lane_ptr[chunk * 32], b = lane_ptr[(chunk + 1) * 32];
    F8 r;
    r.v[0] = __uint_as_float(a.x); r.v[1] = __uint_as_float(a.y); r.v[2] = __uint_as_float(a.z); r.v[3] = __uint_as_float(a.w);
    r.v[4] = __uint_as_float(b.x); r.v[5] = __uint_as_float(b.y); r.v[6] = __uint_as_float(b.z); r.v[7] = __uint_as_float(b.w);
    return r;
}
__device__ __forceinline__ F8 lds8(const uint4 *lane_ptr, int chunk, __nv_bfloat16) {
    const uint4 u = lane_ptr[chunk * 32];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    F8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __uint_as_float(w[i] << 16);
        r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
}

template <typename T, int LANES>
__global__ void __launch_bounds__(EPI_THREADS, 2)
gate_ln_fwd_kernel(const float *__restrict__ agg, const T *__restrict__ xr, const float *__restrict__ x,
                   const float *__restrict__ wbeta, const float *__restrict__ gamma,
                   const float *__restrict__ bias, float *__restrict__ y, T *__restrict__ y_lp,
                   float *__restrict__ beta_out, float *__restrict__ mean_out, float *__restrict__ rstd_out,
                   int64_t n_rows, int hidden, float eps, float p_drop, float inv_keep, uint64_t seed,
                   uint64_t offset, const GateExtra X) {
    using R = EpiRing<T>;
    constexpr int RPW = 32 / LANES, CT = R::CT, S = R::STAGES;
    constexpr int C_AGG = 0, C_XR = 2, C_X = 2 + CT, C_AGGE = 4 + CT;
    extern __shared__ uint4 epi_ring[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const int ch = sub * 8;
    const int64_t slots = (int64_t)gridDim.x * EPI_WARPS * RPW;
    const int64_t slot = ((int64_t)blockIdx.x * EPI_WARPS + warp) * RPW + lane / LANES;
    const int64_t iters = (n_rows + slots - 1) / slots;
    const int64_t agg_rows = X.agg_rows < 0 ? n_rows : X.agg_rows;
    const int C = hidden / X.heads, t = ch / C;
    const T *x_lp = reinterpret_cast<const T *>(X.x_lp);
    const T *agge = reinterpret_cast<const T *>(X.agge);
    const uint4 *my_ring = epi_ring + (size_t)warp * S * R::NCH * 32 + lane;
    const uint32_t my_ring_s = (uint32_t)__cvta_generic_to_shared(my_ring);

    auto issue = [&](int64_t it) {
        const int64_t row = slot + it * slots;
        if (it < iters && row < n_rows) {
            const uint32_t base = my_ring_s + (uint32_t)(it % S) * (R::NCH * 512);
            if (row < agg_rows) {
                cp_row8(base, C_AGG, agg + row * hidden + ch);
                if (agge) cp_row8(base, C_AGGE, agge + ((int64_t)t * agg_rows + row) * C + (ch - t * C));
            }
            cp_row8(base, C_XR, xr + row * X.ldxr + ch);
            if (x) cp_row8(base, C_X, x + row * hidden + ch);
            else cp_row8(base, C_X, x_lp + row * hidden + ch);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < S - 1; ++i) issue(i);

    // gate logit  w1.a + w2.s + w3.(a - s)  =  (w1 + w3).a + (w2 - w3).s : two independent 8-term chains per lane
    F8 w13, w23;
    {
        const F8 w1 = ld8(wbeta + ch), w2 = ld8(wbeta + hidden + ch), w3 = ld8(wbeta + 2 * hidden + ch);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            w13.v[c] = w1.v[c] + w3.v[c];
            w23.v[c] = w2.v[c] - w3.v[c];
        }
    }
    const F8 gm = ld8(gamma + ch), bs = ld8(bias + ch);
    F8 cf;
#pragma unroll
    for (int c = 0; c < 8; ++c) cf.v[c] = 0.f;
    if (agge && X.cvec) cf = ld8(X.cvec + ch);
    const uint64_t rng_off = offset + (X.rng_step ? *X.rng_step : 0ull);
    const float inv_h = 1.0f / (float)hidden;

    for (int64_t it = 0; it < iters; ++it) {
        issue(it + S - 1);
        cp_async_wait<S - 1>();
        const int64_t row = slot + it * slots;
        const bool ok = row < n_rows;
        const bool has_agg = ok && row < agg_rows;
        const uint4 *st = my_ring + (size_t)(it % S) * (R::NCH * 32);
        F8 af, sf, xf;
#pragma unroll
        for (int c = 0; c < 8; ++c) af.v[c] = sf.v[c] = xf.v[c] = 0.f;
        if (ok) {
            sf = lds8(st, C_XR, T());
            xf = x ? lds8(st, C_X, float()) : lds8(st, C_X, T());
        }
        if (has_agg) {
            af = lds8(st, C_AGG, float());
            if (agge) {                 // agg = aggv + (Wc[t] abar_t)  [HEADS, agg_rows, C]  + c_t * S_t
                const F8 ef = lds8(st, C_AGGE, T());
                const float sv = X.cvec ? __ldg(X.stat_s + row * X.heads + t) : 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) af.v[c] += ef.v[c] + cf.v[c] * sv;
            }
            if (X.agg_out) st8(X.agg_out + row * hidden + ch, af);
        }
        float za = 0.f, zs = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            za = fmaf(w13.v[c], af.v[c], za);
            zs = fmaf(w23.v[c], sf.v[c], zs);
        }
        const float zl = group_sum<LANES>(za + zs);
        const float beta = 1.0f / (1.0f + expf(-zl));
        F8 o;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            o.v[c] = beta * sf.v[c] + (1.0f - beta) * af.v[c];
            sum += o.v[c];
        }
        const float mean = group_sum<LANES>(sum) * inv_h;
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float d = o.v[c] - mean;
            sq = fmaf(d, d, sq);
        }
        const float var = group_sum<LANES>(sq) * inv_h;
        const float rstd = 1.0f / sqrtf(var + eps);
        if (!ok) continue;
        float keep[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) keep[c] = 1.f;
        if (p_drop > 0.f) dropout8(seed, rng_off, (uint64_t)row * hidden + ch, p_drop, inv_keep, keep);
        F8 out;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float yv = (o.v[c] - mean) * rstd * gm.v[c] + bs.v[c];
            out.v[c] = xf.v[c] + fmaxf(yv, 0.f) * keep[c];
        }
        st8(y + row * hidden + ch, out);
        if (y_lp) st8(y_lp + row * hidden + ch, out);
        if (sub == 0) {
            beta_out[row] = beta;
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
    }
}

template <typename T, int LANES>
__global__ void __launch_bounds__(EPI_THREADS, 2)
gate_ln_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ agg, const T *__restrict__ xr,
                   const float *__restrict__ wbeta, const float *__restrict__ gamma,
                   const float *__restrict__ bias, const float *__restrict__ beta_in,
                   const float *__restrict__ mean_in, const float *__restrict__ rstd_in,
                   float *__restrict__ dagg, T *__restrict__ dxr, float *__restrict__ partials,
                   int64_t n_rows, int hidden, float p_drop, float inv_keep, uint64_t seed, uint64_t offset,
                   const GateExtra X) {
    constexpr int RPW = 32 / LANES;
    __shared__ float red[EPI_WARPS][LANES * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const int ch = sub * 8;
    const int nvec = X.dcvec ? 6 : 5;                    // 6th vector: dc[ch] = sum_rows dagg[row, ch] * S[row, head(ch)]
    const int head = X.dcvec ? ch / (hidden / X.heads) : 0;
    const int64_t slots = (int64_t)gridDim.x * EPI_WARPS * RPW;
    const int64_t slot = ((int64_t)blockIdx.x * EPI_WARPS + warp) * RPW + lane / LANES;
    const int64_t iters = (n_rows + slots - 1) / slots;

    const F8 w1 = ld8(wbeta + ch), w2 = ld8(wbeta + hidden + ch), w3 = ld8(wbeta + 2 * hidden + ch);
    const F8 gm = ld8(gamma + ch), bs = ld8(bias + ch);
    F8 a1, a2, a3, ag, ab, ac;
#pragma unroll
    for (int c = 0; c < 8; ++c) a1.v[c] = a2.v[c] = a3.v[c] = ag.v[c] = ab.v[c] = ac.v[c] = 0.f;
    const float inv_h = 1.0f / (float)hidden;

    for (int64_t it = 0; it < iters; ++it) {
        const int64_t row = slot + it * slots;
        const bool ok = row < n_rows;
        F8 gf, af, sf;
        float beta = 0.f, mean = 0.f, rstd = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) gf.v[c] = af.v[c] = sf.v[c] = 0.f;
        if (ok) {
            if (dy) gf = ld8(dy + row * hidden + ch);
            if (X.dy2) {
                const F8 g2 = ld8(reinterpret_cast<const T *>(X.dy2) + row * X.lddy2 + ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) gf.v[c] += g2.v[c];
            }
            if (X.agg_rows < 0 || row < X.agg_rows) af = ld8(agg + row * hidden + ch);
            sf = ld8(xr + row * X.ldxr + ch);
            beta = __ldg(beta_in + row);
            mean = __ldg(mean_in + row);
            rstd = __ldg(rstd_in + row);
        }
        float keep[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) keep[c] = 1.f;
        if (p_drop > 0.f && ok)
            dropout8(seed, offset + (X.rng_step ? *X.rng_step : 0ull), (uint64_t)row * hidden + ch, p_drop, inv_keep, keep);
        F8 xh, dxh;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float o = beta * sf.v[c] + (1.0f - beta) * af.v[c];
            xh.v[c] = (o - mean) * rstd;
            const float yv = xh.v[c] * gm.v[c] + bs.v[c];
            const float dyv = yv > 0.f ? gf.v[c] * keep[c] : 0.f;
            ag.v[c] = fmaf(dyv, xh.v[c], ag.v[c]);
            ab.v[c] += dyv;
            dxh.v[c] = dyv * gm.v[c];
            s1 += dxh.v[c];
            s2 = fmaf(dxh.v[c], xh.v[c], s2);
        }
        const float m1 = group_sum<LANES>(s1) * inv_h;
        const float m2 = group_sum<LANES>(s2) * inv_h;
        F8 dof;
        float bp = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            dof.v[c] = rstd * (dxh.v[c] - m1 - xh.v[c] * m2);
            bp = fmaf(dof.v[c], sf.v[c] - af.v[c], bp);
        }
        const float dbeta = group_sum<LANES>(bp);
        const float dz = dbeta * beta * (1.0f - beta);
        if (ok) {
            F8 da, ds;
            const bool has_agg = X.agg_rows < 0 || row < X.agg_rows;
            const float srow = X.dcvec && has_agg ? __ldg(X.stat_s + row * X.heads + head) : 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                da.v[c] = (1.0f - beta) * dof.v[c] + dz * (w1.v[c] + w3.v[c]);
                ds.v[c] = beta * dof.v[c] + dz * (w2.v[c] - w3.v[c]);
                a1.v[c] = fmaf(dz, af.v[c], a1.v[c]);
                a2.v[c] = fmaf(dz, sf.v[c], a2.v[c]);
                a3.v[c] = fmaf(dz, af.v[c] - sf.v[c], a3.v[c]);
                ac.v[c] = fmaf(da.v[c], srow, ac.v[c]);
            }
            if (has_agg) {
                st8(dagg + row * hidden + ch, da);
                if (X.dagg_lp) st8(reinterpret_cast<T *>(X.dagg_lp) + row * hidden + ch, da);
            }
            st8(dxr + row * X.lddxr + ch, ds);
        }
    }
    // fold the row groups of a warp (fixed butterfly order), then the warps of the block (fixed order)
    if (RPW > 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int off = 16; off >= LANES; off >>= 1) {
                a1.v[c] += __shfl_xor_sync(FULL, a1.v[c], off);
                a2.v[c] += __shfl_xor_sync(FULL, a2.v[c], off);
                a3.v[c] += __shfl_xor_sync(FULL, a3.v[c], off);
                ag.v[c] += __shfl_xor_sync(FULL, ag.v[c], off);
                ab.v[c] += __shfl_xor_sync(FULL, ab.v[c], off);
                ac.v[c] += __shfl_xor_sync(FULL, ac.v[c], off);
            }
        }
    }
#pragma unroll
    for (int which = 0; which < 6; ++which) {          // one vector at a time through an 8 KB staging buffer
        if (which >= nvec) break;                      // block-uniform
        const F8 &src = which == 0 ? a1 : which == 1 ? a2 : which == 2 ? a3 : which == 3 ? ag : which == 4 ? ab : ac;
        if (lane < LANES) {
#pragma unroll
            for (int c = 0; c < 8; ++c) red[warp][ch + c] = src.v[c];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < hidden; c += EPI_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < EPI_WARPS; ++w) s += red[w][c];
            partials[((int64_t)blockIdx.x * nvec + which) * hidden + c] = s;
        }
        __syncthreads();
    }
}

// ---- backward, hidden = 256 in bf16 (the training regime): one warp per row, two specialised loops ---------------------
// The generic kernel above spends ~540 instructions per row, of which ~230 are arithmetic: per-row predicates (row in range?
// row has an aggregate?), seven 64-bit row addresses, constants reloaded from local memory (the register budget of 2 CTAs /
// SM does not hold five 8-wide constant vectors next to six accumulator vectors), three beta-gradient accumulators where two
// suffice.  Here a warp walks ITS rows in order, first the rows that have an aggregate (row < agg_rows), then the isolated
// ones (agg = 0: out = beta x_r, no dagg, no d/dc, and dbeta = <dout, x_r>): no predicate inside either loop, only the
// constants each loop needs, dw_beta[2H:3H] = dw_beta[0:H] - dw_beta[H:2H] formed once at the end.  Same mask, same
// statistics, same partial layout and the same fixed-order reduction as the generic kernel.
template <bool HAS_AGG, bool DROP>
__device__ __forceinline__ void glb256_row(int64_t row, int ch, const float *__restrict__ dy, const __nv_bfloat16 *__restrict__ dy2,
                                           int64_t lddy2, const float *__restrict__ agg, const __nv_bfloat16 *__restrict__ xr,
                                           int64_t ldxr, const float *__restrict__ beta_in, const float *__restrict__ mean_in,
                                           const float *__restrict__ rstd_in, const float *__restrict__ stat_s, int heads, int head,
                                           bool dcvec, float *__restrict__ dagg, __nv_bfloat16 *__restrict__ dagg_lp,
                                           __nv_bfloat16 *__restrict__ dxr, int64_t lddxr, const F8 &gm, const F8 &bs, const F8 &w13,
                                           const F8 &w23, float p_drop, float inv_keep, uint64_t seed, uint64_t rng_off,
                                           F8 &a1, F8 &a2, F8 &ag, F8 &ab, F8 &ac) {
    F8 gf;
    if (dy) gf = ld8(dy + row * 256 + ch);
    else {
#pragma unroll
        for (int c = 0; c < 8; ++c) gf.v[c] = 0.f;
    }
    if (dy2) {
        const F8 g2 = ld8(dy2 + row * lddy2 + ch);
#pragma unroll
        for (int c = 0; c < 8; ++c) gf.v[c] += g2.v[c];
    }
    const F8 sf = ld8(xr + row * ldxr + ch);
    F8 af;
    if (HAS_AGG) af = ld8(agg + row * 256 + ch);
    const float beta = __ldg(beta_in + row), mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    if (DROP) {
        float keep[8];
        dropout8(seed, rng_off, (uint64_t)row * 256 + ch, p_drop, inv_keep, keep);
#pragma unroll
        for (int c = 0; c < 8; ++c) gf.v[c] *= keep[c];
    }
    F8 xh, dxh;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        // isolated rows: the forward's  beta * x_r + (1 - beta) * 0  rounds the product ONCE before the mean is subtracted;
        // __fmul_rn keeps the compiler from contracting it into (beta * x_r - mean) here, which would move pre-activations
        // that sit at zero across the ReLU threshold relative to the forward
        const float o = HAS_AGG ? beta * sf.v[c] + (1.0f - beta) * af.v[c] : __fmul_rn(beta, sf.v[c]);
        xh.v[c] = (o - mean) * rstd;
        const float yv = xh.v[c] * gm.v[c] + bs.v[c];
        const float dyv = yv > 0.f ? gf.v[c] : 0.f;
        ag.v[c] = fmaf(dyv, xh.v[c], ag.v[c]);
        ab.v[c] += dyv;
        dxh.v[c] = dyv * gm.v[c];
        s1 += dxh.v[c];
        s2 = fmaf(dxh.v[c], xh.v[c], s2);
    }
    // both row sums through one butterfly
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_xor_sync(FULL, s1, off);
        s2 += __shfl_xor_sync(FULL, s2, off);
    }
    const float m1 = s1 * (1.0f / 256.0f), m2 = s2 * (1.0f / 256.0f);
    F8 dof;
    float bp = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        dof.v[c] = rstd * (dxh.v[c] - m1 - xh.v[c] * m2);
        bp = fmaf(dof.v[c], HAS_AGG ? sf.v[c] - af.v[c] : sf.v[c], bp);
    }
    const float dbeta = group_sum<32>(bp);
    const float dz = dbeta * beta * (1.0f - beta);
    F8 ds;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        ds.v[c] = beta * dof.v[c] + dz * w23.v[c];
        a2.v[c] = fmaf(dz, sf.v[c], a2.v[c]);
    }
    st8(dxr + row * lddxr + ch, ds);
    if (HAS_AGG) {
        const float srow = dcvec ? __ldg(stat_s + row * heads + head) : 0.f;
        F8 da;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            da.v[c] = (1.0f - beta) * dof.v[c] + dz * w13.v[c];
            a1.v[c] = fmaf(dz, af.v[c], a1.v[c]);
            ac.v[c] = fmaf(da.v[c], srow, ac.v[c]);
        }
        st8(dagg + row * 256 + ch, da);
        if (dagg_lp) st8(dagg_lp + row * 256 + ch, da);
    }
}

template <bool DROP>
__global__ void __launch_bounds__(EPI_THREADS, 2)
gate_ln_bwd256_kernel(const float *__restrict__ dy, const float *__restrict__ agg, const __nv_bfloat16 *__restrict__ xr,
                      const float *__restrict__ wbeta, const float *__restrict__ gamma, const float *__restrict__ bias,
                      const float *__restrict__ beta_in, const float *__restrict__ mean_in, const float *__restrict__ rstd_in,
                      float *__restrict__ dagg, __nv_bfloat16 *__restrict__ dxr, float *__restrict__ partials, int64_t n_rows,
                      float p_drop, float inv_keep, uint64_t seed, uint64_t offset, const GateExtra X) {
    constexpr int HID = 256;
    __shared__ float red[EPI_WARPS][HID];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = lane * 8;
    const int nvec = X.dcvec ? 6 : 5;
    const int head = X.dcvec ? ch / (HID / X.heads) : 0;
    const int64_t slots = (int64_t)gridDim.x * EPI_WARPS;
    const int64_t agg_rows = X.agg_rows < 0 ? n_rows : X.agg_rows;
    const uint64_t rng_off = offset + (X.rng_step ? *X.rng_step : 0ull);
    const __nv_bfloat16 *dy2 = reinterpret_cast<const __nv_bfloat16 *>(X.dy2);
    __nv_bfloat16 *dagg_lp = reinterpret_cast<__nv_bfloat16 *>(X.dagg_lp);

    F8 w13, w23;
    {
        const F8 w1 = ld8(wbeta + ch), w2 = ld8(wbeta + HID + ch), w3 = ld8(wbeta + 2 * HID + ch);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            w13.v[c] = w1.v[c] + w3.v[c];
            w23.v[c] = w2.v[c] - w3.v[c];
        }
    }
    const F8 gm = ld8(gamma + ch), bs = ld8(bias + ch);
    F8 a1, a2, ag, ab, ac;
#pragma unroll
    for (int c = 0; c < 8; ++c) a1.v[c] = a2.v[c] = ag.v[c] = ab.v[c] = ac.v[c] = 0.f;

    int64_t row = (int64_t)blockIdx.x * EPI_WARPS + warp;
    for (; row < agg_rows; row += slots)
        glb256_row<true, DROP>(row, ch, dy, dy2, X.lddy2, agg, xr, X.ldxr, beta_in, mean_in, rstd_in, X.stat_s, X.heads, head,
                               X.dcvec, dagg, dagg_lp, dxr, X.lddxr, gm, bs, w13, w23, p_drop, inv_keep, seed, rng_off, a1, a2, ag, ab, ac);
    for (; row < n_rows; row += slots)
        glb256_row<false, DROP>(row, ch, dy, dy2, X.lddy2, agg, xr, X.ldxr, beta_in, mean_in, rstd_in, X.stat_s, X.heads, head,
                                X.dcvec, dagg, dagg_lp, dxr, X.lddxr, gm, bs, w13, w23, p_drop, inv_keep, seed, rng_off, a1, a2, ag, ab, ac);

    // fold the warps of the block in a fixed order, one vector at a time: dw_beta x3 | dgamma | dbias | dc
#pragma unroll
    for (int which = 0; which < 6; ++which) {
        if (which >= nvec) break;                      // block-uniform
        F8 src;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            src.v[c] = which == 0 ? a1.v[c] : which == 1 ? a2.v[c] : which == 2 ? a1.v[c] - a2.v[c] : which == 3 ? ag.v[c]
                       : which == 4 ? ab.v[c] : ac.v[c];
        st8(&red[warp][ch], src);
        __syncthreads();
        for (int c = threadIdx.x; c < HID; c += EPI_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < EPI_WARPS; ++w) s += red[w][c];
            partials[((int64_t)blockIdx.x * nvec + which) * HID + c] = s;
        }
        __syncthreads();
    }
}

// out[i] = sum_b partials[b * width + i]: 8 threads per column stride the blocks, then a fixed-order fold (deterministic)
constexpr int RP_COLS = 32, RP_ROWS = 8;
__global__ void __launch_bounds__(RP_COLS * RP_ROWS)
reduce_partials_kernel(const float *__restrict__ partials, float *__restrict__ out, int n_blocks, int width) {
    __shared__ float red[RP_ROWS][RP_COLS];
    const int tx = threadIdx.x % RP_COLS, ty = threadIdx.x / RP_COLS;
    const int i = blockIdx.x * RP_COLS + tx;
    float s = 0.f;
    if (i < width)
        for (int b = ty; b < n_blocks; b += RP_ROWS) s += partials[(int64_t)b * width + i];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < width) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < RP_ROWS; ++r) t += red[r][tx];
        out[i] = t;
    }
}

// ---- generic (any hidden): one warp per row, lanes stride channels -----------------------------------
template <typename T>
__global__ void __launch_bounds__(EPI_GEN_WARPS * 32)
gate_ln_fwd_generic_kernel(const float *__restrict__ agg, const T *__restrict__ xr, const float *__restrict__ x,
                           const float *__restrict__ wbeta, const float *__restrict__ gamma,
                           const float *__restrict__ bias, float *__restrict__ y, T *__restrict__ y_lp,
                           float *__restrict__ beta_out, float *__restrict__ mean_out,
                           float *__restrict__ rstd_out, int64_t n_rows, int hidden, float eps, float p_drop,
                           float inv_keep, uint64_t seed, uint64_t offset) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * EPI_GEN_WARPS + warp;
    if (row >= n_rows) return;
    const float *ar = agg + row * hidden;
    const T *sr = xr + row * hidden;
    float zp = 0.f;
    for (int c = lane; c < hidden; c += 32) {
        const float a = ar[c], s = ldf(sr + c);
        zp += wbeta[c] * a + wbeta[hidden + c] * s + wbeta[2 * hidden + c] * (a - s);
    }
    const float beta = 1.0f / (1.0f + expf(-warp_sum(zp)));
    float sum = 0.f;
    for (int c = lane; c < hidden; c += 32) sum += beta * ldf(sr + c) + (1.0f - beta) * ar[c];
    const float mean = warp_sum(sum) / (float)hidden;
    float sq = 0.f;
    for (int c = lane; c < hidden; c += 32) {
        const float d = beta * ldf(sr + c) + (1.0f - beta) * ar[c] - mean;
        sq = fmaf(d, d, sq);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)hidden + eps);
    for (int c = lane; c < hidden; c += 32) {
        const float o = beta * ldf(sr + c) + (1.0f - beta) * ar[c];
        const float yv = (o - mean) * rstd * gamma[c] + bias[c];
        float keep = 1.f;
        if (p_drop > 0.f) keep = dropout_scale(seed, offset, (uint64_t)row * hidden + c, p_drop, inv_keep);
        const float out = x[row * hidden + c] + fmaxf(yv, 0.f) * keep;
        y[row * hidden + c] = out;
        if (y_lp) stf(y_lp + row * hidden + c, out);
    }
    if (lane == 0) {
        beta_out[row] = beta;
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
}

template <typename T>
__global__ void __launch_bounds__(EPI_GEN_WARPS * 32)
gate_ln_bwd_generic_kernel(const float *__restrict__ dy, const float *__restrict__ agg, const T *__restrict__ xr,
                           const float *__restrict__ wbeta, const float *__restrict__ gamma,
                           const float *__restrict__ bias, const float *__restrict__ beta_in,
                           const float *__restrict__ mean_in, const float *__restrict__ rstd_in,
                           float *__restrict__ dagg, T *__restrict__ dxr, float *__restrict__ partials,
                           int64_t n_rows, int hidden, float p_drop, float inv_keep, uint64_t seed,
                           uint64_t offset) {
    extern __shared__ float acc[];  // [EPI_GEN_WARPS][5][hidden]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *mine = acc + (size_t)warp * 5 * hidden;
    for (int i = lane; i < 5 * hidden; i += 32) mine[i] = 0.f;
    __syncwarp();
    const int64_t stride = (int64_t)gridDim.x * EPI_GEN_WARPS;
    const float inv_h = 1.0f / (float)hidden;
    for (int64_t row = (int64_t)blockIdx.x * EPI_GEN_WARPS + warp; row < n_rows; row += stride) {
        const float *ar = agg + row * hidden, *gr = dy + row * hidden;
        const T *sr = xr + row * hidden;
        const float beta = beta_in[row], mean = mean_in[row], rstd = rstd_in[row];
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < hidden; c += 32) {
            const float o = beta * ldf(sr + c) + (1.0f - beta) * ar[c];
            const float xh = (o - mean) * rstd;
            const float yv = xh * gamma[c] + bias[c];
            float keep = 1.f;
            if (p_drop > 0.f) keep = dropout_scale(seed, offset, (uint64_t)row * hidden + c, p_drop, inv_keep);
            const float dyv = yv > 0.f ? gr[c] * keep : 0.f;
            mine[3 * hidden + c] = fmaf(dyv, xh, mine[3 * hidden + c]);
            mine[4 * hidden + c] += dyv;
            const float dxh = dyv * gamma[c];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
        }
        const float m1 = warp_sum(s1) * inv_h, m2 = warp_sum(s2) * inv_h;
        float bp = 0.f;
        for (int c = lane; c < hidden; c += 32) {
            const float s = ldf(sr + c), a = ar[c];
            const float o = beta * s + (1.0f - beta) * a;
            const float xh = (o - mean) * rstd;
            const float yv = xh * gamma[c] + bias[c];
            float keep = 1.f;
            if (p_drop > 0.f) keep = dropout_scale(seed, offset, (uint64_t)row * hidden + c, p_drop, inv_keep);
            const float dxh = (yv > 0.f ? gr[c] * keep : 0.f) * gamma[c];
            const float dof = rstd * (dxh - m1 - xh * m2);
            bp = fmaf(dof, s - a, bp);
        }
        const float dbeta = warp_sum(bp);
        const float dz = dbeta * beta * (1.0f - beta);
        for (int c = lane; c < hidden; c += 32) {
            const float s = ldf(sr + c), a = ar[c];
            const float o = beta * s + (1.0f - beta) * a;
            const float xh = (o - mean) * rstd;
            const float yv = xh * gamma[c] + bias[c];
            float keep = 1.f;
            if (p_drop > 0.f) keep = dropout_scale(seed, offset, (uint64_t)row * hidden + c, p_drop, inv_keep);
            const float dxh = (yv > 0.f ? gr[c] * keep : 0.f) * gamma[c];
            const float dof = rstd * (dxh - m1 - xh * m2);
            dagg[row * hidden + c] = (1.0f - beta) * dof + dz * (wbeta[c] + wbeta[2 * hidden + c]);
            stf(dxr + row * hidden + c, beta * dof + dz * (wbeta[hidden + c] - wbeta[2 * hidden + c]));
            mine[c] = fmaf(dz, a, mine[c]);
            mine[hidden + c] = fmaf(dz, s, mine[hidden + c]);
            mine[2 * hidden + c] = fmaf(dz, a - s, mine[2 * hidden + c]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 5 * hidden; i += EPI_GEN_WARPS * 32) {
        float s = 0.f;
        for (int w = 0; w < EPI_GEN_WARPS; ++w) s += acc[(size_t)w * 5 * hidden + i];
        partials[(int64_t)blockIdx.x * 5 * hidden + i] = s;
    }
}

static inline bool epi_fast(int hidden, int *lanes) {
    if (hidden % 8) return false;
    const int L = hidden / 8;
    if (L > 32 || (L & (L - 1))) return false;
    *lanes = L;
    return true;
}

template <typename T, int LANES>
static void launch_epi_fwd(const float *agg, const void *xr, const float *x, const float *wbeta, const float *gamma,
                           const float *bias, float *y, void *y_lp, float *beta, float *mean, float *rstd,
                           int64_t n_rows, int hidden, float eps, float p_drop, uint64_t seed, uint64_t offset,
                           const GateExtra &X, cudaStream_t st) {
    const int rows_per_block = EPI_WARPS * (32 / LANES);
    const int64_t want = (n_rows + rows_per_block - 1) / rows_per_block;
    const unsigned grid = (unsigned)(want < EPI_PARTIAL_BLOCKS ? want : EPI_PARTIAL_BLOCKS);      // persistent: 2 CTAs per SM
    const float inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    cudaFuncSetAttribute(gate_ln_fwd_kernel<T, LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, EpiRing<T>::BYTES);
    gate_ln_fwd_kernel<T, LANES><<<grid, EPI_THREADS, EpiRing<T>::BYTES, st>>>(agg, (const T *)xr, x, wbeta, gamma, bias, y, (T *)y_lp,
                                                               beta, mean, rstd, n_rows, hidden, eps, p_drop, inv_keep,
                                                               seed, offset, X);
}

template <typename T>
static int dispatch_epi_fwd(const float *agg, const void *xr, const float *x, const float *wbeta, const float *gamma,
                            const float *bias, float *y, void *y_lp, float *beta, float *mean, float *rstd,
                            int64_t n_rows, int hidden, float eps, float p_drop, uint64_t seed, uint64_t offset,
                            const GateExtra &X, bool extras, cudaStream_t st) {
    int lanes = 0;
#define EPI_FWD(L) launch_epi_fwd<T, L>(agg, xr, x, wbeta, gamma, bias, y, y_lp, beta, mean, rstd, n_rows, hidden, eps, p_drop, seed, offset, X, st)
    if (extras && !epi_fast(hidden, &lanes)) return ALIGNN_ERR_BAD_SHAPE;
    if (epi_fast(hidden, &lanes)) {
        switch (lanes) {
            case 32: EPI_FWD(32); break;
            case 16: EPI_FWD(16); break;
            case 8: EPI_FWD(8); break;
            case 4: EPI_FWD(4); break;
            case 2: EPI_FWD(2); break;
            default: EPI_FWD(1); break;
        }
    } else {
        const float inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
        const unsigned grid = (unsigned)((n_rows + EPI_GEN_WARPS - 1) / EPI_GEN_WARPS);
        gate_ln_fwd_generic_kernel<T><<<grid, EPI_GEN_WARPS * 32, 0, st>>>(
            agg, (const T *)xr, x, wbeta, gamma, bias, y, (T *)y_lp, beta, mean, rstd, n_rows, hidden, eps, p_drop,
            inv_keep, seed, offset);
    }
#undef EPI_FWD
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

template <typename T, int LANES>
static void launch_epi_bwd(const float *dy, const float *agg, const void *xr, const float *wbeta, const float *gamma,
                           const float *bias, const float *beta, const float *mean, const float *rstd, float *dagg,
                           void *dxr, float *partials, int64_t n_rows, int hidden, float p_drop, uint64_t seed,
                           uint64_t offset, const GateExtra &X, cudaStream_t st) {
    const float inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    gate_ln_bwd_kernel<T, LANES><<<EPI_PARTIAL_BLOCKS, EPI_THREADS, 0, st>>>(
        dy, agg, (const T *)xr, wbeta, gamma, bias, beta, mean, rstd, dagg, (T *)dxr, partials, n_rows, hidden, p_drop,
        inv_keep, seed, offset, X);
}

template <typename T>
static int dispatch_epi_bwd(const float *dy, const float *agg, const void *xr, const float *wbeta, const float *gamma,
                            const float *bias, const float *beta, const float *mean, const float *rstd, float *dagg,
                            void *dxr, float *partials, float *dparams, int64_t n_rows, int hidden, float p_drop,
                            uint64_t seed, uint64_t offset, const GateExtra &X, bool extras, cudaStream_t st) {
    int lanes = 0;
#define EPI_BWD(L) launch_epi_bwd<T, L>(dy, agg, xr, wbeta, gamma, bias, beta, mean, rstd, dagg, dxr, partials, n_rows, hidden, p_drop, seed, offset, X, st)
    if (extras && !epi_fast(hidden, &lanes)) return ALIGNN_ERR_BAD_SHAPE;
    static const bool glb256 = [] { const char *e = getenv("ALIGNN_GLB256"); return !e || atoi(e) != 0; }();
    if (glb256 && sizeof(T) == 2 && hidden == 256 && (!X.dcvec || (X.heads > 0 && 256 % X.heads == 0 && (256 / X.heads) % 8 == 0))) {
        // the training regime: bf16, hidden 256 -> the two-loop kernel
        const float inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
        if (p_drop > 0.f)
            gate_ln_bwd256_kernel<true><<<EPI_PARTIAL_BLOCKS, EPI_THREADS, 0, st>>>(
                dy, agg, (const __nv_bfloat16 *)xr, wbeta, gamma, bias, beta, mean, rstd, dagg, (__nv_bfloat16 *)dxr, partials,
                n_rows, p_drop, inv_keep, seed, offset, X);
        else
            gate_ln_bwd256_kernel<false><<<EPI_PARTIAL_BLOCKS, EPI_THREADS, 0, st>>>(
                dy, agg, (const __nv_bfloat16 *)xr, wbeta, gamma, bias, beta, mean, rstd, dagg, (__nv_bfloat16 *)dxr, partials,
                n_rows, p_drop, inv_keep, seed, offset, X);
    } else if (epi_fast(hidden, &lanes)) {
        switch (lanes) {
            case 32: EPI_BWD(32); break;
            case 16: EPI_BWD(16); break;
            case 8: EPI_BWD(8); break;
            case 4: EPI_BWD(4); break;
            case 2: EPI_BWD(2); break;
            default: EPI_BWD(1); break;
        }
    } else {
        const float inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
        const size_t smem = (size_t)EPI_GEN_WARPS * 5 * hidden * sizeof(float);
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(gate_ln_bwd_generic_kernel<T>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return ALIGNN_ERR_CUDA_BASE + (int)err;
        }
        gate_ln_bwd_generic_kernel<T><<<EPI_PARTIAL_BLOCKS, EPI_GEN_WARPS * 32, smem, st>>>(
            dy, agg, (const T *)xr, wbeta, gamma, bias, beta, mean, rstd, dagg, (T *)dxr, partials, n_rows, hidden,
            p_drop, inv_keep, seed, offset);
    }
#undef EPI_BWD
    ALIGNN_LAUNCH_CHECK();
    const int width = (X.dcvec ? 6 : 5) * hidden;
    reduce_partials_kernel<<<(width + RP_COLS - 1) / RP_COLS, RP_COLS * RP_ROWS, 0, st>>>(partials, dparams, EPI_PARTIAL_BLOCKS, width);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

}  // namespace alignn

using namespace alignn;

extern "C" int64_t alignn_gate_ln_bwd_partial_rows(void) { return EPI_PARTIAL_BLOCKS; }

static GateExtra plain_extra(int hidden) {
    GateExtra X;
    X.agge = nullptr; X.cvec = nullptr; X.stat_s = nullptr; X.agg_out = nullptr; X.dagg_lp = nullptr; X.dcvec = false; X.dy2 = nullptr; X.lddy2 = hidden; X.agg_rows = -1;
    X.heads = 1; X.ldxr = hidden; X.lddxr = hidden; X.rng_step = nullptr; X.x_lp = nullptr;
    return X;
}

static int gate_ln_fwd_impl(const float *agg, const void *xr, const float *x,
                            const float *wbeta, const float *gamma, const float *bias,
                            float *y, void *y_lp, float *beta, float *mean, float *rstd,
                            int64_t n_rows, int hidden, int dtype, float eps,
                            float p_drop, uint64_t seed, uint64_t offset, const GateExtra &X, bool extras, void *stream) {
    if (n_rows < 0 || hidden <= 0 || hidden > 2048) return ALIGNN_ERR_BAD_SHAPE;
    if (!(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (n_rows == 0) return ALIGNN_OK;
    if (!agg || !xr || (!x && !X.x_lp) || !wbeta || !gamma || !bias || !y || !beta || !mean || !rstd) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(agg) || !aligned16(xr) || !aligned16(x) || !aligned16(X.x_lp) || !aligned16(wbeta) || !aligned16(gamma) ||
        !aligned16(bias) || !aligned16(y) || !aligned16(y_lp) || !aligned16(X.agge) || !aligned16(X.cvec) ||
        !aligned16(X.agg_out) || (X.ldxr % 8))
        return ALIGNN_ERR_BAD_ARG;
    if (extras && (X.heads <= 0 || hidden % X.heads || (hidden / X.heads) % 8)) return ALIGNN_ERR_BAD_SHAPE;
    if (X.cvec && !X.stat_s) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ALIGNN_F32)
        return dispatch_epi_fwd<float>(agg, xr, x, wbeta, gamma, bias, y, y_lp, beta, mean, rstd, n_rows, hidden, eps,
                                       p_drop, seed, offset, X, extras, st);
    if (dtype == ALIGNN_BF16)
        return dispatch_epi_fwd<__nv_bfloat16>(agg, xr, x, wbeta, gamma, bias, y, y_lp, beta, mean, rstd, n_rows,
                                               hidden, eps, p_drop, seed, offset, X, extras, st);
    return ALIGNN_ERR_BAD_DTYPE;
}

extern "C" int alignn_gate_ln_fwd(const float *agg, const void *xr, const float *x,
                                  const float *wbeta, const float *gamma, const float *bias,
                                  float *y, void *y_lp, float *beta, float *mean, float *rstd,
                                  int64_t n_rows, int hidden, int dtype, float eps,
                                  float p_drop, uint64_t seed, uint64_t offset, void *stream) {
    return gate_ln_fwd_impl(agg, xr, x, wbeta, gamma, bias, y, y_lp, beta, mean, rstd, n_rows, hidden, dtype, eps,
                            p_drop, seed, offset, plain_extra(hidden), false, stream);
}

extern "C" int alignn_gate_ln_fwd2(const float *aggv, const void *agge, const float *cvec, const float *stat_s,
                                   int heads, const void *xr, int64_t ldxr, const float *x,
                                   const float *wbeta, const float *gamma, const float *bias,
                                   float *agg_out, float *y, void *y_lp, float *beta, float *mean, float *rstd,
                                   int64_t n_rows, int hidden, int dtype, float eps,
                                   float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream) {
    GateExtra X = plain_extra(hidden);
    X.agge = agge; X.cvec = cvec; X.stat_s = stat_s; X.agg_out = agg_out; X.heads = heads; X.ldxr = ldxr;
    X.rng_step = rng_step;
    if (rng_step && (hidden % 8 || hidden > 256 || (256 % hidden))) return ALIGNN_ERR_BAD_SHAPE;   // fast mapping only
    return gate_ln_fwd_impl(aggv, xr, x, wbeta, gamma, bias, y, y_lp, beta, mean, rstd, n_rows, hidden, dtype, eps,
                            p_drop, seed, offset, X, true, stream);
}

static int gate_ln_bwd_impl(const float *dy, const float *agg, const void *xr,
                            const float *wbeta, const float *gamma, const float *bias,
                            const float *beta, const float *mean, const float *rstd,
                            float *dagg, void *dxr, float *partials, float *dparams,
                            int64_t n_rows, int hidden, int dtype,
                            float p_drop, uint64_t seed, uint64_t offset, const GateExtra &X, bool extras, void *stream) {
    if (n_rows < 0 || hidden <= 0 || hidden > 2048) return ALIGNN_ERR_BAD_SHAPE;
    if (!(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (!partials || !dparams || !wbeta || !gamma || !bias) return ALIGNN_ERR_BAD_ARG;
    if (n_rows > 0 && ((!dy && !X.dy2) || !agg || !xr || !beta || !mean || !rstd || !dagg || !dxr)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(dy) || !aligned16(agg) || !aligned16(xr) || !aligned16(wbeta) || !aligned16(gamma) ||
        !aligned16(bias) || !aligned16(dagg) || !aligned16(dxr) || !aligned16(X.dagg_lp) || (X.ldxr % 8) || (X.lddxr % 8))
        return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ALIGNN_F32)
        return dispatch_epi_bwd<float>(dy, agg, xr, wbeta, gamma, bias, beta, mean, rstd, dagg, dxr, partials,
                                       dparams, n_rows, hidden, p_drop, seed, offset, X, extras, st);
    if (dtype == ALIGNN_BF16)
        return dispatch_epi_bwd<__nv_bfloat16>(dy, agg, xr, wbeta, gamma, bias, beta, mean, rstd, dagg, dxr, partials,
                                               dparams, n_rows, hidden, p_drop, seed, offset, X, extras, st);
    return ALIGNN_ERR_BAD_DTYPE;
}

extern "C" int alignn_gate_ln_bwd(const float *dy, const float *agg, const void *xr,
                                  const float *wbeta, const float *gamma, const float *bias,
                                  const float *beta, const float *mean, const float *rstd,
                                  float *dagg, void *dxr, float *partials, float *dparams,
                                  int64_t n_rows, int hidden, int dtype,
                                  float p_drop, uint64_t seed, uint64_t offset, void *stream) {
    return gate_ln_bwd_impl(dy, agg, xr, wbeta, gamma, bias, beta, mean, rstd, dagg, dxr, partials, dparams, n_rows,
                            hidden, dtype, p_drop, seed, offset, plain_extra(hidden), false, stream);
}

extern "C" int alignn_gate_ln_fwd3(const float *aggv, const void *agge, const float *cvec, const float *stat_s,
                                   int heads, int64_t agg_rows, const void *xr, int64_t ldxr, const float *x,
                                   const float *wbeta, const float *gamma, const float *bias,
                                   float *agg_out, float *y, void *y_lp, float *beta, float *mean, float *rstd,
                                   int64_t n_rows, int hidden, int dtype, float eps,
                                   float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream) {
    GateExtra X = plain_extra(hidden);
    X.agge = agge; X.cvec = cvec; X.stat_s = stat_s; X.agg_out = agg_out; X.heads = heads; X.ldxr = ldxr;
    X.rng_step = rng_step;
    X.agg_rows = (agg_rows < 0 || agg_rows > n_rows) ? -1 : agg_rows;
    if (hidden % 8 || hidden > 256 || (256 % hidden)) return ALIGNN_ERR_BAD_SHAPE;   // fast mapping only
    return gate_ln_fwd_impl(aggv, xr, x, wbeta, gamma, bias, y, y_lp, beta, mean, rstd, n_rows, hidden, dtype, eps,
                            p_drop, seed, offset, X, true, stream);
}

extern "C" int alignn_gate_ln_fwd4(const float *aggv, const void *agge, const float *cvec, const float *stat_s,
                                   int heads, int64_t agg_rows, const void *xr, int64_t ldxr, const float *x, const void *x_lp,
                                   const float *wbeta, const float *gamma, const float *bias,
                                   float *agg_out, float *y, void *y_lp, float *beta, float *mean, float *rstd,
                                   int64_t n_rows, int hidden, int dtype, float eps,
                                   float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream) {
    GateExtra X = plain_extra(hidden);
    X.agge = agge; X.cvec = cvec; X.stat_s = stat_s; X.agg_out = agg_out; X.heads = heads; X.ldxr = ldxr;
    X.rng_step = rng_step; X.x_lp = x ? nullptr : x_lp;
    X.agg_rows = (agg_rows < 0 || agg_rows > n_rows) ? -1 : agg_rows;
    if (hidden % 8 || hidden > 256 || (256 % hidden)) return ALIGNN_ERR_BAD_SHAPE;   // fast mapping only
    return gate_ln_fwd_impl(aggv, xr, x, wbeta, gamma, bias, y, y_lp, beta, mean, rstd, n_rows, hidden, dtype, eps,
                            p_drop, seed, offset, X, true, stream);
}

extern "C" int alignn_gate_ln_bwd3(const float *dy, const void *dy2, int64_t lddy2, const float *agg, const void *xr, int64_t ldxr,
                                   const float *wbeta, const float *gamma, const float *bias,
                                   const float *beta, const float *mean, const float *rstd,
                                   const float *stat_s, int heads, int64_t agg_rows,
                                   float *dagg, void *dagg_lp, void *dxr, int64_t lddxr, float *partials, float *dparams,
                                   int64_t n_rows, int hidden, int dtype,
                                   float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream) {
    GateExtra X = plain_extra(hidden);
    X.dagg_lp = dagg_lp; X.ldxr = ldxr; X.lddxr = lddxr; X.rng_step = rng_step;
    X.stat_s = stat_s; X.heads = heads; X.dcvec = true; X.dy2 = dy2; X.lddy2 = lddy2;
    X.agg_rows = (agg_rows < 0 || agg_rows > n_rows) ? -1 : agg_rows;
    if ((!dy && !dy2) || (dy2 && (lddy2 % 8 || !aligned16(dy2)))) return ALIGNN_ERR_BAD_ARG;
    if (!stat_s || heads <= 0 || hidden % heads || (hidden / heads) % 8) return ALIGNN_ERR_BAD_ARG;
    if (hidden % 8 || hidden > 256 || (256 % hidden)) return ALIGNN_ERR_BAD_SHAPE;   // fast mapping only
    return gate_ln_bwd_impl(dy, agg, xr, wbeta, gamma, bias, beta, mean, rstd, dagg, dxr, partials, dparams, n_rows,
                            hidden, dtype, p_drop, seed, offset, X, true, stream);
}

extern "C" int alignn_gate_ln_bwd2(const float *dy, const float *agg, const void *xr, int64_t ldxr,
                                   const float *wbeta, const float *gamma, const float *bias,
                                   const float *beta, const float *mean, const float *rstd,
                                   float *dagg, void *dagg_lp, void *dxr, int64_t lddxr, float *partials, float *dparams,
                                   int64_t n_rows, int hidden, int dtype,
                                   float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream) {
    GateExtra X = plain_extra(hidden);
    X.dagg_lp = dagg_lp; X.ldxr = ldxr; X.lddxr = lddxr; X.rng_step = rng_step;
    if (rng_step && (hidden % 8 || hidden > 256 || (256 % hidden))) return ALIGNN_ERR_BAD_SHAPE;
    return gate_ln_bwd_impl(dy, agg, xr, wbeta, gamma, bias, beta, mean, rstd, dagg, dxr, partials, dparams, n_rows,
                            hidden, dtype, p_drop, seed, offset, X, true, stream);
}

// ---- column sums of a [rows, width] matrix (bias gradients of the stacked node projections) -----------------------
// Replaces `dproj.sum(0)` of torch autograd's AddmmBackward for lin_query/key/value/skip (reference train.py:691 through
// PyG TransformerConv's four Linears).  Deterministic: per-thread register sums over a fixed row stride, fixed-order fold
// of the row groups of a CTA, then reduce_partials_kernel over the CTAs.
namespace alignn {

constexpr int CS_THREADS = 256;
constexpr int CS_BLOCKS = 296;

template <typename T>
__global__ void __launch_bounds__(CS_THREADS)
colsum_kernel(const T *__restrict__ x, int64_t ld, int64_t n_rows, int width, float *__restrict__ partials) {
    __shared__ float red[CS_THREADS * 8];
    const int lanes = width >> 3;                   // threads per row
    const int groups = CS_THREADS / lanes;          // rows per CTA iteration (threads beyond lanes * groups idle)
    const int sub = threadIdx.x % lanes, grp = threadIdx.x / lanes;
    const int64_t stride = (int64_t)gridDim.x * groups;
    F8 acc;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc.v[c] = 0.f;
    int64_t row = grp < groups ? (int64_t)blockIdx.x * groups + grp : n_rows;
    for (; row + 3 * stride < n_rows; row += 4 * stride) {      // 4 independent 16-byte loads in flight per thread
        const F8 a = ld8(x + row * ld + sub * 8), b = ld8(x + (row + stride) * ld + sub * 8);
        const F8 c2 = ld8(x + (row + 2 * stride) * ld + sub * 8), d = ld8(x + (row + 3 * stride) * ld + sub * 8);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc.v[c] += (a.v[c] + b.v[c]) + (c2.v[c] + d.v[c]);
    }
    for (; row < n_rows; row += stride) {
        const F8 a = ld8(x + row * ld + sub * 8);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc.v[c] += a.v[c];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) red[threadIdx.x * 8 + c] = acc.v[c];
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += CS_THREADS) {
        const int s8 = i >> 3, c = i & 7;
        float s = 0.f;
        for (int g2 = 0; g2 < groups; ++g2) s += red[(g2 * lanes + s8) * 8 + c];
        partials[(int64_t)blockIdx.x * width + i] = s;
    }
}

}  // namespace alignn

extern "C" int64_t alignn_colsum_partial_floats(int width) { return (int64_t)CS_BLOCKS * width; }

extern "C" int alignn_colsum_supported(int width) {
    return width >= 8 && width <= 2048 && width % 8 == 0;
}

extern "C" int alignn_colsum(const void *x, int64_t ld, int64_t n_rows, int width, int dtype, float *partials,
                             float *out, void *stream) {
    if (!alignn_colsum_supported(width) || n_rows < 0 || ld < width || (ld % 8)) return ALIGNN_ERR_BAD_SHAPE;
    if (!out || !partials || (n_rows > 0 && !x) || !aligned16(x)) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ALIGNN_F32)
        colsum_kernel<float><<<CS_BLOCKS, CS_THREADS, 0, st>>>((const float *)x, ld, n_rows, width, partials);
    else if (dtype == ALIGNN_BF16)
        colsum_kernel<__nv_bfloat16><<<CS_BLOCKS, CS_THREADS, 0, st>>>((const __nv_bfloat16 *)x, ld, n_rows, width, partials);
    else
        return ALIGNN_ERR_BAD_DTYPE;
    ALIGNN_LAUNCH_CHECK();
    reduce_partials_kernel<<<(width + RP_COLS - 1) / RP_COLS, RP_COLS * RP_ROWS, 0, st>>>(partials, out, CS_BLOCKS, width);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
