// Tensor-core edge attention (hidden = 256, heads = 4, bf16): the streaming conv of edgeattn.cu with the per-edge
// FMA work moved onto mma.sync.m16n8k16.
//
// Same arithmetic and the same reference interfaces as edgeattn.cu (PyG TransformerConv.message + utils.softmax +
// 'add' aggregation at scripts/train.py:315,334, with `lin_edge`, the second angle-encoder Linear (train.py:360-364)
// and `edge_proj` (train.py:324,333) folded into per-NODE operands qt / gt).  A warp owns a contiguous range of
// target rows and walks it in chunks of <= 16 in-edges of ONE target row i:
//
//   phase 1   S[16 edges x 8]  = F[16 x 256] . QT_i[256 x 8]  +  K[16 x 256] . Qbd_i[256 x 8]        (32 MMAs)
//             QT_i column t = qt_i,t (t < 4; columns 4..7 repeat 0..3);  Qbd_i column t = q_i restricted to head t
//   softmax   online over chunks (running max / denominators per head in registers), Philox dropout
//   phase 2   acc[256 ch x 8] += F^T[256 x 16] . P[16 x 8 | cols 0..3]  +  V^T[256 x 16] . P[16 x 8 | cols 4..7]
//             -> columns 0..3 = abar_i,t (all channels), columns 4..7 = sum_j a~ v_j (own-head channels used)  (32 MMAs)
//
// F, K, V rows are gathered with cp.async (16 bytes per lane, one warp instruction per row; per-row TMA bulk copies
// cap at ~45 cycles per copy per SM, measured: profiles/r01_v3_notes.txt) into a warp-private double-buffered ring
// (rows padded to 528 B so ldmatrix is conflict-free); B operands of phase 1 come straight from global memory in a channel
// permutation shared by both operands (one 16-byte load feeds two MMAs); P goes from accumulator to B-fragment
// layout with two movmatrix.  No atomics; every row's result is independent of the launch geometry.
#include <stdlib.h>
#include <math.h>

#include "mma.cuh"

namespace alignn {

constexpr int MM_HID = 256;
constexpr int MM_HEADS = 4;
constexpr int MM_E = 16;                      // edges per chunk
constexpr int MM_ROWB = 528;                  // bytes per staged row (512 + 16: 8 consecutive rows hit 8 distinct 16-byte bank groups)
constexpr int MM_TILE = MM_E * MM_ROWB;       // 8448
constexpr int MM_WARPS = 8;                   // two per scheduler
constexpr int MM_STAGES = 1;                  // (K, V, F) tile sets per warp.  One set x 8 warps fills the SM's shared memory:
                                              // a warp gathers, waits and computes in turn and its scheduler partner covers
                                              // the wait -- measured faster than 4 warps with two sets each (one warp per
                                              // scheduler, nothing to cover MMA latency or the row-start loads)
constexpr int MM_STG = 260;                   // floats per column of the epilogue staging buffer

struct MmFwdParams {
    const __nv_bfloat16 *q, *k, *v;   // strided rows
    const __nv_bfloat16 *qt;          // [4, Nn, 256]
    const __nv_bfloat16 *feat;        // [Ne, 256]
    const int32_t *rowptr, *col, *eid;
    float *aggv;                      // [Nn, 256]
    __nv_bfloat16 *abar;              // [4, Nn, 256]
    float *stat_m, *stat_z, *stat_s;  // [Nn, 4]
    int64_t n_nodes, n_edges;
    int64_t ldq, ldk, ldv;
    int64_t ldqt, hsqt, ldab, hsab;   // qt[row * ldqt + t * hsqt + ch], abar likewise
    float scale_log2, p_drop, inv_keep;
    uint64_t seed, offset;
    const uint64_t *rng_step;         // optional device counter added to `offset`
};

struct FwdRowFrags {
    uint4 qt[8];   // lane (g, q): qt[g & 3][row][32 c + 8 q .. +7], c = 0..7
    uint32_t qv[8];   // own-head q in natural channel order: k-step i -> q[row][64 (g & 3) + 16 i + 2 q, +1] and [.. + 8 + 2 q, +1]
};

__device__ __forceinline__ void load_fwd_row(FwdRowFrags &rf, const MmFwdParams &P, int64_t row, int g, int q) {
    const int t = g & 3;
    const uint4 *pt = reinterpret_cast<const uint4 *>(P.qt + row * P.ldqt + (int64_t)t * P.hsqt) + q;
#pragma unroll
    for (int c = 0; c < 8; ++c) rf.qt[c] = __ldg(pt + 4 * c);
    const uint32_t *pq = reinterpret_cast<const uint32_t *>(P.q + row * P.ldq + 64 * t) + q;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rf.qv[2 * i] = __ldg(pq + 8 * i);
        rf.qv[2 * i + 1] = __ldg(pq + 8 * i + 4);
    }
}

__global__ void __launch_bounds__(MM_WARPS * 32, 1)
edgeattn_mma_fwd_kernel(const MmFwdParams P) {
    constexpr int PER_WARP = MM_STAGES * 3 * MM_TILE + 64;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    unsigned char *base = smem_raw + (size_t)warp * PER_WARP;
    const uint32_t base_u32 = smem_u32(base);

    const int64_t W = (int64_t)gridDim.x * MM_WARPS, w = (int64_t)blockIdx.x * MM_WARPS + warp;
    const int64_t total = P.n_edges + (int64_t)ROW_KAPPA * P.n_nodes;
    const int64_t r0 = row_bound(P.rowptr, P.n_nodes, total * w / W);
    const int64_t r1 = row_bound(P.rowptr, P.n_nodes, total * (w + 1) / W);
    if (r0 >= r1) return;
    const int e_end = __ldg(P.rowptr + r1);

    // padding rows of a chunk are multiplied by exact zeros in phase 2: they must hold finite values
    for (int off = lane * 16; off < MM_STAGES * 3 * MM_TILE; off += 32 * 16) sts128(base_u32 + off, make_uint4(0, 0, 0, 0));
    __syncwarp();

    StageCursor<MM_E> prod, cons;
    prod.init(P.rowptr, r0, r1);
    cons = prod;
    int wbase = prod.pos;
    IndexWindow wcol, weid;
    wcol.init(P.col, wbase, e_end, lane);
    weid.init(P.eid, wbase, e_end, lane);

    // gather one chunk: every lane copies 16 bytes of every row (cp.async, L1 bypass); one commit group per stage
    auto issue = [&](int s) {
        if (!prod.done()) {
            const int n = prod.count();
            const int o = prod.pos + min(lane & 15, n - 1) - wbase;
            const int j = wcol.get(o), id = weid.get(o);
            // one warp instruction copies a 128-byte piece of FOUR rows (lane = 8 * row-in-group + 16-byte slot): the row
            // addresses are formed once per group of four rows, the four pieces of a row are immediate offsets
            const int sub = lane >> 3, slot = (lane & 7) * 8;
            const uint32_t dst = base_u32 + (uint32_t)s * 3 * MM_TILE + (uint32_t)sub * MM_ROWB + (lane & 7) * 16;
#pragma unroll
            for (int grp = 0; grp < MM_E / 4; ++grp) {
                const int u = 4 * grp + sub;
                const int ju = __shfl_sync(FULL, j, u), idu = __shfl_sync(FULL, id, u);
                if (u < n) {
                    const __nv_bfloat16 *pk = P.k + (int64_t)ju * P.ldk + slot, *pv = P.v + (int64_t)ju * P.ldv + slot;
                    const __nv_bfloat16 *pf = P.feat + (int64_t)idu * MM_HID + slot;
                    const uint32_t d = dst + (uint32_t)(4 * grp) * MM_ROWB;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        cp_async16(d + i * 128, pk + i * 64);
                        cp_async16(d + MM_TILE + i * 128, pv + i * 64);
                        cp_async16(d + 2 * MM_TILE + i * 128, pf + i * 64);
                    }
                }
            }
            prod.advance(P.rowptr);
            if (prod.pos - wbase >= 32 && !prod.done()) {
                wbase += 32;
                wcol.shift(P.col, wbase, e_end, lane);
                weid.shift(P.eid, wbase, e_end, lane);
            }
        }
        cp_async_commit();   // always: keeps the group count uniform
    };

    auto zero_rows = [&](int64_t lo, int64_t hi) {   // rows without in-edges: all outputs are zero
        F8 zf;
#pragma unroll
        for (int c = 0; c < 8; ++c) zf.v[c] = 0.f;
        for (int64_t r = lo; r < hi; ++r) {
            st8(P.aggv + r * MM_HID + lane * 8, zf);
#pragma unroll
            for (int t = 0; t < MM_HEADS; ++t) st8(P.abar + r * P.ldab + (int64_t)t * P.hsab + lane * 8, zf);
            if (lane < MM_HEADS) {
                P.stat_m[r * MM_HEADS + lane] = 0.f;
                P.stat_z[r * MM_HEADS + lane] = 0.f;
                P.stat_s[r * MM_HEADS + lane] = 0.f;
            }
        }
    };

    issue(0);

    FwdRowFrags rf;
    load_fwd_row(rf, P, cons.row, g, q);
    float acc[16][4];
    float m0 = -INFINITY, m1 = -INFINITY, z0 = 0.f, z1 = 0.f, zd0 = 0.f, zd1 = 0.f;
    int64_t next_unwritten = r0;
    const int t_own = g & 3;

    for (int it = 0; !cons.done(); ++it) {
        const int s = MM_STAGES == 2 ? (it & 1) : 0;
        if (MM_STAGES == 2) issue(s ^ 1);
        StageCursor<MM_E> nxt = cons;
        nxt.advance(P.rowptr);
        const int n = cons.count();
        const int64_t row = cons.row;
        const bool first = cons.first, last = cons.last();
        const int pos = cons.pos;
        if (first) {
            zero_rows(next_unwritten, row);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            m0 = m1 = -INFINITY;
            z0 = z1 = zd0 = zd1 = 0.f;
        }
        cp_async_wait<MM_STAGES - 1>();   // everything but the group just committed (the next chunk, if prefetched) has landed
        __syncwarp();
        const uint32_t ktile = base_u32 + (uint32_t)s * 3 * MM_TILE, vtile = ktile + MM_TILE, ftile = ktile + 2 * MM_TILE;

        // ---- phase 1: logits ------------------------------------------------------------------------------
        float c[4];
        {
            // four independent accumulator chains (8 MMAs each) instead of one of 32: with one warp per scheduler nothing
            // else hides the MMA latency
            float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb_[4] = {0.f, 0.f, 0.f, 0.f};
            float ck0[4] = {0.f, 0.f, 0.f, 0.f}, ck1[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t fa = ftile + g * MM_ROWB + q * 16;
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                const uint4 x = lds128(fa + cb * 64), y = lds128(fa + 8 * MM_ROWB + cb * 64);
                mma_bf16(ca, x.x, y.x, x.y, y.y, rf.qt[cb].x, rf.qt[cb].y);
                mma_bf16(cb_, x.z, y.z, x.w, y.w, rf.qt[cb].z, rf.qt[cb].w);
            }
            // K A-fragments with ldmatrix: conflict-free on the 528-byte-stride tile (per-lane 128-bit reads are 2-way conflicted)
            const uint32_t aoff = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * MM_ROWB + (lane >> 4) * 16);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                uint32_t a[4];
                ldsm_x4(a, ktile + aoff + kk * 32);
                const bool own = (kk >> 2) == t_own;
                mma_bf16((kk & 1) ? ck1 : ck0, a[0], a[1], a[2], a[3], own ? rf.qv[2 * (kk & 3)] : 0u, own ? rf.qv[2 * (kk & 3) + 1] : 0u);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i] = (ca[i] + cb_[i]) + (ck0[i] + ck1[i]);
        }
        // ---- online softmax over the chunk (columns = heads; rows g and g+8) --------------------------------
        const bool v0 = g < n, v1 = g + 8 < n;
        const float s00 = v0 ? c[0] * P.scale_log2 : -INFINITY, s01 = v0 ? c[1] * P.scale_log2 : -INFINITY;
        const float s10 = v1 ? c[2] * P.scale_log2 : -INFINITY, s11 = v1 ? c[3] * P.scale_log2 : -INFINITY;
        const float mn0 = fmaxf(m0, colmax8(fmaxf(s00, s10))), mn1 = fmaxf(m1, colmax8(fmaxf(s01, s11)));
        const float corr0 = fast_exp2(m0 - mn0), corr1 = fast_exp2(m1 - mn1);   // m = -inf -> 0
        m0 = mn0;
        m1 = mn1;
        float p00 = fast_exp2(s00 - mn0), p01 = fast_exp2(s01 - mn1);
        float p10 = fast_exp2(s10 - mn0), p11 = fast_exp2(s11 - mn1);
        z0 = z0 * corr0 + (p00 + p10);
        z1 = z1 * corr1 + (p01 + p11);
        if (P.p_drop > 0.f) {
            const uint64_t roff = P.offset + (P.rng_step ? *P.rng_step : 0ull);
            float k00, k01, k10, k11;
            dropout_scale_quad(P.seed, roff, (uint64_t)pos, g, q, P.p_drop, P.inv_keep, k00, k01, k10, k11);
            p00 *= k00;
            p01 *= k01;
            p10 *= k10;
            p11 *= k11;
        }
        zd0 = zd0 * corr0 + (p00 + p10);
        zd1 = zd1 * corr1 + (p01 + p11);
        if (!first) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                acc[j][0] *= corr0;
                acc[j][1] *= corr1;
                acc[j][2] *= corr0;
                acc[j][3] *= corr1;
            }
        }
        const uint32_t tlo = movmatrix_trans(pack_bf16(p00, p01)), thi = movmatrix_trans(pack_bf16(p10, p11));
        const uint32_t bf0 = g < 4 ? tlo : 0u, bf1 = g < 4 ? thi : 0u;
        const uint32_t bv0 = g < 4 ? 0u : tlo, bv1 = g < 4 ? 0u : thi;

        // operands of the next target row: in flight while phase 2 runs (the registers are dead after phase 1)
        if (nxt.first && !nxt.done()) load_fwd_row(rf, P, nxt.row, g, q);

        // ---- phase 2: weighted sums --------------------------------------------------------------------------
        {
            const uint32_t toff = (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * MM_ROWB + ((lane >> 3) & 1) * 16);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t a[4];
                ldsm_x4_trans(a, ftile + toff + j * 32);
                mma_bf16(acc[j], a[0], a[1], a[2], a[3], bf0, bf1);
                ldsm_x4_trans(a, vtile + toff + j * 32);
                mma_bf16(acc[j], a[0], a[1], a[2], a[3], bv0, bv1);
            }
        }
        // ---- row epilogue -----------------------------------------------------------------------------------
        if (last) {
            const float zs0 = colsum8(z0), zs1 = colsum8(z1), zds0 = colsum8(zd0), zds1 = colsum8(zd1);
            const float inv0 = 1.0f / (zs0 + 1e-16f), inv1 = 1.0f / (zs1 + 1e-16f);
            if (g == 0 && q < 2) {
                P.stat_m[row * MM_HEADS + 2 * q] = m0;
                P.stat_m[row * MM_HEADS + 2 * q + 1] = m1;
                P.stat_z[row * MM_HEADS + 2 * q] = zs0;
                P.stat_z[row * MM_HEADS + 2 * q + 1] = zs1;
                P.stat_s[row * MM_HEADS + 2 * q] = zds0 * inv0;
                P.stat_s[row * MM_HEADS + 2 * q + 1] = zds1 * inv1;
            }
            // transpose through the (now dead) K slot: stage[col][channel], 260 floats per column
            __syncwarp();
            const uint32_t st0 = ktile + (uint32_t)((2 * q) * MM_STG + g) * 4, st1 = st0 + MM_STG * 4;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                sts32f(st0 + j * 64, acc[j][0] * inv0);
                sts32f(st1 + j * 64, acc[j][1] * inv1);
                sts32f(st0 + j * 64 + 32, acc[j][2] * inv0);
                sts32f(st1 + j * 64 + 32, acc[j][3] * inv1);
            }
            __syncwarp();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int ch = half * 128 + 4 * lane;
#pragma unroll
                for (int t = 0; t < MM_HEADS; ++t) {
                    const float4 x = lds128f(ktile + (uint32_t)(t * MM_STG + ch) * 4);
                    uint2 o;
                    o.x = pack_bf16(x.x, x.y);
                    o.y = pack_bf16(x.z, x.w);
                    *reinterpret_cast<uint2 *>(P.abar + row * P.ldab + (int64_t)t * P.hsab + ch) = o;
                }
                const int head = ch >> 6;
                const float4 y = lds128f(ktile + (uint32_t)((4 + head) * MM_STG + ch) * 4);
                *reinterpret_cast<float4 *>(P.aggv + row * MM_HID + ch) = y;
            }
            next_unwritten = row + 1;
        }
        __syncwarp();          // every lane is done with stage s before it is refilled
        if (MM_STAGES == 1) issue(0);
        cons = nxt;
    }
    zero_rows(next_unwritten, r1);
}

// ------------------------------------------------------------------------------------------------------------
// backward, target-sorted pass
//
//   phase 1   SD[16 edges x 8] = F . [QT_i | GT_i]  +  K . [Qbd_i | 0]  +  V . [0 | Gbd_i]          (48 MMAs)
//             columns 0..3 = raw logits, columns 4..7 = d a~ (before the c-term); lanes q and q^2 swap halves
//   coefficients  a = exp2(s - m)/z,  a~ = a drop,  ds = a ((d + Gc) drop - D) / sqrt(C)  -> coef[eid] = (a~, ds)
//   phase 2   acc[256 ch x 8] += F^T . DS[16 x 8 | cols 0..3]  +  K^T . DS[16 x 8 | cols 4..7]       (32 MMAs)
//             -> columns 0..3 = bbar_i,t, columns 4..7 = dq_i (own-head channels)
//   features  DF[16 edges x 256] = [DS | A~][16 x 8] . [QT_i ; GT_i][8 x 256]  (+ running sum, ReLU mask)  (32 MMAs)
// ------------------------------------------------------------------------------------------------------------
struct MmBwdParams {
    const float *dagg, *agg;                 // [Nn, 256] f32
    const __nv_bfloat16 *dagg_lp;            // [Nn, 256] bf16 copy of dagg (B operand of the V part)
    const __nv_bfloat16 *q, *k, *v;          // strided rows
    const __nv_bfloat16 *qt, *gt;            // [4, Nn, 256]
    const float *cvec;                       // [256] f32 or null
    const __nv_bfloat16 *feat;               // [Ne, 256]
    const float *stat_m, *stat_z;
    const int32_t *rowptr, *col, *eid;
    __nv_bfloat16 *dq;                       // strided rows (lddq)
    __nv_bfloat16 *bbar;                     // [4, Nn, 256]
    float *coef;                             // [Ne, 8] by caller edge id: (a~_0..3, ds_0..3)
    const __nv_bfloat16 *df_in;              // [Ne, 256] running sum or null
    __nv_bfloat16 *df_out;                   // [Ne, 256] (row stride lddf) or null (feature gradient not wanted)
    int64_t lddf;                            // row stride (elements) of df_in / df_out
    int64_t n_nodes, n_edges;
    int64_t ldq, ldk, ldv, lddq;
    int64_t ldqt, hsqt, ldgt, hsgt, ldbb, hsbb;
    float scale, scale_log2, p_drop, inv_keep;
    uint64_t seed, offset;
    const uint64_t *rng_step;
    int relu_mask;
};

struct BwdRowFrags {
    uint4 x[8];    // lane (g, q): g < 4: qt[g][row][32 c + 8 q ..], else gt[g - 4][row][...]
    uint32_t kv[8];   // g < 4: q[row], else dagg_lp[row]: own-head block in natural channel order (see FwdRowFrags::qv)
};

__device__ __forceinline__ void load_bwd_row(BwdRowFrags &rf, const MmBwdParams &P, int64_t row, int g, int q) {
    const int t = g & 3;
    const __nv_bfloat16 *wide = g < 4 ? P.qt + row * P.ldqt + (int64_t)t * P.hsqt : P.gt + row * P.ldgt + (int64_t)t * P.hsgt;
    const uint4 *pt = reinterpret_cast<const uint4 *>(wide) + q;
#pragma unroll
    for (int c = 0; c < 8; ++c) rf.x[c] = __ldg(pt + 4 * c);
    const __nv_bfloat16 *nar = g < 4 ? P.q + row * P.ldq : P.dagg_lp + row * MM_HID;
    const uint32_t *pq = reinterpret_cast<const uint32_t *>(nar + 64 * t) + q;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rf.kv[2 * i] = __ldg(pq + 8 * i);
        rf.kv[2 * i + 1] = __ldg(pq + 8 * i + 4);
    }
}

template <bool ACCUM>
__global__ void __launch_bounds__(MM_WARPS * 32, 1)
edgeattn_mma_bwd_kernel(const MmBwdParams P) {
    constexpr int PER_WARP = MM_STAGES * 3 * MM_TILE + 64 + 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    unsigned char *base = smem_raw + (size_t)warp * PER_WARP;
    int *eid_stash = reinterpret_cast<int *>(base + MM_STAGES * 3 * MM_TILE + 64);   // [MM_STAGES][16]
    const uint32_t base_u32 = smem_u32(base);

    const int64_t W = (int64_t)gridDim.x * MM_WARPS, w = (int64_t)blockIdx.x * MM_WARPS + warp;
    const int64_t total = P.n_edges + (int64_t)ROW_KAPPA * P.n_nodes;
    const int64_t r0 = row_bound(P.rowptr, P.n_nodes, total * w / W);
    const int64_t r1 = row_bound(P.rowptr, P.n_nodes, total * (w + 1) / W);
    if (r0 >= r1) return;
    const int e_end = __ldg(P.rowptr + r1);

    for (int off = lane * 16; off < MM_STAGES * 3 * MM_TILE; off += 32 * 16) sts128(base_u32 + off, make_uint4(0, 0, 0, 0));
    __syncwarp();

    StageCursor<MM_E> prod, cons;
    prod.init(P.rowptr, r0, r1);
    cons = prod;
    int wbase = prod.pos;
    IndexWindow wcol, weid;
    wcol.init(P.col, wbase, e_end, lane);
    weid.init(P.eid, wbase, e_end, lane);

    auto issue = [&](int s) {
        if (!prod.done()) {
            const int n = prod.count();
            const int o = prod.pos + min(lane & 15, n - 1) - wbase;
            const int j = wcol.get(o), id = weid.get(o);
            if (lane < 16) eid_stash[s * 16 + lane] = id;
            // one warp instruction copies a 128-byte piece of FOUR rows (lane = 8 * row-in-group + 16-byte slot): the row
            // addresses are formed once per group of four rows, the four pieces of a row are immediate offsets
            const int sub = lane >> 3, slot = (lane & 7) * 8;
            const uint32_t dst = base_u32 + (uint32_t)s * 3 * MM_TILE + (uint32_t)sub * MM_ROWB + (lane & 7) * 16;
#pragma unroll
            for (int grp = 0; grp < MM_E / 4; ++grp) {
                const int u = 4 * grp + sub;
                const int ju = __shfl_sync(FULL, j, u), idu = __shfl_sync(FULL, id, u);
                if (u < n) {
                    const __nv_bfloat16 *pk = P.k + (int64_t)ju * P.ldk + slot, *pv = P.v + (int64_t)ju * P.ldv + slot;
                    const __nv_bfloat16 *pf = P.feat + (int64_t)idu * MM_HID + slot;
                    const uint32_t d = dst + (uint32_t)(4 * grp) * MM_ROWB;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        cp_async16(d + i * 128, pk + i * 64);
                        cp_async16(d + MM_TILE + i * 128, pv + i * 64);
                        cp_async16(d + 2 * MM_TILE + i * 128, pf + i * 64);
                    }
                }
            }
            prod.advance(P.rowptr);
            if (prod.pos - wbase >= 32 && !prod.done()) {
                wbase += 32;
                wcol.shift(P.col, wbase, e_end, lane);
                weid.shift(P.eid, wbase, e_end, lane);
            }
        }
        cp_async_commit();
    };

    auto zero_rows = [&](int64_t lo, int64_t hi) {
        F8 zf;
#pragma unroll
        for (int c = 0; c < 8; ++c) zf.v[c] = 0.f;
        for (int64_t r = lo; r < hi; ++r) {
            st8(P.dq + r * P.lddq + lane * 8, zf);
#pragma unroll
            for (int t = 0; t < MM_HEADS; ++t) st8(P.bbar + r * P.ldbb + (int64_t)t * P.hsbb + lane * 8, zf);
        }
    };

    issue(0);
    __syncwarp();   // eid stash of the prologue stage is visible to every lane

    BwdRowFrags rf;
    load_bwd_row(rf, P, cons.row, g, q);
    float acc[16][4];
    float D0 = 0.f, D1 = 0.f, G0 = 0.f, G1 = 0.f, mh0 = 0.f, mh1 = 0.f, iz0 = 0.f, iz1 = 0.f;
    int64_t next_unwritten = r0;
    const int hsel = 2 * (q & 1);   // this lane's coefficient columns are heads hsel, hsel + 1

    for (int it = 0; !cons.done(); ++it) {
        const int s = MM_STAGES == 2 ? (it & 1) : 0;
        if (MM_STAGES == 2) issue(s ^ 1);
        StageCursor<MM_E> nxt = cons;
        nxt.advance(P.rowptr);
        const int n = cons.count();
        const int64_t row = cons.row;
        const bool first = cons.first, last = cons.last();
        const int pos = cons.pos;
        if (first) {
            zero_rows(next_unwritten, row);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            // per-head row constants: D = <dagg, agg>, Gc = <dagg, c>, softmax statistics
            const F8 gf = ld8(P.dagg + row * MM_HID + lane * 8), af = ld8(P.agg + row * MM_HID + lane * 8);
            float dpart = 0.f, gpart = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) dpart = fmaf(gf.v[c], af.v[c], dpart);
            if (P.cvec) {
                const F8 cf = ld8(P.cvec + lane * 8);
#pragma unroll
                for (int c = 0; c < 8; ++c) gpart = fmaf(gf.v[c], cf.v[c], gpart);
            }
            const float Dh = group_sum<8>(dpart), Gh = group_sum<8>(gpart);   // lanes 8t..8t+7 hold head t
            D0 = __shfl_sync(FULL, Dh, 8 * hsel);
            D1 = __shfl_sync(FULL, Dh, 8 * hsel + 8);
            G0 = __shfl_sync(FULL, Gh, 8 * hsel);
            G1 = __shfl_sync(FULL, Gh, 8 * hsel + 8);
            mh0 = __ldg(P.stat_m + row * MM_HEADS + hsel);
            mh1 = __ldg(P.stat_m + row * MM_HEADS + hsel + 1);
            iz0 = 1.0f / (__ldg(P.stat_z + row * MM_HEADS + hsel) + 1e-16f);
            iz1 = 1.0f / (__ldg(P.stat_z + row * MM_HEADS + hsel + 1) + 1e-16f);
        }
        cp_async_wait<MM_STAGES - 1>();   // everything but the group just committed (the next chunk, if prefetched) has landed
        __syncwarp();
        const uint32_t ktile = base_u32 + (uint32_t)s * 3 * MM_TILE, vtile = ktile + MM_TILE, ftile = ktile + 2 * MM_TILE;

        // ---- phase 1: logits (cols 0..3) and d a~ (cols 4..7) --------------------------------------------------
        float c[4];
        {
            // four independent accumulator chains instead of one of 48 (see the forward kernel)
            float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb_[4] = {0.f, 0.f, 0.f, 0.f};
            float ck0[4] = {0.f, 0.f, 0.f, 0.f}, ck1[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t fa = ftile + g * MM_ROWB + q * 16;
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                const uint4 x = lds128(fa + cb * 64), y = lds128(fa + 8 * MM_ROWB + cb * 64);
                mma_bf16(ca, x.x, y.x, x.y, y.y, rf.x[cb].x, rf.x[cb].y);
                mma_bf16(cb_, x.z, y.z, x.w, y.w, rf.x[cb].z, rf.x[cb].w);
            }
            // K / V A-fragments with ldmatrix (conflict-free on the 528-byte-stride tiles)
            const uint32_t aoff = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * MM_ROWB + (lane >> 4) * 16);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                uint32_t a[4];
                ldsm_x4(a, ktile + aoff + kk * 32);
                const bool own = (kk >> 2) == g;          // g < 4 and own head
                mma_bf16((kk & 1) ? ck1 : ck0, a[0], a[1], a[2], a[3], own ? rf.kv[2 * (kk & 3)] : 0u, own ? rf.kv[2 * (kk & 3) + 1] : 0u);
            }
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                uint32_t a[4];
                ldsm_x4(a, vtile + aoff + kk * 32);
                const bool own = (kk >> 2) + 4 == g;      // g >= 4 and own head
                mma_bf16((kk & 1) ? ck1 : ck0, a[0], a[1], a[2], a[3], own ? rf.kv[2 * (kk & 3)] : 0u, own ? rf.kv[2 * (kk & 3) + 1] : 0u);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i] = (ca[i] + cb_[i]) + (ck0[i] + ck1[i]);
        }
        // lanes q < 2 hold logits of heads (2q, 2q+1), lanes q >= 2 hold d of heads (2(q-2), 2(q-2)+1): swap halves
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __shfl_xor_sync(FULL, c[i], 2);
        const bool lo_half = q < 2;
        const float sl00 = lo_half ? c[0] : o[0], sl01 = lo_half ? c[1] : o[1];
        const float sl10 = lo_half ? c[2] : o[2], sl11 = lo_half ? c[3] : o[3];
        const float d00 = lo_half ? o[0] : c[0], d01 = lo_half ? o[1] : c[1];
        const float d10 = lo_half ? o[2] : c[2], d11 = lo_half ? o[3] : c[3];
        float dr00 = 1.f, dr01 = 1.f, dr10 = 1.f, dr11 = 1.f;
        if (P.p_drop > 0.f) {
            const uint64_t roff = P.offset + (P.rng_step ? *P.rng_step : 0ull);
            dropout_scale_quad(P.seed, roff, (uint64_t)pos, g, q, P.p_drop, P.inv_keep, dr00, dr01, dr10, dr11);
        }
        const bool v0 = g < n, v1 = g + 8 < n;
        const float a00 = v0 ? fast_exp2(sl00 * P.scale_log2 - mh0) * iz0 : 0.f;
        const float a01 = v0 ? fast_exp2(sl01 * P.scale_log2 - mh1) * iz1 : 0.f;
        const float a10 = v1 ? fast_exp2(sl10 * P.scale_log2 - mh0) * iz0 : 0.f;
        const float a11 = v1 ? fast_exp2(sl11 * P.scale_log2 - mh1) * iz1 : 0.f;
        const float at00 = a00 * dr00, at01 = a01 * dr01, at10 = a10 * dr10, at11 = a11 * dr11;
        // padding rows may hold arbitrary bit patterns (the staging buffer aliases a tile): select, never multiply
        const float ds00 = v0 ? a00 * ((d00 + G0) * dr00 - D0) * P.scale : 0.f;
        const float ds01 = v0 ? a01 * ((d01 + G1) * dr01 - D1) * P.scale : 0.f;
        const float ds10 = v1 ? a10 * ((d10 + G0) * dr10 - D0) * P.scale : 0.f;
        const float ds11 = v1 ? a11 * ((d11 + G1) * dr11 - D1) * P.scale : 0.f;
        const int id0 = eid_stash[s * 16 + min(g, n - 1)], id1 = eid_stash[s * 16 + min(g + 8, n - 1)];
        {   // coef row = (a~_0..3, ds_0..3): lanes q < 2 store a~, lanes q >= 2 store ds
            const float2 w0 = lo_half ? make_float2(at00, at01) : make_float2(ds00, ds01);
            const float2 w1 = lo_half ? make_float2(at10, at11) : make_float2(ds10, ds11);
            if (v0) *reinterpret_cast<float2 *>(P.coef + (int64_t)id0 * 8 + 2 * q) = w0;
            if (v1) *reinterpret_cast<float2 *>(P.coef + (int64_t)id1 * 8 + 2 * q) = w1;
        }
        const uint32_t ds_lo = pack_bf16(ds00, ds01), ds_hi = pack_bf16(ds10, ds11);

        // ---- feature gradient ----------------------------------------------------------------------------------
        if (P.df_out) {
            const uint32_t a_lo = lo_half ? ds_lo : pack_bf16(at00, at01);   // k 0..3: ds (x QT rows), k 4..7: a~ (x GT rows)
            const uint32_t a_hi = lo_half ? ds_hi : pack_bf16(at10, at11);
            __nv_bfloat16 *o0 = P.df_out + (int64_t)id0 * P.lddf + 8 * q, *o1 = P.df_out + (int64_t)id1 * P.lddf + 8 * q;
            const __nv_bfloat16 *i0 = P.df_in + (int64_t)id0 * P.lddf + 8 * q, *i1 = P.df_in + (int64_t)id1 * P.lddf + 8 * q;
            const uint32_t fa = ftile + g * MM_ROWB + q * 16;
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                float e[4][4];
                const uint32_t xw[4] = {rf.x[cb].x, rf.x[cb].y, rf.x[cb].z, rf.x[cb].w};
#pragma unroll
                for (int wd = 0; wd < 4; ++wd) {
                    e[wd][0] = e[wd][1] = e[wd][2] = e[wd][3] = 0.f;
                    mma_bf16(e[wd], a_lo, a_hi, 0u, 0u, movmatrix_trans(xw[wd]), 0u);
                }
                F8 r0f, r1f;
#pragma unroll
                for (int wd = 0; wd < 4; ++wd) {
                    r0f.v[2 * wd] = e[wd][0]; r0f.v[2 * wd + 1] = e[wd][1];
                    r1f.v[2 * wd] = e[wd][2]; r1f.v[2 * wd + 1] = e[wd][3];
                }
                if (ACCUM) {
                    if (v0) { const F8 t0 = ld8(i0 + cb * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) r0f.v[i] += t0.v[i]; }
                    if (v1) { const F8 t1 = ld8(i1 + cb * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) r1f.v[i] += t1.v[i]; }
                }
                if (P.relu_mask) {
                    const F8 f0 = lds8(reinterpret_cast<const __nv_bfloat16 *>(base + (size_t)s * 3 * MM_TILE + 2 * MM_TILE +
                                                                               (size_t)g * MM_ROWB) + 32 * cb + 8 * q);
                    const F8 f1 = lds8(reinterpret_cast<const __nv_bfloat16 *>(base + (size_t)s * 3 * MM_TILE + 2 * MM_TILE +
                                                                               (size_t)(g + 8) * MM_ROWB) + 32 * cb + 8 * q);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        r0f.v[i] = f0.v[i] > 0.f ? r0f.v[i] : 0.f;
                        r1f.v[i] = f1.v[i] > 0.f ? r1f.v[i] : 0.f;
                    }
                }
                if (v0) st8(o0 + cb * 32, r0f);
                if (v1) st8(o1 + cb * 32, r1f);
            }
            (void)fa;
        }

        const uint32_t tlo = movmatrix_trans(ds_lo), thi = movmatrix_trans(ds_hi);
        const uint32_t bf0 = g < 4 ? tlo : 0u, bf1 = g < 4 ? thi : 0u;
        const uint32_t bk0 = g < 4 ? 0u : tlo, bk1 = g < 4 ? 0u : thi;

        if (nxt.first && !nxt.done()) load_bwd_row(rf, P, nxt.row, g, q);

        // ---- phase 2: bbar (cols 0..3) and dq (cols 4..7) --------------------------------------------------------
        {
            const uint32_t toff = (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * MM_ROWB + ((lane >> 3) & 1) * 16);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t a[4];
                ldsm_x4_trans(a, ftile + toff + j * 32);
                mma_bf16(acc[j], a[0], a[1], a[2], a[3], bf0, bf1);
                ldsm_x4_trans(a, ktile + toff + j * 32);
                mma_bf16(acc[j], a[0], a[1], a[2], a[3], bk0, bk1);
            }
        }
        if (last) {
            __syncwarp();
            const uint32_t st0 = vtile + (uint32_t)((2 * q) * MM_STG + g) * 4, st1 = st0 + MM_STG * 4;   // V slot is dead
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                sts32f(st0 + j * 64, acc[j][0]);
                sts32f(st1 + j * 64, acc[j][1]);
                sts32f(st0 + j * 64 + 32, acc[j][2]);
                sts32f(st1 + j * 64 + 32, acc[j][3]);
            }
            __syncwarp();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int ch = half * 128 + 4 * lane;
#pragma unroll
                for (int t = 0; t < MM_HEADS; ++t) {
                    const float4 x = lds128f(vtile + (uint32_t)(t * MM_STG + ch) * 4);
                    uint2 ov;
                    ov.x = pack_bf16(x.x, x.y);
                    ov.y = pack_bf16(x.z, x.w);
                    *reinterpret_cast<uint2 *>(P.bbar + row * P.ldbb + (int64_t)t * P.hsbb + ch) = ov;
                }
                const int head = ch >> 6;
                const float4 y = lds128f(vtile + (uint32_t)((4 + head) * MM_STG + ch) * 4);
                uint2 ov;
                ov.x = pack_bf16(y.x, y.y);
                ov.y = pack_bf16(y.z, y.w);
                *reinterpret_cast<uint2 *>(P.dq + row * P.lddq + ch) = ov;
            }
            next_unwritten = row + 1;
        }
        __syncwarp();
        if (MM_STAGES == 1) issue(0);
        cons = nxt;
    }
    zero_rows(next_unwritten, r1);
}

static int mm_grid(int64_t n_nodes, int64_t n_edges) {
    // one CTA per SM is resident (shared memory): one wave of cost-balanced row ranges (more waves only repeat the
    // per-CTA set-up: 78.9 vs 101.4 us for the backward at 8 192 rows)
    const int64_t work = n_edges + (int64_t)ROW_KAPPA * n_nodes;
    int64_t blocks = 148;
    if (const char *e = getenv("ALIGNN_MM_BLOCKS")) {            // tuning knob
        const int want = atoi(e);
        if (want >= 1) blocks = want;
    }
    const int64_t min_work_per_warp = 64;
    if (work / (blocks * MM_WARPS) < min_work_per_warp) blocks = work / (min_work_per_warp * MM_WARPS) + 1;
    return (int)blocks;
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_edgeattn_mma_supported(int hidden, int heads, int dtype) {
    return hidden == MM_HID && heads == MM_HEADS && dtype == ALIGNN_BF16;
}

extern "C" int alignn_edgeattn_mma_fwd_s(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk,
                                         int64_t ldv, const void *qt, int64_t ldqt, int64_t hsqt, const void *feat,
                                         const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                                         float *aggv, void *abar, int64_t ldab, int64_t hsab,
                                         float *stat_m, float *stat_z, float *stat_s,
                                         int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                         float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step,
                                         void *stream) {
    if (!alignn_edgeattn_mma_supported(hidden, heads, dtype)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes < 0 || n_edges < 0 || !(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 64) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!q || !k || !v || !qt || !rowptr || !aggv || !abar || !stat_m || !stat_z || !stat_s) return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!feat || !col || !eid)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(qt) || !aligned16(feat) || !aligned16(aggv) ||
        !aligned16(abar) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldqt % 8) || (hsqt % 8) || (ldab % 8) || (hsab % 8))
        return ALIGNN_ERR_BAD_ARG;
    MmFwdParams p;
    p.q = (const __nv_bfloat16 *)q; p.k = (const __nv_bfloat16 *)k; p.v = (const __nv_bfloat16 *)v;
    p.qt = (const __nv_bfloat16 *)qt; p.feat = (const __nv_bfloat16 *)feat;
    p.rowptr = rowptr; p.col = col; p.eid = eid;
    p.aggv = aggv; p.abar = (__nv_bfloat16 *)abar; p.stat_m = stat_m; p.stat_z = stat_z; p.stat_s = stat_s;
    p.n_nodes = n_nodes; p.n_edges = n_edges; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
    p.ldqt = ldqt; p.hsqt = hsqt; p.ldab = ldab; p.hsab = hsab; p.rng_step = rng_step;
    p.scale_log2 = LOG2E / sqrtf((float)(hidden / heads));
    p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    p.seed = seed; p.offset = offset;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int SMEM = MM_WARPS * (MM_STAGES * 3 * MM_TILE + 64);
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(edgeattn_mma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    edgeattn_mma_fwd_kernel<<<mm_grid(n_nodes, n_edges), MM_WARPS * 32, SMEM, st>>>(p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int alignn_edgeattn_mma_bwd_dst_s(const float *dagg, const void *dagg_lp, const float *agg,
                                             const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk,
                                             int64_t ldv, const void *qt, int64_t ldqt, int64_t hsqt,
                                             const void *gt, int64_t ldgt, int64_t hsgt, const float *cvec,
                                             const void *feat, const float *stat_m, const float *stat_z,
                                             const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                                             void *dq, int64_t lddq, void *bbar, int64_t ldbb, int64_t hsbb, float *coef,
                                             const void *df_in, void *df_out, int64_t lddf, int relu_mask,
                                             int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                             float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step,
                                             void *stream) {
    if (!alignn_edgeattn_mma_supported(hidden, heads, dtype)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes < 0 || n_edges < 0 || !(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 64) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!dagg || !dagg_lp || !agg || !q || !k || !v || !qt || !gt || !stat_m || !stat_z || !rowptr || !dq || !bbar)
        return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!feat || !col || !eid || !coef)) return ALIGNN_ERR_BAD_ARG;
    if (df_in && !df_out) return ALIGNN_ERR_BAD_ARG;
    if (df_out && (lddf < MM_HID || (lddf % 8))) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(dagg) || !aligned16(dagg_lp) || !aligned16(agg) || !aligned16(q) || !aligned16(k) || !aligned16(v) ||
        !aligned16(qt) || !aligned16(gt) || !aligned16(cvec) || !aligned16(feat) || !aligned16(dq) || !aligned16(bbar) ||
        !aligned16(coef) || !aligned16(df_in) || !aligned16(df_out) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (lddq % 8) ||
        (ldqt % 8) || (hsqt % 8) || (ldgt % 8) || (hsgt % 8) || (ldbb % 8) || (hsbb % 8))
        return ALIGNN_ERR_BAD_ARG;
    MmBwdParams p;
    p.dagg = dagg; p.agg = agg; p.dagg_lp = (const __nv_bfloat16 *)dagg_lp;
    p.q = (const __nv_bfloat16 *)q; p.k = (const __nv_bfloat16 *)k; p.v = (const __nv_bfloat16 *)v;
    p.qt = (const __nv_bfloat16 *)qt; p.gt = (const __nv_bfloat16 *)gt; p.cvec = cvec;
    p.feat = (const __nv_bfloat16 *)feat; p.stat_m = stat_m; p.stat_z = stat_z;
    p.rowptr = rowptr; p.col = col; p.eid = eid;
    p.dq = (__nv_bfloat16 *)dq; p.bbar = (__nv_bfloat16 *)bbar; p.coef = coef;
    p.df_in = (const __nv_bfloat16 *)df_in; p.df_out = (__nv_bfloat16 *)df_out; p.lddf = lddf;
    p.n_nodes = n_nodes; p.n_edges = n_edges; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.lddq = lddq;
    p.ldqt = ldqt; p.hsqt = hsqt; p.ldgt = ldgt; p.hsgt = hsgt; p.ldbb = ldbb; p.hsbb = hsbb; p.rng_step = rng_step;
    p.scale = 1.0f / sqrtf((float)(hidden / heads));
    p.scale_log2 = p.scale * LOG2E;
    p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    p.seed = seed; p.offset = offset; p.relu_mask = relu_mask;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int SMEM = MM_WARPS * (MM_STAGES * 3 * MM_TILE + 64 + 128);
    const int grid = mm_grid(n_nodes, n_edges);
    if (df_in) {
        ALIGNN_CUDA_TRY(cudaFuncSetAttribute(edgeattn_mma_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        edgeattn_mma_bwd_kernel<true><<<grid, MM_WARPS * 32, SMEM, st>>>(p);
    } else {
        ALIGNN_CUDA_TRY(cudaFuncSetAttribute(edgeattn_mma_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        edgeattn_mma_bwd_kernel<false><<<grid, MM_WARPS * 32, SMEM, st>>>(p);
    }
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

// contiguous [heads, Nn, 256] layouts of qt / gt / abar / bbar
extern "C" int alignn_edgeattn_mma_fwd(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk,
                                       int64_t ldv, const void *qt, const void *feat,
                                       const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                                       float *aggv, void *abar, float *stat_m, float *stat_z, float *stat_s,
                                       int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                       float p_drop, uint64_t seed, uint64_t offset, void *stream) {
    return alignn_edgeattn_mma_fwd_s(q, k, v, ldq, ldk, ldv, qt, MM_HID, n_nodes * MM_HID, feat, rowptr, col, eid, aggv,
                                     abar, MM_HID, n_nodes * MM_HID, stat_m, stat_z, stat_s, n_nodes, n_edges, hidden,
                                     heads, dtype, p_drop, seed, offset, nullptr, stream);
}

extern "C" int alignn_edgeattn_mma_bwd_dst(const float *dagg, const void *dagg_lp, const float *agg,
                                           const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk,
                                           int64_t ldv, const void *qt, const void *gt, const float *cvec,
                                           const void *feat, const float *stat_m, const float *stat_z,
                                           const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                                           void *dq, int64_t lddq, void *bbar, float *coef,
                                           const void *df_in, void *df_out, int relu_mask,
                                           int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                           float p_drop, uint64_t seed, uint64_t offset, void *stream) {
    return alignn_edgeattn_mma_bwd_dst_s(dagg, dagg_lp, agg, q, k, v, ldq, ldk, ldv, qt, MM_HID, n_nodes * MM_HID, gt,
                                         MM_HID, n_nodes * MM_HID, cvec, feat, stat_m, stat_z, rowptr, col, eid, dq, lddq,
                                         bbar, MM_HID, n_nodes * MM_HID, coef, df_in, df_out, MM_HID, relu_mask, n_nodes, n_edges,
                                         hidden, heads, dtype, p_drop, seed, offset, nullptr, stream);
}
