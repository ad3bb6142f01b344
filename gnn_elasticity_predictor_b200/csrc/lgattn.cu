// Line-graph edge attention with the angle embedding RECOMPUTED in the kernel (hidden = 256, heads = 4, bf16).
//
// Reference path covered: `angle_emb = angle_encoder(lg_edge_attr)` (scripts/train.py:360-364, 554) feeding
// `EdgeUpdateBlock.conv(edge_state, lg_edge_index, angle_emb)` (train.py:308, 315) = PyG TransformerConv on the line
// graph, forward and backward.  As in edgeattn.cu the second encoder Linear and `lin_edge` are folded into per-NODE
// operands (qt, gt); what is left per angle is h1 = relu(W1 a + b1), a function of 11 numbers.  Storing h1 costs
// 512 B per angle and every layer reads it twice (forward + backward) and accumulates a 512 B gradient row; these
// kernels read the 32-byte padded feature row instead and rebuild the h1 tile with 64 tiny MMAs per 16 angles, in
// both operand orientations the two phases need.  Nothing of size [L, 256] exists in the line-graph path any more.
//
//   a_csr  : [L, 16] bf16, angle features in TARGET-SORTED (CSR) order, column in_dim = 1 (bias), rest 0
//   W1ext  : [256, 16] = (W1 | b1 | 0) in bf16 -- staged once per CTA as ready-made MMA fragments
//
// Pipeline per warp (8 warps / SM): K tile double-buffered, V tile single-buffered (its refill is issued right after
// phase 2 and lands during the next chunk's phase 1), cp.async 16 B per lane.  Attention dropout is keyed by the
// angle's CSR position, so forward and backward agree without an edge-id lookup.
#include <math.h>
#include <stdlib.h>

#include "mma.cuh"

namespace alignn {

constexpr int LG_HID = 256;
constexpr int LG_HEADS = 4;
constexpr int LG_E = 16;
constexpr int LG_ROWB = 528;
constexpr int LG_TILE = LG_E * LG_ROWB;      // 8448
constexpr int LG_STG = 260;
constexpr int LG_IMG1 = 32 * 32 * 8;         // phase-1 W1 fragments (B operand, permuted channel order)
constexpr int LG_IMG2 = 16 * 32 * 16;        // phase-2 W1 fragments (A operand, natural channel order)
constexpr int LG_IMG = LG_IMG1 + LG_IMG2;    // 16 KB per CTA

__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint2 lds64(uint32_t smem_addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_addr));
    return v;
}

constexpr int LG_UNIT = 128;   // scheduling unit: a row range of about this many (edges + ROW_KAPPA * rows)

// work[0]: next unit, work[1]: CTAs that ran out of units.  The last CTA to finish puts both back to zero, so the same
// two words serve every launch on the stream without a memset in between (the buffer must be zero before the first one).
__device__ __forceinline__ void lg_work_done(unsigned int *work) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int d = atomicAdd(work + 1, 1u);
        if (d == gridDim.x - 1) {
            work[0] = 0u;
            work[1] = 0u;
            __threadfence();
        }
    }
}

// W1ext(ch, r): r < in_dim -> W1[ch, r];  r == in_dim -> b1[ch];  else 0
__device__ __forceinline__ float w1ext(const float *__restrict__ w1, const float *__restrict__ b1, int in_dim, int ch, int r) {
    return r < in_dim ? __ldg(w1 + ch * in_dim + r) : (r == in_dim ? __ldg(b1 + ch) : 0.f);
}

// all threads of the CTA; caller syncs
__device__ __forceinline__ void build_w1_images(unsigned char *img, const float *w1, const float *b1, int in_dim) {
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {     // image 1: tile = cb*4+u, lane (g, q)
        const int tile = i >> 5, l = i & 31, g = l >> 2, q = l & 3;
        const int cb = tile >> 2, u = tile & 3;
        const int ch = 32 * cb + 8 * (g >> 1) + 2 * u + (g & 1);
        uint2 w;
        w.x = pack_bf16(w1ext(w1, b1, in_dim, ch, 2 * q), w1ext(w1, b1, in_dim, ch, 2 * q + 1));
        w.y = pack_bf16(w1ext(w1, b1, in_dim, ch, 2 * q + 8), w1ext(w1, b1, in_dim, ch, 2 * q + 9));
        *reinterpret_cast<uint2 *>(img + (size_t)i * 8) = w;
    }
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {     // image 2: m-tile j, lane (g, q)
        const int j = i >> 5, l = i & 31, g = l >> 2, q = l & 3;
        uint4 w;
        w.x = pack_bf16(w1ext(w1, b1, in_dim, 16 * j + g, 2 * q), w1ext(w1, b1, in_dim, 16 * j + g, 2 * q + 1));
        w.y = pack_bf16(w1ext(w1, b1, in_dim, 16 * j + g + 8, 2 * q), w1ext(w1, b1, in_dim, 16 * j + g + 8, 2 * q + 1));
        w.z = pack_bf16(w1ext(w1, b1, in_dim, 16 * j + g, 2 * q + 8), w1ext(w1, b1, in_dim, 16 * j + g, 2 * q + 9));
        w.w = pack_bf16(w1ext(w1, b1, in_dim, 16 * j + g + 8, 2 * q + 8), w1ext(w1, b1, in_dim, 16 * j + g + 8, 2 * q + 9));
        *reinterpret_cast<uint4 *>(img + LG_IMG1 + (size_t)i * 16) = w;
    }
}

// One warp's walk over its target rows in chunks of <= 16 edges that never cross a row.
struct Chunk {
    int row, pos, n;
    bool first, last;
};
struct ChunkCursor {
    int row, row_hi, pos, end;
    bool first;
    __device__ __forceinline__ void init(const int32_t *__restrict__ rowptr, int r0, int r1) {
        row = r0; row_hi = r1;
        pos = r0 < r1 ? __ldg(rowptr + r0) : 0;
        end = pos; first = true;
        seek(rowptr);
    }
    __device__ __forceinline__ void seek(const int32_t *__restrict__ rowptr) {
        while (row < row_hi) {
            end = __ldg(rowptr + row + 1);
            if (end > pos) break;
            ++row;
        }
    }
    __device__ __forceinline__ bool done() const { return row >= row_hi; }
    __device__ __forceinline__ Chunk take(const int32_t *__restrict__ rowptr) {
        Chunk c;
        c.row = row; c.pos = pos; c.n = min(LG_E, end - pos); c.first = first; c.last = end - pos <= LG_E;
        pos += c.n;
        if (pos == end) { ++row; first = true; seek(rowptr); } else first = false;
        return c;
    }
};

// gather 16 rows of 512 B (row u from src + idx_u * ld) into a padded tile; rows >= n are left alone.
// One warp instruction copies a 128-byte piece of FOUR rows (lane = 8 * row-in-group + 16-byte slot): the row address is
// formed once per group of four rows (one shuffle, one 64-bit multiply-add, one predicate) and the four pieces of a row are
// immediate offsets -- 16 cp.async + 4 address set-ups per tile instead of 16 x (shuffle + multiply-add + predicate + copy).
// Every request still covers whole 128-byte lines (same L2 request count as a row per instruction).
__device__ __forceinline__ void gather_rows(uint32_t tile, const __nv_bfloat16 *__restrict__ src, int ld, int jmine, int n,
                                            int lane) {
    const uint32_t stride = (uint32_t)ld * 2u;
    const int sub = lane >> 3;                                   // row within the group of four
    const char *base = reinterpret_cast<const char *>(src) + (lane & 7) * 16;
    const uint32_t dst0 = tile + (uint32_t)sub * LG_ROWB + (lane & 7) * 16;
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
        const int u = 4 * grp + sub;
        const uint32_t ju = (uint32_t)__shfl_sync(FULL, jmine, u);     // ids are non-negative
        if (u < n) {
            const char *p = base + (uint64_t)ju * stride;
            const uint32_t d = dst0 + (uint32_t)(4 * grp) * LG_ROWB;
            cp_async16(d, p);
            cp_async16(d + 128, p + 128);
            cp_async16(d + 256, p + 256);
            cp_async16(d + 384, p + 384);
        }
    }
}

struct LgFwdParams {
    const __nv_bfloat16 *q, *k, *v;     // node rows, strides ldq/ldk/ldv (elements)
    const __nv_bfloat16 *qt;            // qt[row * ldqt + t * hsqt + ch]
    const __nv_bfloat16 *a_csr;         // [Ne, 16]
    const float *w1, *b1;               // fp32 [256, in_dim], [256]
    const int32_t *rowptr, *col;
    float *aggv;                        // [Nn, 256]
    __nv_bfloat16 *abar;                // abar[row * ldab + t * hsab + ch]
    float *stat_m, *stat_z, *stat_s;    // [Nn, 4]
    const uint64_t *rng_step;           // optional device counter added to `offset` (CUDA-graph replays)
    unsigned int *work;                 // [2] zero on entry, zero again on exit (dynamic unit scheduler)
    int n_nodes, n_edges, in_dim;
    int ldq, ldk, ldv;
    int64_t ldqt, hsqt, ldab, hsab;
    float scale_log2, p_drop, inv_keep;
    uint64_t seed, offset;
};

constexpr int LGF_WARPS = 8;
constexpr int LGF_PER_WARP = 3 * LG_TILE;    // K0, K1, V

__global__ void __launch_bounds__(LGF_WARPS * 32, 1)
lgattn_fwd_kernel(const LgFwdParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    build_w1_images(smem_raw, P.w1, P.b1, P.in_dim);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const uint32_t img1 = smem_u32(smem_raw), img2 = img1 + LG_IMG1;
    const uint32_t wbase_u32 = img1 + LG_IMG + (uint32_t)warp * LGF_PER_WARP;
    for (int off = lane * 16; off < LGF_PER_WARP; off += 512) sts128(wbase_u32 + off, make_uint4(0, 0, 0, 0));
    __syncthreads();

    const int64_t total = (int64_t)P.n_edges + (int64_t)ROW_KAPPA * P.n_nodes;
    const int64_t n_units = (total + LG_UNIT - 1) / LG_UNIT;
    const uint64_t rng_off = P.offset + (P.rng_step ? *P.rng_step : 0ull);
    // Dynamic scheduling: every warp draws the next unit (a row range of ~LG_UNIT edges) from a global counter, so that
    // rows of very different length (0 .. thousands of in-edges with PyG-collated batches) balance across the 1 184 warps
    // of the grid.  A row is always processed by ONE warp in CSR order: the result does not depend on the schedule.
    for (;;) {
    int unit = 0;
    if (lane == 0) unit = (int)atomicAdd(P.work, 1u);
    unit = __shfl_sync(FULL, unit, 0);
    if (unit >= n_units) break;
    const int r0 = (int)row_bound(P.rowptr, P.n_nodes, (int64_t)unit * LG_UNIT);
    const int r1 = (int)row_bound(P.rowptr, P.n_nodes, min(total, (int64_t)(unit + 1) * LG_UNIT));
    if (r0 >= r1) continue;
    const int e_end = __ldg(P.rowptr + r1);

    ChunkCursor cur;
    cur.init(P.rowptr, r0, r1);
    if (cur.done()) {   // only rows without in-edges
        F8 zf;
#pragma unroll
        for (int c = 0; c < 8; ++c) zf.v[c] = 0.f;
        for (int r = r0; r < r1; ++r) {
            st8(P.aggv + (int64_t)r * LG_HID + lane * 8, zf);
#pragma unroll
            for (int t = 0; t < LG_HEADS; ++t) st8(P.abar + (int64_t)r * P.ldab + (int64_t)t * P.hsab + lane * 8, zf);
            if (lane < LG_HEADS) {
                P.stat_m[(int64_t)r * LG_HEADS + lane] = 0.f;
                P.stat_z[(int64_t)r * LG_HEADS + lane] = 0.f;
                P.stat_s[(int64_t)r * LG_HEADS + lane] = 0.f;
            }
        }
        continue;
    }
    int wbase = cur.pos;
    IndexWindow wcol;
    wcol.init(P.col, wbase, e_end, lane);
    const uint32_t *ap = reinterpret_cast<const uint32_t *>(P.a_csr);

    auto fetch = [&](Chunk &c, int &jm, uint32_t (&af)[4]) {   // next chunk: descriptor, source ids, angle fragments
        c = cur.take(P.rowptr);
        jm = wcol.get(c.pos + min(lane & 15, c.n - 1) - wbase);
        const int e0 = c.pos + g, e1 = c.pos + g + 8;
        af[0] = e0 < P.n_edges ? __ldg(ap + (int64_t)e0 * 8 + q) : 0u;
        af[1] = e1 < P.n_edges ? __ldg(ap + (int64_t)e1 * 8 + q) : 0u;
        af[2] = e0 < P.n_edges ? __ldg(ap + (int64_t)e0 * 8 + 4 + q) : 0u;
        af[3] = e1 < P.n_edges ? __ldg(ap + (int64_t)e1 * 8 + 4 + q) : 0u;
        if (cur.pos - wbase >= 32 && !cur.done()) {
            wbase += 32;
            wcol.shift(P.col, wbase, e_end, lane);
        }
    };
    auto load_row = [&](uint4 (&qtf)[8], uint32_t (&qvf)[8], int row) {
        const int t = g & 3;
        const uint4 *pt = reinterpret_cast<const uint4 *>(P.qt + (int64_t)row * P.ldqt + (int64_t)t * P.hsqt) + q;
#pragma unroll
        for (int c = 0; c < 8; ++c) qtf[c] = __ldg(pt + 4 * c);
        // own-head q in NATURAL channel order (k-step i: channels 16i + 2q, +1 and 16i + 8 + 2q, +1): pairs with K
        // A fragments read by ldmatrix (conflict-free; per-lane 128-bit reads of the 528-byte-stride tile are not)
        const uint32_t *pq = reinterpret_cast<const uint32_t *>(P.q + (int64_t)row * P.ldq + 64 * t) + q;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            qvf[2 * i] = __ldg(pq + 8 * i);
            qvf[2 * i + 1] = __ldg(pq + 8 * i + 4);
        }
    };
    auto zero_rows = [&](int lo, int hi) {
        F8 zf;
#pragma unroll
        for (int c = 0; c < 8; ++c) zf.v[c] = 0.f;
        for (int r = lo; r < hi; ++r) {
            st8(P.aggv + (int64_t)r * LG_HID + lane * 8, zf);
#pragma unroll
            for (int t = 0; t < LG_HEADS; ++t) st8(P.abar + (int64_t)r * P.ldab + (int64_t)t * P.hsab + lane * 8, zf);
            if (lane < LG_HEADS) {
                P.stat_m[(int64_t)r * LG_HEADS + lane] = 0.f;
                P.stat_z[(int64_t)r * LG_HEADS + lane] = 0.f;
                P.stat_s[(int64_t)r * LG_HEADS + lane] = 0.f;
            }
        }
    };

    const uint32_t vtile = wbase_u32 + 2 * LG_TILE;
    Chunk A, B;
    int jA, jB = 0;
    uint32_t afA[4], afB[4] = {0u, 0u, 0u, 0u};
    fetch(A, jA, afA);
    gather_rows(wbase_u32, P.k, P.ldk, jA, A.n, lane);
    cp_async_commit();
    gather_rows(vtile, P.v, P.ldv, jA, A.n, lane);
    cp_async_commit();

    uint4 qtf[8];
    uint32_t qvf[8];
    load_row(qtf, qvf, A.row);
    float acc[16][4];
    float m0 = -INFINITY, m1 = -INFINITY, z0 = 0.f, z1 = 0.f, zd0 = 0.f, zd1 = 0.f;
    int next_unwritten = r0;
    const int t_own = g & 3;
    B.n = 0; B.row = 0; B.pos = 0; B.first = false; B.last = false;

    for (int it = 0;; ++it) {
        const int s = it & 1;
        const bool have_next = !cur.done();
        if (have_next) {
            fetch(B, jB, afB);
            gather_rows(wbase_u32 + (uint32_t)(s ^ 1) * LG_TILE, P.k, P.ldk, jB, B.n, lane);
        }
        cp_async_commit();
        const uint32_t ktile = wbase_u32 + (uint32_t)s * LG_TILE;
        if (A.first) {
            zero_rows(next_unwritten, A.row);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            m0 = m1 = -INFINITY;
            z0 = z1 = zd0 = zd1 = 0.f;
        }
        // dropout keep-scales depend on (seed, position) only: generated here, off the softmax -> phase-2 critical path
        float k00 = 1.f, k01 = 1.f, k10 = 1.f, k11 = 1.f;
        if (P.p_drop > 0.f)
            dropout_scale_quad(P.seed, rng_off, (uint64_t)A.pos, g, q, P.p_drop, P.inv_keep, k00, k01, k10, k11);
        cp_async_wait<2>();   // K of this chunk (V of this chunk and K of the next may still be in flight)
        __syncwarp();

        // ---- phase 1: logits = relu(a W1^T) . QT  +  K . Qbd ------------------------------------------------
        float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb_[4] = {0.f, 0.f, 0.f, 0.f};
        float ck0[4] = {0.f, 0.f, 0.f, 0.f}, ck1[4] = {0.f, 0.f, 0.f, 0.f};
        {
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                float h[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint2 wf = lds64(img1 + (uint32_t)(((cb * 4 + u) * 32 + lane) * 8));
                    h[u][0] = h[u][1] = h[u][2] = h[u][3] = 0.f;
                    mma_bf16(h[u], afA[0], afA[1], afA[2], afA[3], wf.x, wf.y);
                }
                mma_bf16(ca, pack_relu_bf16(h[0][0], h[0][1]), pack_relu_bf16(h[0][2], h[0][3]),
                         pack_relu_bf16(h[1][0], h[1][1]), pack_relu_bf16(h[1][2], h[1][3]), qtf[cb].x, qtf[cb].y);
                mma_bf16(cb_, pack_relu_bf16(h[2][0], h[2][1]), pack_relu_bf16(h[2][2], h[2][3]),
                         pack_relu_bf16(h[3][0], h[3][1]), pack_relu_bf16(h[3][2], h[3][3]), qtf[cb].z, qtf[cb].w);
            }
            const uint32_t aoff = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * LG_ROWB + (lane >> 4) * 16);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                uint32_t a[4];
                ldsm_x4(a, ktile + aoff + kk * 32);
                const bool own = (kk >> 2) == t_own;
                mma_bf16((kk & 1) ? ck1 : ck0, a[0], a[1], a[2], a[3], own ? qvf[2 * (kk & 3)] : 0u, own ? qvf[2 * (kk & 3) + 1] : 0u);
            }
        }
        const float c0 = (ca[0] + cb_[0]) + (ck0[0] + ck1[0]), c1 = (ca[1] + cb_[1]) + (ck0[1] + ck1[1]);
        const float c2 = (ca[2] + cb_[2]) + (ck0[2] + ck1[2]), c3 = (ca[3] + cb_[3]) + (ck0[3] + ck1[3]);
        // ---- online softmax --------------------------------------------------------------------------------------
        const int n = A.n;
        const bool v0 = g < n, v1 = g + 8 < n;
        const float s00 = v0 ? c0 * P.scale_log2 : -INFINITY, s01 = v0 ? c1 * P.scale_log2 : -INFINITY;
        const float s10 = v1 ? c2 * P.scale_log2 : -INFINITY, s11 = v1 ? c3 * P.scale_log2 : -INFINITY;
        const float mn0 = fmaxf(m0, colmax8(fmaxf(s00, s10))), mn1 = fmaxf(m1, colmax8(fmaxf(s01, s11)));
        const float corr0 = fast_exp2(m0 - mn0), corr1 = fast_exp2(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        float p00 = fast_exp2(s00 - mn0), p01 = fast_exp2(s01 - mn1);
        float p10 = fast_exp2(s10 - mn0), p11 = fast_exp2(s11 - mn1);
        z0 = z0 * corr0 + (p00 + p10);
        z1 = z1 * corr1 + (p01 + p11);
        p00 *= k00;
        p01 *= k01;
        p10 *= k10;
        p11 *= k11;
        zd0 = zd0 * corr0 + (p00 + p10);
        zd1 = zd1 * corr1 + (p01 + p11);
        // the accumulators only need rescaling when a running maximum moved: rare after a row's first chunks (x * 1.0f is
        // exact, so skipping it changes no bit)
        if (!A.first && __any_sync(FULL, corr0 != 1.0f || corr1 != 1.0f)) {
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

        if (have_next && B.first) load_row(qtf, qvf, B.row);   // in flight during phase 2

        cp_async_wait<1>();   // V of this chunk
        __syncwarp();
        // ---- phase 2: acc += relu(W1 a^T) . P|0..3  +  V^T . P|4..7 ---------------------------------------------
        {
            const uint32_t toff = (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * LG_ROWB + ((lane >> 3) & 1) * 16);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint4 wf = lds128(img2 + (uint32_t)((j * 32 + lane) * 16));
                float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
                mma_bf16(t0, wf.x, wf.y, wf.z, wf.w, afA[0], afA[2]);
                mma_bf16(t1, wf.x, wf.y, wf.z, wf.w, afA[1], afA[3]);
                uint32_t a[4];
                ldsm_x4_trans(a, vtile + toff + j * 32);
                mma_bf16(acc[j], pack_relu_bf16(t0[0], t0[1]), pack_relu_bf16(t0[2], t0[3]), pack_relu_bf16(t1[0], t1[1]),
                         pack_relu_bf16(t1[2], t1[3]), bf0, bf1);
                mma_bf16(acc[j], a[0], a[1], a[2], a[3], bv0, bv1);
            }
        }
        __syncwarp();   // all lanes are done with V before it is refilled
        if (have_next) gather_rows(vtile, P.v, P.ldv, jB, B.n, lane);
        cp_async_commit();

        // ---- row epilogue ------------------------------------------------------------------------------------------
        if (A.last) {
            const int row = A.row;
            const float zs0 = colsum8(z0), zs1 = colsum8(z1), zds0 = colsum8(zd0), zds1 = colsum8(zd1);
            const float inv0 = 1.0f / (zs0 + 1e-16f), inv1 = 1.0f / (zs1 + 1e-16f);
            if (g == 0 && q < 2) {
                float *sm = P.stat_m + (int64_t)row * LG_HEADS + 2 * q, *sz = P.stat_z + (int64_t)row * LG_HEADS + 2 * q;
                float *ss = P.stat_s + (int64_t)row * LG_HEADS + 2 * q;
                sm[0] = m0; sm[1] = m1; sz[0] = zs0; sz[1] = zs1; ss[0] = zds0 * inv0; ss[1] = zds1 * inv1;
            }
            // transpose through the K slot of this stage (dead after phase 1; refilled at the top of iteration it+1)
            const uint32_t st0 = ktile + (uint32_t)((2 * q) * LG_STG + g) * 4, st1 = st0 + LG_STG * 4;
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
                for (int t = 0; t < LG_HEADS; ++t) {
                    const float4 x = lds128f(ktile + (uint32_t)(t * LG_STG + ch) * 4);
                    uint2 o;
                    o.x = pack_bf16(x.x, x.y);
                    o.y = pack_bf16(x.z, x.w);
                    *reinterpret_cast<uint2 *>(P.abar + (int64_t)row * P.ldab + (int64_t)t * P.hsab + ch) = o;
                }
                const float4 y = lds128f(ktile + (uint32_t)((4 + (ch >> 6)) * LG_STG + ch) * 4);
                *reinterpret_cast<float4 *>(P.aggv + (int64_t)row * LG_HID + ch) = y;
            }
            next_unwritten = row + 1;
            __syncwarp();
        }
        if (!have_next) break;
        A = B;
        jA = jB;
#pragma unroll
        for (int i = 0; i < 4; ++i) afA[i] = afB[i];
    }
    cp_async_wait<0>();
    zero_rows(next_unwritten, r1);
    __syncwarp();
    }   // unit loop
    lg_work_done(P.work);
}

// ------------------------------------------------------------------------------------------------------------
// backward, target-sorted pass (no feature gradient: the angle-encoder gradient is formed by lg_angle_grad_kernel from
// the per-edge coefficients of all layers)
//   phase 1   SD[16 x 8] = relu(a W1^T) . [QT_i | GT_i] + K . [Qbd_i | 0] + V . [0 | Gbd_i]
//   coef      (a~, ds) per (edge, head) -> coef[p] in CSR order
//   phase 2   acc[256 x 8] += relu(W1 a^T) . DS|0..3 + K^T . DS|4..7   ->  bbar_i (cols 0..3), dq_i (cols 4..7)
// ------------------------------------------------------------------------------------------------------------
struct LgBwdParams {
    const float *dagg, *agg;
    const __nv_bfloat16 *dagg_lp;
    const __nv_bfloat16 *q, *k, *v;
    const __nv_bfloat16 *qt, *gt;         // base[row * ld + t * hs + ch]
    const float *cvec;
    const __nv_bfloat16 *a_csr;
    const float *w1, *b1;
    const float *stat_m, *stat_z;
    const int32_t *rowptr, *col;
    __nv_bfloat16 *dq;                    // strided rows
    __nv_bfloat16 *bbar;                  // base[row * ldbb + t * hsbb + ch]
    float *coef;                          // [Ne, 8] in CSR order: (a~_0..3, ds_0..3)
    const uint64_t *rng_step;
    unsigned int *work;                   // [2] zero on entry / exit
    int n_nodes, n_edges, in_dim;
    int ldq, ldk, ldv, lddq;
    int64_t ldqt, hsqt, ldgt, hsgt, ldbb, hsbb;
    float scale, scale_log2, p_drop, inv_keep;
    uint64_t seed, offset;
};

constexpr int LGB_WARPS = 8;                 // two per scheduler (6 left two schedulers with a single warp)
constexpr int LGB_PER_WARP = 3 * LG_TILE;    // K0, K1, V: K is read by both phases (double-buffered), V by phase 1 only --
                                             // the next chunk's V rows are gathered behind phase 2 into the same tile

__global__ void __launch_bounds__(LGB_WARPS * 32, 1)
lgattn_bwd_kernel(const LgBwdParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    build_w1_images(smem_raw, P.w1, P.b1, P.in_dim);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const uint32_t img1 = smem_u32(smem_raw), img2 = img1 + LG_IMG1;
    const uint32_t wbase_u32 = img1 + LG_IMG + (uint32_t)warp * LGB_PER_WARP;
    for (int off = lane * 16; off < LGB_PER_WARP; off += 512) sts128(wbase_u32 + off, make_uint4(0, 0, 0, 0));
    __syncthreads();

    const int64_t total = (int64_t)P.n_edges + (int64_t)ROW_KAPPA * P.n_nodes;
    const int64_t n_units = (total + LG_UNIT - 1) / LG_UNIT;
    const uint64_t rng_off = P.offset + (P.rng_step ? *P.rng_step : 0ull);
    for (;;) {   // dynamic units, see lgattn_fwd_kernel
    int unit = 0;
    if (lane == 0) unit = (int)atomicAdd(P.work, 1u);
    unit = __shfl_sync(FULL, unit, 0);
    if (unit >= n_units) break;
    const int r0 = (int)row_bound(P.rowptr, P.n_nodes, (int64_t)unit * LG_UNIT);
    const int r1 = (int)row_bound(P.rowptr, P.n_nodes, min(total, (int64_t)(unit + 1) * LG_UNIT));
    if (r0 >= r1) continue;
    const int e_end = __ldg(P.rowptr + r1);

    auto zero_rows = [&](int lo, int hi) {
        F8 zf;
#pragma unroll
        for (int c = 0; c < 8; ++c) zf.v[c] = 0.f;
        for (int r = lo; r < hi; ++r) {
            st8(P.dq + (int64_t)r * P.lddq + lane * 8, zf);
#pragma unroll
            for (int t = 0; t < LG_HEADS; ++t) st8(P.bbar + (int64_t)r * P.ldbb + (int64_t)t * P.hsbb + lane * 8, zf);
        }
    };

    ChunkCursor cur;
    cur.init(P.rowptr, r0, r1);
    if (cur.done()) {
        zero_rows(r0, r1);
        continue;
    }
    int wbase = cur.pos;
    IndexWindow wcol;
    wcol.init(P.col, wbase, e_end, lane);
    const uint32_t *ap = reinterpret_cast<const uint32_t *>(P.a_csr);

    auto fetch = [&](Chunk &c, int &jm, uint32_t (&af)[4]) {
        c = cur.take(P.rowptr);
        jm = wcol.get(c.pos + min(lane & 15, c.n - 1) - wbase);
        const int e0 = c.pos + g, e1 = c.pos + g + 8;
        af[0] = e0 < P.n_edges ? __ldg(ap + (int64_t)e0 * 8 + q) : 0u;
        af[1] = e1 < P.n_edges ? __ldg(ap + (int64_t)e1 * 8 + q) : 0u;
        af[2] = e0 < P.n_edges ? __ldg(ap + (int64_t)e0 * 8 + 4 + q) : 0u;
        af[3] = e1 < P.n_edges ? __ldg(ap + (int64_t)e1 * 8 + 4 + q) : 0u;
        if (cur.pos - wbase >= 32 && !cur.done()) {
            wbase += 32;
            wcol.shift(P.col, wbase, e_end, lane);
        }
    };
    // lane (g, q): g < 4 -> qt[g] / q (head g);  g >= 4 -> gt[g-4] / dagg (head g-4)
    auto load_row = [&](uint4 (&xf)[8], uint32_t (&kvf)[8], int row) {
        const int t = g & 3;
        const __nv_bfloat16 *wide = g < 4 ? P.qt + (int64_t)row * P.ldqt + (int64_t)t * P.hsqt
                                          : P.gt + (int64_t)row * P.ldgt + (int64_t)t * P.hsgt;
        const uint4 *pt = reinterpret_cast<const uint4 *>(wide) + q;
#pragma unroll
        for (int c = 0; c < 8; ++c) xf[c] = __ldg(pt + 4 * c);
        const __nv_bfloat16 *nar = g < 4 ? P.q + (int64_t)row * P.ldq : P.dagg_lp + (int64_t)row * LG_HID;
        // B fragments of the own-head block in NATURAL channel order (k-step i of the head: channels 16i + 2q, +1 and
        // 16i + 8 + 2q, +1), matching A fragments read with ldmatrix
        const uint32_t *pq = reinterpret_cast<const uint32_t *>(nar + 64 * t) + q;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            kvf[2 * i] = __ldg(pq + 8 * i);
            kvf[2 * i + 1] = __ldg(pq + 8 * i + 4);
        }
    };

    Chunk A, B;
    int jA, jB = 0;
    uint32_t afA[4], afB[4] = {0u, 0u, 0u, 0u};
    const uint32_t vtile = wbase_u32 + 2 * LG_TILE;
    fetch(A, jA, afA);
    gather_rows(wbase_u32, P.k, P.ldk, jA, A.n, lane);
    gather_rows(vtile, P.v, P.ldv, jA, A.n, lane);
    cp_async_commit();

    uint4 xf[8];
    uint32_t kvf[8];
    load_row(xf, kvf, A.row);
    float acc[16][4];
    float D0 = 0.f, D1 = 0.f, G0 = 0.f, G1 = 0.f, mh0 = 0.f, mh1 = 0.f, iz0 = 0.f, iz1 = 0.f;
    int next_unwritten = r0;
    const int hsel = 2 * (q & 1);
    const bool lo_half = q < 2;
    B.n = 0; B.row = 0; B.pos = 0; B.first = false; B.last = false;

    for (int it = 0;; ++it) {
        const int s = it & 1;
        const bool have_next = !cur.done();
        if (have_next) {
            fetch(B, jB, afB);
            gather_rows(wbase_u32 + (uint32_t)(s ^ 1) * LG_TILE, P.k, P.ldk, jB, B.n, lane);
        }
        cp_async_commit();
        const uint32_t ktile = wbase_u32 + (uint32_t)s * LG_TILE;
        const int row = A.row, n = A.n;
        if (A.first) {
            zero_rows(next_unwritten, row);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            const F8 gf = ld8(P.dagg + (int64_t)row * LG_HID + lane * 8), af8 = ld8(P.agg + (int64_t)row * LG_HID + lane * 8);
            float dpart = 0.f, gpart = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) dpart = fmaf(gf.v[c], af8.v[c], dpart);
            if (P.cvec) {
                const F8 cf = ld8(P.cvec + lane * 8);
#pragma unroll
                for (int c = 0; c < 8; ++c) gpart = fmaf(gf.v[c], cf.v[c], gpart);
            }
            const float Dh = group_sum<8>(dpart), Gh = group_sum<8>(gpart);
            D0 = __shfl_sync(FULL, Dh, 8 * hsel);
            D1 = __shfl_sync(FULL, Dh, 8 * hsel + 8);
            G0 = __shfl_sync(FULL, Gh, 8 * hsel);
            G1 = __shfl_sync(FULL, Gh, 8 * hsel + 8);
            mh0 = __ldg(P.stat_m + (int64_t)row * LG_HEADS + hsel);
            mh1 = __ldg(P.stat_m + (int64_t)row * LG_HEADS + hsel + 1);
            iz0 = 1.0f / (__ldg(P.stat_z + (int64_t)row * LG_HEADS + hsel) + 1e-16f);
            iz1 = 1.0f / (__ldg(P.stat_z + (int64_t)row * LG_HEADS + hsel + 1) + 1e-16f);
        }
        float dr00 = 1.f, dr01 = 1.f, dr10 = 1.f, dr11 = 1.f;   // keep-scales: off the critical path (see forward)
        if (P.p_drop > 0.f)
            dropout_scale_quad(P.seed, rng_off, (uint64_t)A.pos, g, q, P.p_drop, P.inv_keep, dr00, dr01, dr10, dr11);
        cp_async_wait<1>();              // K and V of this chunk (K of the next may still be in flight)
        __syncwarp();

        // ---- phase 1 -------------------------------------------------------------------------------------------------
        float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb_[4] = {0.f, 0.f, 0.f, 0.f};
        float ck0[4] = {0.f, 0.f, 0.f, 0.f}, ck1[4] = {0.f, 0.f, 0.f, 0.f};
        {
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                float h[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint2 wf = lds64(img1 + (uint32_t)(((cb * 4 + u) * 32 + lane) * 8));
                    h[u][0] = h[u][1] = h[u][2] = h[u][3] = 0.f;
                    mma_bf16(h[u], afA[0], afA[1], afA[2], afA[3], wf.x, wf.y);
                }
                mma_bf16(ca, pack_relu_bf16(h[0][0], h[0][1]), pack_relu_bf16(h[0][2], h[0][3]),
                         pack_relu_bf16(h[1][0], h[1][1]), pack_relu_bf16(h[1][2], h[1][3]), xf[cb].x, xf[cb].y);
                mma_bf16(cb_, pack_relu_bf16(h[2][0], h[2][1]), pack_relu_bf16(h[2][2], h[2][3]),
                         pack_relu_bf16(h[3][0], h[3][1]), pack_relu_bf16(h[3][2], h[3][3]), xf[cb].z, xf[cb].w);
            }
            // A fragments with ldmatrix: 8 consecutive rows of the 528-byte-stride tiles fall in 8 distinct bank groups
            // (the 128-bit per-lane reads this replaces were 2-way conflicted: a third of the kernel's wavefronts)
            const uint32_t aoff = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * LG_ROWB + (lane >> 4) * 16);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                uint32_t a[4];
                ldsm_x4(a, ktile + aoff + kk * 32);
                const bool own = (kk >> 2) == g;
                mma_bf16((kk & 1) ? ck1 : ck0, a[0], a[1], a[2], a[3], own ? kvf[2 * (kk & 3)] : 0u, own ? kvf[2 * (kk & 3) + 1] : 0u);
            }
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                uint32_t a[4];
                ldsm_x4(a, vtile + aoff + kk * 32);
                const bool own = (kk >> 2) + 4 == g;
                mma_bf16((kk & 1) ? ck1 : ck0, a[0], a[1], a[2], a[3], own ? kvf[2 * (kk & 3)] : 0u, own ? kvf[2 * (kk & 3) + 1] : 0u);
            }
        }
        float c[4], o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            c[i] = (ca[i] + cb_[i]) + (ck0[i] + ck1[i]);
            o[i] = __shfl_xor_sync(FULL, c[i], 2);
        }
        const float sl00 = lo_half ? c[0] : o[0], sl01 = lo_half ? c[1] : o[1];
        const float sl10 = lo_half ? c[2] : o[2], sl11 = lo_half ? c[3] : o[3];
        const float d00 = lo_half ? o[0] : c[0], d01 = lo_half ? o[1] : c[1];
        const float d10 = lo_half ? o[2] : c[2], d11 = lo_half ? o[3] : c[3];
        const bool v0 = g < n, v1 = g + 8 < n;
        const float a00 = v0 ? fast_exp2(sl00 * P.scale_log2 - mh0) * iz0 : 0.f;
        const float a01 = v0 ? fast_exp2(sl01 * P.scale_log2 - mh1) * iz1 : 0.f;
        const float a10 = v1 ? fast_exp2(sl10 * P.scale_log2 - mh0) * iz0 : 0.f;
        const float a11 = v1 ? fast_exp2(sl11 * P.scale_log2 - mh1) * iz1 : 0.f;
        const float at00 = a00 * dr00, at01 = a01 * dr01, at10 = a10 * dr10, at11 = a11 * dr11;
        const float ds00 = v0 ? a00 * ((d00 + G0) * dr00 - D0) * P.scale : 0.f;
        const float ds01 = v0 ? a01 * ((d01 + G1) * dr01 - D1) * P.scale : 0.f;
        const float ds10 = v1 ? a10 * ((d10 + G0) * dr10 - D0) * P.scale : 0.f;
        const float ds11 = v1 ? a11 * ((d11 + G1) * dr11 - D1) * P.scale : 0.f;
        {
            const float2 w0 = lo_half ? make_float2(at00, at01) : make_float2(ds00, ds01);
            const float2 w1 = lo_half ? make_float2(at10, at11) : make_float2(ds10, ds11);
            if (v0) *reinterpret_cast<float2 *>(P.coef + (int64_t)(A.pos + g) * 8 + 2 * q) = w0;
            if (v1) *reinterpret_cast<float2 *>(P.coef + (int64_t)(A.pos + g + 8) * 8 + 2 * q) = w1;
        }
        const uint32_t tlo = movmatrix_trans(pack_bf16(ds00, ds01)), thi = movmatrix_trans(pack_bf16(ds10, ds11));
        const uint32_t bf0 = g < 4 ? tlo : 0u, bf1 = g < 4 ? thi : 0u;
        const uint32_t bk0 = g < 4 ? 0u : tlo, bk1 = g < 4 ? 0u : thi;

        if (have_next && B.first) load_row(xf, kvf, B.row);
        __syncwarp();                    // every lane is done with this chunk's V rows: the tile takes the next chunk's
        if (have_next) gather_rows(vtile, P.v, P.ldv, jB, B.n, lane);
        cp_async_commit();

        // ---- phase 2 -------------------------------------------------------------------------------------------------
        {
            const uint32_t toff = (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * LG_ROWB + ((lane >> 3) & 1) * 16);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint4 wf = lds128(img2 + (uint32_t)((j * 32 + lane) * 16));
                float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
                mma_bf16(t0, wf.x, wf.y, wf.z, wf.w, afA[0], afA[2]);
                mma_bf16(t1, wf.x, wf.y, wf.z, wf.w, afA[1], afA[3]);
                uint32_t a[4];
                ldsm_x4_trans(a, ktile + toff + j * 32);
                mma_bf16(acc[j], pack_relu_bf16(t0[0], t0[1]), pack_relu_bf16(t0[2], t0[3]), pack_relu_bf16(t1[0], t1[1]),
                         pack_relu_bf16(t1[2], t1[3]), bf0, bf1);
                mma_bf16(acc[j], a[0], a[1], a[2], a[3], bk0, bk1);
            }
        }
        if (A.last) {
            __syncwarp();
            const uint32_t stg = ktile;      // this chunk's K rows are dead after phase 2
            const uint32_t st0 = stg + (uint32_t)((2 * q) * LG_STG + g) * 4, st1 = st0 + LG_STG * 4;
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
                for (int t = 0; t < LG_HEADS; ++t) {
                    const float4 x = lds128f(stg + (uint32_t)(t * LG_STG + ch) * 4);
                    uint2 ov;
                    ov.x = pack_bf16(x.x, x.y);
                    ov.y = pack_bf16(x.z, x.w);
                    *reinterpret_cast<uint2 *>(P.bbar + (int64_t)row * P.ldbb + (int64_t)t * P.hsbb + ch) = ov;
                }
                const float4 y = lds128f(stg + (uint32_t)((4 + (ch >> 6)) * LG_STG + ch) * 4);
                uint2 ov;
                ov.x = pack_bf16(y.x, y.y);
                ov.y = pack_bf16(y.z, y.w);
                *reinterpret_cast<uint2 *>(P.dq + (int64_t)row * P.lddq + ch) = ov;
            }
            // the fp32 staging values must not stay behind as "K rows": rows past a later chunk's edge count are never
            // gathered but still enter K^T . DS with ds = 0, and 0 x (a NaN bit pattern) would poison dq
            __syncwarp();
            for (int off = lane * 16; off < LG_TILE; off += 512) sts128(stg + off, make_uint4(0, 0, 0, 0));
            next_unwritten = row + 1;
        }
        __syncwarp();
        if (!have_next) break;
        A = B;
        jA = jB;
#pragma unroll
        for (int i = 0; i < 4; ++i) afA[i] = afB[i];
    }
    cp_async_wait<0>();
    zero_rows(next_unwritten, r1);
    __syncwarp();
    }   // unit loop
    lg_work_done(P.work);
}

// ------------------------------------------------------------------------------------------------------------
// Gradient of the first angle-encoder layer from the per-edge coefficients of ALL line-graph layers:
//   df_ij   = sum_l sum_t ( ds^l_ij,t qt^l_i,t + a~^l_ij,t gt^l_i,t )          (never stored)
//   dW1ext  = sum_ij [h1pre_ij > 0] df_ij a_ij^T   ->  out[r * 256 + c] = dW1[c, r] (r < in_dim), out[in_dim*256 + c] = db1[c]
// Per chunk of 16 edges of target i:  DF^T[256 x 16] = X_i^T[256 x 8L] . COEF^T[8L x 16]  (X_i rows: gt^l heads, qt^l
// heads -- staged per target row with cp.async), ReLU mask from the recomputed pre-activation, then
// dW1ext[256 x 16] += mask(DF^T)[256 x 16 edges] . a[16 edges x 16].  Accumulators stay in registers for the whole
// kernel; CTA-level then grid-level reductions run in a fixed order (no atomics).
// ------------------------------------------------------------------------------------------------------------
constexpr int LGA_WARPS = 8;                                // two per scheduler; 8 x 32 x 255 registers is the whole file
constexpr int LGA_MAXL = 4;
constexpr int LGA_IMG_ROWS = 8 * LGA_MAXL;                  // 32 rows of 528 B per target row
constexpr int LGA_PER_WARP = LGA_IMG_ROWS * LG_ROWB;        // ONE image per warp: restaged at a row change (once per ~8 chunks;
                                                            // the other warps cover it), which is what lets 8 warps fit
constexpr int LGA_COEF = LGA_MAXL * LG_E * 32;              // per warp: coef rows of the NEXT chunk, [layer][edge][8] f32 (cp.async)

struct LgAngleParams {
    const __nv_bfloat16 *a_csr;
    const float *w1, *b1;
    const int32_t *rowptr;
    const float *coef[LGA_MAXL];              // [Ne, 8] CSR order
    const __nv_bfloat16 *qt[LGA_MAXL], *gt[LGA_MAXL];
    float *partials;                          // [grid, 16 * 256]
    int n_nodes, n_edges, in_dim, n_layers;
    int64_t ldqt, hsqt, ldgt, hsgt;
};

__global__ void __launch_bounds__(LGA_WARPS * 32, 1)
lg_angle_grad_kernel(const LgAngleParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    build_w1_images(smem_raw, P.w1, P.b1, P.in_dim);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const uint32_t img2 = smem_u32(smem_raw) + LG_IMG1;
    const uint32_t wbase_u32 = smem_u32(smem_raw) + LG_IMG + (uint32_t)warp * LGA_PER_WARP;
    const unsigned char *cbase = smem_raw + LG_IMG + LGA_WARPS * LGA_PER_WARP + warp * LGA_COEF;
    const uint32_t cbase_u32 = smem_u32(cbase);
    for (int off = lane * 16; off < LGA_PER_WARP; off += 512) sts128(wbase_u32 + off, make_uint4(0, 0, 0, 0));
    __syncthreads();

    float dw[16][2][4];
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int t = 0; t < 2; ++t) dw[j][t][0] = dw[j][t][1] = dw[j][t][2] = dw[j][t][3] = 0.f;

    const int64_t W = (int64_t)gridDim.x * LGA_WARPS, w = (int64_t)blockIdx.x * LGA_WARPS + warp;
    const int64_t total = (int64_t)P.n_edges + (int64_t)ROW_KAPPA * P.n_nodes;
    const int r0 = (int)row_bound(P.rowptr, P.n_nodes, total * w / W);
    const int r1 = (int)row_bound(P.rowptr, P.n_nodes, total * (w + 1) / W);
    ChunkCursor cur;
    cur.init(P.rowptr, r0, r1);
    const int n_pairs = (P.n_layers + 1) >> 1;

    if (!cur.done()) {
        const uint32_t *ap = reinterpret_cast<const uint32_t *>(P.a_csr);
        // image rows of layer l: 8l + t = gt^l head t (pairs with a~), 8l + 4 + t = qt^l head t (pairs with ds)
        auto stage_row = [&](int row) {
            const uint32_t dst = wbase_u32 + lane * 16;
#pragma unroll
            for (int l = 0; l < LGA_MAXL; ++l) {
                if (l < P.n_layers) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        cp_async16(dst + (8 * l + t) * LG_ROWB, P.gt[l] + (int64_t)row * P.ldgt + (int64_t)t * P.hsgt + lane * 8);
                        cp_async16(dst + (8 * l + 4 + t) * LG_ROWB, P.qt[l] + (int64_t)row * P.ldqt + (int64_t)t * P.hsqt + lane * 8);
                    }
                }
            }
        };
        auto fetch = [&](Chunk &c, uint32_t (&af)[4]) {
            c = cur.take(P.rowptr);
            const int e0 = c.pos + g, e1 = c.pos + g + 8;
            af[0] = e0 < P.n_edges ? __ldg(ap + (int64_t)e0 * 8 + q) : 0u;
            af[1] = e1 < P.n_edges ? __ldg(ap + (int64_t)e1 * 8 + q) : 0u;
            af[2] = e0 < P.n_edges ? __ldg(ap + (int64_t)e0 * 8 + 4 + q) : 0u;
            af[3] = e1 < P.n_edges ? __ldg(ap + (int64_t)e1 * 8 + 4 + q) : 0u;
        };
        // the 16 coefficient rows of a chunk are 512 contiguous bytes per layer (CSR order): one 16-byte piece per lane,
        // copied one chunk ahead so their latency hides behind the MMAs of the current chunk
        auto stage_coef = [&](int pos) {
            if (pos + (lane >> 1) < P.n_edges) {
#pragma unroll
                for (int l = 0; l < LGA_MAXL; ++l)
                    if (l < P.n_layers) cp_async16(cbase_u32 + l * (LG_E * 32) + lane * 16, P.coef[l] + (int64_t)pos * 8 + lane * 4);
            }
        };
        Chunk A, B;
        uint32_t afA[4], afB[4] = {0u, 0u, 0u, 0u};
        fetch(A, afA);
        stage_row(A.row);
        stage_coef(A.pos);
        cp_async_commit();
        B.n = 0; B.row = 0; B.pos = 0; B.first = false; B.last = false;
        const uint32_t toff = (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * LG_ROWB + ((lane >> 3) & 1) * 16);

        for (;;) {
            const bool have_next = !cur.done();
            if (have_next) fetch(B, afB);
            cp_async_wait<0>();              // this chunk's coefficient rows and its target row's image (both issued a chunk ago)
            __syncwarp();
            const uint32_t image = wbase_u32;
            const int n = A.n;
            const bool v0 = g < n, v1 = g + 8 < n;
            // B fragments of COEF^T: k = coefficient column (2q, 2q+1) of layer 2kp (+8: layer 2kp+1), n = edge g
            uint32_t cf[2][2][2];   // [pair][n-tile][b0/b1]
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const int l = 2 * kp + hlf;
                    float2 x0 = make_float2(0.f, 0.f), x1 = make_float2(0.f, 0.f);
                    if (l < P.n_layers) {
                        if (v0) x0 = *reinterpret_cast<const float2 *>(cbase + l * (LG_E * 32) + g * 32 + q * 8);
                        if (v1) x1 = *reinterpret_cast<const float2 *>(cbase + l * (LG_E * 32) + (g + 8) * 32 + q * 8);
                    }
                    cf[kp][0][hlf] = pack_bf16(x0.x, x0.y);
                    cf[kp][1][hlf] = pack_bf16(x1.x, x1.y);
                }
            }
            __syncwarp();                    // every lane has its fragments: the coefficient buffer can take the next chunk
            if (have_next) stage_coef(B.pos);
            cp_async_commit();
            const uint32_t ba0 = movmatrix_trans(afA[0]), ba1 = movmatrix_trans(afA[1]);
            const uint32_t ba2 = movmatrix_trans(afA[2]), ba3 = movmatrix_trans(afA[3]);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint4 wf = lds128(img2 + (uint32_t)((j * 32 + lane) * 16));
                float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
                mma_bf16(t0, wf.x, wf.y, wf.z, wf.w, afA[0], afA[2]);     // pre-activation^T, edges 0..7
                mma_bf16(t1, wf.x, wf.y, wf.z, wf.w, afA[1], afA[3]);     // edges 8..15
                float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    if (kp < n_pairs) {
                        uint32_t a[4];
                        ldsm_x4_trans(a, image + (uint32_t)(16 * kp) * LG_ROWB + toff + j * 32);
                        mma_bf16(d0, a[0], a[1], a[2], a[3], cf[kp][0][0], cf[kp][0][1]);
                        mma_bf16(d1, a[0], a[1], a[2], a[3], cf[kp][1][0], cf[kp][1][1]);
                    }
                }
                const uint32_t m0 = pack_bf16(t0[0] > 0.f ? d0[0] : 0.f, t0[1] > 0.f ? d0[1] : 0.f);
                const uint32_t m1 = pack_bf16(t0[2] > 0.f ? d0[2] : 0.f, t0[3] > 0.f ? d0[3] : 0.f);
                const uint32_t m2 = pack_bf16(t1[0] > 0.f ? d1[0] : 0.f, t1[1] > 0.f ? d1[1] : 0.f);
                const uint32_t m3 = pack_bf16(t1[2] > 0.f ? d1[2] : 0.f, t1[3] > 0.f ? d1[3] : 0.f);
                mma_bf16(dw[j][0], m0, m1, m2, m3, ba0, ba1);
                mma_bf16(dw[j][1], m0, m1, m2, m3, ba2, ba3);
            }
            __syncwarp();
            if (!have_next) break;
            if (B.first) {                   // every lane is done with this row's image
                stage_row(B.row);
                cp_async_commit();
            }
            A = B;
#pragma unroll
            for (int i = 0; i < 4; ++i) afA[i] = afB[i];
        }
        cp_async_wait<0>();
    }
    // ---- CTA reduction (fixed warp order), then one partial row per CTA ----------------------------------------------
    __syncthreads();
    float *red = reinterpret_cast<float *>(smem_raw);        // [LGA_WARPS][16 * 256], r-major
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            float *dst = red + (size_t)warp * 4096;
            dst[(8 * t + 2 * q) * 256 + 16 * j + g] = dw[j][t][0];
            dst[(8 * t + 2 * q + 1) * 256 + 16 * j + g] = dw[j][t][1];
            dst[(8 * t + 2 * q) * 256 + 16 * j + g + 8] = dw[j][t][2];
            dst[(8 * t + 2 * q + 1) * 256 + 16 * j + g + 8] = dw[j][t][3];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
        float sum = 0.f;
#pragma unroll
        for (int wv = 0; wv < LGA_WARPS; ++wv) sum += red[(size_t)wv * 4096 + i];
        P.partials[(size_t)blockIdx.x * 4096 + i] = sum;
    }
}

__global__ void lg_angle_reduce_kernel(const float *__restrict__ partials, float *__restrict__ out, int n_blocks, int width) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width) return;
    float sum = 0.f;
    for (int b = 0; b < n_blocks; ++b) sum += partials[(size_t)b * 4096 + i];
    out[i] = sum;
}

static int lg_grid(int64_t n_nodes, int64_t n_edges, int warps) {
    const int64_t work = n_edges + (int64_t)ROW_KAPPA * n_nodes;
    int64_t blocks = 148 * 2;          // one CTA is resident per SM; two per SM even out the row-granular static split
    const int64_t min_work_per_warp = 64;
    if (work / (blocks * warps) < min_work_per_warp) blocks = work / (min_work_per_warp * warps) + 1;
    return (int)blocks;
}

// persistent grid of the dynamically scheduled kernels: one CTA per SM, fewer when there is little work
static int lg_persistent_grid(int64_t n_nodes, int64_t n_edges, int warps) {
    const int64_t units = (n_edges + (int64_t)ROW_KAPPA * n_nodes + LG_UNIT - 1) / LG_UNIT;
    const int64_t blocks = (units + warps - 1) / warps;
    return (int)(blocks < 1 ? 1 : (blocks > 148 ? 148 : blocks));
}

// a_csr[p, :] = (bf16(a[eid[p], 0..in_dim-1]), 1, 0, ...)
__global__ void lg_pack_angles_kernel(const float *__restrict__ a, const int32_t *__restrict__ eid,
                                      __nv_bfloat16 *__restrict__ a_csr, int64_t n_edges, int in_dim) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (edge, pair of columns)
    const int64_t p = i >> 3;
    const int c = (int)(i & 7) * 2;
    if (p >= n_edges) return;
    const int64_t e = eid[p];
    const float x0 = c < in_dim ? __ldg(a + e * in_dim + c) : (c == in_dim ? 1.f : 0.f);
    const float x1 = c + 1 < in_dim ? __ldg(a + e * in_dim + c + 1) : (c + 1 == in_dim ? 1.f : 0.f);
    reinterpret_cast<uint32_t *>(a_csr)[i] = pack_bf16(x0, x1);
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_lgattn_supported(int hidden, int heads, int in_dim, int dtype) {
    return hidden == LG_HID && heads == LG_HEADS && in_dim >= 1 && in_dim <= 15 && dtype == ALIGNN_BF16;
}

extern "C" int alignn_lg_pack_angles(const float *a, const int32_t *eid, void *a_csr, int64_t n_edges, int in_dim,
                                     void *stream) {
    if (n_edges < 0 || in_dim < 1 || in_dim > 15) return ALIGNN_ERR_BAD_ARG;
    if (n_edges == 0) return ALIGNN_OK;
    if (!a || !eid || !a_csr || !aligned16(a_csr)) return ALIGNN_ERR_BAD_ARG;
    const int64_t threads = n_edges * 8;
    lg_pack_angles_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        a, eid, (__nv_bfloat16 *)a_csr, n_edges, in_dim);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int alignn_lgattn_fwd(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                                 const void *qt, int64_t ldqt, int64_t hsqt,
                                 const void *a_csr, const float *w1, const float *b1, int in_dim,
                                 const int32_t *rowptr, const int32_t *col,
                                 float *aggv, void *abar, int64_t ldab, int64_t hsab,
                                 float *stat_m, float *stat_z, float *stat_s,
                                 int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                 float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *work,
                                 void *stream) {
    if (!alignn_lgattn_supported(hidden, heads, in_dim, dtype)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes < 0 || n_edges < 0 || !(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 64) return ALIGNN_ERR_BAD_SHAPE;
    if (ldq >= ((int64_t)1 << 30) || ldk >= ((int64_t)1 << 30) || ldv >= ((int64_t)1 << 30)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!q || !k || !v || !qt || !w1 || !b1 || !rowptr || !aggv || !abar || !stat_m || !stat_z || !stat_s || !work)
        return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!a_csr || !col)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(qt) || !aligned16(a_csr) || !aligned16(aggv) ||
        !aligned16(abar) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldqt % 8) || (hsqt % 8) || (ldab % 8) || (hsab % 8))
        return ALIGNN_ERR_BAD_ARG;
    LgFwdParams p;
    p.q = (const __nv_bfloat16 *)q; p.k = (const __nv_bfloat16 *)k; p.v = (const __nv_bfloat16 *)v;
    p.qt = (const __nv_bfloat16 *)qt; p.a_csr = (const __nv_bfloat16 *)a_csr; p.w1 = w1; p.b1 = b1;
    p.rowptr = rowptr; p.col = col; p.aggv = aggv; p.abar = (__nv_bfloat16 *)abar;
    p.stat_m = stat_m; p.stat_z = stat_z; p.stat_s = stat_s; p.rng_step = rng_step; p.work = (unsigned int *)work;
    p.n_nodes = (int)n_nodes; p.n_edges = (int)n_edges; p.in_dim = in_dim;
    p.ldq = (int)ldq; p.ldk = (int)ldk; p.ldv = (int)ldv; p.ldqt = ldqt; p.hsqt = hsqt; p.ldab = ldab; p.hsab = hsab;
    p.scale_log2 = LOG2E / sqrtf((float)(hidden / heads));
    p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    p.seed = seed; p.offset = offset;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int SMEM = LG_IMG + LGF_WARPS * LGF_PER_WARP;
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(lgattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    lgattn_fwd_kernel<<<lg_persistent_grid(n_nodes, n_edges, LGF_WARPS), LGF_WARPS * 32, SMEM, st>>>(p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int alignn_lgattn_bwd_dst(const float *dagg, const void *dagg_lp, const float *agg,
                                     const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                                     const void *qt, int64_t ldqt, int64_t hsqt, const void *gt, int64_t ldgt, int64_t hsgt,
                                     const float *cvec, const void *a_csr, const float *w1, const float *b1, int in_dim,
                                     const float *stat_m, const float *stat_z, const int32_t *rowptr, const int32_t *col,
                                     void *dq, int64_t lddq, void *bbar, int64_t ldbb, int64_t hsbb, float *coef,
                                     int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                     float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *work,
                                     void *stream) {
    if (!alignn_lgattn_supported(hidden, heads, in_dim, dtype) || !work) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes < 0 || n_edges < 0 || !(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 64) return ALIGNN_ERR_BAD_SHAPE;
    if (ldq >= ((int64_t)1 << 30) || ldk >= ((int64_t)1 << 30) || ldv >= ((int64_t)1 << 30) || lddq >= ((int64_t)1 << 30))
        return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!dagg || !dagg_lp || !agg || !q || !k || !v || !qt || !gt || !w1 || !b1 || !stat_m || !stat_z || !rowptr || !dq ||
        !bbar)
        return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!a_csr || !col || !coef)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(dagg) || !aligned16(dagg_lp) || !aligned16(agg) || !aligned16(q) || !aligned16(k) || !aligned16(v) ||
        !aligned16(qt) || !aligned16(gt) || !aligned16(cvec) || !aligned16(a_csr) || !aligned16(dq) || !aligned16(bbar) ||
        !aligned16(coef) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (lddq % 8) || (ldqt % 8) || (hsqt % 8) || (ldgt % 8) ||
        (hsgt % 8) || (ldbb % 8) || (hsbb % 8))
        return ALIGNN_ERR_BAD_ARG;
    LgBwdParams p;
    p.dagg = dagg; p.agg = agg; p.dagg_lp = (const __nv_bfloat16 *)dagg_lp;
    p.q = (const __nv_bfloat16 *)q; p.k = (const __nv_bfloat16 *)k; p.v = (const __nv_bfloat16 *)v;
    p.qt = (const __nv_bfloat16 *)qt; p.gt = (const __nv_bfloat16 *)gt; p.cvec = cvec;
    p.a_csr = (const __nv_bfloat16 *)a_csr; p.w1 = w1; p.b1 = b1; p.stat_m = stat_m; p.stat_z = stat_z;
    p.rowptr = rowptr; p.col = col; p.dq = (__nv_bfloat16 *)dq; p.bbar = (__nv_bfloat16 *)bbar; p.coef = coef;
    p.rng_step = rng_step; p.work = (unsigned int *)work; p.n_nodes = (int)n_nodes; p.n_edges = (int)n_edges; p.in_dim = in_dim;
    p.ldq = (int)ldq; p.ldk = (int)ldk; p.ldv = (int)ldv; p.lddq = (int)lddq;
    p.ldqt = ldqt; p.hsqt = hsqt; p.ldgt = ldgt; p.hsgt = hsgt; p.ldbb = ldbb; p.hsbb = hsbb;
    p.scale = 1.0f / sqrtf((float)(hidden / heads));
    p.scale_log2 = p.scale * LOG2E;
    p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    p.seed = seed; p.offset = offset;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int SMEM = LG_IMG + LGB_WARPS * LGB_PER_WARP;
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(lgattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    lgattn_bwd_kernel<<<lg_persistent_grid(n_nodes, n_edges, LGB_WARPS), LGB_WARPS * 32, SMEM, st>>>(p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int64_t alignn_lg_angle_grad_partial_floats(int64_t n_nodes, int64_t n_edges) {
    return (int64_t)lg_grid(n_nodes, n_edges, LGA_WARPS) * 4096;
}

/* coef / qt / gt: arrays of `n_layers` (<= 4) device pointers (host arrays of pointers). */
extern "C" int alignn_lg_angle_grad2(const void *a_csr, const float *w1, const float *b1, int in_dim,
                                     const int32_t *rowptr, int n_layers, const float *const *coef,
                                     const void *const *qt, const void *const *gt,
                                     int64_t ldqt, int64_t hsqt, int64_t ldgt, int64_t hsgt,
                                     float *partials, float *out, int64_t n_nodes, int64_t n_edges, int max_blocks, void *stream);

extern "C" int alignn_lg_angle_grad(const void *a_csr, const float *w1, const float *b1, int in_dim,
                                    const int32_t *rowptr, int n_layers, const float *const *coef,
                                    const void *const *qt, const void *const *gt,
                                    int64_t ldqt, int64_t hsqt, int64_t ldgt, int64_t hsgt,
                                    float *partials, float *out, int64_t n_nodes, int64_t n_edges, void *stream) {
    return alignn_lg_angle_grad2(a_csr, w1, b1, in_dim, rowptr, n_layers, coef, qt, gt, ldqt, hsqt, ldgt, hsgt, partials, out,
                                 n_nodes, n_edges, 0, stream);
}

/* max_blocks > 0: at most that many CTAs (one per SM): the kernel then leaves the other SMs to concurrent streams. */
extern "C" int alignn_lg_angle_grad2(const void *a_csr, const float *w1, const float *b1, int in_dim,
                                     const int32_t *rowptr, int n_layers, const float *const *coef,
                                     const void *const *qt, const void *const *gt,
                                     int64_t ldqt, int64_t hsqt, int64_t ldgt, int64_t hsgt,
                                     float *partials, float *out, int64_t n_nodes, int64_t n_edges, int max_blocks, void *stream) {
    if (in_dim < 1 || in_dim > 15 || n_layers < 1 || n_layers > LGA_MAXL || n_nodes < 0 || n_edges < 0) return ALIGNN_ERR_BAD_ARG;
    if (!w1 || !b1 || !out || !partials || !coef || !qt || !gt) return ALIGNN_ERR_BAD_ARG;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 64) return ALIGNN_ERR_BAD_SHAPE;
    if ((ldqt % 8) || (hsqt % 8) || (ldgt % 8) || (hsgt % 8)) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int width = (in_dim + 1) * 256;
    if (n_nodes == 0 || n_edges == 0) {
        ALIGNN_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)width * sizeof(float), st));
        return ALIGNN_OK;
    }
    if (!a_csr || !rowptr || !aligned16(a_csr)) return ALIGNN_ERR_BAD_ARG;
    LgAngleParams p;
    p.a_csr = (const __nv_bfloat16 *)a_csr; p.w1 = w1; p.b1 = b1; p.rowptr = rowptr;
    for (int l = 0; l < LGA_MAXL; ++l) {
        p.coef[l] = l < n_layers ? coef[l] : nullptr;
        p.qt[l] = l < n_layers ? (const __nv_bfloat16 *)qt[l] : nullptr;
        p.gt[l] = l < n_layers ? (const __nv_bfloat16 *)gt[l] : nullptr;
        if (l < n_layers && (!coef[l] || !qt[l] || !gt[l] || !aligned16(qt[l]) || !aligned16(gt[l]) || !aligned16(coef[l])))
            return ALIGNN_ERR_BAD_ARG;
    }
    p.partials = partials; p.n_nodes = (int)n_nodes; p.n_edges = (int)n_edges; p.in_dim = in_dim; p.n_layers = n_layers;
    p.ldqt = ldqt; p.hsqt = hsqt; p.ldgt = ldgt; p.hsgt = hsgt;
    constexpr int SMEM = LG_IMG + LGA_WARPS * (LGA_PER_WARP + LGA_COEF);
    static_assert(SMEM >= LGA_WARPS * 4096 * 4, "reduction buffer aliases the pipeline buffers");
    static_assert(SMEM <= 227 * 1024, "over the per-CTA shared-memory limit");
    int grid = lg_grid(n_nodes, n_edges, LGA_WARPS);
    if (max_blocks > 0 && max_blocks < grid) grid = max_blocks;
    if (const char *e = getenv("ALIGNN_LGA_BLOCKS")) {           // tuning knob (never more CTAs than the partials buffer holds)
        const int want = atoi(e);
        if (want >= 1 && want < grid) grid = want;
    }
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(lg_angle_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    lg_angle_grad_kernel<<<grid, LGA_WARPS * 32, SMEM, st>>>(p);
    ALIGNN_LAUNCH_CHECK();
    lg_angle_reduce_kernel<<<(width + 255) / 256, 256, 0, st>>>(partials, out, grid, width);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
