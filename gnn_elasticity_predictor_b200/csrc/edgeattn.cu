// Streaming edge-attention conv with LINEAR edge features: the per-edge projection e = Wc f + c is never
// materialised (forward or backward).
//
// Replaces, for hidden = 256, the same reference arithmetic as conv.cu (PyG TransformerConv.message +
// utils.softmax + 'add' aggregation at scripts/train.py:315,334) PLUS the per-edge dense projections that feed
// it: `lin_edge` (train.py:308,326 -> PyG lin_edge), the second Linear of `angle_encoder` (train.py:360-364)
// and `edge_proj` (train.py:324,333).  With f the per-edge feature row (h1 = relu(W1 a + b1) on the line
// graph, the bond state on the atom graph) and e = Wc f + c (Wc = W_e W2, c = W_e b2):
//
//   s_ij,t  = ( <q_i,t , k_j,t> + <qt_i,t , f_ij> ) / sqrt(C)         qt_i,t = Wc[t]^T q_i,t   (per NODE)
//             (<q_i,t, c_t> is constant over j and cancels in the softmax)
//   agg_i,t = sum_j a~_ij,t v_j,t  +  Wc[t] abar_i,t  +  c_t S_i,t    abar_i,t = sum_j a~_ij,t f_ij ; S = sum_j a~
//
// so the E x H x H per-edge GEMMs become per-node GEMMs (11x fewer rows on the line graph) and each pass
// streams f exactly once.  Backward mirrors it (bbar_i,t = sum_j ds_ij,t f_ij / sqrt(C), gt_i,t = Wc[t]^T g_i,t):
//   df_ij   = sum_t ( dss_ij,t qt_i,t + a~_ij,t gt_i,t )              accumulated over layers in place
//
// Mapping: one warp streams a contiguous, cost-balanced range of target rows through a warp-private ring of
// TMA bulk copies (stream.cuh); lane l owns channels [8l, 8l+8) of every 256-wide row; per-head logits use a
// transposed butterfly (6 shuffles for 4 heads); online softmax; register accumulation in edge order (no
// atomics, bit-reproducible).  The kernels are FP32-FMA/issue bound rather than HBM bound by design
// (~2.6k MAC per edge forward): see DESIGN.md.
#include <math.h>

#include "stream.cuh"

namespace alignn {

constexpr int EA_HIDDEN = 256;
constexpr int EA_WARPS = 4;

template <typename T>
struct EaCfg {
    static constexpr int RB = EA_HIDDEN * (int)sizeof(T);   // bytes per row
    static constexpr int EPS = 2048 / RB;                   // edges per stage: 4 (bf16) / 2 (fp32)
};

template <int HEADS>
__device__ __forceinline__ float bcast_head(float x, int t) {
    return __shfl_sync(FULL, x, t * (32 / HEADS));
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <typename T, int HEADS, int STAGES>
struct FwdSmem {
    static constexpr int RB = EaCfg<T>::RB, EPS = EaCfg<T>::EPS;
    static constexpr int EDGE_RING = STAGES * EPS * 3 * RB;            // f, k, v rows
    static constexpr int ROW_RING = STAGES * (1 + HEADS) * RB;         // q, qt[HEADS]
    static constexpr int META = 128;                                   // barriers + eid stash
    static constexpr int PER_WARP = EDGE_RING + ROW_RING + META;
};

struct EaFwdParams {
    const void *q, *k, *v;       // [Nn, *] with row strides ldq/ldk/ldv (elements)
    const void *qt;              // [HEADS, Nn, 256]
    const void *feat;            // [Ne, 256]
    const int32_t *rowptr, *col, *eid;
    float *aggv;                 // [Nn, 256]
    void *abar;                  // [HEADS, Nn, 256]
    float *stat_m, *stat_z, *stat_s;   // [Nn, HEADS]
    int64_t n_nodes, n_edges;
    int64_t ldq, ldk, ldv;
    float scale_log2, p_drop, inv_keep;
    uint64_t seed, offset;
};

template <typename T, int HEADS, int STAGES>
__global__ void __launch_bounds__(EA_WARPS * 32)
edgeattn_fwd_kernel(const EaFwdParams P) {
    using S = FwdSmem<T, HEADS, STAGES>;
    constexpr int RB = S::RB, EPS = S::EPS, LPH = 32 / HEADS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * S::PER_WARP;
    unsigned char *edge_ring = base;
    unsigned char *row_ring = base + S::EDGE_RING;
    uint64_t *full = reinterpret_cast<uint64_t *>(base + S::EDGE_RING + S::ROW_RING);
    int *eid_stash = reinterpret_cast<int *>(base + S::EDGE_RING + S::ROW_RING + 64);   // [STAGES][EPS]

    const T *q = (const T *)P.q, *k = (const T *)P.k, *v = (const T *)P.v, *qt = (const T *)P.qt;
    const T *feat = (const T *)P.feat;
    T *abar = (T *)P.abar;
    const int myhead = lane / LPH;
    const int ch = lane * 8;

    // this warp's rows
    const int64_t W = (int64_t)gridDim.x * EA_WARPS, w = (int64_t)blockIdx.x * EA_WARPS + warp;
    const int64_t total = P.n_edges + (int64_t)ROW_KAPPA * P.n_nodes;
    const int64_t r0 = row_bound(P.rowptr, P.n_nodes, total * w / W);
    const int64_t r1 = row_bound(P.rowptr, P.n_nodes, total * (w + 1) / W);
    if (r0 >= r1) return;
    const int e_end = __ldg(P.rowptr + r1);

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        mbar_fence_init();
    }
    __syncwarp();

    StageCursor<EPS> prod, cons;
    prod.init(P.rowptr, r0, r1);
    cons = prod;
    int wbase = prod.pos;
    IndexWindow wcol, weid;
    wcol.init(P.col, wbase, e_end, lane);
    weid.init(P.eid, wbase, e_end, lane);

    auto issue = [&](int s) {
        if (prod.done()) return;
        const int n = prod.count();
        const bool first = prod.first;
        const int my_u = lane / 3, my_r = lane - 3 * my_u;
        const int o = prod.pos + min(my_u, EPS - 1) - wbase;
        const int j = wcol.get(o), id = weid.get(o);
        if (lane == 0) mbar_expect_tx(full + s, (uint32_t)(n * 3 * RB + (first ? (1 + HEADS) * RB : 0)));
        __syncwarp();
        if (lane < 3 * EPS && my_u < n) {
            const T *src = my_r == 0 ? feat + (int64_t)id * EA_HIDDEN
                         : my_r == 1 ? k + (int64_t)j * P.ldk : v + (int64_t)j * P.ldv;
            bulk_g2s(edge_ring + ((size_t)(s * EPS + my_u) * 3 + my_r) * RB, src, RB, full + s);
            if (my_r == 0) eid_stash[s * EPS + my_u] = id;
        }
        if (first && lane >= 16 && lane < 16 + 1 + HEADS) {
            const int t = lane - 16;
            const T *src = t == 0 ? q + prod.row * P.ldq : qt + ((int64_t)(t - 1) * P.n_nodes + prod.row) * EA_HIDDEN;
            bulk_g2s(row_ring + ((size_t)s * (1 + HEADS) + t) * RB, src, RB, full + s);
        }
        prod.advance(P.rowptr);
        if (prod.pos - wbase >= 32 && !prod.done()) {
            wbase += 32;
            wcol.shift(P.col, wbase, e_end, lane);
            weid.shift(P.eid, wbase, e_end, lane);
        }
    };

    auto zero_rows = [&](int64_t lo, int64_t hi) {   // rows without in-edges: all outputs are zero
        F8 zf;
#pragma unroll
        for (int c = 0; c < 8; ++c) zf.v[c] = 0.f;
        for (int64_t r = lo; r < hi; ++r) {
            st8(P.aggv + r * EA_HIDDEN + ch, zf);
#pragma unroll
            for (int t = 0; t < HEADS; ++t) st8(abar + ((int64_t)t * P.n_nodes + r) * EA_HIDDEN + ch, zf);
            if (lane < HEADS) {
                P.stat_m[r * HEADS + lane] = 0.f;
                P.stat_z[r * HEADS + lane] = 0.f;
                P.stat_s[r * HEADS + lane] = 0.f;
            }
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    __syncwarp();   // eid stash of the prologue stages is visible to every lane

    // row state
    F8 qf, qtf[HEADS], ab[HEADS], acc;
    float m = -INFINITY, z = 0.f, zd = 0.f;
    int64_t next_unwritten = r0;

    for (int it = 0; !cons.done(); ++it) {
        const int s = it % STAGES;
        issue((it + STAGES - 1) % STAGES);
        mbar_wait(full + s, (uint32_t)((it / STAGES) & 1));
        const int n = cons.count();
        const unsigned char *stage = edge_ring + (size_t)s * EPS * 3 * RB;
        if (cons.first) {
            zero_rows(next_unwritten, cons.row);
            const unsigned char *rs = row_ring + (size_t)s * (1 + HEADS) * RB;
            qf = lds8(reinterpret_cast<const T *>(rs) + ch);
#pragma unroll
            for (int c = 0; c < 8; ++c) qf.v[c] *= P.scale_log2;
#pragma unroll
            for (int t = 0; t < HEADS; ++t) {
                qtf[t] = lds8(reinterpret_cast<const T *>(rs + (size_t)(1 + t) * RB) + ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    qtf[t].v[c] *= P.scale_log2;
                    ab[t].v[c] = 0.f;
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) acc.v[c] = 0.f;
            m = -INFINITY;
            z = zd = 0.f;
        }
        // pass 1: logits of the stage's edges (log2 units), own head per lane
        float sl[EPS];
#pragma unroll
        for (int u = 0; u < EPS; ++u) {
            float p[HEADS];
#pragma unroll
            for (int t = 0; t < HEADS; ++t) p[t] = 0.f;
            if (u < n) {
                const F8 hf = lds8(reinterpret_cast<const T *>(stage + (size_t)(u * 3 + 0) * RB) + ch);
                const F8 kf = lds8(reinterpret_cast<const T *>(stage + (size_t)(u * 3 + 1) * RB) + ch);
                float pk = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    pk = fmaf(qf.v[c], kf.v[c], pk);
#pragma unroll
                    for (int t = 0; t < HEADS; ++t) p[t] = fmaf(qtf[t].v[c], hf.v[c], p[t]);
                }
#pragma unroll
                for (int t = 0; t < HEADS; ++t) p[t] += (t == myhead) ? pk : 0.f;
            }
            sl[u] = reduce_heads<HEADS>(p, lane);
            if (u >= n) sl[u] = -INFINITY;
        }
        float m_new = m;
#pragma unroll
        for (int u = 0; u < EPS; ++u) m_new = fmaxf(m_new, sl[u]);
        const float corr = fast_exp2(m - m_new);     // m = -inf -> 0 (n >= 1, so m_new is finite)
        z *= corr;
        zd *= corr;
        float corr_h[HEADS];
#pragma unroll
        for (int t = 0; t < HEADS; ++t) corr_h[t] = bcast_head<HEADS>(corr, t);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            acc.v[c] *= corr;
#pragma unroll
            for (int t = 0; t < HEADS; ++t) ab[t].v[c] *= corr_h[t];
        }
        m = m_new;
        // pass 2: accumulate
#pragma unroll
        for (int u = 0; u < EPS; ++u) {
            const float wgt = fast_exp2(sl[u] - m_new);   // inactive -> 0
            z += wgt;
            float wd = wgt;
            if (P.p_drop > 0.f && u < n) {
                const int id = eid_stash[s * EPS + u];
                wd = wgt * dropout_scale(P.seed, P.offset, (uint64_t)id * HEADS + myhead, P.p_drop, P.inv_keep);
            }
            zd += wd;
            float wh[HEADS];
#pragma unroll
            for (int t = 0; t < HEADS; ++t) wh[t] = bcast_head<HEADS>(wd, t);
            if (u < n) {
                const F8 hf = lds8(reinterpret_cast<const T *>(stage + (size_t)(u * 3 + 0) * RB) + ch);
                const F8 vf = lds8(reinterpret_cast<const T *>(stage + (size_t)(u * 3 + 2) * RB) + ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    acc.v[c] = fmaf(wd, vf.v[c], acc.v[c]);
#pragma unroll
                    for (int t = 0; t < HEADS; ++t) ab[t].v[c] = fmaf(wh[t], hf.v[c], ab[t].v[c]);
                }
            }
        }
        if (cons.last()) {
            const float inv = 1.0f / (z + 1e-16f);
            const int64_t r = cons.row;
#pragma unroll
            for (int c = 0; c < 8; ++c) acc.v[c] *= inv;
            st8(P.aggv + r * EA_HIDDEN + ch, acc);
#pragma unroll
            for (int t = 0; t < HEADS; ++t) {
                const float inv_t = bcast_head<HEADS>(inv, t);
#pragma unroll
                for (int c = 0; c < 8; ++c) ab[t].v[c] *= inv_t;
                st8(abar + ((int64_t)t * P.n_nodes + r) * EA_HIDDEN + ch, ab[t]);
            }
            if (lane % LPH == 0) {
                P.stat_m[r * HEADS + myhead] = m;
                P.stat_z[r * HEADS + myhead] = z;
                P.stat_s[r * HEADS + myhead] = zd * inv;
            }
            next_unwritten = r + 1;
        }
        __syncwarp();   // every lane is done with stage s before it is refilled
        cons.advance(P.rowptr);
    }
    zero_rows(next_unwritten, r1);
}

// ------------------------------------------------------------------------------------------------------------
// backward, target-sorted pass
// ------------------------------------------------------------------------------------------------------------
template <typename T, int HEADS, int STAGES, bool ACCUM>
struct BwdSmem {
    static constexpr int RB = EaCfg<T>::RB, EPS = EaCfg<T>::EPS;
    static constexpr int ROWS_PER_EDGE = ACCUM ? 4 : 3;                 // f, k, v (+ running df)
    static constexpr int EDGE_RING = STAGES * EPS * ROWS_PER_EDGE * RB;
    static constexpr int ROW_RING = STAGES * (1 + 2 * HEADS) * RB;      // q, qt[HEADS], gt[HEADS]
    static constexpr int META = 128;
    static constexpr int PER_WARP = EDGE_RING + ROW_RING + META;
};

struct EaBwdParams {
    const float *dagg, *agg;     // [Nn, 256] f32
    const void *q, *k, *v;       // strided rows
    const void *qt, *gt;         // [HEADS, Nn, 256]
    const float *cvec;           // [256] f32 (may be null: c = 0)
    const void *feat;            // [Ne, 256]
    const float *stat_m, *stat_z;
    const int32_t *rowptr, *col, *eid;
    void *dq;                    // [Nn, *] stride lddq
    void *bbar;                  // [HEADS, Nn, 256]
    float *coef;                 // [Ne, 2*HEADS]
    const void *df_in;           // [Ne, 256] running sum (ACCUM) or null
    void *df_out;                // [Ne, 256]
    int64_t n_nodes, n_edges;
    int64_t ldq, ldk, ldv, lddq;
    float scale, scale_log2, p_drop, inv_keep;
    uint64_t seed, offset;
    int relu_mask;               // multiply the final df by (feat > 0)
};

template <typename T, int HEADS, int STAGES, bool ACCUM>
__global__ void __launch_bounds__(EA_WARPS * 32)
edgeattn_bwd_dst_kernel(const EaBwdParams P) {
    using S = BwdSmem<T, HEADS, STAGES, ACCUM>;
    constexpr int RB = S::RB, EPS = S::EPS, LPH = 32 / HEADS, RPE = S::ROWS_PER_EDGE, NROW = 1 + 2 * HEADS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * S::PER_WARP;
    unsigned char *edge_ring = base;
    unsigned char *row_ring = base + S::EDGE_RING;
    uint64_t *full = reinterpret_cast<uint64_t *>(base + S::EDGE_RING + S::ROW_RING);
    int *eid_stash = reinterpret_cast<int *>(base + S::EDGE_RING + S::ROW_RING + 64);

    const T *q = (const T *)P.q, *k = (const T *)P.k, *v = (const T *)P.v;
    const T *qt = (const T *)P.qt, *gt = (const T *)P.gt, *feat = (const T *)P.feat;
    const T *df_in = (const T *)P.df_in;
    T *df_out = (T *)P.df_out, *dq = (T *)P.dq, *bbar = (T *)P.bbar;
    const int myhead = lane / LPH;
    const int ch = lane * 8;

    const int64_t W = (int64_t)gridDim.x * EA_WARPS, w = (int64_t)blockIdx.x * EA_WARPS + warp;
    const int64_t total = P.n_edges + (int64_t)ROW_KAPPA * P.n_nodes;
    const int64_t r0 = row_bound(P.rowptr, P.n_nodes, total * w / W);
    const int64_t r1 = row_bound(P.rowptr, P.n_nodes, total * (w + 1) / W);
    if (r0 >= r1) return;
    const int e_end = __ldg(P.rowptr + r1);

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        mbar_fence_init();
    }
    __syncwarp();

    StageCursor<EPS> prod, cons;
    prod.init(P.rowptr, r0, r1);
    cons = prod;
    int wbase = prod.pos;
    IndexWindow wcol, weid;
    wcol.init(P.col, wbase, e_end, lane);
    weid.init(P.eid, wbase, e_end, lane);

    auto issue = [&](int s) {
        if (prod.done()) return;
        const int n = prod.count();
        const bool first = prod.first;
        const int my_u = lane / RPE, my_r = lane - RPE * my_u;
        const int o = prod.pos + min(my_u, EPS - 1) - wbase;
        const int j = wcol.get(o), id = weid.get(o);
        if (lane == 0) mbar_expect_tx(full + s, (uint32_t)(n * RPE * RB + (first ? NROW * RB : 0)));
        __syncwarp();
        if (lane < RPE * EPS && my_u < n) {
            const T *src = my_r == 0 ? feat + (int64_t)id * EA_HIDDEN
                         : my_r == 1 ? k + (int64_t)j * P.ldk
                         : my_r == 2 ? v + (int64_t)j * P.ldv : df_in + (int64_t)id * EA_HIDDEN;
            bulk_g2s(edge_ring + ((size_t)(s * EPS + my_u) * RPE + my_r) * RB, src, RB, full + s);
            if (my_r == 0) eid_stash[s * EPS + my_u] = id;
        }
        if (first && lane >= 16 && lane < 16 + NROW) {
            const int t = lane - 16;
            const T *src = t == 0 ? q + prod.row * P.ldq
                         : t <= HEADS ? qt + ((int64_t)(t - 1) * P.n_nodes + prod.row) * EA_HIDDEN
                                      : gt + ((int64_t)(t - 1 - HEADS) * P.n_nodes + prod.row) * EA_HIDDEN;
            bulk_g2s(row_ring + ((size_t)s * NROW + t) * RB, src, RB, full + s);
        }
        prod.advance(P.rowptr);
        if (prod.pos - wbase >= 32 && !prod.done()) {
            wbase += 32;
            wcol.shift(P.col, wbase, e_end, lane);
            weid.shift(P.eid, wbase, e_end, lane);
        }
    };

    auto zero_rows = [&](int64_t lo, int64_t hi) {
        F8 zf;
#pragma unroll
        for (int c = 0; c < 8; ++c) zf.v[c] = 0.f;
        for (int64_t r = lo; r < hi; ++r) {
            st8(dq + r * P.lddq + ch, zf);
#pragma unroll
            for (int t = 0; t < HEADS; ++t) st8(bbar + ((int64_t)t * P.n_nodes + r) * EA_HIDDEN + ch, zf);
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    __syncwarp();

    F8 qf, gf, qtf[HEADS], gtf[HEADS], bb[HEADS], dqf;
    float D = 0.f, Gc = 0.f, m = 0.f, inv_z = 0.f;
    int64_t next_unwritten = r0;

    for (int it = 0; !cons.done(); ++it) {
        const int s = it % STAGES;
        issue((it + STAGES - 1) % STAGES);
        const int n = cons.count();
        const int64_t row = cons.row;
        if (cons.first) {
            // row vectors that are not TMA-staged: upstream gradient, saved aggregate, bias vector
            gf = ld8(P.dagg + row * EA_HIDDEN + ch);
            const F8 af = ld8(P.agg + row * EA_HIDDEN + ch);
            float dpart = 0.f, gpart = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) dpart = fmaf(gf.v[c], af.v[c], dpart);
            if (P.cvec) {
                const F8 cf = ld8(P.cvec + ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) gpart = fmaf(gf.v[c], cf.v[c], gpart);
            }
            D = group_sum<LPH>(dpart);
            Gc = group_sum<LPH>(gpart);
            m = __ldg(P.stat_m + row * HEADS + myhead);
            inv_z = 1.0f / (__ldg(P.stat_z + row * HEADS + myhead) + 1e-16f);
        }
        mbar_wait(full + s, (uint32_t)((it / STAGES) & 1));
        const unsigned char *stage = edge_ring + (size_t)s * EPS * RPE * RB;
        if (cons.first) {
            zero_rows(next_unwritten, row);
            const unsigned char *rs = row_ring + (size_t)s * NROW * RB;
            qf = lds8(reinterpret_cast<const T *>(rs) + ch);
#pragma unroll
            for (int t = 0; t < HEADS; ++t) {
                qtf[t] = lds8(reinterpret_cast<const T *>(rs + (size_t)(1 + t) * RB) + ch);
                gtf[t] = lds8(reinterpret_cast<const T *>(rs + (size_t)(1 + HEADS + t) * RB) + ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) bb[t].v[c] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) dqf.v[c] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < EPS; ++u) {
            if (u < n) {   // warp-uniform
                const T *frow = reinterpret_cast<const T *>(stage + (size_t)(u * RPE + 0) * RB) + ch;
                const F8 hf = lds8(frow);
                const F8 kf = lds8(reinterpret_cast<const T *>(stage + (size_t)(u * RPE + 1) * RB) + ch);
                const F8 vf = lds8(reinterpret_cast<const T *>(stage + (size_t)(u * RPE + 2) * RB) + ch);
                float p[HEADS], d[HEADS];
                float pk = 0.f, dv = 0.f;
#pragma unroll
                for (int t = 0; t < HEADS; ++t) p[t] = d[t] = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    pk = fmaf(qf.v[c], kf.v[c], pk);
                    dv = fmaf(gf.v[c], vf.v[c], dv);
#pragma unroll
                    for (int t = 0; t < HEADS; ++t) {
                        p[t] = fmaf(qtf[t].v[c], hf.v[c], p[t]);
                        d[t] = fmaf(gtf[t].v[c], hf.v[c], d[t]);
                    }
                }
#pragma unroll
                for (int t = 0; t < HEADS; ++t) {
                    p[t] += (t == myhead) ? pk : 0.f;
                    d[t] += (t == myhead) ? dv : 0.f;
                }
                const float sl = reduce_heads<HEADS>(p, lane);
                const float dat = reduce_heads<HEADS>(d, lane);
                const int id = eid_stash[s * EPS + u];
                float drop = 1.f;
                if (P.p_drop > 0.f)
                    drop = dropout_scale(P.seed, P.offset, (uint64_t)id * HEADS + myhead, P.p_drop, P.inv_keep);
                const float a = fast_exp2(sl * P.scale_log2 - m) * inv_z;
                const float at = a * drop;
                const float dss = a * ((dat + Gc) * drop - D) * P.scale;
                if (lane % LPH == 0) {
                    float *cf = P.coef + (int64_t)id * 2 * HEADS;
                    cf[myhead] = at;
                    cf[HEADS + myhead] = dss;
                }
                float at_h[HEADS], ds_h[HEADS];
#pragma unroll
                for (int t = 0; t < HEADS; ++t) {
                    at_h[t] = bcast_head<HEADS>(at, t);
                    ds_h[t] = bcast_head<HEADS>(dss, t);
                }
                F8 dff;
                if (ACCUM) dff = lds8(reinterpret_cast<const T *>(stage + (size_t)(u * RPE + (RPE - 1)) * RB) + ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    dqf.v[c] = fmaf(dss, kf.v[c], dqf.v[c]);
                    float acc = ACCUM ? dff.v[c] : 0.f;
#pragma unroll
                    for (int t = 0; t < HEADS; ++t) {
                        bb[t].v[c] = fmaf(ds_h[t], hf.v[c], bb[t].v[c]);
                        acc = fmaf(ds_h[t], qtf[t].v[c], acc);
                        acc = fmaf(at_h[t], gtf[t].v[c], acc);
                    }
                    if (P.relu_mask) acc = hf.v[c] > 0.f ? acc : 0.f;
                    dff.v[c] = acc;
                }
                st8(df_out + (int64_t)id * EA_HIDDEN + ch, dff);
            }
        }
        if (cons.last()) {
            st8(dq + row * P.lddq + ch, dqf);
#pragma unroll
            for (int t = 0; t < HEADS; ++t) st8(bbar + ((int64_t)t * P.n_nodes + row) * EA_HIDDEN + ch, bb[t]);
            next_unwritten = row + 1;
        }
        __syncwarp();
        cons.advance(P.rowptr);
    }
    zero_rows(next_unwritten, r1);
}

// ------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------
static int ea_grid(int64_t n_nodes, int64_t n_edges) {
    // static cost-balanced partition; a few CTAs per SM so that SM-to-SM variance evens out
    const int64_t work = n_edges + (int64_t)ROW_KAPPA * n_nodes;
    int64_t blocks = 148 * 8;
    const int64_t min_work_per_warp = 64;
    if (work / (blocks * EA_WARPS) < min_work_per_warp) blocks = work / (min_work_per_warp * EA_WARPS) + 1;
    return (int)blocks;
}

template <typename T, int HEADS>
static int launch_ea_fwd(const EaFwdParams &p, cudaStream_t st) {
    constexpr int STAGES = 3;
    using S = FwdSmem<T, HEADS, STAGES>;
    const size_t smem = (size_t)EA_WARPS * S::PER_WARP;
    auto kern = edgeattn_fwd_kernel<T, HEADS, STAGES>;
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ea_grid(p.n_nodes, p.n_edges), EA_WARPS * 32, smem, st>>>(p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

template <typename T, int HEADS, bool ACCUM>
static int launch_ea_bwd(const EaBwdParams &p, cudaStream_t st) {
    constexpr int STAGES = 2;
    using S = BwdSmem<T, HEADS, STAGES, ACCUM>;
    const size_t smem = (size_t)EA_WARPS * S::PER_WARP;
    auto kern = edgeattn_bwd_dst_kernel<T, HEADS, STAGES, ACCUM>;
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ea_grid(p.n_nodes, p.n_edges), EA_WARPS * 32, smem, st>>>(p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

template <typename T>
static int dispatch_ea_fwd(const EaFwdParams &p, int heads, cudaStream_t st) {
    switch (heads) {
        case 1: return launch_ea_fwd<T, 1>(p, st);
        case 2: return launch_ea_fwd<T, 2>(p, st);
        case 4: return launch_ea_fwd<T, 4>(p, st);
        default: return ALIGNN_ERR_BAD_SHAPE;
    }
}

template <typename T>
static int dispatch_ea_bwd(const EaBwdParams &p, int heads, bool accum, cudaStream_t st) {
#define EA_BWD(H) (accum ? launch_ea_bwd<T, H, true>(p, st) : launch_ea_bwd<T, H, false>(p, st))
    switch (heads) {
        case 1: return EA_BWD(1);
        case 2: return EA_BWD(2);
        case 4: return EA_BWD(4);
        default: return ALIGNN_ERR_BAD_SHAPE;
    }
#undef EA_BWD
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_edgeattn_supported(int hidden, int heads) {
    return hidden == EA_HIDDEN && (heads == 1 || heads == 2 || heads == 4);
}

extern "C" int alignn_edgeattn_fwd(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                                   const void *qt, const void *feat,
                                   const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                                   float *aggv, void *abar, float *stat_m, float *stat_z, float *stat_s,
                                   int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                   float p_drop, uint64_t seed, uint64_t offset, void *stream) {
    if (!alignn_edgeattn_supported(hidden, heads)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes < 0 || n_edges < 0 || !(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 64) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!q || !k || !v || !qt || !rowptr || !aggv || !abar || !stat_m || !stat_z || !stat_s) return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!feat || !col || !eid)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(qt) || !aligned16(feat) || !aligned16(aggv) ||
        !aligned16(abar) || (ldq % 8) || (ldk % 8) || (ldv % 8))
        return ALIGNN_ERR_BAD_ARG;
    EaFwdParams p;
    p.q = q; p.k = k; p.v = v; p.qt = qt; p.feat = feat;
    p.rowptr = rowptr; p.col = col; p.eid = eid;
    p.aggv = aggv; p.abar = abar; p.stat_m = stat_m; p.stat_z = stat_z; p.stat_s = stat_s;
    p.n_nodes = n_nodes; p.n_edges = n_edges; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
    p.scale_log2 = LOG2E / sqrtf((float)(hidden / heads));
    p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    p.seed = seed; p.offset = offset;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ALIGNN_F32) return dispatch_ea_fwd<float>(p, heads, st);
    if (dtype == ALIGNN_BF16) return dispatch_ea_fwd<__nv_bfloat16>(p, heads, st);
    return ALIGNN_ERR_BAD_DTYPE;
}

extern "C" int alignn_edgeattn_bwd_dst(const float *dagg, const float *agg,
                                       const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk,
                                       int64_t ldv, const void *qt, const void *gt, const float *cvec,
                                       const void *feat, const float *stat_m, const float *stat_z,
                                       const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                                       void *dq, int64_t lddq, void *bbar, float *coef,
                                       const void *df_in, void *df_out, int relu_mask,
                                       int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                       float p_drop, uint64_t seed, uint64_t offset, void *stream) {
    if (!alignn_edgeattn_supported(hidden, heads)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes < 0 || n_edges < 0 || !(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 64) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!dagg || !agg || !q || !k || !v || !qt || !gt || !stat_m || !stat_z || !rowptr || !dq || !bbar)
        return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!feat || !col || !eid || !coef || !df_out)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(dagg) || !aligned16(agg) || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(qt) ||
        !aligned16(gt) || !aligned16(cvec) || !aligned16(feat) || !aligned16(dq) || !aligned16(bbar) ||
        !aligned16(df_in) || !aligned16(df_out) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (lddq % 8))
        return ALIGNN_ERR_BAD_ARG;
    EaBwdParams p;
    p.dagg = dagg; p.agg = agg; p.q = q; p.k = k; p.v = v; p.qt = qt; p.gt = gt; p.cvec = cvec; p.feat = feat;
    p.stat_m = stat_m; p.stat_z = stat_z; p.rowptr = rowptr; p.col = col; p.eid = eid;
    p.dq = dq; p.bbar = bbar; p.coef = coef; p.df_in = df_in; p.df_out = df_out;
    p.n_nodes = n_nodes; p.n_edges = n_edges; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.lddq = lddq;
    p.scale = 1.0f / sqrtf((float)(hidden / heads));
    p.scale_log2 = p.scale * LOG2E;
    p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    p.seed = seed; p.offset = offset; p.relu_mask = relu_mask;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ALIGNN_F32) return dispatch_ea_bwd<float>(p, heads, df_in != nullptr, st);
    if (dtype == ALIGNN_BF16) return dispatch_ea_bwd<__nv_bfloat16>(p, heads, df_in != nullptr, st);
    return ALIGNN_ERR_BAD_DTYPE;
}
