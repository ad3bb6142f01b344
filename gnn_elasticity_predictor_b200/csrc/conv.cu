// Fused edge-attention conv core (forward + backward) for sm_100a.
//
// Replaces TransformerConv.message + utils.softmax + 'add' aggregation of PyG 2.7.0 as called at
// reference scripts/train.py:315 (line graph) and :334 (atom graph); math in SURVEY.md Appendix A.
//
// Layout / mapping
//   * edges are visited in stable target-sorted (CSR) order; a "row group" of LANES = hidden/8 lanes
//     owns one target row, each lane 8 consecutive channels (one 16-byte load for bf16, two for
//     fp32), so a gathered row is read as one fully coalesced 128-bit-per-lane request;
//   * per-head logits are reduced with xor-shuffles over the C/8 lanes of a head; the segment
//     softmax is ONLINE (running max / denominator in registers), so each k/v/e row is read
//     exactly once and nothing of size [edges, hidden] is ever written in forward;
//   * the aggregate is a register accumulation in edge order: no atomics, bit-reproducible;
//   * backward: a target-sorted pass (dq, de, per-edge coefficients) and a source-sorted (CSC) pass
//     (dk, dv) -- again plain segmented sums.
//   HBM-bound: algorithmic bytes per conv in DESIGN.md / SURVEY.md section 8(d).
//
// A generic one-warp-per-row kernel family covers every (hidden, heads) the fast mapping cannot
// (hidden/8 or C/8 not a power of two, hidden > 256): same math, shared-memory row state.
#include <math.h>

#include "common.cuh"

namespace alignn {

constexpr int CONV_THREADS = 256;
constexpr int FWD_UNROLL = 4;
constexpr int BWD_UNROLL = 2;

template <int LANES>
__device__ __forceinline__ int warp_max_int(int x) {
    if (LANES < 32) {
#pragma unroll
        for (int off = 16; off >= LANES; off >>= 1) x = max(x, __shfl_xor_sync(FULL, x, off));
    }
    return x;
}

template <int LPH_T>
__device__ __forceinline__ float head_sum(float x, int lph_rt) {
    if (LPH_T > 0) return group_sum<(LPH_T > 0 ? LPH_T : 1)>(x);
    return group_sum_rt(x, lph_rt);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int LANES, int LPH_T>
__global__ void __launch_bounds__(CONV_THREADS)
conv_fwd_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                const T *__restrict__ e, const int32_t *__restrict__ rowptr,
                const int32_t *__restrict__ col, const int32_t *__restrict__ eid,
                float *__restrict__ agg, float *__restrict__ stat_m, float *__restrict__ stat_z,
                int64_t n_nodes, int hidden, int heads, int lph_rt, float scale_log2,
                float p_drop, float inv_keep, uint64_t seed, uint64_t offset) {
    constexpr int RPW = 32 / LANES;
    const int lph = LPH_T > 0 ? LPH_T : lph_rt;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t row = warp_id * RPW + lane / LANES;
    const bool row_ok = row < n_nodes;
    const int head = sub / lph;
    const int ch = sub * 8;

    int beg = 0, end = 0;
    if (row_ok) {
        beg = __ldg(rowptr + row);
        end = __ldg(rowptr + row + 1);
    }
    const int max_deg = warp_max_int<LANES>(end - beg);

    F8 qf;
#pragma unroll
    for (int c = 0; c < 8; ++c) qf.v[c] = 0.f;
    if (row_ok) {
        qf = ld8(q + row * hidden + ch);
#pragma unroll
        for (int c = 0; c < 8; ++c) qf.v[c] *= scale_log2;
    }

    float m = -INFINITY, z = 0.f;
    F8 acc;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc.v[c] = 0.f;

    for (int it = 0; it < max_deg; it += FWD_UNROLL) {
        F8 uf[FWD_UNROLL];
        float s[FWD_UNROLL];
        float drop[FWD_UNROLL];
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u) {
            const int p = beg + it + u;
            const bool act = p < end;
            float part = 0.f;
            drop[u] = 1.f;
            if (act) {
                const int j = __ldg(col + p);
                const int id = __ldg(eid + p);
                const F8 kf = ld8(k + (int64_t)j * hidden + ch);
                const F8 vf = ld8(v + (int64_t)j * hidden + ch);
                const F8 ef = ld8_stream(e + (int64_t)id * hidden + ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    part = fmaf(qf.v[c], kf.v[c] + ef.v[c], part);
                    uf[u].v[c] = vf.v[c] + ef.v[c];
                }
                if (p_drop > 0.f)
                    drop[u] = dropout_scale(seed, offset, (uint64_t)id * heads + head, p_drop, inv_keep);
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) uf[u].v[c] = 0.f;
            }
            s[u] = part;
        }
        float m_new = m;
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u) {
            s[u] = head_sum<LPH_T>(s[u], lph);
            if (beg + it + u >= end) s[u] = -INFINITY;
            m_new = fmaxf(m_new, s[u]);
        }
        const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
        const float corr = fast_exp2(m - m_safe);  // m = -inf -> 0
        z *= corr;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc.v[c] *= corr;
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u) {
            const float w = fast_exp2(s[u] - m_safe);  // inactive -> 0
            z += w;
            const float wd = w * drop[u];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc.v[c] = fmaf(wd, uf[u].v[c], acc.v[c]);
        }
        m = m_new;
    }

    if (row_ok) {
        const float inv = 1.0f / (z + 1e-16f);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc.v[c] *= inv;
        st8(agg + row * hidden + ch, acc);
        if (sub % lph == 0) {
            stat_m[row * heads + head] = (m == -INFINITY) ? 0.f : m;
            stat_z[row * heads + head] = z;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, target-sorted pass: dq, de, coef = (a~, ds/sqrt(C)) per (edge, head)
// ------------------------------------------------------------------------------------------------
template <typename T, int LANES, int LPH_T>
__global__ void __launch_bounds__(CONV_THREADS)
conv_bwd_dst_kernel(const float *__restrict__ dagg, const float *__restrict__ agg,
                    const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                    const T *__restrict__ e, const float *__restrict__ stat_m,
                    const float *__restrict__ stat_z, const int32_t *__restrict__ rowptr,
                    const int32_t *__restrict__ col, const int32_t *__restrict__ eid,
                    T *__restrict__ dq, T *__restrict__ de, float *__restrict__ coef,
                    int64_t n_nodes, int hidden, int heads, int lph_rt, float scale, float scale_log2,
                    float p_drop, float inv_keep, uint64_t seed, uint64_t offset) {
    constexpr int RPW = 32 / LANES;
    const int lph = LPH_T > 0 ? LPH_T : lph_rt;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t row = warp_id * RPW + lane / LANES;
    const bool row_ok = row < n_nodes;
    const int head = sub / lph;
    const int ch = sub * 8;

    int beg = 0, end = 0;
    if (row_ok) {
        beg = __ldg(rowptr + row);
        end = __ldg(rowptr + row + 1);
    }
    const int max_deg = warp_max_int<LANES>(end - beg);

    F8 qf, gf, dqf;
    float d_part = 0.f, m = 0.f, inv_z = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) qf.v[c] = gf.v[c] = dqf.v[c] = 0.f;
    if (row_ok) {
        qf = ld8(q + row * hidden + ch);
        gf = ld8(dagg + row * hidden + ch);
        const F8 af = ld8(agg + row * hidden + ch);
#pragma unroll
        for (int c = 0; c < 8; ++c) d_part = fmaf(gf.v[c], af.v[c], d_part);
        m = __ldg(stat_m + row * heads + head);
        inv_z = 1.0f / (__ldg(stat_z + row * heads + head) + 1e-16f);
    }
    const float D = head_sum<LPH_T>(d_part, lph);

    for (int it = 0; it < max_deg; it += BWD_UNROLL) {
        F8 kef[BWD_UNROLL], uf[BWD_UNROLL];
        float s[BWD_UNROLL], dat[BWD_UNROLL], drop[BWD_UNROLL];
        int ids[BWD_UNROLL];
#pragma unroll
        for (int u = 0; u < BWD_UNROLL; ++u) {
            const int p = beg + it + u;
            const bool act = p < end;
            float sp = 0.f, dp = 0.f;
            drop[u] = 1.f;
            ids[u] = 0;
            if (act) {
                const int j = __ldg(col + p);
                const int id = __ldg(eid + p);
                ids[u] = id;
                const F8 kf = ld8(k + (int64_t)j * hidden + ch);
                const F8 vf = ld8(v + (int64_t)j * hidden + ch);
                const F8 ef = ld8_stream(e + (int64_t)id * hidden + ch);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    kef[u].v[c] = kf.v[c] + ef.v[c];
                    uf[u].v[c] = vf.v[c] + ef.v[c];
                    sp = fmaf(qf.v[c], kef[u].v[c], sp);
                    dp = fmaf(gf.v[c], uf[u].v[c], dp);
                }
                if (p_drop > 0.f)
                    drop[u] = dropout_scale(seed, offset, (uint64_t)id * heads + head, p_drop, inv_keep);
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) kef[u].v[c] = uf[u].v[c] = 0.f;
            }
            s[u] = sp;
            dat[u] = dp;
        }
#pragma unroll
        for (int u = 0; u < BWD_UNROLL; ++u) {
            s[u] = head_sum<LPH_T>(s[u], lph);
            dat[u] = head_sum<LPH_T>(dat[u], lph);
        }
#pragma unroll
        for (int u = 0; u < BWD_UNROLL; ++u) {
            const bool act = beg + it + u < end;
            if (act) {
                const float a = fast_exp2(s[u] * scale_log2 - m) * inv_z;
                const float at = a * drop[u];
                const float ds = a * (dat[u] * drop[u] - D);
                const float dss = ds * scale;
                F8 def;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    dqf.v[c] = fmaf(dss, kef[u].v[c], dqf.v[c]);
                    def.v[c] = fmaf(dss, qf.v[c], at * gf.v[c]);
                }
                st8(de + (int64_t)ids[u] * hidden + ch, def);
                if (sub % lph == 0) {
                    float *cf = coef + (int64_t)ids[u] * 2 * heads;
                    cf[head] = at;
                    cf[heads + head] = dss;
                }
            }
        }
    }
    if (row_ok) st8(dq + row * hidden + ch, dqf);
}

// ------------------------------------------------------------------------------------------------
// backward, source-sorted pass: dk_j = sum_i dss_ij q_i ; dv_j = sum_i a~_ij dagg_i
// ------------------------------------------------------------------------------------------------
template <typename T, int LANES, typename G = float>
__global__ void __launch_bounds__(CONV_THREADS)
conv_bwd_src_kernel(const G *__restrict__ dagg, const T *__restrict__ q, const float *__restrict__ coef,
                    const int32_t *__restrict__ rowptr_t, const int32_t *__restrict__ col_t,
                    const int32_t *__restrict__ eid_t, T *__restrict__ dk, T *__restrict__ dv,
                    int64_t n_nodes, int hidden, int heads, int lph, int64_t ldq, int64_t ldd) {
    constexpr int RPW = 32 / LANES;
    constexpr int U = 8;                       // edges in flight per lane: the pass is bound by L2 gather bandwidth
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t row = warp_id * RPW + lane / LANES;
    if (row >= n_nodes) return;  // no shuffles below: lanes may retire independently
    const int head = sub / lph;
    const int ch = sub * 8;
    const int beg = __ldg(rowptr_t + row), end = __ldg(rowptr_t + row + 1);
    // gather addresses as base + u32 index * u32 byte stride: ONE IMAD.WIDE.U32 per address (the 64-bit index arithmetic
    // of `ptr + (int64)i * ld` was 21 of the 61 instructions per edge, and the pass issues at 60 % of peak)
    const char *dagg_b = reinterpret_cast<const char *>(dagg + ch);
    const char *q_b = reinterpret_cast<const char *>(q + ch);
    const char *at_b = reinterpret_cast<const char *>(coef + head);
    const char *ds_b = reinterpret_cast<const char *>(coef + heads + head);
    const uint32_t dagg_ld = (uint32_t)hidden * (uint32_t)sizeof(G);
    const uint32_t q_ld = (uint32_t)ldq * (uint32_t)sizeof(T);
    const uint32_t coef_ld = 2u * (uint32_t)heads * (uint32_t)sizeof(float);
    auto at_row = [](const char *base, uint32_t idx, uint32_t ld) { return base + (uint64_t)idx * ld; };
    F8 dkf, dvf;
#pragma unroll
    for (int c = 0; c < 8; ++c) dkf.v[c] = dvf.v[c] = 0.f;
    int p = beg;
    for (; p + U <= end; p += U) {
        uint32_t i[U], id[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            i[u] = (uint32_t)__ldg(col_t + p + u);
            id[u] = (uint32_t)__ldg(eid_t + p + u);
        }
        F8 g[U], qq[U];
        float at[U], ds[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            g[u] = ld8(reinterpret_cast<const G *>(at_row(dagg_b, i[u], dagg_ld)));
            qq[u] = ld8(reinterpret_cast<const T *>(at_row(q_b, i[u], q_ld)));
            at[u] = __ldg(reinterpret_cast<const float *>(at_row(at_b, id[u], coef_ld)));
            ds[u] = __ldg(reinterpret_cast<const float *>(at_row(ds_b, id[u], coef_ld)));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {        // accumulation in edge order: same sums as the one-at-a-time loop
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                dvf.v[c] = fmaf(at[u], g[u].v[c], dvf.v[c]);
                dkf.v[c] = fmaf(ds[u], qq[u].v[c], dkf.v[c]);
            }
        }
    }
    for (; p < end; ++p) {
        const uint32_t i0 = (uint32_t)__ldg(col_t + p);
        const uint32_t id0 = (uint32_t)__ldg(eid_t + p);
        const F8 g0 = ld8(reinterpret_cast<const G *>(at_row(dagg_b, i0, dagg_ld)));
        const F8 q0 = ld8(reinterpret_cast<const T *>(at_row(q_b, i0, q_ld)));
        const float at0 = __ldg(reinterpret_cast<const float *>(at_row(at_b, id0, coef_ld)));
        const float ds0 = __ldg(reinterpret_cast<const float *>(at_row(ds_b, id0, coef_ld)));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            dvf.v[c] = fmaf(at0, g0.v[c], dvf.v[c]);
            dkf.v[c] = fmaf(ds0, q0.v[c], dkf.v[c]);
        }
    }
    st8(dk + row * ldd + ch, dkf);
    st8(dv + row * ldd + ch, dvf);
}

// ------------------------------------------------------------------------------------------------
// generic family: one warp per row, lanes stride the channels of each head, row state in smem.
// Correctness path for shapes outside the fast mapping; same math, same saved statistics.
// ------------------------------------------------------------------------------------------------
constexpr int GEN_WARPS = 4;

template <typename T>
__global__ void __launch_bounds__(GEN_WARPS * 32)
conv_fwd_generic_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                        const T *__restrict__ e, const int32_t *__restrict__ rowptr,
                        const int32_t *__restrict__ col, const int32_t *__restrict__ eid,
                        float *__restrict__ agg, float *__restrict__ stat_m, float *__restrict__ stat_z,
                        int64_t n_nodes, int hidden, int heads, float scale_log2,
                        float p_drop, float inv_keep, uint64_t seed, uint64_t offset) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = 2 * hidden + 2 * heads;
    float *qs = smem + warp * per_warp, *acc = qs + hidden, *ms = acc + hidden, *zs = ms + heads;
    const int64_t row = (int64_t)blockIdx.x * GEN_WARPS + warp;
    if (row >= n_nodes) return;
    const int C = hidden / heads;
    for (int c = lane; c < hidden; c += 32) {
        qs[c] = ldf(q + row * hidden + c) * scale_log2;
        acc[c] = 0.f;
    }
    for (int t = lane; t < heads; t += 32) {
        ms[t] = -INFINITY;
        zs[t] = 0.f;
    }
    __syncwarp();
    const int beg = rowptr[row], end = rowptr[row + 1];
    for (int p = beg; p < end; ++p) {
        const int64_t j = col[p], id = eid[p];
        for (int t = 0; t < heads; ++t) {
            float part = 0.f;
            for (int c = t * C + lane; c < (t + 1) * C; c += 32)
                part = fmaf(qs[c], ldf(k + j * hidden + c) + ldf(e + id * hidden + c), part);
            const float s = warp_sum(part);
            const float m_old = ms[t];
            const float m_new = fmaxf(m_old, s);
            const float corr = fast_exp2(m_old - m_new);
            const float w = fast_exp2(s - m_new);
            const float z_new = zs[t] * corr + w;
            float drop = 1.f;
            if (p_drop > 0.f) drop = dropout_scale(seed, offset, (uint64_t)id * heads + t, p_drop, inv_keep);
            __syncwarp();
            if (lane == 0) {
                ms[t] = m_new;
                zs[t] = z_new;
            }
            const float wd = w * drop;
            for (int c = t * C + lane; c < (t + 1) * C; c += 32)
                acc[c] = fmaf(wd, ldf(v + j * hidden + c) + ldf(e + id * hidden + c), acc[c] * corr);
            __syncwarp();
        }
    }
    for (int c = lane; c < hidden; c += 32) agg[row * hidden + c] = acc[c] / (zs[c / C] + 1e-16f);
    for (int t = lane; t < heads; t += 32) {
        stat_m[row * heads + t] = (ms[t] == -INFINITY) ? 0.f : ms[t];
        stat_z[row * heads + t] = zs[t];
    }
}

template <typename T>
__global__ void __launch_bounds__(GEN_WARPS * 32)
conv_bwd_dst_generic_kernel(const float *__restrict__ dagg, const float *__restrict__ agg,
                            const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                            const T *__restrict__ e, const float *__restrict__ stat_m,
                            const float *__restrict__ stat_z, const int32_t *__restrict__ rowptr,
                            const int32_t *__restrict__ col, const int32_t *__restrict__ eid,
                            T *__restrict__ dq, T *__restrict__ de, float *__restrict__ coef,
                            int64_t n_nodes, int hidden, int heads, float scale, float scale_log2,
                            float p_drop, float inv_keep, uint64_t seed, uint64_t offset) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = 3 * hidden;
    float *qs = smem + warp * per_warp, *gs = qs + hidden, *dqs = gs + hidden;
    const int64_t row = (int64_t)blockIdx.x * GEN_WARPS + warp;
    if (row >= n_nodes) return;
    const int C = hidden / heads;
    for (int c = lane; c < hidden; c += 32) {
        qs[c] = ldf(q + row * hidden + c);
        gs[c] = dagg[row * hidden + c];
        dqs[c] = 0.f;
    }
    __syncwarp();
    const int beg = rowptr[row], end = rowptr[row + 1];
    for (int t = 0; t < heads; ++t) {
        float dpart = 0.f;
        for (int c = t * C + lane; c < (t + 1) * C; c += 32) dpart = fmaf(gs[c], agg[row * hidden + c], dpart);
        const float D = warp_sum(dpart);
        const float m = stat_m[row * heads + t];
        const float inv_z = 1.0f / (stat_z[row * heads + t] + 1e-16f);
        for (int p = beg; p < end; ++p) {
            const int64_t j = col[p], id = eid[p];
            float sp = 0.f, dp = 0.f;
            for (int c = t * C + lane; c < (t + 1) * C; c += 32) {
                const float ev = ldf(e + id * hidden + c);
                sp = fmaf(qs[c], ldf(k + j * hidden + c) + ev, sp);
                dp = fmaf(gs[c], ldf(v + j * hidden + c) + ev, dp);
            }
            const float s = warp_sum(sp), dat = warp_sum(dp);
            float drop = 1.f;
            if (p_drop > 0.f) drop = dropout_scale(seed, offset, (uint64_t)id * heads + t, p_drop, inv_keep);
            const float a = fast_exp2(s * scale_log2 - m) * inv_z;
            const float at = a * drop;
            const float dss = a * (dat * drop - D) * scale;
            for (int c = t * C + lane; c < (t + 1) * C; c += 32) {
                dqs[c] = fmaf(dss, ldf(k + j * hidden + c) + ldf(e + id * hidden + c), dqs[c]);
                stf(de + id * hidden + c, fmaf(dss, qs[c], at * gs[c]));
            }
            if (lane == 0) {
                coef[id * 2 * heads + t] = at;
                coef[id * 2 * heads + heads + t] = dss;
            }
        }
    }
    __syncwarp();
    for (int c = lane; c < hidden; c += 32) stf(dq + row * hidden + c, dqs[c]);
}

template <typename T>
__global__ void __launch_bounds__(GEN_WARPS * 32)
conv_bwd_src_generic_kernel(const float *__restrict__ dagg, const T *__restrict__ q,
                            const float *__restrict__ coef, const int32_t *__restrict__ rowptr_t,
                            const int32_t *__restrict__ col_t, const int32_t *__restrict__ eid_t,
                            T *__restrict__ dk, T *__restrict__ dv, int64_t n_nodes, int hidden, int heads) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * GEN_WARPS + warp;
    if (row >= n_nodes) return;
    const int C = hidden / heads;
    const int beg = rowptr_t[row], end = rowptr_t[row + 1];
    for (int c = lane; c < hidden; c += 32) {  // each lane owns its channels: edge-order register sums
        const int t = c / C;
        float dkc = 0.f, dvc = 0.f;
        for (int p = beg; p < end; ++p) {
            const int64_t i = col_t[p], id = eid_t[p];
            dvc = fmaf(coef[id * 2 * heads + t], dagg[i * hidden + c], dvc);
            dkc = fmaf(coef[id * 2 * heads + heads + t], ldf(q + i * hidden + c), dkc);
        }
        stf(dk + row * hidden + c, dkc);
        stf(dv + row * hidden + c, dvc);
    }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
static inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// fast mapping available?  LANES = hidden/8 in {1..32} power of two, C % 8 == 0, C/8 power of two
static inline bool fast_shape(int hidden, int heads, int *lanes, int *lph) {
    if (hidden % 8) return false;
    const int L = hidden / 8, C = hidden / heads;
    if (L > 32 || !is_pow2(L) || C % 8 || !is_pow2(C / 8)) return false;
    *lanes = L;
    *lph = C / 8;
    return true;
}

static inline unsigned fast_grid(int64_t n_nodes, int lanes) {
    const int rows_per_block = (CONV_THREADS / 32) * (32 / lanes);
    return (unsigned)((n_nodes + rows_per_block - 1) / rows_per_block);
}

struct FwdArgs {
    const void *q, *k, *v, *e;
    const int32_t *rowptr, *col, *eid;
    float *agg, *stat_m, *stat_z;
    int64_t n_nodes;
    int hidden, heads;
    float p_drop;
    uint64_t seed, offset;
    cudaStream_t st;
};

template <typename T, int LANES, int LPH_T>
static void launch_fwd(const FwdArgs &a, int lph) {
    const float scale_log2 = LOG2E / sqrtf((float)(a.hidden / a.heads));
    const float inv_keep = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
    conv_fwd_kernel<T, LANES, LPH_T><<<fast_grid(a.n_nodes, LANES), CONV_THREADS, 0, a.st>>>(
        (const T *)a.q, (const T *)a.k, (const T *)a.v, (const T *)a.e, a.rowptr, a.col, a.eid, a.agg,
        a.stat_m, a.stat_z, a.n_nodes, a.hidden, a.heads, lph, scale_log2, a.p_drop, inv_keep, a.seed, a.offset);
}

template <typename T>
static int dispatch_fwd(const FwdArgs &a) {
    int lanes = 0, lph = 0;
    if (fast_shape(a.hidden, a.heads, &lanes, &lph)) {
        if (lanes == 32 && lph == 8) launch_fwd<T, 32, 8>(a, lph);
        else if (lanes == 32) launch_fwd<T, 32, 0>(a, lph);
        else if (lanes == 16) launch_fwd<T, 16, 0>(a, lph);
        else if (lanes == 8) launch_fwd<T, 8, 0>(a, lph);
        else if (lanes == 4) launch_fwd<T, 4, 0>(a, lph);
        else if (lanes == 2) launch_fwd<T, 2, 0>(a, lph);
        else launch_fwd<T, 1, 0>(a, lph);
    } else {
        const float scale_log2 = LOG2E / sqrtf((float)(a.hidden / a.heads));
        const float inv_keep = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
        const size_t smem = (size_t)GEN_WARPS * (2 * a.hidden + 2 * a.heads) * sizeof(float);
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(conv_fwd_generic_kernel<T>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return ALIGNN_ERR_CUDA_BASE + (int)err;
        }
        conv_fwd_generic_kernel<T><<<(unsigned)((a.n_nodes + GEN_WARPS - 1) / GEN_WARPS), GEN_WARPS * 32, smem, a.st>>>(
            (const T *)a.q, (const T *)a.k, (const T *)a.v, (const T *)a.e, a.rowptr, a.col, a.eid, a.agg,
            a.stat_m, a.stat_z, a.n_nodes, a.hidden, a.heads, scale_log2, a.p_drop, inv_keep, a.seed, a.offset);
    }
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

struct BwdArgs {
    const float *dagg, *agg;
    const void *q, *k, *v, *e;
    const float *stat_m, *stat_z;
    const int32_t *rowptr, *col, *eid, *rowptr_t, *col_t, *eid_t;
    void *dq, *dk, *dv, *de;
    float *coef;
    int64_t n_nodes;
    int hidden, heads;
    float p_drop;
    uint64_t seed, offset;
    cudaStream_t st;
};

template <typename T, int LANES, int LPH_T>
static void launch_bwd(const BwdArgs &a, int lph) {
    const float scale = 1.0f / sqrtf((float)(a.hidden / a.heads));
    const float inv_keep = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
    const unsigned grid = fast_grid(a.n_nodes, LANES);
    conv_bwd_dst_kernel<T, LANES, LPH_T><<<grid, CONV_THREADS, 0, a.st>>>(
        a.dagg, a.agg, (const T *)a.q, (const T *)a.k, (const T *)a.v, (const T *)a.e, a.stat_m, a.stat_z,
        a.rowptr, a.col, a.eid, (T *)a.dq, (T *)a.de, a.coef, a.n_nodes, a.hidden, a.heads, lph, scale,
        scale * LOG2E, a.p_drop, inv_keep, a.seed, a.offset);
    conv_bwd_src_kernel<T, LANES><<<grid, CONV_THREADS, 0, a.st>>>(
        a.dagg, (const T *)a.q, a.coef, a.rowptr_t, a.col_t, a.eid_t, (T *)a.dk, (T *)a.dv, a.n_nodes,
        a.hidden, a.heads, lph, (int64_t)a.hidden, (int64_t)a.hidden);
}

template <typename T>
static int dispatch_bwd(const BwdArgs &a) {
    int lanes = 0, lph = 0;
    if (fast_shape(a.hidden, a.heads, &lanes, &lph)) {
        if (lanes == 32 && lph == 8) launch_bwd<T, 32, 8>(a, lph);
        else if (lanes == 32) launch_bwd<T, 32, 0>(a, lph);
        else if (lanes == 16) launch_bwd<T, 16, 0>(a, lph);
        else if (lanes == 8) launch_bwd<T, 8, 0>(a, lph);
        else if (lanes == 4) launch_bwd<T, 4, 0>(a, lph);
        else if (lanes == 2) launch_bwd<T, 2, 0>(a, lph);
        else launch_bwd<T, 1, 0>(a, lph);
    } else {
        const float scale = 1.0f / sqrtf((float)(a.hidden / a.heads));
        const float inv_keep = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
        const size_t smem = (size_t)GEN_WARPS * 3 * a.hidden * sizeof(float);
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(conv_bwd_dst_generic_kernel<T>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return ALIGNN_ERR_CUDA_BASE + (int)err;
        }
        const unsigned grid = (unsigned)((a.n_nodes + GEN_WARPS - 1) / GEN_WARPS);
        conv_bwd_dst_generic_kernel<T><<<grid, GEN_WARPS * 32, smem, a.st>>>(
            a.dagg, a.agg, (const T *)a.q, (const T *)a.k, (const T *)a.v, (const T *)a.e, a.stat_m, a.stat_z,
            a.rowptr, a.col, a.eid, (T *)a.dq, (T *)a.de, a.coef, a.n_nodes, a.hidden, a.heads, scale,
            scale * LOG2E, a.p_drop, inv_keep, a.seed, a.offset);
        conv_bwd_src_generic_kernel<T><<<grid, GEN_WARPS * 32, 0, a.st>>>(
            a.dagg, (const T *)a.q, a.coef, a.rowptr_t, a.col_t, a.eid_t, (T *)a.dk, (T *)a.dv, a.n_nodes,
            a.hidden, a.heads);
    }
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

static int check_shape(int64_t n_nodes, int64_t n_edges, int hidden, int heads, float p_drop) {
    if (n_nodes < 0 || n_edges < 0 || hidden <= 0 || heads <= 0) return ALIGNN_ERR_BAD_ARG;
    if (hidden % heads) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 1) return ALIGNN_ERR_BAD_SHAPE;
    if (!(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (hidden > 4096) return ALIGNN_ERR_BAD_SHAPE;  // generic path keeps 3*hidden floats per warp in smem
    return ALIGNN_OK;
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_conv_fwd(const void *q, const void *k, const void *v, const void *e,
                               const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                               float *agg, float *stat_m, float *stat_z,
                               int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                               float p_drop, uint64_t seed, uint64_t offset, void *stream) {
    int rc = check_shape(n_nodes, n_edges, hidden, heads, p_drop);
    if (rc != ALIGNN_OK) return rc;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!q || !k || !v || !rowptr || !agg || !stat_m || !stat_z) return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!e || !col || !eid)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(e) || !aligned16(agg)) return ALIGNN_ERR_BAD_ARG;
    FwdArgs a{q, k, v, e, rowptr, col, eid, agg, stat_m, stat_z, n_nodes, hidden, heads, p_drop, seed, offset,
              reinterpret_cast<cudaStream_t>(stream)};
    if (dtype == ALIGNN_F32) return dispatch_fwd<float>(a);
    if (dtype == ALIGNN_BF16) return dispatch_fwd<__nv_bfloat16>(a);
    return ALIGNN_ERR_BAD_DTYPE;
}

extern "C" int alignn_conv_bwd(const float *dagg, const float *agg,
                               const void *q, const void *k, const void *v, const void *e,
                               const float *stat_m, const float *stat_z,
                               const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                               const int32_t *rowptr_t, const int32_t *col_t, const int32_t *eid_t,
                               void *dq, void *dk, void *dv, void *de, float *coef,
                               int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                               float p_drop, uint64_t seed, uint64_t offset, void *stream) {
    int rc = check_shape(n_nodes, n_edges, hidden, heads, p_drop);
    if (rc != ALIGNN_OK) return rc;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!dagg || !agg || !q || !k || !v || !stat_m || !stat_z || !rowptr || !rowptr_t || !dq || !dk || !dv)
        return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!e || !col || !eid || !col_t || !eid_t || !de || !coef)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(dagg) || !aligned16(agg) || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(e) ||
        !aligned16(dq) || !aligned16(dk) || !aligned16(dv) || !aligned16(de))
        return ALIGNN_ERR_BAD_ARG;
    BwdArgs a{dagg, agg, q, k, v, e, stat_m, stat_z, rowptr, col, eid, rowptr_t, col_t, eid_t, dq, dk, dv, de,
              coef, n_nodes, hidden, heads, p_drop, seed, offset, reinterpret_cast<cudaStream_t>(stream)};
    if (dtype == ALIGNN_F32) return dispatch_bwd<float>(a);
    if (dtype == ALIGNN_BF16) return dispatch_bwd<__nv_bfloat16>(a);
    return ALIGNN_ERR_BAD_DTYPE;
}

// same pass with the upstream gradient in STORAGE dtype (bf16): the pass is bound by L2 gather bandwidth
// (1.5 KB per edge with an fp32 dagg row, 1 KB with a bf16 one)
extern "C" int alignn_edgeattn_bwd_src_lp(const void *dagg_lp, const void *q, int64_t ldq, const float *coef,
                                          const int32_t *rowptr_t, const int32_t *col_t, const int32_t *eid_t,
                                          void *dk, void *dv, int64_t ldd, int64_t n_nodes, int64_t n_edges,
                                          int hidden, int heads, int dtype, void *stream) {
    int rc = check_shape(n_nodes, n_edges, hidden, heads, 0.f);
    if (rc != ALIGNN_OK) return rc;
    int lanes = 0, lph = 0;
    if (!fast_shape(hidden, heads, &lanes, &lph) || lanes != 32) return ALIGNN_ERR_BAD_SHAPE;
    if (dtype != ALIGNN_BF16) return ALIGNN_ERR_BAD_DTYPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!dagg_lp || !q || !rowptr_t || !dk || !dv || (n_edges > 0 && (!coef || !col_t || !eid_t))) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(dagg_lp) || !aligned16(q) || !aligned16(dk) || !aligned16(dv) || (ldq % 8) || (ldd % 8)) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    conv_bwd_src_kernel<__nv_bfloat16, 32, __nv_bfloat16><<<fast_grid(n_nodes, 32), CONV_THREADS, 0, st>>>(
        (const __nv_bfloat16 *)dagg_lp, (const __nv_bfloat16 *)q, coef, rowptr_t, col_t, eid_t, (__nv_bfloat16 *)dk,
        (__nv_bfloat16 *)dv, n_nodes, hidden, heads, lph, ldq, ldd);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

// source-sorted pass on strided operands (q, dk, dv are column slices of [N, 4H] projection buffers)
extern "C" int alignn_edgeattn_bwd_src(const float *dagg, const void *q, int64_t ldq, const float *coef,
                                       const int32_t *rowptr_t, const int32_t *col_t, const int32_t *eid_t,
                                       void *dk, void *dv, int64_t ldd, int64_t n_nodes, int64_t n_edges,
                                       int hidden, int heads, int dtype, void *stream) {
    int rc = check_shape(n_nodes, n_edges, hidden, heads, 0.f);
    if (rc != ALIGNN_OK) return rc;
    int lanes = 0, lph = 0;
    if (!fast_shape(hidden, heads, &lanes, &lph) || lanes != 32) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!dagg || !q || !rowptr_t || !dk || !dv || (n_edges > 0 && (!coef || !col_t || !eid_t))) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(dagg) || !aligned16(q) || !aligned16(dk) || !aligned16(dv) || (ldq % 8) || (ldd % 8)) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned grid = fast_grid(n_nodes, 32);
    if (dtype == ALIGNN_F32)
        conv_bwd_src_kernel<float, 32><<<grid, CONV_THREADS, 0, st>>>(dagg, (const float *)q, coef, rowptr_t, col_t, eid_t,
                                                                     (float *)dk, (float *)dv, n_nodes, hidden, heads, lph, ldq, ldd);
    else if (dtype == ALIGNN_BF16)
        conv_bwd_src_kernel<__nv_bfloat16, 32><<<grid, CONV_THREADS, 0, st>>>(
            dagg, (const __nv_bfloat16 *)q, coef, rowptr_t, col_t, eid_t, (__nv_bfloat16 *)dk, (__nv_bfloat16 *)dv, n_nodes,
            hidden, heads, lph, ldq, ldd);
    else
        return ALIGNN_ERR_BAD_DTYPE;
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
