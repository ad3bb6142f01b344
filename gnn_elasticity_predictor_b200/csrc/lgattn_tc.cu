// Line-graph edge attention, forward, on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as lgattn_fwd_kernel (lgattn.cu) -- reference path: `angle_encoder` first layer + line-graph
// `TransformerConv.message` + `utils.softmax` + aggregation (scripts/train.py:360-364, 554, 308, 315; PyG 2.7.0) with the
// per-edge projections folded onto the target rows (fused.py) -- but organised around 128-edge TILES instead of 16-edge
// warp chunks, so that every contraction is one UMMA with M = 128:
//
//   tile      = up to 128 CSR-consecutive angles = up to TC_R whole target rows, or one 128-angle piece of a long row
//   MMA1      H1acc[128 x 256]   = A[128 x 16] . W1ext^T                      (angle features -> first encoder layer)
//   cvt       H1 = relu(H1acc) -> bf16 -> shared memory (K-major, 128-byte swizzle)       TMEM -> registers -> smem
//   MMA2      S[128 x 16]        = [H1 | K] [128 x 512] . [QT_r,t ; Qbd_r,t]^T          (logits of all (row, head) columns)
//   softmax   each thread owns one angle: picks the 4 columns of its own target row, segmented max / sum over the tile
//   MMA3      D[512 x 16]       += [H1 | V]^T [512 x 128] . P[128 x 16]                  (abar_i,t and agg_i,t, transposed)
//   epilogue  thread = channel: D * 1/z -> abar (bf16), aggv (fp32); (m, z, S) statistics
//
// The SAME shared-memory image of H1 (and the gathered V rows) serves MMA2 as a K-major A operand and MMA3 as an
// MN-major (transposed) A operand: 8 rows x 128 bytes swizzle atoms are identical in both canonical layouts.
// K / V rows are gathered with 16-byte cp.async (one 512-byte row per warp instruction) straight into the swizzled layout;
// the gathers of tile i+1 are issued as soon as the MMA that reads the buffer for tile i has completed.
// Rows longer than 128 angles are accumulated in TMEM across their pieces with the usual online-softmax rescale.
// No atomics, fixed reduction orders: deterministic.  One CTA (128 threads) per SM, static row-range partition.
#include <math.h>

#include "tc.cuh"

namespace alignn {

constexpr int TC_E = 128;     // angles per tile = UMMA M
constexpr int TC_R = 4;       // target rows per tile
constexpr int TC_N = 16;      // logit columns: TC_R rows x 4 heads
constexpr int TC_THREADS = 128;

// shared-memory map (bytes); the swizzled operands are 1024-byte aligned
constexpr uint32_t TS_H1 = 0;            // [4 blocks of 64 ch][128 rows][128 B]  bf16, SW128
constexpr uint32_t TS_K = 65536;         // same shape, gathered K rows
constexpr uint32_t TS_V = 131072;        // same shape, gathered V rows
constexpr uint32_t TS_B2 = 196608;       // [8 blocks of 64 k][16 rows][128 B]: QT (blocks 0..3) | block-diagonal q (4..7)
constexpr uint32_t TS_W1 = 212992;       // [256 rows][2 chunks] W1ext, no swizzle (8-row groups of 256 B)
constexpr uint32_t TS_A = 221184;        // [128 rows][2 chunks] packed angle rows, no swizzle
constexpr uint32_t TS_P = 225280;        // [128 angles][16 cols] bf16, MN-major, no swizzle
constexpr uint32_t TS_MISC = 229376;     // barriers, TMEM pointer, per-warp softmax partials, 1/z
constexpr uint32_t TS_TOTAL = TS_MISC + 2048;

// TMEM columns
constexpr uint32_t TM_H1 = 0;            // 256 columns
constexpr uint32_t TM_S = 256;           // 4 partial blocks x 16 (one per issuing warp, summed on read)
constexpr uint32_t TM_D = 320;           // 4 blocks x 16: (H1 ch 0..127, H1 ch 128..255, V ch 0..127, V ch 128..255)

struct Misc {
    uint64_t bar[3];                     // completion of MMA1 / MMA2 / MMA3
    uint32_t tmem_base;
    uint32_t pad;
    float pmax[4][TC_N];                 // per warp, per (row, head) column
    float psum[4][TC_N];
    float pzd[4][TC_N];
    float inv[TC_N];                     // 1 / (z + 1e-16) per column, for the epilogue
};

__device__ __forceinline__ uint32_t pack_relu_bf16_tc(uint32_t lo_bits, uint32_t hi_bits) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
    return d;
}
// order-preserving float <-> uint (for redux max)
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// ---- tiles -----------------------------------------------------------------------------------------------------------
struct Tile {
    int row0, R;            // first target row, rows in the tile (1..TC_R)
    int e0, ne;             // first CSR position, angles in the tile (1..128)
    int rs[TC_R + 1];       // row starts relative to e0 (rs[R] = ne; rs[k > R] = 1 << 20)
    bool first, last;       // single-row tiles: first / last piece of the row (multi-row tiles: both true)
    bool valid;
};

struct TileCursor {
    int row, row_hi, pos;   // next row, end of this CTA's row range, next CSR position inside `row` (piece cursor)
    __device__ __forceinline__ void init(const int32_t *__restrict__ rowptr, int r0, int r1) {
        row = r0; row_hi = r1; pos = r0 < r1 ? __ldg(rowptr + r0) : 0;
    }
    // rows without in-edges in front of the cursor are skipped here; the caller zero-fills [*skip_lo, *skip_hi)
    __device__ __forceinline__ Tile take(const int32_t *__restrict__ rowptr, int *skip_lo, int *skip_hi) {
        Tile t;
        *skip_lo = row;
        while (row < row_hi && __ldg(rowptr + row + 1) == pos && __ldg(rowptr + row) == pos) ++row;   // empty rows
        *skip_hi = row;
        t.valid = row < row_hi;
        t.row0 = row; t.R = 1; t.e0 = pos; t.ne = 0; t.first = t.last = true;
#pragma unroll
        for (int k = 0; k <= TC_R; ++k) t.rs[k] = 1 << 20;
        if (!t.valid) return t;
        const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        t.rs[0] = 0;
        if (end - pos > TC_E || pos > start) {          // a piece of a long row
            t.first = pos == start;
            t.ne = min(TC_E, end - pos);
            t.last = pos + t.ne == end;
            t.rs[1] = t.ne;
            pos += t.ne;
            if (t.last) ++row;
            return t;
        }
        int total = end - pos, r = 1;
        t.rs[1] = total;
        while (r < TC_R && row + r < row_hi) {
            const int e = __ldg(rowptr + row + r + 1);
            if (e - pos > TC_E) break;
            total = e - pos;
            ++r;
            t.rs[r] = total;
        }
        t.R = r; t.ne = total;
        row += r; pos += total;
        return t;
    }
};

struct TcFwdParams {
    const __nv_bfloat16 *q, *k, *v, *qt, *a_csr;
    const float *w1, *b1;
    const int32_t *rowptr, *col;
    float *aggv;
    __nv_bfloat16 *abar;
    float *stat_m, *stat_z, *stat_s;
    const uint64_t *rng_step;
    int n_nodes, n_edges, in_dim;
    int ldq, ldk, ldv;
    int64_t ldqt, hsqt, ldab, hsab;
    float scale_log2, p_drop, inv_keep;
    uint64_t seed, offset;
    long long *dbg;          // optional [16] cycle counters of CTA 0 (profiling builds of the caller; NULL otherwise)
};

// gather 128 rows of 512 B (row u from src + col[e0 + u] * ld) into a [4][128][128 B] SW128 image; rows >= ne keep stale data
__device__ __forceinline__ void tc_gather(uint32_t base, const __nv_bfloat16 *__restrict__ src, int ld,
                                          const int32_t *__restrict__ col, int e0, int ne, int warp, int lane) {
    const int r0 = warp * 32;
    if (r0 >= ne) return;
    const int jmine = __ldg(col + e0 + min(r0 + lane, ne - 1));
    const char *g = reinterpret_cast<const char *>(src) + lane * 16;
    const uint32_t blk = base + (uint32_t)(lane >> 3) * 16384u;
    const uint32_t cc = (uint32_t)(lane & 7);
    const uint64_t stride = (uint64_t)ld * 2u;
#pragma unroll 8
    for (int u = 0; u < 32; ++u) {
        const uint32_t ju = (uint32_t)__shfl_sync(FULL, jmine, u);
        const int row = r0 + u;
        if (row < ne) cp_async16(blk + (uint32_t)row * 128u + ((cc ^ (uint32_t)(row & 7)) << 4), g + (uint64_t)ju * stride);
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
lgattn_fwd_tc_kernel(const TcFwdParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sb = smem_u32(smem);
    Misc *misc = reinterpret_cast<Misc *>(smem + TS_MISC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: zero the operand images, W1ext image, barriers, TMEM ----------------------------------------
    for (uint32_t off = tid * 16; off < TS_MISC; off += TC_THREADS * 16) sts128(sb + off, make_uint4(0, 0, 0, 0));
    __syncthreads();
    for (int i = tid; i < 256 * 2; i += TC_THREADS) {          // W1ext(ch, c): 8 consecutive features of chunk c
        const int ch = i >> 1, c = i & 1;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = 8 * c + j;
            f[j] = r < P.in_dim ? __ldg(P.w1 + ch * P.in_dim + r) : (r == P.in_dim ? __ldg(P.b1 + ch) : 0.f);
        }
        sts128(sb + TS_W1 + (uint32_t)(ch >> 3) * 256u + (uint32_t)c * 128u + (uint32_t)(ch & 7) * 16u,
               make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
    }
    if (tid == 0) {
        mbar_init(&misc->bar[0], 1);
        mbar_init(&misc->bar[1], 4);          // four issuing threads each commit their own MMAs
        mbar_init(&misc->bar[2], 4);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&misc->tmem_base))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_sync();
    const uint32_t tmem = misc->tmem_base;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);     // this warp's 32 TMEM lanes

    // ---- static partition of the target rows by cost = angles + KAPPA * rows -----------------------------------------
    const int64_t total = (int64_t)P.n_edges + (int64_t)ROW_KAPPA * P.n_nodes;
    const int r_lo = (int)row_bound(P.rowptr, P.n_nodes, total * blockIdx.x / gridDim.x);
    const int r_hi = blockIdx.x + 1 == gridDim.x ? P.n_nodes
                                                 : (int)row_bound(P.rowptr, P.n_nodes, total * (blockIdx.x + 1) / gridDim.x);
    const uint64_t rng_off = P.offset + (P.rng_step ? *P.rng_step : 0ull);
    constexpr uint32_t IDESC1 = tc_idesc(128, 256, 0, 0), IDESC2 = tc_idesc(128, TC_N, 0, 0), IDESC3 = tc_idesc(128, TC_N, 1, 1);

    auto zero_rows = [&](int lo, int hi) {       // rows without in-edges: agg = 0, abar = 0, statistics 0 (as lgattn_fwd_kernel)
        for (int r = lo; r < hi; ++r) {
            for (int c = tid; c < 256; c += TC_THREADS) {
                P.aggv[(int64_t)r * 256 + c] = 0.f;
#pragma unroll
                for (int t = 0; t < 4; ++t) P.abar[(int64_t)r * P.ldab + (int64_t)t * P.hsab + c] = __float2bfloat16_rn(0.f);
            }
            if (tid < 4) {
                P.stat_m[(int64_t)r * 4 + tid] = 0.f; P.stat_z[(int64_t)r * 4 + tid] = 0.f; P.stat_s[(int64_t)r * 4 + tid] = 0.f;
            }
        }
    };
    auto load_a = [&](const Tile &t) {           // packed angle rows of the tile (rows beyond the tile: zeros)
        const uint32_t dst = sb + TS_A + (uint32_t)(tid >> 3) * 256u + (uint32_t)(tid & 7) * 16u;
        if (tid < t.ne) {
            const char *src = reinterpret_cast<const char *>(P.a_csr) + (int64_t)(t.e0 + tid) * 32;
            cp_async16(dst, src);
            cp_async16(dst + 128u, src + 16);
        } else {
            sts128(dst, make_uint4(0, 0, 0, 0));
            sts128(dst + 128u, make_uint4(0, 0, 0, 0));
        }
    };
    auto load_b2 = [&](const Tile &t) {          // QT rows (k-blocks 0..3) and block-diagonal q rows (k-blocks 4..7)
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            const int idx = tid + it * TC_THREADS;
            const int r = idx / 160, rem = idx - r * 160;
            if (r >= t.R) break;
            const int64_t row = t.row0 + r;
            if (rem < 128) {
                const int h = rem >> 5, c = rem & 31, n = 4 * r + h;
                cp_async16(sb + TS_B2 + (uint32_t)(c >> 3) * 2048u + (uint32_t)n * 128u + (uint32_t)(((c & 7) ^ (n & 7)) << 4),
                           P.qt + row * P.ldqt + (int64_t)h * P.hsqt + c * 8);
            } else {
                const int h = (rem - 128) >> 3, c = (rem - 128) & 7, n = 4 * r + h;
                cp_async16(sb + TS_B2 + (uint32_t)(4 + h) * 2048u + (uint32_t)n * 128u + (uint32_t)((c ^ (n & 7)) << 4),
                           P.q + row * (int64_t)P.ldq + 64 * h + c * 8);
            }
        }
    };

    TileCursor cur;
    cur.init(P.rowptr, r_lo, r_hi);
    int sk_lo, sk_hi;
    Tile T = cur.take(P.rowptr, &sk_lo, &sk_hi);
    zero_rows(sk_lo, sk_hi);
    // prologue: everything tile 0 needs, as three commit groups (A | K + B2 | V) -- the steady state keeps this order
    if (T.valid) load_a(T);
    cp_async_commit();
    if (T.valid) { tc_gather(sb + TS_K, P.k, P.ldk, P.col, T.e0, T.ne, warp, lane); load_b2(T); }
    cp_async_commit();
    if (T.valid) tc_gather(sb + TS_V, P.v, P.ldv, P.col, T.e0, T.ne, warp, lane);
    cp_async_commit();

    float m_run[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, z_run[4] = {0.f, 0.f, 0.f, 0.f},
          zd_run[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t phase = 0;
    long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define TC_MARK(i) do { if (P.dbg) { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } } while (0)
    while (T.valid) {
        int nsk_lo, nsk_hi;
        const Tile Nx = cur.take(P.rowptr, &nsk_lo, &nsk_hi);        // next tile (uniform over the CTA)

        TC_MARK(0);
        // ---- P1: H1acc = A . W1ext^T ---------------------------------------------------------------------------------
        cp_async_wait<2>();                                          // A of this tile
        tc_sync();
        if (tid == 0) {
            tc_mma(tmem + TM_H1, tc_desc(sb + TS_A, 128, 256, 0), tc_desc(sb + TS_W1, 128, 256, 0), IDESC1, 0);
            tc_commit(&misc->bar[0]);
        }
        // dropout keep-scales of this thread's angle (position-keyed Philox, same masks as the mma.sync kernels)
        float keep[4] = {1.f, 1.f, 1.f, 1.f};
        if (P.p_drop > 0.f) dropout_scale4(P.seed, rng_off, (uint64_t)(T.e0 + tid), P.p_drop, P.inv_keep, keep);
        mbar_wait(&misc->bar[0], phase);
        tc_fence_after();
        TC_MARK(1);

        // ---- P2: relu, bf16, shared memory (thread = angle row) ---------------------------------------------------------
        {
            const uint32_t rowoff = sb + TS_H1 + (uint32_t)tid * 128u;
            const uint32_t sw = (uint32_t)(tid & 7);
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                uint32_t v[4][16];
#pragma unroll
                for (int s = 0; s < 4; ++s) TC_LD16(tlane + TM_H1 + cb * 64 + s * 16, v[s]);
                tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t *x = &v[c >> 1][(c & 1) * 8];
                    sts128(rowoff + (uint32_t)cb * 16384u + ((uint32_t)(c ^ sw) << 4),
                           make_uint4(pack_relu_bf16_tc(x[0], x[1]), pack_relu_bf16_tc(x[2], x[3]),
                                      pack_relu_bf16_tc(x[4], x[5]), pack_relu_bf16_tc(x[6], x[7])));
                }
            }
        }

        TC_MARK(2);
        // ---- P3: S = [H1 | K] . [QT ; Qbd]^T -------------------------------------------------------------------------------
        cp_async_wait<1>();                                          // K rows and B2 of this tile
        tc_sync();
        TC_MARK(3);
        if (lane == 0) {                      // four issuing threads: warp w accumulates k-blocks {w, w + 4} into S block w
            const uint64_t hi = ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int kb = warp + 4 * half;
                const uint32_t abase = (kb < 4 ? sb + TS_H1 + kb * 16384u : sb + TS_K + (kb - 4) * 16384u);
                const uint64_t ad = hi | (uint64_t)((abase & 0x3FFFFu) >> 4), bd = hi | (uint64_t)(((sb + TS_B2 + kb * 2048u) & 0x3FFFFu) >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    tc_mma(tmem + TM_S + warp * TC_N, ad + 2 * ks, bd + 2 * ks, IDESC2, (half | ks) ? 1u : 0u);
            }
            tc_commit(&misc->bar[1]);
        }
        mbar_wait(&misc->bar[1], phase);
        tc_fence_after();
        TC_MARK(4);
        // K, B2 and A are free: prefetch for the next tile (commit group "A", then "K + B2")
        if (Nx.valid) load_a(Nx);
        cp_async_commit();
        if (Nx.valid) { tc_gather(sb + TS_K, P.k, P.ldk, P.col, Nx.e0, Nx.ne, warp, lane); load_b2(Nx); }
        cp_async_commit();

        TC_MARK(10);
        // ---- P4: segmented softmax over the tile (thread = angle) ---------------------------------------------------------
        const bool valid = tid < T.ne;
        const int r_e = (tid >= T.rs[1]) + (tid >= T.rs[2]) + (tid >= T.rs[3]);
        float s[4];
        {
            uint32_t v[4][16];
#pragma unroll
            for (int b = 0; b < 4; ++b) TC_LD16(tlane + TM_S + b * TC_N, v[b]);
            tc_wait_ld();
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float x = 0.f;
#pragma unroll
                for (int b = 0; b < 4; ++b)      // fixed order: k-blocks (0,4) + (1,5) + (2,6) + (3,7)
                    x += __uint_as_float(r_e == 0 ? v[b][t] : (r_e == 1 ? v[b][4 + t] : (r_e == 2 ? v[b][8 + t] : v[b][12 + t])));
                s[t] = valid ? x * P.scale_log2 : -INFINITY;
            }
        }
        TC_MARK(11);
        if (lane < TC_N) { misc->pmax[warp][lane] = -INFINITY; misc->psum[warp][lane] = 0.f; misc->pzd[warp][lane] = 0.f; }
        __syncwarp();
        const unsigned seg = __match_any_sync(FULL, valid ? r_e : TC_R);
        const int seg_lo = __ffs(seg) - 1, seg_hi = 31 - __clz(seg);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float mx = ord2f(__reduce_max_sync(seg, f2ord(s[t])));
            if (valid && lane == seg_lo) misc->pmax[warp][4 * r_e + t] = mx;
        }
        __syncthreads();
        TC_MARK(12);
        float m_new[4], corr[4], p[4], pd[4];
        {
            const int cbase = valid ? 4 * r_e : 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float mt = fmaxf(fmaxf(misc->pmax[0][cbase + t], misc->pmax[1][cbase + t]),
                                       fmaxf(misc->pmax[2][cbase + t], misc->pmax[3][cbase + t]));
                // single-row tiles continue the row's running maximum; multi-row tiles start fresh
                const float mo = T.first ? -INFINITY : m_run[t];
                m_new[t] = fmaxf(mo, mt);
                corr[t] = T.first ? 0.f : fast_exp2(mo - m_new[t]);
                p[t] = valid ? fast_exp2(s[t] - m_new[t]) : 0.f;
                pd[t] = p[t] * keep[t];
            }
        }
        // segment sums (lanes of one row are contiguous): the lowest lane of the segment ends up with the total
        float zs[4], zds[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { zs[t] = p[t]; zds[t] = pd[t]; }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float a = __shfl_down_sync(FULL, zs[t], off), b = __shfl_down_sync(FULL, zds[t], off);
                if (lane + off <= seg_hi) { zs[t] += a; zds[t] += b; }
            }
        }
        if (valid && lane == seg_lo) {
#pragma unroll
            for (int t = 0; t < 4; ++t) { misc->psum[warp][4 * r_e + t] = zs[t]; misc->pzd[warp][4 * r_e + t] = zds[t]; }
        }
        {   // P row of this angle: 16 bf16 = (row, head) columns, zero outside its own row
            const uint32_t lo = pack_bf16(pd[0], pd[1]), hi = pack_bf16(pd[2], pd[3]);
            const uint32_t dst = sb + TS_P + (uint32_t)(tid >> 3) * 256u + (uint32_t)(tid & 7) * 16u;
            uint4 g0 = make_uint4(0, 0, 0, 0), g1 = make_uint4(0, 0, 0, 0);
            if (valid) {
                if (r_e == 0) { g0.x = lo; g0.y = hi; } else if (r_e == 1) { g0.z = lo; g0.w = hi; }
                else if (r_e == 2) { g1.x = lo; g1.y = hi; } else { g1.z = lo; g1.w = hi; }
            }
            sts128(dst, g0);
            sts128(dst + 128u, g1);
        }
        // long rows: rescale the accumulated columns 0..3 of all four D blocks (thread = channel lane)
        if (!T.first && (corr[0] != 1.f || corr[1] != 1.f || corr[2] != 1.f || corr[3] != 1.f)) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint32_t d[4];
                tc_ld4(tlane + TM_D + b * 16, d);
                tc_wait_ld();
#pragma unroll
                for (int t = 0; t < 4; ++t) d[t] = __float_as_uint(__uint_as_float(d[t]) * corr[t]);
                tc_st4(tlane + TM_D + b * 16, d);
            }
            tc_wait_st();
        }

        TC_MARK(5);
        // ---- P5: D += [H1 | V]^T . P -----------------------------------------------------------------------------------------
        cp_async_wait<2>();                                          // V rows of this tile (A and K + B2 of the next may be in flight)
        tc_sync();
        TC_MARK(6);
        if (lane == 0) {                      // four issuing threads: warp b owns D block b
            const int ksteps = (T.ne + 15) >> 4;
            const int b = warp;
            const uint32_t abase = (b < 2 ? sb + TS_H1 : sb + TS_V) + (uint32_t)(b & 1) * 32768u;
            const uint64_t ad = tc_desc(abase, 16384, 1024, 2), bd = tc_desc(sb + TS_P, 256, 128, 0);
            for (int ks = 0; ks < ksteps; ++ks)
                tc_mma(tmem + TM_D + b * 16, ad + (uint64_t)(ks * (2048 >> 4)), bd + (uint64_t)(ks * (512 >> 4)), IDESC3,
                       (ks > 0 || !T.first) ? 1u : 0u);
            tc_commit(&misc->bar[2]);
        }
        // running statistics of the row(s): fixed-order sums over the four warps
        float z_t[4], zd_t[4];
        if (T.R == 1) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float zz = (misc->psum[0][t] + misc->psum[1][t]) + (misc->psum[2][t] + misc->psum[3][t]);
                const float zd = (misc->pzd[0][t] + misc->pzd[1][t]) + (misc->pzd[2][t] + misc->pzd[3][t]);
                z_run[t] = z_run[t] * corr[t] + zz;
                zd_run[t] = zd_run[t] * corr[t] + zd;
                m_run[t] = m_new[t];
                z_t[t] = z_run[t]; zd_t[t] = zd_run[t];
            }
        }
        if (T.last && tid < TC_N && (tid >> 2) < T.R) {               // thread = (row, head) column: final statistics
            const int r = tid >> 2, t = tid & 3;
            float zz, zd, mm;
            if (T.R == 1) {
                zz = z_t[t]; zd = zd_t[t]; mm = m_run[t];
            } else {
                zz = (misc->psum[0][tid] + misc->psum[1][tid]) + (misc->psum[2][tid] + misc->psum[3][tid]);
                zd = (misc->pzd[0][tid] + misc->pzd[1][tid]) + (misc->pzd[2][tid] + misc->pzd[3][tid]);
                mm = fmaxf(fmaxf(misc->pmax[0][tid], misc->pmax[1][tid]), fmaxf(misc->pmax[2][tid], misc->pmax[3][tid]));
            }
            const float inv = 1.0f / (zz + 1e-16f);
            const int64_t o = (int64_t)(T.row0 + r) * 4 + t;
            const bool empty = zz == 0.f;                              // a row without in-edges inside a multi-row tile
            P.stat_m[o] = empty ? 0.f : mm;
            P.stat_z[o] = zz;
            P.stat_s[o] = zd * inv;
            misc->inv[tid] = inv;
        }
        mbar_wait(&misc->bar[2], phase);
        tc_fence_after();
        TC_MARK(7);
        // V is free: gather the next tile's V rows (commit group "V")
        if (Nx.valid) tc_gather(sb + TS_V, P.v, P.ldv, P.col, Nx.e0, Nx.ne, warp, lane);
        cp_async_commit();

        TC_MARK(13);
        // ---- P6: epilogue (thread = channel) ---------------------------------------------------------------------------------
        if (T.last) {
            __syncthreads();                                         // misc->inv
            float inv[TC_N];
#pragma unroll
            for (int c = 0; c < TC_N; ++c) inv[c] = misc->inv[c];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint32_t d[16];
                TC_LD16(tlane + TM_D + b * 16, d);
                tc_wait_ld();
                const int ch = (b & 1) * 128 + tid;
                if (b < 2) {
                    for (int r = 0; r < T.R; ++r)
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            P.abar[(int64_t)(T.row0 + r) * P.ldab + (int64_t)t * P.hsab + ch] =
                                __float2bfloat16_rn(__uint_as_float(d[4 * r + t]) * inv[4 * r + t]);
                } else {
                    const int h = ch >> 6;
                    for (int r = 0; r < T.R; ++r) {
                        const uint32_t x = h == 0 ? d[4 * r] : (h == 1 ? d[4 * r + 1] : (h == 2 ? d[4 * r + 2] : d[4 * r + 3]));
                        const float iv = h == 0 ? inv[4 * r] : (h == 1 ? inv[4 * r + 1] : (h == 2 ? inv[4 * r + 2] : inv[4 * r + 3]));
                        P.aggv[(int64_t)(T.row0 + r) * 256 + ch] = __uint_as_float(x) * iv;
                    }
                }
            }
            if (T.R == 1) {
#pragma unroll
                for (int t = 0; t < 4; ++t) { m_run[t] = -INFINITY; z_run[t] = 0.f; zd_run[t] = 0.f; }
            }
        }
        zero_rows(nsk_lo, nsk_hi);
        T = Nx;
        phase ^= 1u;
        TC_MARK(8);
        tacc[9] += 1;
    }
    if (P.dbg && blockIdx.x == 0 && tid == 0)
        for (int i = 0; i < 16; ++i) P.dbg[i] = tacc[i];
    cp_async_wait<0>();
    tc_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_lgattn_supported(int hidden, int heads, int in_dim, int dtype);

static long long *g_tc_dbg = nullptr;
// profiling hook (not part of the public header): device buffer of 16 int64 receiving CTA 0's per-phase cycle counts
extern "C" void alignn_lgattn_tc_debug(long long *buf) { g_tc_dbg = buf; }

extern "C" int alignn_lgattn_fwd_tc(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                                    const void *qt, int64_t ldqt, int64_t hsqt,
                                    const void *a_csr, const float *w1, const float *b1, int in_dim,
                                    const int32_t *rowptr, const int32_t *col,
                                    float *aggv, void *abar, int64_t ldab, int64_t hsab,
                                    float *stat_m, float *stat_z, float *stat_s,
                                    int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                    float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream) {
    if (!alignn_lgattn_supported(hidden, heads, in_dim, dtype)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes < 0 || n_edges < 0 || !(p_drop >= 0.f && p_drop < 1.f)) return ALIGNN_ERR_BAD_ARG;
    if (n_nodes >= ((int64_t)1 << 31) - 1 || n_edges >= ((int64_t)1 << 31) - 256) return ALIGNN_ERR_BAD_SHAPE;
    if (ldq >= ((int64_t)1 << 30) || ldk >= ((int64_t)1 << 30) || ldv >= ((int64_t)1 << 30)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_nodes == 0) return ALIGNN_OK;
    if (!q || !k || !v || !qt || !w1 || !b1 || !rowptr || !aggv || !abar || !stat_m || !stat_z || !stat_s)
        return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!a_csr || !col)) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(qt) || !aligned16(a_csr) || (ldq % 8) || (ldk % 8) ||
        (ldv % 8) || (ldqt % 8) || (hsqt % 8))
        return ALIGNN_ERR_BAD_ARG;
    TcFwdParams p;
    p.q = (const __nv_bfloat16 *)q; p.k = (const __nv_bfloat16 *)k; p.v = (const __nv_bfloat16 *)v;
    p.qt = (const __nv_bfloat16 *)qt; p.a_csr = (const __nv_bfloat16 *)a_csr; p.w1 = w1; p.b1 = b1;
    p.rowptr = rowptr; p.col = col; p.aggv = aggv; p.abar = (__nv_bfloat16 *)abar;
    p.stat_m = stat_m; p.stat_z = stat_z; p.stat_s = stat_s; p.rng_step = rng_step;
    p.n_nodes = (int)n_nodes; p.n_edges = (int)n_edges; p.in_dim = in_dim;
    p.ldq = (int)ldq; p.ldk = (int)ldk; p.ldv = (int)ldv; p.ldqt = ldqt; p.hsqt = hsqt; p.ldab = ldab; p.hsab = hsab;
    p.scale_log2 = LOG2E / sqrtf((float)(hidden / heads));
    p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    p.seed = seed; p.offset = offset;
    p.dbg = g_tc_dbg;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(lgattn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_TOTAL));
    int dev = 0, sms = 148;
    ALIGNN_CUDA_TRY(cudaGetDevice(&dev));
    ALIGNN_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t tiles = (n_edges + TC_E - 1) / TC_E + 1;
    const int grid = (int)(tiles < sms ? tiles : sms);
    lgattn_fwd_tc_kernel<<<grid, TC_THREADS, TS_TOTAL, st>>>(p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
