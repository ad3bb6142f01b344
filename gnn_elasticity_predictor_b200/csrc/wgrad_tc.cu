// Weight + bias gradient of a node projection in ONE pass over the activations, on tcgen05 / TMEM (sm_100a):
//
//     C[M, 256] = A^T B     (fp32)            A : [K, M] bf16  gradient of the projection output  (dq|dk|dv|bbar.. or dx_r)
//     s[M]      = sum_k A[k, :]  (fp32)        B : [K, 256] bf16 block input x
//
// i.e. `dW = dproj^T x` and `db = dproj.sum(0)` of autograd's AddmmBackward for lin_query / lin_key / lin_value / lin_skip
// (reference scripts/train.py:691 through PyG TransformerConv's Linears) -- a reduction over K = number of bonds / atoms
// (98 304 at BASELINE config 2) with a tiny output: HBM-bound, each operand must be read exactly once.  cuBLAS runs it as a
// split-K GEMM at ~35 % of the copy roofline and the bias gradient needs a second pass over A (colsum_kernel).
//
// Mapping: grid = (k_splits, ceil(M / 256)).  A CTA owns up to 256 gradient channels (two M = 128 accumulators = all 512
// TMEM columns) and a K range; it streams 64-row chunks of A and B through a 3-stage ring of 128-byte-swizzled shared
// memory tiles filled by TMA (cp.async.bulk.tensor.2d, 64 x 64 boxes, completion on mbarriers; rows / channels out of
// range arrive as zeros).  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = column sums + epilogue;
// stages are handed over through full / empty mbarriers only (no CTA-wide barrier in the loop, MMAs issue back to back).
// Both operands are "MN-major" for the UMMA (the contraction index k is the slow index of both row-major tensors), so no
// transposition happens anywhere.  One thread
// issues 8 tcgen05.mma (128 x 256 x 16) per chunk; the column sums are formed from the shared-memory tile by the other
// threads while the MMAs run.  Per-CTA partials go to a [k_splits][256][M] (+ [M]) buffer (column-major per split: a warp
// stores 128 contiguous bytes straight from its TMEM lanes), reduced in a fixed order by wgrad_reduce_kernel: deterministic.
#include <cuda.h>
#include <stdlib.h>

#include "tc.cuh"

namespace alignn {

constexpr int WG_THREADS = 192;                  // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: column sums + epilogue
constexpr int WG_ROWS = 64;                      // K rows per chunk
constexpr int WG_N = 256;                        // width of B (hidden)
constexpr int WG_STAGES = 3;
constexpr uint32_t WG_TILE = 32768;              // one operand, one stage: [4 blocks of 64 ch][64 rows][128 B], 128-byte swizzle
constexpr uint32_t WG_STAGE = 2 * WG_TILE;       // A then B
constexpr uint32_t WG_MISC = WG_STAGES * WG_STAGE;
constexpr uint32_t WG_SMEM = WG_MISC + 1024;

struct WgMisc {
    uint64_t full[WG_STAGES];                    // TMA bytes of the stage have landed            (1 arrival + tx bytes)
    uint64_t empty[WG_STAGES];                   // MMAs done with the stage + 4 reader warps done (5 arrivals)
    uint64_t accum;                              // all MMAs of the CTA have completed
    uint32_t tmem_base;
};

struct WgParams {
    float *partials;                             // [k_splits][256 columns][M] (+ [M] column sums): column-major per split
    int64_t K;
    int M;
    int ch_per_cta;                              // 256 (two accumulators per CTA) or 128 (one; twice the CTAs stream)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c_inner, int c_outer, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const WgParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sb = smem_u32(smem);
    WgMisc *misc = reinterpret_cast<WgMisc *>(smem + WG_MISC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(&misc->full[s], 1);
            mbar_init(&misc->empty[s], 5);
        }
        mbar_init(&misc->accum, 1);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    }
    if (warp == 0) tc_alloc_512(&misc->tmem_base);
    tc_sync();
    const uint32_t tmem = misc->tmem_base;

    const int splits = gridDim.x;
    const int64_t chunks_total = (P.K + WG_ROWS - 1) / WG_ROWS;
    const int64_t c_lo = chunks_total * blockIdx.x / splits, c_hi = chunks_total * (blockIdx.x + 1) / splits;
    const int64_t n_chunks = c_hi - c_lo;
    const int m0 = blockIdx.y * P.ch_per_cta;                     // first gradient channel of this CTA
    const int m_end = min(P.M, m0 + P.ch_per_cta);
    const int n_tiles = (m_end - m0 + 127) / 128;                 // 1 or 2 accumulators

    if (warp == 0) {
        // ---- TMA producer: one stage = 4 + 4 boxes of 64 channels x 64 rows (out-of-range channels / rows arrive as zeros)
        if (lane == 0)
            for (int64_t c = 0; c < n_chunks; ++c) {
                const int stage = (int)(c % WG_STAGES);
                if (c >= WG_STAGES) mbar_wait(&misc->empty[stage], (uint32_t)((c / WG_STAGES - 1) & 1));
                uint64_t *bar = &misc->full[stage];
                mbar_expect_tx(bar, WG_TILE + (uint32_t)n_tiles * 16384u);
                const int row = (int)((c_lo + c) * WG_ROWS);
                const uint32_t base = sb + (uint32_t)stage * WG_STAGE;
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                    if (blk < 2 * n_tiles) tma_load_2d(base + (uint32_t)blk * 8192u, &map_a, m0 + 64 * blk, row, bar);
                    tma_load_2d(base + WG_TILE + (uint32_t)blk * 8192u, &map_b, 64 * blk, row, bar);
                }
            }
    } else if (warp == 1) {
        // ---- MMA issuer: 8 x (128 x 256 x 16) per chunk, back to back; the commit releases the stage
        if (lane == 0) {
            constexpr uint32_t IDESC = tc_idesc(128, WG_N, 1, 1);
            for (int64_t c = 0; c < n_chunks; ++c) {
                const int stage = (int)(c % WG_STAGES);
                mbar_wait(&misc->full[stage], (uint32_t)((c / WG_STAGES) & 1));
                tc_fence_after();
                const uint32_t abase = sb + (uint32_t)stage * WG_STAGE, bbase = abase + WG_TILE;
                for (int t = 0; t < n_tiles; ++t) {
                    const uint64_t ad = tc_desc(abase + (uint32_t)t * 16384u, 8192, 1024, 2);
                    const uint64_t bd = tc_desc(bbase, 8192, 1024, 2);
#pragma unroll
                    for (int ks = 0; ks < WG_ROWS / 16; ++ks)
                        tc_mma(tmem + (uint32_t)t * WG_N, ad + (uint64_t)(ks * (2048 >> 4)), bd + (uint64_t)(ks * (2048 >> 4)),
                               IDESC, (c > 0 || ks > 0) ? 1u : 0u);
                }
                tc_commit(&misc->empty[stage]);
            }
            if (n_chunks > 0) tc_commit(&misc->accum); else mbar_arrive(&misc->accum);
        }
    } else {
        // ---- readers: column sums from the landed tiles, then the epilogue
        const int rt = tid - 64;                                  // 0..127
        float cs0 = 0.f, cs1 = 0.f;                               // channels m0 + 2 rt, m0 + 2 rt + 1
        const int c2 = 2 * rt;
        const uint32_t chunk16 = (uint32_t)((c2 & 63) >> 3), within = (uint32_t)((c2 & 7) * 2);
        for (int64_t c = 0; c < n_chunks; ++c) {
            const int stage = (int)(c % WG_STAGES);
            mbar_wait(&misc->full[stage], (uint32_t)((c / WG_STAGES) & 1));
            const uint32_t blk = sb + (uint32_t)stage * WG_STAGE + (uint32_t)(c2 >> 6) * 8192u;
            float a0 = 0.f, a1 = 0.f;
            if (c2 < 128 * n_tiles)
#pragma unroll 8
            for (int r = 0; r < WG_ROWS; ++r) {
                uint32_t w;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(blk + (uint32_t)r * 128u + ((chunk16 ^ (uint32_t)(r & 7)) << 4) + within));
                a0 += __uint_as_float(w << 16);
                a1 += __uint_as_float(w & 0xffff0000u);
            }
            cs0 += a0;
            cs1 += a1;
            __syncwarp();
            if (lane == 0) mbar_arrive(&misc->empty[stage]);
        }
        mbar_wait(&misc->accum, 0u);
        tc_fence_after();
        float *out = P.partials + (int64_t)blockIdx.x * ((int64_t)P.M * WG_N + P.M);
        const int q = warp & 3;                                   // this warp's TMEM lane quarter
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        for (int t = 0; t < n_tiles; ++t) {
            const int m = m0 + t * 128 + q * 32 + lane;           // thread = accumulator lane = gradient channel
#pragma unroll 1
            for (int cb = 0; cb < WG_N / 16; ++cb) {
                uint32_t v[16];
                if (n_chunks > 0) {
                    TC_LD16(tlane + (uint32_t)t * WG_N + cb * 16, v);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0u;
                }
                if (m < m_end) {                                  // partial layout [col][m]: a warp stores 128 contiguous bytes
#pragma unroll
                    for (int j = 0; j < 16; ++j) out[(int64_t)(cb * 16 + j) * P.M + m] = __uint_as_float(v[j]);
                }
            }
        }
        const int m = m0 + c2;
        float *so = out + (int64_t)P.M * WG_N;
        if (m < m_end) so[m] = cs0;
        if (m + 1 < m_end) so[m + 1] = cs1;
    }
    tc_sync();
    if (warp == 0) tc_dealloc_512(tmem);
}

// out[i] = sum_s partials[s * width + i]: 8 threads per group of 4 columns stride the splits, then fold in a fixed order
constexpr int WR_COLS = 32, WR_ROWS = 8;         // 32 column groups x 8 split lanes per CTA
__global__ void __launch_bounds__(WR_COLS * WR_ROWS)
wgrad_reduce_kernel(const float *__restrict__ partials, int splits, int64_t width, int64_t w_mat, int M,
                    float *__restrict__ c_out, float *__restrict__ s_out) {
    __shared__ float4 red[WR_ROWS][WR_COLS];
    const int cx = threadIdx.x % WR_COLS, ry = threadIdx.x / WR_COLS;
    const int64_t i = ((int64_t)blockIdx.x * WR_COLS + cx) * 4;   // width % 4 == 0
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < width)
#pragma unroll 4
        for (int s = ry; s < splits; s += WR_ROWS) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(partials + (int64_t)s * width + i));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && i < width) {
        float4 t = red[0][cx];
#pragma unroll
        for (int r = 1; r < WR_ROWS; ++r) { t.x += red[r][cx].x; t.y += red[r][cx].y; t.z += red[r][cx].z; t.w += red[r][cx].w; }
        const float r4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t idx = i + j;                        // partial index = col * M + m
            if (idx < w_mat) c_out[(idx % M) * WG_N + idx / M] = r4[j];
            else if (s_out) s_out[idx - w_mat] = r4[j];
        }
    }
}

static int wg_ch_per_cta(int M) { return M <= 256 ? 128 : 256; }

static int wg_splits(int64_t K, int M) {
    const int groups = (M + wg_ch_per_cta(M) - 1) / wg_ch_per_cta(M);
    if (const char *e = getenv("ALIGNN_WG_SPLITS")) {            // tuning knob (profiling only)
        const int v = atoi(e);
        if (v > 0) return v;
    }
    int64_t chunks = (K + WG_ROWS - 1) / WG_ROWS;
    int s = (M <= 256 ? 148 : 74) / groups;      // M <= 256: every SM streams, one 128-channel accumulator each (partials as small
                                                 // as with 74 two-accumulator CTAs; the second reader of a B row hits L2)
    if (s < 1) s = 1;
    if (s > chunks) s = (int)(chunks > 0 ? chunks : 1);
    return s;
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_wgrad_supported(int n, int dtype) { return n == WG_N && dtype == ALIGNN_BF16; }

extern "C" int64_t alignn_wgrad_partial_floats(int64_t K, int M) {
    if (K < 0 || M <= 0) return 0;
    return (int64_t)wg_splits(K, M) * ((int64_t)M * WG_N + M);
}

typedef CUresult (*WgEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static WgEncodeFn wg_encode_fn() {
    static WgEncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<WgEncodeFn>(p);
    }
    return fn;
}

// [rows, cols] bf16 row-major window (row stride ld elements) as a 2-D tensor map with 64 x 64 boxes, 128-byte swizzle
static int wg_make_map(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int64_t ld) {
    WgEncodeFn enc = wg_encode_fn();
    if (!enc) return ALIGNN_ERR_CUDA_BASE + (int)cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2u};
    const cuuint32_t box[2] = {64u, (cuuint32_t)WG_ROWS};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ALIGNN_OK : ALIGNN_ERR_BAD_ARG;
}

extern "C" int alignn_wgrad(const void *a, int64_t lda, const void *b, int64_t ldb, int64_t K, int M, int N, int dtype,
                            float *partials, float *c, float *colsum, void *stream) {
    if (!alignn_wgrad_supported(N, dtype)) return ALIGNN_ERR_BAD_SHAPE;
    if (K <= 0 || M <= 0 || (M % 8) || lda < M || ldb < N || (lda % 8) || (ldb % 8)) return ALIGNN_ERR_BAD_ARG;
    if (!c || !partials || !a || !b || !aligned16(a) || !aligned16(b) || !aligned16(partials) || !aligned16(c))
        return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int splits = wg_splits(K, M);
    CUtensorMap map_a, map_b;
    int rc = wg_make_map(&map_a, a, K, M, lda);
    if (rc != ALIGNN_OK) return rc;
    rc = wg_make_map(&map_b, b, K, N, ldb);
    if (rc != ALIGNN_OK) return rc;
    WgParams p;
    p.partials = partials; p.K = K; p.M = M; p.ch_per_cta = wg_ch_per_cta(M);
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    wgrad_tc_kernel<<<dim3((unsigned)splits, (unsigned)((M + p.ch_per_cta - 1) / p.ch_per_cta)), WG_THREADS, WG_SMEM, st>>>(map_a, map_b, p);
    ALIGNN_LAUNCH_CHECK();
    const int64_t w_mat = (int64_t)M * WG_N, width = w_mat + M;
    wgrad_reduce_kernel<<<(unsigned)((width / 4 + WR_COLS - 1) / WR_COLS), WR_COLS * WR_ROWS, 0, st>>>(partials, splits, width,
                                                                                                       w_mat, M, c, colsum);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
