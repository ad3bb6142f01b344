// Per-node projection GEMM on tcgen05 / TMEM with TMA-fed operands (sm_100a):
//
//     C[M, N] = A[M, 256] . W[N, 256]^T + bias[N]        A, W, bias, C bf16; fp32 accumulation in TMEM
//
// This is `nn.Linear` on the bond / atom state: the skip projection `x_r = lin_skip(x)` and the stacked
// `q | k | v | qt_0..3 = x [Wq; Wk; Wv; WQT]^T` of PyG TransformerConv (reference scripts/train.py:308, 326 through
// `lin_query / lin_key / lin_value / lin_skip`, and the folded `lin_edge` products of fused.py).  K = hidden = 256 is tiny, so
// the op is HBM-bound (A read once, C written once); what the kernel has to do is keep the weight block resident and stream.
//
// Mapping: persistent CTAs over tiles of 128 rows x 256 columns, enumerated column-block-major, a contiguous range per
// CTA -- so the 128 KB weight block [256, 256] stays in shared memory across a CTA's row tiles and is re-fetched only when
// the column block changes (never for N = 256).  A goes through a four-stage ring of [128 rows x 64 k] boxes (16 KB,
// 128-byte swizzle) filled by TMA (cp.async.bulk.tensor.2d; rows past M arrive as zeros): the ring runs across tile
// boundaries, so the loads of the next row tile are in flight while the MMAs of this one issue.  Warp roles: warp 0 = TMA
// producer, warp 1 = MMA issuer (4 tcgen05.mma 128 x 256 x 16 per box, both operands K-major), warps 2..5 = epilogue: TMEM ->
// registers -> + bias -> bf16 -> 128-byte-swizzled staging rows -> TMA store (cp.async.bulk.tensor.2d.global.shared; rows
// past M are clipped by the tensor map).  Two accumulators (2 x 256 TMEM columns) alternate, so the MMAs of tile i + 1 run
// under the epilogue of tile i; every hand-over is an mbarrier (no CTA-wide barrier in the loop).
#include <cuda.h>
#include <stdlib.h>

#include "tc.cuh"

namespace alignn {

constexpr int PJ_THREADS = 192;
constexpr int PJ_K = 256;
constexpr int PJ_BM = 128, PJ_BN = 256;
constexpr uint32_t PJ_ABOX = 16384;                   // A: one [128 rows x 64 bf16] box, 128-byte swizzle
constexpr uint32_t PJ_WBLK = 32768;                   // W: one k-block [256 rows x 64 bf16] = two boxes
constexpr int PJ_STAGES = 4;                          // A boxes in flight
constexpr uint32_t PJ_OFF_A = 4 * PJ_WBLK;            // smem: W block (128 KB) | A ring (64 KB) | staging (32 KB) | misc
constexpr uint32_t PJ_OFF_STG = PJ_OFF_A + PJ_STAGES * PJ_ABOX;
constexpr uint32_t PJ_STG_WARP = 8192;                // per epilogue warp: 2 boxes of [32 rows x 64 bf16]
constexpr uint32_t PJ_MISC = PJ_OFF_STG + 4 * PJ_STG_WARP;
constexpr uint32_t PJ_SMEM = PJ_MISC + 1024;
static_assert(PJ_SMEM <= 227 * 1024, "over the per-CTA shared-memory limit");

struct PjMisc {
    uint64_t a_full[PJ_STAGES], a_empty[PJ_STAGES];   // TMA landed / MMAs done with the box
    uint64_t w_full, w_empty;                          // weight block landed / all MMAs that read it have completed
    uint64_t acc_full[2], acc_empty[2];                // accumulator ready for the epilogue / drained
    uint32_t tmem_base;
};

// Two column groups share one pass over A: group 0 ("full") is computed for all M rows into C0, group 1 ("prefix") for
// the first M_pre rows only into C1 (the trunk's x_r over every bond row + q|k|v|qt over the rows that have line-graph
// neighbours).  Column block b of a group reads W rows [w_row + 256 b, +256) and the bias entries of the same rows.
struct PjParams {
    const __nv_bfloat16 *bias;                         // indexed like the rows of W, or null
    int64_t M, M_pre;
    int nb_full, nb_pre;                               // column blocks (of 256) per group
    int w_row_full, w_row_pre;
};

struct PjTile {
    int wrow, ccol, group;
    int64_t mt;
};
__device__ __forceinline__ PjTile pj_decode(const PjParams &P, int64_t t, int64_t mt_full, int64_t mt_pre) {
    PjTile r;
    const int64_t tiles_full = (int64_t)P.nb_full * mt_full;
    if (t < tiles_full) {
        const int nb = (int)(t / mt_full);
        r.group = 0; r.mt = t % mt_full; r.wrow = P.w_row_full + nb * PJ_BN; r.ccol = nb * PJ_BN;
    } else {
        const int64_t u = t - tiles_full;
        const int nb = (int)(u / mt_pre);
        r.group = 1; r.mt = u % mt_pre; r.wrow = P.w_row_pre + nb * PJ_BN; r.ccol = nb * PJ_BN;
    }
    return r;
}

__device__ __forceinline__ void pj_tma_2d(uint32_t dst, const CUtensorMap *map, int c_inner, int c_outer, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void pj_tma_store_2d(const CUtensorMap *map, int c_inner, int c_outer, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c_inner), "r"(c_outer), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void pj_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(PJ_THREADS, 1)
proj_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c1, const PjParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sb = smem_u32(smem);
    PjMisc *misc = reinterpret_cast<PjMisc *>(smem + PJ_MISC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < PJ_STAGES; ++s) {
            mbar_init(&misc->a_full[s], 1);
            mbar_init(&misc->a_empty[s], 1);
        }
        mbar_init(&misc->w_full, 1);
        mbar_init(&misc->w_empty, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&misc->acc_full[b], 1);
            mbar_init(&misc->acc_empty[b], 4);
        }
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c1)) : "memory");
    }
    if (warp == 0) tc_alloc_512(&misc->tmem_base);
    tc_sync();
    const uint32_t tmem = misc->tmem_base;

    const int64_t mt_full = (P.M + PJ_BM - 1) / PJ_BM, mt_pre = (P.M_pre + PJ_BM - 1) / PJ_BM;
    const int64_t n_tiles = (int64_t)P.nb_full * mt_full + (int64_t)P.nb_pre * mt_pre;   // column-block-major within a group
    const int64_t t_lo = n_tiles * blockIdx.x / gridDim.x, t_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) {
        // ---- TMA producer -------------------------------------------------------------------------------------------------
        if (lane == 0) {
            int cur_w = -1;
            uint32_t w_loads = 0;
            int64_t box = 0;                                              // running A-box counter (ring position)
            for (int64_t t = t_lo; t < t_hi; ++t) {
                const PjTile T = pj_decode(P, t, mt_full, mt_pre);
                const int64_t mt = T.mt;
                if (T.wrow != cur_w) {
                    // the MMAs of the previous column block signal w_empty once they have all completed
                    if (w_loads > 0) mbar_wait(&misc->w_empty, (w_loads - 1) & 1u);
                    mbar_expect_tx(&misc->w_full, 4 * PJ_WBLK);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        pj_tma_2d(sb + (uint32_t)kb * PJ_WBLK, &map_w, 64 * kb, T.wrow, &misc->w_full);
                        pj_tma_2d(sb + (uint32_t)kb * PJ_WBLK + 16384u, &map_w, 64 * kb, T.wrow + 128, &misc->w_full);
                    }
                    cur_w = T.wrow;
                    ++w_loads;
                }
#pragma unroll 1
                for (int kb = 0; kb < 4; ++kb, ++box) {
                    const int stage = (int)(box % PJ_STAGES);
                    if (box >= PJ_STAGES) mbar_wait(&misc->a_empty[stage], (uint32_t)((box / PJ_STAGES - 1) & 1));
                    mbar_expect_tx(&misc->a_full[stage], PJ_ABOX);
                    pj_tma_2d(sb + PJ_OFF_A + (uint32_t)stage * PJ_ABOX, &map_a, 64 * kb, (int)(mt * PJ_BM), &misc->a_full[stage]);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ---------------------------------------------------------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t IDESC = tc_idesc(PJ_BM, PJ_BN, 0, 0);
            // K-major, 128-byte swizzle: 8-row groups 1024 B apart; a k-step of 16 bf16 advances the start address by 32 B
            const uint64_t hi = ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
            int cur_w = -1;
            uint32_t w_loads = 0;
            int64_t box = 0;
            for (int64_t t = t_lo; t < t_hi; ++t) {
                const PjTile T = pj_decode(P, t, mt_full, mt_pre);
                const int64_t i = t - t_lo;
                const int buf = (int)(i & 1);
                if (T.wrow != cur_w) {
                    mbar_wait(&misc->w_full, w_loads & 1u);
                    cur_w = T.wrow;
                    ++w_loads;
                }
                if (i >= 2) mbar_wait(&misc->acc_empty[buf], (uint32_t)((i / 2 - 1) & 1));
#pragma unroll 1
                for (int kb = 0; kb < 4; ++kb, ++box) {
                    const int stage = (int)(box % PJ_STAGES);
                    mbar_wait(&misc->a_full[stage], (uint32_t)((box / PJ_STAGES) & 1));
                    tc_fence_after();
                    const uint64_t ad = hi | (uint64_t)(((sb + PJ_OFF_A + (uint32_t)stage * PJ_ABOX) & 0x3FFFFu) >> 4);
                    const uint64_t bd = hi | (uint64_t)(((sb + (uint32_t)kb * PJ_WBLK) & 0x3FFFFu) >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc_mma(tmem + (uint32_t)buf * PJ_BN, ad + 2 * ks, bd + 2 * ks, IDESC, (kb | ks) ? 1u : 0u);
                    tc_commit(&misc->a_empty[stage]);      // the box may be refilled once these MMAs have read it
                }
                tc_commit(&misc->acc_full[buf]);           // the accumulator is complete
                const bool last_of_block = (t + 1 == t_hi) || pj_decode(P, t + 1, mt_full, mt_pre).wrow != T.wrow;
                if (last_of_block) tc_commit(&misc->w_empty);
            }
        }
    } else {
        // ---- epilogue: thread = accumulator lane = output row ------------------------------------------------------------------
        const int qtr = warp & 3;                                     // TMEM lane quarter this warp may read
        const uint32_t tlane = tmem + ((uint32_t)(qtr * 32) << 16);
        const uint32_t stg = sb + PJ_OFF_STG + (uint32_t)qtr * PJ_STG_WARP;
        const uint32_t srow = stg + (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
        for (int64_t t = t_lo; t < t_hi; ++t) {
            const PjTile T = pj_decode(P, t, mt_full, mt_pre);
            const int64_t i = t - t_lo;
            const int buf = (int)(i & 1);
            mbar_wait(&misc->acc_full[buf], (uint32_t)((i / 2) & 1));
            tc_fence_after();
            const int row0 = (int)(T.mt * PJ_BM + qtr * 32);
            const int n0 = T.ccol;                                    // column in the group's output
            const int b0col = T.wrow;                                 // bias entries follow the rows of W
            const CUtensorMap *cmap = T.group == 0 ? &map_c : &map_c1;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {                    // 128 columns per pass through the staging rows
                // the TMA stores of the previous pass must have READ the staging rows before they are overwritten
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll 2
                for (int cb = 0; cb < 8; ++cb) {
                    const int col = half * 128 + cb * 16;
                    uint32_t v[16];
                    TC_LD16(tlane + (uint32_t)buf * PJ_BN + col, v);
                    tc_wait_ld();
                    float b[16];
                    if (P.bias) {
                        const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(P.bias + b0col + col));
                        const uint4 b1 = __ldg(reinterpret_cast<const uint4 *>(P.bias + b0col + col) + 1);
                        const uint32_t w[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            b[2 * j] = __uint_as_float(w[j] << 16);
                            b[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) b[j] = 0.f;
                    }
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = pack_bf16(__uint_as_float(v[2 * j]) + b[2 * j], __uint_as_float(v[2 * j + 1]) + b[2 * j + 1]);
                    // staging box (cb >> 2) = 64 columns; 16-byte chunk c of row r sits at chunk c ^ (r & 7) (128-byte swizzle)
                    const uint32_t base = srow + (uint32_t)(cb >> 2) * 4096u;
                    const uint32_t c0 = (uint32_t)((cb & 3) * 2);
                    sts128(base + (((c0) ^ sw) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                    sts128(base + (((c0 + 1) ^ sw) << 4), make_uint4(o[4], o[5], o[6], o[7]));
                }
                if (half == 1) {                                       // the accumulator is drained: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) pj_arrive(&misc->acc_empty[buf]);
                }
                fence_proxy_async();                                   // generic-proxy staging writes -> visible to the TMA store
                __syncwarp();
                if (lane == 0) {
                    pj_tma_store_2d(cmap, n0 + half * 128, row0, stg);
                    pj_tma_store_2d(cmap, n0 + half * 128 + 64, row0, stg + 4096u);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // stores complete before the CTA exits
    }
    tc_sync();
    if (warp == 0) tc_dealloc_512(tmem);
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_proj_tc_supported(int K, int N, int dtype) { return K == PJ_K && N > 0 && (N % PJ_BN) == 0 && dtype == ALIGNN_BF16; }

typedef CUresult (*PjEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PjEncodeFn pj_encode_fn() {
    static PjEncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PjEncodeFn>(p);
    }
    return fn;
}

// [rows, cols] bf16 row-major (row stride ld elements) as a 2-D tensor map with [64 columns x box_rows rows] boxes, 128-byte swizzle
static int pj_make_map(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    PjEncodeFn enc = pj_encode_fn();
    if (!enc) return ALIGNN_ERR_CUDA_BASE + (int)cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2u};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ALIGNN_OK : ALIGNN_ERR_BAD_ARG;
}

/* One pass over A[M, 256] (row stride lda) for two column groups of W (rows of W = output columns; row stride ldw):
     C0[M, n_full]         = A          . W[w_row_full : w_row_full + n_full]^T + bias[w_row_full : ...]
     C1[M_pre, n_pre]      = A[:M_pre]  . W[w_row_pre  : w_row_pre  + n_pre ]^T + bias[w_row_pre  : ...]
   n_full, n_pre multiples of 256 (n_pre may be 0: c1 unused), M_pre <= M, bias bf16 indexed like the rows of W (may be null).
   The trunk calls it once per block: x_r = lin_skip(x) for every bond row + q | k | v | qt_0..3 for the rows with
   line-graph neighbours (reference scripts/train.py:308, 326: the Linears of TransformerConv). */
extern "C" int alignn_proj_tc2(const void *a, int64_t lda, const void *w, int64_t ldw, int64_t w_rows, const void *bias,
                               void *c0, int64_t ldc0, int n_full, int w_row_full,
                               void *c1, int64_t ldc1, int n_pre, int w_row_pre,
                               int64_t M, int64_t M_pre, int K, int dtype, void *stream) {
    if (K != PJ_K || dtype != ALIGNN_BF16 || n_full < 0 || n_pre < 0 || (n_full % PJ_BN) || (n_pre % PJ_BN) || n_full + n_pre == 0)
        return ALIGNN_ERR_BAD_SHAPE;
    if (M < 0 || M_pre < 0 || M_pre > M || lda < K || ldw < K || (lda % 8) || (ldw % 8)) return ALIGNN_ERR_BAD_ARG;
    if (w_row_full < 0 || w_row_pre < 0 || w_row_full + n_full > w_rows || w_row_pre + n_pre > w_rows) return ALIGNN_ERR_BAD_ARG;
    if ((n_full && (!c0 || ldc0 < n_full || (ldc0 % 8) || !aligned16(c0))) || (n_pre && (!c1 || ldc1 < n_pre || (ldc1 % 8) || !aligned16(c1))))
        return ALIGNN_ERR_BAD_ARG;
    if (M == 0 || (n_full == 0 && M_pre == 0)) return ALIGNN_OK;
    if (M >= ((int64_t)1 << 31) - PJ_BM) return ALIGNN_ERR_BAD_SHAPE;
    if (!a || !w || !aligned16(a) || !aligned16(w) || (bias && !aligned16(bias))) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap map_a, map_w, map_c0, map_c1;
    int rc = pj_make_map(&map_a, a, M, PJ_K, lda, PJ_BM);
    if (rc != ALIGNN_OK) return rc;
    rc = pj_make_map(&map_w, w, w_rows, PJ_K, ldw, 128);
    if (rc != ALIGNN_OK) return rc;
    const bool has_pre = n_pre > 0 && M_pre > 0;
    rc = n_full ? pj_make_map(&map_c0, c0, M, n_full, ldc0, 32) : pj_make_map(&map_c0, c1, M_pre, n_pre, ldc1, 32);
    if (rc != ALIGNN_OK) return rc;
    rc = has_pre ? pj_make_map(&map_c1, c1, M_pre, n_pre, ldc1, 32) : pj_make_map(&map_c1, n_full ? c0 : c1, n_full ? M : M_pre,
                                                                                  n_full ? n_full : n_pre, n_full ? ldc0 : ldc1, 32);
    if (rc != ALIGNN_OK) return rc;
    PjParams p;
    p.bias = (const __nv_bfloat16 *)bias; p.M = M; p.M_pre = has_pre ? M_pre : 0;
    p.nb_full = n_full / PJ_BN; p.nb_pre = has_pre ? n_pre / PJ_BN : 0;
    p.w_row_full = w_row_full; p.w_row_pre = w_row_pre;
    const int64_t tiles = (int64_t)p.nb_full * ((M + PJ_BM - 1) / PJ_BM) + (int64_t)p.nb_pre * ((p.M_pre + PJ_BM - 1) / PJ_BM);
    if (tiles == 0) return ALIGNN_OK;
    const int grid = (int)(tiles < 148 ? tiles : 148);
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(proj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PJ_SMEM));
    proj_tc_kernel<<<grid, PJ_THREADS, PJ_SMEM, st>>>(map_a, map_w, map_c0, map_c1, p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

/* C[M, N] (bf16, row stride ldc) = A[M, 256] (row stride lda) . W[N, 256]^T (row stride ldw) + bias[N] (bf16, may be null). */
extern "C" int alignn_proj_tc(const void *a, int64_t lda, const void *w, int64_t ldw, const void *bias, void *c, int64_t ldc,
                              int64_t M, int N, int K, int dtype, void *stream) {
    if (!alignn_proj_tc_supported(K, N, dtype)) return ALIGNN_ERR_BAD_SHAPE;
    return alignn_proj_tc2(a, lda, w, ldw, N, bias, c, ldc, N, 0, nullptr, 0, 0, 0, M, 0, K, dtype, stream);
}
