// Per-node projection GEMM on tcgen05 / TMEM with TMA-fed operands (sm_100a):
//
//     C[M, N] = A[M, 256] . W[N, 256]^T + bias[N]        A, W, bias, C bf16; fp32 accumulation in TMEM
//
// This is `nn.Linear` on the bond / atom state: the skip projection `x_r = lin_skip(x)` and the stacked
// `q | k | v | qt_0..3 = x [Wq; Wk; Wv; WQT]^T` of PyG TransformerConv (reference scripts/train.py:308, 326 through
// `lin_query / lin_key / lin_value / lin_skip`, and the folded `lin_edge` products of fused.py).  K = hidden = 256 is tiny, so
// the op is HBM-bound (A read once, C written once); what the kernel has to do is keep the weight block resident and stream.
//
// Mapping: persistent CTAs over tiles of 128 rows x 128 columns, enumerated column-block-major, a contiguous range per
// CTA -- so the 64 KB weight block [128, 256] stays in shared memory across a CTA's row tiles and is re-fetched only when
// the column block changes.  Row tiles of A ([128, 256] = four 128-byte-swizzled [128 x 64] boxes, 64 KB) go through a
// two-stage ring filled by TMA (cp.async.bulk.tensor.2d; rows past M arrive as zeros).  Warp roles: warp 0 = TMA producer,
// warp 1 = MMA issuer (16 tcgen05.mma 128 x 128 x 16 per tile, both operands K-major), warps 2..5 = epilogue (TMEM ->
// registers -> + bias -> bf16 -> 16-byte stores, thread = row).  Two accumulators (2 x 128 TMEM columns) alternate, so the
// MMAs of tile i + 1 run under the epilogue of tile i; every hand-over is an mbarrier (no CTA-wide barrier in the loop).
#include <cuda.h>
#include <stdlib.h>

#include "tc.cuh"

namespace alignn {

constexpr int PJ_THREADS = 192;
constexpr int PJ_K = 256;
constexpr int PJ_BM = 128, PJ_BN = 128;
constexpr uint32_t PJ_KBLK = 16384;                   // one [128 rows x 64 bf16] box, 128-byte swizzle
constexpr uint32_t PJ_TILE = 4 * PJ_KBLK;             // [128 x 256] operand tile
constexpr int PJ_STAGES = 2;
constexpr uint32_t PJ_OFF_A = PJ_TILE;                // smem: W block | A stage 0 | A stage 1 | misc
constexpr uint32_t PJ_MISC = PJ_TILE * (1 + PJ_STAGES);
constexpr uint32_t PJ_SMEM = PJ_MISC + 1024;

struct PjMisc {
    uint64_t a_full[PJ_STAGES], a_empty[PJ_STAGES];   // TMA landed / MMAs done with the stage
    uint64_t w_full, w_empty;                          // weight block landed / all MMAs that read it have completed
    uint64_t acc_full[2], acc_empty[2];                // accumulator ready for the epilogue / drained
    uint32_t tmem_base;
};

struct PjParams {
    const __nv_bfloat16 *bias;                         // [N] or null
    __nv_bfloat16 *c;
    int64_t ldc, M;
    int N;
};

__device__ __forceinline__ void pj_tma_2d(uint32_t dst, const CUtensorMap *map, int c_inner, int c_outer, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void pj_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(PJ_THREADS, 1)
proj_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const PjParams P) {
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
    }
    if (warp == 0) tc_alloc_512(&misc->tmem_base);
    tc_sync();
    const uint32_t tmem = misc->tmem_base;

    const int64_t m_tiles = (P.M + PJ_BM - 1) / PJ_BM;
    const int64_t n_tiles = (int64_t)(P.N / PJ_BN) * m_tiles;              // tile id = column block * m_tiles + row tile
    const int64_t t_lo = n_tiles * blockIdx.x / gridDim.x, t_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) {
        // ---- TMA producer -------------------------------------------------------------------------------------------------
        if (lane == 0) {
            int64_t cur_nb = -1;
            uint32_t w_loads = 0;
            for (int64_t t = t_lo; t < t_hi; ++t) {
                const int64_t nb = t / m_tiles, mt = t % m_tiles;
                const int64_t i = t - t_lo;
                if (nb != cur_nb) {
                    // the MMAs of the previous column block signal w_empty once they have all completed
                    if (w_loads > 0) mbar_wait(&misc->w_empty, (w_loads - 1) & 1u);
                    mbar_expect_tx(&misc->w_full, PJ_TILE);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) pj_tma_2d(sb + (uint32_t)kb * PJ_KBLK, &map_w, 64 * kb, (int)(nb * PJ_BN), &misc->w_full);
                    cur_nb = nb;
                    ++w_loads;
                }
                const int stage = (int)(i % PJ_STAGES);
                if (i >= PJ_STAGES) mbar_wait(&misc->a_empty[stage], (uint32_t)((i / PJ_STAGES - 1) & 1));
                mbar_expect_tx(&misc->a_full[stage], PJ_TILE);
                const uint32_t dst = sb + PJ_OFF_A + (uint32_t)stage * PJ_TILE;
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) pj_tma_2d(dst + (uint32_t)kb * PJ_KBLK, &map_a, 64 * kb, (int)(mt * PJ_BM), &misc->a_full[stage]);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ---------------------------------------------------------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t IDESC = tc_idesc(PJ_BM, PJ_BN, 0, 0);
            // K-major, 128-byte swizzle: 8-row groups 1024 B apart; a k-step of 16 bf16 advances the start address by 32 B
            const uint64_t hi = ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
            int64_t cur_nb = -1;
            uint32_t w_loads = 0;
            for (int64_t t = t_lo; t < t_hi; ++t) {
                const int64_t nb = t / m_tiles;
                const int64_t i = t - t_lo;
                const int stage = (int)(i % PJ_STAGES), buf = (int)(i & 1);
                if (nb != cur_nb) {
                    mbar_wait(&misc->w_full, w_loads & 1u);
                    cur_nb = nb;
                    ++w_loads;
                }
                if (i >= 2) mbar_wait(&misc->acc_empty[buf], (uint32_t)((i / 2 - 1) & 1));
                mbar_wait(&misc->a_full[stage], (uint32_t)((i / PJ_STAGES) & 1));
                tc_fence_after();
                const uint32_t abase = sb + PJ_OFF_A + (uint32_t)stage * PJ_TILE;
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    const uint64_t ad = hi | (uint64_t)(((abase + (uint32_t)kb * PJ_KBLK) & 0x3FFFFu) >> 4);
                    const uint64_t bd = hi | (uint64_t)(((sb + (uint32_t)kb * PJ_KBLK) & 0x3FFFFu) >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc_mma(tmem + (uint32_t)buf * PJ_BN, ad + 2 * ks, bd + 2 * ks, IDESC, (kb | ks) ? 1u : 0u);
                }
                tc_commit(&misc->a_empty[stage]);          // the stage may be refilled once these MMAs have read it
                tc_commit(&misc->acc_full[buf]);           // ... and the accumulator is complete
                const bool last_of_block = (t + 1 == t_hi) || ((t + 1) / m_tiles != nb);
                if (last_of_block) tc_commit(&misc->w_empty);
            }
        }
    } else {
        // ---- epilogue: thread = accumulator lane = output row ------------------------------------------------------------------
        const int qtr = warp & 3;                                     // TMEM lane quarter this warp may read
        const uint32_t tlane = tmem + ((uint32_t)(qtr * 32) << 16);
        for (int64_t t = t_lo; t < t_hi; ++t) {
            const int64_t nb = t / m_tiles, mt = t % m_tiles;
            const int64_t i = t - t_lo;
            const int buf = (int)(i & 1);
            mbar_wait(&misc->acc_full[buf], (uint32_t)((i / 2) & 1));
            tc_fence_after();
            const int64_t row = mt * PJ_BM + qtr * 32 + lane;
            const int n0 = (int)(nb * PJ_BN);
            __nv_bfloat16 *crow = P.c + row * P.ldc + n0;
#pragma unroll 2
            for (int cb = 0; cb < PJ_BN / 16; ++cb) {
                uint32_t v[16];
                TC_LD16(tlane + (uint32_t)buf * PJ_BN + cb * 16, v);
                tc_wait_ld();
                float b[16];
                if (P.bias) {
                    const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(P.bias + n0 + cb * 16));
                    const uint4 b1 = __ldg(reinterpret_cast<const uint4 *>(P.bias + n0 + cb * 16) + 1);
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
                if (row < P.M) {
                    *reinterpret_cast<uint4 *>(crow + cb * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4 *>(crow + cb * 16 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) pj_arrive(&misc->acc_empty[buf]);
        }
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

// [rows, 256] bf16 row-major (row stride ld elements) as a 2-D tensor map with [64 columns x 128 rows] boxes, 128-byte swizzle
static int pj_make_map(CUtensorMap *map, const void *base, int64_t rows, int64_t ld) {
    PjEncodeFn enc = pj_encode_fn();
    if (!enc) return ALIGNN_ERR_CUDA_BASE + (int)cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)PJ_K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2u};
    const cuuint32_t box[2] = {64u, (cuuint32_t)PJ_BM};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ALIGNN_OK : ALIGNN_ERR_BAD_ARG;
}

/* C[M, N] (bf16, row stride ldc) = A[M, 256] (row stride lda) . W[N, 256]^T (row stride ldw) + bias[N] (bf16, may be null).
   The `nn.Linear`s of TransformerConv on the node state (reference scripts/train.py:308, 326: lin_query / lin_key /
   lin_value / lin_skip; stacked and folded as in gnn_elasticity_predictor_b200/fused.py). */
extern "C" int alignn_proj_tc(const void *a, int64_t lda, const void *w, int64_t ldw, const void *bias, void *c, int64_t ldc,
                              int64_t M, int N, int K, int dtype, void *stream) {
    if (!alignn_proj_tc_supported(K, N, dtype)) return ALIGNN_ERR_BAD_SHAPE;
    if (M < 0 || lda < K || ldw < K || ldc < N || (lda % 8) || (ldw % 8) || (ldc % 8)) return ALIGNN_ERR_BAD_ARG;
    if (M == 0) return ALIGNN_OK;
    if (M >= ((int64_t)1 << 31) - PJ_BM) return ALIGNN_ERR_BAD_SHAPE;
    if (!a || !w || !c || !aligned16(a) || !aligned16(w) || !aligned16(c) || (bias && !aligned16(bias))) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap map_a, map_w;
    int rc = pj_make_map(&map_a, a, M, lda);
    if (rc != ALIGNN_OK) return rc;
    rc = pj_make_map(&map_w, w, N, ldw);
    if (rc != ALIGNN_OK) return rc;
    PjParams p;
    p.bias = (const __nv_bfloat16 *)bias; p.c = (__nv_bfloat16 *)c; p.ldc = ldc; p.M = M; p.N = N;
    const int64_t tiles = ((M + PJ_BM - 1) / PJ_BM) * (N / PJ_BN);
    const int grid = (int)(tiles < 148 ? tiles : 148);
    ALIGNN_CUDA_TRY(cudaFuncSetAttribute(proj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PJ_SMEM));
    proj_tc_kernel<<<grid, PJ_THREADS, PJ_SMEM, st>>>(map_a, map_w, p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
