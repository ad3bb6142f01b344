// Graph plan: stable target-sort (CSR) and source-sort (CSC) of an edge list, on device.
//
// Replaces the index handling inside PyG's propagate/softmax/scatter for the conv calls at
// reference scripts/train.py:315,334.  Output is bit-identical to torch.sort(index, stable=True):
// a least-significant-digit radix sort, 8 bits per pass, whose scatter step ranks equal digits in
// input order (warp match + ordered per-warp bases), so no atomics decide any output position.
//
// HBM-bound integer work: ~3 passes x (2 reads + 1 write) of (key,value) int32 pairs per sort --
// a few % of one conv layer's traffic, and the plan is shared by all layers and by fwd+bwd.
#include "common.cuh"

namespace alignn {

constexpr int SORT_THREADS = 256;              // == number of 8-bit digits
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_TILE = 4096;                // keys per block
constexpr int SORT_PER_WARP = SORT_TILE / SORT_WARPS;

// keys for both sorts + identity values; edges with any index outside [0, n_nodes) get the sentinel
// key n_nodes in BOTH sorts (they end up after rowptr[n_nodes] and are never visited).
__global__ void plan_prep_kernel(const int64_t *__restrict__ src_row, const int64_t *__restrict__ dst_row,
                                 int64_t n_edges, int64_t n_nodes, int32_t *__restrict__ key_dst,
                                 int32_t *__restrict__ key_src, int32_t *__restrict__ vals,
                                 int32_t *__restrict__ status, int check_src_sorted, int64_t n_valid) {
    // n_nodes here is the KEY BOUND (sentinel); indices in [bound, n_valid) are legal node ids the caller's bound excluded
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int64_t s = src_row[e], d = dst_row[e];
    const bool ok = s >= 0 && s < n_nodes && d >= 0 && d < n_nodes;
    key_dst[e] = ok ? (int32_t)d : (int32_t)n_nodes;
    key_src[e] = ok ? (int32_t)s : (int32_t)n_nodes;
    vals[e] = (int32_t)e;
    if (!ok) atomicOr(status, (s >= 0 && s < n_valid && d >= 0 && d < n_valid) ? 4 : 1);
    if (check_src_sorted && e > 0) {   // the caller's "already source-sorted" hint, verified on the sort KEYS
        const int64_t sp = src_row[e - 1], dp = dst_row[e - 1];
        const bool okp = sp >= 0 && sp < n_nodes && dp >= 0 && dp < n_nodes;
        const int64_t kp = okp ? sp : n_nodes, kc = ok ? s : n_nodes;
        if (kc < kp) atomicOr(status, 2);
    }
}

__global__ void radix_hist_kernel(const int32_t *__restrict__ keys, int64_t n, int shift,
                                  int32_t *__restrict__ hist, int nblocks) {
    __shared__ int32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) {
        const int64_t idx = base + i;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255], 1);  // integer counts: order-independent
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix and
// the block total through `total`.
__device__ __forceinline__ int block_exclusive_scan_256(int x, int *total) {
    __shared__ int warp_tot[SORT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = x;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
        const int t = warp_tot[w];
        if (w < warp) wbase += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return wbase + incl - x;
}

// one block per digit: exclusive scan of that digit's per-tile counts (in place) + digit total
__global__ void radix_scan_tiles_kernel(int32_t *__restrict__ hist, int32_t *__restrict__ totals, int nblocks) {
    int32_t *row = hist + (int64_t)blockIdx.x * nblocks;
    int carry = 0;
    for (int base = 0; base < nblocks; base += SORT_THREADS) {
        const int i = base + threadIdx.x;
        const int x = i < nblocks ? row[i] : 0;
        int tot;
        const int ex = block_exclusive_scan_256(x, &tot);
        if (i < nblocks) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void radix_scan_totals_kernel(int32_t *__restrict__ totals) {
    int tot;
    const int ex = block_exclusive_scan_256(totals[threadIdx.x], &tot);
    totals[threadIdx.x] = ex;
}

__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const int32_t *__restrict__ keys_in, const int32_t *__restrict__ vals_in,
                     int32_t *__restrict__ keys_out, int32_t *__restrict__ vals_out, int64_t n, int shift,
                     const int32_t *__restrict__ hist, const int32_t *__restrict__ totals, int nblocks) {
    __shared__ int32_t cnt[SORT_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * SORT_TILE + (int64_t)warp * SORT_PER_WARP;
    for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    // phase 1: digit counts of each warp's contiguous sub-tile
    for (int it = 0; it < SORT_PER_WARP / 32; ++it) {
        const int64_t idx = w0 + it * 32 + lane;
        if (idx < n) atomicAdd(&cnt[warp][(keys_in[idx] >> shift) & 255], 1);
    }
    __syncthreads();
    // phase 2: turn counts into ordered bases: digit total prefix + earlier tiles + earlier warps
    {
        const int d = threadIdx.x;
        int running = totals[d] + hist[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const int c = cnt[w][d];
            cnt[w][d] = running;
            running += c;
        }
    }
    __syncthreads();
    // phase 3: in-order placement; equal digits inside a 32-key group are ranked by lane id
    for (int it = 0; it < SORT_PER_WARP / 32; ++it) {
        const int64_t idx = w0 + it * 32 + lane;
        const bool valid = idx < n;
        const unsigned act = __ballot_sync(FULL, valid);
        if (valid) {
            const int32_t key = keys_in[idx];
            const int32_t val = vals_in[idx];
            const int d = (key >> shift) & 255;
            const unsigned peers = __match_any_sync(act, d);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int base = cnt[warp][d];
            __syncwarp(act);
            if (rank == 0) cnt[warp][d] = base + __popc(peers);
            __syncwarp(act);
            keys_out[base + rank] = key;
            vals_out[base + rank] = val;
        }
    }
}

// rowptr[i] = first sorted position whose key is >= i (binary search: empty-row runs of any length
// cost the same).  One thread per row 0..n_nodes.
__global__ void plan_rowptr_kernel(const int32_t *__restrict__ keys, int64_t n_edges, int64_t n_nodes,
                                   int32_t *__restrict__ rowptr, int64_t bound) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_nodes) return;
    const int64_t target = i < bound ? i : bound;      // rows at / beyond the key bound are empty: all point at the first sentinel
    int64_t lo = 0, hi = n_edges;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    rowptr[i] = (int32_t)lo;
}

// permutation + gather of the opposite endpoint.  One thread per sorted position.
__global__ void plan_finish_kernel(const int32_t *__restrict__ keys, const int32_t *__restrict__ vals,
                                   const int64_t *__restrict__ other_row, int64_t n_edges, int64_t n_nodes,
                                   int32_t *__restrict__ col, int32_t *__restrict__ eid) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_edges) return;
    const int32_t id = vals[p];
    const bool live = keys[p] < n_nodes;
    eid[p] = id;                                       // dropped edges keep their id: eid stays a permutation of [0, Ne)
    col[p] = live ? (int32_t)other_row[id] : 0;        // ... and get a safe endpoint (never visited: beyond rowptr[Nn])
}

// inv[eid[q]] = q, then pos_t[p] = inv[eid_t[p]]: CSR position of the p-th edge of the CSC order
__global__ void plan_invert_kernel(const int32_t *__restrict__ eid, int64_t n, int32_t *__restrict__ inv) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) inv[eid[q]] = (int32_t)q;
}
__global__ void plan_compose_kernel(const int32_t *__restrict__ inv, const int32_t *__restrict__ eid_t, int64_t n,
                                    int32_t *__restrict__ pos_t) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) pos_t[p] = inv[eid_t[p]];
}

static inline int num_sort_tiles(int64_t n) { return (int)((n + SORT_TILE - 1) / SORT_TILE); }

struct PlanWorkspace {
    int32_t *key_dst, *key_src, *vals0, *keys_tmp, *vals_a, *vals_b, *hist, *totals;
};

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t carve(void *ws, int64_t n_edges, PlanWorkspace *out) {
    const size_t per = align_up((size_t)(n_edges > 0 ? n_edges : 1) * sizeof(int32_t));
    const size_t hist = align_up((size_t)256 * (size_t)(num_sort_tiles(n_edges) > 0 ? num_sort_tiles(n_edges) : 1) * sizeof(int32_t));
    const size_t tot = align_up(256 * sizeof(int32_t));
    char *p = reinterpret_cast<char *>(ws);
    if (out) {
        out->key_dst = (int32_t *)(p);
        out->key_src = (int32_t *)(p + per);
        out->vals0 = (int32_t *)(p + 2 * per);
        out->keys_tmp = (int32_t *)(p + 3 * per);
        out->vals_a = (int32_t *)(p + 4 * per);
        out->vals_b = (int32_t *)(p + 5 * per);
        out->hist = (int32_t *)(p + 6 * per);
        out->totals = (int32_t *)(p + 6 * per + hist);
    }
    return 6 * per + hist + tot;
}

// sorts (keys, vals0) by key; the sorted keys end in `keys` or `keys_tmp`; returns which via pointers
static int radix_sort_pairs(int32_t *keys, const int32_t *vals0, const PlanWorkspace &w, int64_t n, int64_t max_key,
                            cudaStream_t st, int32_t **sorted_keys, int32_t **sorted_vals) {
    int bits = 1;
    while (((int64_t)1 << bits) <= max_key) ++bits;
    const int passes = (bits + 7) / 8;
    const int nb = num_sort_tiles(n);
    int32_t *kin = keys, *kout = w.keys_tmp;
    const int32_t *vin = vals0;
    int32_t *vout = w.vals_a, *vnext = w.vals_b;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        radix_hist_kernel<<<nb, SORT_THREADS, 0, st>>>(kin, n, shift, w.hist, nb);
        radix_scan_tiles_kernel<<<256, SORT_THREADS, 0, st>>>(w.hist, w.totals, nb);
        radix_scan_totals_kernel<<<1, SORT_THREADS, 0, st>>>(w.totals);
        radix_scatter_kernel<<<nb, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, w.hist, w.totals, nb);
        ALIGNN_LAUNCH_CHECK();
        int32_t *t = kin; kin = kout; kout = t;
        vin = vout;
        int32_t *tv = vout; vout = vnext; vnext = tv;
    }
    *sorted_keys = kin;
    *sorted_vals = const_cast<int32_t *>(vin);
    return ALIGNN_OK;
}

}  // namespace alignn

using namespace alignn;

extern "C" size_t alignn_plan_workspace_bytes(int64_t n_edges, int64_t n_nodes) {
    (void)n_nodes;
    if (n_edges < 0) return 0;
    return carve(nullptr, n_edges, nullptr);
}

extern "C" int alignn_build_plan(const int64_t *edge_index, int64_t n_edges, int64_t n_nodes,
                                 int32_t *rowptr, int32_t *col, int32_t *eid,
                                 int32_t *rowptr_t, int32_t *col_t, int32_t *eid_t,
                                 int32_t *status, void *workspace, size_t workspace_bytes, void *stream) {
    return alignn_build_plan_ex(edge_index, n_edges, n_nodes, rowptr, col, eid, rowptr_t, col_t, eid_t, status, workspace,
                                workspace_bytes, 0, stream);
}

extern "C" int alignn_build_plan_bounded(const int64_t *edge_index, int64_t n_edges, int64_t n_nodes, int64_t key_bound,
                                         int32_t *rowptr, int32_t *col, int32_t *eid,
                                         int32_t *rowptr_t, int32_t *col_t, int32_t *eid_t,
                                         int32_t *status, void *workspace, size_t workspace_bytes, int flags, void *stream);

extern "C" int alignn_build_plan_ex(const int64_t *edge_index, int64_t n_edges, int64_t n_nodes,
                                    int32_t *rowptr, int32_t *col, int32_t *eid,
                                    int32_t *rowptr_t, int32_t *col_t, int32_t *eid_t,
                                    int32_t *status, void *workspace, size_t workspace_bytes, int flags, void *stream) {
    return alignn_build_plan_bounded(edge_index, n_edges, n_nodes, n_nodes, rowptr, col, eid, rowptr_t, col_t, eid_t, status,
                                     workspace, workspace_bytes, flags, stream);
}

extern "C" int alignn_build_plan_bounded(const int64_t *edge_index, int64_t n_edges, int64_t n_nodes, int64_t key_bound,
                                         int32_t *rowptr, int32_t *col, int32_t *eid,
                                         int32_t *rowptr_t, int32_t *col_t, int32_t *eid_t,
                                         int32_t *status, void *workspace, size_t workspace_bytes, int flags, void *stream) {
    const int src_sorted = flags & ALIGNN_PLAN_SOURCE_SORTED;
    const int64_t kb = (key_bound < 0 || key_bound > n_nodes) ? n_nodes : key_bound;
    if (n_edges < 0 || n_nodes < 0 || !rowptr || !rowptr_t || !status) return ALIGNN_ERR_BAD_ARG;
    if (n_edges >= ((int64_t)1 << 31) - SORT_TILE || n_nodes >= ((int64_t)1 << 31) - 1) return ALIGNN_ERR_BAD_SHAPE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ALIGNN_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    if (n_edges == 0) {
        ALIGNN_CUDA_TRY(cudaMemsetAsync(rowptr, 0, (size_t)(n_nodes + 1) * sizeof(int32_t), st));
        ALIGNN_CUDA_TRY(cudaMemsetAsync(rowptr_t, 0, (size_t)(n_nodes + 1) * sizeof(int32_t), st));
        return ALIGNN_OK;
    }
    if (!edge_index || !col || !eid || !col_t || !eid_t || !workspace) return ALIGNN_ERR_BAD_ARG;
    if (workspace_bytes < carve(nullptr, n_edges, nullptr)) return ALIGNN_ERR_WORKSPACE;
    PlanWorkspace w;
    carve(workspace, n_edges, &w);
    const int64_t *src_row = edge_index, *dst_row = edge_index + n_edges;
    const int tb = 256;
    plan_prep_kernel<<<(unsigned)((n_edges + tb - 1) / tb), tb, 0, st>>>(src_row, dst_row, n_edges, kb,
                                                                          w.key_dst, w.key_src, w.vals0, status,
                                                                          src_sorted, n_nodes);
    ALIGNN_LAUNCH_CHECK();
    const unsigned fin_blocks = (unsigned)((n_edges + tb - 1) / tb);
    const unsigned row_blocks = (unsigned)((n_nodes + 1 + tb - 1) / tb);
    int32_t *sk = nullptr, *sv = nullptr;
    int rc = radix_sort_pairs(w.key_dst, w.vals0, w, n_edges, kb, st, &sk, &sv);
    if (rc != ALIGNN_OK) return rc;
    plan_rowptr_kernel<<<row_blocks, tb, 0, st>>>(sk, n_edges, n_nodes, rowptr, kb);
    plan_finish_kernel<<<fin_blocks, tb, 0, st>>>(sk, sv, src_row, n_edges, kb, col, eid);
    ALIGNN_LAUNCH_CHECK();
    if (src_sorted) {   // the stable source sort of a source-sorted list is the identity (verified in plan_prep_kernel)
        sk = w.key_src;
        sv = w.vals0;
    } else {
        rc = radix_sort_pairs(w.key_src, w.vals0, w, n_edges, kb, st, &sk, &sv);
        if (rc != ALIGNN_OK) return rc;
    }
    plan_rowptr_kernel<<<row_blocks, tb, 0, st>>>(sk, n_edges, n_nodes, rowptr_t, kb);
    plan_finish_kernel<<<fin_blocks, tb, 0, st>>>(sk, sv, dst_row, n_edges, kb, col_t, eid_t);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int alignn_plan_csc_positions(const int32_t *eid, const int32_t *eid_t, int64_t n_edges, int32_t *scratch,
                                         int32_t *pos_t, void *stream) {
    if (n_edges < 0) return ALIGNN_ERR_BAD_ARG;
    if (n_edges == 0) return ALIGNN_OK;
    if (!eid || !eid_t || !scratch || !pos_t) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int tb = 256;
    const unsigned blocks = (unsigned)((n_edges + tb - 1) / tb);
    plan_invert_kernel<<<blocks, tb, 0, st>>>(eid, n_edges, scratch);
    plan_compose_kernel<<<blocks, tb, 0, st>>>(scratch, eid_t, n_edges, pos_t);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
