// Data side of the hot path, on the device:
//
//  * collate: the PyG `Batch.from_data_list` rules the reference's DataLoader applies to its `Data` objects
//    (scripts/train.py:2037; scripts/fetch.py:614-651 for the fields) as ONE call over a device-resident store of
//    concatenated graphs -- segmented copies of the feature rows, index rows shifted by the running atom count
//    (`edge_index`; and `lg_edge_index` too under PyG's default `__inc__`) or by the running bond count
//    (`lg_inc = bonds`), the `batch` vector, per-graph rows (`global_x`, `sg_one_hot`, `y`), and -- when the output
//    buffers are larger than the selection -- the shape-bucket padding of batching.py (zero rows, index -1, dummy graph).
//    It replaces `PtGraphDataset.__getitem__` (train.py:130-172: one torch.load per sample per epoch) + the host collate.
//  * bond features and line graph: `_edge_geom` / `_rbf_expand` / the bond loop (fetch.py:250-263, 311-316, 385-396) and
//    the line-graph loop with its angle basis (fetch.py:266-273, 417-447), in float64 like the reference (numpy), results
//    stored as float32 / int64 exactly as `to_pyg_data` stores them (fetch.py:629-633).
//
// All of it is HBM-bound integer / byte / fp64-scalar work: coalesced word copies (16-byte vectors when both sides are
// aligned), one warp per bond, no atomics deciding any position (angle slots come from an exclusive scan of exact counts).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace alignn {

constexpr int DP_THREADS = 256;

// ---- collate ---------------------------------------------------------------------------------------------------------
// seg_ptr[0..2][g] = exclusive prefix of the selected graphs' atom / bond / angle counts, g = 0..n_sel
__global__ void collate_offsets_kernel(const int64_t *__restrict__ sel, int64_t n_sel, int64_t n_graphs,
                                       const int64_t *__restrict__ node_ptr, const int64_t *__restrict__ bond_ptr,
                                       const int64_t *__restrict__ angle_ptr, int64_t *__restrict__ seg_ptr,
                                       int64_t cap_nodes, int64_t cap_bonds, int64_t cap_angles, int64_t max_nodes,
                                       int64_t max_bonds, int64_t max_angles, int32_t *__restrict__ status) {
    __shared__ int64_t warp_tot[3][32];
    __shared__ int64_t carry[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x < 3) carry[threadIdx.x] = 0;
    __syncthreads();
    const int64_t *ptrs[3] = {node_ptr, bond_ptr, angle_ptr};
    const int64_t maxs[3] = {max_nodes, max_bonds, max_angles};
    for (int64_t base = 0; base < n_sel; base += blockDim.x) {
        const int64_t g = base + threadIdx.x;
        int64_t sz[3] = {0, 0, 0};
        if (g < n_sel) {
            const int64_t s = sel[g];
            if (s < 0 || s >= n_graphs) {
                atomicOr(status, 1);
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    sz[k] = ptrs[k][s + 1] - ptrs[k][s];
                    if (sz[k] > maxs[k]) atomicOr(status, 4);
                }
            }
        }
        int64_t incl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int64_t v = sz[k];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t y = __shfl_up_sync(FULL, v, off);
                if (lane >= off) v += y;
            }
            incl[k] = v;
            if (lane == 31) warp_tot[k][warp] = v;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int64_t v = lane < nw ? warp_tot[k][lane] : 0;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int64_t y = __shfl_up_sync(FULL, v, off);
                    if (lane >= off) v += y;
                }
                warp_tot[k][lane] = v;      // inclusive over warps
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int64_t before = carry[k] + (warp > 0 ? warp_tot[k][warp - 1] : 0) + incl[k] - sz[k];
            if (g < n_sel) seg_ptr[k * (n_sel + 1) + g] = before;
        }
        __syncthreads();
        if (threadIdx.x < 3) carry[threadIdx.x] += warp_tot[threadIdx.x][nw - 1];
        __syncthreads();
    }
    if (threadIdx.x < 3) {
        const int64_t tot = carry[threadIdx.x];
        seg_ptr[threadIdx.x * (n_sel + 1) + n_sel] = tot;
        const int64_t cap = threadIdx.x == 0 ? cap_nodes : (threadIdx.x == 1 ? cap_bonds : cap_angles);
        if (tot != cap) atomicOr(status, 2);     // the host-stated total (launch sizing, buffer bound) is wrong
    }
}

// One segment per blockIdx.x (segment n_sel = the padding tail), blockIdx.y strides over the segment's 4-byte words.
//   ragged rows   (src_ptr != NULL): segment g copies rows [src_ptr[sel[g]], +count) to rows [seg_ptr[g], +count)
//   per-graph row (src_ptr == NULL): segment g copies row sel[g] to row g
// `seg0` = first segment of this launch: 0 for the copy launch (gridDim.x = n_sel), n_sel for the padding launch (gridDim.x = 1)
__global__ void collate_rows_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst,
                                    const int64_t *__restrict__ sel, const int64_t *__restrict__ src_ptr,
                                    const int64_t *__restrict__ seg_ptr, int64_t n_sel, int64_t seg0, int64_t row_words,
                                    int64_t rows_out, uint32_t pad_word, const int32_t *__restrict__ status) {
    if (*status & 3) return;                                   // bad selection / overflow: leave the outputs untouched
    const int64_t g = seg0 + blockIdx.x;
    const int64_t stride = (int64_t)gridDim.y * blockDim.x;
    const int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (g == n_sel) {                                          // padding tail (launched only when the bucket is larger)
        const int64_t used = seg_ptr ? seg_ptr[n_sel] : n_sel;
        const int64_t d0 = used * row_words, n = (rows_out - used) * row_words;
        for (int64_t w = t; w < n; w += stride) dst[d0 + w] = pad_word;
        return;
    }
    int64_t s0, d0, n;
    const int64_t s = sel[g];
    if (src_ptr) {
        s0 = src_ptr[s] * row_words;
        d0 = seg_ptr[g] * row_words;
        n = (seg_ptr[g + 1] - seg_ptr[g]) * row_words;
    } else {
        s0 = s * row_words;
        d0 = g * row_words;
        n = row_words;
    }
    if (((s0 | d0) & 3) == 0) {                                // both sides 16-byte aligned: vector body + word tail
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + s0);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + d0);
        const int64_t n4 = n >> 2;
        for (int64_t w = t; w < n4; w += stride) d4[w] = __ldg(s4 + w);
        for (int64_t w = (n4 << 2) + t; w < n; w += stride) dst[d0 + w] = __ldg(src + s0 + w);
    } else if (((s0 | d0) & 1) == 0) {
        const uint2 *s2 = reinterpret_cast<const uint2 *>(src + s0);
        uint2 *d2 = reinterpret_cast<uint2 *>(dst + d0);
        const int64_t n2 = n >> 1;
        for (int64_t w = t; w < n2; w += stride) d2[w] = __ldg(s2 + w);
        for (int64_t w = (n2 << 1) + t; w < n; w += stride) dst[d0 + w] = __ldg(src + s0 + w);
    } else {
        for (int64_t w = t; w < n; w += stride) dst[d0 + w] = __ldg(src + s0 + w);
    }
}

// index rows: out[r, seg_ptr[g] + e] = in[r, src_ptr[sel[g]] + e] + inc_ptr[g]; padding = -1
__global__ void collate_index_kernel(const int64_t *__restrict__ src, int64_t src_ld, int64_t *__restrict__ dst,
                                     int64_t dst_ld, const int64_t *__restrict__ sel,
                                     const int64_t *__restrict__ src_ptr, const int64_t *__restrict__ seg_ptr,
                                     const int64_t *__restrict__ inc_ptr, int64_t n_sel, int64_t seg0,
                                     const int32_t *__restrict__ status) {
    if (*status & 3) return;
    const int64_t g = seg0 + blockIdx.x;
    const int64_t stride = (int64_t)gridDim.y * blockDim.x;
    const int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (g == n_sel) {
        const int64_t d0 = seg_ptr[n_sel];
        for (int64_t e = d0 + t; e < dst_ld; e += stride) {
            dst[e] = -1;
            dst[dst_ld + e] = -1;
        }
        return;
    }
    const int64_t s0 = src_ptr[sel[g]], d0 = seg_ptr[g], n = seg_ptr[g + 1] - d0, inc = inc_ptr[g];
    for (int64_t e = t; e < n; e += stride) {
        dst[d0 + e] = __ldg(src + s0 + e) + inc;
        dst[dst_ld + d0 + e] = __ldg(src + src_ld + s0 + e) + inc;
    }
}

// batch[n] = g for the atoms of segment g (padding atoms -> the dummy graph n_sel); train_idx[g] = sel[g] (padding -1)
__global__ void collate_batch_vector_kernel(int64_t *__restrict__ batch, int64_t *__restrict__ train_idx,
                                            const int64_t *__restrict__ sel, const int64_t *__restrict__ seg_ptr,
                                            int64_t n_sel, int64_t seg0, int64_t rows_out, int64_t graphs_out,
                                            const int32_t *__restrict__ status) {
    if (*status & 3) return;
    const int64_t g = seg0 + blockIdx.x;
    const int64_t stride = (int64_t)gridDim.y * blockDim.x;
    const int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t d0 = seg_ptr[g], d1 = g == n_sel ? rows_out : seg_ptr[g + 1];
    for (int64_t r = d0 + t; r < d1; r += stride) batch[r] = g;
    if (train_idx && blockIdx.y == 0) {
        if (g < n_sel) {
            if (threadIdx.x == 0) train_idx[g] = sel[g];
        } else {
            for (int64_t r = n_sel + threadIdx.x; r < graphs_out; r += blockDim.x) train_idx[r] = -1;
        }
    }
}

static inline unsigned chunks_for(int64_t words, int per_thread) {
    int64_t c = (words + (int64_t)DP_THREADS * per_thread - 1) / ((int64_t)DP_THREADS * per_thread);
    if (c < 1) c = 1;
    if (c > 4096) c = 4096;
    return (unsigned)c;
}

// ---- bond geometry / features ------------------------------------------------------------------------------------
struct Vec3 {
    double x, y, z;
};

// fetch.py:250-263: unit vector and length of a -> b (+ image), cartesian = dfrac @ lattice (row vector times matrix)
__device__ __forceinline__ Vec3 edge_dir(const double *__restrict__ frac, const double *__restrict__ lat, int64_t a,
                                         int64_t b, int ia, int ib, int ic, double *dist_out) {
    const double d0 = (frac[3 * b] + (double)ia) - frac[3 * a];
    const double d1 = (frac[3 * b + 1] + (double)ib) - frac[3 * a + 1];
    const double d2 = (frac[3 * b + 2] + (double)ic) - frac[3 * a + 2];
    Vec3 c;
    c.x = __dadd_rn(__dadd_rn(__dmul_rn(d0, lat[0]), __dmul_rn(d1, lat[3])), __dmul_rn(d2, lat[6]));
    c.y = __dadd_rn(__dadd_rn(__dmul_rn(d0, lat[1]), __dmul_rn(d1, lat[4])), __dmul_rn(d2, lat[7]));
    c.z = __dadd_rn(__dadd_rn(__dmul_rn(d0, lat[2]), __dmul_rn(d1, lat[5])), __dmul_rn(d2, lat[8]));
    const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(c.x, c.x), __dmul_rn(c.y, c.y)), __dmul_rn(c.z, c.z)));
    if (dist > 0.0) {
        c.x /= dist; c.y /= dist; c.z /= dist;
    } else {
        c.x = c.y = c.z = 0.0;
    }
    *dist_out = dist;
    return c;
}

// one warp per bond: lanes stride the radial basis, lane 0 writes dEN and the direction
__global__ void bond_features_kernel(const double *__restrict__ frac, const double *__restrict__ lattice,
                                     const int64_t *__restrict__ atom_graph, const double *__restrict__ en,
                                     const int64_t *__restrict__ bond_src, const int64_t *__restrict__ bond_dst,
                                     const int32_t *__restrict__ bond_image, int64_t n_bonds,
                                     const double *__restrict__ centers, int n_rbf, double gamma,
                                     double *__restrict__ dirv, float *__restrict__ edge_attr) {
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= n_bonds) return;
    const int64_t i = bond_src[e], j = bond_dst[e];
    const double *lat = lattice + 9 * (atom_graph ? atom_graph[i] : 0);
    double dist;
    const Vec3 d = edge_dir(frac, lat, i, j, bond_image[3 * e], bond_image[3 * e + 1], bond_image[3 * e + 2], &dist);
    float *row = edge_attr + e * (int64_t)(n_rbf + 4);
    for (int k = lane; k < n_rbf; k += 32) {
        const double t = dist - centers[k];
        row[k] = (float)exp(-gamma * __dmul_rn(t, t));
    }
    if (lane == 0) {
        row[n_rbf] = (float)fabs(en[i] - en[j]);
        row[n_rbf + 1] = (float)d.x;
        row[n_rbf + 2] = (float)d.y;
        row[n_rbf + 3] = (float)d.z;
        dirv[3 * e] = d.x; dirv[3 * e + 1] = d.y; dirv[3 * e + 2] = d.z;
    }
}

// ---- line graph -------------------------------------------------------------------------------------------------------
// candidates of bond e1 = (i -> j, im): the bonds leaving j, in bond order = [out_ptr[j], out_ptr[j + 1]) for i-major bond
// lists; the exact reverse image (k == i and kimage == -im) is skipped (fetch.py:424-428)
__device__ __forceinline__ bool is_backtrack(const int64_t *__restrict__ bond_dst, const int32_t *__restrict__ img,
                                             int64_t e2, int64_t i, int a, int b, int c) {
    return bond_dst[e2] == i && img[3 * e2] == -a && img[3 * e2 + 1] == -b && img[3 * e2 + 2] == -c;
}

__global__ void linegraph_count_kernel(const int64_t *__restrict__ bond_src, const int64_t *__restrict__ bond_dst,
                                       const int32_t *__restrict__ img, const int64_t *__restrict__ out_ptr,
                                       int64_t n_bonds, int64_t *__restrict__ counts) {
    const int64_t e1 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e1 >= n_bonds) return;
    const int64_t i = bond_src[e1], j = bond_dst[e1];
    const int a = img[3 * e1], b = img[3 * e1 + 1], c = img[3 * e1 + 2];
    const int64_t lo = out_ptr[j], hi = out_ptr[j + 1];
    int cnt = 0;
    for (int64_t e2 = lo + lane; e2 < hi; e2 += 32) cnt += is_backtrack(bond_dst, img, e2, i, a, b, c) ? 0 : 1;
#pragma unroll
    for (int off = 16; off; off >>= 1) cnt += __shfl_xor_sync(FULL, cnt, off);
    if (lane == 0) counts[e1] = cnt;
}

__global__ void linegraph_fill_kernel(const double *__restrict__ frac, const double *__restrict__ lattice,
                                      const int64_t *__restrict__ atom_graph, const int64_t *__restrict__ graph_bond_ptr,
                                      const int64_t *__restrict__ bond_src, const int64_t *__restrict__ bond_dst,
                                      const int32_t *__restrict__ img, const int64_t *__restrict__ out_ptr,
                                      const double *__restrict__ dirv, const int64_t *__restrict__ angle_ptr,
                                      int64_t n_bonds, const double *__restrict__ centers, int n_ang, double gamma,
                                      int64_t *__restrict__ lg_index, int64_t n_angles, float *__restrict__ lg_attr) {
    const int64_t e1 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e1 >= n_bonds) return;
    const int64_t i = bond_src[e1], j = bond_dst[e1];
    const int a = img[3 * e1], b = img[3 * e1 + 1], c = img[3 * e1 + 2];
    const int64_t gidx = atom_graph ? atom_graph[i] : 0;
    const double *lat = lattice + 9 * gidx;
    const int64_t base = graph_bond_ptr ? graph_bond_ptr[gidx] : 0;        // local bond ids within the graph
    double dist;
    const Vec3 u = edge_dir(frac, lat, j, i, -a, -b, -c, &dist);             // j -> i through the exact reverse image
    const double nu = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(u.x, u.x), __dmul_rn(u.y, u.y)), __dmul_rn(u.z, u.z)));
    const int64_t lo = out_ptr[j], hi = out_ptr[j + 1];
    int64_t slot = angle_ptr[e1];
    const int width = n_ang + 3;
    for (int64_t c0 = lo; c0 < hi; c0 += 32) {
        const int64_t e2 = c0 + lane;
        const bool keep = e2 < hi && !is_backtrack(bond_dst, img, e2, i, a, b, c);
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) {
            const int64_t p = slot + __popc(m & ((1u << lane) - 1u));
            const double vx = dirv[3 * e2], vy = dirv[3 * e2 + 1], vz = dirv[3 * e2 + 2];
            const double nv = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
            double angle = 0.0;
            if (nu != 0.0 && nv != 0.0) {                                     // fetch.py:266-273
                const double dot = __dadd_rn(__dadd_rn(__dmul_rn(u.x, vx), __dmul_rn(u.y, vy)), __dmul_rn(u.z, vz));
                double ct = dot / __dmul_rn(nu, nv);
                ct = fmin(1.0, fmax(-1.0, ct));
                angle = acos(ct);
            }
            float *row = lg_attr + p * width;
            for (int k = 0; k < n_ang; ++k) {
                const double t = angle - centers[k];
                row[k] = (float)exp(-gamma * __dmul_rn(t, t));
            }
            row[n_ang] = (float)angle;
            row[n_ang + 1] = (float)cos(angle);
            row[n_ang + 2] = (float)sin(angle);
            lg_index[p] = e1 - base;
            lg_index[n_angles + p] = e2 - base;
        }
        slot += __popc(m);
    }
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_collate(const alignn_graph_store *store, const int64_t *sel, int64_t n_sel, int lg_inc_bonds,
                              const alignn_batch_out *out, int64_t *seg_ptr, const int64_t *totals, const int64_t *maxima,
                              int32_t *status, void *stream) {
    if (!store || !out || !seg_ptr || !status || !totals || !maxima || n_sel < 0 || (n_sel > 0 && !sel))
        return ALIGNN_ERR_BAD_ARG;
    if (n_sel >= ((int64_t)1 << 31) - 2 || out->n_graphs < n_sel) return ALIGNN_ERR_BAD_SHAPE;
    const int64_t caps[3] = {out->n_nodes, out->n_bonds, out->n_angles};
    for (int k = 0; k < 3; ++k)
        if (totals[k] < 0 || maxima[k] < 0 || totals[k] > caps[k]) return ALIGNN_ERR_BAD_SHAPE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ALIGNN_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    // the host-stated totals size the launches; the device recomputes them from the store and refuses a mismatch
    collate_offsets_kernel<<<1, 1024, 0, st>>>(sel, n_sel, store->n_graphs, store->node_ptr, store->bond_ptr,
                                               store->angle_ptr, seg_ptr, totals[0], totals[1], totals[2], maxima[0],
                                               maxima[1], maxima[2], status);
    ALIGNN_LAUNCH_CHECK();
    const int64_t *node_seg = seg_ptr, *bond_seg = seg_ptr + (n_sel + 1), *angle_seg = seg_ptr + 2 * (n_sel + 1);
    const unsigned segs = (unsigned)n_sel;
    auto rows = [&](const float *src, float *dst, const int64_t *src_ptr, const int64_t *seg, int64_t width,
                    int64_t max_rows, int64_t used_rows, int64_t rows_out, float pad) -> int {
        if (!dst || width == 0) return ALIGNN_OK;
        if (!src) return ALIGNN_ERR_BAD_ARG;
        uint32_t pw;
        memcpy(&pw, &pad, 4);
        if (segs > 0 && max_rows > 0)
            collate_rows_kernel<<<dim3(segs, chunks_for(max_rows * width, 16)), DP_THREADS, 0, st>>>(
                reinterpret_cast<const uint32_t *>(src), reinterpret_cast<uint32_t *>(dst), sel, src_ptr, seg, n_sel, 0,
                width, rows_out, pw, status);
        if (rows_out > used_rows)
            collate_rows_kernel<<<dim3(1, chunks_for((rows_out - used_rows) * width, 8)), DP_THREADS, 0, st>>>(
                reinterpret_cast<const uint32_t *>(src), reinterpret_cast<uint32_t *>(dst), sel, src_ptr, seg, n_sel, n_sel,
                width, rows_out, pw, status);
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? ALIGNN_OK : ALIGNN_ERR_CUDA_BASE + (int)e;
    };
    int rc;
    if ((rc = rows(store->x, out->x, store->node_ptr, node_seg, store->node_dim, maxima[0], totals[0], out->n_nodes, 0.f))) return rc;
    if ((rc = rows(store->edge_attr, out->edge_attr, store->bond_ptr, bond_seg, store->edge_dim, maxima[1], totals[1], out->n_bonds, 0.f))) return rc;
    if ((rc = rows(store->lg_edge_attr, out->lg_edge_attr, store->angle_ptr, angle_seg, store->angle_dim, maxima[2], totals[2], out->n_angles, 0.f))) return rc;
    if ((rc = rows(store->global_x, out->global_x, nullptr, nullptr, store->global_dim, 1, n_sel, out->n_graphs, 0.f))) return rc;
    if ((rc = rows(store->sg_one_hot, out->sg_one_hot, nullptr, nullptr, store->sg_dim, 1, n_sel, out->n_graphs, 0.f))) return rc;
    if ((rc = rows(store->y, out->y, nullptr, nullptr, store->target_dim, 1, n_sel, out->n_graphs, 1.f))) return rc;
    auto index = [&](const int64_t *src, int64_t src_ld, int64_t *dst, int64_t dst_ld, const int64_t *src_ptr,
                     const int64_t *seg, const int64_t *inc, int64_t max_rows, int64_t used) {
        if (!dst) return;
        if (segs > 0 && max_rows > 0)
            collate_index_kernel<<<dim3(segs, chunks_for(max_rows, 4)), DP_THREADS, 0, st>>>(src, src_ld, dst, dst_ld, sel,
                                                                                           src_ptr, seg, inc, n_sel, 0, status);
        if (dst_ld > used)
            collate_index_kernel<<<dim3(1, chunks_for(dst_ld - used, 4)), DP_THREADS, 0, st>>>(src, src_ld, dst, dst_ld, sel,
                                                                                             src_ptr, seg, inc, n_sel, n_sel,
                                                                                             status);
    };
    index(store->edge_index, store->n_bonds, out->edge_index, out->n_bonds, store->bond_ptr, bond_seg, node_seg, maxima[1],
          totals[1]);
    index(store->lg_edge_index, store->n_angles, out->lg_edge_index, out->n_angles, store->angle_ptr, angle_seg,
          lg_inc_bonds ? bond_seg : node_seg, maxima[2], totals[2]);
    if (out->batch) {
        if (segs > 0)
            collate_batch_vector_kernel<<<dim3(segs, chunks_for(maxima[0] > 0 ? maxima[0] : 1, 4)), DP_THREADS, 0, st>>>(
                out->batch, out->train_idx, sel, node_seg, n_sel, 0, out->n_nodes, out->n_graphs, status);
        if (out->n_nodes > totals[0] || out->n_graphs > n_sel)
            collate_batch_vector_kernel<<<dim3(1, chunks_for(out->n_nodes - totals[0] + 1, 4)), DP_THREADS, 0, st>>>(
                out->batch, out->train_idx, sel, node_seg, n_sel, n_sel, out->n_nodes, out->n_graphs, status);
    }
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int alignn_bond_features(const double *frac, const double *lattice, const int64_t *atom_graph, const double *en,
                                    const int64_t *bond_src, const int64_t *bond_dst, const int32_t *bond_image,
                                    int64_t n_bonds, const double *rbf_centers, int n_rbf, double rbf_gamma,
                                    double *dirv, float *edge_attr, void *stream) {
    if (n_bonds < 0 || n_rbf < 0) return ALIGNN_ERR_BAD_ARG;
    if (n_bonds == 0) return ALIGNN_OK;
    if (!frac || !lattice || !en || !bond_src || !bond_dst || !bond_image || !dirv || !edge_attr || (n_rbf && !rbf_centers))
        return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((n_bonds * 32 + DP_THREADS - 1) / DP_THREADS);
    bond_features_kernel<<<blocks, DP_THREADS, 0, st>>>(frac, lattice, atom_graph, en, bond_src, bond_dst, bond_image,
                                                        n_bonds, rbf_centers, n_rbf, rbf_gamma, dirv, edge_attr);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int alignn_linegraph_count(const int64_t *bond_src, const int64_t *bond_dst, const int32_t *bond_image,
                                      const int64_t *out_ptr, int64_t n_bonds, int64_t *counts, void *stream) {
    if (n_bonds < 0) return ALIGNN_ERR_BAD_ARG;
    if (n_bonds == 0) return ALIGNN_OK;
    if (!bond_src || !bond_dst || !bond_image || !out_ptr || !counts) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((n_bonds * 32 + DP_THREADS - 1) / DP_THREADS);
    linegraph_count_kernel<<<blocks, DP_THREADS, 0, st>>>(bond_src, bond_dst, bond_image, out_ptr, n_bonds, counts);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int alignn_linegraph_fill(const double *frac, const double *lattice, const int64_t *atom_graph,
                                     const int64_t *graph_bond_ptr, const int64_t *bond_src, const int64_t *bond_dst,
                                     const int32_t *bond_image, const int64_t *out_ptr, const double *dirv,
                                     const int64_t *angle_ptr, int64_t n_bonds, const double *angle_centers, int n_ang,
                                     double angle_gamma, int64_t *lg_edge_index, int64_t n_angles, float *lg_edge_attr,
                                     void *stream) {
    if (n_bonds < 0 || n_angles < 0 || n_ang < 0) return ALIGNN_ERR_BAD_ARG;
    if (n_bonds == 0 || n_angles == 0) return ALIGNN_OK;
    if (!frac || !lattice || !bond_src || !bond_dst || !bond_image || !out_ptr || !dirv || !angle_ptr || !lg_edge_index ||
        !lg_edge_attr || (n_ang && !angle_centers))
        return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((n_bonds * 32 + DP_THREADS - 1) / DP_THREADS);
    linegraph_fill_kernel<<<blocks, DP_THREADS, 0, st>>>(frac, lattice, atom_graph, graph_bond_ptr, bond_src, bond_dst,
                                                         bond_image, out_ptr, dirv, angle_ptr, n_bonds, angle_centers,
                                                         n_ang, angle_gamma, lg_edge_index, n_angles, lg_edge_attr);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
