// First layer of the angle encoder and its weight gradient, fused around the streaming conv.
//
// Replaces `angle_encoder[0..1]` = Linear(angle_dim -> H) + ReLU applied to `lg_edge_attr`
// (reference scripts/train.py:360-364, 554) and the matching autograd (weight/bias gradients of that Linear).
// The second Linear of the encoder is folded into the per-layer edge projection (see edgeattn.cu), so
// h1 = relu(W1 a + b1) is the only [L, H] tensor the line-graph path keeps.
//
//   angle_h1_fwd   : h1[e, :] = relu(W1 a_e + b1)        K = angle_dim (11) is far too small for tensor cores:
//                    HBM-bound write of L x H (one 16-byte store per lane), weights staged in shared memory.
//   colsum_outer   : dW1[c, f] = sum_e dpre[e, c] a[e, f];  db1[c] = sum_e dpre[e, c]   (dpre already ReLU-masked)
//                    register accumulation per (channel, feature) over a persistent grid, fixed-order
//                    cross-CTA reduction (no atomics).
#include "common.cuh"

namespace alignn {

constexpr int ANG_THREADS = 256;
constexpr int ANG_WARPS = ANG_THREADS / 32;
constexpr int ANG_MAX_IN = 16;
constexpr int ANG_EDGES_PER_ITER = 4;
constexpr int ANG_PARTIAL_BLOCKS = 296;   // 2 x 148 SMs

template <typename T>
__global__ void __launch_bounds__(ANG_THREADS)
angle_h1_fwd_kernel(const float *__restrict__ a, const float *__restrict__ w1, const float *__restrict__ b1,
                    T *__restrict__ h1, int64_t n_edges, int in_dim, int hidden) {
    // w1 is [hidden, in_dim] (nn.Linear layout); stage it transposed: wt[f][c]
    extern __shared__ float wt[];
    for (int i = threadIdx.x; i < hidden * in_dim; i += ANG_THREADS) {
        const int c = i / in_dim, f = i - c * in_dim;
        wt[f * hidden + c] = w1[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = (int64_t)blockIdx.x * ANG_WARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * ANG_WARPS;
    const int groups = hidden / 256;   // hidden is a multiple of 256 on this path: lane owns 8 channels per group
    for (int g = 0; g < groups; ++g) {
        const int ch = g * 256 + lane * 8;
        const F8 bf = ld8(b1 + ch);
        for (int64_t e0 = warp_id * ANG_EDGES_PER_ITER; e0 < n_edges; e0 += n_warps * ANG_EDGES_PER_ITER) {
            F8 acc[ANG_EDGES_PER_ITER];
#pragma unroll
            for (int u = 0; u < ANG_EDGES_PER_ITER; ++u) acc[u] = bf;
            for (int f = 0; f < in_dim; ++f) {
                const float4 wa = *reinterpret_cast<const float4 *>(wt + f * hidden + ch);
                const float4 wb = *reinterpret_cast<const float4 *>(wt + f * hidden + ch + 4);
#pragma unroll
                for (int u = 0; u < ANG_EDGES_PER_ITER; ++u) {
                    const int64_t e = e0 + u;
                    const float x = e < n_edges ? __ldg(a + e * in_dim + f) : 0.f;   // warp-broadcast load
                    acc[u].v[0] = fmaf(wa.x, x, acc[u].v[0]); acc[u].v[1] = fmaf(wa.y, x, acc[u].v[1]);
                    acc[u].v[2] = fmaf(wa.z, x, acc[u].v[2]); acc[u].v[3] = fmaf(wa.w, x, acc[u].v[3]);
                    acc[u].v[4] = fmaf(wb.x, x, acc[u].v[4]); acc[u].v[5] = fmaf(wb.y, x, acc[u].v[5]);
                    acc[u].v[6] = fmaf(wb.z, x, acc[u].v[6]); acc[u].v[7] = fmaf(wb.w, x, acc[u].v[7]);
                }
            }
#pragma unroll
            for (int u = 0; u < ANG_EDGES_PER_ITER; ++u) {
                const int64_t e = e0 + u;
                if (e < n_edges) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[u].v[c] = fmaxf(acc[u].v[c], 0.f);
                    st8(h1 + e * hidden + ch, acc[u]);
                }
            }
        }
    }
}

// partials: [ANG_PARTIAL_BLOCKS][in_dim + 1][256] ; row f < in_dim: dW1[:, f], row in_dim: db1
template <typename T, int IN>
__global__ void __launch_bounds__(ANG_THREADS)
colsum_outer_kernel(const T *__restrict__ dpre, const float *__restrict__ a, float *__restrict__ partials,
                    int64_t n_edges) {
    __shared__ float red[ANG_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = lane * 8;
    const int64_t warp_id = (int64_t)blockIdx.x * ANG_WARPS + warp;
    const int64_t n_warps = (int64_t)gridDim.x * ANG_WARPS;
    float acc[IN + 1][8];
#pragma unroll
    for (int f = 0; f <= IN; ++f)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[f][c] = 0.f;
    // contiguous slab per warp: fixed summation order
    const int64_t per = (n_edges + n_warps - 1) / n_warps;
    const int64_t lo = warp_id * per, hi = min(n_edges, lo + per);
    for (int64_t e = lo; e < hi; ++e) {
        const F8 d = ld8_stream(dpre + e * 256 + ch);
        float x[IN];
#pragma unroll
        for (int f = 0; f < IN; ++f) x[f] = __ldg(a + e * IN + f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int f = 0; f < IN; ++f) acc[f][c] = fmaf(d.v[c], x[f], acc[f][c]);
            acc[IN][c] += d.v[c];
        }
    }
#pragma unroll
    for (int f = 0; f <= IN; ++f) {
#pragma unroll
        for (int c = 0; c < 8; ++c) red[warp][ch + c] = acc[f][c];
        __syncthreads();
        {
            const int c = threadIdx.x;   // 256 threads, 256 channels
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < ANG_WARPS; ++w) s += red[w][c];
            partials[((int64_t)blockIdx.x * (IN + 1) + f) * 256 + c] = s;
        }
        __syncthreads();
    }
}

__global__ void reduce_blocks_kernel(const float *__restrict__ partials, float *__restrict__ out, int n_blocks, int width) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width) return;
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += partials[(int64_t)b * width + i];
    out[i] = s;
}

template <typename T>
static int launch_colsum(const void *dpre, const float *a, float *partials, float *out, int64_t n_edges, int in_dim,
                         cudaStream_t st) {
#define CS(IN) colsum_outer_kernel<T, IN><<<ANG_PARTIAL_BLOCKS, ANG_THREADS, 0, st>>>((const T *)dpre, a, partials, n_edges)
    switch (in_dim) {
        case 1: CS(1); break;   case 2: CS(2); break;   case 3: CS(3); break;   case 4: CS(4); break;
        case 5: CS(5); break;   case 6: CS(6); break;   case 7: CS(7); break;   case 8: CS(8); break;
        case 9: CS(9); break;   case 10: CS(10); break; case 11: CS(11); break; case 12: CS(12); break;
        case 13: CS(13); break; case 14: CS(14); break; case 15: CS(15); break; case 16: CS(16); break;
        default: return ALIGNN_ERR_BAD_SHAPE;
    }
#undef CS
    ALIGNN_LAUNCH_CHECK();
    const int width = (in_dim + 1) * 256;
    reduce_blocks_kernel<<<(width + 255) / 256, 256, 0, st>>>(partials, out, ANG_PARTIAL_BLOCKS, width);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_angle_supported(int in_dim, int hidden) {
    return in_dim >= 1 && in_dim <= ANG_MAX_IN && hidden == 256;
}

extern "C" int alignn_angle_h1_fwd(const float *a, const float *w1, const float *b1, void *h1, int64_t n_edges,
                                   int in_dim, int hidden, int dtype, void *stream) {
    if (!alignn_angle_supported(in_dim, hidden)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_edges < 0) return ALIGNN_ERR_BAD_ARG;
    if (n_edges == 0) return ALIGNN_OK;
    if (!a || !w1 || !b1 || !h1 || !aligned16(b1) || !aligned16(h1)) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)hidden * in_dim * sizeof(float);
    const int64_t warps_needed = (n_edges + ANG_EDGES_PER_ITER - 1) / ANG_EDGES_PER_ITER;
    const int64_t blocks = warps_needed / ANG_WARPS + 1;
    const unsigned grid = (unsigned)(blocks < 148 * 4 ? blocks : 148 * 4);
    if (dtype == ALIGNN_F32)
        angle_h1_fwd_kernel<float><<<grid, ANG_THREADS, smem, st>>>(a, w1, b1, (float *)h1, n_edges, in_dim, hidden);
    else if (dtype == ALIGNN_BF16)
        angle_h1_fwd_kernel<__nv_bfloat16><<<grid, ANG_THREADS, smem, st>>>(a, w1, b1, (__nv_bfloat16 *)h1, n_edges, in_dim, hidden);
    else
        return ALIGNN_ERR_BAD_DTYPE;
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}

extern "C" int64_t alignn_angle_partial_floats(int in_dim) { return (int64_t)ANG_PARTIAL_BLOCKS * (in_dim + 1) * 256; }

// out: f32 [(in_dim + 1) * 256] = dW1^T rows (out[f*256 + c] = dW1[c, f]) followed by db1
extern "C" int alignn_angle_h1_bwd(const void *dpre, const float *a, float *partials, float *out, int64_t n_edges,
                                   int in_dim, int hidden, int dtype, void *stream) {
    if (!alignn_angle_supported(in_dim, hidden)) return ALIGNN_ERR_BAD_SHAPE;
    if (n_edges < 0 || !partials || !out) return ALIGNN_ERR_BAD_ARG;
    if (n_edges > 0 && (!dpre || !a || !aligned16(dpre))) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == ALIGNN_F32) return launch_colsum<float>(dpre, a, partials, out, n_edges, in_dim, st);
    if (dtype == ALIGNN_BF16) return launch_colsum<__nv_bfloat16>(dpre, a, partials, out, n_edges, in_dim, st);
    return ALIGNN_ERR_BAD_DTYPE;
}
