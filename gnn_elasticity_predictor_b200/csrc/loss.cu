// Heteroscedastic Gaussian NLL + log-sigma L2 of the reference's training loop, forward and gradient in ONE launch.
//
// Replaces the ~20 elementwise / reduction kernels (and as many again in autograd's backward) of
// `train_epoch_hetero` (reference scripts/train.py:655-681):
//     lv = clamp(logvar, min=floor);  nll = 0.5 (lv + (mean - y)^2 / exp(lv)) [* w_b];
//     loss = mean_b mean_t nll + l2 * mean_{b,t} (0.5 lv)^2
// with an optional per-graph mask (dummy graphs of a shape-bucket-padded batch count for nothing: means run over the real
// graphs) and optional per-sample weights (`sample_weights`, train.py:661-675).  Outputs the scalar loss and
// d loss / d mean, d loss / d logvar (clamp passes the gradient where logvar >= floor, like torch.clamp).
// One CTA, fixed-order tree reductions: deterministic.  B x T is a few hundred numbers -- the point is the launch count.
#include <math.h>

#include "common.cuh"

namespace alignn {

constexpr int NLL_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float x, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NLL_THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(NLL_THREADS)
gaussian_nll_kernel(const float *__restrict__ mean, const float *__restrict__ logvar, const float *__restrict__ target,
                    const float *__restrict__ mask, const float *__restrict__ weight, int64_t n_graphs, int n_targets,
                    float floor_, float l2, float *__restrict__ loss, float *__restrict__ dmean,
                    float *__restrict__ dlogvar) {
    __shared__ float red[NLL_THREADS / 32];
    const int64_t n = n_graphs * n_targets;
    float cnt = 0.f;
    if (mask) {
        for (int64_t b = threadIdx.x; b < n_graphs; b += NLL_THREADS) cnt += mask[b];
        cnt = fmaxf(block_sum_256(cnt, red), 1.0f);
    } else {
        cnt = (float)n_graphs;
    }
    const float inv = 1.0f / (cnt * (float)n_targets);
    float acc_nll = 0.f, acc_l2 = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += NLL_THREADS) {
        const int64_t b = i / n_targets;
        const float mk = mask ? mask[b] : 1.0f, w = weight ? weight[b] : 1.0f;
        const float raw = logvar[i];
        const float lv = fmaxf(raw, floor_);
        const float diff = mean[i] - target[i];
        const float iv = expf(-lv);
        acc_nll += mk * w * 0.5f * (lv + diff * diff * iv);
        acc_l2 += mk * 0.25f * lv * lv;
        if (dmean) dmean[i] = mk * w * diff * iv * inv;
        if (dlogvar) dlogvar[i] = raw >= floor_ ? mk * (w * 0.5f * (1.0f - diff * diff * iv) + l2 * 0.5f * lv) * inv : 0.f;
    }
    const float s_nll = block_sum_256(acc_nll, red);
    const float s_l2 = block_sum_256(acc_l2, red);
    if (threadIdx.x == 0) loss[0] = (s_nll + l2 * s_l2) * inv;
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_gaussian_nll(const float *mean, const float *logvar, const float *target, const float *mask,
                                   const float *weight, int64_t n_graphs, int n_targets, float min_logvar_floor,
                                   float log_sigma_l2, float *loss, float *dmean, float *dlogvar, void *stream) {
    if (n_graphs <= 0 || n_targets <= 0 || !mean || !logvar || !target || !loss) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    gaussian_nll_kernel<<<1, NLL_THREADS, 0, st>>>(mean, logvar, target, mask, weight, n_graphs, n_targets, min_logvar_floor,
                                                  log_sigma_l2, loss, dmean, dlogvar);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
