// Fused gradient-norm clip + AdamW over ONE flat fp32 parameter bucket (+ bf16 shadow copy for the tensor-core
// operands), deterministic and CUDA-graph capturable.
//
// Replaces the optimizer step that follows the hot path in the reference (scripts/train.py:690-699:
// `clip_grad_norm_(model.parameters(), 5.0)`; `AdamW(param_groups, weight_decay, fused=True).step()` built at
// :1516-1540 with two learning-rate groups: base + mean heads, and the log-variance heads) -- SURVEY.md 8(f) N1.
//
//   kernel 1  grad_sumsq:  per-CTA partial sums of g^2 in a fixed order; thread 0 of CTA 0 advances the step counter
//   kernel 2  adamw_step:  every CTA folds the partials in the same order -> total norm -> clip coefficient
//                          min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_), then the decoupled
//                          weight-decay Adam update of its slice; p, m, v in place, bf16 shadow written.
// Hyper-parameters that change during training (the two learning rates) are read from DEVICE memory so a captured
// graph sees the scheduler's updates; the step count lives on the device for the same reason.
#include <math.h>

#include "common.cuh"

namespace alignn {

constexpr int OPT_THREADS = 256;
constexpr int OPT_BLOCKS = 296;   // 2 x 148 SMs

__global__ void __launch_bounds__(OPT_THREADS)
grad_sumsq_kernel(const float *__restrict__ g, int64_t n, float *__restrict__ partials, float *__restrict__ step) {
    __shared__ float red[OPT_THREADS / 32];
    float acc = 0.f;
    const int64_t n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * OPT_THREADS) {
        const float4 x = __ldg(g4 + i);
        acc = fmaf(x.x, x.x, acc); acc = fmaf(x.y, x.y, acc); acc = fmaf(x.z, x.z, acc); acc = fmaf(x.w, x.w, acc);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float x = g[(n4 << 2) + threadIdx.x];
        acc = fmaf(x, x, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
        partials[blockIdx.x] = t;
        if (blockIdx.x == 0 && step) *step += 1.0f;
    }
}

struct AdamwParams {
    float *p, *m, *v;
    const float *g;
    __nv_bfloat16 *shadow;      // may be null
    const float *partials;      // [OPT_BLOCKS]
    const float *step;          // [1] (already advanced)
    const float *lr;            // [2] device: lr of [0, split) and of [split, n)
    float *norm_out;            // [1] total gradient norm before clipping (may be null)
    int64_t n, split;
    float beta1, beta2, eps, weight_decay, max_norm, grad_scale;
};

__global__ void __launch_bounds__(OPT_THREADS)
adamw_step_kernel(const AdamwParams P) {
    __shared__ float s_coef;
    if (threadIdx.x < 32) {
        float t = 0.f;
        for (int i = threadIdx.x; i < OPT_BLOCKS; i += 32) t += P.partials[i];   // fixed order per lane ...
        t = warp_sum(t);                                                         // ... and fixed butterfly
        if (threadIdx.x == 0) {
            const float norm = sqrtf(t) * P.grad_scale;
            s_coef = P.max_norm > 0.f ? fminf(1.0f, P.max_norm / (norm + 1e-6f)) * P.grad_scale : P.grad_scale;
            if (blockIdx.x == 0 && P.norm_out) *P.norm_out = norm;
        }
    }
    __syncthreads();
    const float coef = s_coef;
    const float t = *P.step;
    const float bc1 = 1.0f - powf(P.beta1, t), bc2_rsqrt = rsqrtf(1.0f - powf(P.beta2, t));
    const float lr0 = P.lr[0], lr1 = P.lr[1];
    for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * OPT_THREADS) {
        const float lr = i < P.split ? lr0 : lr1;
        const float g = P.g[i] * coef;
        float p = P.p[i];
        p *= 1.0f - lr * P.weight_decay;
        const float m = P.beta1 * P.m[i] + (1.0f - P.beta1) * g;
        const float v = P.beta2 * P.v[i] + (1.0f - P.beta2) * g * g;
        const float denom = sqrtf(v) * bc2_rsqrt + P.eps;
        p -= (lr / bc1) * (m / denom);
        P.p[i] = p;
        P.m[i] = m;
        P.v[i] = v;
        if (P.shadow) P.shadow[i] = __float2bfloat16_rn(p);
    }
}

}  // namespace alignn

using namespace alignn;

extern "C" int64_t alignn_adamw_partial_floats(void) { return OPT_BLOCKS; }

extern "C" int alignn_clip_adamw_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                                      void *shadow_bf16, float *partials, float *step, const float *lr2,
                                      float *norm_out, int64_t n, int64_t split,
                                      float beta1, float beta2, float eps, float weight_decay, float max_norm,
                                      float grad_scale, void *stream) {
    if (n < 0 || split < 0 || split > n) return ALIGNN_ERR_BAD_ARG;
    if (n == 0) return ALIGNN_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !partials || !step || !lr2) return ALIGNN_ERR_BAD_ARG;
    if (!aligned16(grads)) return ALIGNN_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    grad_sumsq_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, st>>>(grads, n, partials, step);
    ALIGNN_LAUNCH_CHECK();
    AdamwParams p;
    p.p = params; p.m = exp_avg; p.v = exp_avg_sq; p.g = grads; p.shadow = (__nv_bfloat16 *)shadow_bf16;
    p.partials = partials; p.step = step; p.lr = lr2; p.norm_out = norm_out; p.n = n; p.split = split;
    p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.weight_decay = weight_decay; p.max_norm = max_norm;
    p.grad_scale = grad_scale;
    adamw_step_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, st>>>(p);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
