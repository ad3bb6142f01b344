// Ensemble post-processing as one reduction over the members.
//
// Replaces, per batch, the tail of the member loop of `ensemble_collect` (reference scripts/train.py:876-894; the same
// arithmetic in predict.ensemble_predict, scripts/predict.py:604-623, and evaluate.collect_member_predictions,
// scripts/evaluate.py:244-261) + `sqrt(clamp(var, 1e-12))` (train.py:903) + `apply_conformal_intervals`
// (train.py:1053-1076) + `LogTransformer.inverse_transform_tensor` (train.py:281-296):
//
//   mean_z = mean_m mu_m ;  var_z = mean_m exp(max(logvar_m, floor)) + mean_m mu_m^2 - mean_z^2 ;  std_z = sqrt(max(var_z, 1e-12))
//   lower_z / upper_z = mean_z -+ q * std_z ("scaled") or mean_z -+ q ("absolute")
//   *_orig = exp(z * log_std + log_mean)  (when the log-transform statistics are given, else the z-space values)
//
// One thread per (graph, target); members are read in order 0..M-1 and summed in that order (the order torch's mean over
// dim 0 of an [M, B, T] tensor uses for small M): deterministic.  A few KB of traffic -- the point is that the M member
// outputs never leave the device and no per-member host round trip remains.
#include <math.h>

#include "common.cuh"

namespace alignn {

__global__ void ensemble_post_kernel(const float *__restrict__ means, const float *__restrict__ logvars, int n_members,
                                     int64_t n, int n_targets, float floor_, const float *__restrict__ q, int scaled,
                                     const float *__restrict__ log_means, const float *__restrict__ log_stds,
                                     float *__restrict__ mean_z, float *__restrict__ var_z, float *__restrict__ std_z,
                                     float *__restrict__ mean_o, float *__restrict__ lower_o, float *__restrict__ upper_o) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s_mu = 0.f, s_var = 0.f, s_sq = 0.f;
    for (int m = 0; m < n_members; ++m) {
        const float mu = means[(int64_t)m * n + i];
        s_mu += mu;
        s_sq += mu * mu;
        if (logvars) s_var += expf(fmaxf(logvars[(int64_t)m * n + i], floor_));
    }
    const float inv = 1.0f / (float)n_members;
    const float mz = s_mu * inv;
    const float vz = s_var * inv + s_sq * inv - mz * mz;
    const float sz = sqrtf(fmaxf(vz, 1e-12f));
    mean_z[i] = mz;
    if (var_z) var_z[i] = vz;
    if (std_z) std_z[i] = sz;
    if (!mean_o) return;
    const int t = (int)(i % n_targets);
    float lo = mz, hi = mz;
    if (q) {
        const float w = scaled ? q[t] * sz : q[t];
        lo = mz - w;
        hi = mz + w;
    }
    float mo = mz;
    if (log_means) {
        const float sd = log_stds[t], mu0 = log_means[t];
        mo = expf(mz * sd + mu0);
        lo = expf(lo * sd + mu0);
        hi = expf(hi * sd + mu0);
    }
    mean_o[i] = mo;
    if (lower_o) lower_o[i] = lo;
    if (upper_o) upper_o[i] = hi;
}

}  // namespace alignn

using namespace alignn;

extern "C" int alignn_ensemble_post(const float *member_means, const float *member_logvars, int n_members, int64_t n_graphs,
                                    int n_targets, float min_logvar_floor, const float *q, int scaled,
                                    const float *log_means, const float *log_stds, float *mean_z, float *var_z,
                                    float *std_z, float *mean_orig, float *lower_orig, float *upper_orig, void *stream) {
    if (n_members <= 0 || n_graphs < 0 || n_targets <= 0) return ALIGNN_ERR_BAD_ARG;
    if (n_graphs == 0) return ALIGNN_OK;
    if (!member_means || !mean_z || ((log_means == nullptr) != (log_stds == nullptr))) return ALIGNN_ERR_BAD_ARG;
    const int64_t n = n_graphs * n_targets;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ensemble_post_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(member_means, member_logvars, n_members, n, n_targets,
                                                                     min_logvar_floor, q, scaled, log_means, log_stds,
                                                                     mean_z, var_z, std_z, mean_orig, lower_orig, upper_orig);
    ALIGNN_LAUNCH_CHECK();
    return ALIGNN_OK;
}
