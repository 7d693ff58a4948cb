// §8f-3: fused Adam over the reference's parameter groups (takes the place of torch.optim.Adam in
// /root/reference/model/gaussian.py:389-412 / train.py:156-157; same update rule as torch's default
// Adam: no weight decay, no amsgrad, bias-corrected, eps added after the sqrt).  One launch for all groups:
// 28 B/element of HBM traffic (read p, g, m, v; write p, m, v) and nothing else.
#include "egs_common.cuh"

namespace egs {
constexpr int kAdamThreads = 256;
constexpr int kAdamMaxGroups = 8;

struct AdamGroups {
  float* p[kAdamMaxGroups];
  const float* g[kAdamMaxGroups];
  float* m[kAdamMaxGroups];
  float* v[kAdamMaxGroups];
  int64_t end[kAdamMaxGroups];  // exclusive end of the group in the concatenated index space
  float lr[kAdamMaxGroups];
  int n_groups;
};

__global__ void __launch_bounds__(kAdamThreads) fused_adam_kernel(const AdamGroups gs, float beta1, float beta2, float eps,
                                                                   float bias1, float bias2_sqrt) {
  const int64_t total = gs.end[gs.n_groups - 1];
  for (int64_t i = (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kAdamThreads) {
    int k = 0;
#pragma unroll
    for (int j = 0; j < kAdamMaxGroups - 1; ++j) k += (j < gs.n_groups - 1 && i >= gs.end[j]) ? 1 : 0;
    const int64_t off = i - (k > 0 ? gs.end[k - 1] : 0);
    const float g = gs.g[k][off];
    const float m = beta1 * gs.m[k][off] + (1.0f - beta1) * g;
    const float v = beta2 * gs.v[k][off] + (1.0f - beta2) * g * g;
    gs.m[k][off] = m;
    gs.v[k][off] = v;
    // torch: denom = sqrt(v) / sqrt(bias2) + eps ; p -= (lr / bias1) * m / denom
    const float denom = sqrtf(v) / bias2_sqrt + eps;
    gs.p[k][off] -= (gs.lr[k] / bias1) * (m / denom);
  }
}
}  // namespace egs

using namespace egs;

extern "C" int egs_fused_adam(int32_t n_groups, float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, const int64_t* numels, const float* lrs, float beta1,
                              float beta2, float eps, int64_t step, egs_stream_t stream) {
  EGS_REQUIRE(n_groups >= 1 && n_groups <= kAdamMaxGroups, "fused_adam: n_groups=%d out of [1,%d]", n_groups, kAdamMaxGroups);
  EGS_REQUIRE(step >= 1, "fused_adam: step must be >= 1");
  AdamGroups gs;
  int64_t run = 0;
  for (int k = 0; k < n_groups; ++k) {
    EGS_REQUIRE(numels[k] >= 0, "fused_adam: negative numel");
    gs.p[k] = params[k]; gs.g[k] = grads[k]; gs.m[k] = exp_avg[k]; gs.v[k] = exp_avg_sq[k];
    run += numels[k];
    gs.end[k] = run;
    gs.lr[k] = lrs[k];
  }
  for (int k = n_groups; k < kAdamMaxGroups; ++k) { gs.p[k] = nullptr; gs.g[k] = nullptr; gs.m[k] = nullptr; gs.v[k] = nullptr; gs.end[k] = run; gs.lr[k] = 0.f; }
  gs.n_groups = n_groups;
  if (run == 0) return 0;
  const double b1 = 1.0 - pow((double)beta1, (double)step), b2 = 1.0 - pow((double)beta2, (double)step);
  int64_t blocks = ceil_div(run, kAdamThreads);
  if (blocks > 148 * 32) blocks = 148 * 32;
  fused_adam_kernel<<<(unsigned)blocks, kAdamThreads, 0, (cudaStream_t)stream>>>(gs, beta1, beta2, eps, (float)b1, (float)sqrt(b2));
  return check_launch("fused_adam_kernel");
}
