// §8f-3: fused Adam over the reference's parameter groups (takes the place of torch.optim.Adam in
// /root/reference/model/gaussian.py:389-412 / train.py:156-157; same update rule as torch's default
// Adam: no weight decay, no amsgrad, bias-corrected, eps added after the sqrt).  One launch for all groups, 128-bit accesses:
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
  int64_t end[kAdamMaxGroups];    // exclusive end of the group in the concatenated index space (units of 4 elements)
  int64_t numel[kAdamMaxGroups];
  float lr[kAdamMaxGroups];
  bool aligned[kAdamMaxGroups];   // p, g, m, v all 16-byte aligned
  int n_groups;
};

// One Adam update, torch's default rule: denom = sqrt(v) / sqrt(bias2) + eps ; p -= (lr / bias1) * m / denom
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float beta1, float beta2, float eps,
                                            float step_size, float bias2_sqrt) {
  m = beta1 * m + (1.0f - beta1) * g;
  v = beta2 * v + (1.0f - beta2) * g * g;
  const float denom = sqrtf(v) / bias2_sqrt + eps;
  p -= step_size * (m / denom);
}

// The index space is the groups' elements in units of FOUR (every group rounded up to whole units): a thread moves
// 7 x 16 bytes per unit with 128-bit loads and stores (the scalar form reached 3.6 TB/s on the 177 M parameters of
// the 3 M-Gaussian training step, 1.37 ms of an 18 ms step on 8 GPUs).  A unit that crosses the end of its group, or a
// group whose tensors are not 16-byte aligned, takes the scalar path.
__global__ void __launch_bounds__(kAdamThreads) fused_adam_kernel(const AdamGroups gs, float beta1, float beta2, float eps,
                                                                   float bias1, float bias2_sqrt) {
  const int64_t total = gs.end[gs.n_groups - 1];  // in units of 4 elements
  for (int64_t u = (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; u < total; u += (int64_t)gridDim.x * kAdamThreads) {
    int k = 0;
#pragma unroll
    for (int j = 0; j < kAdamMaxGroups - 1; ++j) k += (j < gs.n_groups - 1 && u >= gs.end[j]) ? 1 : 0;
    const int64_t off = (u - (k > 0 ? gs.end[k - 1] : 0)) * 4;
    const float step_size = gs.lr[k] / bias1;
    float* pp = gs.p[k] + off;
    const float* gp = gs.g[k] + off;
    float* mp = gs.m[k] + off;
    float* vp = gs.v[k] + off;
    const int64_t left = gs.numel[k] - off;
    if (left >= 4 && gs.aligned[k]) {
      float4 p4 = *reinterpret_cast<float4*>(pp), m4 = *reinterpret_cast<float4*>(mp), v4 = *reinterpret_cast<float4*>(vp);
      const float4 g4 = *reinterpret_cast<const float4*>(gp);
      adam_update(p4.x, g4.x, m4.x, v4.x, beta1, beta2, eps, step_size, bias2_sqrt);
      adam_update(p4.y, g4.y, m4.y, v4.y, beta1, beta2, eps, step_size, bias2_sqrt);
      adam_update(p4.z, g4.z, m4.z, v4.z, beta1, beta2, eps, step_size, bias2_sqrt);
      adam_update(p4.w, g4.w, m4.w, v4.w, beta1, beta2, eps, step_size, bias2_sqrt);
      *reinterpret_cast<float4*>(mp) = m4;
      *reinterpret_cast<float4*>(vp) = v4;
      *reinterpret_cast<float4*>(pp) = p4;
    } else {
      for (int e = 0; e < 4 && e < left; ++e) {
        float p = pp[e], m = mp[e], v = vp[e];
        adam_update(p, gp[e], m, v, beta1, beta2, eps, step_size, bias2_sqrt);
        mp[e] = m; vp[e] = v; pp[e] = p;
      }
    }
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
    run += ceil_div(numels[k], (int64_t)4);
    gs.end[k] = run;
    gs.numel[k] = numels[k];
    gs.lr[k] = lrs[k];
    gs.aligned[k] = ((reinterpret_cast<uintptr_t>(params[k]) | reinterpret_cast<uintptr_t>(grads[k]) |
                      reinterpret_cast<uintptr_t>(exp_avg[k]) | reinterpret_cast<uintptr_t>(exp_avg_sq[k])) & 15) == 0;
  }
  for (int k = n_groups; k < kAdamMaxGroups; ++k) { gs.p[k] = nullptr; gs.g[k] = nullptr; gs.m[k] = nullptr; gs.v[k] = nullptr; gs.end[k] = run; gs.numel[k] = 0; gs.lr[k] = 0.f; gs.aligned[k] = false; }
  gs.n_groups = n_groups;
  if (run == 0) return 0;
  const double b1 = 1.0 - pow((double)beta1, (double)step), b2 = 1.0 - pow((double)beta2, (double)step);
  int64_t blocks = ceil_div(run, kAdamThreads);
  if (blocks > 148 * 32) blocks = 148 * 32;
  fused_adam_kernel<<<(unsigned)blocks, kAdamThreads, 0, (cudaStream_t)stream>>>(gs, beta1, beta2, eps, (float)b1, (float)sqrt(b2));
  return check_launch("fused_adam_kernel");
}
