// §8f-1: fused, sync-free densification statistics.  Takes the place of
// GaussianModel.update_statistics (/root/reference/model/gaussian.py:188-197), which costs ~10 torch
// kernels and >= 3 host syncs (bool-mask indexing) per step, and extends it to C cameras per call
// with exactly the per-view semantics (each view that sees a Gaussian adds |absgrad|_2 * max_hw
// and 1, and raises max_radii to radii / max_hw).  HBM-bound: 12 B/(camera, Gaussian) read +
// 24 B/Gaussian read-modify-write.
#include "egs_common.cuh"

namespace egs {
constexpr int kStatsThreads = 256;

__global__ void __launch_bounds__(kStatsThreads) densify_stats_kernel(int C, int N, const int32_t* __restrict__ radii,
                                                                       const float2* __restrict__ absgrad, float max_hw,
                                                                       float* __restrict__ max_radii,
                                                                       float* __restrict__ grad_norm_accum,
                                                                       float* __restrict__ counts) {
  const int n = blockIdx.x * kStatsThreads + threadIdx.x;
  if (n >= N) return;
  float mr = max_radii[n], acc = grad_norm_accum[n], cnt = counts[n];
  bool any = false;
  for (int c = 0; c < C; ++c) {
    const size_t idx = (size_t)c * N + n;
    const float r = (float)radii[idx] / max_hw;
    if (r > 0.0f) {
      const float2 g = absgrad[idx];
      mr = fmaxf(mr, r);
      acc = acc + sqrtf(g.x * g.x + g.y * g.y) * max_hw;
      cnt = cnt + 1.0f;
      any = true;
    }
  }
  if (any) {
    max_radii[n] = mr;
    grad_norm_accum[n] = acc;
    counts[n] = cnt;
  }
}
}  // namespace egs

using namespace egs;

extern "C" int egs_densify_stats_update(int32_t C, int32_t N, const int32_t* radii, const float* absgrad, float max_hw,
                                        float* max_radii, float* grad_norm_accum, float* collecting_counts,
                                        egs_stream_t stream) {
  EGS_REQUIRE(C >= 0 && N >= 0, "densify_stats_update: negative sizes");
  EGS_REQUIRE(max_hw > 0.f, "densify_stats_update: max_hw must be positive");
  if (C == 0 || N == 0) return 0;
  densify_stats_kernel<<<(unsigned)ceil_div(N, kStatsThreads), kStatsThreads, 0, (cudaStream_t)stream>>>(
      C, N, radii, reinterpret_cast<const float2*>(absgrad), max_hw, max_radii, grad_norm_accum, collecting_counts);
  return check_launch("densify_stats_kernel");
}
