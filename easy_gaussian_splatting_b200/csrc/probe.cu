// Measurement utility (bench.py only): dependent-FMA throughput probe used as the FP32-SIMT roofline
// denominator for the blending kernels, because MEASURED_PEAKS.json only records HBM and bf16
// tensor peaks (BASELINE.md §3 asks the builder to measure the FP32 peak on the box).
#include "egs_common.cuh"

namespace egs {
constexpr int kProbeThreads = 256;
constexpr int kProbeChains = 8;  // independent FMA chains per thread (hides the 4-cycle FMA latency)

__global__ void __launch_bounds__(kProbeThreads) fp32_fma_probe_kernel(int iters, float seed, float* __restrict__ out) {
  float a[kProbeChains];
#pragma unroll
  for (int i = 0; i < kProbeChains; ++i) a[i] = seed + (float)(threadIdx.x + i);
  const float m = 1.0000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < kProbeChains; ++i) a[i] = fmaf(a[i], m, c);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kProbeChains; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;  // never true; keeps the chain alive
}
}  // namespace egs

using namespace egs;

// Launches `blocks` CTAs of 256 threads, each thread doing iters * 32 FMAs.  Returns the number of
// floating point operations issued (2 per FMA) in *host_flops.
extern "C" int egs_probe_fp32_fma(int32_t blocks, int32_t iters, float* out, double* host_flops, egs_stream_t stream) {
  EGS_REQUIRE(blocks > 0 && iters > 0, "probe_fp32_fma: blocks and iters must be positive");
  fp32_fma_probe_kernel<<<blocks, kProbeThreads, 0, (cudaStream_t)stream>>>(iters, 1.0f, out);
  if (host_flops) *host_flops = 2.0 * 32.0 * (double)iters * (double)blocks * kProbeThreads;
  return check_launch("fp32_fma_probe_kernel");
}
