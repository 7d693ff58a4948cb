// §8f-4: fused photometric loss of the reference's LossComputer (/root/reference/model/gaussian.py:415-453):
//   x      = mask * gt + (1 - mask) * render                      (gaussian.py:428-429)
//   l1     = mean |x - gt|                                         (gaussian.py:447-448)
//   ssim   = 1 - SSIM(gt, x), torchmetrics StructuralSimilarityIndexMeasure(data_range=1.0) defaults:
//            11-tap Gaussian window (sigma 1.5), k1 = 0.01, k2 = 0.03, reflect padding followed by a crop of the
//            padded border — i.e. exactly a VALID 11x11 window over the unpadded image, mean over the
//            (H-10) x (W-10) x 3 interior values                    (gaussian.py:419, 450-453)
//   total  = (1 - lambda) * l1 + lambda * ssim                     (gaussian.py:437)
// upstream this is ~25 torch kernels forward + as many backward (five depthwise 11x11 convolutions over a
// 5-image stack, pads, crops, elementwise chains).  Here: ONE forward kernel (separable window in shared memory,
// per-image sums, and the three partial-derivative maps dS/dmu_x, dS/dE[x^2], dS/dE[xy]) and ONE backward kernel
// (separable window over the three maps -> d total / d render).  Images are [C,H,W,3] fp32 exactly as
// rasterization() returns them (no permute), HBM traffic 28 B/pixel read + 36 B/pixel maps forward,
// 64 B/pixel read + 12 B/pixel written backward.
#include <math.h>

#include "egs_common.cuh"

namespace egs {

constexpr int kWin = 11;          // window taps
constexpr int kHalo = kWin - 1;   // 10
constexpr int kTX = 32, kTY = 16; // output tile (pixels)
constexpr int kLossThreads = 256;
constexpr int kInCols = (kTX + kHalo) * 3;   // 126 interleaved floats per staged row
constexpr int kInRows = kTY + kHalo;         // 26
constexpr int kOutCols = kTX * 3;            // 96
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;  // (k1 * data_range)^2, (k2 * data_range)^2, data_range = 1

struct Window { float w[kWin]; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums[c] = {sum |x - gt| over H*W*3, sum of the SSIM map over (H-10)*(W-10)*3}; maps (nullable) = 3 planes of
// [C, H-10, W-10, 3]: dS/dmu_x, dS/dE[x^2], dS/dE[xy]
__global__ void __launch_bounds__(kLossThreads) l1_ssim_fwd_kernel(int H, int W, const float* __restrict__ render,
                                                                   const float* __restrict__ gt,
                                                                   const float* __restrict__ mask, float* __restrict__ maps,
                                                                   double* __restrict__ sums, const Window win) {
  extern __shared__ __align__(16) float smem_loss[];
  float* sx = smem_loss;                          // [kInRows][kInCols]
  float* sy = sx + kInRows * kInCols;             // [kInRows][kInCols]
  float* hz = sy + kInRows * kInCols;             // [5][kInRows][kOutCols]
  __shared__ float red[2][kLossThreads / 32];
  const int Hi = H - kHalo, Wi = W - kHalo;
  const int c = blockIdx.z;
  const int y0 = blockIdx.y * kTY, x0 = blockIdx.x * kTX;
  const bool last_y = blockIdx.y == gridDim.y - 1, last_x = blockIdx.x == gridDim.x - 1;
  const size_t img = (size_t)c * H * W * 3;
  const float* rp = render + img;
  const float* gp = gt + img;
  const float* mp = mask ? mask + (size_t)c * H * W : nullptr;
  const int tid = threadIdx.x;

  // stage the (kTY+10) x (kTX+10) window of x and gt; every image pixel is counted into the L1 sum by exactly one block
  float l1 = 0.f;
  for (int i = tid; i < kInRows * kInCols; i += kLossThreads) {
    const int row = i / kInCols, col = i - row * kInCols;
    const int gy = y0 + row, gx3 = x0 * 3 + col;
    float x = 0.f, y = 0.f;
    if (gy < H && gx3 < W * 3) {
      const size_t o = (size_t)gy * W * 3 + gx3;
      const float r = __ldg(rp + o);
      y = __ldg(gp + o);
      const float m = mp ? __ldg(mp + (size_t)gy * W + gx3 / 3) : 0.f;
      x = m * y + (1.0f - m) * r;
      if ((row < kTY || last_y) && (col < kOutCols || last_x)) l1 += fabsf(x - y);
    }
    // staged values are centred on 0.5: variances / covariances are shift invariant and lose ~4x fewer bits
    // to the E[x^2] - mu^2 cancellation for images in [0, 1]
    sx[i] = x - 0.5f;
    sy[i] = y - 0.5f;
  }
  __syncthreads();
  // horizontal pass: 5 windowed sums per staged row and output column (interleaved channels: tap stride 3)
  for (int i = tid; i < kInRows * kOutCols; i += kLossThreads) {
    const int row = i / kOutCols, j = i - row * kOutCols;
    const float* ax = sx + row * kInCols + j;
    const float* ay = sy + row * kInCols + j;
    float hx = 0.f, hy = 0.f, hxx = 0.f, hyy = 0.f, hxy = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float a = ax[3 * k], b = ay[3 * k], w = win.w[k];
      const float wa = w * a, wb = w * b;
      hx += wa; hy += wb; hxx += wa * a; hyy += wb * b; hxy += wa * b;
    }
    hz[0 * kInRows * kOutCols + i] = hx;
    hz[1 * kInRows * kOutCols + i] = hy;
    hz[2 * kInRows * kOutCols + i] = hxx;
    hz[3 * kInRows * kOutCols + i] = hyy;
    hz[4 * kInRows * kOutCols + i] = hxy;
  }
  __syncthreads();
  // vertical pass + SSIM + derivative maps
  float ssum = 0.f;
  for (int i = tid; i < kTY * kOutCols; i += kLossThreads) {
    const int y = i / kOutCols, j = i - y * kOutCols;
    const int py = y0 + y, px = x0 + j / 3;
    if (py >= Hi || px >= Wi) continue;
    float v[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float* col = hz + q * kInRows * kOutCols + y * kOutCols + j;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) acc += win.w[k] * col[k * kOutCols];
      v[q] = acc;
    }
    const float sxx = v[2] - v[0] * v[0], syy = v[3] - v[1] * v[1], sxy = v[4] - v[0] * v[1];
    const float mux = v[0] + 0.5f, muy = v[1] + 0.5f;
    const float A1 = 2.f * mux * muy + kC1, A2 = 2.f * sxy + kC2;
    const float B1 = mux * mux + muy * muy + kC1, B2 = sxx + syy + kC2;
    const float rB1 = 1.0f / B1, rB2 = 1.0f / B2;
    const float S = A1 * A2 * rB1 * rB2;
    ssum += S;
    if (maps) {
      const size_t plane = (size_t)gridDim.z * Hi * Wi * 3;
      const size_t o = ((size_t)c * Hi + py) * Wi * 3 + (size_t)x0 * 3 + j;
      maps[o] = 2.f * muy * (A2 - A1) * rB1 * rB2 - 2.f * mux * S * (rB1 - rB2);  // dS/dmu_x (through sxx, sxy too)
      maps[plane + o] = -S * rB2;                                                // dS/dE[x^2]
      maps[2 * plane + o] = 2.f * A1 * rB1 * rB2;                                // dS/dE[xy]
    }
  }
  l1 = warp_sum(l1);
  ssum = warp_sum(ssum);
  if ((tid & 31) == 0) { red[0][tid >> 5] = l1; red[1][tid >> 5] = ssum; }
  __syncthreads();
  if (tid < 2) {
    double t = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) t += (double)red[tid][w];
    atomicAdd(sums + 2 * c + tid, t);
  }
}

// v_render[c,y,x,ch] = v_total[c] * (1 - mask) * ( (1 - lambda) sign(x - gt) / (3 H W)
//                                                 - lambda / (3 Hi Wi) * (win*dmu + 2 x win*dExx + gt win*dExy) )
__global__ void __launch_bounds__(kLossThreads) l1_ssim_bwd_kernel(int H, int W, const float* __restrict__ render,
                                                                   const float* __restrict__ gt,
                                                                   const float* __restrict__ mask,
                                                                   const float* __restrict__ maps, float lambda_ssim,
                                                                   const float* __restrict__ v_total,
                                                                   float* __restrict__ v_render, const Window win) {
  __shared__ float sm[kInRows * kInCols];
  __shared__ float hm[kInRows * kOutCols];
  const int Hi = H - kHalo, Wi = W - kHalo;
  const int c = blockIdx.z;
  const int y0 = blockIdx.y * kTY, x0 = blockIdx.x * kTX;
  const size_t img = (size_t)c * H * W * 3;
  const size_t plane = (size_t)gridDim.z * Hi * Wi * 3;
  const int tid = threadIdx.x;
  constexpr int kPer = kTY * kOutCols / kLossThreads;  // 6 outputs per thread
  float acc[kPer], xs[kPer], ys[kPer];
  bool ok[kPer];
#pragma unroll
  for (int t = 0; t < kPer; ++t) {
    const int i = tid + t * kLossThreads;
    const int y = i / kOutCols, j = i - y * kOutCols;
    const int gy = y0 + y, gx3 = x0 * 3 + j;
    ok[t] = gy < H && gx3 < W * 3;
    acc[t] = 0.f; xs[t] = 0.f; ys[t] = 0.f;
    if (ok[t]) {
      const size_t o = img + (size_t)gy * W * 3 + gx3;
      const float r = __ldg(render + o);
      ys[t] = __ldg(gt + o);
      const float m = mask ? __ldg(mask + ((size_t)c * H + gy) * W + gx3 / 3) : 0.f;
      xs[t] = m * ys[t] + (1.0f - m) * r;
    }
  }
  for (int q = 0; q < 3; ++q) {
    const float* mq = maps + q * plane + (size_t)c * Hi * Wi * 3;
    // map values at interior coordinates [y0-10, y0+kTY) x [x0-10, x0+kTX), zero outside the map
    for (int i = tid; i < kInRows * kInCols; i += kLossThreads) {
      const int row = i / kInCols, col = i - row * kInCols;
      const int py = y0 - kHalo + row, px3 = (x0 - kHalo) * 3 + col;
      float v = 0.f;
      if (py >= 0 && py < Hi && px3 >= 0 && px3 < Wi * 3) v = __ldg(mq + (size_t)py * Wi * 3 + px3);
      sm[i] = v;
    }
    __syncthreads();
    for (int i = tid; i < kInRows * kOutCols; i += kLossThreads) {
      const int row = i / kOutCols, j = i - row * kOutCols;
      const float* a = sm + row * kInCols + j;
      float h = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) h += win.w[k] * a[3 * k];
      hm[i] = h;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kPer; ++t) {
      const int i = tid + t * kLossThreads;
      const float* col = hm + i;  // row y, column j; the window walks down the rows
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) v += win.w[k] * col[k * kOutCols];
      const float coef = q == 0 ? 1.0f : (q == 1 ? 2.0f * xs[t] : ys[t]);
      acc[t] += coef * v;
    }
    __syncthreads();
  }
  const float vt = __ldg(v_total + c);
  const float k_l1 = (1.0f - lambda_ssim) / (3.0f * (float)H * (float)W);
  const float k_ss = lambda_ssim / (3.0f * (float)Hi * (float)Wi);
#pragma unroll
  for (int t = 0; t < kPer; ++t) {
    if (!ok[t]) continue;
    const int i = tid + t * kLossThreads;
    const int y = i / kOutCols, j = i - y * kOutCols;
    const int gy = y0 + y, gx3 = x0 * 3 + j;
    const float m = mask ? __ldg(mask + ((size_t)c * H + gy) * W + gx3 / 3) : 0.f;
    const float d = xs[t] - ys[t];
    const float sgn = d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f);
    v_render[img + (size_t)gy * W * 3 + gx3] = vt * (1.0f - m) * (k_l1 * sgn - k_ss * acc[t]);
  }
}

static Window make_window() {
  // torchmetrics _gaussian(kernel_size = 11, sigma = 1.5): exp(-(d / sigma)^2 / 2), d = -5..5, normalised, in fp32
  Window win;
  float s = 0.f;
  for (int k = 0; k < kWin; ++k) {
    const float d = (float)(k - kWin / 2) / 1.5f;
    win.w[k] = expf(-(d * d) / 2.0f);
    s += win.w[k];
  }
  for (int k = 0; k < kWin; ++k) win.w[k] /= s;
  return win;
}

}  // namespace egs

using namespace egs;

extern "C" int egs_l1_ssim_fwd(int32_t C, int32_t H, int32_t W, const float* render, const float* gt, const float* mask,
                               float* maps, double* sums, egs_stream_t stream) {
  EGS_REQUIRE(C >= 0 && C <= 65535, "l1_ssim_fwd: C=%d out of range", C);
  EGS_REQUIRE(H > kHalo && W > kHalo, "l1_ssim_fwd: images must be larger than the 11x11 SSIM window (got %d x %d)", W, H);
  if (C == 0) return 0;
  constexpr int kSmem = (2 * kInRows * kInCols + 5 * kInRows * kOutCols) * (int)sizeof(float);
  const cudaError_t rc = cudaFuncSetAttribute(l1_ssim_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  if (rc != cudaSuccess) return fail((int)rc, "l1_ssim_fwd: shared memory opt-in failed: %s", cudaGetErrorString(rc));
  dim3 grid((unsigned)ceil_div(W - kHalo, kTX), (unsigned)ceil_div(H - kHalo, kTY), (unsigned)C);
  l1_ssim_fwd_kernel<<<grid, kLossThreads, kSmem, (cudaStream_t)stream>>>(H, W, render, gt, mask, maps, sums, make_window());
  return check_launch("l1_ssim_fwd_kernel");
}

extern "C" int egs_l1_ssim_bwd(int32_t C, int32_t H, int32_t W, const float* render, const float* gt, const float* mask,
                               const float* maps, float lambda_ssim, const float* v_total, float* v_render,
                               egs_stream_t stream) {
  EGS_REQUIRE(C >= 0 && C <= 65535, "l1_ssim_bwd: C=%d out of range", C);
  EGS_REQUIRE(H > kHalo && W > kHalo, "l1_ssim_bwd: images must be larger than the 11x11 SSIM window (got %d x %d)", W, H);
  EGS_REQUIRE(maps != nullptr && v_total != nullptr, "l1_ssim_bwd: maps and v_total are required");
  if (C == 0) return 0;
  dim3 grid((unsigned)ceil_div(W, kTX), (unsigned)ceil_div(H, kTY), (unsigned)C);
  l1_ssim_bwd_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(H, W, render, gt, mask, maps, lambda_ssim, v_total,
                                                                     v_render, make_window());
  return check_launch("l1_ssim_bwd_kernel");
}
