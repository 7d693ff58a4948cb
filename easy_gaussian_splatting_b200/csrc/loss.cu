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
// 64 B/pixel read + 12 B/pixel written backward.  The maps are an internal, channel-planar buffer.
#include <math.h>

#include "egs_common.cuh"

namespace egs {

constexpr int kWin = 11;          // window taps
constexpr int kHalo = kWin - 1;   // 10
constexpr int kTX = 32, kTY = 16; // output tile (pixels)
constexpr int kLossThreads = 256;
constexpr int kInRows = kTY + kHalo;         // 26 staged rows
constexpr int kInW = kTX + kHalo;            // 42 staged columns per channel
constexpr int kPitch = 44;                   // staged row pitch (floats): 16-byte aligned groups of 4 columns
constexpr int kPlane = kInRows * kPitch;     // one staged channel plane
constexpr int kXG = 4, kYG = 4;              // outputs per thread along x (horizontal pass) / y (vertical pass)
constexpr int kHItems = kInRows * 3 * (kTX / kXG);  // 624 horizontal work items per tile
constexpr int kVItems = kTX * 3 * (kTY / kYG);      // 384 vertical work items per tile
constexpr int kVPer = (kVItems + kLossThreads - 1) / kLossThreads;  // 2
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;  // (k1 * data_range)^2, (k2 * data_range)^2, data_range = 1

struct Window { float w[kWin]; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// kXG + 10 consecutive staged values of one row (16-byte aligned start): 3 LDS.128 + 1 LDS.64
__device__ __forceinline__ void load_span(const float* __restrict__ p, float (&v)[kXG + kHalo]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4),
               c = *reinterpret_cast<const float4*>(p + 8);
  const float2 d = *reinterpret_cast<const float2*>(p + 12);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y;
}

// Both kernels are separable 11-tap windows in shared memory with REGISTER sliding windows: a thread produces 4
// neighbouring outputs from 14 inputs it loads once (the first version re-read every input 11 times and was bound
// by shared-memory bandwidth: 0.57 ms for four 1080p images, ncu/bench r1j), on channel-planar tiles so that the
// spans are 128-bit shared loads and the lanes of the vertical pass walk consecutive columns (conflict free).
//
// sums[c] = {sum |x - gt| over H*W*3, sum of the SSIM map over (H-10)*(W-10)*3}; maps (nullable) = 3 planes of
// [C, 3, H-10, W-10] (channel-planar): dS/dmu_x, dS/dE[x^2], dS/dE[xy]
__global__ void __launch_bounds__(kLossThreads, 2) l1_ssim_fwd_kernel(int H, int W, const float* __restrict__ render,
                                                                      const float* __restrict__ gt,
                                                                      const float* __restrict__ mask, float* __restrict__ maps,
                                                                      double* __restrict__ sums, const Window win) {
  extern __shared__ __align__(16) float smem_loss[];
  float* sx = smem_loss;                 // [3][kInRows][kPitch]  x - 0.5
  float* sy = sx + 3 * kPlane;           // [3][kInRows][kPitch]  gt - 0.5
  float* hz = sy + 3 * kPlane;           // [5][3][kInRows][kTX]  row-filtered x, y, x^2, y^2, xy
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

  // stage the (kTY+10) x (kTX+10) window of x and gt, de-interleaving the channels (global reads stay contiguous);
  // every image pixel is counted into the L1 sum by exactly one block
  // A warp stages whole rows (126 contiguous floats = 4 per lane), kRowsPerWarp rows per trip with every global load
  // of the trip in flight before the first use (a load -> use -> load chain per element left the 16 resident warps
  // per SM waiting on HBM latency), and almost no index arithmetic (ncu r1j: a third of the first version's
  // instructions were integer address math).
  float l1 = 0.f;
  {
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kLossThreads / 32, kPerLane = (kInW * 3 + 31) / 32;  // 8, 4
    constexpr int kTrips = (kInRows + kWarps - 1) / kWarps;                     // 4
    float rv[kTrips][kPerLane], yv[kTrips][kPerLane], mv[kTrips][kPerLane];
#pragma unroll
    for (int t = 0; t < kTrips; ++t) {
      const int row = warp + t * kWarps, gy = y0 + row;
      const bool row_ok = row < kInRows && gy < H;
      const float* rrow = rp + (size_t)(row_ok ? gy : 0) * W * 3 + (size_t)x0 * 3;
      const float* grow = gp + (size_t)(row_ok ? gy : 0) * W * 3 + (size_t)x0 * 3;
      const float* mrow = mp ? mp + (size_t)(row_ok ? gy : 0) * W + x0 : nullptr;
#pragma unroll
      for (int j = 0; j < kPerLane; ++j) {
        const int col = lane + 32 * j, px = col / 3;
        const bool ok = row_ok && col < kInW * 3 && x0 + px < W;
        rv[t][j] = ok ? __ldg(rrow + col) : 0.f;
        yv[t][j] = ok ? __ldg(grow + col) : 0.f;
        mv[t][j] = (ok && mrow) ? __ldg(mrow + px) : 0.f;
      }
    }
#pragma unroll
    for (int t = 0; t < kTrips; ++t) {
      const int row = warp + t * kWarps, gy = y0 + row;
      if (row < kInRows) {
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
          const int col = lane + 32 * j, px = col / 3, ch = col - px * 3;
          if (col < kInW * 3) {
            float x = 0.5f, y = 0.5f;
            if (gy < H && x0 + px < W) {
              y = yv[t][j];
              x = mv[t][j] * y + (1.0f - mv[t][j]) * rv[t][j];
              if ((row < kTY || last_y) && (px < kTX || last_x)) l1 += fabsf(x - y);
            }
            // staged values are centred on 0.5: variances / covariances are shift invariant and lose ~4x fewer
            // bits to the E[x^2] - mu^2 cancellation for images in [0, 1]
            sx[ch * kPlane + row * kPitch + px] = x - 0.5f;
            sy[ch * kPlane + row * kPitch + px] = y - 0.5f;
          }
        }
      }
    }
  }
  __syncthreads();
  // horizontal pass: item = (row, channel, group of 4 columns); 5 windowed sums per output
  for (int i = tid; i < kHItems; i += kLossThreads) {
    const int g = i & 7, rc = i >> 3;   // rc = row * 3 + ch
    const int row = rc / 3, ch = rc - row * 3;
    float a[kXG + kHalo], b[kXG + kHalo];
    load_span(sx + ch * kPlane + row * kPitch + g * kXG, a);
    load_span(sy + ch * kPlane + row * kPitch + g * kXG, b);
    float2 s01[kXG], s23[kXG];
    float s4[kXG];
#pragma unroll
    for (int o = 0; o < kXG; ++o) { s01[o] = make_float2(0.f, 0.f); s23[o] = make_float2(0.f, 0.f); s4[o] = 0.f; }
#pragma unroll
    for (int j = 0; j < kXG + kHalo; ++j) {
      const float2 ab = make_float2(a[j], b[j]);
      const float2 sq = __fmul2_rn(ab, ab);
      const float xy = a[j] * b[j];
#pragma unroll
      for (int o = 0; o < kXG; ++o) {
        const int k = j - o;  // tap of output o that input j feeds
        if (k >= 0 && k < kWin) {
          const float2 w2 = make_float2(win.w[k], win.w[k]);
          s01[o] = __ffma2_rn(w2, ab, s01[o]);
          s23[o] = __ffma2_rn(w2, sq, s23[o]);
          s4[o] = fmaf(win.w[k], xy, s4[o]);
        }
      }
    }
    float* dst = hz + (ch * kInRows + row) * kTX + g * kXG;
    constexpr int kQ = 3 * kInRows * kTX;
    *reinterpret_cast<float4*>(dst + 0 * kQ) = make_float4(s01[0].x, s01[1].x, s01[2].x, s01[3].x);
    *reinterpret_cast<float4*>(dst + 1 * kQ) = make_float4(s01[0].y, s01[1].y, s01[2].y, s01[3].y);
    *reinterpret_cast<float4*>(dst + 2 * kQ) = make_float4(s23[0].x, s23[1].x, s23[2].x, s23[3].x);
    *reinterpret_cast<float4*>(dst + 3 * kQ) = make_float4(s23[0].y, s23[1].y, s23[2].y, s23[3].y);
    *reinterpret_cast<float4*>(dst + 4 * kQ) = make_float4(s4[0], s4[1], s4[2], s4[3]);
  }
  __syncthreads();
  // vertical pass + SSIM + derivative maps: item = (group of 4 rows, channel, column)
  float ssum = 0.f;
  for (int i = tid; i < kVItems; i += kLossThreads) {
    const int x = i & 31, t = i >> 5;   // t = yg * 3 + ch
    const int yg = t / 3, ch = t - yg * 3;
    const int px = x0 + x;
    constexpr int kQ = 3 * kInRows * kTX;
    const float* col = hz + (ch * kInRows + yg * kYG) * kTX + x;
    float2 v01[kYG], v23[kYG];
    float v4[kYG];
#pragma unroll
    for (int o = 0; o < kYG; ++o) { v01[o] = make_float2(0.f, 0.f); v23[o] = make_float2(0.f, 0.f); v4[o] = 0.f; }
#pragma unroll
    for (int j = 0; j < kYG + kHalo; ++j) {
      const float2 q01 = make_float2(col[0 * kQ + j * kTX], col[1 * kQ + j * kTX]);
      const float2 q23 = make_float2(col[2 * kQ + j * kTX], col[3 * kQ + j * kTX]);
      const float q4 = col[4 * kQ + j * kTX];
#pragma unroll
      for (int o = 0; o < kYG; ++o) {
        const int k = j - o;
        if (k >= 0 && k < kWin) {
          const float2 w2 = make_float2(win.w[k], win.w[k]);
          v01[o] = __ffma2_rn(w2, q01, v01[o]);
          v23[o] = __ffma2_rn(w2, q23, v23[o]);
          v4[o] = fmaf(win.w[k], q4, v4[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < kYG; ++o) {
      const int py = y0 + yg * kYG + o;
      if (py >= Hi || px >= Wi) continue;
      const float sxx = v23[o].x - v01[o].x * v01[o].x, syy = v23[o].y - v01[o].y * v01[o].y,
                  sxy = v4[o] - v01[o].x * v01[o].y;
      const float mux = v01[o].x + 0.5f, muy = v01[o].y + 0.5f;
      const float A1 = 2.f * mux * muy + kC1, A2 = 2.f * sxy + kC2;
      const float B1 = mux * mux + muy * muy + kC1, B2 = sxx + syy + kC2;
      const float rB1 = 1.0f / B1, rB2 = 1.0f / B2;
      const float S = A1 * A2 * rB1 * rB2;
      ssum += S;
      if (maps) {
        const size_t plane = (size_t)gridDim.z * 3 * Hi * Wi;
        const size_t off = (((size_t)c * 3 + ch) * Hi + py) * Wi + px;
        maps[off] = 2.f * muy * (A2 - A1) * rB1 * rB2 - 2.f * mux * S * (rB1 - rB2);  // dS/dmu_x (through sxx, sxy too)
        maps[plane + off] = -S * rB2;                                                // dS/dE[x^2]
        maps[2 * plane + off] = 2.f * A1 * rB1 * rB2;                                // dS/dE[xy]
      }
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
__global__ void __launch_bounds__(kLossThreads, 2) l1_ssim_bwd_kernel(int H, int W, const float* __restrict__ render,
                                                                      const float* __restrict__ gt,
                                                                      const float* __restrict__ mask,
                                                                      const float* __restrict__ maps, float lambda_ssim,
                                                                      const float* __restrict__ v_total,
                                                                      float* __restrict__ v_render, const Window win) {
  __shared__ __align__(16) float sm[3 * kPlane];           // one map's halo, channel planar
  __shared__ __align__(16) float hm[3 * kInRows * kTX];    // row-filtered
  __shared__ float sxv[kTY * kTX * 3], syv[kTY * kTX * 3]; // x and gt of the tile's pixels (interleaved, as in HBM)
  __shared__ float outv[kTY * kTX * 3];                    // windowed sums, interleaved for a contiguous store
  const int Hi = H - kHalo, Wi = W - kHalo;
  const int c = blockIdx.z;
  const int y0 = blockIdx.y * kTY, x0 = blockIdx.x * kTX;
  const size_t img = (size_t)c * H * W * 3;
  const size_t plane = (size_t)gridDim.z * 3 * Hi * Wi;
  const int tid = threadIdx.x;
  // the tile's own pixels: contiguous rows of 96 floats (loads batched, see the forward kernel)
  {
    constexpr int kOwn = kTY * kTX * 3 / kLossThreads;  // 6
    float rv[kOwn], gv[kOwn], mv[kOwn];
#pragma unroll
    for (int it = 0; it < kOwn; ++it) {
      const int i = tid + it * kLossThreads;
      const int y = i / (kTX * 3), j = i - y * (kTX * 3);
      const int gy = y0 + y, gx3 = x0 * 3 + j;
      const bool ok = gy < H && gx3 < W * 3;
      const size_t o = ok ? img + (size_t)gy * W * 3 + gx3 : 0;
      rv[it] = ok ? __ldg(render + o) : 0.f;
      gv[it] = ok ? __ldg(gt + o) : 0.f;
      mv[it] = (ok && mask) ? __ldg(mask + ((size_t)c * H + gy) * W + gx3 / 3) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < kOwn; ++it) {
      const int i = tid + it * kLossThreads;
      sxv[i] = mv[it] * gv[it] + (1.0f - mv[it]) * rv[it];
      syv[i] = gv[it];
    }
  }
  float acc[kVPer][kYG];
#pragma unroll
  for (int s = 0; s < kVPer; ++s)
#pragma unroll
    for (int o = 0; o < kYG; ++o) acc[s][o] = 0.f;
  for (int q = 0; q < 3; ++q) {
    __syncthreads();  // previous map's tiles are free (and sxv / syv are complete)
    // map values at interior coordinates [y0-10, y0+kTY) x [x0-10, x0+kTX), zero outside the map
    {
      // a warp stages whole (channel, row) lines of the map halo: 42 contiguous floats, 78 lines = 10 trips of 8 warps
      const int lane = tid & 31, warp = tid >> 5;
      constexpr int kWarps = kLossThreads / 32, kLines = 3 * kInRows, kTrips = (kLines + kWarps - 1) / kWarps;
      const float* mq = maps + q * plane + (size_t)c * 3 * Hi * Wi;
      float mvv[kTrips][2];
#pragma unroll
      for (int t = 0; t < kTrips; ++t) {
        const int line = warp + t * kWarps;              // = ch * kInRows + row
        const int ch = line / kInRows, row = line - ch * kInRows;
        const int my = y0 - kHalo + row;
        const bool line_ok = line < kLines && my >= 0 && my < Hi;
        const float* src = mq + ((size_t)ch * Hi + (line_ok ? my : 0)) * Wi;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int px = lane + 32 * j, mx = x0 - kHalo + px;
          mvv[t][j] = (line_ok && px < kInW && mx >= 0 && mx < Wi) ? __ldg(src + mx) : 0.f;
        }
      }
#pragma unroll
      for (int t = 0; t < kTrips; ++t) {
        const int line = warp + t * kWarps;
        if (line < kLines) {
          const int ch = line / kInRows, row = line - ch * kInRows;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int px = lane + 32 * j;
            if (px < kInW) sm[ch * kPlane + row * kPitch + px] = mvv[t][j];
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < kHItems; i += kLossThreads) {
      const int g = i & 7, rc = i >> 3;
      const int row = rc / 3, ch = rc - row * 3;
      float a[kXG + kHalo];
      load_span(sm + ch * kPlane + row * kPitch + g * kXG, a);
      float h[kXG] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kXG + kHalo; ++j)
#pragma unroll
        for (int o = 0; o < kXG; ++o) {
          const int k = j - o;
          if (k >= 0 && k < kWin) h[o] = fmaf(win.w[k], a[j], h[o]);
        }
      *reinterpret_cast<float4*>(hm + (ch * kInRows + row) * kTX + g * kXG) = make_float4(h[0], h[1], h[2], h[3]);
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kVPer; ++s) {
      const int i = tid + s * kLossThreads;
      if (i < kVItems) {
        const int x = i & 31, t = i >> 5;
        const int yg = t / 3, ch = t - yg * 3;
        const float* col = hm + (ch * kInRows + yg * kYG) * kTX + x;
        float v[kYG] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < kYG + kHalo; ++j) {
          const float hv = col[j * kTX];
#pragma unroll
          for (int o = 0; o < kYG; ++o) {
            const int k = j - o;
            if (k >= 0 && k < kWin) v[o] = fmaf(win.w[k], hv, v[o]);
          }
        }
#pragma unroll
        for (int o = 0; o < kYG; ++o) {
          const int e = ((yg * kYG + o) * kTX + x) * 3 + ch;
          const float coef = q == 0 ? 1.0f : (q == 1 ? 2.0f * sxv[e] : syv[e]);
          acc[s][o] = fmaf(coef, v[o], acc[s][o]);
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < kVPer; ++s) {
    const int i = tid + s * kLossThreads;
    if (i < kVItems) {
      const int x = i & 31, t = i >> 5;
      const int yg = t / 3, ch = t - yg * 3;
#pragma unroll
      for (int o = 0; o < kYG; ++o) outv[((yg * kYG + o) * kTX + x) * 3 + ch] = acc[s][o];
    }
  }
  __syncthreads();
  const float vt = __ldg(v_total + c);
  const float k_l1 = (1.0f - lambda_ssim) / (3.0f * (float)H * (float)W);
  const float k_ss = lambda_ssim / (3.0f * (float)Hi * (float)Wi);
  for (int i = tid; i < kTY * kTX * 3; i += kLossThreads) {
    const int y = i / (kTX * 3), j = i - y * (kTX * 3);
    const int gy = y0 + y, gx3 = x0 * 3 + j;
    if (gy < H && gx3 < W * 3) {
      const float m = mask ? __ldg(mask + ((size_t)c * H + gy) * W + gx3 / 3) : 0.f;
      const float d = sxv[i] - syv[i];
      const float sgn = d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f);
      v_render[img + (size_t)gy * W * 3 + gx3] = vt * (1.0f - m) * (k_l1 * sgn - k_ss * outv[i]);
    }
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
  constexpr int kSmem = (6 * kPlane + 5 * 3 * kInRows * kTX) * (int)sizeof(float);
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
