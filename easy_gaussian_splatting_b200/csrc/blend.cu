// g6 / g7: alpha blending forward and backward (takes the place of gsplat's
// rasterize_to_pixels fwd/bwd).  FP32 SIMT + MUFU bound; no dense contraction, so
// no tensor cores.
//
// One CTA per 16x16 pixel tile, one pixel per thread, each warp owning a compact 8x4 pixel block
// (more coherent accept / terminate decisions than two 16-pixel rows).  The tile's depth-sorted
// Gaussians are staged through shared memory in batches of 256 packed 48-byte splat records with
// cp.async (LDGSTS, three 16-byte copies per record, double buffered, with the flatten ids of the
// batch after next prefetched into registers), so the gather latency of batch b+1 hides behind the
// blending of batch b.
//
// Backward: per-pixel back-to-front replay; the 11 per-Gaussian partial gradients of a warp are
// combined with a 16-slot shuffle reduce-scatter (16 SHFL instead of the 55 of a per-value
// butterfly), after which 11 lanes issue one coalesced RED.ADD.F32 into the packed 48-byte gradient
// record of the Gaussian.
#include "egs_common.cuh"

namespace egs {

constexpr int kTileSize = 16;
constexpr int kBlendThreads = kTileSize * kTileSize;  // 256
constexpr int kBatch = kBlendThreads;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTMin = 1e-4f;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct TileCoord {
  int cam, tile_id, px, py;
  bool inside;
  int range_start, range_end;
};

__device__ __forceinline__ TileCoord tile_setup(int width, int height, int tile_w, int tile_h, int64_t n_isects,
                                                const int32_t* __restrict__ tile_offsets, int n_tiles_total) {
  TileCoord tc;
  tc.cam = blockIdx.z;
  tc.tile_id = (tc.cam * tile_h + blockIdx.y) * tile_w + blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  tc.px = blockIdx.x * kTileSize + (warp & 1) * 8 + (lane & 7);
  tc.py = blockIdx.y * kTileSize + (warp >> 1) * 4 + (lane >> 3);
  tc.inside = tc.px < width && tc.py < height;
  tc.range_start = tile_offsets[tc.tile_id];
  tc.range_end = (tc.tile_id == n_tiles_total - 1) ? (int)n_isects : tile_offsets[tc.tile_id + 1];
  return tc;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kBlendThreads) rasterize_fwd_kernel(
    int64_t n_isects, const float4* __restrict__ splats, const int32_t* __restrict__ tile_offsets,
    const int32_t* __restrict__ flatten_ids, const float* __restrict__ backgrounds, int width, int height, int tile_w,
    int tile_h, int n_tiles_total, float* __restrict__ render_colors, float* __restrict__ render_alphas,
    int32_t* __restrict__ last_ids, unsigned long long* __restrict__ pair_counters) {
  __shared__ __align__(16) float4 sb[2][kBatch * 3];
  const int tid = threadIdx.x;
  const TileCoord tc = tile_setup(width, height, tile_w, tile_h, n_isects, tile_offsets, n_tiles_total);
  const float px = (float)tc.px + 0.5f, py = (float)tc.py + 0.5f;
  const int range = tc.range_end - tc.range_start;
  const int nb = (range + kBatch - 1) / kBatch;

  float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f;
  int last = 0;
  bool done = !tc.inside;
  unsigned int n_eval = 0, n_acc = 0;

  auto load_id = [&](int b) -> int {
    const int idx = tc.range_start + b * kBatch + tid;
    return (b < nb && idx < tc.range_end) ? __ldg(flatten_ids + idx) : -1;
  };
  auto issue = [&](int buf, int id) {
    if (id >= 0) {
      const float4* src = splats + (size_t)id * 3;
      float4* dst = &sb[buf][tid * 3];
      cp_async16(dst + 0, src + 0);
      cp_async16(dst + 1, src + 1);
      cp_async16(dst + 2, src + 2);
    }
    cp_async_commit();
  };

  if (nb > 0) {
    issue(0, load_id(0));
    int id_next = load_id(1);
    for (int b = 0; b < nb; ++b) {
      if (b + 1 < nb) {
        issue((b + 1) & 1, id_next);
        id_next = load_id(b + 2);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      // barrier (makes batch b visible to everyone) + vote: stop when every pixel is finished
      if (__syncthreads_and(done)) break;
      const int batch_start = tc.range_start + b * kBatch;
      const int batch_size = min(kBatch, tc.range_end - batch_start);
      const float4* s = sb[b & 1];
      for (int t = 0; t < batch_size && !done; ++t) {
        const float4 g0 = s[t * 3 + 0];  // x, y, conic_a, conic_b
        const float4 g1 = s[t * 3 + 1];  // conic_c, opacity, r, g
        const float dx = g0.x - px, dy = g0.y - py;
        const float sigma = 0.5f * (g0.z * dx * dx + g1.x * dy * dy) + g0.w * dx * dy;
        const float alpha = fminf(kAlphaMax, g1.y * __expf(-sigma));
        if (COUNT) ++n_eval;
        if (sigma < 0.f || alpha < kAlphaMin) continue;
        const float next_T = T * (1.0f - alpha);
        if (next_T <= kTMin) { done = true; break; }
        const float w = alpha * T;
        cr += g1.z * w;
        cg += g1.w * w;
        cb += s[t * 3 + 2].x * w;
        last = batch_start + t;
        T = next_T;
        if (COUNT) ++n_acc;
      }
      __syncthreads();  // everyone is done with buffer b&1 before batch b+2 overwrites it
    }
    cp_async_wait<0>();
  }

  if (tc.inside) {
    const size_t pix = ((size_t)tc.cam * height + tc.py) * width + tc.px;
    if (backgrounds != nullptr) {
      const float* bg = backgrounds + tc.cam * 3;
      cr += T * bg[0]; cg += T * bg[1]; cb += T * bg[2];
    }
    render_colors[pix * 3 + 0] = cr;
    render_colors[pix * 3 + 1] = cg;
    render_colors[pix * 3 + 2] = cb;
    render_alphas[pix] = 1.0f - T;
    last_ids[pix] = last;
  }
  if (COUNT) {
    // warp-reduce, then one atomic pair per warp
    for (int d = 16; d > 0; d >>= 1) {
      n_eval += __shfl_xor_sync(0xffffffffu, n_eval, d);
      n_acc += __shfl_xor_sync(0xffffffffu, n_acc, d);
    }
    if ((tid & 31) == 0) {
      atomicAdd(pair_counters + 0, (unsigned long long)n_eval);
      atomicAdd(pair_counters + 1, (unsigned long long)n_acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------

// 16-slot reduce-scatter across the warp: on return every lane holds the warp total of slot
// (lane >> 1).  16 shuffles.
__device__ __forceinline__ float warp_reduce_scatter16(float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = b4 ? v[j] : v[j + 8];
    const float keep = b4 ? v[j + 8] : v[j];
    a8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = b3 ? a8[j] : a8[j + 4];
    const float keep = b3 ? a8[j + 4] : a8[j];
    a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = b2 ? a4[j] : a4[j + 2];
    const float keep = b2 ? a4[j + 2] : a4[j];
    a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? a2[0] : a2[1];
  const float keep = b1 ? a2[1] : a2[0];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

__global__ void __launch_bounds__(kBlendThreads) rasterize_bwd_kernel(
    int64_t n_isects, const float4* __restrict__ splats, const int32_t* __restrict__ tile_offsets,
    const int32_t* __restrict__ flatten_ids, const float* __restrict__ backgrounds, int width, int height, int tile_w,
    int tile_h, int n_tiles_total, const float* __restrict__ render_alphas, const int32_t* __restrict__ last_ids,
    const float* __restrict__ v_render_colors, const float* __restrict__ v_render_alphas,
    float* __restrict__ v_splats) {
  __shared__ __align__(16) float4 sb[2][kBatch * 3];
  __shared__ int s_id[2][kBatch];
  __shared__ int s_warp_last[kBlendThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const TileCoord tc = tile_setup(width, height, tile_w, tile_h, n_isects, tile_offsets, n_tiles_total);
  if (tc.range_end <= tc.range_start) return;  // uniform for the block
  const float px = (float)tc.px + 0.5f, py = (float)tc.py + 0.5f;

  float T_final = 1.f, vcr = 0.f, vcg = 0.f, vcb = 0.f, va = 0.f;
  int bin_final = -1;  // pixels outside the image never match any index
  if (tc.inside) {
    const size_t pix = ((size_t)tc.cam * height + tc.py) * width + tc.px;
    T_final = 1.0f - render_alphas[pix];
    bin_final = last_ids[pix];
    vcr = v_render_colors[pix * 3 + 0];
    vcg = v_render_colors[pix * 3 + 1];
    vcb = v_render_colors[pix * 3 + 2];
    va = v_render_alphas[pix];
  }
  float bg_dot = 0.f;
  if (backgrounds != nullptr) {
    const float* bg = backgrounds + tc.cam * 3;
    bg_dot = bg[0] * vcr + bg[1] * vcg + bg[2] * vcb;
  }
  float T = T_final, br = 0.f, bgc = 0.f, bb = 0.f;  // running T and colour accumulated behind

  const int warp_last = __reduce_max_sync(0xffffffffu, bin_final);
  if (lane == 0) s_warp_last[warp] = warp_last;
  __syncthreads();
  int block_last = s_warp_last[0];
#pragma unroll
  for (int w = 1; w < kBlendThreads / 32; ++w) block_last = max(block_last, s_warp_last[w]);
  // nothing behind the last blended Gaussian of any pixel of the tile can receive gradient
  const int end_idx = min(tc.range_end - 1, block_last);
  if (end_idx < tc.range_start) return;
  const int nb = (end_idx - tc.range_start + 1 + kBatch - 1) / kBatch;

  auto load_id = [&](int b) -> int {
    const int idx = end_idx - b * kBatch - tid;
    return (b < nb && idx >= tc.range_start) ? __ldg(flatten_ids + idx) : -1;
  };
  auto issue = [&](int buf, int id) {
    if (id >= 0) {
      const float4* src = splats + (size_t)id * 3;
      float4* dst = &sb[buf][tid * 3];
      cp_async16(dst + 0, src + 0);
      cp_async16(dst + 1, src + 1);
      cp_async16(dst + 2, src + 2);
    }
    s_id[buf][tid] = id;
    cp_async_commit();
  };

  issue(0, load_id(0));
  int id_next = load_id(1);
  for (int b = 0; b < nb; ++b) {
    if (b + 1 < nb) {
      issue((b + 1) & 1, id_next);
      id_next = load_id(b + 2);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int batch_end = end_idx - b * kBatch;  // sorted index held in slot 0 (the one furthest back)
    const int batch_size = min(kBatch, batch_end + 1 - tc.range_start);
    const float4* s = sb[b & 1];
    const int* ids = s_id[b & 1];
    for (int t = max(0, batch_end - warp_last); t < batch_size; ++t) {
      bool valid = (batch_end - t) <= bin_final;
      float4 g0, g1;
      float dx = 0.f, dy = 0.f, vis = 0.f, alpha = 0.f;
      if (valid) {
        g0 = s[t * 3 + 0];
        g1 = s[t * 3 + 1];
        dx = g0.x - px; dy = g0.y - py;
        const float sigma = 0.5f * (g0.z * dx * dx + g1.x * dy * dy) + g0.w * dx * dy;
        vis = __expf(-sigma);
        alpha = fminf(kAlphaMax, g1.y * vis);
        if (sigma < 0.f || alpha < kAlphaMin) valid = false;
      }
      if (!__any_sync(0xffffffffu, valid)) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
      if (valid) {
        const float cb_ = s[t * 3 + 2].x;
        const float ra = __fdividef(1.0f, 1.0f - alpha);
        T *= ra;
        const float fac = alpha * T;
        v[6] = fac * vcr; v[7] = fac * vcg; v[8] = fac * vcb;
        float v_alpha = (g1.z * T - br * ra) * vcr + (g1.w * T - bgc * ra) * vcg + (cb_ * T - bb * ra) * vcb;
        v_alpha += T_final * ra * va;
        v_alpha -= T_final * ra * bg_dot;
        const float ov = g1.y * vis;
        if (ov <= kAlphaMax) {
          const float v_sigma = -ov * v_alpha;
          v[2] = 0.5f * v_sigma * dx * dx;
          v[3] = v_sigma * dx * dy;
          v[4] = 0.5f * v_sigma * dy * dy;
          v[0] = v_sigma * (g0.z * dx + g0.w * dy);
          v[1] = v_sigma * (g0.w * dx + g1.x * dy);
          v[9] = fabsf(v[0]);
          v[10] = fabsf(v[1]);
          v[5] = vis * v_alpha;
        }
        br += g1.z * fac; bgc += g1.w * fac; bb += cb_ * fac;
      }
      const float total = warp_reduce_scatter16(v, lane);
      const int slot = lane >> 1;
      if ((lane & 1) == 0 && slot < 11) atomicAdd(v_splats + (size_t)ids[t] * EGS_SPLAT_FLOATS + slot, total);
    }
    __syncthreads();
  }
}

}  // namespace egs

using namespace egs;

static int check_raster_args(const char* who, int32_t C, int64_t n_isects, int32_t width, int32_t height,
                             int32_t tile_width, int32_t tile_height) {
  EGS_REQUIRE(C >= 0 && C <= 65535, "%s: C=%d out of [0,65535]", who, C);
  EGS_REQUIRE(width >= 1 && height >= 1, "%s: width/height must be >= 1", who);
  EGS_REQUIRE(tile_width == (width + kTileSize - 1) / kTileSize && tile_height == (height + kTileSize - 1) / kTileSize,
              "%s: tile grid %dx%d does not match %dx%d pixels at tile_size 16", who, tile_width, tile_height, width, height);
  EGS_REQUIRE(tile_height <= 65535, "%s: tile_height=%d exceeds the grid limit", who, tile_height);
  EGS_REQUIRE(n_isects >= 0 && n_isects < 0x7fffffffLL, "%s: n_isects=%lld out of int32 range", who, (long long)n_isects);
  EGS_REQUIRE((int64_t)C * tile_width * tile_height < 0x7fffffffLL, "%s: too many tiles", who);
  return 0;
}

extern "C" int egs_rasterize_fwd(int32_t C, int32_t N, int64_t n_isects, const float* splats,
                                 const int32_t* tile_offsets, const int32_t* flatten_ids, const float* backgrounds,
                                 int32_t width, int32_t height, int32_t tile_width, int32_t tile_height,
                                 float* render_colors, float* render_alphas, int32_t* last_ids, egs_stream_t stream) {
  (void)N;
  if (int rc = check_raster_args("rasterize_fwd", C, n_isects, width, height, tile_width, tile_height)) return rc;
  if (C == 0) return 0;
  dim3 grid(tile_width, tile_height, C);
  rasterize_fwd_kernel<false><<<grid, kBlendThreads, 0, (cudaStream_t)stream>>>(
      n_isects, reinterpret_cast<const float4*>(splats), tile_offsets, flatten_ids, backgrounds, width, height,
      tile_width, tile_height, C * tile_width * tile_height, render_colors, render_alphas, last_ids, nullptr);
  return check_launch("rasterize_fwd_kernel");
}

// Instrumented variant for the roofline model: also accumulates P_eval and P_acc
// into pair_counters[2] (device, uint64, caller-zeroed).  Not used on the product path.
extern "C" int egs_rasterize_fwd_count(int32_t C, int32_t N, int64_t n_isects, const float* splats,
                                       const int32_t* tile_offsets, const int32_t* flatten_ids,
                                       const float* backgrounds, int32_t width, int32_t height, int32_t tile_width,
                                       int32_t tile_height, float* render_colors, float* render_alphas,
                                       int32_t* last_ids, uint64_t* pair_counters, egs_stream_t stream) {
  (void)N;
  if (int rc = check_raster_args("rasterize_fwd_count", C, n_isects, width, height, tile_width, tile_height)) return rc;
  if (C == 0) return 0;
  dim3 grid(tile_width, tile_height, C);
  rasterize_fwd_kernel<true><<<grid, kBlendThreads, 0, (cudaStream_t)stream>>>(
      n_isects, reinterpret_cast<const float4*>(splats), tile_offsets, flatten_ids, backgrounds, width, height,
      tile_width, tile_height, C * tile_width * tile_height, render_colors, render_alphas, last_ids,
      reinterpret_cast<unsigned long long*>(pair_counters));
  return check_launch("rasterize_fwd_kernel<count>");
}

extern "C" int egs_rasterize_bwd(int32_t C, int32_t N, int64_t n_isects, const float* splats,
                                 const int32_t* tile_offsets, const int32_t* flatten_ids, const float* backgrounds,
                                 int32_t width, int32_t height, int32_t tile_width, int32_t tile_height,
                                 const float* render_alphas, const int32_t* last_ids, const float* v_render_colors,
                                 const float* v_render_alphas, float* v_splats, egs_stream_t stream) {
  (void)N;
  if (int rc = check_raster_args("rasterize_bwd", C, n_isects, width, height, tile_width, tile_height)) return rc;
  if (C == 0 || n_isects == 0) return 0;
  dim3 grid(tile_width, tile_height, C);
  rasterize_bwd_kernel<<<grid, kBlendThreads, 0, (cudaStream_t)stream>>>(
      n_isects, reinterpret_cast<const float4*>(splats), tile_offsets, flatten_ids, backgrounds, width, height,
      tile_width, tile_height, C * tile_width * tile_height, render_alphas, last_ids, v_render_colors,
      v_render_alphas, v_splats);
  return check_launch("rasterize_bwd_kernel");
}
