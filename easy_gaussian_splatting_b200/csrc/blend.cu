// g6 / g7: alpha blending forward and backward (takes the place of gsplat's
// rasterize_to_pixels fwd/bwd).  FP32 SIMT + MUFU bound; no dense contraction, so
// no tensor cores.
//
// One CTA per 16x16 pixel tile.  A thread owns a PXW x PYH block of pixels (default 2x2: 64 threads,
// two warps, each warp a compact 16x8 pixel region): the per-Gaussian shared-memory reads, the loop
// overhead, the separable parts of the quadratic form and — in the backward pass — the warp
// reduction are amortised over 4 pixels, and the 4 independent pixel chains give the scheduler ILP.
// The tile's depth-sorted Gaussians are staged through shared memory in batches of 128/256 packed
// 48-byte splat records with cp.async (LDGSTS, three 16-byte copies per record, double buffered,
// flatten ids of the batch after next prefetched into registers), so the gather latency of batch
// b+1 hides behind the blending of batch b.
//
// All loops over a batch are WARP-UNIFORM (finished pixels are predicated off, a vote at the top of
// the body is the reconvergence point).  A per-lane break/continue lets the lanes of a warp drift
// apart under independent thread scheduling: measured with ncu on the first version of the forward
// kernel, 1.9 active threads per instruction and a 20x slowdown (profiles/r1a_*).
//
// Backward: per-pixel back-to-front replay; the 11 per-Gaussian partial gradients are first summed
// over the thread's own pixels in registers, then combined across the warp with a 16-slot shuffle
// reduce-scatter (16 SHFL instead of the 55 of a per-value butterfly), after which 11 lanes issue one
// coalesced RED.ADD.F32 into the packed 48-byte gradient record of the Gaussian.
#include <stdlib.h>

#include "egs_common.cuh"

namespace egs {

constexpr int kTileSize = 16;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTMin = 1e-4f;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// exp(-sigma) as one FMUL + MUFU.EX2 (flush-to-zero: anything that small is rejected as alpha < 1/255)
__device__ __forceinline__ float fast_exp_neg(float sigma) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(sigma * -1.4426950408889634f));
  return e;
}

// ex2 of an argument that is already scaled by -log2(e)
__device__ __forceinline__ float fast_ex2(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}
// Stops ptxas from re-deriving a loop-invariant value inside the loop (it otherwise rematerialises
// the pixel-centre coordinates with I2FP + FADD in every iteration to save two registers).
__device__ __forceinline__ float opaque(float x) {
  asm volatile("" : "+f"(x));
  return x;
}
constexpr float kNegLog2e = -1.4426950408889634f;

// Thread <-> pixel mapping of one tile.
template <int PXW, int PYH>
struct TileMap {
  static constexpr int NP = PXW * PYH;            // pixels per thread
  static constexpr int TX = kTileSize / PXW;      // thread grid
  static constexpr int TY = kTileSize / PYH;
  static constexpr int NT = TX * TY;              // threads per CTA
  static constexpr int NW = NT / 32;              // warps per CTA
  static constexpr int WPR = TX / 8;              // a warp spans 8 x 4 threads
  // splat records staged per batch: small CTAs use small batches so that shared memory (2 buffers x
  // 48 B x BATCH) does not cap the number of resident CTAs below what registers allow
  static constexpr int BATCH = NT >= 256 ? 256 : 128;
  static constexpr int RPT = BATCH / NT;          // records staged per thread per batch
  static_assert(TX % 8 == 0 && TY % 4 == 0, "unsupported pixel block");
};

struct TileRange {
  int cam, tile_id, x0, y0;  // x0,y0: first pixel of this thread's block
  int range_start, range_end;
  float wx_lo, wx_hi, wy_lo, wy_hi;  // pixel-centre rectangle covered by this thread's WARP
};

template <class M>
__device__ __forceinline__ TileRange tile_setup(int tile_w, int tile_h, int64_t n_isects,
                                                const int32_t* __restrict__ tile_offsets, int n_tiles_total) {
  TileRange tc;
  tc.cam = blockIdx.z;
  tc.tile_id = (tc.cam * tile_h + blockIdx.y) * tile_w + blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qx = (warp % M::WPR) * 8 + (lane & 7);
  const int qy = (warp / M::WPR) * 4 + (lane >> 3);
  tc.x0 = blockIdx.x * kTileSize + qx * (kTileSize / M::TX);
  tc.y0 = blockIdx.y * kTileSize + qy * (kTileSize / M::TY);
  tc.range_start = tile_offsets[tc.tile_id];
  tc.range_end = (tc.tile_id == n_tiles_total - 1) ? (int)n_isects : tile_offsets[tc.tile_id + 1];
  constexpr int PXW = kTileSize / M::TX, PYH = kTileSize / M::TY;
  tc.wx_lo = (float)(blockIdx.x * kTileSize + (warp % M::WPR) * 8 * PXW) + 0.5f;
  tc.wx_hi = tc.wx_lo + (float)(8 * PXW - 1);
  tc.wy_lo = (float)(blockIdx.y * kTileSize + (warp / M::WPR) * 4 * PYH) + 0.5f;
  tc.wy_hi = tc.wy_lo + (float)(4 * PYH - 1);
  return tc;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int PXW, int PYH, bool COUNT>
__global__ void __launch_bounds__(TileMap<PXW, PYH>::NT) rasterize_fwd_kernel(
    int64_t n_isects, const float4* __restrict__ splats, const int32_t* __restrict__ tile_offsets,
    const int32_t* __restrict__ flatten_ids, const float* __restrict__ backgrounds, int width, int height, int tile_w,
    int tile_h, int n_tiles_total, float* __restrict__ render_colors, float* __restrict__ render_alphas,
    int32_t* __restrict__ last_ids, unsigned long long* __restrict__ pair_counters) {
  using M = TileMap<PXW, PYH>;
  constexpr int NP = M::NP, NT = M::NT, RPT = M::RPT, kBatch = M::BATCH;
  __shared__ __align__(16) float4 sb[2][kBatch * 3];
  const int tid = threadIdx.x;
  const TileRange tc = tile_setup<M>(tile_w, tile_h, n_isects, tile_offsets, n_tiles_total);
  const int nb = (tc.range_end - tc.range_start + kBatch - 1) / kBatch;

  float pxf[PXW], pyf[PYH];
#pragma unroll
  for (int i = 0; i < PXW; ++i) pxf[i] = opaque((float)(tc.x0 + i) + 0.5f);
#pragma unroll
  for (int i = 0; i < PYH; ++i) pyf[i] = opaque((float)(tc.y0 + i) + 0.5f);
  // T[j] > 0: transmittance of a live pixel; T[j] < 0: pixel finished, |T[j]| is its final transmittance
  // (the "done" flag lives in the sign bit, so liveness is one more FSETP in the accept test).
  const float wx_lo = opaque(tc.wx_lo), wx_hi = opaque(tc.wx_hi), wy_lo = opaque(tc.wy_lo), wy_hi = opaque(tc.wy_hi);
  float T[NP], cr[NP], cg[NP], cb[NP];
  int last[NP];
  bool all_done = true;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const bool inside = (tc.x0 + j % PXW) < width && (tc.y0 + j / PXW) < height;
    T[j] = inside ? 1.0f : -1.0f; cr[j] = 0.f; cg[j] = 0.f; cb[j] = 0.f; last[j] = 0;
    all_done = all_done && !inside;
  }
  unsigned int n_eval = 0, n_acc = 0;

  auto load_ids = [&](int b, int (&ids)[RPT]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int idx = tc.range_start + b * kBatch + r * NT + tid;
      ids[r] = (b < nb && idx < tc.range_end) ? __ldg(flatten_ids + idx) : -1;
    }
  };
  auto issue = [&](int buf, const int (&ids)[RPT]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (ids[r] >= 0) {
        const float4* src = splats + (size_t)ids[r] * 3;
        float4* dst = &sb[buf][(r * NT + tid) * 3];
        cp_async16(dst + 0, src + 0);
        cp_async16(dst + 1, src + 1);
        cp_async16(dst + 2, src + 2);
      }
    }
    cp_async_commit();
  };

  if (nb > 0) {
    int ids[RPT];
    load_ids(0, ids);
    issue(0, ids);
    load_ids(1, ids);
    for (int b = 0; b < nb; ++b) {
      if (b + 1 < nb) {
        issue((b + 1) & 1, ids);
        load_ids(b + 2, ids);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      // barrier (makes batch b visible to everyone) + vote: stop when every pixel of the tile is finished
      if (__syncthreads_and(all_done)) break;
      const int batch_start = tc.range_start + b * kBatch;
      const int batch_size = min(kBatch, tc.range_end - batch_start);
      const float4* s = sb[b & 1];
      for (int t = 0; t < batch_size; ++t) {
        if (__all_sync(0xffffffffu, all_done)) break;  // warp-uniform exit + reconvergence point
        const float4 g0 = s[t * 3 + 0];  // x, y, conic_a, conic_b
        const float4 g2 = s[t * 3 + 2];  // b, depth, r_eff^2, -
        {
          // warp-uniform skip: the warp's pixel rectangle lies outside the circle in which alpha >= 1/255
          const float ex = fmaxf(fmaxf(wx_lo - g0.x, g0.x - wx_hi), 0.f);
          const float ey = fmaxf(fmaxf(wy_lo - g0.y, g0.y - wy_hi), 0.f);
          if (fmaf(ex, ex, ey * ey) > g2.z) {
            if (COUNT) {
#pragma unroll
              for (int j = 0; j < NP; ++j) n_eval += (T[j] > 0.f) ? 1u : 0u;
            }
            continue;
          }
        }
        const float4 g1 = s[t * 3 + 1];  // conic_c, opacity, r, g
        const float cbl = g2.x;
        // q(dx,dy) = -log2(e) * sigma, separable parts shared by the rows / columns of the pixel block
        const float la = (0.5f * kNegLog2e) * g0.z, lc = (0.5f * kNegLog2e) * g1.x, lb = kNegLog2e * g0.w;
        float qx[PXW], bx[PXW], dy[PYH], qy[PYH];
#pragma unroll
        for (int i = 0; i < PXW; ++i) { const float dx = g0.x - pxf[i]; qx[i] = la * dx * dx; bx[i] = lb * dx; }
#pragma unroll
        for (int i = 0; i < PYH; ++i) { dy[i] = g0.y - pyf[i]; qy[i] = lc * dy[i] * dy[i]; }
        float tmax = -1.0f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float q = fmaf(bx[j % PXW], dy[j / PXW], qx[j % PXW] + qy[j / PXW]);
          const float alpha = fminf(kAlphaMax, g1.y * fast_ex2(q));
          if (COUNT && T[j] > 0.f) ++n_eval;
          if (T[j] > 0.f && q <= 0.f && alpha >= kAlphaMin) {  // sigma >= 0  <=>  q <= 0
            const float next_T = T[j] * (1.0f - alpha);
            if (next_T <= kTMin) {
              T[j] = -T[j];  // finished: this Gaussian is not blended
            } else {
              const float w = alpha * T[j];
              cr[j] = fmaf(g1.z, w, cr[j]);
              cg[j] = fmaf(g1.w, w, cg[j]);
              cb[j] = fmaf(cbl, w, cb[j]);
              last[j] = batch_start + t;
              T[j] = next_T;
              if (COUNT) ++n_acc;
            }
          }
          tmax = fmaxf(tmax, T[j]);
        }
        all_done = !(tmax > 0.f);
      }
      __syncthreads();  // everyone is done with buffer b&1 before batch b+2 overwrites it
    }
    cp_async_wait<0>();
  }

  float bgr = 0.f, bgg = 0.f, bgb = 0.f;
  if (backgrounds != nullptr) {
    bgr = backgrounds[tc.cam * 3 + 0]; bgg = backgrounds[tc.cam * 3 + 1]; bgb = backgrounds[tc.cam * 3 + 2];
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int x = tc.x0 + j % PXW, y = tc.y0 + j / PXW;
    if (x < width && y < height) {
      const size_t pix = ((size_t)tc.cam * height + y) * width + x;
      const float Tf = fabsf(T[j]);
      render_colors[pix * 3 + 0] = fmaf(Tf, bgr, cr[j]);
      render_colors[pix * 3 + 1] = fmaf(Tf, bgg, cg[j]);
      render_colors[pix * 3 + 2] = fmaf(Tf, bgb, cb[j]);
      render_alphas[pix] = 1.0f - Tf;
      last_ids[pix] = last[j];
    }
  }
  if (COUNT) {
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

template <int PXW, int PYH>
__global__ void __launch_bounds__(TileMap<PXW, PYH>::NT) rasterize_bwd_kernel(
    int64_t n_isects, const float4* __restrict__ splats, const int32_t* __restrict__ tile_offsets,
    const int32_t* __restrict__ flatten_ids, const float* __restrict__ backgrounds, int width, int height, int tile_w,
    int tile_h, int n_tiles_total, const float* __restrict__ render_alphas, const int32_t* __restrict__ last_ids,
    const float* __restrict__ v_render_colors, const float* __restrict__ v_render_alphas,
    float* __restrict__ v_splats) {
  using M = TileMap<PXW, PYH>;
  constexpr int NP = M::NP, NT = M::NT, RPT = M::RPT, NW = M::NW, kBatch = M::BATCH;
  __shared__ __align__(16) float4 sb[2][kBatch * 3];
  __shared__ int s_id[2][kBatch];
  __shared__ int s_warp_last[NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const TileRange tc = tile_setup<M>(tile_w, tile_h, n_isects, tile_offsets, n_tiles_total);
  if (tc.range_end <= tc.range_start) return;  // uniform for the block

  float pxf[PXW], pyf[PYH];
#pragma unroll
  for (int i = 0; i < PXW; ++i) pxf[i] = opaque((float)(tc.x0 + i) + 0.5f);
#pragma unroll
  for (int i = 0; i < PYH; ++i) pyf[i] = opaque((float)(tc.y0 + i) + 0.5f);

  const float wx_lo = opaque(tc.wx_lo), wx_hi = opaque(tc.wx_hi), wy_lo = opaque(tc.wy_lo), wy_hi = opaque(tc.wy_hi);
  // per-pixel replay state.  bdot = sum over the Gaussians behind of fac * (rgb . v_colour), which is all
  // the backward pass needs of the colour accumulated behind; tfv = T_final * (v_alpha_out - bg . v_colour).
  float T[NP], bdot[NP], vcr[NP], vcg[NP], vcb[NP], tfv[NP];
  int bin_final[NP];
  int my_last = -1;
  float bgr = 0.f, bgg = 0.f, bgb = 0.f;
  if (backgrounds != nullptr) {
    bgr = backgrounds[tc.cam * 3 + 0]; bgg = backgrounds[tc.cam * 3 + 1]; bgb = backgrounds[tc.cam * 3 + 2];
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int x = tc.x0 + j % PXW, y = tc.y0 + j / PXW;
    T[j] = 1.f; bdot[j] = 0.f; vcr[j] = 0.f; vcg[j] = 0.f; vcb[j] = 0.f; tfv[j] = 0.f;
    bin_final[j] = -1;  // pixels outside the image never match any index
    if (x < width && y < height) {
      const size_t pix = ((size_t)tc.cam * height + y) * width + x;
      const float T_final = 1.0f - render_alphas[pix];
      T[j] = T_final;
      bin_final[j] = last_ids[pix];
      vcr[j] = v_render_colors[pix * 3 + 0];
      vcg[j] = v_render_colors[pix * 3 + 1];
      vcb[j] = v_render_colors[pix * 3 + 2];
      tfv[j] = T_final * (v_render_alphas[pix] - (bgr * vcr[j] + bgg * vcg[j] + bgb * vcb[j]));
    }
    my_last = max(my_last, bin_final[j]);
  }

  const int warp_last = __reduce_max_sync(0xffffffffu, my_last);
  if (lane == 0) s_warp_last[warp] = warp_last;
  __syncthreads();
  int block_last = s_warp_last[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) block_last = max(block_last, s_warp_last[w]);
  // nothing behind the last blended Gaussian of any pixel of the tile can receive gradient
  const int end_idx = min(tc.range_end - 1, block_last);
  if (end_idx < tc.range_start) return;
  const int nb = (end_idx - tc.range_start + 1 + kBatch - 1) / kBatch;

  auto load_ids = [&](int b, int (&ids)[RPT]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int idx = end_idx - b * kBatch - (r * NT + tid);
      ids[r] = (b < nb && idx >= tc.range_start) ? __ldg(flatten_ids + idx) : -1;
    }
  };
  auto issue = [&](int buf, const int (&ids)[RPT]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int slot = r * NT + tid;
      if (ids[r] >= 0) {
        const float4* src = splats + (size_t)ids[r] * 3;
        float4* dst = &sb[buf][slot * 3];
        cp_async16(dst + 0, src + 0);
        cp_async16(dst + 1, src + 1);
        cp_async16(dst + 2, src + 2);
      }
      s_id[buf][slot] = ids[r];
    }
    cp_async_commit();
  };

  int ids[RPT];
  load_ids(0, ids);
  issue(0, ids);
  load_ids(1, ids);
  for (int b = 0; b < nb; ++b) {
    if (b + 1 < nb) {
      issue((b + 1) & 1, ids);
      load_ids(b + 2, ids);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int batch_end = end_idx - b * kBatch;  // sorted index held in slot 0 (the one furthest back)
    const int batch_size = min(kBatch, batch_end + 1 - tc.range_start);
    const float4* s = sb[b & 1];
    const int* sid = s_id[b & 1];
    for (int t = max(0, batch_end - warp_last); t < batch_size; ++t) {  // warp-uniform bounds
      const int idx = batch_end - t;
      const float4 g0 = s[t * 3 + 0];  // x, y, conic_a, conic_b
      const float4 g2 = s[t * 3 + 2];  // b, depth, r_eff^2, -
      {
        // warp-uniform skip (same test as the forward pass)
        const float ex = fmaxf(fmaxf(wx_lo - g0.x, g0.x - wx_hi), 0.f);
        const float ey = fmaxf(fmaxf(wy_lo - g0.y, g0.y - wy_hi), 0.f);
        if (fmaf(ex, ex, ey * ey) > g2.z) continue;
      }
      const float4 g1 = s[t * 3 + 1];  // conic_c, opacity, r, g
      const float la = (0.5f * kNegLog2e) * g0.z, lc = (0.5f * kNegLog2e) * g1.x, lb = kNegLog2e * g0.w;
      float dx[PXW], qx[PXW], bx[PXW], dy[PYH], qy[PYH];
#pragma unroll
      for (int i = 0; i < PXW; ++i) { dx[i] = g0.x - pxf[i]; qx[i] = la * dx[i] * dx[i]; bx[i] = lb * dx[i]; }
#pragma unroll
      for (int i = 0; i < PYH; ++i) { dy[i] = g0.y - pyf[i]; qy[i] = lc * dy[i] * dy[i]; }
      float ov[NP];  // opacity * exp(-sigma), before the 0.999 clamp
      bool valid[NP];
      bool any_valid = false;
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float q = fmaf(bx[j % PXW], dy[j / PXW], qx[j % PXW] + qy[j / PXW]);  // -log2(e) * sigma
        ov[j] = g1.y * fast_ex2(q);
        valid[j] = idx <= bin_final[j] && q <= 0.f && ov[j] >= kAlphaMin;  // min(.999, ov) >= 1/255 <=> ov >= 1/255
        any_valid = any_valid || valid[j];
      }
      if (!__any_sync(0xffffffffu, any_valid)) continue;  // warp-uniform
      const float cbl = g2.x;
      const float inv_o = __fdividef(1.0f, g1.y);
      // v[2], v[3], v[4] accumulate sx*dx, sx*dy, sy*dy (the 0.5 of the conic gradient is applied once, after
      // the warp reduction); v[5] accumulates ov * v_alpha (the 1/opacity is applied after the reduction).
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.f;
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        if (valid[j]) {
          const float ddx = dx[j % PXW], ddy = dy[j / PXW];
          const float alpha = fminf(kAlphaMax, ov[j]);
          const float ra = __fdividef(1.0f, 1.0f - alpha);
          T[j] *= ra;  // transmittance in front of this Gaussian
          const float fac = alpha * T[j];
          v[6] = fmaf(fac, vcr[j], v[6]);
          v[7] = fmaf(fac, vcg[j], v[7]);
          v[8] = fmaf(fac, vcb[j], v[8]);
          const float cdot = fmaf(cbl, vcb[j], fmaf(g1.w, vcg[j], g1.z * vcr[j]));
          const float v_alpha = fmaf(T[j], cdot, -ra * (bdot[j] - tfv[j]));
          bdot[j] = fmaf(cdot, fac, bdot[j]);
          if (ov[j] <= kAlphaMax) {  // the clamp was inactive: alpha depends on sigma and opacity
            const float w = ov[j] * v_alpha;  // = opacity * d(alpha)/d(opacity) * v_alpha = -v_sigma
            const float sx = -w * ddx, sy = -w * ddy;
            v[2] = fmaf(sx, ddx, v[2]);
            v[3] = fmaf(sx, ddy, v[3]);
            v[4] = fmaf(sy, ddy, v[4]);
            const float gx = fmaf(g0.w, sy, g0.z * sx);
            const float gy = fmaf(g1.x, sy, g0.w * sx);
            v[0] += gx;
            v[1] += gy;
            v[9] += fabsf(gx);
            v[10] += fabsf(gy);
            v[5] += w;
          }
        }
      }
      v[2] *= 0.5f; v[4] *= 0.5f; v[5] *= inv_o;
      const float total = warp_reduce_scatter16(v, lane);
      const int slot = lane >> 1;
      if ((lane & 1) == 0 && slot < 11) atomicAdd(v_splats + (size_t)sid[t] * EGS_SPLAT_FLOATS + slot, total);
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

// Pixel-block variant: 22 = 2x2 pixels per thread (default), 24 = 2x4 (one warp per tile), 21 = 2x1, 11 = 1x1.  The environment
// variable EGS_BLEND_VARIANT exists for tuning runs only.
static int blend_variant() {
  static int v = [] {
    const char* e = getenv("EGS_BLEND_VARIANT");
    int x = e ? atoi(e) : 22;
    return (x == 11 || x == 21 || x == 22 || x == 24) ? x : 22;
  }();
  return v;
}

template <bool COUNT>
static int launch_fwd(int32_t C, int64_t n_isects, const float* splats, const int32_t* tile_offsets,
                      const int32_t* flatten_ids, const float* backgrounds, int32_t width, int32_t height,
                      int32_t tile_width, int32_t tile_height, float* render_colors, float* render_alphas,
                      int32_t* last_ids, uint64_t* pair_counters, egs_stream_t stream) {
  dim3 grid(tile_width, tile_height, C);
  const float4* sp = reinterpret_cast<const float4*>(splats);
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(pair_counters);
  const int nt = C * tile_width * tile_height;
  cudaStream_t st = (cudaStream_t)stream;
#define EGS_LAUNCH_FWD(PX, PY)                                                                                    \
  rasterize_fwd_kernel<PX, PY, COUNT><<<grid, TileMap<PX, PY>::NT, 0, st>>>(                                       \
      n_isects, sp, tile_offsets, flatten_ids, backgrounds, width, height, tile_width, tile_height, nt,           \
      render_colors, render_alphas, last_ids, pc)
  switch (blend_variant()) {
    case 11: EGS_LAUNCH_FWD(1, 1); break;
    case 21: EGS_LAUNCH_FWD(2, 1); break;
    case 24: EGS_LAUNCH_FWD(2, 4); break;
    default: EGS_LAUNCH_FWD(2, 2); break;
  }
#undef EGS_LAUNCH_FWD
  return check_launch("rasterize_fwd_kernel");
}

extern "C" int egs_rasterize_fwd(int32_t C, int32_t N, int64_t n_isects, const float* splats,
                                 const int32_t* tile_offsets, const int32_t* flatten_ids, const float* backgrounds,
                                 int32_t width, int32_t height, int32_t tile_width, int32_t tile_height,
                                 float* render_colors, float* render_alphas, int32_t* last_ids, egs_stream_t stream) {
  (void)N;
  if (int rc = check_raster_args("rasterize_fwd", C, n_isects, width, height, tile_width, tile_height)) return rc;
  if (C == 0) return 0;
  return launch_fwd<false>(C, n_isects, splats, tile_offsets, flatten_ids, backgrounds, width, height, tile_width,
                           tile_height, render_colors, render_alphas, last_ids, nullptr, stream);
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
  return launch_fwd<true>(C, n_isects, splats, tile_offsets, flatten_ids, backgrounds, width, height, tile_width,
                          tile_height, render_colors, render_alphas, last_ids, pair_counters, stream);
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
  const float4* sp = reinterpret_cast<const float4*>(splats);
  const int nt = C * tile_width * tile_height;
  cudaStream_t st = (cudaStream_t)stream;
#define EGS_LAUNCH_BWD(PX, PY)                                                                                   \
  rasterize_bwd_kernel<PX, PY><<<grid, TileMap<PX, PY>::NT, 0, st>>>(                                             \
      n_isects, sp, tile_offsets, flatten_ids, backgrounds, width, height, tile_width, tile_height, nt,          \
      render_alphas, last_ids, v_render_colors, v_render_alphas, v_splats)
  switch (blend_variant()) {
    case 11: EGS_LAUNCH_BWD(1, 1); break;
    case 21: EGS_LAUNCH_BWD(2, 1); break;
    case 24: EGS_LAUNCH_BWD(2, 4); break;
    default: EGS_LAUNCH_BWD(2, 2); break;
  }
#undef EGS_LAUNCH_BWD
  return check_launch("rasterize_bwd_kernel");
}
