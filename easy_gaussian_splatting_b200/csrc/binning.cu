// g3 (scan + key emission) and g5 (tile offsets): integer / bitwise HBM-bound kernels.
// Compiled with -fmad=false like projection.cu because tile_rect() must reproduce the tile
// rectangle the projection kernel counted.
#include "egs_common.cuh"
#include "egs_math.cuh"

namespace egs {

// ------------------------------------------------------------------------------------------------
// Exclusive scan int32 -> int64.  Three small launches (block sums, scan of block sums, apply):
// 12 B/element of traffic, launch-latency class at the sizes of this path (N <= a few million).
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                         // per thread
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 per block

__device__ __forceinline__ int64_t warp_inclusive_scan(int64_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// inclusive scan over the block; returns this thread's inclusive prefix and the block total
__device__ __forceinline__ int64_t block_inclusive_scan(int64_t v, int64_t* smem /*[33]*/, int64_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t inc = warp_inclusive_scan(v, lane);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int64_t w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0;
    int64_t wi = warp_inclusive_scan(w, lane);
    smem[lane] = wi - w;  // exclusive prefix of warp totals
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  inc += smem[warp];
  total = smem[32];
  __syncthreads();
  return inc;
}

__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                        int64_t* __restrict__ block_sums) {
  __shared__ int64_t smem[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t j = base + (int64_t)i * kScanThreads + threadIdx.x;
    if (j < n) s += in[j];
  }
  int64_t total;
  block_inclusive_scan(s, smem, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums in place, grand total to *total
__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(int64_t* __restrict__ block_sums, int64_t nblocks,
                                                                   int64_t* __restrict__ total_out) {
  __shared__ int64_t smem[33];
  int64_t carry = 0;
  for (int64_t base = 0; base < nblocks; base += kScanThreads) {
    int64_t j = base + threadIdx.x;
    int64_t v = (j < nblocks) ? block_sums[j] : 0;
    int64_t total;
    int64_t inc = block_inclusive_scan(v, smem, total);
    if (j < nblocks) block_sums[j] = carry + inc - v;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                   const int64_t* __restrict__ block_sums,
                                                                   int64_t* __restrict__ out) {
  __shared__ int64_t smem[33];
  // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int64_t total;
  int64_t inc = block_inclusive_scan(s, smem, total);
  int64_t run = block_sums[blockIdx.x] + inc - s;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Key emission: one thread per (camera, Gaussian); a visible Gaussian writes its tile rectangle
// row-major at its scanned offset.
// ------------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;

__global__ void __launch_bounds__(kEmitThreads) isect_emit_kernel(
    int C, int N, const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
    const float* __restrict__ depths, const int64_t* __restrict__ cum_excl, float tile_size, int tile_w, int tile_h,
    int tile_n_bits, int64_t n_isects, int64_t* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids) {
  const int64_t idx = (int64_t)blockIdx.x * kEmitThreads + threadIdx.x;
  if (idx >= (int64_t)C * N) return;
  const int32_t r = radii[idx];
  if (r <= 0) return;
  const float2 m = means2d[idx];
  int32_t x0, y0, x1, y1;
  tile_rect(m.x, m.y, r, tile_size, tile_w, tile_h, x0, y0, x1, y1);
  const int64_t cam = idx / N;
  const uint64_t hi = (uint64_t)cam << tile_n_bits;
  const uint64_t depth_bits = (uint64_t)__float_as_uint(depths[idx]);
  int64_t out = cum_excl[idx];
  for (int32_t y = y0; y < y1; ++y) {
    for (int32_t x = x0; x < x1; ++x) {
      if (out < n_isects) {  // guards against a caller passing a short buffer
        uint64_t tile = (uint64_t)(y * tile_w + x);
        isect_ids[out] = (int64_t)(((hi | tile) << 32) | depth_bits);
        flatten_ids[out] = (int32_t)idx;
      }
      ++out;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tile offsets from the sorted keys.
// ------------------------------------------------------------------------------------------------
constexpr int kOffThreads = 256;

__global__ void __launch_bounds__(kOffThreads) isect_offset_encode_kernel(int64_t n_isects,
                                                                          const int64_t* __restrict__ ids,
                                                                          int n_tiles, int tile_n_bits, int64_t n_slots,
                                                                          int32_t* __restrict__ offsets) {
  const int64_t i = (int64_t)blockIdx.x * kOffThreads + threadIdx.x;
  if (i >= n_isects) return;
  const uint64_t tile_mask = (1ull << tile_n_bits) - 1ull;
  const uint64_t hi = (uint64_t)ids[i] >> 32;
  const int64_t cur = (int64_t)(hi >> tile_n_bits) * n_tiles + (int64_t)(hi & tile_mask);
  if (i == 0) {
    for (int64_t t = 0; t <= cur && t < n_slots; ++t) offsets[t] = 0;
  } else {
    const uint64_t hp = (uint64_t)ids[i - 1] >> 32;
    const int64_t prev = (int64_t)(hp >> tile_n_bits) * n_tiles + (int64_t)(hp & tile_mask);
    for (int64_t t = prev + 1; t <= cur && t < n_slots; ++t) offsets[t] = (int32_t)i;
  }
  if (i == n_isects - 1) {
    for (int64_t t = cur + 1; t < n_slots; ++t) offsets[t] = (int32_t)n_isects;
  }
}


// ================================================================================================
// Fast path of g3-g5: same sorted (isect_ids, flatten_ids, offsets) as emit + 64-bit sort + offset
// encode, with ~3x less HBM traffic.  A stable sort on (cam | tile | depth) keys equals
//   (1) a stable sort of the VISIBLE Gaussians of all cameras on depth  [N_vis pairs, not n_isects; 32-bit keys]
//   (2) emitting their tiles in that order, and
//   (3) a stable sort of the emitted pairs on the 13..19-bit (cam, tile) index alone [2-3 passes of 8 B pairs]
// because (3) keeps, inside every tile, the (depth, flat index) order that (1) established.
// Launches of the route (egs_isect_visible_keys + egs_isect_sorted): visible keys | histogram + 4 passes | scan + emit |
// histogram + 2-3 passes (the last one records every key's first position) | tile offsets | tile order = 12-13, with
// one memset of the control block and one of the offsets.  The separate operators further down (three-launch scans,
// egs_isect_emit_sorted, egs_isect_finalize with its offsets4 / 64-bit-key kernels) are the same steps as C-ABI
// entries of their own; isect_finalize also rebuilds the 64-bit keys for `meta["isect_ids"]`.
// ================================================================================================

__device__ __forceinline__ int64_t block_reduce_sum(int64_t v, int64_t* smem /*[33]*/) {
  int64_t total;
  block_inclusive_scan(v, smem, total);
  return total;
}

// ---- decoupled look-back chain shared by the single-launch scans below --------------------------------------------
constexpr uint64_t kChainAggregate = 1ull << 62, kChainPrefix = 2ull << 62, kChainValue = (1ull << 62) - 1ull;

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix of `total` over the blocks before `blk` (ticket order).  Called by one full warp; status words
// start at zero.  Per trip a lane polls two blocks, blk-1-lane and blk-33-lane (both loads in flight together: the chain
// advances up to 64 blocks per L2 round trip); everything up to the first block that has not published yet is consumed,
// the walk ends at the first inclusive prefix.
__device__ __forceinline__ int64_t chain_lookback(uint64_t* __restrict__ status, uint32_t blk, int64_t total, int lane) {
  if (lane == 0) st_volatile_u64(status + blk, (blk == 0 ? kChainPrefix : kChainAggregate) | (uint64_t)total);
  int64_t excl = 0;
  if (blk > 0) {
    int64_t j = (int64_t)blk - 1;
    bool done = false;
    while (!done) {
      uint64_t v[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int64_t jj = j - 32 * q - lane;
        v[q] = jj >= 0 ? ld_volatile_u64(status + jj) : kChainPrefix;  // before block 0: an empty prefix
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (done) break;
        const uint32_t flag = (uint32_t)(v[q] >> 62);
        const uint32_t not_ready = __ballot_sync(0xffffffffu, flag == 0u);
        const uint32_t is_prefix = __ballot_sync(0xffffffffu, flag == 2u);
        const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
        const int first_pf = is_prefix ? __ffs(is_prefix) - 1 : 32;
        const int take = first_pf < first_nr ? first_pf + 1 : first_nr;
        int64_t c = lane < take ? (int64_t)(v[q] & kChainValue) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        excl += c;
        j -= take;
        if (first_pf < first_nr) done = true;
        else if (take < 32) break;  // a block in this half has nothing yet: poll again from it
      }
    }
    if (lane == 0) st_volatile_u64(status + blk, kChainPrefix | (uint64_t)(excl + total));
  }
  return excl;
}

// The same chain with MIN instead of SUM (value = a 32-bit position; 0xffffffff = nothing): the minimum over the blocks
// before `blk` in ticket order.
__device__ __forceinline__ uint32_t chain_lookback_min(uint64_t* __restrict__ status, uint32_t blk, uint32_t own, int lane) {
  if (lane == 0) st_volatile_u64(status + blk, (blk == 0 ? kChainPrefix : kChainAggregate) | (uint64_t)own);
  uint32_t excl = 0xffffffffu;
  if (blk > 0) {
    int64_t j = (int64_t)blk - 1;
    bool done = false;
    while (!done) {
      const int64_t jj = j - lane;
      const uint64_t v = jj >= 0 ? ld_volatile_u64(status + jj) : (kChainPrefix | 0xffffffffull);
      const uint32_t flag = (uint32_t)(v >> 62);
      const uint32_t not_ready = __ballot_sync(0xffffffffu, flag == 0u);
      const uint32_t is_prefix = __ballot_sync(0xffffffffu, flag == 2u);
      const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
      const int first_pf = is_prefix ? __ffs(is_prefix) - 1 : 32;
      const int take = first_pf < first_nr ? first_pf + 1 : first_nr;
      const uint32_t c = lane < take ? (uint32_t)(v & 0xffffffffull) : 0xffffffffu;
      excl = min(excl, __reduce_min_sync(0xffffffffu, c));
      j -= take;
      done = first_pf < first_nr;
    }
    if (lane == 0) st_volatile_u64(status + blk, kChainPrefix | (uint64_t)min(excl, own));
  }
  return excl;
}

// ---- visible compaction + level-1 sort input in ONE launch --------------------------------------------------------
// key = bits(depth), value = flat index (camera * N + Gaussian), compacted; totals = {n_vis, sum of tiles, 0, 0}.
// The camera needs no key bits: level 2 sorts stably on the (camera, tile) index, so sorting the visible entries of
// ALL cameras on depth alone (ties keep the flat-index order they are written in here) leaves every (camera, tile)
// bucket in (depth, flat index) order — the order of a sort on cam | tile | depth.
// control: {ticket u32, finished u32, tile sum u64} then one chain status word per block, all zero at launch.
// A block takes 8192 entries (warp w: 512 consecutive ones, 32 per round, so loads and the compacted stores are
// both coalesced; a thread that owns consecutive entries scatters 4-byte stores over 32 sectors per instruction).
struct VisibleControl { uint32_t ticket, finished; unsigned long long tile_sum; };
constexpr int kVisThreads = 512;
constexpr int kVisRounds = 16;
constexpr int kVisTile = kVisThreads * kVisRounds;

__global__ void __launch_bounds__(kVisThreads) visible_keys_kernel(const int32_t* __restrict__ tiles, int64_t n,
                                                                    const float* __restrict__ depths,
                                                                    uint32_t* __restrict__ keys1, uint32_t* __restrict__ vals1,
                                                                    int64_t* __restrict__ totals, VisibleControl* __restrict__ control,
                                                                    uint64_t* __restrict__ status) {
  __shared__ int32_t warp_excl[kVisThreads / 32];
  __shared__ int64_t warp_tiles[kVisThreads / 32];
  __shared__ int64_t block_base;
  __shared__ uint32_t block_ticket;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) block_ticket = atomicAdd(&control->ticket, 1u);
  __syncthreads();
  const uint32_t blk = block_ticket;
  const int64_t warp_first = (int64_t)blk * kVisTile + (int64_t)warp * (32 * kVisRounds);
  int32_t v[kVisRounds];
  float z[kVisRounds];
#pragma unroll
  for (int r = 0; r < kVisRounds; ++r) {
    const int64_t i = warp_first + r * 32 + lane;
    v[r] = i < n ? __ldg(tiles + i) : 0;
  }
#pragma unroll
  for (int r = 0; r < kVisRounds; ++r) {
    const int64_t i = warp_first + r * 32 + lane;
    z[r] = i < n ? __ldg(depths + i) : 0.f;  // unconditional: no dependence on the count (culled entries hold 0)
  }
  uint32_t mask[kVisRounds];
  int32_t warp_total = 0;
  int64_t t = 0;
#pragma unroll
  for (int r = 0; r < kVisRounds; ++r) {
    mask[r] = __ballot_sync(0xffffffffu, v[r] > 0);
    warp_total += __popc(mask[r]);
    t += v[r];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  if (lane == 0) { warp_excl[warp] = warp_total; warp_tiles[warp] = t; }
  __syncthreads();
  if (warp == 0) {
    const int32_t wt = lane < kVisThreads / 32 ? warp_excl[lane] : 0;
    int64_t tt = lane < kVisThreads / 32 ? warp_tiles[lane] : 0;
    int32_t wi = wt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t u = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += u;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, d);
    const int32_t total = __shfl_sync(0xffffffffu, wi, 31);
    if (lane < kVisThreads / 32) warp_excl[lane] = wi - wt;
    const int64_t excl = chain_lookback(status, blk, total, lane);
    if (lane == 0) {
      block_base = excl;
      if ((int64_t)(blk + 1) * kVisTile >= n) totals[0] = excl + total;  // the last block of the chain
      atomicAdd(&control->tile_sum, (unsigned long long)tt);
      __threadfence();
      if (atomicAdd(&control->finished, 1u) == gridDim.x - 1) {  // every block's sum has landed
        totals[1] = (int64_t)atomicAdd(&control->tile_sum, 0ull);
        totals[2] = 0; totals[3] = 0;
      }
    }
  }
  __syncthreads();
  int64_t run = block_base + warp_excl[warp];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kVisRounds; ++r) {
    if ((mask[r] >> lane) & 1u) {
      const int64_t dst = run + __popc(mask[r] & lt);
      keys1[dst] = __float_as_uint(z[r]);
      vals1[dst] = (uint32_t)(warp_first + r * 32 + lane);
    }
    run += __popc(mask[r]);
  }
}

// ---- exclusive scan of src[gather[i]] (tile counts in depth order) -------------------------------------
__global__ void __launch_bounds__(kScanThreads) gscan_block_sums_kernel(const int32_t* __restrict__ src,
                                                                         const uint32_t* __restrict__ gather, int64_t n,
                                                                         const int64_t* __restrict__ n_dev,
                                                                         int64_t* __restrict__ block_sums) {
  __shared__ int64_t smem[33];
  n = live_count(n, n_dev);
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t j = base + (int64_t)i * kScanThreads + threadIdx.x;
    if (j < n) s += src[gather[j]];
  }
  int64_t total;
  block_inclusive_scan(s, smem, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) gscan_apply_kernel(const int32_t* __restrict__ src,
                                                                    const uint32_t* __restrict__ gather, int64_t n,
                                                                    const int64_t* __restrict__ n_dev,
                                                                    const int64_t* __restrict__ block_sums,
                                                                    int64_t* __restrict__ out) {
  __shared__ int64_t smem[33];
  n = live_count(n, n_dev);
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? src[gather[base + i]] : 0;
    s += v[i];
  }
  int64_t total;
  int64_t inc = block_inclusive_scan(s, smem, total);
  int64_t run = block_sums[blockIdx.x] + inc - s;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

// ---- warp-cooperative emission in depth order: coalesced, balanced -------------------------------------
__global__ void __launch_bounds__(kEmitThreads) isect_emit_sorted_kernel(
    int N, int64_t n_vis, const uint32_t* __restrict__ order, const int64_t* __restrict__ cum_excl,
    const float2* __restrict__ means2d, const int32_t* __restrict__ radii, float tile_size, int tile_w, int tile_h,
    int64_t n_isects, uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ flat_vals,
    const int64_t* __restrict__ counts_dev /* nullable: {n_vis, n_isects} live counts, the arguments are capacities */,
    const float4* __restrict__ splats /* nullable: emit the TIGHT rectangles (tighten_tile_rect) */) {
  if (counts_dev != nullptr) {
    n_vis = live_count(n_vis, counts_dev);
    n_isects = live_count(n_isects, counts_dev + 1);
  }
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kEmitThreads + threadIdx.x;  // position in (cam, depth) order
  const bool valid = i < n_vis;
  uint32_t g = 0;
  int32_t x0 = 0, y0 = 0, w = 1, cnt = 0;
  int64_t excl = n_isects;
  if (valid) {
    g = order[i];
    const float2 m = means2d[g];
    int32_t x1, y1;
    tile_rect(m.x, m.y, radii[g], tile_size, tile_w, tile_h, x0, y0, x1, y1);
    if (splats != nullptr) {
      const float4 g0 = splats[(size_t)g * 3 + 0];
      tighten_tile_rect(g0.x, g0.y, g0.z, g0.w, splats[(size_t)g * 3 + 1].x, splats[(size_t)g * 3 + 2].w, tile_size, x0, y0, x1, y1);
    }
    w = max(x1 - x0, 1);
    cnt = (x1 - x0) * (y1 - y0);
    excl = cum_excl[i];
  }
  // j / w for every tile j of the rectangle without a division per pair: one reciprocal per Gaussian,
  // floor(j * ceil(2^32 / w) / 2^32) == j / w exactly while j * w < 2^32 (j < tile_w * tile_h < 2^31 / ... holds: the
  // entry point refuses grids of 2^16 tiles or more per side product, see EGS_REQUIRE below)
  const uint32_t wmagic = w > 1 ? (uint32_t)((0x100000000ull + (uint32_t)w - 1u) / (uint32_t)w) : 0u;
  const int64_t warp_base = __shfl_sync(0xffffffffu, excl, 0);
  const int32_t lexcl = (int32_t)(excl - warp_base);
  const int32_t total = __shfl_sync(0xffffffffu, lexcl + cnt, 31);
  const uint32_t key_base = (uint32_t)(g / (uint32_t)N) * (uint32_t)(tile_w * tile_h);
  for (int32_t k0 = 0; k0 < total; k0 += 32) {
    const int32_t k = k0 + lane;
    // owner = last lane whose run starts at or before k (runs of valid lanes are non-empty, so starts increase)
    int o = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int32_t e = __shfl_sync(0xffffffffu, lexcl, o + step);
      if (e <= k) o += step;
    }
    const int32_t j = k - __shfl_sync(0xffffffffu, lexcl, o);
    const int32_t ow = __shfl_sync(0xffffffffu, w, o);
    const uint32_t omagic = __shfl_sync(0xffffffffu, wmagic, o);
    const int32_t ox0 = __shfl_sync(0xffffffffu, x0, o);
    const int32_t oy0 = __shfl_sync(0xffffffffu, y0, o);
    const uint32_t okb = __shfl_sync(0xffffffffu, key_base, o);
    const uint32_t og = __shfl_sync(0xffffffffu, g, o);
    if (k < total) {
      const int32_t ty = ow > 1 ? (int32_t)__umulhi((uint32_t)j, omagic) : j, tx = j - ty * ow;
      const int64_t dst = warp_base + k;
      if (dst < n_isects) {
        tile_keys[dst] = okb + (uint32_t)((oy0 + ty) * tile_w + ox0 + tx);
        flat_vals[dst] = og;
      }
    }
  }
}

// ---- the same emission with the scan folded in (the product route) --------------------------------------------------
// One launch instead of four (block sums, spine, apply, emit): a thread block scans the tile counts of its 256
// Gaussians, chains its total to the blocks before it with decoupled look-back (one 64-bit status word per block:
// 2 flag bits | running count), and emits.  The tight rectangle arrives packed in 8 bytes from the projection kernel
// (pack_tile_rect) — one gather per Gaussian instead of the 60 bytes of record, mean and radius the rectangle was
// recomputed from (that set-up was two thirds of the old kernel's instructions, profiles/r2p).
// A block takes kEmitRounds x 256 Gaussians (warp w: 128 consecutive ones, 32 per round): the chain advances 32 blocks
// per L2 round trip, so the Gaussians per block set how fast the prefix can travel (256 per block: 119 us for the
// 3.3 M visible Gaussians of the benchmark step, chain-bound; gpurun_out/launches_r2r.csv).
#ifndef EGS_EMIT_ROUNDS
#define EGS_EMIT_ROUNDS 4
#endif
constexpr int kEmitRounds = EGS_EMIT_ROUNDS;
constexpr int kEmitBlock = kEmitThreads * kEmitRounds;

__global__ void __launch_bounds__(kEmitThreads) isect_scan_emit_kernel(
    int N, int64_t n_vis, const int64_t* __restrict__ n_vis_dev, const uint32_t* __restrict__ order,
    const int2* __restrict__ tight_rects /* nullable: classic rectangles from means2d / radii */,
    const float2* __restrict__ means2d, const int32_t* __restrict__ radii, float tile_size, int tile_w, int tile_h,
    int64_t capacity, uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ flat_vals,
    int64_t* __restrict__ n_isects_out, uint32_t* __restrict__ ticket, uint64_t* __restrict__ status) {
  __shared__ int32_t warp_excl[kEmitThreads / 32];
  __shared__ int64_t block_base;
  __shared__ uint32_t block_ticket;
  __shared__ __align__(16) uint4 run_table[kEmitThreads / 32][32];
  __shared__ uint32_t run_gauss[kEmitThreads / 32][32];
  n_vis = live_count(n_vis, n_vis_dev);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the grid covers C * N: blocks past the visible entries leave before they take a ticket, so exactly the live
  // blocks draw tickets 0 .. live-1 (block 0 is spare only when nothing is visible)
  if ((int64_t)blockIdx.x * kEmitBlock >= n_vis) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_isects_out = 0;
    return;
  }
  // ticket order = chain order: a block only ever waits for blocks that are already running
  if (threadIdx.x == 0) block_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t blk = block_ticket;
  const int64_t first = (int64_t)blk * kEmitBlock;
  const int64_t warp_first = first + warp * (32 * kEmitRounds);  // position in depth order of this warp's first Gaussian
  uint32_t g[kEmitRounds];
#pragma unroll
  for (int r = 0; r < kEmitRounds; ++r) {
    const int64_t i = warp_first + r * 32 + lane;
    g[r] = i < n_vis ? order[i] : 0xffffffffu;
  }
  int2 rect[kEmitRounds];
  if (tight_rects != nullptr) {  // all gathers of the block in flight together
#pragma unroll
    for (int r = 0; r < kEmitRounds; ++r) rect[r] = g[r] != 0xffffffffu ? __ldg(tight_rects + g[r]) : make_int2(0, 0);
  } else {
#pragma unroll
    for (int r = 0; r < kEmitRounds; ++r) {
      rect[r] = make_int2(0, 0);
      if (g[r] != 0xffffffffu) {
        const float2 m = means2d[g[r]];
        int32_t x0, y0, x1, y1;
        tile_rect(m.x, m.y, radii[g[r]], tile_size, tile_w, tile_h, x0, y0, x1, y1);
        pack_tile_rect(x0, y0, x1, y1, rect[r].x, rect[r].y);
      }
    }
  }
  int32_t cnt[kEmitRounds], lexcl[kEmitRounds], round_total[kEmitRounds];
  int32_t warp_total = 0;
#pragma unroll
  for (int r = 0; r < kEmitRounds; ++r) {
    if (g[r] == 0xffffffffu) g[r] = 0;
    cnt[r] = (rect[r].y & 0xffff) * (int32_t)((uint32_t)rect[r].y >> 16);
    // counts of one block stay far below 2^31 (1024 rectangles of < 2^21 tiles each: the entry point checks the grid)
    int32_t inc = cnt[r];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    lexcl[r] = inc - cnt[r];
    round_total[r] = __shfl_sync(0xffffffffu, inc, 31);
    warp_total += round_total[r];
  }
  if (lane == 0) warp_excl[warp] = warp_total;
  __syncthreads();
  if (warp == 0) {
    const int32_t wt = lane < kEmitThreads / 32 ? warp_excl[lane] : 0;
    int32_t wi = wt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    const int32_t block_total = __shfl_sync(0xffffffffu, wi, 31);
    if (lane < kEmitThreads / 32) warp_excl[lane] = wi - wt;
    const int64_t excl = chain_lookback(status, blk, block_total, lane);
    if (lane == 0) {
      block_base = excl;
      if (first + kEmitBlock >= n_vis) *n_isects_out = excl + block_total;  // the block that holds the last entry
    }
  }
  __syncthreads();
  // Emission, 32 consecutive entries per trip.  The non-empty runs of a round are compacted into a per-warp table
  // {start, row reciprocal, key base - start, row stride} + Gaussian; the owner of entry k is then
  // (#runs that start before the trip's window) + (#runs that start inside it at or before k) - 1: one ballot, one
  // warp-wide OR (redux.sync) and two popc instead of a five-step shuffle search, and two shared-memory loads
  // instead of five shuffles for the owner's data.
  int64_t round_base = block_base + warp_excl[warp];
  uint4* tab = run_table[warp];
  uint32_t* tab_g = run_gauss[warp];
  const uint32_t le_mask = (2u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kEmitRounds; ++r) {
    const int32_t total = round_total[r];
    const uint32_t nonempty = __ballot_sync(0xffffffffu, cnt[r] > 0);
    if (cnt[r] > 0) {
      const int32_t w = rect[r].y & 0xffff;
      // floor(j * ceil(2^32 / w) / 2^32) == j / w exactly while j * w < 2^32 (the entry point checks the tile grid);
      // 0 marks w == 1 (the reciprocal would be 2^32)
      const uint32_t magic = w > 1 ? 0xffffffffu / (uint32_t)w + 1u : 0u;
      const uint32_t key_base = (g[r] / (uint32_t)N) * (uint32_t)(tile_w * tile_h) +
                                (uint32_t)((int32_t)((uint32_t)rect[r].x >> 16) * tile_w + (rect[r].x & 0xffff));
      const int slot = __popc(nonempty & (le_mask >> 1));
      tab[slot] = make_uint4((uint32_t)lexcl[r], magic, key_base - (uint32_t)lexcl[r], (uint32_t)(tile_w - w));
      tab_g[slot] = g[r];
    }
    __syncwarp();
    const int32_t start = lane < __popc(nonempty) ? (int32_t)tab[lane].x : 0x7fffffff;  // start of the lane-th run
    for (int32_t k0 = 0; k0 < total; k0 += 32) {
      const int32_t k = k0 + lane;
      const uint32_t before = __popc(__ballot_sync(0xffffffffu, start < k0));
      const uint32_t d = (uint32_t)(start - k0);
      const uint32_t inside = __reduce_or_sync(0xffffffffu, d < 32u ? 1u << d : 0u);
      const int o = (int)(before + __popc(inside & le_mask)) - 1;  // >= 0: the first run starts at entry 0
      const int64_t dst = round_base + k;
      if (k < total && dst < capacity) {
        const uint4 run = tab[o];
        const uint32_t j = (uint32_t)k - run.x;
        const uint32_t ty = run.y ? __umulhi(j, run.y) : j;
        tile_keys[dst] = run.z + (uint32_t)k + ty * run.w;  // key base + j + row * (tile_w - w)
        flat_vals[dst] = tab_g[o];
      }
    }
    __syncwarp();  // the table is rewritten by the next round
    round_base += total;
  }
}

// ---- tile offsets + 64-bit keys from the (cam,tile)-sorted pairs ---------------------------------------
__global__ void __launch_bounds__(kOffThreads) isect_finalize_kernel(int64_t n_isects, const uint32_t* __restrict__ tile_keys,
                                                                     const uint32_t* __restrict__ flat_vals,
                                                                     const float* __restrict__ depths, int n_tiles,
                                                                     int tile_n_bits, int64_t n_slots,
                                                                     int64_t* __restrict__ isect_ids,
                                                                     int32_t* __restrict__ offsets) {
  const int64_t i = (int64_t)blockIdx.x * kOffThreads + threadIdx.x;
  if (i >= n_isects) return;
  const int64_t cur = tile_keys[i];
  if (isect_ids != nullptr) {
    const uint64_t cam = (uint64_t)(cur / n_tiles), tile = (uint64_t)(cur % n_tiles);
    const uint64_t depth_bits = (uint64_t)__float_as_uint(depths[flat_vals[i]]);
    isect_ids[i] = (int64_t)(((((cam << tile_n_bits) | tile)) << 32) | depth_bits);
  }
  if (offsets == nullptr) return;
  if (i == 0) {
    for (int64_t t = 0; t <= cur && t < n_slots; ++t) offsets[t] = 0;
  } else {
    const int64_t prev = tile_keys[i - 1];
    for (int64_t t = prev + 1; t <= cur && t < n_slots; ++t) offsets[t] = (int32_t)i;
  }
  if (i == n_isects - 1) {
    for (int64_t t = cur + 1; t < n_slots; ++t) offsets[t] = (int32_t)n_isects;
  }
}

// Offsets only (the product path: the 64-bit keys are rebuilt lazily, on the rare read of meta["isect_ids"]).
// HBM-bound shape: one 128-bit load of 4 sorted keys per thread plus the key in front of them, 32-bit arithmetic;
// tile t's offset is written by the thread that sees the first key >= t (runs of empty tiles are filled by it too).
constexpr int kOff4Threads = 256;
__global__ void __launch_bounds__(kOff4Threads) isect_offsets4_kernel(int32_t n_isects, const uint32_t* __restrict__ tile_keys,
                                                                       int32_t n_slots, int32_t* __restrict__ offsets,
                                                                       const int64_t* __restrict__ n_dev, int sentinel) {
  // n_dev: the live count (the argument is then the capacity the grid was sized for); sentinel: offsets has
  // n_slots + 1 entries and the last one receives the live count (the blend kernels read it as the end of the last tile)
  n_isects = (int32_t)live_count(n_isects, n_dev);
  const int32_t i0 = (blockIdx.x * kOff4Threads + threadIdx.x) * 4;
  if (i0 == 0 && sentinel) offsets[n_slots] = n_isects;
  if (n_isects == 0) {  // nothing visible: every tile is empty (grid-uniform branch)
    for (int32_t t = i0 / 4; t < n_slots; t += gridDim.x * kOff4Threads) offsets[t] = 0;
    return;
  }
  if (i0 >= n_isects) return;
  uint32_t k[4];
  if (i0 + 4 <= n_isects) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(tile_keys + i0));
    k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = tile_keys[min(i0 + j, n_isects - 1)];
  }
  uint32_t prev = i0 > 0 ? __ldg(tile_keys + i0 - 1) : 0xffffffffu;  // 0xffffffff + 1 = 0: tile 0 starts the fill
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int32_t i = i0 + j;
    if (i < n_isects && k[j] != prev) {
      for (uint32_t t = prev + 1u; t <= k[j] && t < (uint32_t)n_slots; ++t) offsets[t] = i;
    }
    prev = k[j];
    if (i == n_isects - 1)
      for (uint32_t t = k[j] + 1u; t < (uint32_t)n_slots; ++t) offsets[t] = n_isects;
  }
}

}  // namespace egs

using namespace egs;

extern "C" int64_t egs_exclusive_scan_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return (ceil_div(n > 0 ? n : 1, kScanTile) + 1) * (int64_t)sizeof(int64_t);
}

extern "C" int egs_exclusive_scan(int64_t n, const int32_t* in, int64_t* out, int64_t* total, void* workspace,
                                  int64_t workspace_bytes, egs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EGS_REQUIRE(n >= 0, "exclusive_scan: n=%lld < 0", (long long)n);
  if (n == 0) {
    EGS_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), stream));
    return 0;
  }
  if (workspace_bytes < egs_exclusive_scan_workspace_bytes(n))
    return fail(EGS_ERR_WORKSPACE_TOO_SMALL, "exclusive_scan: workspace %lld < %lld bytes", (long long)workspace_bytes,
                (long long)egs_exclusive_scan_workspace_bytes(n));
  const int64_t nblocks = ceil_div(n, kScanTile);
  EGS_REQUIRE(nblocks <= 0x7fffffff, "exclusive_scan: n too large");
  int64_t* block_sums = reinterpret_cast<int64_t*>(workspace);
  scan_block_sums_kernel<<<(unsigned)nblocks, kScanThreads, 0, stream>>>(in, n, block_sums);
  scan_spine_kernel<<<1, kScanThreads, 0, stream>>>(block_sums, nblocks, total);
  scan_apply_kernel<<<(unsigned)nblocks, kScanThreads, 0, stream>>>(in, n, block_sums, out);
  return check_launch("exclusive_scan", 3);
}

extern "C" int egs_isect_emit(int32_t C, int32_t N, const float* means2d, const int32_t* radii, const float* depths,
                              const int64_t* cum_tiles_excl, int32_t tile_size, int32_t tile_width,
                              int32_t tile_height, int32_t tile_n_bits, int64_t n_isects, int64_t* isect_ids,
                              int32_t* flatten_ids, egs_stream_t stream) {
  EGS_REQUIRE(C >= 0 && N >= 0, "isect_emit: negative sizes");
  EGS_REQUIRE((int64_t)C * N < 0x7fffffffLL, "isect_emit: C*N=%lld does not fit the int32 flatten id", (long long)C * N);
  EGS_REQUIRE(tile_n_bits >= 1 && tile_n_bits <= 30, "isect_emit: tile_n_bits=%d out of range", tile_n_bits);
  if ((int64_t)C * N == 0 || n_isects == 0) return 0;
  const int64_t nblocks = ceil_div((int64_t)C * N, kEmitThreads);
  isect_emit_kernel<<<(unsigned)nblocks, kEmitThreads, 0, (cudaStream_t)stream>>>(
      C, N, reinterpret_cast<const float2*>(means2d), radii, depths, cum_tiles_excl, (float)tile_size, tile_width,
      tile_height, tile_n_bits, n_isects, isect_ids, flatten_ids);
  return check_launch("isect_emit_kernel");
}

extern "C" int egs_isect_offset_encode(int64_t n_isects, const int64_t* isect_ids_sorted, int32_t C, int32_t n_tiles,
                                       int32_t tile_n_bits, int32_t* offsets, egs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EGS_REQUIRE(n_isects >= 0 && n_isects < 0x7fffffffLL, "isect_offset_encode: n_isects=%lld out of int32 range", (long long)n_isects);
  const int64_t n_slots = (int64_t)C * n_tiles;
  if (n_slots == 0) return 0;
  if (n_isects == 0) {
    EGS_CUDA(cudaMemsetAsync(offsets, 0, n_slots * sizeof(int32_t), stream));
    return 0;
  }
  isect_offset_encode_kernel<<<(unsigned)ceil_div(n_isects, kOffThreads), kOffThreads, 0, stream>>>(
      n_isects, isect_ids_sorted, n_tiles, tile_n_bits, n_slots, offsets);
  return check_launch("isect_offset_encode_kernel");
}

// ---- fast path entry points (see the block comment above isect_emit_sorted_kernel) -------------------------
extern "C" int64_t egs_isect_scan_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return 256 + ceil_div(n > 0 ? n : 1, kVisTile) * (int64_t)sizeof(uint64_t);  // control block + chain status words
}

extern "C" int egs_isect_visible_keys(int32_t C, int32_t N, const int32_t* tiles_per_gauss, const float* depths,
                                      uint32_t* keys1, uint32_t* vals1, int64_t* totals, void* workspace,
                                      int64_t workspace_bytes, egs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EGS_REQUIRE(C >= 0 && N >= 0, "isect_visible_keys: negative sizes");
  const int64_t n = (int64_t)C * N;
  EGS_REQUIRE(n < 0x7fffffffLL, "isect_visible_keys: C*N=%lld does not fit the int32 flatten id", (long long)n);
  if (n == 0) {
    EGS_CUDA(cudaMemsetAsync(totals, 0, 4 * sizeof(int64_t), stream));
    return 0;
  }
  const int64_t need = egs_isect_scan_workspace_bytes(n);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(EGS_ERR_WORKSPACE_TOO_SMALL, "isect_visible_keys: workspace too small");
  EGS_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "isect_visible_keys: workspace must be 16-byte aligned");
  EGS_CUDA(cudaMemsetAsync(workspace, 0, need, stream));
  const int64_t nblocks = ceil_div(n, kVisTile);
  char* ws = reinterpret_cast<char*>(workspace);
  visible_keys_kernel<<<(unsigned)nblocks, kVisThreads, 0, stream>>>(tiles_per_gauss, n, depths, keys1, vals1, totals,
                                                                      reinterpret_cast<VisibleControl*>(ws),
                                                                      reinterpret_cast<uint64_t*>(ws + 256));
  return check_launch("visible_keys_kernel");
}

extern "C" int egs_exclusive_scan_gather(int64_t n, const int32_t* src, const uint32_t* gather, int64_t* out,
                                         int64_t* total, void* workspace, int64_t workspace_bytes,
                                         egs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EGS_REQUIRE(n >= 0, "exclusive_scan_gather: n < 0");
  if (n == 0) {
    EGS_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), stream));
    return 0;
  }
  if (workspace_bytes < egs_exclusive_scan_workspace_bytes(n))
    return fail(EGS_ERR_WORKSPACE_TOO_SMALL, "exclusive_scan_gather: workspace too small");
  const int64_t nblocks = ceil_div(n, kScanTile);
  int64_t* block_sums = reinterpret_cast<int64_t*>(workspace);
  gscan_block_sums_kernel<<<(unsigned)nblocks, kScanThreads, 0, stream>>>(src, gather, n, nullptr, block_sums);
  scan_spine_kernel<<<1, kScanThreads, 0, stream>>>(block_sums, nblocks, total);
  gscan_apply_kernel<<<(unsigned)nblocks, kScanThreads, 0, stream>>>(src, gather, n, nullptr, block_sums, out);
  return check_launch("exclusive_scan_gather", 3);
}

extern "C" int egs_isect_emit_sorted(int32_t C, int32_t N, int64_t n_vis, const uint32_t* order,
                                     const int64_t* cum_excl, const float* means2d, const int32_t* radii,
                                     int32_t tile_size, int32_t tile_width, int32_t tile_height, int64_t n_isects,
                                     uint32_t* tile_keys, uint32_t* flat_vals, egs_stream_t stream) {
  EGS_REQUIRE(C >= 0 && N >= 0 && n_vis >= 0, "isect_emit_sorted: negative sizes");
  EGS_REQUIRE((int64_t)C * tile_width * tile_height < 0x7fffffffLL, "isect_emit_sorted: too many tiles");
  EGS_REQUIRE((int64_t)tile_width * tile_height * tile_width < 0x100000000LL,
              "isect_emit_sorted: tile grid %dx%d too large for the reciprocal row split", tile_width, tile_height);
  if (n_vis == 0 || n_isects == 0) return 0;
  isect_emit_sorted_kernel<<<(unsigned)ceil_div(n_vis, kEmitThreads), kEmitThreads, 0, (cudaStream_t)stream>>>(
      N, n_vis, order, cum_excl, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, tile_width,
      tile_height, n_isects, tile_keys, flat_vals, nullptr, nullptr);
  return check_launch("isect_emit_sorted_kernel");
}

extern "C" int egs_isect_finalize(int64_t n_isects, const uint32_t* tile_keys_sorted, const uint32_t* flat_sorted,
                                  const float* depths, int32_t C, int32_t n_tiles, int32_t tile_n_bits,
                                  int64_t* isect_ids, int32_t* offsets, egs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EGS_REQUIRE(n_isects >= 0 && n_isects < 0x7fffffffLL, "isect_finalize: n_isects=%lld out of int32 range", (long long)n_isects);
  const int64_t n_slots = (int64_t)C * n_tiles;
  if (n_slots == 0) return 0;
  if (n_isects == 0) {
    if (offsets != nullptr) EGS_CUDA(cudaMemsetAsync(offsets, 0, n_slots * sizeof(int32_t), stream));
    return 0;
  }
  if (isect_ids == nullptr && offsets != nullptr && reinterpret_cast<uintptr_t>(tile_keys_sorted) % 16 == 0) {
    isect_offsets4_kernel<<<(unsigned)ceil_div(n_isects, 4 * kOff4Threads), kOff4Threads, 0, stream>>>(
        (int32_t)n_isects, tile_keys_sorted, (int32_t)n_slots, offsets, nullptr, 0);
    return check_launch("isect_offsets4_kernel");
  }
  isect_finalize_kernel<<<(unsigned)ceil_div(n_isects, kOffThreads), kOffThreads, 0, stream>>>(
      n_isects, tile_keys_sorted, flat_sorted, depths, n_tiles, tile_n_bits, n_slots, isect_ids, offsets);
  return check_launch("isect_finalize_kernel");
}

// Launch order of the blend kernels: the tiles sorted by list length, longest first — by length CLASS
// (class = position of the highest set bit of the length, so a class spans a factor of two), tiles of one class staying
// roughly in grid order so that neighbouring tiles, which share Gaussians, still run close together.  The GPU starts
// thread blocks roughly in index order; in grid order the last blocks to start are the bottom rows of the last view,
// whatever their length, and a launch can end with a few SMs walking long lists while the rest idle.  Measured: one
// view per call of the object scene 1.117 -> 1.042 ms, batched benchmark step unchanged (gpurun_out/r2l).
// Two tiny launches (class histogram; bases + placement; warp-aggregated atomics) instead of one single-CTA kernel,
// which took 93 us for the 32 640 tiles of the benchmark step (profiles/r2n_launches.csv).
namespace egs {
constexpr int kSchedThreads = 256;
constexpr int kSchedClasses = 32;

__device__ __forceinline__ int length_class(int32_t len) { return len > 0 ? 32 - __clz(len) : 0; }  // 0 .. 31

// hist[c] = tiles of class c (caller-zeroed); *max_len = longest list
__global__ void __launch_bounds__(kSchedThreads) tile_classes_kernel(const int32_t* __restrict__ offsets, int32_t n_slots,
                                                                     int32_t* __restrict__ hist,
                                                                     unsigned long long* __restrict__ max_len) {
  const int32_t t = blockIdx.x * kSchedThreads + threadIdx.x;
  const bool valid = t < n_slots;
  const int32_t len = valid ? offsets[t + 1] - offsets[t] : 0;
  const int lane = threadIdx.x & 31;
  const uint32_t active = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const int cls = length_class(len);
    const uint32_t group = __match_any_sync(active, cls);
    if (lane == __ffs(group) - 1) atomicAdd(&hist[cls], __popc(group));
  }
  const int32_t m = __reduce_max_sync(0xffffffffu, len);
  if (lane == 0 && m > 0 && max_len != nullptr) atomicMax(max_len, (unsigned long long)m);
}

// Tile offsets from the first positions the last sort pass left (radix_sort.cu, onesweep_tile (h)), the sentinel, the
// class histogram of the list lengths and the longest list, in one launch: offsets[t] = the first position of the
// first non-empty tile at or behind t = a suffix minimum, chained from the LAST chunk of tiles to the first (ticket 0
// takes the last chunk).  Takes the place of a pass over all sorted keys (isect_offsets4_kernel: 74 MB read for the 18 M
// entries of the benchmark step) and of tile_classes_kernel.
constexpr int kFillItems = 8;
constexpr int kFillChunk = kSchedThreads * kFillItems;  // 2048 tiles per block

__global__ void __launch_bounds__(kSchedThreads) tile_offsets_fill_kernel(int32_t* __restrict__ offsets /* in: first positions */,
                                                                          int32_t n_slots, int64_t capacity,
                                                                          const int64_t* __restrict__ n_dev,
                                                                          int32_t* __restrict__ hist,
                                                                          unsigned long long* __restrict__ max_len,
                                                                          uint32_t* __restrict__ ticket, uint64_t* __restrict__ status) {
  __shared__ uint32_t warp_min[kSchedThreads / 32];
  __shared__ uint32_t carry_s;
  __shared__ uint32_t block_ticket;
  const uint32_t n_live = (uint32_t)live_count(capacity, n_dev);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) block_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t blk = block_ticket;
  const int32_t chunk = (int32_t)(gridDim.x - 1 - blk);  // chain order: from the last tiles to the first
  const int32_t base = chunk * kFillChunk + threadIdx.x * kFillItems;
  uint32_t v[kFillItems];
#pragma unroll
  for (int i = 0; i < kFillItems; ++i) v[i] = base + i < n_slots ? (uint32_t)offsets[base + i] : 0xffffffffu;
  // suffix minimum inside the thread, then across the threads behind it (higher thread index = later tiles)
#pragma unroll
  for (int i = kFillItems - 2; i >= 0; --i) v[i] = min(v[i], v[i + 1]);
  uint32_t suf = v[0];  // inclusive suffix minimum over this thread and the later threads of the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_down_sync(0xffffffffu, suf, d);
    if (lane + d < 32) suf = min(suf, t);
  }
  if (lane == 0) warp_min[warp] = suf;
  __syncthreads();
  if (warp == 0) {
    uint32_t block_min = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < kSchedThreads / 32; ++w) block_min = min(block_min, warp_min[w]);
    const uint32_t carry = chain_lookback_min(status, blk, block_min, lane);  // minimum over all later chunks
    if (lane == 0) carry_s = min(carry, n_live);  // behind the last non-empty tile: the end of the list
  }
  __syncthreads();
  uint32_t behind = carry_s;  // suffix minimum over everything behind this thread's tiles
#pragma unroll
  for (int w = kSchedThreads / 32 - 1; w >= 0; --w) behind = w > warp ? min(behind, warp_min[w]) : behind;
  const uint32_t next_lane = __shfl_down_sync(0xffffffffu, suf, 1);
  if (lane < 31) behind = min(behind, next_lane);
  int32_t longest = 0;
  uint32_t after = behind;  // offsets[t + 1] while walking the thread's tiles backwards
#pragma unroll
  for (int i = kFillItems - 1; i >= 0; --i) {
    const int32_t t = base + i;
    const bool valid = t < n_slots;
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    const uint32_t mine = min(v[i], behind);
    if (valid) {
      offsets[t] = (int32_t)mine;
      const int32_t len = (int32_t)(after - mine);
      longest = max(longest, len);
      const int cls = length_class(len);
      const uint32_t group = __match_any_sync(active, cls);
      if (lane == __ffs(group) - 1) atomicAdd(&hist[cls], __popc(group));
      after = mine;
    }
  }
  if (blk == 0 && threadIdx.x == 0) offsets[n_slots] = (int32_t)n_live;  // the sentinel the blend kernels read
  const int32_t m = __reduce_max_sync(0xffffffffu, longest);
  if (lane == 0 && m > 0 && max_len != nullptr) atomicMax(max_len, (unsigned long long)m);
}

// Placement: class c starts behind all longer classes (suffix sums of the class histogram, recomputed by every block —
// 32 values — rather than by a launch of their own); `taken` counts what each class has handed out (zero at launch).
__global__ void __launch_bounds__(kSchedThreads) tile_order_kernel(const int32_t* __restrict__ offsets, int32_t n_slots,
                                                                   const int32_t* __restrict__ hist, int32_t* __restrict__ taken,
                                                                   int32_t* __restrict__ order) {
  __shared__ int32_t class_base[kSchedClasses];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const int32_t n = hist[lane];
    int32_t after = n;  // inclusive suffix sum over classes >= lane
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t v = __shfl_down_sync(0xffffffffu, after, d);
      if (lane + d < 32) after += v;
    }
    class_base[lane] = after - n;  // tiles of strictly longer classes come first
  }
  __syncthreads();
  const int32_t t = blockIdx.x * kSchedThreads + threadIdx.x;
  const bool valid = t < n_slots;
  const uint32_t active = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const int cls = length_class(offsets[t + 1] - offsets[t]);
  const uint32_t group = __match_any_sync(active, cls);
  const int leader = __ffs(group) - 1;
  int32_t base = 0;
  if (lane == leader) base = class_base[cls] + atomicAdd(&taken[cls], __popc(group));
  base = __shfl_sync(group, base, leader);
  order[base + __popc(group & ((1u << lane) - 1u))] = t;
}
}  // namespace egs

// ---- the whole route behind egs_isect_visible_keys in ONE call, with the two counts read on the device -----------------
// Workspace layout (all 256-byte aligned): keys1_b | vals1_b | level-2 ping-pong keys | level-2 ping-pong values |
// control block = {level-1 sort workspace | emission chain (ticket + one status word per block) | level-2 sort
// workspace | tile-schedule histogram | offsets chain}.  The control block is what must start at zero: ONE memset per
// call (plus one of 0xFF over the tile offsets, which receive the keys' first positions from the last sort pass).
namespace {
struct SortedLayout {
  int64_t keys1_b, vals1_b, keys2, vals2, control, sort1, chain, sort2, sched, fill, total;
};
SortedLayout sorted_layout(int64_t n, int64_t capacity, int end_bit2, int64_t n_slots) {
  SortedLayout L;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { const int64_t at = off; off += egs::align_up(bytes > 0 ? bytes : 16, 256); return at; };
  L.keys1_b = take(n * 4);
  L.vals1_b = take(n * 4);
  L.keys2 = take(capacity * 4);
  L.vals2 = take(capacity * 4);
  L.control = off;
  L.sort1 = take(egs::radix_sort_workspace_bytes(n, 32));
  L.chain = take(256 + ceil_div(n, (int64_t)kEmitBlock) * 8);  // ticket (padded) + status words
  L.sort2 = take(egs::radix_sort_workspace_bytes(capacity, end_bit2));
  L.sched = take(2 * 32 * 4);  // class histogram + class cursors of the tile schedule
  L.fill = take(256 + ceil_div(n_slots > 0 ? n_slots : 1, (int64_t)egs::kFillChunk) * 8);  // offsets chain: ticket + status words
  L.total = off;
  return L;
}
int level2_end_bit(int64_t n_slots) {
  int b = 1;
  while (b < 31 && (1ll << b) < n_slots) ++b;
  return b;
}
}  // namespace

// offsets[n_slots + 1] (with the sentinel) -> *max_len (nullable), order[n_slots] (nullable); scratch: 64 int32
static int tile_schedule(const int32_t* offsets, int32_t n_slots, unsigned long long* max_len, int32_t* order, int32_t* scratch,
                         cudaStream_t stream, bool scratch_is_zero = false) {
  if (n_slots <= 0 || (max_len == nullptr && order == nullptr)) return 0;
  int32_t* hist = scratch;
  int32_t* taken = scratch + kSchedClasses;
  if (!scratch_is_zero) EGS_CUDA(cudaMemsetAsync(scratch, 0, 2 * kSchedClasses * sizeof(int32_t), stream));
  const unsigned blocks = (unsigned)ceil_div(n_slots, kSchedThreads);
  tile_classes_kernel<<<blocks, kSchedThreads, 0, stream>>>(offsets, n_slots, hist, max_len);
  if (order == nullptr) return check_launch("tile_classes_kernel");
  tile_order_kernel<<<blocks, kSchedThreads, 0, stream>>>(offsets, n_slots, hist, taken, order);
  return check_launch("tile_schedule", 2);
}

extern "C" int64_t egs_isect_sorted_workspace_bytes(int32_t C, int32_t N, int32_t n_tiles, int64_t capacity) {
  if (C < 0 || N < 0 || n_tiles < 0 || capacity < 0) return 0;
  const int64_t n = (int64_t)C * N;
  return sorted_layout(n > 0 ? n : 1, capacity > 0 ? capacity : 1, level2_end_bit((int64_t)C * n_tiles), (int64_t)C * n_tiles).total;
}

extern "C" int egs_isect_sorted(int32_t C, int32_t N, const int32_t* tight_rects, const float* means2d, const int32_t* radii,
                                uint32_t* keys1, uint32_t* vals1, int64_t* stats, int32_t tile_size, int32_t tile_width,
                                int32_t tile_height, int64_t capacity, void* workspace, int64_t workspace_bytes,
                                uint32_t* tile_keys, uint32_t* flatten_ids, int32_t* offsets, int32_t* tile_order,
                                egs_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EGS_REQUIRE(C >= 0 && N >= 0, "isect_sorted: negative sizes");
  const int64_t n = (int64_t)C * N;
  const int64_t n_slots = (int64_t)C * tile_width * tile_height;
  EGS_REQUIRE(n < 0x7fffffffLL, "isect_sorted: C*N=%lld does not fit the int32 flatten id", (long long)n);
  EGS_REQUIRE(n_slots < 0x7fffffffLL, "isect_sorted: too many tiles");
  EGS_REQUIRE(tile_width < 65536 && tile_height < 65536 && (int64_t)tile_width * tile_height < (1 << 21),
              "isect_sorted: tile grid %dx%d too large (packed rectangles, 32-bit block scan)", tile_width, tile_height);
  EGS_REQUIRE((int64_t)tile_width * tile_height * tile_width < 0x100000000LL,
              "isect_sorted: tile grid %dx%d too large for the reciprocal row split", tile_width, tile_height);
  EGS_REQUIRE(capacity >= 1 && capacity < 0x7fffffffLL, "isect_sorted: capacity=%lld out of int32 range", (long long)capacity);
  EGS_REQUIRE(stats != nullptr, "isect_sorted: the device counts {n_vis, n_isects, ..} of egs_isect_visible_keys are required");
  const int64_t* counts = stats;       // [0] n_vis (input), [1] intersection count (rewritten by the emission below)
  int64_t* n_isects_dev = stats + 1;
  stats += 2;                          // slot 2: longest tile list (output)
  if (n_slots == 0) return 0;
  if (n == 0) {  // nothing to bin: all offsets (and the sentinel) are zero, any launch order will do
    EGS_CUDA(cudaMemsetAsync(offsets, 0, (n_slots + 1) * sizeof(int32_t), stream));
    EGS_REQUIRE(tile_order == nullptr || (workspace != nullptr && workspace_bytes >= 256), "isect_sorted: workspace too small");
    return tile_schedule(offsets, (int32_t)n_slots, nullptr, tile_order, reinterpret_cast<int32_t*>(workspace), stream);
  }
  const int end_bit2 = level2_end_bit(n_slots);
  const SortedLayout L = sorted_layout(n, capacity, end_bit2, n_slots);
  if (workspace == nullptr || workspace_bytes < L.total)
    return fail(EGS_ERR_WORKSPACE_TOO_SMALL, "isect_sorted: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)L.total);
  char* ws = reinterpret_cast<char*>(workspace);
  uint32_t* keys1_b = reinterpret_cast<uint32_t*>(ws + L.keys1_b);
  uint32_t* vals1_b = reinterpret_cast<uint32_t*>(ws + L.vals1_b);
  EGS_CUDA(cudaMemsetAsync(ws + L.control, 0, L.total - L.control, stream));  // histograms, look-back words, tickets
  // level 1: visible entries of all cameras in (depth, flat index) order — 4 passes over n_vis 8-byte pairs
  int in_b = 0;
  if (int rc = radix_sort_pairs_u32(n, counts, keys1, vals1, keys1_b, vals1_b, 32, ws + L.sort1, L.chain - L.sort1, &in_b, stream, true))
    return rc;
  const uint32_t* order = in_b ? vals1_b : vals1;
  // scan of the tile counts in that order + emission, one launch; into whichever side of the level-2 ping-pong makes
  // the sorted pairs end in the caller's buffers.  (tight_rects: the projection kernel's tight rectangles for the blend
  // kernels' own lists; null: gsplat's rectangles.  The emitted total becomes stats[1], the count every kernel below
  // works with.)
  const int passes2 = (end_bit2 + 7) / 8;
  uint32_t* ka = (passes2 & 1) ? reinterpret_cast<uint32_t*>(ws + L.keys2) : tile_keys;
  uint32_t* va = (passes2 & 1) ? reinterpret_cast<uint32_t*>(ws + L.vals2) : flatten_ids;
  uint32_t* kb = (passes2 & 1) ? tile_keys : reinterpret_cast<uint32_t*>(ws + L.keys2);
  uint32_t* vb = (passes2 & 1) ? flatten_ids : reinterpret_cast<uint32_t*>(ws + L.vals2);
  isect_scan_emit_kernel<<<(unsigned)ceil_div(n, kEmitBlock), kEmitThreads, 0, stream>>>(
      N, n, counts, order, reinterpret_cast<const int2*>(tight_rects), reinterpret_cast<const float2*>(means2d), radii,
      (float)tile_size, tile_width, tile_height, capacity, ka, va, n_isects_dev, reinterpret_cast<uint32_t*>(ws + L.chain),
      reinterpret_cast<uint64_t*>(ws + L.chain + 256));
  if (int rc = check_launch("isect_scan_emit_kernel")) return rc;
  // level 2: stable sort on the dense (camera, tile) index.  Its last pass leaves every occurring key's first position
  // in `offsets` (0xffffffff = "no entry"), which one small launch turns into the tile offsets, the sentinel, the class
  // histogram of the list lengths and the longest list.
  EGS_CUDA(cudaMemsetAsync(offsets, 0xFF, (n_slots + 1) * sizeof(int32_t), stream));
  if (int rc = radix_sort_pairs_u32(capacity, n_isects_dev, ka, va, kb, vb, end_bit2, ws + L.sort2, L.sched - L.sort2, &in_b, stream, true,
                                    reinterpret_cast<uint32_t*>(offsets)))
    return rc;
  if ((in_b != 0) != ((passes2 & 1) != 0)) return fail(EGS_ERR_INVALID_ARGUMENT, "isect_sorted: internal ping-pong mismatch");
  int32_t* sched = reinterpret_cast<int32_t*>(ws + L.sched);  // class histogram | handed-out counters (zero: control block)
  const unsigned fill_blocks = (unsigned)ceil_div(n_slots, (int64_t)kFillChunk);
  tile_offsets_fill_kernel<<<fill_blocks, kSchedThreads, 0, stream>>>(
      offsets, (int32_t)n_slots, capacity, n_isects_dev, sched, reinterpret_cast<unsigned long long*>(stats),
      reinterpret_cast<uint32_t*>(ws + L.fill), reinterpret_cast<uint64_t*>(ws + L.fill + 256));
  if (tile_order == nullptr) return check_launch("tile_offsets_fill_kernel");
  tile_order_kernel<<<(unsigned)ceil_div(n_slots, (int64_t)kSchedThreads), kSchedThreads, 0, stream>>>(
      offsets, (int32_t)n_slots, sched, sched + kSchedClasses, tile_order);
  return check_launch("tile offsets + schedule", 2);
}
