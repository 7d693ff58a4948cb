// g4: hand-written onesweep radix sort of (u64 key, u32 value) pairs — stable, LSD, 8-bit digits,
// decoupled look-back (takes the place of cub::DeviceRadixSort::SortPairs in gsplat's isect_tiles).
//
// Launch sequence for P = ceil(end_bit / 8) passes over n pairs:
//   1 memset   : histograms, dynamic tile counters and look-back status words
//   1 histogram: reads the keys once (8 B/pair) and builds all P digit histograms; the block that finishes last
//                turns each 256-bin histogram into exclusive global digit bases
//   P passes   : each reads 12 B/pair and writes 12 B/pair; a tile of 4096 pairs is ranked in
//                shared memory (warp-level shared-memory atomicOr multisplit, stable), its per-digit counts are
//                chained to the preceding tiles with decoupled look-back, and the tile is written
//                out digit-run by digit-run so stores are coalesced.
// HBM traffic: 8 + 24*P bytes per pair = 152 B at P = 6.  Integer work only.
#include "egs_common.cuh"

namespace egs {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 8;
constexpr int kSortTilePairs = 4096;  // pairs per tile in every configuration: threads x ITEMS
// threads per block = kSortTilePairs / ITEMS (template parameter ITEMS = pairs per thread):
//   64-bit keys: ITEMS = 8, 512 threads (16 keys per thread do not fit the registers of two resident blocks)
//   32-bit keys: EGS_SORT_ITEMS_U32 (build-time A/B knob)

#ifndef EGS_SORT_ITEMS_U32
#define EGS_SORT_ITEMS_U32 16
#endif
#ifndef EGS_SORT_BLOCKS_16  // resident blocks per SM asked of ptxas for the 256-thread, 16-item form
#define EGS_SORT_BLOCKS_16 4
#endif

#ifndef EGS_SORT_LOOK_WINDOW
#define EGS_SORT_LOOK_WINDOW 4
#endif
constexpr int kLookWindow = EGS_SORT_LOOK_WINDOW;
constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;

struct SortWorkspace {
  uint32_t* hist;      // [kMaxPasses][256]  digit counts, then exclusive digit bases
  uint32_t* counters;  // [kMaxPasses]       dynamic tile ids
  uint32_t* status;    // [passes][ntiles][256]
};

__host__ __device__ inline int64_t sort_ntiles(int64_t n) { return (n + kSortTilePairs - 1) / kSortTilePairs; }

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- histogram of every digit position in one read of the keys -----------------------------------
constexpr int kHistThreads = 256;
constexpr int kHistItems = 16;

template <typename KeyT>
__global__ void __launch_bounds__(kHistThreads) radix_histogram_kernel(const KeyT* __restrict__ keys, int64_t n,
                                                                        const int64_t* __restrict__ n_dev, int passes,
                                                                        uint32_t* __restrict__ hist,
                                                                        uint32_t* __restrict__ blocks_done) {
  n = live_count(n, n_dev);
  // Each thread walks kHistItems CONSECUTIVE keys and combines runs of equal digits in a register before
  // touching shared memory.  The keys of this path come out of isect_emit in runs that share the depth
  // word and most of the tile id (one Gaussian = one run), so almost every shared-memory atomic
  // is saved and the hot-bin serialisation of a plain atomic histogram disappears.
  __shared__ uint32_t sh[kMaxPasses * kRadix];
  for (int i = threadIdx.x; i < passes * kRadix; i += kHistThreads) sh[i] = 0;
  __syncthreads();
  constexpr int kVec = 16 / sizeof(KeyT);  // keys per 128-bit load
  const int64_t chunk = (int64_t)kHistThreads * kHistItems;
  for (int64_t base = (int64_t)blockIdx.x * chunk; base < n; base += (int64_t)gridDim.x * chunk) {
    const int64_t first = base + (int64_t)threadIdx.x * kHistItems;
    KeyT k[kHistItems];
    int cnt = 0;
    if (first + kHistItems <= n) {
      const uint4* p4 = reinterpret_cast<const uint4*>(keys + first);  // first is a multiple of 16 keys: 16B aligned
#pragma unroll
      for (int i = 0; i < kHistItems / kVec; ++i) {
        const uint4 v = __ldg(p4 + i);
        if constexpr (sizeof(KeyT) == 8) {
          k[2 * i] = (KeyT)(((uint64_t)v.y << 32) | v.x);
          k[2 * i + 1] = (KeyT)(((uint64_t)v.w << 32) | v.z);
        } else {
          k[4 * i] = (KeyT)v.x; k[4 * i + 1] = (KeyT)v.y; k[4 * i + 2] = (KeyT)v.z; k[4 * i + 3] = (KeyT)v.w;
        }
      }
      cnt = kHistItems;
    } else {
#pragma unroll
      for (int i = 0; i < kHistItems; ++i) {
        if (first + i < n) { k[i] = keys[first + i]; cnt = i + 1; }
        else k[i] = 0;
      }
    }
    if (cnt == 0) continue;
    for (int p = 0; p < passes; ++p) {
      const int shift = p * kRadixBits;
      uint32_t run_d = (uint32_t)(k[0] >> shift) & (kRadix - 1);
      uint32_t run_n = 1;
#pragma unroll
      for (int i = 1; i < kHistItems; ++i) {
        if (i < cnt) {
          const uint32_t d = (uint32_t)(k[i] >> shift) & (kRadix - 1);
          if (d == run_d) {
            ++run_n;
          } else {
            atomicAdd(&sh[p * kRadix + run_d], run_n);
            run_d = d;
            run_n = 1;
          }
        }
      }
      atomicAdd(&sh[p * kRadix + run_d], run_n);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kRadix; i += kHistThreads) {
    const uint32_t v = sh[i];
    if (v) atomicAdd(&hist[i], v);
  }
  // The block that finishes last turns the counts into exclusive digit bases, in place (kHistThreads = 256 = one
  // thread per digit): the scan needs no launch of its own.
  __shared__ bool last_block;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_block = atomicAdd(blocks_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last_block) return;
  __threadfence();
  __shared__ uint32_t warp_tot[kRadix / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = 0; p < passes; ++p) {
    volatile uint32_t* h = hist + (size_t)p * kRadix;
    const uint32_t v = h[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    h[threadIdx.x] = base + inc - v;
    __syncthreads();
  }
}

// ---- one onesweep pass ------------------------------------------------------------------------------
// Dynamic shared memory layout of one tile (67.6 KB with 64-bit keys -> 2 resident CTAs of 16 warps per SM):
template <typename KeyT, int ITEMS>
struct SortSmem {
  static constexpr int kSortThreads = kSortTilePairs / ITEMS, kSortWarps = kSortThreads / 32;
  static constexpr int kSortTile = kSortTilePairs;
  static_assert(kSortThreads >= kRadix && kSortThreads % 32 == 0, "one thread per digit owns its look-back");
  KeyT keys[kSortTile];                     // 32 KB (u64) / 16 KB (u32)  tile-sorted keys
  uint32_t vals[kSortTile];                 // 16 KB  tile-sorted values
  uint32_t warp_hist[kSortWarps][kRadix];   // 16 KB  per-warp digit counts, then per-warp exclusive offsets
  uint32_t match[kSortWarps][kRadix];       // 16 KB  per-warp, per-digit lane masks (zero between items)
  uint32_t digit_start[kRadix];             // first slot of each digit inside the tile
  uint32_t dst_base[kRadix];                // global base - digit_start (mod 2^32)
  uint32_t warp_tot[kRadix / 32];
  uint32_t tile;
};

// Register budget: the 16 keys (32 registers) are only live until they are scattered into shared memory;
// ranks / slots are packed two per register; the values are loaded after the keys have left, and the
// global destination of a slot is recomputed from the key's digit instead of being kept.  That keeps
// the kernel at 64 registers (2 CTAs of 512 threads per SM) with every global load of a phase in flight at once.
// One tile of a pass (FULL: all kSortTile slots are live — every tile but the last; the ragged form clamps its loads
// and predicates every shared-memory and global store, ~15 % more instructions).
template <typename KeyT, int ITEMS, bool FULL, bool FIRST_POS>
__device__ __forceinline__ void onesweep_tile(SortSmem<KeyT, ITEMS>& sm, const KeyT* __restrict__ keys_in,
                                              const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                                              uint32_t* __restrict__ vals_out, int shift,
                                              const uint32_t* __restrict__ digit_base, uint32_t* __restrict__ status,
                                              uint32_t tile, int64_t tile_base, int tile_count,
                                              uint32_t* __restrict__ first_pos) {
  constexpr int kSortItems = ITEMS, kSortTile = kSortTilePairs, kWarpSpan = 32 * ITEMS;
  constexpr int kSortThreads = kSortTilePairs / ITEMS, kSortWarps = kSortThreads / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (a) load, warp-striped: item i of lane l in warp w is element w*256 + i*32 + l of the tile
  // (loads are unpredicated — out-of-range items re-read the tile's last element and are masked by `valid`
  //  where it matters — because per-load predicates exhaust the 7 predicate registers and make ptxas
  //  serialise the loads behind the ranking loop: profiles/r1c)
  KeyT key[kSortItems];
  const int warp_off = warp * kWarpSpan + lane;
  const int last_local = tile_count - 1;
  if constexpr (FULL) {  // every tile but the last: plain strided loads off one pointer
    const KeyT* kp = keys_in + tile_base + warp_off;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) key[i] = kp[i * 32];
  } else {
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) key[i] = keys_in[tile_base + min(warp_off + i * 32, last_local)];
  }

  // (b) stable rank of every item among the items of its warp with the same digit (two ranks per register)
  uint32_t packed[kSortItems / 2];
  uint32_t* wh = sm.warp_hist[warp];
  const uint32_t lanemask_lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const bool valid = FULL || (warp_off + i * 32) < tile_count;
    const uint32_t d = (uint32_t)(key[i] >> shift) & (kRadix - 1);
    // lanes holding the same digit: each lane ORs its bit into a per-warp, per-digit mask word in shared
    // memory (ATOMS.OR), then reads the word back.  ~12 instructions per item; 8 ballots cost ~48 (ALU bound,
    // ncu r1g) and match.any stalls 10-14 warps per issue on high-entropy digits (ncu r1e).  The group leader
    // clears the word again, so the masks never need a bulk reset.
    uint32_t* mm = sm.match[warp];
    if (valid) atomicOr(&mm[d], 1u << lane);
    __syncwarp();
    const uint32_t m = valid ? mm[d] : 0u;
    const uint32_t before = __popc(m & lanemask_lt);
    const uint32_t prev = wh[d];
    __syncwarp();
    if (valid && before == 0) {
      wh[d] = prev + (uint32_t)__popc(m);
      mm[d] = 0u;
    }
    __syncwarp();
    const uint32_t rank = prev + before;
    if (i & 1) packed[i / 2] |= rank << 16;
    else packed[i / 2] = rank;
  }
  __syncthreads();

  // (c) thread d (< 256) owns digit d: exclusive prefix over the warps, tile total
  const bool owner = tid < kRadix;
  uint32_t count = 0;
  uint32_t* my_status = status + (size_t)tile * kRadix + (owner ? tid : 0);
  if (owner) {
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t t = sm.warp_hist[w][tid];
      sm.warp_hist[w][tid] = count;
      count += t;
    }
    // publish the tile aggregate (or the inclusive prefix for tile 0) as early as possible
    st_volatile_u32(my_status, (tile == 0 ? kFlagPrefix : kFlagAggregate) | count);
  }

  // (d) exclusive scan of the digit counts across the 256 owner threads -> slot of each digit in the tile
  uint32_t inc = count;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (owner && lane == 31) sm.warp_tot[warp] = inc;
  __syncthreads();
  uint32_t digit_start = 0;
  if (owner) {
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w) wbase += (w < warp) ? sm.warp_tot[w] : 0u;
    digit_start = wbase + inc - count;
    sm.digit_start[tid] = digit_start;
  }
  __syncthreads();

  // (e) scatter the keys into tile-sorted order in shared memory; keep the slots (packed)
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & (kRadix - 1);
    const uint32_t rank = (i & 1) ? (packed[i / 2] >> 16) : (packed[i / 2] & 0xffffu);
    const uint32_t pos = sm.digit_start[d] + wh[d] + rank;
    if (FULL || (warp_off + i * 32) < tile_count) sm.keys[pos] = key[i];
    if (i & 1) packed[i / 2] = (packed[i / 2] & 0xffffu) | (pos << 16);
    else packed[i / 2] = (packed[i / 2] & 0xffff0000u) | pos;
  }

  // the values can start travelling now (their registers are the ones the keys just released)
  uint32_t val[kSortItems];
  if constexpr (FULL) {
    const uint32_t* vp = vals_in + tile_base + warp_off;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) val[i] = vp[i * 32];
  } else {
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) val[i] = vals_in[tile_base + min(warp_off + i * 32, last_local)];
  }

  // (f) decoupled look-back for digit tid.  kLookWindow predecessors are polled with independent loads per
  // round (one L2 latency per round instead of one per predecessor), then consumed in order.
  uint32_t excl = 0;
  if (owner && tile > 0) {
    int64_t j = (int64_t)tile - 1;
    bool finished = false;
    while (!finished) {
      uint32_t v[kLookWindow];
#pragma unroll
      for (int w = 0; w < kLookWindow; ++w) {
        const int64_t jj = j - w;
        const uint32_t x = ld_volatile_u32(status + (size_t)(jj > 0 ? jj : 0) * kRadix + tid);
        v[w] = (jj >= 0) ? x : kFlagPrefix;  // before tile 0: an empty prefix
      }
#pragma unroll
      for (int w = 0; w < kLookWindow; ++w) {
        if (!finished) {
          if ((v[w] >> 30) == 0u) break;  // predecessor has not published yet: poll again from it
          excl += v[w] & kValueMask;
          --j;
          if ((v[w] & kFlagPrefix) != 0u) finished = true;
        }
      }
    }
    st_volatile_u32(my_status, kFlagPrefix | ((excl + count) & kValueMask));
  }
  if (owner) sm.dst_base[tid] = digit_base[tid] + excl - digit_start;

  // (g) values into tile-sorted order
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t pos = (i & 1) ? (packed[i / 2] >> 16) : (packed[i / 2] & 0xffffu);
    if (FULL || (warp_off + i * 32) < tile_count) sm.vals[pos] = val[i];
  }
  __syncthreads();

  // (h) write out; consecutive slots of one digit go to consecutive addresses
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int p = tid + i * kSortThreads;
    if (FULL || p < tile_count) {
      const KeyT k = sm.keys[p];
      const uint32_t d = (uint32_t)(k >> shift) & (kRadix - 1);
      const uint32_t dst = sm.dst_base[d] + (uint32_t)p;
      keys_out[dst] = k;
      vals_out[dst] = sm.vals[p];
      if constexpr (sizeof(KeyT) == 4 && FIRST_POS) {
        // LAST pass of a complete sort (first_pos given): the tile is then sorted on the whole key (its input was
        // sorted on the lower bits, the ranking is stable), so a key's first slot in the tile is a key boundary.  A
        // boundary INSIDE a digit run is the key's first position in the whole output (an earlier tile cannot hold the
        // key behind a smaller one of the same digit): plain store.  The first slot of a digit run may continue a key
        // of an earlier tile: atomicMin.  first_pos starts at 0xffffffff; keys without entries keep it.
        const KeyT kp = sm.keys[p > 0 ? p - 1 : 0];
        const bool run_start = p == 0 || ((uint32_t)(kp >> shift) & (kRadix - 1)) != d;
        if (run_start) atomicMin(first_pos + k, dst);
        else if (kp != k) first_pos[k] = dst;
      }
    }
  }
}

template <typename KeyT, int ITEMS, bool FIRST_POS>
__global__ void __launch_bounds__(kSortTilePairs / ITEMS, (ITEMS > 8) ? EGS_SORT_BLOCKS_16 : 2) radix_onesweep_kernel(
    const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, int64_t n, const int64_t* __restrict__ n_dev, int shift,
    const uint32_t* __restrict__ digit_base /*[256]*/, uint32_t* __restrict__ tile_counter,
    uint32_t* __restrict__ status /*[ntiles][256]*/, uint32_t* __restrict__ first_pos /* nullable, see onesweep_tile (h) */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  n = live_count(n, n_dev);  // the grid covers the capacity; tiles past the live count leave at once (below)
  constexpr int kSortItems = ITEMS, kSortTile = kSortTilePairs, kWarpSpan = 32 * ITEMS;
  constexpr int kSortThreads = kSortTilePairs / ITEMS, kSortWarps = kSortThreads / 32;
  SortSmem<KeyT, ITEMS>& sm = *reinterpret_cast<SortSmem<KeyT, ITEMS>*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // The grid covers the capacity.  Blocks past the live tiles leave before they take a ticket, so exactly the live
  // tiles' worth of blocks draw tickets 0 .. live-1 (a buffer sized generously costs a load per spare block, not an atomic).
  if ((int64_t)blockIdx.x * kSortTile >= n) return;
  // dynamic tile id: a tile only waits on tiles that have already started (no look-back deadlock)
  if (tid == 0) sm.tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < 2 * kSortWarps * kRadix; i += kSortThreads) (&sm.warp_hist[0][0])[i] = 0;  // warp_hist + match
  __syncthreads();
  const uint32_t tile = sm.tile;
  const int64_t tile_base = (int64_t)tile * kSortTile;
  const int tile_count = (int)min((int64_t)kSortTile, n - tile_base);

  if (tile_count == kSortTile)
    onesweep_tile<KeyT, ITEMS, true, FIRST_POS>(sm, keys_in, vals_in, keys_out, vals_out, shift, digit_base, status, tile, tile_base, tile_count, first_pos);
  else
    onesweep_tile<KeyT, ITEMS, false, FIRST_POS>(sm, keys_in, vals_in, keys_out, vals_out, shift, digit_base, status, tile, tile_base, tile_count, first_pos);
}

static int carve_workspace(void* ws, int64_t ws_bytes, int64_t n, int passes, SortWorkspace& w, int64_t& clear_bytes) {
  const int64_t ntiles = sort_ntiles(n);
  const int64_t hist_b = (int64_t)kMaxPasses * kRadix * 4;
  const int64_t cnt_b = 256;  // kMaxPasses counters, padded
  const int64_t status_b = (int64_t)passes * ntiles * kRadix * 4;
  clear_bytes = hist_b + cnt_b + status_b;
  if (ws == nullptr) return 0;
  if (ws_bytes < clear_bytes) return -1;
  char* base = reinterpret_cast<char*>(ws);
  w.hist = reinterpret_cast<uint32_t*>(base);
  w.counters = reinterpret_cast<uint32_t*>(base + hist_b);
  w.status = reinterpret_cast<uint32_t*>(base + hist_b + cnt_b);
  return 0;
}

}  // namespace egs

using namespace egs;

extern "C" int64_t egs_radix_sort_workspace_bytes(int64_t n, int32_t end_bit) {
  if (n < 0 || end_bit < 0) return 0;
  int passes = (end_bit + kRadixBits - 1) / kRadixBits;
  if (passes > kMaxPasses) passes = kMaxPasses;
  SortWorkspace w;
  int64_t bytes = 0;
  carve_workspace(nullptr, 0, n, passes, w, bytes);
  return bytes;
}

template <typename KeyT, int ITEMS>
static int run_passes(int64_t n, const int64_t* n_dev, KeyT* keys_a, uint32_t* vals_a, KeyT* keys_b, uint32_t* vals_b,
                      int passes, const SortWorkspace& w, cudaStream_t stream, uint32_t* first_pos);

template <typename KeyT>
static int radix_sort_pairs_impl(int64_t n, const int64_t* n_dev, KeyT* keys_a, uint32_t* vals_a, KeyT* keys_b,
                                 uint32_t* vals_b, int32_t end_bit, void* workspace, int64_t workspace_bytes,
                                 int32_t* host_result_in_b, cudaStream_t stream, bool workspace_is_zero = false,
                                 uint32_t* first_pos = nullptr) {
  constexpr int kKeyBits = (int)sizeof(KeyT) * 8;
  EGS_REQUIRE(n >= 0, "radix_sort: n=%lld < 0", (long long)n);
  EGS_REQUIRE(n < (1ll << 30), "radix_sort: n=%lld exceeds the 2^30 pairs the look-back words can count", (long long)n);
  EGS_REQUIRE(end_bit >= 0 && end_bit <= kKeyBits, "radix_sort: end_bit=%d out of [0,%d]", end_bit, kKeyBits);
  EGS_REQUIRE(reinterpret_cast<uintptr_t>(keys_a) % 16 == 0 && reinterpret_cast<uintptr_t>(keys_b) % 16 == 0,
              "radix_sort: key buffers must be 16-byte aligned");
  const int passes = (end_bit + kRadixBits - 1) / kRadixBits;
  if (host_result_in_b) *host_result_in_b = (passes & 1);
  if (n == 0 || passes == 0) {
    if (host_result_in_b) *host_result_in_b = 0;
    return 0;
  }
  SortWorkspace w;
  int64_t clear_bytes = 0;
  if (carve_workspace(workspace, workspace_bytes, n, passes, w, clear_bytes) != 0 || workspace == nullptr)
    return fail(EGS_ERR_WORKSPACE_TOO_SMALL, "radix_sort: workspace %lld < %lld bytes", (long long)workspace_bytes,
                (long long)clear_bytes);
  if (!workspace_is_zero) EGS_CUDA(cudaMemsetAsync(workspace, 0, clear_bytes, stream));
  int64_t hist_blocks = ceil_div(n, (int64_t)kHistThreads * kHistItems);
  if (hist_blocks > 148 * 8) hist_blocks = 148 * 8;
  static_assert(kHistThreads == kRadix, "the last histogram block scans one digit per thread");
  radix_histogram_kernel<KeyT><<<(unsigned)hist_blocks, kHistThreads, 0, stream>>>(keys_a, n, n_dev, passes, w.hist,
                                                                                    w.counters + kMaxPasses);
  // pairs per thread: 8 for 64-bit keys (a 16-item tile would need 2 x the shared memory and drop to one CTA per SM);
  // EGS_SORT_ITEMS_U32 for 32-bit keys (build-time A/B knob, scripts/build_variant.py)
  if constexpr (sizeof(KeyT) == 4) return run_passes<KeyT, EGS_SORT_ITEMS_U32>(n, n_dev, keys_a, vals_a, keys_b, vals_b, passes, w, stream, first_pos);
  else return run_passes<KeyT, 8>(n, n_dev, keys_a, vals_a, keys_b, vals_b, passes, w, stream, nullptr);
}

template <typename KeyT, int ITEMS>
static int run_passes(int64_t n, const int64_t* n_dev, KeyT* keys_a, uint32_t* vals_a, KeyT* keys_b, uint32_t* vals_b,
                      int passes, const SortWorkspace& w, cudaStream_t stream, uint32_t* first_pos) {
  constexpr int kSmem = (int)sizeof(SortSmem<KeyT, ITEMS>);
  // the variant that also records every key's first position (last pass of the level-2 sort) is its own kernel:
  // the plain passes keep their register budget
  auto plain = radix_onesweep_kernel<KeyT, ITEMS, false>;
  auto with_first_pos = radix_onesweep_kernel<KeyT, ITEMS, sizeof(KeyT) == 4>;
  for (auto kernel : {plain, with_first_pos}) {
    const cudaError_t attr_rc =  // per-device attribute: set on every call (cheap), not once per process
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (attr_rc != cudaSuccess)
      return fail((int)attr_rc, "radix_sort: cannot opt in to %d bytes of shared memory: %s", kSmem, cudaGetErrorString(attr_rc));
  }
  const int64_t ntiles = sort_ntiles(n);
  const int64_t status_stride = ntiles * kRadix;
  KeyT* kin = keys_a; uint32_t* vin = vals_a;
  KeyT* kout = keys_b; uint32_t* vout = vals_b;
  for (int p = 0; p < passes; ++p) {
    auto kernel = (p == passes - 1 && first_pos != nullptr) ? with_first_pos : plain;
    kernel<<<(unsigned)ntiles, kSortTilePairs / ITEMS, kSmem, stream>>>(
        kin, vin, kout, vout, n, n_dev, p * kRadixBits, w.hist + (size_t)p * kRadix, w.counters + p,
        w.status + (size_t)p * status_stride, first_pos);
    KeyT* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  return check_launch("radix_sort_pairs", passes + 1);  // + histogram (its last block scans)
}

namespace egs {
int64_t radix_sort_workspace_bytes(int64_t capacity, int end_bit) { return egs_radix_sort_workspace_bytes(capacity, end_bit); }
int radix_sort_pairs_u32(int64_t capacity, const int64_t* count_dev, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b,
                         uint32_t* vals_b, int end_bit, void* workspace, int64_t workspace_bytes, int* result_in_b,
                         cudaStream_t stream, bool workspace_is_zero, uint32_t* first_pos) {
  return radix_sort_pairs_impl<uint32_t>(capacity, count_dev, keys_a, vals_a, keys_b, vals_b, end_bit, workspace,
                                         workspace_bytes, result_in_b, stream, workspace_is_zero, first_pos);
}
}  // namespace egs

extern "C" int egs_radix_sort_pairs_u64_u32(int64_t n, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b,
                                            uint32_t* vals_b, int32_t end_bit, void* workspace,
                                            int64_t workspace_bytes, int32_t* host_result_in_b,
                                            egs_stream_t stream) {
  return radix_sort_pairs_impl<uint64_t>(n, nullptr, keys_a, vals_a, keys_b, vals_b, end_bit, workspace, workspace_bytes,
                                         host_result_in_b, (cudaStream_t)stream);
}

extern "C" int egs_radix_sort_pairs_u32_u32(int64_t n, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b,
                                            uint32_t* vals_b, int32_t end_bit, void* workspace,
                                            int64_t workspace_bytes, int32_t* host_result_in_b,
                                            egs_stream_t stream) {
  return radix_sort_pairs_impl<uint32_t>(n, nullptr, keys_a, vals_a, keys_b, vals_b, end_bit, workspace, workspace_bytes,
                                         host_result_in_b, (cudaStream_t)stream);
}
