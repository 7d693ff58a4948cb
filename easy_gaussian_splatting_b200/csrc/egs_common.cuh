// Shared host-side helpers for the C-ABI translation units: error reporting, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/egs_raster.h"

namespace egs {

// thread-local error text behind egs_last_error_string()
char* error_buffer();
int fail(int code, const char* fmt, ...);

// Adds to the process-wide count of kernels this library has launched (egs_kernel_launch_count()).
void note_kernel_launches(int n);

// n_kernels = how many kernels the caller enqueued since its last check (counted only when they were accepted).
inline int check_launch(const char* what, int n_kernels = 1) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  note_kernel_launches(n_kernels);
  return 0;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Element counts that only the device knows (the sync-free binning route): a kernel is launched for a CAPACITY and
// reads the live count itself.  count_dev == nullptr: the capacity is the count.
__device__ __forceinline__ int64_t live_count(int64_t capacity, const int64_t* __restrict__ count_dev) {
  if (count_dev == nullptr) return capacity;
  const int64_t n = *count_dev;
  return n < capacity ? n : capacity;
}

// radix_sort.cu: the onesweep sort of (u32 key, u32 value) pairs on key bits [0, end_bit) for up to `capacity` pairs,
// min(capacity, *count_dev) of which are live.  *result_in_b (host) = the sorted data ended in the b buffers.
// workspace_is_zero: the caller has already cleared the workspace (one memset for several stages).
// first_pos (nullable, [number of distinct keys], filled with 0xffffffff by the caller): receives the position of the
// first pair of every key that occurs — only meaningful when end_bit covers every set key bit (a complete sort).
int64_t radix_sort_workspace_bytes(int64_t capacity, int end_bit);
int radix_sort_pairs_u32(int64_t capacity, const int64_t* count_dev, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b,
                         uint32_t* vals_b, int end_bit, void* workspace, int64_t workspace_bytes, int* result_in_b,
                         cudaStream_t stream, bool workspace_is_zero = false, uint32_t* first_pos = nullptr);
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

}  // namespace egs

#define EGS_REQUIRE(cond, ...)                                         \
  do {                                                                 \
    if (!(cond)) return egs::fail(EGS_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

#define EGS_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) return egs::fail((int)e__, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)
