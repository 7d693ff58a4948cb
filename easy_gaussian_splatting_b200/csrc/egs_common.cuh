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
