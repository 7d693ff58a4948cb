// Micro-probe: issue cost of packed fp32 (FFMA2 / FADD2 / FMUL2, sm_100) against scalar FFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 4096;
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float seed) {
  float2 a[8];
  const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(seed, -seed);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  unsigned k = threadIdx.x;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // 2 scalar FFMA per pair
        a[i].x = fmaf(a[i].x, m.x, c.x);
        a[i].y = fmaf(a[i].y, m.y, c.y);
      } else if (MODE == 1) {  // 1 FFMA2 per pair
        a[i] = __ffma2_rn(a[i], m, c);
      } else if (MODE == 2) {  // scalar + an integer op per pair
        a[i].x = fmaf(a[i].x, m.x, c.x);
        a[i].y = fmaf(a[i].y, m.y, c.y);
        k = k * 1664525u + 1013904223u;
      } else if (MODE == 3) {  // FFMA2 + an integer op per pair
        a[i] = __ffma2_rn(a[i], m, c);
        k = k * 1664525u + 1013904223u;
      } else if (MODE == 4) {  // FADD2 + FMUL2
        a[i] = __fadd2_rn(__fmul2_rn(a[i], m), c);
      } else if (MODE == 5) {  // scalar FADD+FMUL
        a[i].x = __fadd_rn(__fmul_rn(a[i].x, m.x), c.x);
        a[i].y = __fadd_rn(__fmul_rn(a[i].y, m.y), c.y);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)k;
}
template <int MODE>
void run(const char* name, float* out, double pair_ops) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8;
  probe<MODE><<<grid, 256>>>(out, 1e-7f);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) probe<MODE><<<grid, 256>>>(out, 1e-7f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double pairs = (double)grid * 256 * kIters * 8;
  printf("%-28s %.3f ms  %.2f T pair-updates/s  (%.1f TFLOP/s at %g flop/pair)\n", name, ms, pairs / ms / 1e9, pairs * pair_ops / ms / 1e9, pair_ops);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("scalar FFMA x2", out, 4);
  run<1>("FFMA2", out, 4);
  run<2>("scalar FFMA x2 + IMAD", out, 4);
  run<3>("FFMA2 + IMAD", out, 4);
  run<4>("FMUL2 + FADD2", out, 4);
  run<5>("scalar FMUL,FADD x2", out, 4);
  return 0;
}
