// Micro-probe: is MUFU.RCP(1.0f) exactly 1.0f (so T * rcp(1 - 0) == T bit for bit)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rcp_one rcp_one.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void k(float one, unsigned* out) {
  float r;
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(one));
  out[0] = __float_as_uint(r);
  unsigned bad = 0;
  for (unsigned i = 0; i < 100000; ++i) {
    const float T = __uint_as_float(0x30000000u + i * 2654435u % 0x0f000000u);
    if (__float_as_uint(T * r) != __float_as_uint(T)) ++bad;
  }
  out[1] = bad;
  float om = fmaf(0.0f, -1.0f, one);  // 1 - 0
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(om));
  out[2] = __float_as_uint(r);
}
int main() {
  unsigned* d; unsigned h[3];
  cudaMalloc(&d, 12);
  k<<<1, 1>>>(1.0f, d);
  cudaMemcpy(h, d, 12, cudaMemcpyDeviceToHost);
  printf("rcp.approx(1.0f) bits = 0x%08x (1.0f = 0x3f800000), mismatches T*r != T: %u, rcp(1-0) = 0x%08x\n", h[0], h[1], h[2]);
  return (h[0] == 0x3f800000u && h[1] == 0 && h[2] == 0x3f800000u) ? 0 : 1;
}
