// f32x2_bench.cu -- issue rate of packed fp32 (FFMA2) vs scalar FFMA on one SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/f32x2_bench.cu -o tools/f32x2_bench.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float s[16]; uint64_t p[8];
  for (int i = 0; i < 16; ++i) s[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = ((uint64_t)__float_as_uint(s[2 * i]) << 32) | __float_as_uint(s[2 * i + 1]);
  const float b = 1.0001f, c = 0.5f;
  const uint64_t b2 = ((uint64_t)__float_as_uint(b) << 32) | __float_as_uint(b), c2 = ((uint64_t)__float_as_uint(c) << 32) | __float_as_uint(c);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = fma1(s[i], b, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], b2, c2);
    } else {           // mixed: 8 FFMA2 + 8 FFMA
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], b2, c2); s[i] = fma1(s[i], b, c); }
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 16; ++i) acc += s[i];
  for (int i = 0; i < 8; ++i) acc += __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps : {1, 2, 4, 8, 16}) {
    long long h[3];
    k<0><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h[0], cyc, 8, cudaMemcpyDeviceToHost);
    k<1><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h[1], cyc, 8, cudaMemcpyDeviceToHost);
    k<2><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h[2], cyc, 8, cudaMemcpyDeviceToHost);
    // per SMSP: warps/4 warps (min 1).  flops-lanes per cycle per SMSP
    double wps = warps < 4 ? 1 : warps / 4.0;
    printf("warps %2d: FFMA x16: %.2f cyc/iter (%.1f fma-lanes/clk/SMSP) | FFMA2 x8: %.2f cyc/iter (%.1f) | 8 FFMA2 + 8 FFMA: %.2f cyc/iter (%.1f)\n", warps,
           (double)h[0] / iters, 16 * 32 * wps / ((double)h[0] / iters), (double)h[1] / iters, 16 * 32 * wps / ((double)h[1] / iters),
           (double)h[2] / iters, 24 * 32 * wps / ((double)h[2] / iters));
  }
  return 0;
}
