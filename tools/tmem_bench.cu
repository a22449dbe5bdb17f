// tmem_bench.cu -- microbenchmark: tcgen05.ld throughput of one SM as a function of the number of warps reading
// and of the load shape (32x32b.x16 / .x32 / .x64).  The window-attention backward reads 144 KB of TMEM per item
// (S, dP twice, three gradient accumulators); this says how many cycles that costs at best.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multimodal_neuroimage_b200/csrc \
//        tools/tmem_bench.cu -o tools/tmem_bench.bin
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

using namespace mmn::tc;

template <int X>
__device__ __forceinline__ uint32_t ld_once(uint32_t taddr) {
  uint32_t acc = 0;
  if constexpr (X == 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else if constexpr (X == 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  } else {
    uint32_t r[32], q[32];
    tmem_ld_32x32b_x32(taddr, r);
    tmem_ld_32x32b_x32(taddr + 32, q);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i] ^ q[i];
  }
  return acc;
}

// two loads in flight per warp before the wait (x16 + x16 at different columns), as the backward's softmax does
__device__ __forceinline__ uint32_t ld_pair16(uint32_t taddr) {
  uint32_t r[16], q[16], acc = 0;
  tmem_ld_32x32b_x16(taddr, r);
  tmem_ld_32x32b_x16(taddr + 64, q);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) acc ^= r[i] ^ q[i];
  return acc;
}

// N loads of 16 columns in flight, one wait (N * 16 registers)
template <int N>
__device__ __forceinline__ uint32_t ld_multi16(uint32_t taddr) {
  uint32_t r[N][16], acc = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) tmem_ld_32x32b_x16(taddr + k * 16, r[k]);
  tmem_ld_wait();
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[k][i];
  return acc;
}
// N loads of 32 columns in flight, one wait
template <int N>
__device__ __forceinline__ uint32_t ld_multi32(uint32_t taddr) {
  uint32_t r[N][32], acc = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) tmem_ld_32x32b_x32(taddr + k * 32, r[k]);
  tmem_ld_wait();
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[k][i];
  return acc;
}
// N loads of 8 columns, one wait
template <int N>
__device__ __forceinline__ uint32_t ld_multi8(uint32_t taddr) {
  uint32_t r[N][8], acc = 0;
#pragma unroll
  for (int k = 0; k < N; ++k)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[k][0]), "=r"(r[k][1]), "=r"(r[k][2]), "=r"(r[k][3]), "=r"(r[k][4]), "=r"(r[k][5]), "=r"(r[k][6]), "=r"(r[k][7])
                 : "r"(taddr + k * 8));
  tmem_ld_wait();
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[k][i];
  return acc;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_bench_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t col = ((warp >> 2) * 128) & 255;            // warps sharing a lane quadrant read different columns
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = tmem + lane_base + col + (it & 1) * 256;
    if constexpr (MODE == 0) acc ^= ld_once<16>(a);
    else if constexpr (MODE == 1) acc ^= ld_once<32>(a);
    else if constexpr (MODE == 2) acc ^= ld_once<64>(a);
    else if constexpr (MODE == 3) acc ^= ld_pair16(a);
    else if constexpr (MODE == 4) acc ^= ld_multi16<4>(a);
    else if constexpr (MODE == 5) acc ^= ld_multi16<8>(a);
    else if constexpr (MODE == 6) acc ^= ld_multi32<4>(a);
    else if constexpr (MODE == 7) acc ^= ld_multi8<8>(a);
    else acc ^= ld_multi8<16>(a);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 8 * 148);
  cudaMalloc(&sink, 4096);
  const int iters = 2000;
  const char* names[9] = {"32x32b.x16", "32x32b.x32", "2 x 32x32b.x32", "2 x 32x32b.x16 (one wait)", "4 x x16 (one wait)", "8 x x16 (one wait)",
                          "4 x x32 (one wait)", "8 x x8 (one wait)", "16 x x8 (one wait)"};
  const int bytes_per_warp_iter[9] = {32 * 16 * 4, 32 * 32 * 4, 32 * 64 * 4, 32 * 32 * 4, 32 * 64 * 4, 32 * 128 * 4, 32 * 128 * 4, 32 * 64 * 4, 32 * 128 * 4};
  for (int mode = 0; mode < 9; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tmem_bench_kernel<0><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 1) tmem_bench_kernel<1><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 2) tmem_bench_kernel<2><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 3) tmem_bench_kernel<3><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 4) tmem_bench_kernel<4><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 5) tmem_bench_kernel<5><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 6) tmem_bench_kernel<6><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 7) tmem_bench_kernel<7><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 8) tmem_bench_kernel<8><<<1, warps * 32>>>(iters, cyc, sink);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)bytes_per_warp_iter[mode] * warps * iters;
      printf("%-28s warps %2d: %8.1f cycles/iter  %7.1f B/cycle/SM  (%.1f B/cycle per lane quadrant in use)\n", names[mode], warps,
             (double)c / iters, bytes / c, bytes / c / (warps < 4 ? warps : 4));
    }
  return 0;
}
