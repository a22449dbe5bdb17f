// tma_bench.cu -- microbenchmark: how fast can one SM's TMA unit gather 4x4x4-token window
// boxes out of a (B,32,32,32,C) bf16 volume, as a function of the box's inner (channel) extent?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multimodal_neuroimage_b200/csrc \
//        tools/tma_bench.cu -o /tmp/tma_bench
// Each CTA walks windows; per window it issues `nb` loads of a (inner ch x 4 x 4 x 4) box into a
// ring of smem stages (and optionally stores them back), nothing else.  Prints GB/s.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace mmn::tc;

struct Params {
  CUtensorMap in, out;
  int n_windows, inner_bytes, boxes_per_window, do_store, stages;
};

__global__ void __launch_bounds__(64, 1) tma_bench_kernel(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[8];
  const int box_bytes = 64 * P.inner_bytes;
  const int stage_bytes = box_bytes * P.boxes_per_window;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // thread 0: producer + consumer (wait for the data, optionally store it, recycle the stage)
  int issued = 0, consumed = 0;
  const int total = (P.n_windows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int n) {
    int w = blockIdx.x + n * gridDim.x;
    int b = w / 512, wl = w % 512, i2 = wl % 8, i1 = (wl / 8) % 8, i0 = wl / 64;
    int st = n % P.stages;
    mbar_arrive_expect_tx(&full[st], stage_bytes);
    for (int q = 0; q < P.boxes_per_window; ++q)
      tma_load_5d(&P.in, &full[st], smem + st * stage_bytes + q * box_bytes, q * (P.inner_bytes / 2), i2 * 4, i1 * 4, i0 * 4, b);
  };
  while (issued < total && issued < P.stages) issue(issued++);
  while (consumed < total) {
    int st = consumed % P.stages;
    mbar_wait(&full[st], (consumed / P.stages) & 1);
    if (P.do_store) {
      int w = blockIdx.x + consumed * gridDim.x;
      int b = w / 512, wl = w % 512, i2 = wl % 8, i1 = (wl / 8) % 8, i0 = wl / 64;
      for (int q = 0; q < P.boxes_per_window; ++q)
        tma_store_5d(&P.out, smem + st * stage_bytes + q * box_bytes, q * (P.inner_bytes / 2), i2 * 4, i1 * 4, i0 * 4, b);
      tma_store_commit();
      tma_store_wait_read<0>();
    }
    ++consumed;
    if (issued < total) issue(issued++);
  }
  if (P.do_store) tma_store_wait_all<0>();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int B = 16, C = 288;                       // packed qkv rows: 288 bf16 = 576 B per token
  size_t elems = (size_t)B * 32 * 32 * 32 * C;
  void *din, *dout;
  cudaMalloc(&din, elems * 2);
  cudaMalloc(&dout, elems * 2);
  cudaMemset(din, 1, elems * 2);
  struct Cfg { int inner_bytes, boxes, store, stages; };
  std::vector<Cfg> cfgs = {{64, 1, 0, 4}, {64, 3, 0, 4}, {64, 9, 0, 2}, {128, 1, 0, 4}, {128, 2, 0, 4}, {128, 4, 0, 2},
                           {64, 3, 1, 4}, {128, 2, 1, 4}, {32, 4, 0, 4}, {64, 6, 0, 3}};
  for (auto c : cfgs) {
    Params P;
    cuuint64_t dims[5] = {(cuuint64_t)C, 32, 32, 32, (cuuint64_t)B};
    cuuint64_t rs = (cuuint64_t)C * 2;
    cuuint64_t strides[4] = {rs, rs * 32, rs * 32 * 32, rs * 32 * 32 * 32};
    cuuint32_t box[5] = {(cuuint32_t)(c.inner_bytes / 2), 4, 4, 4, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUtensorMapSwizzle sw = c.inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (c.inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r1 = enc(&P.in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, din, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&P.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dout, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 || r2) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
    P.n_windows = B * 512; P.inner_bytes = c.inner_bytes; P.boxes_per_window = c.boxes; P.do_store = c.store; P.stages = c.stages;
    size_t smem = 1024 + (size_t)c.stages * c.boxes * 64 * c.inner_bytes;
    cudaFuncSetAttribute(tma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148, 296}) {
      if (grid == 296 && smem > 110 * 1024) continue;
      tma_bench_kernel<<<grid, 64, smem>>>(P);
      cudaEventRecord(e0);
      for (int i = 0; i < 5; ++i) tma_bench_kernel<<<grid, 64, smem>>>(P);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      double bytes = (double)P.n_windows * c.boxes * 64 * c.inner_bytes * (c.store ? 2 : 1);
      printf("inner %3d B x %d boxes/window, store %d, stages %d, grid %3d: %8.3f ms  %8.1f GB/s  (%s)\n", c.inner_bytes, c.boxes,
             c.store, c.stages, grid, ms, bytes / ms / 1e6, cudaGetErrorString(e));
    }
  }
  return 0;
}
