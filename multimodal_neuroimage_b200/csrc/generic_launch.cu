// generic_launch.cu -- instantiates and launches the shape-generic kernels.
#include "generic_launch.h"
#include "zero_fill.h"

#include "../../include/mmn_b200.h"

namespace mmn {
namespace {

// rows: extent of the thread-per-row side; other: extent of the staged side;
// floats_per_staged_row: shared floats per staged row excluding the 12-byte offset/rid.
mmn::GenericLaunch plan(int rows, int other, int bytes_per_staged_row) {
  mmn::GenericLaunch L{};
  L.rows_per_slot = rows < mmn::kGenericThreads ? rows : mmn::kGenericThreads;
  L.slots = mmn::kGenericThreads / L.rows_per_slot;
  const size_t budget = 96 * 1024;
  int cap = (int)(budget / ((size_t)L.slots * bytes_per_staged_row));
  if (cap > other) cap = other;
  if (cap > 256) cap = 256;
  if (cap < 1) cap = 1;
  L.chunk = cap;
  L.smem_bytes = (size_t)L.slots * cap * bytes_per_staged_row;
  return L;
}

constexpr int kMaxSmem = 100 * 1024;

template <typename T, int DMAX>
cudaError_t launch_fwd(const mmn::GenericProblem& P, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t st, int* launches) {
  mmn::GenericLaunch L = plan(P.nq, P.nk, 2 * P.d * 4 + 12);
  auto kern = mmn::attn_fwd_generic<T, DMAX>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  dim3 grid((P.n_items + L.slots - 1) / L.slots, (P.nq + L.rows_per_slot - 1) / L.rows_per_slot);
  kern<<<grid, mmn::kGenericThreads, L.smem_bytes, st>>>(P, L, (const T*)q, (const T*)k, (const T*)v, (T*)out, lse);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) ++*launches;
  return e;
}

template <typename T, int DMAX>
cudaError_t launch_bwd(const mmn::GenericProblem& P, const void* q, const void* k, const void* v, const float* lse,
               const void* dout, void* dq, void* dk, void* dv, float* dbias, float* dhs, float* ws, float* dcs, cudaStream_t st, int* launches) {
  {
    mmn::GenericLaunch L = plan(P.nq, P.nk, 2 * P.d * 4 + 12);
    auto kern = mmn::attn_bwd_dq_generic<T, DMAX>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    dim3 grid((P.n_items + L.slots - 1) / L.slots, (P.nq + L.rows_per_slot - 1) / L.rows_per_slot);
    kern<<<grid, mmn::kGenericThreads, L.smem_bytes, st>>>(P, L, (const T*)q, (const T*)k, (const T*)v, lse,
                                                          (const T*)dout, (T*)dq, dbias, dhs, ws, dcs);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  {
    mmn::GenericLaunch L = plan(P.nk, P.nq, 2 * P.d * 4 + 24);
    auto kern = mmn::attn_bwd_dkv_generic<T, DMAX>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    dim3 grid((P.n_items + L.slots - 1) / L.slots, (P.nk + L.rows_per_slot - 1) / L.rows_per_slot);
    kern<<<grid, mmn::kGenericThreads, L.smem_bytes, st>>>(P, L, (const T*)q, (const T*)k, (const T*)v, lse, ws,
                                                          (const T*)dout, (T*)dk, (T*)dv, dcs);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) ++*launches;
    return e;
  }
}

#define MMN_DISPATCH_D(T, FN, ...)                                            \
  do {                                                                         \
    if (P.d <= 4) return FN<T, 4>(__VA_ARGS__);                                \
    if (P.d <= 8) return FN<T, 8>(__VA_ARGS__);                                \
    if (P.d <= 16) return FN<T, 16>(__VA_ARGS__);                              \
    if (P.d <= 32) return FN<T, 32>(__VA_ARGS__);                              \
    if (P.d <= 64) return FN<T, 64>(__VA_ARGS__);                              \
    return FN<T, 128>(__VA_ARGS__);                                            \
  } while (0)

}  // namespace

cudaError_t generic_fwd(const GenericProblem& P, int dt, const void* q, const void* k, const void* v, void* out, float* lse,
                        cudaStream_t st, int* launches) {
  if (dt == MMN_DT_F32) MMN_DISPATCH_D(float, launch_fwd, P, q, k, v, out, lse, st, launches);
  MMN_DISPATCH_D(__nv_bfloat16, launch_fwd, P, q, k, v, out, lse, st, launches);
}

cudaError_t generic_bwd(const GenericProblem& P, int dt, const void* q, const void* k, const void* v, const float* lse,
                        const void* dout, void* dq, void* dk, void* dv, float* dbias, float* dhs, float* ws, float* dcs, cudaStream_t st,
                        int* launches) {
  if (dt == MMN_DT_F32) MMN_DISPATCH_D(float, launch_bwd, P, q, k, v, lse, dout, dq, dk, dv, dbias, dhs, ws, dcs, st, launches);
  MMN_DISPATCH_D(__nv_bfloat16, launch_bwd, P, q, k, v, lse, dout, dq, dk, dv, dbias, dhs, ws, dcs, st, launches);
}

cudaError_t generic_avg_weights(const GenericProblem& P, int dt, int batch, const void* q, const void* k, const float* lse,
                                float* avg, cudaStream_t st, int* launches) {
  long long total = (long long)batch * P.nq * P.nk;
  int blocks = (int)((total + 255) / 256);
  if (dt == MMN_DT_F32)
    mha_avg_weights_generic<float><<<blocks, 256, 0, st>>>(P, batch, (const float*)q, (const float*)k, lse, avg);
  else
    mha_avg_weights_generic<__nv_bfloat16><<<blocks, 256, 0, st>>>(P, batch, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, lse, avg);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) ++*launches;
  return e;
}

cudaError_t colsum(int dt, const void* x, long long rows, int cols, long long row_stride, float* out, cudaStream_t st, int* launches) {
  const int vcols = cols / 8;
  const int rpb = 256 / vcols;
  long long want = (rows + rpb - 1) / rpb;
  // four blocks per SM: with eight, the per-block flush (one atomic per column and block, all on the same few cache lines)
  // weighs more than the extra loads in flight help -- 10.0 -> 7.4 us at 110 592 x 96, 30.8 -> 25.7 us at x 384, equal at 8x the rows
  int blocks = (int)(want < 148 * 4 ? want : 148 * 4);
  if (dt == MMN_DT_F32) colsum_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, rows, cols, row_stride, out);
  else colsum_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, rows, cols, row_stride, out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) ++*launches;
  return e;
}

__global__ void zero_words_kernel(uint32_t* __restrict__ p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}

cudaError_t zero_words_async(void* p, size_t words, cudaStream_t st) {
  if (words == 0) return cudaSuccess;
  const size_t want = (words + 255) / 256;
  zero_words_kernel<<<(unsigned)(want < 592 ? want : 592), 256, 0, st>>>(static_cast<uint32_t*>(p), words);
  return cudaGetLastError();
}

}  // namespace mmn


