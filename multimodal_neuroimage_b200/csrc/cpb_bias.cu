// cpb_bias.cu -- SwinV2's continuous relative-position bias (swin_v2_module.py:158-162) as two small kernels
// forward and three backward, instead of the ~25 ATen/cuBLAS launches autograd spends on it every step:
//
//   tab[t][h]   = sum_j W2[h][j] relu(W1[j] . coords[t] + b1[j])          t < T = prod(2w-1), hidden j < J (512)
//   bias[h][e]  = 16 sigmoid(tab[index[e]][h])                            e < N*N
//
// The sizes are tiny (T = 343, J = 512, nH = 3 at BASELINE cfg2); the point is launch count, not flops.
// fp32 throughout (the bias feeds logits that reach +-100; the parity bar for this row is 1e-5).
#include <cuda_runtime.h>
#include <stdint.h>

#include "generic_launch.h"

namespace mmn {

constexpr int kCpbMaxIn = 3;      // spatial dims of the coordinate table
constexpr int kCpbMaxHeads = 64;

// grid = T blocks, 128 threads: tab16[t][h] = 16 sigmoid(...)
__global__ void __launch_bounds__(128)
cpb_table_kernel(const float* __restrict__ coords, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, int n_in, int J, int nH, float* __restrict__ tab16) {
  __shared__ float red[4][kCpbMaxHeads];
  const int t = blockIdx.x, tid = threadIdx.x;
  float c[kCpbMaxIn];
  for (int a = 0; a < kCpbMaxIn; ++a) c[a] = a < n_in ? __ldg(coords + t * n_in + a) : 0.f;
  // each thread: a strided subset of the hidden units, partial dot products with every head's W2 row
  for (int h0 = 0; h0 < nH; h0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = tid; j < J; j += 128) {
      float z = __ldg(b1 + j);
      for (int a = 0; a < n_in; ++a) z = fmaf(__ldg(w1 + j * n_in + a), c[a], z);
      z = fmaxf(z, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (h0 + k < nH) acc[k] = fmaf(__ldg(w2 + (h0 + k) * J + j), z, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0 && h0 + k < nH) red[tid >> 5][h0 + k] = v;
    }
  }
  __syncthreads();
  if (tid < nH) {
    const float s = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    tab16[t * nH + tid] = 16.f / (1.f + __expf(-s));
  }
}

// bias[h][e] = tab16[index[e]][h]
__global__ void cpb_gather_kernel(const float* __restrict__ tab16, const long long* __restrict__ index, int NN, int nH,
                                  float* __restrict__ bias) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NN) return;
  const long long t = index[e];
  for (int h = 0; h < nH; ++h) bias[(size_t)h * NN + e] = __ldg(tab16 + t * nH + h);
}

// dtab16[index[e]][h] += dbias[h][e]   (dtab16 zeroed by the caller of this kernel)
__global__ void cpb_scatter_kernel(const float* __restrict__ dbias, const long long* __restrict__ index, int NN, int nH,
                                   float* __restrict__ dtab16) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NN) return;
  const long long t = index[e];
  for (int h = 0; h < nH; ++h) atomicAdd(dtab16 + t * nH + h, __ldg(dbias + (size_t)h * NN + e));
}

// One thread per hidden unit j: walks the T table entries, recomputes its activation, and accumulates
// dW2[:, j], dW1[j, :], db1[j] in registers (no atomics: deterministic).  g[t][h] = d tab[t][h] is staged in
// shared memory by the block: d/dx 16 sigmoid(x) = tab16 (1 - tab16 / 16).
__global__ void __launch_bounds__(128)
cpb_mlp_bwd_kernel(const float* __restrict__ coords, const float* __restrict__ w1, const float* __restrict__ b1,
                   const float* __restrict__ w2, const float* __restrict__ tab16, const float* __restrict__ dtab16, int T, int n_in,
                   int J, int nH, float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2) {
  extern __shared__ float g[];                 // [T][nH]
  for (int i = threadIdx.x; i < T * nH; i += blockDim.x) {
    const float s16 = __ldg(tab16 + i);
    g[i] = __ldg(dtab16 + i) * s16 * (1.f - s16 * (1.f / 16.f));
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  float wj[kCpbMaxIn], dwj[kCpbMaxIn] = {0.f, 0.f, 0.f};
  for (int a = 0; a < kCpbMaxIn; ++a) wj[a] = a < n_in ? __ldg(w1 + j * n_in + a) : 0.f;
  const float bj = __ldg(b1 + j);
  float dbj = 0.f;
  for (int h0 = 0; h0 < nH; h0 += 8) {         // heads in groups of 8 register accumulators
    float w2j[8], dw2j[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { w2j[k] = h0 + k < nH ? __ldg(w2 + (h0 + k) * J + j) : 0.f; dw2j[k] = 0.f; }
    for (int t = 0; t < T; ++t) {
      float c[kCpbMaxIn], z = bj;
      for (int a = 0; a < kCpbMaxIn; ++a) { c[a] = a < n_in ? __ldg(coords + t * n_in + a) : 0.f; z = fmaf(wj[a], c[a], z); }
      const float act = fmaxf(z, 0.f);
      float dh = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (h0 + k < nH) {
          const float gt = g[t * nH + h0 + k];
          dw2j[k] = fmaf(gt, act, dw2j[k]);
          dh = fmaf(gt, w2j[k], dh);
        }
      if (z > 0.f) {
        dbj += dh;
        for (int a = 0; a < kCpbMaxIn; ++a) dwj[a] = fmaf(dh, c[a], dwj[a]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (h0 + k < nH) dw2[(h0 + k) * J + j] = dw2j[k];
  }
  db1[j] = dbj;
  for (int a = 0; a < n_in; ++a) dw1[j * n_in + a] = dwj[a];
}

cudaError_t cpb_bias_fwd(const float* coords, const float* w1, const float* b1, const float* w2, const long long* index, int T,
                         int n_in, int J, int nH, int NN, float* tab16, float* bias, cudaStream_t st, int* launches) {
  cpb_table_kernel<<<T, 128, 0, st>>>(coords, w1, b1, w2, n_in, J, nH, tab16);
  cpb_gather_kernel<<<(NN + 255) / 256, 256, 0, st>>>(tab16, index, NN, nH, bias);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) *launches += 2;
  return e;
}

cudaError_t cpb_bias_bwd(const float* coords, const float* w1, const float* b1, const float* w2, const long long* index,
                         const float* tab16, const float* dbias, int T, int n_in, int J, int nH, int NN, float* dtab16, float* dw1,
                         float* db1, float* dw2, cudaStream_t st, int* launches) {
  cudaError_t e = cudaMemsetAsync(dtab16, 0, (size_t)T * nH * sizeof(float), st);
  if (e != cudaSuccess) return e;
  cpb_scatter_kernel<<<(NN + 255) / 256, 256, 0, st>>>(dbias, index, NN, nH, dtab16);
  cpb_mlp_bwd_kernel<<<(J + 127) / 128, 128, (size_t)T * nH * sizeof(float), st>>>(coords, w1, b1, w2, tab16, dtab16, T, n_in, J, nH, dw1,
                                                                                  db1, dw2);
  e = cudaGetLastError();
  if (e == cudaSuccess) *launches += 2;
  return e;
}

}  // namespace mmn
