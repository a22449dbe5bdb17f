// layernorm.cu -- LayerNorm fused with the residual add around it, forward and backward, at HBM speed.
//
// The blocks of the path interleave every attention / Mlp call with a LayerNorm and a residual add
// (swin_v2_module.py:299-302 post-norm, swinfusion_module.py:345,377-378,491-492,535-539 pre-norm,
// crossmodal_transformer.py:143-165).  PyTorch runs them as separate kernels on an fp32 residual stream (autocast), plus the
// bf16 casts feeding the next GEMM; at the Swin widths (C = 96 ... 1536, ~1e5 rows) ATen's layer_norm kernels are far
// from the memory roofline (one block per 96-element row).  Here one kernel does, per row:
//
//   mode PRE  (pre-norm blocks):   s = resid + delta;   out_sum = s;   out_norm = LN(s) * gamma + beta
//   mode POST (SwinV2 res-post-norm):   s = resid + LN(delta) * gamma + beta;   out_sum = s;   out_norm = s (cast)
//
// either input may be absent (PRE with delta == null is a plain LayerNorm), either output may be skipped, and every
// tensor is fp32 or bf16 independently (residual stream fp32, activations bf16 under autocast).  mean / rstd per row are
// kept for the backward, which returns d resid, d delta and (accumulated with atomics over row blocks) d gamma, d beta.
//
// One warp per row; lane l owns the column pairs 2 (l + 32 k): 4-byte (bf16x2) or 8-byte (float2) accesses, fully
// coalesced, values stay in registers between the statistics pass and the normalisation pass.  fp32 arithmetic.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmn_b200.h"
#include "generic_launch.h"

namespace mmn {

namespace {

struct LnTensor {            // (rows, cols) matrix, contiguous rows
  const void* p;
  int dt;                    // MMN_DT_F32 / MMN_DT_BF16
};
struct LnOut {
  void* p;
  int dt;
};

__device__ __forceinline__ float2 ld2(const LnTensor& t, long long idx) {       // idx: element index of an even column
  if (t.dt == MMN_DT_F32) return __ldg(reinterpret_cast<const float2*>(static_cast<const float*>(t.p) + idx));
  const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const __nv_bfloat16*>(t.p) + idx));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void st2(const LnOut& t, long long idx, float2 v) {
  if (t.dt == MMN_DT_F32) {
    *reinterpret_cast<float2*>(static_cast<float*>(t.p) + idx) = v;
  } else {
    __nv_bfloat162 b = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(static_cast<__nv_bfloat16*>(t.p) + idx) = b;
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kLnWarps = 8;

// K = column pairs per lane = ceil(cols / 64)
template <int K>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_fwd_kernel(LnTensor resid, LnTensor delta, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int mode,
              LnOut out_sum, LnOut out_norm, float* __restrict__ mean, float* __restrict__ rstd, long long rows, int cols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2 g[K], b[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int c = 2 * (lane + 32 * k);
    g[k] = c < cols ? __ldg(reinterpret_cast<const float2*>(gamma + c)) : make_float2(0.f, 0.f);
    b[k] = c < cols && beta ? __ldg(reinterpret_cast<const float2*>(beta + c)) : make_float2(0.f, 0.f);
  }
  const float inv_n = 1.f / (float)cols;
  for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < rows; row += (long long)gridDim.x * kLnWarps) {
    const long long base = row * cols;
    float2 x[K], r[K];          // x: the normalised quantity (PRE: resid + delta; POST: delta); r: the residual (POST)
    float s1 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = 2 * (lane + 32 * k);
      x[k] = make_float2(0.f, 0.f);
      r[k] = make_float2(0.f, 0.f);
      if (c < cols) {
        if (mode == 0) {
          if (resid.p) x[k] = ld2(resid, base + c);
          if (delta.p) { const float2 d = ld2(delta, base + c); x[k].x += d.x; x[k].y += d.y; }
        } else {
          x[k] = ld2(delta, base + c);
          if (resid.p) r[k] = ld2(resid, base + c);
        }
        s1 += x[k].x + x[k].y;
      }
    }
    const float mu = warp_sum(s1) * inv_n;
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = 2 * (lane + 32 * k);
      if (c < cols) { const float a = x[k].x - mu, d = x[k].y - mu; s2 += a * a + d * d; }
    }
    const float rs = rsqrtf(warp_sum(s2) * inv_n + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = 2 * (lane + 32 * k);
      if (c < cols) {
        const float2 n = make_float2((x[k].x - mu) * rs * g[k].x + b[k].x, (x[k].y - mu) * rs * g[k].y + b[k].y);
        if (mode == 0) {
          if (out_sum.p) st2(out_sum, base + c, x[k]);
          if (out_norm.p) st2(out_norm, base + c, n);
        } else {
          const float2 s = make_float2(r[k].x + n.x, r[k].y + n.y);
          if (out_sum.p) st2(out_sum, base + c, s);
          if (out_norm.p) st2(out_norm, base + c, s);
        }
      }
    }
  }
}

// Backward.  gs: gradient w.r.t. out_sum (optional), gn: gradient w.r.t. out_norm (optional), x: the normalised quantity
// (PRE: out_sum; POST: delta).
//   PRE :  g_x = gs + LN'(gn)      d resid = d delta = g_x          d gamma += sum_rows gn * xhat,  d beta += sum_rows gn
//   POST:  g   = gs + gn           d resid = g,  d delta = LN'(g)   d gamma += sum_rows g * xhat,   d beta += sum_rows g
// LN'(u)_c = rstd * (u_c gamma_c - mean_c(u gamma) - xhat_c mean_c(u gamma xhat)).
template <int K>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_bwd_kernel(LnTensor gs, LnTensor gn, LnTensor x, const float* __restrict__ gamma, const float* __restrict__ mean,
              const float* __restrict__ rstd, int mode, LnOut d_resid, LnOut d_delta, float* __restrict__ dgamma,
              float* __restrict__ dbeta, long long rows, int cols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2 g[K], ag[K], ab[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int c = 2 * (lane + 32 * k);
    g[k] = c < cols ? __ldg(reinterpret_cast<const float2*>(gamma + c)) : make_float2(0.f, 0.f);
    ag[k] = make_float2(0.f, 0.f);
    ab[k] = make_float2(0.f, 0.f);
  }
  const float inv_n = 1.f / (float)cols;
  for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < rows; row += (long long)gridDim.x * kLnWarps) {
    const long long base = row * cols;
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float2 u[K], xh[K], pass[K];     // u: gradient entering the LayerNorm; pass: gradient bypassing it
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = 2 * (lane + 32 * k);
      u[k] = xh[k] = pass[k] = make_float2(0.f, 0.f);
      if (c < cols) {
        const float2 xv = ld2(x, base + c);
        xh[k] = make_float2((xv.x - mu) * rs, (xv.y - mu) * rs);
        float2 a = gs.p ? ld2(gs, base + c) : make_float2(0.f, 0.f);
        float2 n = gn.p ? ld2(gn, base + c) : make_float2(0.f, 0.f);
        if (mode == 0) { u[k] = n; pass[k] = a; }
        else { u[k] = make_float2(a.x + n.x, a.y + n.y); pass[k] = u[k]; }
        ag[k].x += u[k].x * xh[k].x; ag[k].y += u[k].y * xh[k].y;
        ab[k].x += u[k].x; ab[k].y += u[k].y;
        const float ux = u[k].x * g[k].x, uy = u[k].y * g[k].y;
        m1 += ux + uy;
        m2 += ux * xh[k].x + uy * xh[k].y;
      }
    }
    m1 = warp_sum(m1) * inv_n;
    m2 = warp_sum(m2) * inv_n;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = 2 * (lane + 32 * k);
      if (c < cols) {
        const float2 ln = make_float2(rs * (u[k].x * g[k].x - m1 - xh[k].x * m2), rs * (u[k].y * g[k].y - m1 - xh[k].y * m2));
        if (mode == 0) {
          const float2 t = make_float2(pass[k].x + ln.x, pass[k].y + ln.y);
          if (d_resid.p) st2(d_resid, base + c, t);
          if (d_delta.p) st2(d_delta, base + c, t);
        } else {
          if (d_resid.p) st2(d_resid, base + c, pass[k]);
          if (d_delta.p) st2(d_delta, base + c, ln);
        }
      }
    }
  }
  // column sums over this block's rows: warps -> shared memory -> one atomic per column and block (64 columns per round)
  __shared__ float red[kLnWarps][64];
  for (int pass_id = 0; pass_id < 2; ++pass_id) {
    float* dst = pass_id == 0 ? dgamma : dbeta;
    if (!dst) continue;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float2 v = pass_id == 0 ? ag[k] : ab[k];
      red[warp][2 * lane] = v.x;
      red[warp][2 * lane + 1] = v.y;
      __syncthreads();
      const int c = 64 * k + threadIdx.x;
      if (threadIdx.x < 64 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) s += red[w][threadIdx.x];
        atomicAdd(dst + c, s);
      }
      __syncthreads();
    }
  }
}

int ln_grid(long long rows) {
  long long want = (rows + kLnWarps - 1) / kLnWarps;
  const long long cap = 148ll * 8;
  return (int)(want < cap ? want : cap);
}

}  // namespace

#define MMN_LN_DISPATCH(K_, call)                       \
  switch (K_) {                                         \
    case 1: { constexpr int K = 1; call; } break;       \
    case 2: { constexpr int K = 2; call; } break;       \
    case 3: { constexpr int K = 3; call; } break;       \
    case 4: { constexpr int K = 4; call; } break;       \
    case 6: { constexpr int K = 6; call; } break;       \
    case 8: { constexpr int K = 8; call; } break;       \
    case 12: { constexpr int K = 12; call; } break;     \
    case 16: { constexpr int K = 16; call; } break;     \
    case 24: { constexpr int K = 24; call; } break;     \
    default: return cudaErrorInvalidValue;              \
  }

static int ln_pairs(int cols) {
  const int k = (cols + 63) / 64;
  const int sizes[] = {1, 2, 3, 4, 6, 8, 12, 16, 24};
  for (int s : sizes) if (k <= s) return s;
  return 0;
}

bool layernorm_supported(int cols) { return cols >= 2 && cols % 2 == 0 && ln_pairs(cols) > 0; }

cudaError_t layernorm_fwd(const void* resid, int resid_dt, const void* delta, int delta_dt, const float* gamma, const float* beta,
                          float eps, int mode, void* out_sum, int sum_dt, void* out_norm, int norm_dt, float* mean, float* rstd,
                          long long rows, int cols, cudaStream_t st, int* launches) {
  const int kk = ln_pairs(cols);
  LnTensor r{resid, resid_dt}, d{delta, delta_dt};
  LnOut os{out_sum, sum_dt}, on{out_norm, norm_dt};
  MMN_LN_DISPATCH(kk, (ln_fwd_kernel<K><<<ln_grid(rows), kLnWarps * 32, 0, st>>>(r, d, gamma, beta, eps, mode, os, on, mean, rstd, rows, cols)));
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) ++*launches;
  return e;
}

cudaError_t layernorm_bwd(const void* g_sum, int gs_dt, const void* g_norm, int gn_dt, const void* x, int x_dt, const float* gamma,
                          const float* mean, const float* rstd, int mode, void* d_resid, int dr_dt, void* d_delta, int dd_dt,
                          float* dgamma, float* dbeta, long long rows, int cols, cudaStream_t st, int* launches) {
  const int kk = ln_pairs(cols);
  LnTensor gs{g_sum, gs_dt}, gn{g_norm, gn_dt}, xx{x, x_dt};
  LnOut dr{d_resid, dr_dt}, dd{d_delta, dd_dt};
  MMN_LN_DISPATCH(kk, (ln_bwd_kernel<K><<<ln_grid(rows), kLnWarps * 32, 0, st>>>(gs, gn, xx, gamma, mean, rstd, mode, dr, dd, dgamma, dbeta, rows, cols)));
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) ++*launches;
  return e;
}

}  // namespace mmn
