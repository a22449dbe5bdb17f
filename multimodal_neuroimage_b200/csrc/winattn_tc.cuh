// winattn_tc.cuh -- shifted-window attention on the Blackwell tensor cores
// (tcgen05.mma + TMEM accumulators + TMA), bf16 in / bf16 out, fp32 softmax.
//
// Tuned shape: 64-token windows (4x4x4, 8x8, or 64 pre-windowed tokens), head_dim 32.
// A CTA tile is a PAIR of windows x one head: 128 query rows = the 128 TMEM lanes.
//
//   warp 8   TMA producer   q/k/v tiles of the pair, gathered straight out of the un-windowed,
//                           un-shifted (B,D,H,W,3C) tensor with 5-D tensor maps.  The cyclic
//                           shift is a coordinate offset; a window that wraps around the volume
//                           edge is fetched as 2 / 2*w0 / 2*w0*w1 boxes (split along the
//                           innermost wrapping axis), each landing at its window-order rows.
//   warp 9   MMA issuer     S = Q K^T  (M128 N128 K32, block diagonal = the two windows),
//                           O = P V    (M128 N32 K128); accumulators in TMEM.
//   warps 0-7 softmax       two threads per query row (32 keys each): tcgen05.ld the logits, apply
//                           cosine normalisation / logit scale (or q scale), relative position
//                           bias, shift mask (region ids computed from coordinates), softmax in
//                           fp32, write P (bf16, 128B-swizzled K-major) for the second MMA,
//                           then scale O by 1/l, stage it and TMA-store it back through the
//                           same boxes (= window_reverse + roll back).
//
// Shared memory (dynamic, 1024-aligned): 3 stages x (Q 8K | K 8K | V 8K), P 32K, O stage 8K.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <mutex>

#include "../../include/mmn_b200.h"
#include "tc_common.cuh"

namespace mmn { namespace tc {

constexpr int kN = 64;            // tokens per window
constexpr int kD = 32;            // head_dim
constexpr int kStages = 3;
constexpr int kTile = 128 * 64;   // bytes of one operand tile: 2 windows x 64 rows x 64 B
constexpr int kWinBytes = 64 * 64;
constexpr int kSoftmaxThreads = 256;   // warps 0-7: two threads per query row (32 keys each)
constexpr int kProducerWarp = 8, kMmaWarp = 9;
constexpr int kFwdThreads = 320;
constexpr int kBiasLd = 68;             // padded row of the shared bias table (conflict-free float4 rows)
constexpr int kTmemCols = 256;    // S: columns [0,128), O: columns [128,160)
constexpr float kLog2e = 1.4426950408889634f;

struct FwdParams {
  CUtensorMap q[4], k[4], v[4], o[4];   // box shapes: 0 whole window, 1 half along axis 0, 2 along axis 1, 3 along axis 2
  int nH, n_pairs, nW;
  int grid[3], win[3], shift[3], nwin[3];
  int cosine, mask_kind, mask_windows;
  float scale;
  const float* bias;
  const float* head_scale;
  const float* mask;
  float* lse;
};

struct WinGeom {
  int b, start[3], idx[3];
  int aw;        // innermost wrapping axis, -1 if the window does not wrap
  int nbox;
};

__device__ __forceinline__ WinGeom decode_window(const FwdParams& P, int w) {
  WinGeom g;
  g.b = w / P.nW;
  int wl = w - g.b * P.nW;
  g.idx[2] = wl % P.nwin[2];
  int t = wl / P.nwin[2];
  g.idx[1] = t % P.nwin[1];
  g.idx[0] = t / P.nwin[1];
  g.aw = -1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    g.start[a] = g.idx[a] * P.win[a] + P.shift[a];
    if (g.start[a] + P.win[a] > P.grid[a]) g.aw = a;
  }
  g.nbox = g.aw == 2 ? 2 * P.win[0] * P.win[1] : (g.aw == 1 ? 2 * P.win[0] : (g.aw == 0 ? 2 : 1));
  return g;
}

// Issue this lane's share of the TMA boxes of one window (load into / store from the
// window's 4 KB slot `tile`).  Box `bi` covers window positions [p0, p0 + box tokens).
template <bool LOAD>
__device__ __forceinline__ void issue_boxes(const FwdParams& P, const CUtensorMap* maps, const WinGeom& g, int chan,
                                            uint8_t* tile, uint64_t* bar, int lane) {
  const int w0 = P.win[0], w1 = P.win[1], w2 = P.win[2];
  for (int bi = lane; bi < g.nbox; bi += 32) {
    int a0 = 0, a1 = 0, a2 = 0, shape = 0;
    if (g.aw == 0) { shape = 1; a0 = bi * (w0 >> 1); }
    else if (g.aw == 1) { shape = 2; a0 = bi >> 1; a1 = (bi & 1) * (w1 >> 1); }
    else if (g.aw == 2) { shape = 3; a0 = bi / (2 * w1); a1 = (bi >> 1) % w1; a2 = (bi & 1) * (w2 >> 1); }
    int c0 = g.start[0] + a0; if (c0 >= P.grid[0]) c0 -= P.grid[0];
    int c1 = g.start[1] + a1; if (c1 >= P.grid[1]) c1 -= P.grid[1];
    int c2 = g.start[2] + a2; if (c2 >= P.grid[2]) c2 -= P.grid[2];
    uint8_t* p = tile + ((a0 * w1 + a1) * w2 + a2) * 64;
    if (LOAD) tma_load_5d(&maps[shape], bar, p, chan, c2, c1, c0, g.b);
    else tma_store_5d(&maps[shape], p, chan, c2, c1, c0, g.b);
  }
}

// Region id of in-window position p of window g in the shifted frame (swin_v2_module.py:247-258).
__device__ __forceinline__ int region_id(const FwdParams& P, const WinGeom& g, int p) {
  int a2 = p % P.win[2]; int t = p / P.win[2];
  int a1 = t % P.win[1]; int a0 = t / P.win[1];
  int a[3] = {a0, a1, a2};
  int rid = 0;
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    int v = g.idx[x] * P.win[x] + a[x];
    int r = P.shift[x] == 0 ? 0 : (v < P.grid[x] - P.win[x] ? 0 : (v < P.grid[x] - P.shift[x] ? 1 : 2));
    rid = rid * 3 + r;
  }
  return rid;
}

__global__ void __launch_bounds__(kFwdThreads, 1)
winattn_fwd_tc_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024-B alignment
  uint8_t* sQKV = smem;                                  // kStages x 3 x kTile
  uint8_t* sP = sQKV + kStages * 3 * kTile;              // 2 x 16 KB (key halves)
  uint8_t* sO = sP + 2 * 16384;                          // 8 KB
  float* sBias = reinterpret_cast<float*>(sO + kTile);   // [64][kBiasLd] fp32: this CTA's head
  float* sRq = sBias + kN * kBiasLd;                     // 128: per-row logit multiplier
  float* sRk = sRq + 128;                                // 128: per-key 1/||k||
  float* sMax = sRk + 128;                               // [2][128] partial row maxima
  float* sSum = sMax + 256;                              // [2][128] partial row sums
  int* sRid = reinterpret_cast<int*>(sSum + 256);        // 128 region ids
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRid + 128);
  uint64_t* full = bars;                                 // [kStages]
  uint64_t* empty = bars + kStages;                      // [kStages]
  uint64_t* s_full = bars + 2 * kStages;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = s_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A CTA serves ONE head (its bias table stays in shared memory) and strides over window pairs.
  const int h = blockIdx.x % P.nH;
  const int pair0 = blockIdx.x / P.nH, pair_step = gridDim.x / P.nH;

  // ---- one-time setup
  for (int i = tid; i < 2 * 16384 / 16; i += kFwdThreads) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
  if (P.bias)
    for (int i = tid; i < kN * kN; i += kFwdThreads) sBias[(i >> 6) * kBiasLd + (i & 63)] = __ldg(P.bias + (size_t)h * kN * kN + i);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, kSoftmaxThreads); mbar_init(p_full, kSoftmaxThreads); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0) {
    for (int i = 0; i < 4; ++i) { tma_prefetch_desc(&P.q[i]); tma_prefetch_desc(&P.k[i]); tma_prefetch_desc(&P.v[i]); tma_prefetch_desc(&P.o[i]); }
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  fence_proxy_async_smem();            // zeroed P must be visible to the tensor-core (async) proxy
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kProducerWarp) {
    // ============================== TMA producer ==============================
    int it = 0;
    for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it) {
      const int stage = it % kStages, phase = (it / kStages) & 1;
      mbar_wait(&empty[stage], phase ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full[stage], 3 * kTile);
      __syncwarp();
      uint8_t* base = sQKV + stage * 3 * kTile;
#pragma unroll
      for (int slot = 0; slot < 2; ++slot) {
        WinGeom g = decode_window(P, pair * 2 + slot);
        issue_boxes<true>(P, P.q, g, h * kD, base + slot * kWinBytes, &full[stage], lane);
        issue_boxes<true>(P, P.k, g, h * kD, base + kTile + slot * kWinBytes, &full[stage], lane);
        issue_boxes<true>(P, P.v, g, h * kD, base + 2 * kTile + slot * kWinBytes, &full[stage], lane);
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idescS = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idescO = umma_idesc_bf16(128, 32, 0, 1);    // P (K-major) x V (MN-major)
    const uint32_t tS = tmem, tO = tmem + 128;
    const uint32_t pAddr = smem_u32(sP);
    int it = 0;
    for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it) {
      const int stage = it % kStages, phase = (it / kStages) & 1;
      const uint32_t qAddr = smem_u32(sQKV + stage * 3 * kTile), kAddr = qAddr + kTile, vAddr = qAddr + 2 * kTile;
      mbar_wait(&full[stage], phase);
      mbar_wait(s_empty, (it & 1) ^ 1);
      tcgen05_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_bf16_ss(tS, umma_smem_desc(qAddr + ks * 32, 0, 512, kSwz64), umma_smem_desc(kAddr + ks * 32, 0, 512, kSwz64),
                       idescS, ks);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, it & 1);
      tcgen05_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16_ss(tO, umma_smem_desc(pAddr + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024, kSwz128),
                       umma_smem_desc(vAddr + ks * 1024, 8192, 512, kSwz64), idescO, ks);
        umma_commit(o_full);
        umma_commit(&empty[stage]);
      }
      __syncwarp();
    }
  } else {
    // ============================== softmax / epilogue (256 threads: 2 per query row) ==============================
    const int r = tid & 127, half = tid >> 7;            // row of the pair tile; which 32 of its 64 keys
    const int slot = r >> 6, i = r & 63;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float hscale = P.cosine ? __ldg(P.head_scale + h) : 1.f;
    int it = 0;
    for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it) {
      const int stage = it % kStages, phase = (it / kStages) & 1;
      const int w = pair * 2 + slot;
      const WinGeom g = decode_window(P, w);
      const bool masked = P.mask_kind == MMN_MASK_SHIFT && g.aw >= 0;   // uniform over the threads of a window
      const uint8_t* base = sQKV + stage * 3 * kTile;

      mbar_wait(&full[stage], phase);
      if (P.cosine) {
        // half 0 owns ||q_r||, half 1 owns ||k_r|| (the swizzle only permutes 16-byte chunks inside the 64-byte row)
        const uint4* row = reinterpret_cast<const uint4*>(base + half * kTile + r * 64);
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 a = row[c];
          const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(pa[e]); ss += f.x * f.x + f.y * f.y; }
        }
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        if (half == 0) sRq[r] = inv * hscale; else sRk[r] = inv;
      }
      int rid_i = 0;
      if (masked) { rid_i = region_id(P, g, i); if (half == 0) sRid[r] = rid_i; }
      named_bar_sync(1, kSoftmaxThreads);

      // ---- logits of this thread's 32 keys
      mbar_wait(s_full, it & 1);
      tcgen05_fence_after();
      uint32_t raw[32];
      tmem_ld_32x32b_x32(tmem + lane_base + slot * 64 + half * 32, raw);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(s_empty);

      float s[32];
      const float a_i = P.cosine ? sRq[r] : P.scale;
      const float4* brow = reinterpret_cast<const float4*>(sBias + i * kBiasLd + half * 32);
      const float4* krow = reinterpret_cast<const float4*>(sRk + slot * 64 + half * 32);
      const float4* mrow = P.mask_kind == MMN_MASK_TENSOR
                               ? reinterpret_cast<const float4*>(P.mask + ((size_t)(w % P.mask_windows) * kN + i) * kN + half * 32)
                               : nullptr;
      float mx = -INFINITY;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 bb = P.bias ? brow[j4] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (mrow) { float4 mm = __ldg(mrow + j4); bb.x += mm.x; bb.y += mm.y; bb.z += mm.z; bb.w += mm.w; }
        float4 kk = P.cosine ? krow[j4] : make_float4(1.f, 1.f, 1.f, 1.f);
        const float add[4] = {bb.x, bb.y, bb.z, bb.w};
        const float rk[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j4 * 4 + e;
          float v = fmaf(__uint_as_float(raw[j]) * rk[e], a_i, add[e]);
          if (masked && sRid[slot * 64 + half * 32 + j] != rid_i) v -= 100.f;
          s[j] = v;
          mx = fmaxf(mx, v);
        }
      }
      sMax[half * 128 + r] = mx;
      named_bar_sync(2, kSoftmaxThreads);
      mx = fmaxf(mx, sMax[(half ^ 1) * 128 + r]);
      float l = 0.f;
      const float mneg = -mx * kLog2e;
#pragma unroll
      for (int j = 0; j < 32; ++j) { s[j] = fast_exp2(fmaf(s[j], kLog2e, mneg)); l += s[j]; }
      sSum[half * 128 + r] = l;

      // ---- P (bf16) into the 128B-swizzled K-major tile of this window's key half
      {
        uint8_t* prow = sP + slot * 16384 + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v4 = make_uint4(pack_bf16x2(s[c * 8 + 0], s[c * 8 + 1]), pack_bf16x2(s[c * 8 + 2], s[c * 8 + 3]),
                                pack_bf16x2(s[c * 8 + 4], s[c * 8 + 5]), pack_bf16x2(s[c * 8 + 6], s[c * 8 + 7]));
          *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (r & 7)) << 4)) = v4;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(p_full);

      // ---- O epilogue: half 0 takes output channels [0,16), half 1 [16,32)
      if (warp == 0) tma_store_wait_read<0>();        // previous pair's stores have drained sO
      mbar_wait(o_full, it & 1);
      tcgen05_fence_after();
      uint32_t oraw[16];
      tmem_ld_32x32b_x16(tmem + lane_base + 128 + half * 16, oraw);
      tmem_ld_wait();
      tcgen05_fence_before();
      named_bar_sync(3, kSoftmaxThreads);             // sO free (warp 0 waited) and sSum complete
      {
        l = sSum[r] + sSum[128 + r];
        if (half == 0) P.lse[((size_t)w * P.nH + h) * kN + i] = mx + __logf(l);
        const float inv = 1.f / l;
        uint8_t* orow = sO + r * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 v4 = make_uint4(pack_bf16x2(__uint_as_float(oraw[c * 8 + 0]) * inv, __uint_as_float(oraw[c * 8 + 1]) * inv),
                                pack_bf16x2(__uint_as_float(oraw[c * 8 + 2]) * inv, __uint_as_float(oraw[c * 8 + 3]) * inv),
                                pack_bf16x2(__uint_as_float(oraw[c * 8 + 4]) * inv, __uint_as_float(oraw[c * 8 + 5]) * inv),
                                pack_bf16x2(__uint_as_float(oraw[c * 8 + 6]) * inv, __uint_as_float(oraw[c * 8 + 7]) * inv));
          *reinterpret_cast<uint4*>(orow + (((half * 2 + c) ^ ((r >> 1) & 3)) << 4)) = v4;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(4, kSoftmaxThreads);
      if (warp == 0) {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          WinGeom gs = decode_window(P, pair * 2 + sl);
          issue_boxes<false>(P, P.o, gs, h * kD, sO + sl * kWinBytes, nullptr, lane);
        }
        tma_store_commit();
      }
    }
    if (warp == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem);
}

constexpr size_t kFwdSmemBytes = 1024 /*align slack*/ + kStages * 3 * kTile + 2 * 16384 + kTile + kN * kBiasLd * 4 +
                                 (128 + 128 + 256 + 256 + 128) * 4 + 16 * 8;

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

struct RGeom { int grid[3], win[3], shift[3], nwin[3]; };   // right-aligned geometry
inline RGeom right_align(const mmn_winattn_desc* d) {
  RGeom g;
  for (int a = 0; a < 3; ++a) { g.grid[a] = 1; g.win[a] = 1; g.shift[a] = 0; g.nwin[a] = 1; }
  for (int a = 0; a < d->ndim; ++a) {
    int t = 3 - d->ndim + a;
    g.grid[t] = d->grid[a]; g.win[t] = d->window[a]; g.shift[t] = d->shift[a]; g.nwin[t] = d->grid[a] / d->window[a];
  }
  return g;
}

inline const char* why_not(const mmn_winattn_desc* d) {
  if (d->io_dtype != MMN_DT_BF16) return "io dtype is not bf16";
  if (d->head_dim != kD) return "head_dim != 32";
  if (d->dropout_p > 0.f) return "attention dropout is only implemented in the generic path";
  RGeom g = right_align(d);
  if (g.win[0] * g.win[1] * g.win[2] != kN) return "window does not hold 64 tokens";
  for (int a = 0; a < 3; ++a)
    if (g.shift[a] != 0 && (2 * g.shift[a] != g.win[a])) return "shift is neither 0 nor window/2";
  if (g.shift[2] != 0 && g.win[2] % 4 != 0) return "innermost window extent not a multiple of 4 (TMA 128-byte smem alignment)";
  if (g.shift[1] != 0 && ((g.win[1] / 2) * g.win[2]) % 2 != 0) return "half-window rows not 128-byte aligned";
  if (g.shift[0] != 0 && ((g.win[0] / 2) * g.win[1] * g.win[2]) % 2 != 0) return "half-window slabs not 128-byte aligned";
  long long nwin = (long long)d->batch * g.nwin[0] * g.nwin[1] * g.nwin[2];
  if (nwin % 2 != 0) return "odd number of windows";
  if (d->q_row_stride % 8 || d->k_row_stride % 8 || d->v_row_stride % 8 || d->o_row_stride % 8) return "row stride not 16-byte aligned";
  if (!encode_fn()) return "cuTensorMapEncodeTiled unavailable";
  return nullptr;
}
inline bool winattn_supported(const mmn_winattn_desc* d) { return why_not(d) == nullptr; }
inline bool winattn_bwd_supported(const mmn_winattn_desc*) { return false; }

// Four tensor maps (one per box shape) over a (B, g0, g1, g2, C) bf16 tensor with token stride `row_stride`.
inline bool make_maps(CUtensorMap* out, const void* ptr, long long row_stride, int batch, int channels, const RGeom& g) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  if (reinterpret_cast<uintptr_t>(ptr) % 16) return false;
  cuuint64_t dims[5] = {(cuuint64_t)channels, (cuuint64_t)g.grid[2], (cuuint64_t)g.grid[1], (cuuint64_t)g.grid[0], (cuuint64_t)batch};
  cuuint64_t rs = (cuuint64_t)row_stride * 2;
  cuuint64_t strides[4] = {rs, rs * g.grid[2], rs * g.grid[2] * g.grid[1], rs * g.grid[2] * g.grid[1] * g.grid[0]};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  auto half = [](int w) { return (cuuint32_t)(w / 2 > 0 ? w / 2 : 1); };
  cuuint32_t boxes[4][5] = {
      {kD, (cuuint32_t)g.win[2], (cuuint32_t)g.win[1], (cuuint32_t)g.win[0], 1},
      {kD, (cuuint32_t)g.win[2], (cuuint32_t)g.win[1], half(g.win[0]), 1},
      {kD, (cuuint32_t)g.win[2], half(g.win[1]), 1, 1},
      {kD, half(g.win[2]), 1, 1, 1}};
  for (int s = 0; s < 4; ++s) {
    CUresult r = enc(&out[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, boxes[s], estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
  }
  return true;
}

inline int winattn_fwd(const mmn_winattn_desc* d, const void* q, const void* k, const void* v, const float* bias,
                       const float* head_scale, const float* mask, void* out, float* lse, cudaStream_t st, char* err,
                       size_t errlen) {
  RGeom g = right_align(d);
  FwdParams P;
  const int C = d->num_heads * d->head_dim;
  if (!make_maps(P.q, q, d->q_row_stride, d->batch, C, g) || !make_maps(P.k, k, d->k_row_stride, d->batch, C, g) ||
      !make_maps(P.v, v, d->v_row_stride, d->batch, C, g) || !make_maps(P.o, out, d->o_row_stride, d->batch, C, g)) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (pointer alignment or strides)");
    return MMN_ERR_CUDA;
  }
  P.nH = d->num_heads;
  P.nW = g.nwin[0] * g.nwin[1] * g.nwin[2];
  P.n_pairs = d->batch * P.nW / 2;
  for (int a = 0; a < 3; ++a) { P.grid[a] = g.grid[a]; P.win[a] = g.win[a]; P.shift[a] = g.shift[a]; P.nwin[a] = g.nwin[a]; }
  P.cosine = d->score_kind == MMN_SCORE_COSINE;
  P.mask_kind = d->mask_kind;
  P.mask_windows = d->mask_windows > 0 ? d->mask_windows : 1;
  P.scale = d->scale;
  P.bias = bias; P.head_scale = head_scale; P.mask = mask; P.lse = lse;

  static std::once_flag once;
  static int num_sms = 148;
  std::call_once(once, [] {
    cudaFuncSetAttribute(winattn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemBytes);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
  int per_head = num_sms / P.nH;                 // CTAs per head (each CTA keeps one head's bias table resident)
  if (per_head < 1) per_head = 1;
  if (per_head > P.n_pairs) per_head = P.n_pairs;
  const int grid = per_head * P.nH;
  winattn_fwd_tc_kernel<<<grid, kFwdThreads, kFwdSmemBytes, st>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "winattn_fwd_tc_kernel: %s", cudaGetErrorString(e));
    return MMN_ERR_CUDA;
  }
  return MMN_OK;
}

inline int winattn_bwd(const mmn_winattn_desc*, const void*, const void*, const void*, const float*, const float*,
                       const float*, const void*, const float*, const void*, void*, void*, void*, float*, float*, float*,
                       cudaStream_t, char* err, size_t errlen) {
  snprintf(err, errlen, "tcgen05 backward not built");
  return MMN_ERR_UNSUPPORTED;
}

}}  // namespace mmn::tc
