// winattn_tc_bwd.cuh -- shifted-window attention backward on the Blackwell tensor cores.
//
// Same work decomposition as the forward (winattn_tc_fwd.cuh): item = pair of 64-token
// windows x one head, 128 rows = 128 TMEM lanes, one head per CTA, 352 threads:
//
//   warp 8    TMA producer   Q, K, V, dO tiles (piece-major boxes out of the un-windowed tensors)
//   warp 9    MMA issuer     S  = Q K^T,  dP = dO V^T                       (M128 N128 K32)
//                            dV = P^T dO, dQ = dS Ks, dK = dS^T Qs          (M128 N32 K128)
//                            where P / dS are the softmax warps' bf16 tiles (block diagonal over
//                            the two windows), read K-major for dQ and MN-major (= transposed,
//                            same bytes) for dV / dK, and Qs / Ks are the q / k tiles pre-multiplied
//                            by the logit scale (x 1/||.|| for cosine attention).
//   warps 0-7 softmax        two threads per query row (32 keys each): recompute P = exp(S - lse)
//                            from the saved log-sum-exp, delta = rowsum(P o dP) exactly (the whole
//                            row is in the tile), dS = P o (dP - delta); accumulate dbias (registers
//                            for window-ordered tiles, shared-memory atomics for piece-major ones)
//                            and d(logit scale); then the epilogue: gradients through the cosine
//                            normalisation, bf16, staging tiles.
//   warp 10   TMA store      dQ, dK, dV staging tiles -> global through the same boxes.
//
// The forward output is never read (delta is recomputed), matching the generic path.
#pragma once

#include "winattn_tc_fwd.cuh"

namespace mmn { namespace tc {

constexpr int kBwdStages = 2;
constexpr int kBwdTmemCols = 512;   // S [0,128) dP [128,256) dV [256,288) dQ [288,320) dK [320,352)

struct BwdParams {
  CUtensorMap q[8], k[8], v[8], dout[8], dq[8], dk[8], dv[8];
  WinShape S;
  int nH, n_pairs;
  int cosine, mask_kind, mask_windows;
  float scale;
  const float* bias;
  const float* head_scale;
  const float* mask;
  const float* lse;
  float* dbias;         // (nH, 64, 64) accumulated, may be null
  float* dhead_scale;   // (nH) accumulated, may be null
  float* dcolsum;       // (3, nH*32) accumulated column sums of dq, dk, dv (= projection bias grads), may be null
  long long* trace;     // debug: clock64 stamps of CTA 0 (MMN_TC_TRACE_BWD=<file>)
};

__device__ __forceinline__ void trace_evb(const BwdParams& P, int role, int item, int ev) {
  if (P.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && item < 32) P.trace[(role * 32 + item) * 16 + ev] = clock64();
}

// COS: cosine attention (else scaled dot product); MASK: MMN_MASK_NONE / _SHIFT / _TENSOR.  Compile-time so that
// each variant carries only its own code (the kernel is instruction-cache sensitive).
template <bool COS, int MASK>
__global__ void __launch_bounds__(kFwdThreads, 1)
winattn_bwd_tc_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sIn = smem;                                   // kBwdStages x (Q | K | V | dO) x kTile
  uint8_t* sP = sIn + kBwdStages * 4 * kTile;            // 2 x 16 KB (key halves)
  uint8_t* sDS = sP + 2 * 16384;                         // 2 x 16 KB
  uint8_t* sQs = sDS + 2 * 16384;                        // scaled q tile (64B-swizzled rows)
  uint8_t* sKs = sQs + kTile;                            // scaled k tile
  uint8_t* sOut = sKs + kTile;                           // dQ | dK | dV staging, 3 x kTile
  float* sBias = reinterpret_cast<float*>(sOut + 3 * kTile);   // [64][kBiasLd]
  float* sDb = sBias + kN * kBiasLd;                     // [64][kBiasLd] dbias accumulator (piece-major windows)
  float* sRq = sDb + kN * kBiasLd;                       // 128: logit multiplier per row
  float* sRk = sRq + 128;                                // 128: 1/||k|| per key
  float* sDelta = sRk + 128;                             // [2][128] partial deltas
  float* sRed = sDelta + 256;                            // 8 floats: dhead_scale per softmax warp
  int* sRid = reinterpret_cast<int*>(sRed + 8);          // 128
  uint8_t* sPos = reinterpret_cast<uint8_t*>(sRid + 128); // [8][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPos + 512);
  uint64_t* full = bars;                                 // [kBwdStages]
  uint64_t* empty = bars + kBwdStages;                   // [kBwdStages] (256 arrivals: softmax threads)
  uint64_t* sdp_full = bars + 2 * kBwdStages;
  uint64_t* sdp_empty = sdp_full + 1;                    // 256 arrivals
  uint64_t* pds_full = sdp_full + 2;                     // 256 arrivals
  uint64_t* out_full = sdp_full + 3;
  uint64_t* so_ready = sdp_full + 4;                     // 256 arrivals
  uint64_t* so_free = sdp_full + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sdp_full + 6);

  const WinShape& S = P.S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % P.nH;
  const int pair0 = blockIdx.x / P.nH, pair_step = gridDim.x / P.nH;

  // ---- one-time setup
  for (int i = tid; i < 4 * 16384 / 16; i += kFwdThreads) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);   // P and dS
  for (int i = tid; i < kN * kBiasLd; i += kFwdThreads) sDb[i] = 0.f;
  if (P.bias)
    for (int i = tid; i < kN * kN; i += kFwdThreads) sBias[(i >> 6) * kBiasLd + (i & 63)] = __ldg(P.bias + (size_t)h * kN * kN + i);
  for (int i = tid; i < 512; i += kFwdThreads) sPos[i] = (uint8_t)piece_position(S, i >> 6, i & 63);
  if (tid == 0) {
    for (int s = 0; s < kBwdStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kSoftmaxThreads); }
    mbar_init(sdp_full, 1); mbar_init(sdp_empty, kSoftmaxThreads); mbar_init(pds_full, kSoftmaxThreads);
    mbar_init(out_full, 1); mbar_init(so_ready, kSoftmaxThreads); mbar_init(so_free, 1);
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0)
    for (int i = 0; i < 8; ++i) { tma_prefetch_desc(&P.q[i]); tma_prefetch_desc(&P.k[i]); tma_prefetch_desc(&P.v[i]); tma_prefetch_desc(&P.dout[i]); }
  if (warp == kStoreWarp && lane == 0)
    for (int i = 0; i < 8; ++i) { tma_prefetch_desc(&P.dq[i]); tma_prefetch_desc(&P.dk[i]); tma_prefetch_desc(&P.dv[i]); }
  if (warp == kMmaWarp) tmem_alloc<kBwdTmemCols>(tmem_slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  WinCursor step;
  step.init(S, 2 * pair_step);
  WinCursor one;
  one.b = 0; one.i0 = 0; one.i1 = 0; one.i2 = 1;

  if (warp == kProducerWarp) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      WinCursor c0;
      c0.init(S, 2 * pair0);
      int it = 0;
      for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it, c0.advance(S, step)) {
        const int stage = it % kBwdStages, phase = (it / kBwdStages) & 1;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], 4 * kTile);
        uint8_t* base = sIn + stage * 4 * kTile;
        WinCursor c = c0;
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
          const WinGeom g = window_geom(S, c);
          issue_window_boxes<true>(S, P.q, g, h * kD, base + slot * kWinBytes, &full[stage]);
          issue_window_boxes<true>(S, P.k, g, h * kD, base + kTile + slot * kWinBytes, &full[stage]);
          issue_window_boxes<true>(S, P.v, g, h * kD, base + 2 * kTile + slot * kWinBytes, &full[stage]);
          issue_window_boxes<true>(S, P.dout, g, h * kD, base + 3 * kTile + slot * kWinBytes, &full[stage]);
          c.advance(S, one);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idescS = umma_idesc_bf16(128, 128, 0, 0);    // A K-major, B K-major
    constexpr uint32_t idescKM = umma_idesc_bf16(128, 32, 0, 1);    // A K-major (dS),      B MN-major (Ks)
    constexpr uint32_t idescMM = umma_idesc_bf16(128, 32, 1, 1);    // A MN-major (P^T/dS^T), B MN-major (dO / Qs)
    const uint32_t pAddr = smem_u32(sP), dsAddr = smem_u32(sDS), qsAddr = smem_u32(sQs), ksAddr = smem_u32(sKs);
    auto issue_sdp = [&](int n) {
      const int stage = n % kBwdStages, phase = (n / kBwdStages) & 1;
      const uint32_t qAddr = smem_u32(sIn + stage * 4 * kTile), kAddr = qAddr + kTile, vAddr = qAddr + 2 * kTile, doAddr = qAddr + 3 * kTile;
      mbar_wait(&full[stage], phase);
      mbar_wait(sdp_empty, (n & 1) ^ 1);
      tcgen05_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_bf16_ss(tmem, umma_smem_desc(qAddr + ks * 32, 0, 512, kSwz64), umma_smem_desc(kAddr + ks * 32, 0, 512, kSwz64), idescS, ks);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_bf16_ss(tmem + 128, umma_smem_desc(doAddr + ks * 32, 0, 512, kSwz64), umma_smem_desc(vAddr + ks * 32, 0, 512, kSwz64), idescS, ks);
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    int it = 0;
    if (pair0 < P.n_pairs) issue_sdp(0);
    for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it) {
      const int stage = it % kBwdStages;
      const uint32_t doAddr = smem_u32(sIn + stage * 4 * kTile) + 3 * kTile;
      mbar_wait(pds_full, it & 1);
      tcgen05_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 queries (dV, dK) or 16 keys (dQ) per step
          // dV[key][d] += P[query][key]^T dO[query][d]
          umma_bf16_ss(tmem + 256, umma_smem_desc(pAddr + ks * 2048, 16384, 1024, kSwz128),
                       umma_smem_desc(doAddr + ks * 1024, 8192, 512, kSwz64), idescMM, ks);
          // dQ[query][d] += dS[query][key] Ks[key][d]
          umma_bf16_ss(tmem + 288, umma_smem_desc(dsAddr + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024, kSwz128),
                       umma_smem_desc(ksAddr + ks * 1024, 8192, 512, kSwz64), idescKM, ks);
          // dK[key][d] += dS[query][key]^T Qs[query][d]
          umma_bf16_ss(tmem + 320, umma_smem_desc(dsAddr + ks * 2048, 16384, 1024, kSwz128),
                       umma_smem_desc(qsAddr + ks * 1024, 8192, 512, kSwz64), idescMM, ks);
        }
        umma_commit(out_full);
      }
      __syncwarp();
      // S / dP of the next pair: their TMEM columns were read out long ago; Q,K,V,dO of the next stage are prefetched
      if (pair + pair_step < P.n_pairs) issue_sdp(it + 1);
    }
  } else if (warp == kStoreWarp) {
    // ============================== TMA store (+ projection-bias gradients) ==============================
    // Lane 0 issues the stores; meanwhile all 32 lanes sum the columns of the three staging tiles
    // (dq, dk, dv over all tokens = the bias gradients of the q/k/v projections).  Lane l owns the
    // 8 channels of logical 16-byte chunk l%4 for rows == l/4 (mod 8).
    {
      WinCursor c0;
      c0.init(S, 2 * pair0);
      float cs[3][8];
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) cs[t][e] = 0.f;
      int it = 0;
      for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it, c0.advance(S, step)) {
        mbar_wait(so_ready, it & 1);
        if (lane == 0) {
          WinCursor c = c0;
#pragma unroll
          for (int slot = 0; slot < 2; ++slot) {
            const WinGeom g = window_geom(S, c);
            issue_window_boxes<false>(S, P.dq, g, h * kD, sOut + slot * kWinBytes, nullptr);
            issue_window_boxes<false>(S, P.dk, g, h * kD, sOut + kTile + slot * kWinBytes, nullptr);
            issue_window_boxes<false>(S, P.dv, g, h * kD, sOut + 2 * kTile + slot * kWinBytes, nullptr);
            c.advance(S, one);
          }
          tma_store_commit();
        }
        if (P.dcolsum) {
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll 4
            for (int r0 = 0; r0 < 128; r0 += 8) {
              const int row = r0 + (lane >> 2);
              const uint4 a = *reinterpret_cast<const uint4*>(sOut + t * kTile + row * 64 + (((lane & 3) ^ ((row >> 1) & 3)) << 4));
              const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
              for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(pa[e]); cs[t][2 * e] += f.x; cs[t][2 * e + 1] += f.y; }
            }
        }
        __syncwarp();
        if (lane == 0) {
          tma_store_wait_read<0>();
          mbar_arrive(so_free);
        }
      }
      if (lane == 0) tma_store_wait_all<0>();
      if (P.dcolsum) {
        const int C = P.nH * kD;
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float v = cs[t][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 4) atomicAdd(P.dcolsum + t * C + h * kD + lane * 8 + e, v);
          }
      }
    }
  } else {
    // ============================== softmax / epilogue (256 threads: 2 per row) ==============================
    const int r = tid & 127, half = tid >> 7;
    const int slot = r >> 6, i = r & 63;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float hscale = COS ? __ldg(P.head_scale + h) : 1.f;
    const int rsw = (r >> 1) & 3;                       // 64B-swizzle phase of this thread's tile row
    float dbacc[32];                                    // dbias[i][32*half + j] over window-ordered tiles
#pragma unroll
    for (int j = 0; j < 32; ++j) dbacc[j] = 0.f;
    float dscale_acc = 0.f;
    float* const gdb_head = P.dbias ? P.dbias + (size_t)h * kN * kN : nullptr;
    const int trole = warp == 0 ? 0 : (warp == 7 ? 1 : -1);
#define TRB(item, ev) do { if (trole >= 0) trace_evb(P, trole, item, ev); } while (0)

    WinCursor cur;
    cur.init(S, 2 * pair0 + slot);
    int it = 0;
    for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it, cur.advance(S, step)) {
      const int stage = it % kBwdStages, phase = (it / kBwdStages) & 1;
      const int w = pair * 2 + slot;
      const WinGeom g = window_geom(S, cur);
      const bool masked = (MASK == MMN_MASK_SHIFT) && g.cls != 0;
      const bool permuted = (MASK == MMN_MASK_SHIFT) && (g.cls & 6) != 0;
      const uint8_t* pos = sPos + g.cls * 64;
      const int ipos = permuted ? pos[i] : i;
      const uint8_t* base = sIn + stage * 4 * kTile;
      const float lse_i = __ldg(P.lse + ((size_t)w * P.nH + h) * kN + ipos);

      // ---- (a) row norm; scaled copy of this thread's q row (half 0) / k row (half 1)
      TRB(it, 0);
      mbar_wait(&full[stage], phase);
      TRB(it, 1);
      float rinv = 1.f;                                 // 1 / max(||row||, eps)
      {
        const uint8_t* rowp = base + half * kTile + r * 64;
        uint4 raw4[4];
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          raw4[c] = *reinterpret_cast<const uint4*>(rowp + (c << 4));   // physical chunk order: fine for a sum and a copy
          const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&raw4[c]);
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(pa[e]); ss += f.x * f.x + f.y * f.y; }
        }
        if (COS) rinv = rsqrtf(fmaxf(ss, 1e-24f));
        const float mul = COS ? rinv * hscale : P.scale;
        uint8_t* dst = (half == 0 ? sQs : sKs) + r * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&raw4[c]);
          float2 f0 = __bfloat1622float2(pa[0]), f1 = __bfloat1622float2(pa[1]), f2 = __bfloat1622float2(pa[2]), f3 = __bfloat1622float2(pa[3]);
          *reinterpret_cast<uint4*>(dst + (c << 4)) = make_uint4(pack_bf16x2(f0.x * mul, f0.y * mul), pack_bf16x2(f1.x * mul, f1.y * mul),
                                                                 pack_bf16x2(f2.x * mul, f2.y * mul), pack_bf16x2(f3.x * mul, f3.y * mul));
        }
        if (COS) { if (half == 0) sRq[r] = rinv * hscale; else sRk[r] = rinv; }
      }
      int rid_i = 0;
      if (masked) { rid_i = region_id(S, g, ipos); if (half == 0) sRid[r] = rid_i; }
      named_bar_sync(1, kSoftmaxThreads);
      TRB(it, 2);

      // ---- (b) additive terms of this thread's 32 logits (bias, mask) while S / dP finish
      float p[32];
      const float* mtile = (MASK == MMN_MASK_TENSOR) ? P.mask + (size_t)(w % P.mask_windows) * kN * kN : nullptr;
      if (!permuted) {
        const float4* brow = reinterpret_cast<const float4*>(sBias + ipos * kBiasLd + half * 32);
        const float4* mrow = mtile ? reinterpret_cast<const float4*>(mtile + ipos * kN + half * 32) : nullptr;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 bb = P.bias ? brow[j4] : make_float4(0.f, 0.f, 0.f, 0.f);
          if (mrow) { float4 mm = __ldg(mrow + j4); bb.x += mm.x; bb.y += mm.y; bb.z += mm.z; bb.w += mm.w; }
          p[j4 * 4 + 0] = bb.x; p[j4 * 4 + 1] = bb.y; p[j4 * 4 + 2] = bb.z; p[j4 * 4 + 3] = bb.w;
        }
      } else {
        const uint32_t* pj4 = reinterpret_cast<const uint32_t*>(pos + half * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const uint32_t pk = pj4[j4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int jp = (pk >> (8 * e)) & 0xff;
            float add = P.bias ? sBias[ipos * kBiasLd + jp] : 0.f;
            if (mtile) add += __ldg(mtile + ipos * kN + jp);
            p[j4 * 4 + e] = add;
          }
        }
      }
      if (masked) {
        const int* rids = sRid + slot * 64 + half * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (rids[j] != rid_i) p[j] -= 100.f;
      }

      // ---- (c) S and dP from TMEM; P = exp(S - lse); partial delta and d(logit scale) sums
      TRB(it, 3);
      mbar_wait(sdp_full, it & 1);
      tcgen05_fence_after();
      TRB(it, 4);
      uint32_t raw[32], dpr[32];
      tmem_ld_32x32b_x32(tmem + lane_base + slot * 64 + half * 32, raw);
      tmem_ld_32x32b_x32(tmem + lane_base + 128 + slot * 64 + half * 32, dpr);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(sdp_empty);
      TRB(it, 5);
      const float a_i = COS ? sRq[r] : P.scale;
      const float4* krow = reinterpret_cast<const float4*>(sRk + slot * 64 + half * 32);
      const float lneg = -lse_i * kLog2e;
      float delta = 0.f, acc_pdt = 0.f, acc_pt = 0.f;   // sum p dp, sum p dp t, sum p t   (t = raw * rk)
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 kk = COS ? krow[j4] : make_float4(1.f, 1.f, 1.f, 1.f);
        const float rk[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j4 * 4 + e;
          const float t = __uint_as_float(raw[j]) * rk[e];
          const float pj = fast_exp2(fmaf(fmaf(t, a_i, p[j]), kLog2e, lneg));
          const float pd = pj * __uint_as_float(dpr[j]);
          p[j] = pj;
          delta += pd;
          acc_pdt = fmaf(pd, t, acc_pdt);
          acc_pt = fmaf(pj, t, acc_pt);
        }
      }
      sDelta[half * 128 + r] = delta;
      // P (bf16) can go out before delta is known
      {
        uint8_t* prow = sP + slot * 16384 + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16x2(p[c * 8 + 0], p[c * 8 + 1]), pack_bf16x2(p[c * 8 + 2], p[c * 8 + 3]),
                         pack_bf16x2(p[c * 8 + 4], p[c * 8 + 5]), pack_bf16x2(p[c * 8 + 6], p[c * 8 + 7]));
      }
      TRB(it, 6);
      named_bar_sync(2, kSoftmaxThreads);
      TRB(it, 7);
      delta += sDelta[(half ^ 1) * 128 + r];

      // ---- (d) dS = P o (dP - delta), in place of P; dbias and d(logit scale) reductions
#pragma unroll
      for (int j = 0; j < 32; ++j) p[j] *= __uint_as_float(dpr[j]) - delta;
      if (COS) dscale_acc += (acc_pdt - delta * acc_pt) * (a_i / hscale);   // sum_j dS_ij cos_ij, cos = raw rk_j rq_i
      if (P.dbias) {
        if (!permuted) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dbacc[j] += p[j];
        } else {
          const uint32_t* pj4 = reinterpret_cast<const uint32_t*>(pos + half * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const uint32_t pk = pj4[j4];
#pragma unroll
            for (int e = 0; e < 4; ++e) atomicAdd(&sDb[ipos * kBiasLd + ((pk >> (8 * e)) & 0xff)], p[j4 * 4 + e]);
          }
        }
      }

      // ---- (e) dS (bf16) into its 128B-swizzled tile
      {
        uint8_t* drow = sDS + slot * 16384 + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(drow + (((half * 4 + c) ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16x2(p[c * 8 + 0], p[c * 8 + 1]), pack_bf16x2(p[c * 8 + 2], p[c * 8 + 3]),
                         pack_bf16x2(p[c * 8 + 4], p[c * 8 + 5]), pack_bf16x2(p[c * 8 + 6], p[c * 8 + 7]));
      }
      fence_proxy_async_smem();
      mbar_arrive(pds_full);
      TRB(it, 8);

      // ---- (f) epilogue: half 0 -> dQ row r and dV channels [0,16); half 1 -> dK row r and dV channels [16,32)
      mbar_wait(out_full, it & 1);
      tcgen05_fence_after();
      TRB(it, 9);
      uint32_t g32[32], gv[16];
      tmem_ld_32x32b_x32(tmem + lane_base + (half == 0 ? 288 : 320), g32);
      tmem_ld_32x32b_x16(tmem + lane_base + 256 + half * 16, gv);
      tmem_ld_wait();
      tcgen05_fence_before();
      float outv[32];
      if (COS) {
        // d/dx of x / max(||x||, eps): (g - xhat (xhat . g)) / ||x||, xhat = x * rinv; x re-read from the stage tile
        float xrow[32];
        const uint8_t* rowp = base + half * kTile + r * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 a = *reinterpret_cast<const uint4*>(rowp + ((c ^ rsw) << 4));
          const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(pa[e]); xrow[c * 8 + 2 * e] = f.x * rinv; xrow[c * 8 + 2 * e + 1] = f.y * rinv; }
        }
        float proj = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) proj = fmaf(xrow[c], __uint_as_float(g32[c]), proj);
        if (rinv >= 1e12f) proj = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) outv[c] = (__uint_as_float(g32[c]) - xrow[c] * proj) * rinv;
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) outv[c] = __uint_as_float(g32[c]);
      }
      mbar_arrive(&empty[stage]);                       // this thread is done with the stage's tiles
      TRB(it, 10);
      mbar_wait(so_free, (it & 1) ^ 1);
      TRB(it, 11);                 // previous pair's stores have drained the staging tiles
      {
        uint8_t* orow = sOut + half * kTile + r * 64;   // dQ tile (half 0) or dK tile (half 1)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(orow + ((c ^ rsw) << 4)) = make_uint4(pack_bf16x2(outv[c * 8 + 0], outv[c * 8 + 1]), pack_bf16x2(outv[c * 8 + 2], outv[c * 8 + 3]),
                                                                          pack_bf16x2(outv[c * 8 + 4], outv[c * 8 + 5]), pack_bf16x2(outv[c * 8 + 6], outv[c * 8 + 7]));
        uint8_t* vrow = sOut + 2 * kTile + r * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c)
          *reinterpret_cast<uint4*>(vrow + (((half * 2 + c) ^ rsw) << 4)) =
              make_uint4(pack_bf16x2(__uint_as_float(gv[c * 8 + 0]), __uint_as_float(gv[c * 8 + 1])), pack_bf16x2(__uint_as_float(gv[c * 8 + 2]), __uint_as_float(gv[c * 8 + 3])),
                         pack_bf16x2(__uint_as_float(gv[c * 8 + 4]), __uint_as_float(gv[c * 8 + 5])), pack_bf16x2(__uint_as_float(gv[c * 8 + 6]), __uint_as_float(gv[c * 8 + 7])));
      }
      fence_proxy_async_smem();
      mbar_arrive(so_ready);
      TRB(it, 12);
    }
#undef TRB

    // ---- cross-window reductions: dbias (registers + shared table) and d(logit scale)
    if (P.dbias) {
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(gdb_head + i * kN + half * 32 + j, dbacc[j]);
      named_bar_sync(1, kSoftmaxThreads);               // all shared-memory atomics done
      for (int e = tid; e < kN * kN; e += kSoftmaxThreads) {
        float v = sDb[(e >> 6) * kBiasLd + (e & 63)];
        if (v != 0.f) atomicAdd(gdb_head + e, v);
      }
    }
    if (COS && P.dhead_scale) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dscale_acc += __shfl_xor_sync(0xffffffffu, dscale_acc, o);
      if (lane == 0) sRed[warp] = dscale_acc;
      named_bar_sync(2, kSoftmaxThreads);
      if (tid == 0) {
        float tot = 0.f;
        for (int x = 0; x < 8; ++x) tot += sRed[x];
        atomicAdd(P.dhead_scale + h, tot);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<kBwdTmemCols>(tmem);
}

constexpr size_t kBwdSmemBytes = 1024 + kBwdStages * 4 * kTile + 4 * 16384 + 2 * kTile + 3 * kTile + 2 * kN * kBiasLd * 4 +
                                 (128 + 128 + 256 + 8 + 128) * 4 + 512 + 16 * 8;

inline const char* bwd_why_not_impl(const mmn_winattn_desc* d) {
  const char* w = fwd_why_not_impl(d);
  if (w) return w;
  if (d->do_row_stride % 8 || d->dq_row_stride % 8 || d->dk_row_stride % 8 || d->dv_row_stride % 8) return "gradient row stride not 16-byte aligned";
  return nullptr;
}

inline int winattn_bwd_launch(const mmn_winattn_desc* d, const void* q, const void* k, const void* v, const float* bias,
                              const float* head_scale, const float* mask, const float* lse, const void* dout, void* dq, void* dk,
                              void* dv, float* dbias, float* dhead_scale, float* dcolsum, cudaStream_t st, char* err, size_t errlen) {
  BwdParams P;
  P.S = shape_from(d);
  const int C = d->num_heads * d->head_dim;
  const int B = d->batch;
  if (!make_window_maps(P.q, q, d->q_row_stride, B, C, P.S) || !make_window_maps(P.k, k, d->k_row_stride, B, C, P.S) ||
      !make_window_maps(P.v, v, d->v_row_stride, B, C, P.S) || !make_window_maps(P.dout, dout, d->do_row_stride, B, C, P.S) ||
      !make_window_maps(P.dq, dq, d->dq_row_stride, B, C, P.S) || !make_window_maps(P.dk, dk, d->dk_row_stride, B, C, P.S) ||
      !make_window_maps(P.dv, dv, d->dv_row_stride, B, C, P.S)) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (pointer alignment or strides)");
    return MMN_ERR_CUDA;
  }
  P.nH = d->num_heads;
  P.n_pairs = P.S.n_windows / 2;
  P.cosine = d->score_kind == MMN_SCORE_COSINE;
  P.mask_kind = d->mask_kind;
  P.mask_windows = d->mask_windows > 0 ? d->mask_windows : 1;
  P.scale = d->scale;
  P.bias = bias; P.head_scale = head_scale; P.mask = mask; P.lse = lse;
  P.dbias = bias ? dbias : nullptr;
  P.dhead_scale = P.cosine ? dhead_scale : nullptr;
  P.dcolsum = dcolsum;
  P.trace = nullptr;
  const char* trace_path = getenv("MMN_TC_TRACE_BWD");
  if (trace_path && *trace_path) {
    cudaMalloc(&P.trace, 5 * 32 * 16 * sizeof(long long));
    cudaMemsetAsync(P.trace, 0, 5 * 32 * 16 * sizeof(long long), st);
  }

  using Kern = void (*)(const BwdParams);
  static const Kern kernels[2][3] = {
      {winattn_bwd_tc_kernel<false, MMN_MASK_NONE>, winattn_bwd_tc_kernel<false, MMN_MASK_SHIFT>, winattn_bwd_tc_kernel<false, MMN_MASK_TENSOR>},
      {winattn_bwd_tc_kernel<true, MMN_MASK_NONE>, winattn_bwd_tc_kernel<true, MMN_MASK_SHIFT>, winattn_bwd_tc_kernel<true, MMN_MASK_TENSOR>}};
  static std::once_flag once;
  static int num_sms = 148;
  std::call_once(once, [] {
    for (int c = 0; c < 2; ++c)
      for (int m = 0; m < 3; ++m) cudaFuncSetAttribute(kernels[c][m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
  const Kern kern = kernels[P.cosine ? 1 : 0][P.mask_kind];
  int per_head = num_sms / P.nH;
  if (per_head < 1) per_head = 1;
  if (per_head > P.n_pairs) per_head = P.n_pairs;
  kern<<<per_head * P.nH, kFwdThreads, kBwdSmemBytes, st>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "winattn_bwd_tc_kernel: %s", cudaGetErrorString(e));
    return MMN_ERR_CUDA;
  }
  if (P.trace) {
    static long long host[5 * 32 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, P.trace, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(P.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int role = 0; role < 5; ++role)
        for (int item = 0; item < 32; ++item) {
          fprintf(f, "%d %d", role, item);
          for (int ev = 0; ev < 16; ++ev) fprintf(f, " %lld", host[(role * 32 + item) * 16 + ev]);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
  return MMN_OK;
}

}}  // namespace mmn::tc
