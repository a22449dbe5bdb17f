// winattn_tc_fwd.cuh -- shifted-window attention forward on the Blackwell tensor cores
// (tcgen05.mma + TMEM accumulators + TMA), bf16 in / bf16 out, fp32 softmax.
//
// Tuned shape: 64-token windows (4x4x4, 8x8, or 64 pre-windowed tokens), head_dim 32.
// A work item is a PAIR of consecutive windows x one head: 128 query rows = the 128 TMEM
// lanes.  A CTA serves one head (its 64x64 bias table lives in shared memory) and strides
// over window pairs.  352 threads, one CTA per SM, warp-specialised:
//
//   warp 8    TMA producer   q/k/v tiles of the pair, gathered straight out of the un-windowed,
//                            un-shifted (B,D,H,W,3C) tensor with 5-D tensor maps: the cyclic shift
//                            is a coordinate offset, a window that wraps around the volume edge is
//                            fetched as its 2/4/8 contiguous pieces (tc_window.cuh).
//   warp 9    MMA issuer     S = Q K^T  (M128 N128 K32; the block diagonal is the two windows),
//                            O = P V    (M128 N32 K128); accumulators in TMEM, S(n+1) issued as
//                            soon as S(n) has been read out, O double-buffered.
//   warps 0-7 softmax        two threads per query row (32 keys each): tcgen05.ld the logits,
//                            cosine normalisation x logit scale (or q scale), relative position
//                            bias, shift mask from region ids, exp2 softmax in fp32, P (bf16) into
//                            the 128B-swizzled K-major tile of the second MMA; one pair later,
//                            O / l -> bf16 -> 64B-swizzled staging tile.
//   warp 10   TMA store      staging tile -> global through the same boxes
//                            (= window_reverse + roll back), double-buffered.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "tc_window.cuh"

namespace mmn { namespace tc {

constexpr int kStages = 3;
constexpr int kTile = 128 * 64;   // bytes of one operand tile: 2 windows x 64 rows x 64 B
constexpr int kWinBytes = 64 * 64;
constexpr int kSoftmaxThreads = 256;
constexpr int kProducerWarp = 8, kMmaWarp = 9, kStoreWarp = 10;
constexpr int kFwdThreads = 352;
constexpr int kBiasLd = 68;       // padded row of the shared bias table (conflict-free float4 rows)
constexpr int kTmemCols = 256;    // S: columns [0,128), O double-buffered: [128,160) and [160,192)
constexpr float kLog2e = 1.4426950408889634f;

struct FwdParams {
  CUtensorMap q[8], k[8], v[8], o[8];   // one box shape per wrap class
  WinShape S;
  int nH, n_pairs;
  int cosine, mask_kind, mask_windows;
  float scale;
  const float* bias;
  const float* head_scale;
  const float* mask;
  float* lse;
  long long* trace;   // debug: per-phase clock64 stamps of CTA 0 (MMN_TC_TRACE=<file>), else null
};

// trace[(role*32 + item)*16 + ev]; roles: 0 softmax warp 0, 1 softmax warp 7, 2 producer, 3 MMA, 4 store
__device__ __forceinline__ void trace_ev(const FwdParams& P, int role, int item, int ev) {
  if (P.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && item < 32) P.trace[(role * 32 + item) * 16 + ev] = clock64();
}

// COS: cosine attention (else scaled dot product); MASK: MMN_MASK_NONE / _SHIFT / _TENSOR.  Compile-time so that
// each variant carries only its own code (the kernel is instruction-cache sensitive).
template <bool COS, int MASK>
__global__ void __launch_bounds__(kFwdThreads, 1)
winattn_fwd_tc_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024-B alignment
  uint8_t* sQKV = smem;                                  // kStages x 3 x kTile
  uint8_t* sP = sQKV + kStages * 3 * kTile;              // 2 x 16 KB (key halves)
  uint8_t* sO = sP + 2 * 16384;                          // 2 x 8 KB output staging
  float* sBias = reinterpret_cast<float*>(sO + 2 * kTile);   // [64][kBiasLd] fp32: this CTA's head
  float* sRq = sBias + kN * kBiasLd;                     // 128: per-row logit multiplier
  float* sRk = sRq + 128;                                // 128: per-key 1/||k||
  float* sMax = sRk + 128;                               // [2][128] partial row maxima
  float* sSum = sMax + 256;                              // [2 pairs in flight][2][128] partial row sums
  int* sRid = reinterpret_cast<int*>(sSum + 512);        // 128 region ids
  uint8_t* sPos = reinterpret_cast<uint8_t*>(sRid + 128); // [8 wrap classes][64]: tile row -> window position
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPos + 512);
  uint64_t* full = bars;                                 // [kStages]  TMA -> MMA, softmax
  uint64_t* empty = bars + kStages;                      // [kStages]  MMA -> TMA
  uint64_t* s_full = bars + 2 * kStages;                 // S in TMEM
  uint64_t* s_empty = s_full + 1;                        // S read out (256 arrivals)
  uint64_t* p_full = s_full + 2;                         // P in smem (256 arrivals)
  uint64_t* o_full = s_full + 3;                         // O in TMEM / P consumed
  uint64_t* so_ready = s_full + 4;                       // [2] staging tile written (256 arrivals)
  uint64_t* so_free = s_full + 6;                        // [2] staging tile drained by the store warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);

  const WinShape& S = P.S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % P.nH;
  const int pair0 = blockIdx.x / P.nH, pair_step = gridDim.x / P.nH;

  // ---- one-time setup
  for (int i = tid; i < 2 * 16384 / 16; i += kFwdThreads) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
  if (P.bias)
    for (int i = tid; i < kN * kN; i += kFwdThreads) sBias[(i >> 6) * kBiasLd + (i & 63)] = __ldg(P.bias + (size_t)h * kN * kN + i);
  for (int i = tid; i < 512; i += kFwdThreads) sPos[i] = (uint8_t)piece_position(S, i >> 6, i & 63);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, kSoftmaxThreads); mbar_init(p_full, kSoftmaxThreads); mbar_init(o_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&so_ready[s], kSoftmaxThreads); mbar_init(&so_free[s], 1); }
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0)
    for (int i = 0; i < 8; ++i) { tma_prefetch_desc(&P.q[i]); tma_prefetch_desc(&P.k[i]); tma_prefetch_desc(&P.v[i]); }
  if (warp == kStoreWarp && lane == 0)
    for (int i = 0; i < 8; ++i) tma_prefetch_desc(&P.o[i]);
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  fence_proxy_async_smem();            // zeroed P must be visible to the tensor-core (async) proxy
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  // window cursors: slot 0 / slot 1 of the current pair, stepped by 2*pair_step windows
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
        const int stage = it % kStages, phase = (it / kStages) & 1;
        trace_ev(P, 2, it, 0);
        mbar_wait(&empty[stage], phase ^ 1);
        trace_ev(P, 2, it, 1);
        mbar_arrive_expect_tx(&full[stage], 3 * kTile);
        uint8_t* base = sQKV + stage * 3 * kTile;
        WinCursor c = c0;
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
          const WinGeom g = window_geom(S, c);
          issue_window_boxes<true>(S, P.q, g, h * kD, base + slot * kWinBytes, &full[stage]);
          issue_window_boxes<true>(S, P.k, g, h * kD, base + kTile + slot * kWinBytes, &full[stage]);
          issue_window_boxes<true>(S, P.v, g, h * kD, base + 2 * kTile + slot * kWinBytes, &full[stage]);
          c.advance(S, one);
        }
        trace_ev(P, 2, it, 2);
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idescS = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idescO = umma_idesc_bf16(128, 32, 0, 1);    // P (K-major) x V (MN-major)
    const uint32_t tS = tmem;
    const uint32_t pAddr = smem_u32(sP);
    auto issue_s = [&](int n) {
      const int stage = n % kStages, phase = (n / kStages) & 1;
      const uint32_t qAddr = smem_u32(sQKV + stage * 3 * kTile), kAddr = qAddr + kTile;
      mbar_wait(&full[stage], phase);
      mbar_wait(s_empty, (n & 1) ^ 1);
      tcgen05_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_bf16_ss(tS, umma_smem_desc(qAddr + ks * 32, 0, 512, kSwz64), umma_smem_desc(kAddr + ks * 32, 0, 512, kSwz64),
                       idescS, ks);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    int it = 0;
    if (pair0 < P.n_pairs) issue_s(0);
    for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it) {
      const int stage = it % kStages;
      const uint32_t vAddr = smem_u32(sQKV + stage * 3 * kTile) + 2 * kTile;
      trace_ev(P, 3, it, 0);
      if (pair + pair_step < P.n_pairs) issue_s(it + 1);
      trace_ev(P, 3, it, 1);
      mbar_wait(p_full, it & 1);
      trace_ev(P, 3, it, 2);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t tO = tmem + 128 + (it & 1) * 32;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16_ss(tO, umma_smem_desc(pAddr + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024, kSwz128),
                       umma_smem_desc(vAddr + ks * 1024, 8192, 512, kSwz64), idescO, ks);
        umma_commit(o_full);
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      trace_ev(P, 3, it, 3);
    }
  } else if (warp == kStoreWarp) {
    // ============================== TMA store ==============================
    if (lane == 0) {
      WinCursor c0;
      c0.init(S, 2 * pair0);
      int it = 0;
      for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it, c0.advance(S, step)) {
        const int buf = it & 1;
        trace_ev(P, 4, it, 0);
        mbar_wait(&so_ready[buf], (it >> 1) & 1);
        trace_ev(P, 4, it, 1);
        WinCursor c = c0;
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
          issue_window_boxes<false>(S, P.o, window_geom(S, c), h * kD, sO + buf * kTile + slot * kWinBytes, nullptr);
          c.advance(S, one);
        }
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&so_free[buf]);
        trace_ev(P, 4, it, 2);
      }
      tma_store_wait_all<0>();
    }
  } else {
    // ============================== softmax / epilogue (256 threads: 2 per query row) ==============================
    const int r = tid & 127, half = tid >> 7;            // row of the pair tile; which 32 of its 64 keys
    const int slot = r >> 6, i = r & 63;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float hscale = COS ? __ldg(P.head_scale + h) : 1.f;
    const int trole = warp == 0 ? 0 : (warp == 7 ? 1 : -1);
#define TR(item, ev) do { if (trole >= 0) trace_ev(P, trole, item, ev); } while (0)

    // O epilogue of iteration `ie`, deferred behind the next pair's softmax so that the PV MMA
    // runs under useful work: half 0 takes output channels [0,16), half 1 [16,32).
    auto epilogue = [&](int ie, float mx_e, long long lse_index) {
      const int buf = ie & 1;
      uint32_t oraw[16];
      tmem_ld_32x32b_x16(tmem + lane_base + 128 + buf * 32 + half * 16, oraw);
      tmem_ld_wait();
      tcgen05_fence_before();
      const float* sums = sSum + buf * 256;
      const float l = sums[r] + sums[128 + r];
      if (half == 0) P.lse[lse_index] = mx_e + __logf(l);
      const float inv = __frcp_rn(l);
      TR(ie + 1, 10);
      mbar_wait(&so_free[buf], ((ie >> 1) & 1) ^ 1);  // the store warp has drained this staging buffer
      TR(ie + 1, 11);
      uint8_t* orow = sO + buf * kTile + r * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint4 v4 = make_uint4(pack_bf16x2(__uint_as_float(oraw[c * 8 + 0]) * inv, __uint_as_float(oraw[c * 8 + 1]) * inv),
                              pack_bf16x2(__uint_as_float(oraw[c * 8 + 2]) * inv, __uint_as_float(oraw[c * 8 + 3]) * inv),
                              pack_bf16x2(__uint_as_float(oraw[c * 8 + 4]) * inv, __uint_as_float(oraw[c * 8 + 5]) * inv),
                              pack_bf16x2(__uint_as_float(oraw[c * 8 + 6]) * inv, __uint_as_float(oraw[c * 8 + 7]) * inv));
        *reinterpret_cast<uint4*>(orow + (((half * 2 + c) ^ ((r >> 1) & 3)) << 4)) = v4;
      }
      fence_proxy_async_smem();
      mbar_arrive(&so_ready[buf]);
      TR(ie + 1, 12);
    };

    WinCursor cur;
    cur.init(S, 2 * pair0 + slot);
    int it = 0;
    bool have_prev = false;
    float prev_mx = 0.f;
    long long prev_lse = 0;
    for (int pair = pair0; pair < P.n_pairs; pair += pair_step, ++it, cur.advance(S, step)) {
      const int stage = it % kStages, phase = (it / kStages) & 1;
      const int w = pair * 2 + slot;
      const WinGeom g = window_geom(S, cur);
      const bool masked = (MASK == MMN_MASK_SHIFT) && g.cls != 0;  // uniform over the threads of a window
      const bool permuted = (MASK == MMN_MASK_SHIFT) && (g.cls & 6) != 0;         // splitting the slowest axis only keeps window order
      const uint8_t* pos = sPos + g.cls * 64;
      const int ipos = permuted ? pos[i] : i;         // window position of this thread's query row
      const uint8_t* base = sQKV + stage * 3 * kTile;

      TR(it, 0);
      mbar_wait(&full[stage], phase);
      TR(it, 1);
      if (COS) {
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
        const float inv = rsqrtf(fmaxf(ss, 1e-24f));     // 1 / max(||x||, 1e-12)
        if (half == 0) sRq[r] = inv * hscale; else sRk[r] = inv;
      }
      int rid_i = 0;
      if (masked) { rid_i = region_id(S, g, ipos); if (half == 0) sRid[r] = rid_i; }
      named_bar_sync(1, kSoftmaxThreads);
      TR(it, 2);

      // ---- logits of this thread's 32 keys
      mbar_wait(s_full, it & 1);
      tcgen05_fence_after();
      TR(it, 3);
      uint32_t raw[32];
      tmem_ld_32x32b_x32(tmem + lane_base + slot * 64 + half * 32, raw);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(s_empty);
      TR(it, 4);

      float s[32];
      const float a_i = COS ? sRq[r] : P.scale;
      const float4* krow = reinterpret_cast<const float4*>(sRk + slot * 64 + half * 32);
      const float* mtile = (MASK == MMN_MASK_TENSOR) ? P.mask + (size_t)(w % P.mask_windows) * kN * kN : nullptr;
      if (!permuted) {
        const float4* brow = reinterpret_cast<const float4*>(sBias + ipos * kBiasLd + half * 32);
        const float4* mrow = mtile ? reinterpret_cast<const float4*>(mtile + ipos * kN + half * 32) : nullptr;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 bb = P.bias ? brow[j4] : make_float4(0.f, 0.f, 0.f, 0.f);
          if (mrow) { float4 mm = __ldg(mrow + j4); bb.x += mm.x; bb.y += mm.y; bb.z += mm.z; bb.w += mm.w; }
          float4 kk = COS ? krow[j4] : make_float4(1.f, 1.f, 1.f, 1.f);
          const float add[4] = {bb.x, bb.y, bb.z, bb.w};
          const float rk[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = j4 * 4 + e;
            s[j] = fmaf(__uint_as_float(raw[j]) * rk[e], a_i, add[e]);
          }
        }
      } else {
        // piece-major tile: key column j of the tile is window position pos[j]
        const uint32_t* pj4 = reinterpret_cast<const uint32_t*>(pos + half * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const uint32_t pk = pj4[j4];
          float4 kk = COS ? krow[j4] : make_float4(1.f, 1.f, 1.f, 1.f);
          const float rk[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = j4 * 4 + e, jp = (pk >> (8 * e)) & 0xff;
            float add = P.bias ? sBias[ipos * kBiasLd + jp] : 0.f;
            if (mtile) add += __ldg(mtile + ipos * kN + jp);
            s[j] = fmaf(__uint_as_float(raw[j]) * rk[e], a_i, add);
          }
        }
      }
      if (masked) {
        const int* rids = sRid + slot * 64 + half * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (rids[j] != rid_i) s[j] -= 100.f;
      }
      float mx = s[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) mx = fmaxf(mx, s[j]);
      sMax[half * 128 + r] = mx;
      TR(it, 5);
      named_bar_sync(2, kSoftmaxThreads);
      TR(it, 6);
      mx = fmaxf(mx, sMax[(half ^ 1) * 128 + r]);
      float l = 0.f;
      const float mneg = -mx * kLog2e;
#pragma unroll
      for (int j = 0; j < 32; ++j) { s[j] = fast_exp2(fmaf(s[j], kLog2e, mneg)); l += s[j]; }
      sSum[(it & 1) * 256 + half * 128 + r] = l;

      // ---- P (bf16) into the 128B-swizzled K-major tile of this window's key half.  The previous
      // pair's PV MMA must have finished reading the tile (it has had this whole softmax to do so).
      TR(it, 7);
      if (have_prev) { mbar_wait(o_full, (it - 1) & 1); tcgen05_fence_after(); }
      TR(it, 8);
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
      TR(it, 9);

      if (have_prev) epilogue(it - 1, prev_mx, prev_lse);
      TR(it, 13);
      have_prev = true;
      prev_mx = mx;
      prev_lse = ((long long)w * P.nH + h) * kN + ipos;
    }
    if (have_prev) {
      mbar_wait(o_full, (it - 1) & 1);
      tcgen05_fence_after();
      epilogue(it - 1, prev_mx, prev_lse);
    }
#undef TR
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem);
}

constexpr size_t kFwdSmemBytes = 1024 /*align slack*/ + kStages * 3 * kTile + 2 * 16384 + 2 * kTile + kN * kBiasLd * 4 +
                                 (128 + 128 + 256 + 512 + 128) * 4 + 512 + 16 * 8;

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
inline const char* fwd_why_not_impl(const mmn_winattn_desc* d) {
  if (d->io_dtype != MMN_DT_BF16) return "io dtype is not bf16";
  if (d->head_dim != kD) return "head_dim != 32";
  if (d->dropout_p > 0.f) return "attention dropout is only implemented in the generic path";
  WinShape g = shape_from(d);
  if (g.win[0] * g.win[1] * g.win[2] != kN) return "window does not hold 64 tokens";
  for (int a = 0; a < 3; ++a)
    if (g.shift[a] != 0 && (2 * g.shift[a] != g.win[a])) return "shift is neither 0 nor window/2";
  if (g.n_windows % 2 != 0) return "odd number of windows";
  if (d->q_row_stride % 8 || d->k_row_stride % 8 || d->v_row_stride % 8 || d->o_row_stride % 8) return "row stride not 16-byte aligned";
  if (!encode_fn()) return "cuTensorMapEncodeTiled unavailable";
  return nullptr;
}

inline int winattn_fwd_launch(const mmn_winattn_desc* d, const void* q, const void* k, const void* v, const float* bias,
                              const float* head_scale, const float* mask, void* out, float* lse, cudaStream_t st, char* err,
                              size_t errlen) {
  FwdParams P;
  P.S = shape_from(d);
  const int C = d->num_heads * d->head_dim;
  if (!make_window_maps(P.q, q, d->q_row_stride, d->batch, C, P.S) || !make_window_maps(P.k, k, d->k_row_stride, d->batch, C, P.S) ||
      !make_window_maps(P.v, v, d->v_row_stride, d->batch, C, P.S) || !make_window_maps(P.o, out, d->o_row_stride, d->batch, C, P.S)) {
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
  P.trace = nullptr;
  const char* trace_path = getenv("MMN_TC_TRACE");
  if (trace_path && *trace_path) {
    cudaMalloc(&P.trace, 5 * 32 * 16 * sizeof(long long));
    cudaMemsetAsync(P.trace, 0, 5 * 32 * 16 * sizeof(long long), st);
  }

  using Kern = void (*)(const FwdParams);
  static const Kern kernels[2][3] = {
      {winattn_fwd_tc_kernel<false, MMN_MASK_NONE>, winattn_fwd_tc_kernel<false, MMN_MASK_SHIFT>, winattn_fwd_tc_kernel<false, MMN_MASK_TENSOR>},
      {winattn_fwd_tc_kernel<true, MMN_MASK_NONE>, winattn_fwd_tc_kernel<true, MMN_MASK_SHIFT>, winattn_fwd_tc_kernel<true, MMN_MASK_TENSOR>}};
  static std::once_flag once;
  static int num_sms = 148;
  std::call_once(once, [] {
    for (int c = 0; c < 2; ++c)
      for (int m = 0; m < 3; ++m) cudaFuncSetAttribute(kernels[c][m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemBytes);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
  const Kern kern = kernels[P.cosine ? 1 : 0][P.mask_kind];
  int per_head = num_sms / P.nH;                 // CTAs per head (each CTA keeps one head's bias table resident)
  if (per_head < 1) per_head = 1;
  if (per_head > P.n_pairs) per_head = P.n_pairs;
  const int grid = per_head * P.nH;
  kern<<<grid, kFwdThreads, kFwdSmemBytes, st>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "winattn_fwd_tc_kernel: %s", cudaGetErrorString(e));
    return MMN_ERR_CUDA;
  }
  if (P.trace) {      // debug only: synchronous dump of CTA 0's timeline
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
