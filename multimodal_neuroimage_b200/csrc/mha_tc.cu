// mha_tc.cu -- cross-modal multi-head attention (modules/multihead_attention.py:85-127: q*scaling, bmm, + attn_mask, fp32
// softmax, bmm) on the Blackwell tensor cores, flash style: nothing T x S ever reaches HBM.
//
// Tuned shapes: bf16, head_dim 32 or 64, any T / S / batch / heads, masks NONE / FUTURE (generated from indices,
// crossmodal_transformer.py:179-186) / TENSOR ((T,S) additive fp32).  q, k, v, out are addressed in place in the reference's
// (len, batch, embed) layout (row strides from the descriptor: q/k/v may be column slices of one packed projection), through
// 4-D tensor maps (element-in-panel, row, 32-channel panel, batch) whose boxes land as 64B-swizzled [panel][128 rows][64 B]
// tiles -- the K-major operand of Q K^T and, read transposed (MN-major), the operand of P V, dS^T Q, P^T dO and dS K alike.
//
//   forward   CTA = (128 queries, head, batch); K/V tiles of 128 keys stream through a 2-stage ring.  Per key tile:
//             S = Q K^T (tcgen05, TMEM) -> one thread per query row: online softmax in fp32 (two passes over the TMEM tile:
//             max, then exp2 / sum), P (bf16) back into TMEM as the A operand of P V; the tile's P V lands in its own TMEM
//             columns and is folded into the row's output in REGISTERS (O = O * alpha + PV) -- no TMEM rescale pass.
//             Two CTAs per SM (256 TMEM columns, 80 KB of shared memory each) overlap one's softmax with the other's MMAs.
//   backward  two kernels, both recompute S and dP = dO V^T per 128 x 128 block and form P = exp2(s - lse),
//             dS = P o (dP - delta) * scale in registers -> bf16 tiles in shared memory:
//               mode dKdV: CTA = (128 keys, head, batch), query tiles stream;  dV += P^T dO,  dK += dS^T Q   (TMEM)
//               mode dQ  : CTA = (128 queries, head, batch), key tiles stream; dQ += dS K                    (TMEM)
//             delta = rowsum(dO o O) comes from a small pre-pass.  No atomics: deterministic.
#include <cstdio>
#include <mutex>

#include "tc_window.cuh"
#include "winattn_tc.h"

namespace mmn { namespace tc {

constexpr float kMLog2e = 1.4426950408889634f;
constexpr float kMLn2 = 0.6931471805599453f;
constexpr int kMThreads = 192;                 // 4 softmax warps + TMA producer + MMA issuer
constexpr int kMStages = 2;

struct MhaParams {
  CUtensorMap q, k, v, dout;                   // 4-D maps (32, rows, E / 32, batch); box (32, 128, D / 32, 1)
  int T, S, B, nH;
  int mask_kind, mask_diag;
  float scale;
  const float* mask;                           // (T, S) additive, MMN_MASK_TENSOR
  __nv_bfloat16* out; long long o_st, o_sb;    // forward output rows (t, b): out + t * o_st + b * o_sb + h * D
  float* lse;                                  // (B * nH, T) natural-log log-sum-exp
  const float* delta;                          // (B * nH, T) rowsum(dO o O)       (backward)
  __nv_bfloat16 *dq, *dk, *dv;                 // backward outputs
  long long dq_st, dq_sb, dk_st, dk_sb, dv_st, dv_sb;
};

// number of key tiles a query tile starting at t0 can see / first query tile that sees key tile j (FUTURE mask:
// key s is visible to query t iff s - t < diag)
__device__ __forceinline__ int visible_key_tiles(const MhaParams& P, int t0) {
  const int nkt = (P.S + 127) >> 7;
  if (P.mask_kind != MMN_MASK_FUTURE) return nkt;
  const long long last = (long long)t0 + 127 + P.mask_diag - 1;          // largest visible key index of the tile's last row
  if (last < 0) return 0;
  return (int)min((long long)nkt, (last >> 7) + 1);
}
__device__ __forceinline__ int first_query_tile(const MhaParams& P, int s0) {
  if (P.mask_kind != MMN_MASK_FUTURE) return 0;
  const long long t_min = (long long)s0 - P.mask_diag + 1;               // smallest t with s0 - t < diag
  return t_min <= 0 ? 0 : (int)(t_min >> 7);
}

// additive mask term (log2 domain) for logit (t, s); -inf outside the valid / visible range
__device__ __forceinline__ float mask_term(const MhaParams& P, int t, int s) {
  if (s >= P.S) return -INFINITY;
  if (P.mask_kind == MMN_MASK_FUTURE) return (s - t >= P.mask_diag) ? -INFINITY : 0.f;
  if (P.mask_kind == MMN_MASK_TENSOR) return t < P.T ? __ldg(P.mask + (long long)t * P.S + s) * kMLog2e : 0.f;
  return 0.f;
}
__device__ __forceinline__ bool tile_needs_mask(const MhaParams& P, int t0, int s0) {
  if (s0 + 128 > P.S || P.mask_kind == MMN_MASK_TENSOR) return true;
  if (P.mask_kind == MMN_MASK_FUTURE) return (s0 + 127) - t0 >= P.mask_diag;   // some (t, s) of the block is masked
  return false;
}

// ------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kMThreads, 2)
mha_fwd_tc_kernel(const __grid_constant__ MhaParams P) {
  constexpr int kTileB = D * 256;                       // 128 rows x D bf16
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileB;                            // [kMStages]
  uint8_t* sV = sK + kMStages * kTileB;                 // [kMStages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kMStages * kTileB);
  uint64_t* q_full = bars;
  uint64_t* full = bars + 1;                            // [kMStages]
  uint64_t* empty = full + kMStages;                    // [kMStages]
  uint64_t* s_full = empty + kMStages;
  uint64_t* p_ready = s_full + 1;                       // 4 arrivals
  uint64_t* pv_full = s_full + 2;
  uint64_t* pv_empty = s_full + 3;                      // 4 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = visible_key_tiles(P, t0);

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kMStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(pv_full, 1); mbar_init(pv_empty, 4);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&P.q); tma_prefetch_desc(&P.k); tma_prefetch_desc(&P.v); }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tP = tmem + 128, tPV = tmem + 192;

  if (warp == 4) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kTileB);
      tma_load_4d(&P.q, q_full, sQ, 0, t0, h * (D / 32), b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kMStages;
        mbar_wait(&empty[s], ((j / kMStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], 2 * kTileB);
        tma_load_4d(&P.k, &full[s], sK + s * kTileB, 0, j * 128, h * (D / 32), b);
        tma_load_4d(&P.v, &full[s], sV + s * kTileB, 0, j * 128, h * (D / 32), b);
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t idescS = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idescPV = umma_idesc_bf16(128, D, 0, 1);
    const uint64_t dK = umma_smem_desc(0, 0, 512, kSwz64);            // K-major tiles (Q, K)
    const uint64_t dVm = umma_smem_desc(0, 8192, 512, kSwz64);        // V read MN-major: channel panels 8 KB apart
    const uint32_t q0 = smem_u32(sQ) >> 4, k0 = smem_u32(sK) >> 4, v0 = smem_u32(sV) >> 4;
    auto issue_S = [&](int j) {
      const int s = j % kMStages;
      mbar_wait(&full[s], (j / kMStages) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint32_t o = (ks >> 1) * (8192 >> 4) + (ks & 1) * 2;
          umma_bf16_ss(tS, dK + (q0 + o), dK + (k0 + s * (kTileB >> 4) + o), idescS, ks > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (n_tiles > 0) {
      mbar_wait(q_full, 0);
      issue_S(0);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % kMStages;
      mbar_wait(p_ready, j & 1);                    // P(j) is in TMEM, S(j) has been read out
      mbar_wait(pv_empty, (j & 1) ^ 1);             // the rows have folded PV(j - 1) into their outputs
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)              // 16 keys = 8 TMEM columns of P per step
          umma_bf16_ts(tPV, tP + ks * 8, dVm + (v0 + s * (kTileB >> 4) + ks * 64), idescPV, ks > 0 ? 1u : 0u);
        umma_commit(pv_full);
        umma_commit(&empty[s]);
      }
      __syncwarp();
      if (j + 1 < n_tiles) issue_S(j + 1);
    }
  } else {
    // ============================== softmax: thread = query row ==============================
    const int r = tid;                              // 0..127
    const int t = t0 + r;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const float sc = P.scale * kMLog2e;
    float m = -INFINITY, l = 0.f;
    float O[D];
#pragma unroll
    for (int e = 0; e < D; ++e) O[e] = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int s0 = j * 128;
      const bool masked = tile_needs_mask(P, t0, s0);
      mbar_wait(s_full, j & 1);
      tcgen05_fence_after();
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + lane_base + c * 32, v);
        tmem_ld_wait();
        if (masked) {
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, fmaf(__uint_as_float(v[e]), sc, mask_term(P, t, s0 + c * 32 + e)));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]) * sc);
        }
      }
      const float m_new = fmaxf(m, mx);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = fast_exp2(m - m_use);      // m = -inf: 0
      // pass 2: P = exp2(s - m) -> bf16 pairs -> TMEM (A operand of P V), row sum
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32], pk[16];
        tmem_ld_32x32b_x32(tS + lane_base + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float a = fmaf(__uint_as_float(v[e]), sc, -m_use), bb = fmaf(__uint_as_float(v[e + 1]), sc, -m_use);
          if (masked) { a += mask_term(P, t, s0 + c * 32 + e); bb += mask_term(P, t, s0 + c * 32 + e + 1); }
          a = fast_exp2(a); bb = fast_exp2(bb);
          rs += a + bb;
          pk[e >> 1] = pack_bf16x2(a, bb);
        }
        tmem_st_32x32b_x16(tP + lane_base + c * 16, pk);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive_warp(p_ready);
      l = l * alpha + rs;
      m = m_new;
      // fold this tile's P V into the row's output
      mbar_wait(pv_full, j & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tPV + lane_base + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) O[c * 32 + e] = fmaf(O[c * 32 + e], alpha, __uint_as_float(v[e]));
      }
      tcgen05_fence_before();
      mbar_arrive_warp(pv_empty);
    }
    if (t < P.T) {
      const float inv = l > 0.f ? __frcp_rn(l) : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(P.out + (long long)t * P.o_st + (long long)b * P.o_sb + h * D);
#pragma unroll
      for (int e = 0; e < D / 8; ++e)
        dst[e] = make_uint4(pack_bf16x2(O[8 * e] * inv, O[8 * e + 1] * inv), pack_bf16x2(O[8 * e + 2] * inv, O[8 * e + 3] * inv),
                            pack_bf16x2(O[8 * e + 4] * inv, O[8 * e + 5] * inv), pack_bf16x2(O[8 * e + 6] * inv, O[8 * e + 7] * inv));
      P.lse[((long long)b * P.nH + h) * P.T + t] = l > 0.f ? (m + __log2f(l)) * kMLn2 : -INFINITY;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------
// delta[(b, h), t] = sum_e dO[t, b, h D + e] * O[t, b, h D + e]
template <int D>
__global__ void mha_delta_kernel(const __nv_bfloat16* __restrict__ o, long long o_st, long long o_sb, const __nv_bfloat16* __restrict__ dout,
                                 long long do_st, long long do_sb, int T, int B, int nH, float* __restrict__ delta) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // ((b, h), t)
  if (idx >= (long long)B * nH * T) return;
  const int t = (int)(idx % T);
  const int bh = (int)(idx / T), b = bh / nH, h = bh - b * nH;
  const uint4* po = reinterpret_cast<const uint4*>(o + (long long)t * o_st + (long long)b * o_sb + h * D);
  const uint4* pd = reinterpret_cast<const uint4*>(dout + (long long)t * do_st + (long long)b * do_sb + h * D);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < D / 8; ++e) {
    const uint4 a = __ldg(po + e), c = __ldg(pd + e);
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, uc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      s += __uint_as_float(ua[k] << 16) * __uint_as_float(uc[k] << 16) + __uint_as_float(ua[k] & 0xffff0000u) * __uint_as_float(uc[k] & 0xffff0000u);
  }
  delta[idx] = s;
}

// MODE 0: dK, dV (CTA = key tile, query tiles stream).  MODE 1: dQ (CTA = query tile, key tiles stream).
template <int D, int MODE>
__global__ void __launch_bounds__(kMThreads, 1)
mha_bwd_tc_kernel(const __grid_constant__ MhaParams P) {
  constexpr int kTileB = D * 256;
  constexpr int kPB = 128 * 128 * 2;                    // P / dS tile: [4 panels of 32 keys][128 query rows][64 B]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sR0 = smem;                                  // resident: K (mode 0) / Q (mode 1)
  uint8_t* sR1 = sR0 + kTileB;                          // resident: V (mode 0) / dO (mode 1)
  uint8_t* sS0 = sR1 + kTileB;                          // [kMStages] streamed: Q (mode 0) / K (mode 1)
  uint8_t* sS1 = sS0 + kMStages * kTileB;               // [kMStages] streamed: dO (mode 0) / V (mode 1)
  uint8_t* sdS = sS1 + kMStages * kTileB;
  uint8_t* sP = sdS + kPB;                              // mode 0 only
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (MODE == 0 ? kPB : 0));
  uint64_t* r_full = bars;
  uint64_t* full = bars + 1;                            // [kMStages]
  uint64_t* empty = full + kMStages;                    // [kMStages]
  uint64_t* s_full = empty + kMStages;                  // S and dP in TMEM
  uint64_t* ps_ready = s_full + 1;                      // P / dS tiles in shared memory (4 arrivals)
  uint64_t* ps_free = s_full + 2;                       // the gradient MMAs have read them
  uint64_t* acc_done = s_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;     // first key (mode 0) / query (mode 1) of this CTA
  int n_begin, n_end;                                                   // streamed tiles
  if (MODE == 0) { n_begin = first_query_tile(P, o0); n_end = (P.T + 127) >> 7; }
  else { n_begin = 0; n_end = visible_key_tiles(P, o0); }
  const int n_tiles = max(0, n_end - n_begin);

  if (tid == 0) {
    mbar_init(r_full, 1);
    for (int s = 0; s < kMStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(ps_ready, 4); mbar_init(ps_free, 1); mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&P.q); tma_prefetch_desc(&P.k); tma_prefetch_desc(&P.v); tma_prefetch_desc(&P.dout); }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tA0 = tmem + 256, tA1 = tmem + 256 + D;

  if (warp == 4) {
    if (elect_one() && n_tiles > 0) {
      mbar_arrive_expect_tx(r_full, 2 * kTileB);
      tma_load_4d(MODE == 0 ? &P.k : &P.q, r_full, sR0, 0, o0, h * (D / 32), b);
      tma_load_4d(MODE == 0 ? &P.v : &P.dout, r_full, sR1, 0, o0, h * (D / 32), b);
      for (int n = 0; n < n_tiles; ++n) {
        const int s = n % kMStages, row0 = (n_begin + n) * 128;
        mbar_wait(&empty[s], ((n / kMStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], 2 * kTileB);
        tma_load_4d(MODE == 0 ? &P.q : &P.k, &full[s], sS0 + s * kTileB, 0, row0, h * (D / 32), b);
        tma_load_4d(MODE == 0 ? &P.dout : &P.v, &full[s], sS1 + s * kTileB, 0, row0, h * (D / 32), b);
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t idescS = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idescG0 = umma_idesc_bf16(128, D, 1, 1);       // mode 0: A = P / dS transposed, B = dO / Q transposed
    constexpr uint32_t idescG1 = umma_idesc_bf16(128, D, 0, 1);       // mode 1: A = dS, B = K transposed
    const uint64_t dKm = umma_smem_desc(0, 0, 512, kSwz64);           // K-major
    const uint64_t dMn = umma_smem_desc(0, 8192, 512, kSwz64);        // MN-major, 32-wide panels 8 KB (128 rows) apart
    const uint32_t r0 = smem_u32(sR0) >> 4, r1 = smem_u32(sR1) >> 4, s0b = smem_u32(sS0) >> 4, s1b = smem_u32(sS1) >> 4;
    const uint32_t ds0 = smem_u32(sdS) >> 4, p0 = smem_u32(sP) >> 4;
    auto issue_grad = [&](int n) {                                     // gradient MMAs of streamed tile n
      const int s = n % kMStages;
      mbar_wait(ps_ready, n & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t acc = n > 0 ? 1u : 0u;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (MODE == 0) {   // 16 query rows per step
            umma_bf16_ss(tA0, dMn + (p0 + ks * 64), dMn + (s1b + s * (kTileB >> 4) + ks * 64), idescG0, acc | (ks > 0));    // dV += P^T dO
            umma_bf16_ss(tA1, dMn + (ds0 + ks * 64), dMn + (s0b + s * (kTileB >> 4) + ks * 64), idescG0, acc | (ks > 0));   // dK += dS^T Q
          } else {           // 16 keys per step
            umma_bf16_ss(tA0, dKm + (ds0 + (ks >> 1) * (8192 >> 4) + (ks & 1) * 2), dMn + (s0b + s * (kTileB >> 4) + ks * 64), idescG1,
                         acc | (ks > 0));                                                                                   // dQ += dS K
          }
        }
        umma_commit(ps_free);
        umma_commit(&empty[s]);
      }
      __syncwarp();
    };
    if (n_tiles > 0) mbar_wait(r_full, 0);
    for (int n = 0; n < n_tiles; ++n) {
      const int s = n % kMStages;
      mbar_wait(&full[s], (n / kMStages) & 1);
      if (n > 0) issue_grad(n - 1);                                    // also: the rows have read S(n - 1), dP(n - 1)
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t qa = MODE == 0 ? s0b + s * (kTileB >> 4) : r0, kb = MODE == 0 ? r0 : s0b + s * (kTileB >> 4);
        const uint32_t da = MODE == 0 ? s1b + s * (kTileB >> 4) : r1, vb = MODE == 0 ? r1 : s1b + s * (kTileB >> 4);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint32_t o = (ks >> 1) * (8192 >> 4) + (ks & 1) * 2;
          umma_bf16_ss(tS, dKm + (qa + o), dKm + (kb + o), idescS, ks > 0 ? 1u : 0u);       // S = Q K^T
        }
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint32_t o = (ks >> 1) * (8192 >> 4) + (ks & 1) * 2;
          umma_bf16_ss(tdP, dKm + (da + o), dKm + (vb + o), idescS, ks > 0 ? 1u : 0u);      // dP = dO V^T
        }
        umma_commit(s_full);
      }
      __syncwarp();
    }
    if (n_tiles > 0) {
      issue_grad(n_tiles - 1);
      if (elect_one()) umma_commit(acc_done);
      __syncwarp();
    }
  } else {
    // ============================== P / dS: thread = query row of the current block ==============================
    const int r = tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const float sc = P.scale * kMLog2e;
    const long long item = (long long)b * P.nH + h;
    const int rsw = (r >> 1) & 3;
    float lse2 = INFINITY, dl = 0.f;
    if (MODE == 1 && o0 + r < P.T) { lse2 = __ldg(P.lse + item * P.T + o0 + r) * kMLog2e; dl = __ldg(P.delta + item * P.T + o0 + r); }
    for (int n = 0; n < n_tiles; ++n) {
      const int t0 = MODE == 0 ? (n_begin + n) * 128 : o0, s0 = MODE == 0 ? o0 : (n_begin + n) * 128;
      const int t = t0 + r;
      if (MODE == 0) {
        lse2 = INFINITY; dl = 0.f;
        if (t < P.T) { lse2 = __ldg(P.lse + item * P.T + t) * kMLog2e; dl = __ldg(P.delta + item * P.T + t); }
      }
      if (lse2 == -INFINITY) lse2 = INFINITY;       // a fully masked row has P = 0
      const bool masked = tile_needs_mask(P, t0, s0);
      mbar_wait(s_full, n & 1);
      tcgen05_fence_after();
      if (n > 0) mbar_wait(ps_free, (n - 1) & 1);   // the gradient MMAs of the previous block have read the P / dS tiles
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t vs[32], vd[32];
        tmem_ld_32x32b_x32(tS + lane_base + c * 32, vs);
        tmem_ld_32x32b_x32(tdP + lane_base + c * 32, vd);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float a = fmaf(__uint_as_float(vs[e]), sc, -lse2), bb = fmaf(__uint_as_float(vs[e + 1]), sc, -lse2);
          if (masked) { a += mask_term(P, t, s0 + c * 32 + e); bb += mask_term(P, t, s0 + c * 32 + e + 1); }
          a = fast_exp2(a); bb = fast_exp2(bb);
          pp[e >> 1] = pack_bf16x2(a, bb);
          dd[e >> 1] = pack_bf16x2(a * (__uint_as_float(vd[e]) - dl) * P.scale, bb * (__uint_as_float(vd[e + 1]) - dl) * P.scale);
        }
        uint8_t* prow = sdS + c * 8192 + r * 64;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          *reinterpret_cast<uint4*>(prow + ((q4 ^ rsw) << 4)) = make_uint4(dd[4 * q4], dd[4 * q4 + 1], dd[4 * q4 + 2], dd[4 * q4 + 3]);
          if (MODE == 0)
            *reinterpret_cast<uint4*>(prow + (sP - sdS) + ((q4 ^ rsw) << 4)) = make_uint4(pp[4 * q4], pp[4 * q4 + 1], pp[4 * q4 + 2], pp[4 * q4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive_warp(ps_ready);
    }
    // ---- epilogue: the accumulators -> bf16 rows
    const int row = o0 + r;
    const int limit = MODE == 0 ? P.S : P.T;
    if (n_tiles > 0) {
      mbar_wait(acc_done, 0);
      tcgen05_fence_after();
    }
#pragma unroll
    for (int which = 0; which < (MODE == 0 ? 2 : 1); ++which) {
      __nv_bfloat16* base = MODE == 0 ? (which == 0 ? P.dv : P.dk) : P.dq;
      const long long st = MODE == 0 ? (which == 0 ? P.dv_st : P.dk_st) : P.dq_st, sb = MODE == 0 ? (which == 0 ? P.dv_sb : P.dk_sb) : P.dq_sb;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[32];
        if (n_tiles > 0) {
          tmem_ld_32x32b_x32((which == 0 ? tA0 : tA1) + lane_base + c * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
        if (row < limit) {
          uint4* dst = reinterpret_cast<uint4*>(base + (long long)row * st + (long long)b * sb + h * D + c * 32);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            dst[e] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * e]), __uint_as_float(v[8 * e + 1])),
                                pack_bf16x2(__uint_as_float(v[8 * e + 2]), __uint_as_float(v[8 * e + 3])),
                                pack_bf16x2(__uint_as_float(v[8 * e + 4]), __uint_as_float(v[8 * e + 5])),
                                pack_bf16x2(__uint_as_float(v[8 * e + 6]), __uint_as_float(v[8 * e + 7])));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
static bool row_map(CUtensorMap* out, const void* ptr, int len, int batch, int embed, long long st, long long sb, int D) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[4] = {32, (cuuint64_t)len, (cuuint64_t)(embed / 32), (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)st * 2, 64, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {32, 128, (cuuint32_t)(D / 32), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool aligned(const void* p, long long st, long long sb) {
  return reinterpret_cast<uintptr_t>(p) % 16 == 0 && st % 8 == 0 && sb % 8 == 0;
}

const char* mha_why_not(const mmn_mha_desc* d, bool backward) {
  if (d->io_dtype != MMN_DT_BF16) return "io dtype is not bf16";
  if (d->head_dim != 32 && d->head_dim != 64) return "head_dim is not 32 or 64";
  if (d->dropout_p > 0.f) return "attention dropout is only implemented in the generic path";
  if (d->q_stride_t % 8 || d->q_stride_b % 8 || d->k_stride_t % 8 || d->k_stride_b % 8 || d->v_stride_t % 8 || d->v_stride_b % 8 ||
      d->o_stride_t % 8 || d->o_stride_b % 8)
    return "row strides not 16-byte aligned";
  if (backward && (d->do_stride_t % 8 || d->do_stride_b % 8 || d->dq_stride_t % 8 || d->dq_stride_b % 8 || d->dk_stride_t % 8 ||
                   d->dk_stride_b % 8 || d->dv_stride_t % 8 || d->dv_stride_b % 8))
    return "gradient row strides not 16-byte aligned";
  if (d->batch > 65535 || d->num_heads > 65535) return "batch or head count beyond the grid limits";
  if (!encode_fn()) return "cuTensorMapEncodeTiled unavailable";
  return nullptr;
}

static void fill_common(MhaParams& P, const mmn_mha_desc* d, const float* mask) {
  P.T = d->tgt_len; P.S = d->src_len; P.B = d->batch; P.nH = d->num_heads;
  P.mask_kind = d->mask_kind; P.mask_diag = d->mask_diagonal; P.scale = d->scale; P.mask = mask;
}

template <int D>
static int mha_fwd_launch(const MhaParams& P, cudaStream_t st) {
  constexpr size_t smem = 1024 + (size_t)(1 + 2 * kMStages) * D * 256 + 16 * 8 + 16;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(mha_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
  dim3 grid((P.T + 127) / 128, P.nH, P.B);
  mha_fwd_tc_kernel<D><<<grid, kMThreads, smem, st>>>(P);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int mha_fwd(const mmn_mha_desc* d, const void* q, const void* k, const void* v, const float* mask, void* out, float* lse, cudaStream_t st,
            char* err, size_t errlen, int* launches) {
  MhaParams P{};
  fill_common(P, d, mask);
  const int E = d->num_heads * d->head_dim, D = d->head_dim;
  if (!aligned(q, d->q_stride_t, d->q_stride_b) || !aligned(k, d->k_stride_t, d->k_stride_b) || !aligned(v, d->v_stride_t, d->v_stride_b) ||
      !aligned(out, d->o_stride_t, d->o_stride_b) || !row_map(&P.q, q, P.T, P.B, E, d->q_stride_t, d->q_stride_b, D) ||
      !row_map(&P.k, k, P.S, P.B, E, d->k_stride_t, d->k_stride_b, D) || !row_map(&P.v, v, P.S, P.B, E, d->v_stride_t, d->v_stride_b, D)) {
    snprintf(err, errlen, "mha: cuTensorMapEncodeTiled failed (pointer alignment or strides)");
    return MMN_ERR_CUDA;
  }
  P.out = static_cast<__nv_bfloat16*>(out); P.o_st = d->o_stride_t; P.o_sb = d->o_stride_b; P.lse = lse;
  const int rc = D == 32 ? mha_fwd_launch<32>(P, st) : mha_fwd_launch<64>(P, st);
  if (rc) { snprintf(err, errlen, "mha_fwd_tc_kernel: %s", cudaGetErrorString(cudaGetLastError())); return MMN_ERR_CUDA; }
  ++*launches;
  return MMN_OK;
}

template <int D, int MODE>
static int mha_bwd_launch(const MhaParams& P, cudaStream_t st) {
  constexpr size_t smem = 1024 + (size_t)(2 + 2 * kMStages) * D * 256 + (MODE == 0 ? 2 : 1) * 32768 + 16 * 8 + 16;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(mha_bwd_tc_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
  dim3 grid(((MODE == 0 ? P.S : P.T) + 127) / 128, P.nH, P.B);
  mha_bwd_tc_kernel<D, MODE><<<grid, kMThreads, smem, st>>>(P);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int mha_bwd(const mmn_mha_desc* d, const void* q, const void* k, const void* v, const float* mask, const void* out, const float* lse,
            const void* dout, void* dq, void* dk, void* dv, float* workspace, cudaStream_t st, char* err, size_t errlen, int* launches) {
  MhaParams P{};
  fill_common(P, d, mask);
  const int E = d->num_heads * d->head_dim, D = d->head_dim;
  if (!out) { snprintf(err, errlen, "mha backward (tcgen05) needs the forward output"); return MMN_ERR_INVALID; }
  if (!aligned(q, d->q_stride_t, d->q_stride_b) || !aligned(k, d->k_stride_t, d->k_stride_b) || !aligned(v, d->v_stride_t, d->v_stride_b) ||
      !aligned(out, d->o_stride_t, d->o_stride_b) || !aligned(dout, d->do_stride_t, d->do_stride_b) || !aligned(dq, d->dq_stride_t, d->dq_stride_b) ||
      !aligned(dk, d->dk_stride_t, d->dk_stride_b) || !aligned(dv, d->dv_stride_t, d->dv_stride_b) ||
      !row_map(&P.q, q, P.T, P.B, E, d->q_stride_t, d->q_stride_b, D) || !row_map(&P.k, k, P.S, P.B, E, d->k_stride_t, d->k_stride_b, D) ||
      !row_map(&P.v, v, P.S, P.B, E, d->v_stride_t, d->v_stride_b, D) || !row_map(&P.dout, dout, P.T, P.B, E, d->do_stride_t, d->do_stride_b, D)) {
    snprintf(err, errlen, "mha: cuTensorMapEncodeTiled failed (pointer alignment or strides)");
    return MMN_ERR_CUDA;
  }
  P.lse = const_cast<float*>(lse);
  P.delta = workspace;
  P.dq = static_cast<__nv_bfloat16*>(dq); P.dk = static_cast<__nv_bfloat16*>(dk); P.dv = static_cast<__nv_bfloat16*>(dv);
  P.dq_st = d->dq_stride_t; P.dq_sb = d->dq_stride_b; P.dk_st = d->dk_stride_t; P.dk_sb = d->dk_stride_b;
  P.dv_st = d->dv_stride_t; P.dv_sb = d->dv_stride_b;
  const long long n = (long long)P.B * P.nH * P.T;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (D == 32)
    mha_delta_kernel<32><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out), d->o_stride_t, d->o_stride_b,
                                                 static_cast<const __nv_bfloat16*>(dout), d->do_stride_t, d->do_stride_b, P.T, P.B, P.nH, workspace);
  else
    mha_delta_kernel<64><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(out), d->o_stride_t, d->o_stride_b,
                                                 static_cast<const __nv_bfloat16*>(dout), d->do_stride_t, d->do_stride_b, P.T, P.B, P.nH, workspace);
  if (cudaGetLastError() != cudaSuccess) { snprintf(err, errlen, "mha_delta_kernel launch failed"); return MMN_ERR_CUDA; }
  ++*launches;
  int rc = D == 32 ? mha_bwd_launch<32, 0>(P, st) : mha_bwd_launch<64, 0>(P, st);
  if (!rc) { ++*launches; rc = D == 32 ? mha_bwd_launch<32, 1>(P, st) : mha_bwd_launch<64, 1>(P, st); }
  if (rc) { snprintf(err, errlen, "mha_bwd_tc_kernel: %s", cudaGetErrorString(cudaGetLastError())); return MMN_ERR_CUDA; }
  ++*launches;
  return MMN_OK;
}

}}  // namespace mmn::tc
