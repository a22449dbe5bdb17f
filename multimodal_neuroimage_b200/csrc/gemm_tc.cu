// gemm_tc.cu -- the projections around the attention kernels on the Blackwell tensor cores (tcgen05.mma, accumulators in
// TMEM, operands by TMA), for every layer width the path uses (multiples of 32: Swin stages C = 96 ... 1536, MLP 4C):
//
//   forward    y = act(x W^T + b)            F.linear of  swin_v2_module.py:148,176 (qkv, proj), :27-31 (Mlp fc1 -> GELU -> fc2),
//                                            swinfusion_module.py:121,143,221-222,244, crossmodal_transformer.py:158-160 (relu)
//   backward   dx = (dy W) o act'(pre)       the same layers' dgrad; act'(pre) of the layer BELOW was written by ITS forward epilogue
//              dW = dy^T x  (fp32)           wgrad, split over the token dimension, partials summed by a second kernel
//
// ONE persistent kernel, three operand layouts.  A 128 x BN output tile; operands arrive as 64B-swizzled panels of 32
// elements of the contiguous dimension (3-D tensor maps: element-in-panel, row, panel):
//   K-major operand   (reduction dim contiguous: x and W in the forward, dy in dgrad)   stage = [KP panels][rows][64 B]
//   MN-major operand  (output dim contiguous:    W in dgrad, dy^T and x in wgrad)        stage = [rows/32 panels][32 KP][64 B]
// Warp roles: warps 0-15 epilogue, warp 16 TMA producer, warp 17 MMA issuer.  The epilogue warps form TWO groups of eight
// (two per TMEM lane quarter, each half of the tile's columns); group g owns accumulator g, i.e. every other tile of the
// CTA, with its own staging slots, named barriers and TMA-store leader.  At the small-K layers of the path (K = 96: one
// k-block per tile) the epilogue IS the kernel -- TMEM load, bias / GELU arithmetic, swizzled smem stores, store issue --
// and with one group all of it was a serial chain per tile whose waits nothing covered (fc1 + GELU: 50 us at cfg3 where
// the same GEMM without GELU took 22: ALU time simply added to the memory time); two groups overlap one tile's arithmetic
// with the other's loads and stores.  Three pipelines: smem ring (TMA <-> MMA), two TMEM accumulators (MMA <-> epilogue
// groups), static round-robin tile schedule with the n tiles of one row block adjacent (x is read from HBM once, W stays
// in L2).
#include <cstdio>
#include <mutex>

#include "tc_window.cuh"
#include "winattn_tc.h"

namespace mmn { namespace tc {

enum { kEpiNone = 0, kEpiRelu = 1, kEpiGelu = 2, kEpiMulAux = 3 };

constexpr int kGEpiWarps = 16;               // two groups of eight: two per TMEM lane quarter, each half of the tile's columns
constexpr int kGThreads = 32 * (kGEpiWarps + 2);
constexpr int kGProducerWarp = kGEpiWarps, kGMmaWarp = kGEpiWarps + 1;
constexpr int kGTmemCols = 256;                // two accumulators of BN <= 128 columns

struct GemmParams {
  CUtensorMap a, b, d, d_pre, aux_map;         // aux_map: kEpiMulAux, boxes = output tiles (L2 prefetch only)
  int M, N;                                    // output rows / columns
  int tiles_n, n_tiles, k_blocks, splits, kb_per_split;
  int epi, has_pre;
  const float* bias;                           // [N] fp32 or null
  const __nv_bfloat16* aux;                    // kEpiMulAux: act'(pre) of the layer below, (M, N) with row stride ld_aux
  long long ld_aux;
  float* out32;                                // fp32 output: [splits][M][N]
};

// Exact (erf) GELU and its derivative for the epilogues.  erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the
// bf16 rounding of the result): one MUFU.RCP, one MUFU.EX2 and seven FMAs; exp(-v^2 / 2) is shared between the CDF and the PDF.
// libm's erff + expf cost ~50 instructions per element, which made the 128 x BN epilogue of the K = 96 layers -- not HBM -- the
// bound of fc1 forward and fc2 dgrad (tools/bench_blocks.py: 84 / 104 us against a 29 us memory floor at cfg3).
__device__ __forceinline__ void normal_cdf_pdf(float v, float& cdf, float& pdf) {
  const float z = fabsf(v) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  const float e = fast_exp2(-1.4426950408889634f * z * z);                 // exp(-v^2 / 2)
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.f);
  cdf = fmaf(0.5f, copysignf(erf_abs, v), 0.5f);
  pdf = 0.3989422804014327f * e;
}
template <int BN, int KP, int STAGES, bool A_MN, bool B_MN, bool F32OUT>
__global__ void __launch_bounds__(kGThreads, 1)
gemm_tc_kernel(const __grid_constant__ GemmParams P) {
  constexpr int kABytes = 8192 * KP, kBBytes = BN * 64 * KP;
  constexpr int kPanelOut = 128 * 64;                       // one 32-column panel of the output tile
  constexpr int kOutBytes = F32OUT ? 0 : (BN / 32) * kPanelOut;
  constexpr int kSlots = BN <= 64 ? 4 : 2;                  // staging slots: one per epilogue group, two with a second output
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * kABytes;
  uint8_t* sOut = sB + STAGES * kBBytes;                    // staging slots of [BN/32 panels][128 rows][64 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + kSlots * kOutBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;                   // [2]
  uint64_t* acc_empty = acc_full + 2;                       // [2], one arrival per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_units = P.n_tiles * P.splits;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kGEpiWarps / 2); }
    fence_barrier_init();
  }
  if (warp == kGProducerWarp && lane == 0) { tma_prefetch_desc(&P.a); tma_prefetch_desc(&P.b); if (!F32OUT && P.epi == kEpiMulAux) tma_prefetch_desc(&P.aux_map); }
  if (!F32OUT && (warp == 0 || warp == 4) && lane == 0) { tma_prefetch_desc(&P.d); if (P.has_pre) tma_prefetch_desc(&P.d_pre); }
  if (warp == kGMmaWarp) tmem_alloc<kGTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kGProducerWarp) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      int it = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int tile = u / P.splits, sp = u - tile * P.splits;
        const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
        const int m0 = tm * 128, n0 = tn * BN;
        const int kb0 = sp * P.kb_per_split, kb1 = min(kb0 + P.kb_per_split, P.k_blocks);
        // the epilogue multiplies this tile by aux with plain loads: pull that box into L2 now, a pipeline depth ahead
        // (the loads were ~1 us of exposed HBM latency per tile: long-scoreboard stalls 13.6 per issue in the fc2 dgrad)
        if (!F32OUT && P.epi == kEpiMulAux) tma_prefetch_3d(&P.aux_map, 0, m0, n0 >> 5);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], kABytes + kBBytes);
          if (A_MN) tma_load_3d(&P.a, &full[s], sA + s * kABytes, 0, kb * 32 * KP, m0 >> 5);
          else tma_load_3d(&P.a, &full[s], sA + s * kABytes, 0, m0, kb * KP);
          if (B_MN) tma_load_3d(&P.b, &full[s], sB + s * kBBytes, 0, kb * 32 * KP, n0 >> 5);
          else tma_load_3d(&P.b, &full[s], sB + s * kBBytes, 0, n0, kb * KP);
        }
      }
    }
  } else if (warp == kGMmaWarp) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    // MN-major: panels of 32 output-dim elements are 32*KP rows (= 2 KB * KP) apart; K-major: every 16-element step
    // addresses its own panel, so the leading-dimension offset is unused
    const uint64_t dA = umma_smem_desc(0, A_MN ? 2048 * KP : 0, 512, kSwz64);
    const uint64_t dB = umma_smem_desc(0, B_MN ? 2048 * KP : 0, 512, kSwz64);
    const uint32_t a_base = smem_u32(sA) >> 4, b_base = smem_u32(sB) >> 4;
    int it = 0, un = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++un) {
      const int tile = u / P.splits, sp = u - tile * P.splits;
      const int kb0 = sp * P.kb_per_split, kb1 = min(kb0 + P.kb_per_split, P.k_blocks);
      const int as = un & 1;
      mbar_wait(&acc_empty[as], ((un >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
      tcgen05_fence_after();
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a0 = a_base + s * (kABytes >> 4), b0 = b_base + s * (kBBytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 2 * KP; ++ks) {
            const uint32_t ao = A_MN ? ks * 64 : (ks >> 1) * (8192 >> 4) + (ks & 1) * 2;
            const uint32_t bo = B_MN ? ks * 64 : (ks >> 1) * ((BN * 64) >> 4) + (ks & 1) * 2;
            umma_bf16_ss(tmem + as * BN, dA + (a0 + ao), dB + (b0 + bo), idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          if (kb == kb1 - 1) umma_commit(&acc_full[as]);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================== epilogue: warp w -> TMEM lanes 32 (w % 4) .., group (w / 4) % 2, column half w / 8 ==============================
    const int q = warp & 3, grp = (warp >> 2) & 1, cg = warp >> 3;
    const int r = q * 32 + lane;                          // row of the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    constexpr int kChunks = BN / 32;                      // 32-column chunks: column half cg takes chunks cg, cg + 2, ...
    constexpr int kGroupThreads = 32 * kGEpiWarps / 2;
    const int rsw = (r >> 1) & 3;
    const bool leader = warp == 4 * grp && lane == 0;     // issues and retires this group's TMA stores
    const uint32_t bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;
    uint8_t* slot_act = sOut + (P.has_pre ? 2 * grp : grp) * kOutBytes;
    uint8_t* slot_pre = slot_act + kOutBytes;             // has_pre only (BN <= 64: four slots)
    const int as = grp;
    for (int un = grp, u = blockIdx.x + grp * gridDim.x; u < total_units; un += 2, u += 2 * gridDim.x) {
      const int tile = u / P.splits, sp = u - tile * P.splits;
      const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
      const int m0 = tm * 128, n0 = tn * BN;
      mbar_wait(&acc_full[as], (un >> 1) & 1);
      tcgen05_fence_after();
      if (!F32OUT) {
        if (leader) tma_store_wait_read<0>();              // the stores of this group's previous tile have drained its slot(s)
        named_bar_sync(bar_a, kGroupThreads);
      }
      for (int c = cg; c < kChunks; c += 2) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + lane_base + as * BN + c * 32, v);
        tmem_ld_wait();
        if (F32OUT) {
          if (m0 + r < P.M) {
            float4* dst = reinterpret_cast<float4*>(P.out32 + ((size_t)sp * P.M + (m0 + r)) * P.N + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
          }
        } else {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (P.bias) {
            const float4* bp = reinterpret_cast<const float4*>(P.bias + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(bp + j);
              f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
            }
          }
          const int ooff = c * kPanelOut + r * 64;
          if (P.epi == kEpiRelu || P.epi == kEpiGelu) {
            // activation, and (has_pre) its derivative at the pre-activation: what the layer's backward multiplies with --
            // computed here, where exp(-v^2 / 2) is already at hand, so that the dgrad epilogue is one multiply per element
            float d[32];
            if (P.epi == kEpiRelu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { d[j] = f[j] > 0.f ? 1.f : 0.f; f[j] = fmaxf(f[j], 0.f); }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float cdf, pdf;
                normal_cdf_pdf(f[j], cdf, pdf);
                d[j] = fmaf(f[j], pdf, cdf);
                f[j] *= cdf;
              }
            }
            if (P.has_pre) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(slot_pre + ooff + ((j ^ rsw) << 4)) =
                    make_uint4(pack_bf16x2(d[8 * j], d[8 * j + 1]), pack_bf16x2(d[8 * j + 2], d[8 * j + 3]),
                               pack_bf16x2(d[8 * j + 4], d[8 * j + 5]), pack_bf16x2(d[8 * j + 6], d[8 * j + 7]));
            }
          } else if (P.epi == kEpiMulAux) {
            uint32_t x[16];
            if (m0 + r < P.M) {
              const uint4* ap = reinterpret_cast<const uint4*>(P.aux + (size_t)(m0 + r) * P.ld_aux + n0 + c * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j) { const uint4 t = __ldg(ap + j); x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w; }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) x[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              f[2 * j] *= __uint_as_float(x[j] << 16);
              f[2 * j + 1] *= __uint_as_float(x[j] & 0xffff0000u);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(slot_act + ooff + ((j ^ rsw) << 4)) =
                make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                           pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
        }
      }
      tcgen05_fence_before();
      mbar_arrive_warp(&acc_empty[as]);
      if (!F32OUT) {
        fence_proxy_async_smem();
        named_bar_sync(bar_b, kGroupThreads);
        if (leader) {
          tma_store_3d(&P.d, slot_act, 0, m0, n0 >> 5);
          if (P.has_pre) tma_store_3d(&P.d_pre, slot_pre, 0, m0, n0 >> 5);
          tma_store_commit();
        }
      }
    }
    if (!F32OUT && leader) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kGMmaWarp) tmem_dealloc<kGTmemCols>(tmem);
}

// out[e] = sum over splits of part[s][e].  Block = 128 elements (one float4 per lane) x 32 slices of the split range (pure
// latency: a thread issues at most ceil(splits / 32) independent loads); fixed summation order.
constexpr int kGRedSlices = 32;
__global__ void __launch_bounds__(kGRedSlices * 32)
gemm_reduce_kernel(const float* __restrict__ part, int splits, long long n, float* __restrict__ out) {
  __shared__ float4 red[kGRedSlices][32];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const long long e = ((long long)blockIdx.x * 32 + lane) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e < n) {
#pragma unroll 4
    for (int s = sl; s < splits; s += kGRedSlices) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)s * n + e));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[sl][lane] = acc;
  __syncthreads();
  if (sl < 4) {                                                // slices 8 sl .. 8 sl + 7
    acc = red[8 * sl][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 v = red[8 * sl + k][lane]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    red[8 * sl][lane] = acc;
  }
  __syncthreads();
  if (sl == 0 && e < n) {
#pragma unroll
    for (int k = 1; k < 4; ++k) { const float4 v = red[8 * k][lane]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    *reinterpret_cast<float4*>(out + e) = acc;
  }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
// (32, rows, cols / 32) view of a row-major bf16 matrix with leading dimension ld; box = (32, box_rows, box_panels)
static bool panel_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_rows, int box_panels) {
  EncodeTiledFn enc = encode_fn();
  if (!enc || reinterpret_cast<uintptr_t>(ptr) % 16 || ld % 8 || cols % 32 || rows < 1) return false;
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 64};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_panels};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int pick_bn(long long n) { return n % 128 == 0 ? 128 : n % 96 == 0 ? 96 : n % 64 == 0 ? 64 : 32; }
// panels of the reduction dimension per stage: 2 (64 elements) unless that leaves a ragged tail and 3 does not
static int pick_kp(long long k) { const long long p = k / 32; return p % 2 == 0 ? 2 : p % 3 == 0 ? 3 : 1; }

template <int BN, int KP, bool A_MN, bool B_MN, bool F32OUT>
struct GemmCfg {
  static constexpr int kStageBytes = 8192 * KP + BN * 64 * KP;
  static constexpr int kOut = F32OUT ? 0 : (BN <= 64 ? 4 : 2) * (BN / 32) * 128 * 64;
  static constexpr int kStages = (200 * 1024 - kOut) / kStageBytes >= 6 ? 6 : (200 * 1024 - kOut) / kStageBytes;
  static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + kOut + (2 * kStages + 4) * 8 + 16;
  static int launch(const GemmParams& P, int grid, cudaStream_t st) {
    auto kern = gemm_tc_kernel<BN, KP, kStages, A_MN, B_MN, F32OUT>;
    static std::once_flag once;
    std::call_once(once, [&] { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem); });
    kern<<<grid, kGThreads, kSmem, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
};

template <bool A_MN, bool B_MN, bool F32OUT>
static int launch_gemm(int bn, int kp, const GemmParams& P, int grid, cudaStream_t st) {
#define MMN_G(BN_, KP_) if (bn == BN_ && kp == KP_) return GemmCfg<BN_, KP_, A_MN, B_MN, F32OUT>::launch(P, grid, st);
  MMN_G(128, 2) MMN_G(128, 3) MMN_G(128, 1) MMN_G(96, 2) MMN_G(96, 3) MMN_G(96, 1)
  MMN_G(64, 2) MMN_G(64, 3) MMN_G(64, 1) MMN_G(32, 2) MMN_G(32, 3) MMN_G(32, 1)
#undef MMN_G
  return 2;
}

const char* linear_why_not(int io_dtype, long long rows, int in_features, int out_features, long long ld_x, long long ld_y) {
  if (io_dtype != MMN_DT_BF16) return "io dtype is not bf16";
  if (rows < 1) return "no rows";
  if (in_features < 32 || in_features % 32 || out_features < 32 || out_features % 32) return "feature counts are not multiples of 32";
  if (ld_x % 8 || ld_y % 8) return "leading dimension not 16-byte aligned";
  if (rows >= (1ll << 31) - 256) return "too many rows";
  if (!encode_fn()) return "cuTensorMapEncodeTiled unavailable";
  return nullptr;
}

static void fill_schedule(GemmParams& P, long long M, int N, int bn, long long k_blocks, int splits_wanted) {
  P.M = (int)M; P.N = N;
  P.tiles_n = N / bn;
  P.n_tiles = (int)((M + 127) / 128) * P.tiles_n;
  P.k_blocks = (int)k_blocks;
  int splits = splits_wanted < 1 ? 1 : splits_wanted;
  if (splits > k_blocks) splits = (int)k_blocks;
  P.kb_per_split = (int)((k_blocks + splits - 1) / splits);
  P.splits = (int)((k_blocks + P.kb_per_split - 1) / P.kb_per_split);
}

// y (rows, out) = act(x (rows, in) w^T (out, in) + bias); y_pre (optional) = the value before the activation
int linear_fwd(const void* x, const void* w, const float* bias, void* y, void* y_pre, int act, long long rows, int in_features,
               int out_features, long long ld_x, long long ld_y, cudaStream_t st, char* err, size_t errlen, int* launches) {
  // a second output (act') takes two staging slots per epilogue group: tiles of at most 64 columns
  const int bn = y_pre ? (out_features % 64 == 0 ? 64 : 32) : pick_bn(out_features), kp = pick_kp(in_features);
  GemmParams P{};
  if (!panel_map(&P.a, x, rows, in_features, ld_x, 128, kp) || !panel_map(&P.b, w, out_features, in_features, in_features, bn, kp) ||
      !panel_map(&P.d, y, rows, out_features, ld_y, 128, bn / 32) ||
      (y_pre && !panel_map(&P.d_pre, y_pre, rows, out_features, ld_y, 128, bn / 32))) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (pointer alignment or strides)");
    return MMN_ERR_CUDA;
  }
  fill_schedule(P, rows, out_features, bn, (in_features / 32 + kp - 1) / kp, 1);
  P.epi = act; P.has_pre = y_pre ? 1 : 0; P.bias = bias;
  int grid = num_sms_cached();
  if (grid > P.n_tiles) grid = P.n_tiles;
  const int rc = launch_gemm<false, false, false>(bn, kp, P, grid, st);
  if (rc) { snprintf(err, errlen, "gemm_tc_kernel (forward): %s", rc == 2 ? "no tile configuration" : cudaGetErrorString(cudaGetLastError())); return MMN_ERR_CUDA; }
  ++*launches;
  return MMN_OK;
}

size_t linear_wgrad_workspace_bytes(long long rows, int in_features, int out_features) {
  const int bn = pick_bn(in_features);
  const long long tiles = ((out_features + 127) / 128) * (long long)(in_features / bn);
  long long splits = num_sms_cached() / tiles;             // one unit per SM (one round, no ragged second one)
  const long long kb = (rows + 63) / 64;
  if (splits > kb) splits = kb;
  if (splits < 1) splits = 1;
  return splits <= 1 ? 16 : (size_t)splits * out_features * in_features * sizeof(float);
}

// dx (rows, in) = (dy (rows, out) w (out, in)) o act'(aux);  dw (out, in) fp32 = dy^T x.  Either half may be skipped (null).
int linear_bwd_general(const void* dy, const void* x, const void* w, void* dx, float* dw, float* workspace, const void* aux, long long ld_aux,
                       int act_grad, long long rows, int in_features, int out_features, long long ld_dy, long long ld_x, long long ld_dx,
                       cudaStream_t st, char* err, size_t errlen, int* launches) {
  if (dx) {
    // dgrad: A = dy (K-major over out), B = w (out rows, in contiguous: MN-major), D = dx
    const int bn = pick_bn(in_features), kp = pick_kp(out_features);
    GemmParams P{};
    if (!panel_map(&P.a, dy, rows, out_features, ld_dy, 128, kp) || !panel_map(&P.b, w, out_features, in_features, in_features, 32 * kp, bn / 32) ||
        !panel_map(&P.d, dx, rows, in_features, ld_dx, 128, bn / 32) ||
        (act_grad && !panel_map(&P.aux_map, aux, rows, in_features, ld_aux, 128, bn / 32))) {
      snprintf(err, errlen, "cuTensorMapEncodeTiled failed (pointer alignment or strides)");
      return MMN_ERR_CUDA;
    }
    fill_schedule(P, rows, in_features, bn, (out_features / 32 + kp - 1) / kp, 1);
    P.epi = act_grad; P.aux = static_cast<const __nv_bfloat16*>(aux); P.ld_aux = ld_aux;
    int grid = num_sms_cached();
    if (grid > P.n_tiles) grid = P.n_tiles;
    const int rc = launch_gemm<false, true, false>(bn, kp, P, grid, st);
    if (rc) { snprintf(err, errlen, "gemm_tc_kernel (dgrad): %s", rc == 2 ? "no tile configuration" : cudaGetErrorString(cudaGetLastError())); return MMN_ERR_CUDA; }
    ++*launches;
  }
  if (dw) {
    // wgrad: A = dy^T (M = out, contiguous in dy: MN-major), B = x (N = in, contiguous: MN-major), reduction = tokens
    const int bn = pick_bn(in_features), kp = 2;
    GemmParams P{};
    if (!panel_map(&P.a, dy, rows, out_features, ld_dy, 32 * kp, 4) || !panel_map(&P.b, x, rows, in_features, ld_x, 32 * kp, bn / 32)) {
      snprintf(err, errlen, "cuTensorMapEncodeTiled failed (pointer alignment or strides)");
      return MMN_ERR_CUDA;
    }
    const long long tiles = ((out_features + 127) / 128) * (long long)(in_features / bn);
    const int want = (int)(num_sms_cached() / tiles);        // one unit per SM: one round of the persistent grid
    fill_schedule(P, out_features, in_features, bn, (rows + 32 * kp - 1) / (32 * kp), want);
    P.out32 = P.splits > 1 ? workspace : dw;
    int grid = num_sms_cached();
    if (grid > P.n_tiles * P.splits) grid = P.n_tiles * P.splits;
    const int rc = launch_gemm<true, true, true>(bn, kp, P, grid, st);
    if (rc) { snprintf(err, errlen, "gemm_tc_kernel (wgrad): %s", rc == 2 ? "no tile configuration" : cudaGetErrorString(cudaGetLastError())); return MMN_ERR_CUDA; }
    ++*launches;
    if (P.splits > 1) {
      const long long n = (long long)out_features * in_features;
      gemm_reduce_kernel<<<(unsigned)((n / 4 + 31) / 32), kGRedSlices * 32, 0, st>>>(workspace, P.splits, n, dw);
      if (cudaGetLastError() != cudaSuccess) { snprintf(err, errlen, "gemm_reduce_kernel launch failed"); return MMN_ERR_CUDA; }
      ++*launches;
    }
  }
  return MMN_OK;
}

}}  // namespace mmn::tc
