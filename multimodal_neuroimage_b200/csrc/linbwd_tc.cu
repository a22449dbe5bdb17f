// linbwd_tc.cu -- fused backward of a bf16 projection  y = x W^T + b  on the Blackwell tensor cores:
//
//     dX = dY W            (tokens x in)      tokens x out  .  out x in
//     dW = dY^T X          (out x in, fp32)   accumulated over all tokens
//     db = colsum(dY)      (out, fp32)
//
// in ONE pass over dY and X.  The reference gets these from three separate autograd kernels per
// projection (F.linear backward: swin_v2_module.py:148,176, swinfusion_module.py:121,143,221-222,244), each
// re-reading the (tokens x out) gradient from HBM; at BASELINE cfg2 those GEMMs are pure HBM streaming
// (in = 96, out = 96 / 288), so reading dY once instead of two or three times is the whole game.
//
// Tuned shape: in = 96, out = 96 G (G = 1 .. 4: proj, kv, qkv, Mlp fc1).  One persistent CTA per SM walks 128-token tiles:
//   warp 4   TMA producer: X tile (128 x 96) and the G chunks of the dY tile (128 x 96 each), as 64B-swizzled
//            panels of 32 channels (3-D tensor maps: channel-in-panel, token, panel), into two rings.
//   warp 5   MMA issuer, per chunk c:
//              dX   (+)= dY_c (K-major A)       x  W_c (MN-major B)      M128 N96 K96   -> TMEM cols [384, 480)
//              dW_c^T += [X | 1]^T (MN-major A) x  dY_c (MN-major B)     M128 N96 K128  -> TMEM cols [96 c, 96 c + 96)
//            The X tile carries a constant fourth panel whose first channel is 1, so lane 96 of the dW^T
//            accumulators is colsum(dY) = db for free (lanes 97..127 are zero).  Accumulating the TRANSPOSE keeps all of
//            dW in 96 G <= 384 columns (dW_c as M = out-channel tiles of 128 x 128 took 128 G: G = 3 at most).
//   warps 0-3  epilogue: dX accumulators -> bf16 -> staging panels -> TMA store; at the end the dW^T / db
//            accumulators -> this CTA's slice of the workspace.  A second tiny kernel sums the per-CTA
//            partials (deterministic, no atomics) and transposes.
#include <cstdio>
#include <mutex>

#include "tc_window.cuh"
#include "winattn_tc.h"

namespace mmn { namespace tc {

constexpr int kLC = 96;                        // in-features of the tuned path (3 panels of 32)
constexpr int kLTile = 128;                    // tokens per tile
constexpr int kPanel = kLTile * 64;            // 8 KB: 128 tokens x 32 channels bf16
constexpr int kLTileBytes = 3 * kPanel;        // 24 KB
constexpr int kXSlot = 4 * kPanel;             // X tile + the ones panel
constexpr int kXStages = 2;
constexpr int kMaxDyStages = 3;              // 3 dY stages for G <= 3, 2 for G = 4 (its four weight chunks take the room)
constexpr int kWChunk = 3 * 96 * 64;           // one 96 x 96 weight chunk as 3 panels of [96 out][32 in]
constexpr int kLThreads = 192;                 // 4 epilogue warps + producer + MMA
constexpr int kLTmemCols = 512;                // dW_c^T at 96 c; dX at 384

struct LinBwdParams {
  CUtensorMap dy, x, w, dx;                    // 3-D maps (32, rows, panels), 64B swizzle
  int M, G, n_tiles;
  float* ws;                                   // [gridDim.x][97][G * 96] partial dW^T (rows 0..95 = in-channel) | db (row 96)
};

__global__ void __launch_bounds__(kLThreads, 1)
linbwd_tc_kernel(const __grid_constant__ LinBwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int G = P.G;
  const int kDyStages = G == 4 ? 2 : 3;
  uint8_t* sDY = smem;                                   // kDyStages x 24 KB
  uint8_t* sX = sDY + kDyStages * kLTileBytes;           // kXStages x 32 KB (X tile + ones panel)
  uint8_t* sW = sX + kXStages * kXSlot;                  // G x 18 KB
  uint8_t* sOut = sW + G * kWChunk;                      // 24 KB staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + kLTileBytes);
  uint64_t* x_full = bars;                               // [kXStages]
  uint64_t* x_empty = bars + kXStages;                   // [kXStages]
  uint64_t* dy_full = bars + 2 * kXStages;               // [kDyStages]
  uint64_t* dy_empty = dy_full + kMaxDyStages;           // [kDyStages]
  uint64_t* w_full = dy_empty + kMaxDyStages;
  uint64_t* dx_full = w_full + 1;
  uint64_t* dx_empty = w_full + 2;                       // 4 arrivals (one per epilogue warp)
  uint64_t* dw_full = w_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // tiles of this CTA: blockIdx.x, + gridDim.x, ...
  const int my_tiles = P.n_tiles > (int)blockIdx.x ? (P.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // ---- one-time setup: ones panel of every X slot (channel 0 of the panel = 1.0, the rest 0), barriers, TMEM
  for (int i = tid; i < kXStages * kPanel / 16; i += kLThreads) {
    const int s = i / (kPanel / 16), o = i % (kPanel / 16);
    const int row = o >> 2, chunk = o & 3;               // 4 x 16-byte chunks per 64-byte row
    uint4 v = make_uint4(0, 0, 0, 0);
    if (chunk == ((row >> 1) & 3)) v.x = 0x3F80u;        // logical chunk 0 sits at physical chunk 0 ^ swizzle phase
    reinterpret_cast<uint4*>(sX + s * kXSlot + 3 * kPanel)[o] = v;
  }
  if (tid == 0) {
    for (int s = 0; s < kXStages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    for (int s = 0; s < kDyStages; ++s) { mbar_init(&dy_full[s], 1); mbar_init(&dy_empty[s], 1); }
    mbar_init(w_full, 1); mbar_init(dx_full, 1); mbar_init(dx_empty, 4); mbar_init(dw_full, 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&P.dy); tma_prefetch_desc(&P.x); tma_prefetch_desc(&P.w); }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&P.dx);
  if (warp == 5) tmem_alloc<kLTmemCols>(tmem_slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, G * kWChunk);
      for (int c = 0; c < G; ++c) tma_load_3d(&P.w, w_full, sW + c * kWChunk, 0, c * 96, 0);
      int dn = 0;
      for (int n = 0; n < my_tiles; ++n) {
        const int row0 = ((int)blockIdx.x + n * (int)gridDim.x) * kLTile;
        const int xs = n % kXStages;
        mbar_wait(&x_empty[xs], ((n / kXStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&x_full[xs], kLTileBytes);
        tma_load_3d(&P.x, &x_full[xs], sX + xs * kXSlot, 0, row0, 0);
        for (int c = 0; c < G; ++c, ++dn) {
          const int ds = dn % kDyStages;
          mbar_wait(&dy_empty[ds], ((dn / kDyStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&dy_full[ds], kLTileBytes);
          tma_load_3d(&P.dy, &dy_full[ds], sDY + ds * kLTileBytes, 0, row0, 3 * c);
        }
      }
    }
  } else if (warp == 5) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idescDX = umma_idesc_bf16(128, 96, 0, 1);     // A K-major (dY), B MN-major (W)
    constexpr uint32_t idescDW = umma_idesc_bf16(128, 96, 1, 1);     // A MN-major ([X | 1]^T), B MN-major (dY)
    const uint64_t dAk = umma_smem_desc(0, 0, 512, kSwz64);          // dY tile, K-major
    const uint64_t dMn = umma_smem_desc(0, kPanel, 512, kSwz64);     // dY / X tile, MN-major: channel panels 8 KB apart
    const uint64_t dBw = umma_smem_desc(0, 96 * 64, 512, kSwz64);    // W chunk, MN-major: in-channel panels 6 KB apart
    const uint32_t dy0 = smem_u32(sDY) >> 4, x0 = smem_u32(sX) >> 4, w0 = smem_u32(sW) >> 4;
    mbar_wait(w_full, 0);
    int dn = 0;
    for (int n = 0; n < my_tiles; ++n) {
      const int xs = n % kXStages;
      mbar_wait(&x_full[xs], (n / kXStages) & 1);
      mbar_wait(dx_empty, (n & 1) ^ 1);                   // the epilogue has read the previous tile's dX accumulators
      for (int c = 0; c < G; ++c, ++dn) {
        const int ds = dn % kDyStages;
        mbar_wait(&dy_full[ds], (dn / kDyStages) & 1);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a0 = dy0 + ds * (kLTileBytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 6; ++ks)                 // 16 out-channels per step
            umma_bf16_ss(tmem + 384, dAk + (a0 + (ks >> 1) * (kPanel >> 4) + (ks & 1) * 2),
                         dBw + (w0 + c * (kWChunk >> 4) + ks * 64), idescDX, (c | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)                 // 16 tokens per step
            umma_bf16_ss(tmem + c * 96, dMn + (x0 + xs * (kXSlot >> 4) + ks * 64), dMn + (a0 + ks * 64), idescDW, (n | ks) != 0);
          umma_commit(&dy_empty[ds]);
          if (c == G - 1) { umma_commit(dx_full); umma_commit(&x_empty[xs]); }
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(dw_full);
    __syncwarp();
  } else {
    // ============================== epilogue (warps 0-3): one thread per token row / per out-channel row ==============================
    const int r = tid;                                    // 0..127 = TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int rsw = (r >> 1) & 3;
    for (int n = 0; n < my_tiles; ++n) {
      const int row0 = ((int)blockIdx.x + n * (int)gridDim.x) * kLTile;
      mbar_wait(dx_full, n & 1);
      tcgen05_fence_after();
      named_bar_sync(1, 128);                             // thread 0 has seen the previous store read the staging tile out
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + lane_base + 384 + p * 32, v);
        tmem_ld_wait();
        uint8_t* orow = sOut + p * kPanel + r * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(orow + ((c ^ rsw) << 4)) =
              make_uint4(pack_bf16x2(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1])),
                         pack_bf16x2(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3])),
                         pack_bf16x2(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5])),
                         pack_bf16x2(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7])));
      }
      tcgen05_fence_before();
      mbar_arrive_warp(dx_empty);
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (tid == 0) {
        tma_store_3d(&P.dx, sOut, 0, row0, 0);
        tma_store_commit();
        tma_store_wait_read<0>();
      }
    }
    if (tid == 0) tma_store_wait_all<0>();
    // ---- dW^T | db partials of this CTA: thread r = in-channel r (r < 96) or the db row (r == 96)
    mbar_wait(dw_full, 0);
    tcgen05_fence_after();
    const int ncol = G * 96;
    float* wsb = P.ws + (size_t)blockIdx.x * 97 * ncol;
    for (int p = 0; p < 3 * G; ++p) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + lane_base + p * 32, v);
      tmem_ld_wait();
      if (r < 97) {
        float4* dst = reinterpret_cast<float4*>(wsb + (size_t)r * ncol + p * 32);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          dst[e] = my_tiles > 0 ? make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                              __uint_as_float(v[4 * e + 3]))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<kLTmemCols>(tmem);
}

// out[e] = sum over CTAs of ws[cta][e]; element e = (row i in 0..96, out-channel o): rows 0..95 -> dW[o][i], row 96 -> db[o].
// Block = 128 elements (one float4 of four out-channels per lane) x 32 slices of the CTA range: with 148 partials a thread
// issues at most five independent loads (the kernel is pure latency: 8 slices = 19 loads per thread took 6.2 us per launch,
// 192 launches in a cfg3 step).  Summation order is fixed: deterministic.
constexpr int kRedSlices = 32;
__global__ void __launch_bounds__(kRedSlices * 32)
linbwd_reduce_kernel(const float* __restrict__ ws, int n_cta, int ncol, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float4 part[kRedSlices][32];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int e = (blockIdx.x * 32 + lane) * 4;
  const int n = 97 * ncol;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e < n) {
#pragma unroll 5
    for (int c = sl; c < n_cta; c += kRedSlices) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ws + (size_t)c * n + e));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  part[sl][lane] = s;
  __syncthreads();
  if (sl < 4) {                                                // slices 8 sl .. 8 sl + 7
    s = part[8 * sl][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 v = part[8 * sl + k][lane]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    part[8 * sl][lane] = s;
  }
  __syncthreads();
  if (sl == 0 && e < n) {
#pragma unroll
    for (int k = 1; k < 4; ++k) { const float4 v = part[8 * k][lane]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    const int i = e / ncol, o = e - i * ncol;                // ncol % 4 == 0: the four elements share the row
    if (i < 96) { dw[o * 96 + i] = s.x; dw[(o + 1) * 96 + i] = s.y; dw[(o + 2) * 96 + i] = s.z; dw[(o + 3) * 96 + i] = s.w; }
    else if (db) *reinterpret_cast<float4*>(db + o) = s;
  }
}

static size_t linbwd_smem_bytes(int G) { return 1024 + (size_t)(G == 4 ? 2 : 3) * kLTileBytes + kXStages * kXSlot + (size_t)G * kWChunk + kLTileBytes + 32 * 8; }

static bool make_panel_map(CUtensorMap* out, const void* ptr, long long rows, int cols, long long ld, int box_rows, int box_panels) {
  EncodeTiledFn enc = encode_fn();
  if (!enc || reinterpret_cast<uintptr_t>(ptr) % 16 || ld % 8 || cols % 32) return false;
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 64};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_panels};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

const char* linbwd_why_not(int io_dtype, long long rows, int in_features, int out_features, long long ld_dy, long long ld_x, long long ld_dx) {
  if (io_dtype != MMN_DT_BF16) return "io dtype is not bf16";
  if (in_features != kLC) return "in_features != 96";
  if (out_features != 96 && out_features != 192 && out_features != 288 && out_features != 384) return "out_features is not 96, 192, 288 or 384";
  if (rows < 1) return "no rows";
  if (ld_dy % 8 || ld_x % 8 || ld_dx % 8) return "leading dimension not 16-byte aligned";
  if (!encode_fn()) return "cuTensorMapEncodeTiled unavailable";
  return nullptr;
}

size_t linbwd_workspace_bytes(int out_features) { return (size_t)num_sms_cached() * out_features * 97 * sizeof(float); }

int linbwd(const void* dy, const void* x, const void* w, void* dx, float* dw, float* db, float* workspace, long long rows,
           int in_features, int out_features, long long ld_dy, long long ld_x, long long ld_dx, cudaStream_t st, char* err,
           size_t errlen, int* launches) {
  LinBwdParams P;
  P.M = (int)rows;
  P.G = out_features / 96;
  P.n_tiles = (int)((rows + kLTile - 1) / kLTile);
  P.ws = workspace;
  if (!make_panel_map(&P.dy, dy, rows, out_features, ld_dy, kLTile, 3) || !make_panel_map(&P.x, x, rows, in_features, ld_x, kLTile, 3) ||
      !make_panel_map(&P.w, w, out_features, in_features, in_features, 96, 3) || !make_panel_map(&P.dx, dx, rows, in_features, ld_dx, kLTile, 3)) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (pointer alignment or strides)");
    return MMN_ERR_CUDA;
  }
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(linbwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(linbwd_smem_bytes(3) > linbwd_smem_bytes(4) ? linbwd_smem_bytes(3) : linbwd_smem_bytes(4))); });
  int grid = num_sms_cached();
  if (grid > P.n_tiles) grid = P.n_tiles;
  linbwd_tc_kernel<<<grid, kLThreads, linbwd_smem_bytes(P.G), st>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(err, errlen, "linbwd_tc_kernel: %s", cudaGetErrorString(e)); return MMN_ERR_CUDA; }
  ++*launches;
  const int n = out_features * 97;
  linbwd_reduce_kernel<<<(n / 4 + 31) / 32, kRedSlices * 32, 0, st>>>(workspace, grid, out_features, dw, db);   // ncol = out_features
  e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(err, errlen, "linbwd_reduce_kernel: %s", cudaGetErrorString(e)); return MMN_ERR_CUDA; }
  ++*launches;
  return MMN_OK;
}

}}  // namespace mmn::tc
