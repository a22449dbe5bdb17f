// tc_window.cuh -- window geometry shared by the tensor-core forward and backward kernels:
// the decomposition of a (possibly wrapped) shifted window into contiguous TMA boxes and the
// piece-major token permutation that goes with it.  All closed forms restate SURVEY.md 8a (a1-a4).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/mmn_b200.h"
#include "tc_common.cuh"

namespace mmn { namespace tc {

constexpr int kN = 64;            // tokens per window (tuned path)
constexpr int kD = 32;            // head_dim (tuned path)

struct WinShape {                 // right-aligned geometry: unused leading axes have extent 1
  int grid[3], win[3], shift[3], nwin[3];
  int nW;                         // windows per sample
  int n_windows;                  // batch * nW
};

struct WinGeom {
  int b, start[3], idx[3];
  int cls;       // wrap class: bit x set if the window wraps around the volume edge along axis x
};

// A window that wraps along k axes is 2^k contiguous pieces of the source volume.  Piece q is
// one TMA box and lands at rows [q*psize, (q+1)*psize) of the window's 64-row tile, so the
// tile holds the window's tokens in PIECE-MAJOR order.  Attention is equivariant to that
// permutation as long as bias / mask / lse are addressed through it (piece_position) and the
// outputs are stored through the same boxes.  Axis 2 is the least significant wrapped axis.
// The tensor maps come one per wrap class (box = half extent along every wrapped axis); the boxes of an item are issued by
// tc_sched.cuh (BoxPlan / issue_item_boxes).

// Window position (row-major over the window) of tile row `row` for wrap class `cls`.
__device__ __forceinline__ int piece_position(const WinShape& S, int cls, int row) {
  int seg[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) seg[a] = (cls >> a) & 1 ? S.win[a] >> 1 : S.win[a];
  const int psize = seg[0] * seg[1] * seg[2];
  int q = row / psize, rem = row - q * psize;
  int b2 = rem % seg[2]; rem /= seg[2];
  int b1 = rem % seg[1]; int b0 = rem / seg[1];
  int a2 = b2, a1 = b1, a0 = b0;
  if (cls & 4) { a2 += (q & 1) * seg[2]; q >>= 1; }
  if (cls & 2) { a1 += (q & 1) * seg[1]; q >>= 1; }
  if (cls & 1) { a0 += (q & 1) * seg[0]; }
  return (a0 * S.win[1] + a1) * S.win[2] + a2;
}

// ------------------------------------------------------------------------------------------
// Host helpers
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();

inline WinShape shape_from(const mmn_winattn_desc* d) {
  WinShape g;
  for (int a = 0; a < 3; ++a) { g.grid[a] = 1; g.win[a] = 1; g.shift[a] = 0; g.nwin[a] = 1; }
  for (int a = 0; a < d->ndim; ++a) {
    int t = 3 - d->ndim + a;
    g.grid[t] = d->grid[a]; g.win[t] = d->window[a]; g.shift[t] = d->shift[a]; g.nwin[t] = d->grid[a] / d->window[a];
  }
  g.nW = g.nwin[0] * g.nwin[1] * g.nwin[2];
  g.n_windows = d->batch * g.nW;
  return g;
}

inline int num_sms_cached() {
  static std::once_flag once;
  static int num_sms = 148;
  std::call_once(once, [] {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
  return num_sms;
}

// Eight tensor maps (one box shape per wrap class) over a (B, g0, g1, g2, channels) bf16 tensor
// whose tokens are `row_stride` elements apart; boxes are 32 channels (one head) wide, 64B-swizzled.
bool make_window_maps(CUtensorMap* out, const void* ptr, long long row_stride, int batch, int channels, const WinShape& g);

}}  // namespace mmn::tc
