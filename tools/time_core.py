"""Time the tcgen05 window-attention core kernels alone at the cfg2 shape (CUDA events, median of N launches,
tensors far larger than L2).  MMN_LIB=<path> times another build of the library (A/B on the same box).
Usage: python tools/time_core.py [B] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import _lib, ops  # noqa: E402,F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
grid, nH, d = (32, 32, 32), 3, 32
C = nH * d
torch.manual_seed(0)
qkv = torch.randn(B, *grid, 3 * C, device="cuda", dtype=torch.bfloat16)
dout = torch.randn(B, *grid, C, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(nH, 64, 64, device="cuda")
hs = torch.rand(nH, device="cuda") * 10 + 1
args = (list(grid), [4, 4, 4], [2, 2, 2], nH, 1, 1, 1.0, 0.0, 0, 0, 0)
out, lse = torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, None, *args)


def timeit(fn, burst=10):
    for _ in range(30):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(burst):
            fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / burst)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


f_med, f_min = timeit(lambda: torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, None, *args))
b_med, b_min = timeit(lambda: torch.ops.mmn_b200.winattn_bwd(dout, qkv, None, bias, hs, None, out, lse, *args, True))
nwin = B * 512
print(f"{os.environ.get('MMN_LIB', 'default')}: fwd {f_med:.4f} ms (min {f_min:.4f}, {nwin * 49920 / f_med / 1e6:.0f} GB/s)  "
      f"bwd {b_med:.4f} ms (min {b_min:.4f}, {nwin * 99840 / b_med / 1e6:.0f} GB/s)")
