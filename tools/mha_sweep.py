"""Tensor-core MHA kernels: forward / backward time over sequence length and mask kind (scaled shape: 12 heads x 64).
  python tools/mha_sweep.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nH, d = 12, 64
E = nH * d
torch.manual_seed(0)


def timeit(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for T in (512, 1024, 2048, 4096):
    b = max(1, B * 2048 // T)
    q, k, v, do = (torch.randn(T, b, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    for mask_kind in (0, 3):
        out, lse = ops.mha_fwd(q, k, v, None, nH, mask_kind, 1, d ** -0.5, 0.0, 0, 0)
        f = timeit(lambda: ops.mha_fwd(q, k, v, None, nH, mask_kind, 1, d ** -0.5, 0.0, 0, 0))
        g = timeit(lambda: ops.mha_bwd(do, q, k, v, None, out, lse, nH, mask_kind, 1, d ** -0.5, 0.0, 0, 0))
        work = 0.5 + 64.0 / T if mask_kind == 3 else 1.0          # fraction of 128 x 128 blocks a causal pass touches
        fl = 4 * T * T * E * b * work
        print(f"T={T} batch={b} mask={mask_kind}: fwd {f:.3f} ms ({fl / f / 1e9:.0f} TFLOP/s of the visible blocks)  "
              f"bwd {g:.3f} ms ({2.5 * fl / g / 1e9:.0f})")
