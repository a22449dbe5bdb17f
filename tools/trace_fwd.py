"""Debug helper: dump the per-phase clock trace of CTA 0 of the tcgen05 forward kernel
(MMN_TC_TRACE) for the cfg2 shape.  Usage on the GPU box: python tools/trace_fwd.py out.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import _lib, ops  # noqa: E402,F401

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_fwd.txt"
B, grid, nH, d = int(os.environ.get("MMN_TRACE_B", "32")), (32, 32, 32), 3, 32
C = nH * d
qkv = torch.randn(B, *grid, 3 * C, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(nH, 64, 64, device="cuda")
hs = torch.rand(nH, device="cuda") * 10 + 1
args = (list(grid), [4, 4, 4], [2, 2, 2], nH, 1, 1, 1.0, 0.0, 0, 0, 0)
for _ in range(3):
    torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, None, *args)
torch.cuda.synchronize()
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
os.environ["MMN_TC_TRACE"] = out
torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, None, *args)
torch.cuda.synchronize()
print("traced ->", out)
