"""Debug helper: per-phase clock trace of CTA 0 of the tcgen05 backward kernel (MMN_TC_TRACE_BWD)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import _lib, ops  # noqa: E402,F401

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_bwd.txt"
B, grid, nH, d = int(os.environ.get("MMN_TRACE_B", "32")), (32, 32, 32), 3, 32
C = nH * d
qkv = torch.randn(B, *grid, 3 * C, device="cuda", dtype=torch.bfloat16)
dout = torch.randn(B, *grid, C, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(nH, 64, 64, device="cuda")
hs = torch.rand(nH, device="cuda") * 10 + 1
args = (list(grid), [4, 4, 4], [2, 2, 2], nH, 1, 1, 1.0, 0.0, 0, 0, 0)
out, lse = torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, None, *args)
for _ in range(3):
    torch.ops.mmn_b200.winattn_bwd(dout, qkv, None, bias, hs, None, out, lse, *args, True)
torch.cuda.synchronize()
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
os.environ["MMN_TC_TRACE_BWD"] = out_path
torch.ops.mmn_b200.winattn_bwd(dout, qkv, None, bias, hs, None, out, lse, *args, True)
torch.cuda.synchronize()
print("traced ->", out_path)
