"""Run the tensor-core forward/backward repeatedly on the same inputs: outputs and input gradients must be
bit-identical from run to run (no atomics on their path); dbias / dhead_scale / dcolsum (atomic flushes) must agree
to fp32 round-off.  A race in the kernels' pipelines shows up here as a sporadic mismatch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import _lib, ops  # noqa: E402,F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
grid, nH, d = (32, 32, 32), 3, 32
C = nH * d
torch.manual_seed(0)
qkv = torch.randn(B, *grid, 3 * C, device="cuda", dtype=torch.bfloat16)
dout = torch.randn(B, *grid, C, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(nH, 64, 64, device="cuda")
hs = torch.rand(nH, device="cuda") * 10 + 1
args = (list(grid), [4, 4, 4], [2, 2, 2], nH, 1, 1, 1.0, 0.0, 0, 0, 0)
ref = None
bad = 0
for it in range(iters):
    out, lse = torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, None, *args)
    da, _, dbias, dhs, dcs = torch.ops.mmn_b200.winattn_bwd(dout, qkv, None, bias, hs, None, out, lse, *args, True)
    cur = (out, lse, da, dbias, dhs, dcs)
    if ref is None:
        ref = [t.clone() for t in cur]
        continue
    names = ["out", "lse", "dqkv", "dbias", "dhead_scale", "dcolsum"]
    for n, a, b in zip(names, cur, ref):
        if n in ("out", "lse", "dqkv"):
            if not torch.equal(a, b):
                diff = (a.float() - b.float()).abs()
                print(f"iter {it}: {n} differs: max abs {diff.max().item():.3e}, {int((diff > 0).sum())} elements, first at {torch.nonzero(diff.reshape(-1) > 0)[0].item()}")
                bad += 1
        else:
            rel = ((a - b).abs().max() / b.abs().max()).item()
            if rel > 1e-4:
                print(f"iter {it}: {n} rel diff {rel:.3e}")
                bad += 1
print("mismatches:", bad)
