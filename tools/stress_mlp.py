"""Run the tensor-core projection kernels (csrc/gemm_tc.cu: forward with activation epilogue, dgrad with act' epilogue,
wgrad, column sums) repeatedly on the same inputs: every output must be bit-identical from run to run except the
atomically accumulated bias gradients.  A race in the kernels' pipelines shows up as a sporadic mismatch; the report
says which tensor, how many elements and which rows / columns.    tools/stress_mlp.py [C] [rows] [iters] [act]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import _lib, ops  # noqa: E402,F401

C = int(sys.argv[1]) if len(sys.argv) > 1 else 768
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2100
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
act = {"relu": _lib.ACT_RELU, "gelu": _lib.ACT_GELU}[sys.argv[4] if len(sys.argv) > 4 else "relu"]
torch.manual_seed(0)
dev = "cuda"
x = torch.randn(rows, C, device=dev).bfloat16()
w1 = (torch.randn(4 * C, C, device=dev) * C ** -0.5).bfloat16()
w2 = (torch.randn(C, 4 * C, device=dev) * (4 * C) ** -0.5).bfloat16()
b1 = torch.randn(4 * C, device=dev) * 0.1
b2 = torch.randn(C, device=dev) * 0.1
dy = torch.randn(rows, C, device=dev).bfloat16()
names = ["h", "dact", "y", "dpre", "dw2", "db2", "dx", "dw1", "db1"]
exact = {"h", "dact", "y", "dpre", "dw2", "dx", "dw1"}
ref, bad = None, 0
poison = os.environ.get("POISON", "0") == "1"       # every buffer the ops allocate starts out as NaN bytes: a read of memory
for it in range(iters):                             # the kernels should have written (or never read) turns into a NaN
    if poison:
        del_me = torch.full((1 << 30,), 0xFF, dtype=torch.uint8, device=dev)
        del del_me
    h, dact = torch.ops.mmn_b200.linear_fwd(x, w1, b1, act, True)
    y, _ = torch.ops.mmn_b200.linear_fwd(h, w2, b2, _lib.ACT_NONE, False)
    dpre, dw2, db2 = torch.ops.mmn_b200.linear_bwd(dy, h, w2, dact, act, True, True)
    dx, dw1, db1 = torch.ops.mmn_b200.linear_bwd(dpre, x, w1, None, 0, True, True)
    cur = (h, dact, y, dpre, dw2, db2, dx, dw1, db1)
    if it % 7 == 3:                                   # perturb the timing: a burst of unrelated work on the same stream
        torch.randn(1 << 22, device=dev).sum()
    if ref is None:
        ref = [t.clone() for t in cur]
        for n, a in zip(names, cur):
            if not torch.isfinite(a.float()).all():
                print(f"iter 0: {n} has non-finite elements", flush=True)
        del cur, h, dact, y, dpre, dw2, db2, dx, dw1, db1
        continue
    for n, a, b in zip(names, cur, ref):
        if not torch.isfinite(a.float()).all():
            nz = torch.nonzero(~torch.isfinite(a.float()).reshape(a.shape[0], -1))
            print(f"iter {it}: {n} {tuple(a.shape)} has {nz.shape[0]} non-finite elements, rows {nz[:, 0].min().item()}..{nz[:, 0].max().item()}, "
                  f"cols {nz[:, 1].min().item()}..{nz[:, 1].max().item()}", flush=True)
            bad += 1
            continue
        diff = (a.float() - b.float()).abs()
        if n in exact:
            if not torch.equal(a, b):
                nz = torch.nonzero(diff > 0)
                r0, r1, c0, c1 = nz[:, 0].min().item(), nz[:, 0].max().item(), nz[:, 1].min().item(), nz[:, 1].max().item()
                print(f"iter {it}: {n} {tuple(a.shape)} differs: {nz.shape[0]} elements, rows {r0}..{r1}, cols {c0}..{c1}, "
                      f"max abs {diff.max().item():.3e} (ref max {b.float().abs().max().item():.3e})", flush=True)
                bad += 1
        elif (diff.max() / b.abs().max()).item() > 1e-4:
            print(f"iter {it}: {n} differs beyond round-off: {(diff.max() / b.abs().max()).item():.3e}", flush=True)
            bad += 1
    del cur, h, dact, y, dpre, dw2, db2, dx, dw1, db1
torch.cuda.synchronize()
print(f"C={C} rows={rows} act={act}: {iters} iterations, {bad} mismatches")
