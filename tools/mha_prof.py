"""Run the tensor-core MHA kernels a few times at the scaled shape (E=768, 12 heads x 64, T=S=2048) -- the process that
ncu profiles (tools: ncu --set full -k regex:mha_ ... python tools/mha_prof.py [batch] [mask_kind] [dropout_p])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mask_kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
drop = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0          # attention dropout probability
T = S = 2048
nH, d = 12, 64
E = nH * d
torch.manual_seed(0)
q, k, v, do = (torch.randn(T, B, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
for _ in range(3):
    out, lse = ops.mha_fwd(q, k, v, None, nH, mask_kind, 1, d ** -0.5, drop, 11, 5)
    dq, dk, dv = ops.mha_bwd(do, q, k, v, None, out, lse, nH, mask_kind, 1, d ** -0.5, drop, 11, 5)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(5):
    out, lse = ops.mha_fwd(q, k, v, None, nH, mask_kind, 1, d ** -0.5, drop, 11, 5)
e1.record()
for _ in range(5):
    dq, dk, dv = ops.mha_bwd(do, q, k, v, None, out, lse, nH, mask_kind, 1, d ** -0.5, drop, 11, 5)
e2.record()
torch.cuda.synchronize()
fl = 4 * T * S * E * B
print(f"B={B} mask={mask_kind} dropout={drop}: fwd {e0.elapsed_time(e1) / 5:.3f} ms ({fl / e0.elapsed_time(e1) * 5 / 1e9:.0f} TFLOP/s)  "
      f"bwd {e1.elapsed_time(e2) / 5:.3f} ms ({2.5 * fl / e1.elapsed_time(e2) * 5 / 1e9:.0f} TFLOP/s)")
