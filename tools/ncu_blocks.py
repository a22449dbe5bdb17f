#!/usr/bin/env python
"""One launch of every building-block kernel at a workload's shapes between cudaProfilerStart/Stop, for
  ncu --set full --profile-from-start off --import-source on -o prof python tools/ncu_blocks.py [rows] [C]
(tools/bench_blocks.py times the same calls)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_neuroimage_b200 import _lib, ops  # noqa: E402,F401  (ops registers torch.ops.mmn_b200.*)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 110592
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    dev = torch.device("cuda", 0)
    _lib.load()
    bf, f32 = torch.bfloat16, torch.float32
    r = lambda *s, dt=bf: torch.randn(*s, device=dev, dtype=dt)
    O = torch.ops.mmn_b200
    resid, delta, gamma, beta = r(rows, C, dt=f32), r(rows, C), r(C, dt=f32), r(C, dt=f32)
    gs, gn = r(rows, C, dt=f32), r(rows, C)
    x, h, dy = r(rows, C), r(rows, 4 * C), r(rows, C)
    dact = r(rows, 4 * C)
    w_qkv, w_p, w1, w2 = r(3 * C, C), r(C, C), r(4 * C, C), r(C, 4 * C)
    b_qkv, b1, b2 = r(3 * C, dt=f32), r(4 * C, dt=f32), r(C, dt=f32)
    dqkv, dh = r(rows, 3 * C), r(rows, 4 * C)
    calls = []
    s, n, mean, rstd = O.layernorm_fwd(resid, delta, gamma, beta, 1e-5, 0, True, _lib.DT_BF16)
    calls.append(lambda: O.layernorm_fwd(resid, delta, gamma, beta, 1e-5, 0, True, _lib.DT_BF16))
    calls.append(lambda: O.layernorm_bwd(gs, gn, s, gamma, mean, rstd, 0, _lib.DT_F32, _lib.DT_BF16, True))
    calls.append(lambda: O.linear_fwd(x, w_qkv, b_qkv, 0, False))
    calls.append(lambda: O.linear_fwd(x, w_p, b2, 0, False))
    calls.append(lambda: O.linear_fwd(x, w1, b1, _lib.ACT_GELU, True))
    calls.append(lambda: O.linear_fwd(h, w2, b2, 0, False))
    calls.append(lambda: O.linear_bwd(dy, h, w2, dact, _lib.ACT_GELU, True, True))
    calls.append(lambda: O.linear_bwd(dh, x, w1, None, 0, True, True))
    calls.append(lambda: O.linear_bwd(dqkv, x, w_qkv, None, 0, True, True))
    calls.append(lambda: O.linear_bwd(dy, x, w_p, None, 0, True, True))
    if C == 96 and rows % 13824 == 0:
        B = rows // 13824
        grid, nH = (24, 24, 24), 3
        qkv = r(B, *grid, 3 * C)
        bias = r(nH, 64, 64, dt=f32)
        hs = torch.full((nH,), 10.0, device=dev)
        args = (list(grid), [4, 4, 4], [2, 2, 2], nH, _lib.SCORE_COSINE, _lib.MASK_SHIFT, 1.0, 0.0, 0, 0, _lib.PATH_AUTO)
        o, lse = O.winattn_fwd(qkv, None, bias, hs, None, *args)
        do = r(B, *grid, C)
        calls.append(lambda: O.winattn_fwd(qkv, None, bias, hs, None, *args))
        calls.append(lambda: O.winattn_bwd(do, qkv, None, bias, hs, None, o, lse, *args, True))
    for _ in range(2):
        for c in calls:
            c()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for c in calls:
        c()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
