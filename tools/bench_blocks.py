#!/usr/bin/env python
"""Per-op times of the building blocks of a Swin block at one workload's shapes (default cfg3: 8 x 24^3 = 110 592 tokens of
96 channels), against each op's HBM floor (algorithmic bytes / measured peak).  CUDA events around bursts of launches.

  python tools/bench_blocks.py [rows] [C]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_neuroimage_b200 import _lib, ops  # noqa: E402

PEAK = 6548.2e9


def timeit(fn, n=10, reps=5):
    """GPU time per call: n calls captured in one CUDA graph (no host launch overhead between them), best of reps replays."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    del g
    return best * 1e3                                              # us


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 110592
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    dev = torch.device("cuda", 0)
    _lib.load()
    bf, f32 = torch.bfloat16, torch.float32
    r = lambda *s, dt=bf: torch.randn(*s, device=dev, dtype=dt)
    out = []

    def row(name, us, byts):
        floor = byts / PEAK * 1e6
        out.append((name, us, byts / 1e6, floor, floor / us))

    resid, delta = r(rows, C, dt=f32), r(rows, C)
    gamma, beta = r(C, dt=f32), r(C, dt=f32)
    for mode, nm in ((_lib.LN_PRE, "pre"), (_lib.LN_POST, "post")):
        us = timeit(lambda: torch.ops.mmn_b200.layernorm_fwd(resid, delta, gamma, beta, 1e-5, mode, True, _lib.DT_BF16))
        row(f"layernorm_fwd {nm} (resid f32 + delta bf16 -> sum f32, norm bf16)", us, rows * C * (4 + 2 + 4 + 2))
        s, n, mean, rstd = torch.ops.mmn_b200.layernorm_fwd(resid, delta, gamma, beta, 1e-5, mode, True, _lib.DT_BF16)
        gs, gn = r(rows, C, dt=f32), r(rows, C)
        x = s if mode == _lib.LN_PRE else delta
        us = timeit(lambda: torch.ops.mmn_b200.layernorm_bwd(gs, gn, x, gamma, mean, rstd, mode, _lib.DT_F32, _lib.DT_BF16, True))
        row(f"layernorm_bwd {nm} (gs f32, gn bf16, x -> d_resid f32, d_delta bf16)", us, rows * C * (4 + 2 + (4 if mode == 0 else 2) + 4 + 2))
    us = timeit(lambda: torch.ops.mmn_b200.layernorm_fwd(resid, None, gamma, beta, 1e-5, 0, False, _lib.DT_BF16))
    row("layernorm_fwd plain (f32 -> bf16)", us, rows * C * (4 + 2))

    x = r(rows, C)
    for n_out, act, pre, nm in ((3 * C, 0, False, "qkv"), (C, 0, False, "proj"), (4 * C, _lib.ACT_GELU, True, "fc1+gelu+pre"),
                                (4 * C, _lib.ACT_GELU, False, "fc1+gelu (no pre)")):
        w, b = r(n_out, C), r(n_out, dt=f32)
        us = timeit(lambda: torch.ops.mmn_b200.linear_fwd(x, w, b, act, pre))
        row(f"linear_fwd {C}->{n_out} {nm}", us, rows * 2 * (C + n_out * (2 if pre else 1)))
    h = r(rows, 4 * C)
    w2, b2 = r(C, 4 * C), r(C, dt=f32)
    us = timeit(lambda: torch.ops.mmn_b200.linear_fwd(h, w2, b2, 0, False))
    row(f"linear_fwd {4 * C}->{C} fc2", us, rows * 2 * (5 * C))

    dy = r(rows, C)
    pre = r(rows, 4 * C)
    us = timeit(lambda: torch.ops.mmn_b200.linear_bwd(dy, h, w2, pre, _lib.ACT_GELU, True, True))
    row("linear_bwd fc2 (dgrad+gelu' , wgrad)", us, rows * 2 * (C + 4 * C + 4 * C + 4 * C))
    us = timeit(lambda: torch.ops.mmn_b200.linear_bwd(dy, h, w2, pre, _lib.ACT_GELU, True, False))
    row("  of which dgrad+gelu'", us, rows * 2 * (C + 4 * C + 4 * C))
    us = timeit(lambda: torch.ops.mmn_b200.linear_bwd(dy, h, w2, None, 0, False, True))
    row("  of which wgrad", us, rows * 2 * (C + 4 * C))
    for n_out, nm in ((4 * C, "fc1"), (3 * C, "qkv"), (C, "proj")):
        w, dyo = r(n_out, C), r(rows, n_out)
        us = timeit(lambda: torch.ops.mmn_b200.linear_bwd(dyo, x, w, None, 0, True, True))
        row(f"linear_bwd {nm} (dy {n_out}, x {C} -> dx, dw, db)", us, rows * 2 * (n_out + 2 * C))
    us = timeit(lambda: torch.ops.mmn_b200.colsum(dy))
    row("colsum", us, rows * 2 * C)
    a = r(rows, C, dt=f32)
    b = r(rows, C, dt=f32)
    us = timeit(lambda: torch.add(a, b))
    row("torch.add f32", us, rows * C * 12)
    us = timeit(lambda: a.to(bf))
    row("torch f32->bf16 cast", us, rows * C * 6)

    if C == 96 and rows % 13824 == 0:
        B = rows // 13824
        grid, nH = (24, 24, 24), 3
        qkv = r(B, *grid, 3 * C)
        bias = r(nH, 64, 64, dt=f32)
        hs = torch.full((nH,), 10.0, device=dev)
        args = (list(grid), [4, 4, 4], [2, 2, 2], nH, _lib.SCORE_COSINE, _lib.MASK_SHIFT, 1.0, 0.0, 0, 0, _lib.PATH_AUTO)
        us = timeit(lambda: torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, None, *args))
        row("winattn_fwd (cosine, shift)", us, rows * C * 2 * 4)
        o, lse = torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, None, *args)
        do = r(B, *grid, C)
        us = timeit(lambda: torch.ops.mmn_b200.winattn_bwd(do, qkv, None, bias, hs, None, o, lse, *args, True))
        row("winattn_bwd", us, rows * C * 2 * 7)

    print(f"# building blocks at rows = {rows}, C = {C} (us per launch, 10 calls per CUDA-graph replay; floor = bytes / 6548 GB/s)\n")
    print("| op | us | MB | floor us | frac of HBM peak |\n|---|---|---|---|---|")
    for name, us, mb, fl, fr in out:
        print(f"| {name} | {us:.1f} | {mb:.0f} | {fl:.1f} | {fr:.2f} |")


if __name__ == "__main__":
    main()
