#!/usr/bin/env python
"""Kernel-time breakdown of one training step of a workload (cfg3 / cfg4 / cfg5) with torch.profiler (CUPTI): which kernels the
step spends its GPU time in, own kernels vs PyTorch's.  Eager launches (a CUDA graph hides the kernels from the profiler's
per-op view); times are GPU durations, so host launch gaps do not count.

  python tools/profile_step.py cfg3 [batch] > profiles/<round>_cfg3_kernels.md
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multimodal_neuroimage_b200 import train_step as TS  # noqa: E402
from multimodal_neuroimage_b200 import workloads as W  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else (1 if name == "cfg5" else 8)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = (W.SwinFusion3D() if name == "cfg3" else W.FuncStructCross3D() if name == "cfg4" else W.SwinV2CrossModal3D())
    W.randomise_norms(model)
    model = model.to(dev)
    img = 128 if name == "cfg5" else 96
    if name == "cfg4":
        *ins, y = W.synthetic_batch_cfg4(batch, img, dev, pinned=True)
        ins = tuple(ins)
    else:
        A, B, y = W.synthetic_batch(batch, img, dev, pinned=True)
        ins = (A, B)
    ts = TS.TrainStep(model, torch.nn.functional.binary_cross_entropy_with_logits, ins, y, use_graph=False)
    for _ in range(2):
        ts()
    torch.cuda.synchronize()
    steps = 2
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            ts()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t > 0:
            rows.append((t / steps, e.count // steps, e.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    own = sum(r[0] for r in rows if "mmn" in r[2])
    print(f"# {name} training step, batch {batch}: GPU time by kernel (eager, {steps} steps averaged)\n")
    print(f"total GPU time per step {total / 1e3:.2f} ms in {sum(r[1] for r in rows)} launches; own kernels (namespace mmn) {own / 1e3:.2f} ms = {100 * own / total:.1f} %\n")
    print("| ms / step | % | launches / step | kernel |\n|---|---|---|---|")
    for t, n, k in rows[:45]:
        print(f"| {t / 1e3:.3f} | {100 * t / total:.1f} | {n} | `{k[:150]}` |")


if __name__ == "__main__":
    main()
