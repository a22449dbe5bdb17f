#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph replay of the cfg2 module step (what bench.py times): start offset, duration and the
gap before every kernel, from torch.profiler (CUPTI).  Shows what the step spends outside its four big kernels.
    python tools/graph_timeline.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200.modules import swin_v2_module as v2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
C, WINDOW, HEADS, GRID, SHIFT = 96, 4, 3, (32, 32, 32), 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
attn = v2.WindowAttention(C, (WINDOW,) * 3, HEADS).to(dev)
L = 32 ** 3
x = torch.randn(B, L, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
dy = torch.randn(B, L, C, device=dev, dtype=torch.bfloat16)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = attn.forward_grid(x, GRID, (SHIFT,) * 3)
    y.backward(dy)
    return y


torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
side = torch.cuda.Stream(device=dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    for _ in range(3):
        for p in attn.parameters():
            p.grad = None
        x.grad = None
        step()
torch.cuda.current_stream(dev).wait_stream(side)
for p in attn.parameters():
    p.grad = None
x.grad = None
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(10):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"replay: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per step")
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "emcpy" not in e.name and "emset" not in e.name or
              (e.device_type == torch.autograd.DeviceType.CUDA)), key=lambda e: e.time_range.start)
# the last replay: kernels after the last long gap
n = len(evs) // 3
evs = evs[-n:]
t0, prev_end, busy = evs[0].time_range.start, evs[0].time_range.start, 0.0
print(f"{'start':>8} {'dur':>7} {'gap':>6}  kernel")
for e in evs:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    print(f"{s:8.1f} {d:7.1f} {e.time_range.start - prev_end:6.1f}  {e.name[:90]}")
    prev_end = max(prev_end, e.time_range.end)
    busy += d
print(f"span {prev_end - t0:.1f} us, sum of kernel durations {busy:.1f} us, {len(evs)} kernels")
