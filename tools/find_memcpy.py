#!/usr/bin/env python
"""Which host ops of a training step issue device-to-device MEMCPYs (or memsets)?  Inside a CUDA graph a memcpy / memset
node between kernel nodes costs ~8 us of dependency latency (a kernel: ~1 us), so they are worth replacing by kernels.
One eager step of a workload under torch.profiler with Python stacks; prints, per source line, how many memcpys it issues.
    python tools/find_memcpy.py [cfg3|cfg4|cfg5] [batch]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import train_step as TS  # noqa: E402
from multimodal_neuroimage_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (1 if name == "cfg5" else 2)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = (W.SwinFusion3D() if name == "cfg3" else W.FuncStructCross3D() if name == "cfg4" else W.SwinV2CrossModal3D()).to(dev)
img = 128 if name == "cfg5" else 96
if name == "cfg4":
    *ins, y = W.synthetic_batch_cfg4(batch, img, dev)
else:
    *ins, y = W.synthetic_batch(batch, img, dev)
ts = TS.TrainStep(model, torch.nn.functional.binary_cross_entropy_with_logits, tuple(ins), y, use_graph=False)
for _ in range(2):
    ts()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    ts._fwd_bwd()
    torch.cuda.synchronize()
by_site = collections.Counter()
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CPU or not e.kernels:
        continue
    n = sum(1 for k in e.kernels if "emcpy" in k.name or "emset" in k.name)
    if not n:
        continue
    stack = [s for s in (e.stack or []) if "multimodal_neuroimage_b200" in s or "autograd" in s][:3]
    chain, p = [e.name], e.cpu_parent
    while p is not None and len(chain) < 5:
        chain.append(p.name)
        p = p.cpu_parent
    by_site[(" <- ".join(chain), " | ".join(s.split("/")[-1] for s in stack))] += n
for (chain, stack), n in by_site.most_common(25):
    print(f"{n:4d}  {chain}\n        {stack}")
print("total", sum(by_site.values()))
