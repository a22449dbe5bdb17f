"""Auxiliary measurement for SURVEY.md 8a rows a12-a15 (BASELINE configs[0], "cfg1"): the cross-modal
TransformerEncoder of the reference's sex-classification model -- 12 pre-LN layers, E = 84, 12 heads (head_dim 7),
T = S = 368, batch 2, causal mask, fp32 -- forward + backward through the drop-in modules (generic CUDA kernels
behind mmn_mha_fwd / mmn_mha_bwd), next to the CPU oracle port on the host cores.  Prints one JSON object.

  python tools/bench_mha.py > profiles/<round>_mha_cfg1.json
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_neuroimage_b200 import _lib  # noqa: E402
from multimodal_neuroimage_b200.modules import crossmodal_transformer as cm  # noqa: E402
from oracle import ref_nd as R  # noqa: E402

T, B, E, H, LAYERS = 368, 2, 84, 12, 12
torch.manual_seed(0)
enc = cm.TransformerEncoder(E, H, LAYERS, attn_mask=True).eval()
xq, xk = torch.randn(T, B, E), torch.randn(T, B, E)
flops_attn = LAYERS * 3 * 4 * T * T * E * B                    # QK^T + PV, fwd + bwd (no recompute)

# --- CPU oracle port
sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
for v in sd.values():
    if v.is_floating_point():
        v.requires_grad_(True)
torch.set_num_threads(os.cpu_count() or 1)


def cpu_step():
    xi = xq.clone().requires_grad_(True)
    out = R.transformer_encoder(xi, sd, H, LAYERS, True, xk, xk)
    torch.autograd.grad(out.sum(), [xi] + [v for v in sd.values() if v.requires_grad], allow_unused=True)


cpu_step()
t0 = time.perf_counter()
for _ in range(3):
    cpu_step()
cpu_ms = (time.perf_counter() - t0) / 3 * 1e3

# --- GPU: drop-in modules
dev = torch.device("cuda:0")
_lib.load()
encg = enc.to(dev)
xqg, xkg = xq.to(dev), xk.to(dev)


def gpu_step():
    for p in encg.parameters():
        p.grad = None
    xi = xqg.clone().requires_grad_(True)
    out = encg(xi, xkg, xkg)
    out.sum().backward()
    return out


before = _lib.launch_count()
for _ in range(5):
    gpu_step()
torch.cuda.synchronize()
launches = (_lib.launch_count() - before) // 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gpu_step()
e1.record()
torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 20
g = torch.cuda.CUDAGraph()
static_in = xqg.clone().requires_grad_(True)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        for p in encg.parameters():
            p.grad = None
        static_in.grad = None
        encg(static_in, xkg, xkg).sum().backward()
torch.cuda.current_stream().wait_stream(s)
for p in encg.parameters():
    p.grad = None
static_in.grad = None
graph_ms = None
try:
    with torch.cuda.graph(g):
        encg(static_in, xkg, xkg).sum().backward()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / 50
except Exception as exc:  # noqa: BLE001
    print(f"graph capture failed: {exc}", file=sys.stderr)

with torch.no_grad():
    want = R.transformer_encoder(xq.double(), {k: (v.detach().double() if v.is_floating_point() else v) for k, v in sd.items()},
                                 H, LAYERS, True, xk.double(), xk.double())
    got = encg(xqg, xkg, xkg).double().cpu()
rel = float((got - want).abs().max() / want.abs().max())
print(json.dumps({"workload": "cfg1 cross-modal TransformerEncoder: 12 layers, E=84, 12 heads x 7, T=S=368, batch 2, causal mask, fp32, fwd+bwd",
                  "gpu_ms_per_step_eager": gpu_ms, "gpu_ms_per_step_cuda_graph": graph_ms, "our_kernel_launches_per_step": launches,
                  "cpu_port_ms_per_step": cpu_ms, "cpu_cores": os.cpu_count(),
                  "attention_gflop_per_step": flops_attn / 1e9, "forward_rel_err_vs_fp64_oracle": rel,
                  "note": "head_dim 7 and 24 (batch x head) pairs: latency-bound on any GPU; the generic fp32 kernels are used "
                          "(no tensor-core path at these sizes)"}))
