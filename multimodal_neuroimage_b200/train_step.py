"""One training step of a model built on the drop-in modules, the way the hot loop of the reference's trainer runs it
(trainer.py:363-453: zero_grad -> autocast forward -> loss -> backward -> [DDP all-reduce] -> optimizer step), restructured
for a B200: no host synchronisation inside the step (the reference calls `.item()` / `.cpu()` several times
per step, trainer.py:548,560,563,763-773), forward + backward replayed from ONE CUDA graph, gradients living in one flat
fp32 buffer so that the data-parallel exchange is a single NCCL all-reduce over NVLink (trainer.py:280-290's DDP moves
the same bytes in 25 MB buckets), optimizer step replayed from a second graph.

A step launches ~3000 kernels for the cfg3 model; issued eagerly the host (Python + ~70 us per custom-op call) is the
bottleneck, not the GPU.  Graph replay removes that; `use_graph=False` gives the eager step for comparison, and
`ddp="torch"` wraps the model in torch's DistributedDataParallel instead (bucketed overlap, eager).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn


class TrainStep:
    def __init__(self, model: nn.Module, loss_fn: Callable, example_inputs: Sequence[torch.Tensor], example_target: torch.Tensor,
                 lr: float = 1e-4, weight_decay: float = 0.05, autocast_dtype=torch.bfloat16, world: int = 1,
                 use_graph: bool = True, ddp: str = "flat", warmup: int = 3, bf16_weight_copies: bool = True):
        self.model, self.loss_fn, self.world, self.dtype = model, loss_fn, world, autocast_dtype
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.dev = dev
        self.inputs = [torch.empty_like(t, device=dev) for t in example_inputs]          # static device buffers
        self.target = torch.empty_like(example_target, device=dev)
        self.use_graph, self.ddp = use_graph, ddp
        self.fwd_model = model
        if ddp == "torch" and world > 1:
            from torch.nn.parallel import DistributedDataParallel as DDP
            self.fwd_model = DDP(model, device_ids=[dev.index], gradient_as_bucket_view=True, broadcast_buffers=False,
                                 static_graph=True)
            self.use_graph = False
        # all gradients in one flat fp32 buffer (the exchange and the optimizer read it through per-parameter views).
        # Backward does not accumulate into the views -- that costs one tiny add kernel per parameter and step (~700 for
        # cfg3) plus the zeroing; it leaves fresh gradients (p.grad is None going in) and ONE batched copy gathers them.
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=torch.float32)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, fused=True, capturable=self.use_graph)
        self.loss = torch.zeros((), device=dev)
        # bf16 copies of the weights, refreshed by one multi-tensor copy after every optimizer step instead of one cast
        # kernel per weight inside the forward (ops.weight_bf16)
        from . import ops
        self.shadows = ops.Bf16Shadows(self.params) if autocast_dtype == torch.bfloat16 and bf16_weight_copies else None
        self.g_fb = self.g_opt = None
        self._load(example_inputs, example_target)
        if world > 1 and self.fwd_model is model:
            for p in model.parameters():
                dist.broadcast(p.data, 0)
            if self.shadows is not None:          # `.data` updates do not move the version counter the copies are guarded by
                self.shadows.refresh()
        if self.use_graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(warmup):                      # allocator / autotune / lazy-init warm-up off the capture
                    self._fwd_bwd()
                    self._exchange()
                    self._opt_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            self.g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fb):
                self._fwd_bwd()
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
                self._opt_step()

    # -- pieces -------------------------------------------------------------------------------------------------------
    def _load(self, inputs, target):
        for dst, src in zip(self.inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.target.copy_(target, non_blocking=True)

    def _fwd_bwd(self):
        own = self.fwd_model is self.model
        if own:
            for p in self.params:
                p.grad = None
        else:
            self.opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=self.dtype):
            out = self.fwd_model(*self.inputs)
        loss = self.loss_fn(out.float(), self.target)
        loss.backward()
        self.loss.copy_(loss.detach())
        if own:
            torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params], out=self.flat)
            for p, v in zip(self.params, self.views):
                p.grad = v

    def _opt_step(self):
        self.opt.step()
        if self.shadows is not None:
            self.shadows.refresh()

    def _exchange(self):
        """The path's one collective: average the weight gradients over the ranks."""
        if self.world > 1 and self.fwd_model is self.model:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)

    # -- one step -----------------------------------------------------------------------------------------------------
    def __call__(self, inputs: Optional[Sequence[torch.Tensor]] = None, target: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`inputs` / `target`: this step's batch (pinned host or device tensors), copied into the static buffers; None
        re-uses what is there.  Returns the loss as a DEVICE scalar (read it with .item() only when you need it)."""
        if inputs is not None:
            self._load(inputs, target)
        if self.g_fb is not None:
            self.g_fb.replay()
            self._exchange()
            self.g_opt.replay()
        else:
            self._fwd_bwd()
            self._exchange()
            self._opt_step()
        return self.loss

    def grad_bytes(self) -> int:
        return self.flat.numel() * 4
