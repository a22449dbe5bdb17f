"""world_size-2 CPU (gloo) tests of the host-side multi-rank logic (the path
shards by batch only; the one collective is the gradient all-reduce that trainer.py's DDP
performs).  No GPU, no kernels: ranks shard a global batch with geometry.shard_range, run the
CPU oracle on their shard, all-reduce parameter gradients, and must reproduce the
single-process result -- which is the property bench.py --gpus N relies on (weak scaling,
no data-path collective)."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_neuroimage_b200 import geometry
from oracle import ref_nd as R


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    g = torch.Generator().manual_seed(0)
    grid, w, s, C, nH, B = (8, 8), 4, 2, 16, 2, 6
    N = w * w
    p = {
        "qkv.weight": torch.randn(3 * C, C, generator=g, dtype=torch.float64) * 0.3,
        "qkv.bias": torch.randn(3 * C, generator=g, dtype=torch.float64) * 0.1,
        "proj.weight": torch.randn(C, C, generator=g, dtype=torch.float64) * 0.3,
        "proj.bias": torch.zeros(C, dtype=torch.float64),
        "relative_position_bias_table": torch.randn((2 * w - 1) ** 2, nH, generator=g, dtype=torch.float64),
        "relative_position_index": geometry.relative_position_index((w, w)),
    }
    x = torch.randn(B, math.prod(grid), C, generator=g, dtype=torch.float64)
    return grid, w, s, nH, p, x


def _loss_and_grads(p, x, grid, w, s, nH):
    """Mean-over-samples loss of the windowed attention on x; returns (loss_sum, n, grads)."""
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in p.items()}
    B, L, C = x.shape
    xs = R.cyclic_shift_nd(x.view(B, *grid, C), (s, s))
    xw = R.window_partition_nd(xs, (w, w)).view(-1, w * w, C)
    y = R.window_attention_scaled(xw, params, (w, w), nH, R.shift_mask_nd(grid, (w, w), (s, s), x.dtype))
    loss = (y ** 2).sum()
    wrt = [v for v in params.values() if v.requires_grad]
    return loss.detach(), torch.autograd.grad(loss, wrt)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        grid, w, s, nH, p, x = _problem()
        lo, hi = geometry.shard_range(x.shape[0], world, rank)
        loss, grads = _loss_and_grads(p, x[lo:hi], grid, w, s, nH)
        dist.all_reduce(loss)
        for g in grads:
            dist.all_reduce(g)          # sum of per-shard gradients == gradient of the global-batch loss
        if rank == 0:
            torch.save({"loss": loss, "grads": grads, "shard": (lo, hi)}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_batch_sharding_matches_single_process(tmp_path):
    world = 2
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    grid, w, s, nH, p, x = _problem()
    loss, grads = _loss_and_grads(p, x, grid, w, s, nH)
    torch.testing.assert_close(got["loss"], loss, rtol=1e-12, atol=1e-12)
    for a, b in zip(got["grads"], grads):
        torch.testing.assert_close(a, b, rtol=1e-10, atol=1e-12)
    assert got["shard"] == (0, 3)


def test_bench_workload_is_rank_independent():
    """bench.py's per-rank workload (weak scaling): every rank processes `batch` volumes, so the
    units processed by N ranks are N * batch * windows_per_sample, with no data exchanged."""
    import bench
    assert bench.WINDOWS_PER_SAMPLE == 512 and bench.FLOP_PER_WINDOW == 18874368
    for world in (1, 2, 4, 8):
        cfg = bench.workload_config(32, world)
        assert cfg["windows_per_step"] == 32 * world * 512 and cfg["parallelism"] == f"dp{world}"
