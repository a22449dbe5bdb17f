"""GPU parity tests added in round 2 (VERDICT r1 "What's weak" #1/#6): full-size cfg2 against the oracle
itself, cfg1-size and bf16 multi-head attention, cfg5 stage head counts, the in-kernel shift mask pinned bit-exactly,
fp16 autocast + GradScaler as the reference trainer runs it (trainer.py:84,378-409), colsum and large CPB tables,
the eval-time CPB cache, and the residual groups (BasicLayer / RSTB / CRSTB) of SURVEY.md 8a row a11.
"""
import math
import os

import pytest
import torch

from oracle import ref_nd as R
from test_gpu_parity import (BF16_TOL, FP32_TOL, _core_inputs, _oracle_core, _randomise, _run_core, _sd64, check, mm,  # noqa: F401
                                   rel_err)

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------
# cfg2 at FULL size (B=1, 32^3 tokens, 512 windows, 3 heads x 32): tcgen05 path vs the fp64 oracle directly
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cross", [False, True], ids=["self", "cross"])
def test_cfg2_full_size_tcgen05_vs_oracle(mm, cross):
    case = ((32, 32, 32), (4, 4, 4), (2, 2, 2), 3, 32, True, "shift", 1)
    a, b, *_ = _core_inputs(case, torch.bfloat16, cross)
    assert mm.ops.winattn_path_name(a.cuda().bfloat16(), None if b is None else b.cuda().bfloat16(), case[0], case[1], case[2],
                                    3, mm.lib.SCORE_COSINE, mm.lib.MASK_SHIFT) == "tcgen05"
    _run_core(mm, case, torch.bfloat16, cross, mm.lib.PATH_AUTO)


# cfg5 stage shapes (embed 192 -> 1536, heads C/32 = 6/12/24/48, 4x4x4 windows; windows/sample shrink with the stage)
@pytest.mark.parametrize("grid,nH", [((16, 8, 8), 6), ((8, 8, 8), 12), ((8, 8, 4), 24), ((4, 4, 4), 48)],
                         ids=["C192", "C384", "C768", "C1536"])
def test_cfg5_stage_heads_tcgen05(mm, grid, nH):
    shift = tuple(2 if g > 4 else 0 for g in grid)
    case = (grid, (4, 4, 4), shift, nH, 32, True, "shift" if any(shift) else "none", 2)
    a, *_ = _core_inputs(case, torch.bfloat16, False)
    kind = mm.lib.MASK_SHIFT if any(shift) else mm.lib.MASK_NONE
    assert mm.ops.winattn_path_name(a.cuda().bfloat16(), None, grid, (4, 4, 4), shift, nH, mm.lib.SCORE_COSINE, kind) == "tcgen05"
    _run_core(mm, case, torch.bfloat16, False, mm.lib.PATH_AUTO)


# ------------------------------------------------------------------------------------------
# Shift mask generated in the kernel == the oracle's {0,-100} mask tensor, BIT for BIT on one code path
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("grid,window,shift,nH,d", [((12, 12), (6, 6), (3, 3), 3, 4), ((8, 16), (4, 4), (2, 2), 2, 16),
                                                    ((8, 8, 8), (4, 4, 4), (2, 2, 2), 3, 32), ((4, 8, 12), (2, 4, 4), (1, 2, 2), 2, 8)],
                         ids=["2d_w6", "2d_w4", "3d_w4", "3d_aniso"])
def test_in_kernel_shift_mask_bit_exact_generic(mm, grid, window, shift, nH, d):
    """north_star: masks bit-exact.  MMN_MASK_SHIFT (region ids from coordinates in the kernel) must give the same bits as
    MMN_MASK_TENSOR fed with the oracle's mask (swin_v2_module.py:244-266 restated n-D), forward AND backward."""
    g = torch.Generator().manual_seed(4)
    C, N, B = nH * d, math.prod(window), 2
    qkv = torch.randn(B, *grid, 3 * C, generator=g).cuda().requires_grad_(True)
    bias = torch.randn(nH, N, N, generator=g).cuda()
    hs = (torch.rand(nH, generator=g) * 20 + 0.5).cuda()
    mask = R.shift_mask_nd(grid, window, shift, torch.float32).cuda().contiguous()
    dout = torch.randn(B, *grid, C, generator=g).cuda()
    res = []
    for kind, m in ((mm.lib.MASK_SHIFT, None), (mm.lib.MASK_TENSOR, mask)):
        out, lse = torch.ops.mmn_b200.winattn_fwd(qkv, None, bias, hs, m, list(grid), list(window), list(shift), nH,
                                                  mm.lib.SCORE_COSINE, kind, 1.0, 0.0, 0, 0, mm.lib.PATH_GENERIC)
        dq, = torch.autograd.grad((out * dout).sum(), qkv)
        res.append((out, lse[0], dq))
    for x, y, name in zip(res[0], res[1], ("out", "lse", "dqkv")):
        assert torch.equal(x, y), f"{name}: in-kernel shift mask differs from the oracle mask tensor"


@pytest.mark.parametrize("grid,window,shift", [((24, 32), (8, 8), (4, 4)), ((12, 12, 12), (4, 4, 4), (2, 2, 2)), ((8, 8, 4), (4, 4, 4), (2, 2, 0))],
                         ids=["2d", "3d_all_classes", "3d_one_axis_unshifted"])
def test_in_kernel_shift_mask_bit_exact_tcgen05(mm, grid, window, shift):
    """Same on the tensor-core path, where the mask is folded per wrap class into the shared-memory table (tc_sched.cuh:
    class_region_id).  Reference = the same kernel on the PRE-ROLLED volume, shift 0, fed the oracle's mask tensor (no
    window wraps there, so tile order = window order).  With no bias both routes add exactly -100*log2(e) or 0 to a logit:
    windows that do not wrap must agree bit for bit; wrapped windows sum their keys in piece-major order, so they agree to
    fp32-accumulation rounding of bf16 outputs only."""
    g = torch.Generator().manual_seed(5)
    nH, d, B, n = 3, 32, 2, len(grid)
    C = nH * d
    qkv = torch.randn(B, *grid, 3 * C, generator=g).bfloat16().cuda()
    mask = R.shift_mask_nd(grid, window, shift, torch.float32).cuda().contiguous()
    assert mm.ops.winattn_path_name(qkv, None, grid, window, shift, nH, mm.lib.SCORE_SCALED, mm.lib.MASK_SHIFT) == "tcgen05"
    zero = [0] * n
    assert mm.ops.winattn_path_name(qkv, None, grid, window, zero, nH, mm.lib.SCORE_SCALED, mm.lib.MASK_TENSOR) == "tcgen05"
    # a mask tensor plus a shift is not a tensor-core combination (tile order != window order for wrapped windows)
    assert mm.ops.winattn_path_name(qkv, None, grid, window, shift, nH, mm.lib.SCORE_SCALED, mm.lib.MASK_TENSOR) == "generic"
    dims = tuple(range(1, n + 1))
    out_s, lse_s = torch.ops.mmn_b200.winattn_fwd(qkv, None, None, None, None, list(grid), list(window), list(shift), nH,
                                                  mm.lib.SCORE_SCALED, mm.lib.MASK_SHIFT, d ** -0.5, 0.0, 0, 0, mm.lib.PATH_TCGEN05)
    rolled = torch.roll(qkv, [-s for s in shift], dims).contiguous()
    out_t, lse_t = torch.ops.mmn_b200.winattn_fwd(rolled, None, None, None, mask, list(grid), list(window), zero, nH,
                                                  mm.lib.SCORE_SCALED, mm.lib.MASK_TENSOR, d ** -0.5, 0.0, 0, 0, mm.lib.PATH_TCGEN05)
    out_sr = torch.roll(out_s, [-s for s in shift], dims)
    inner = torch.ones(grid, dtype=torch.bool, device="cuda")       # tokens (rolled frame) of windows that do not wrap
    for a in range(n):
        if shift[a]:
            idx = torch.arange(grid[a], device="cuda") // window[a] < grid[a] // window[a] - 1
            inner &= idx.view([-1 if i == a else 1 for i in range(n)])
    assert inner.any() and (~inner).any()
    assert torch.equal(out_sr[:, inner], out_t[:, inner]), "unwrapped windows: in-kernel mask differs from the oracle mask tensor"
    assert rel_err(out_sr, out_t) < 1e-2
    assert torch.equal(lse_s[0].isfinite(), lse_t[0].isfinite())


# ------------------------------------------------------------------------------------------
# Multi-head attention: cfg1 sizes (T = S = 368, 12 heads, d = 7 cross / d = 14 self, causal), fp32 and bf16
# ------------------------------------------------------------------------------------------
def _mha_case(mm, E, nH, T, S, B, cross, causal, dtype, tol):
    g = torch.Generator().manual_seed(E + T)
    m = mm.mh.MultiheadAttention(E, nH)
    with torch.no_grad():
        m.in_proj_bias.normal_(0, 0.3, generator=g)
        m.out_proj.bias.normal_(0, 0.3, generator=g)
    q = torch.randn(T, B, E, generator=g)
    k = torch.randn(S, B, E, generator=g) if cross else q
    v = torch.randn(S, B, E, generator=g) if cross else q
    cot = torch.randn(T, B, E, generator=g)
    if dtype == torch.bfloat16:
        q, k, v = q.bfloat16().float(), k.bfloat16().float(), v.bfloat16().float()
    sd = {n: p.detach().double() for n, p in m.state_dict().items()}
    ins = [t.double().requires_grad_(True) for t in ((q, k, v) if cross else (q,))]
    qo, ko, vo = ins if cross else (ins[0], ins[0], ins[0])
    want, wavg = R.multihead_attention(qo, ko, vo, sd, nH, R.future_mask(T, S, torch.float64) if causal else None)
    gwant = torch.autograd.grad((want * cot.double()).sum(), ins)
    m = m.cuda()
    insc = [t.cuda().requires_grad_(True) for t in ((q, k, v) if cross else (q,))]
    qc, kc, vc = insc if cross else (insc[0], insc[0], insc[0])
    mask = mm.cm.buffered_future_mask(qc, kc) if causal else None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        got, gavg = m(qc, kc, vc, attn_mask=mask, need_weights=True)
    ggot = torch.autograd.grad((got.float() * cot.cuda()).sum(), insc + [m.in_proj_weight, m.in_proj_bias])
    check(got, want, tol, "mha out")
    check(gavg, wavg, tol, "mha averaged weights")
    for a, b in zip(ggot[:len(ins)], gwant):
        check(a, b, tol * 3, "mha d input")
    assert all(torch.isfinite(t).all() for t in ggot)


@pytest.mark.parametrize("E,cross", [(84, True), (168, False)], ids=["cross_E84_d7", "self_E168_d14"])
def test_mha_cfg1_size_fp32(mm, E, cross):
    _mha_case(mm, E, 12, 368, 368, 2, cross, True, torch.float32, FP32_TOL)


@pytest.mark.parametrize("E,nH,T,S,cross,causal", [(84, 12, 368, 368, True, True), (168, 12, 368, 368, False, True),
                                                  (256, 4, 200, 333, True, False), (64, 2, 96, 96, False, True)],
                         ids=["cfg1_cross_d7", "cfg1_self_d14", "d64_TneS", "d32_causal"])
def test_mha_bf16(mm, E, nH, T, S, cross, causal, monkeypatch):
    # every bf16 case must reach the tensor-core kernels: d = 32 / 64 directly, the reference's own d = 7 / 14 through the
    # zero-padded projections (multihead_attention.py: _padded_head_dim)
    import ctypes
    lib, seen = mm.lib.load(), []
    real = lib.mmn_mha_fwd

    def spy(dref, *a):
        d = ctypes.cast(dref, ctypes.POINTER(mm.lib.MhaDesc)).contents
        seen.append((d.head_dim, lib.mmn_mha_path(dref).decode()))
        return real(dref, *a)
    monkeypatch.setattr(lib, "mmn_mha_fwd", spy)
    _mha_case(mm, E, nH, T, S, 2, cross, causal, torch.bfloat16, BF16_TOL)
    assert seen and all(p == "tcgen05" and hd in (32, 64) for hd, p in seen), seen


# ------------------------------------------------------------------------------------------
# fp16 autocast + GradScaler, exactly as the reference trainer drives the model (trainer.py:84,378,385,402-409)
# ------------------------------------------------------------------------------------------
def test_fp16_autocast_gradscaler_like_the_reference_trainer(mm):
    grid, C, nH, B = (8, 8, 8), 96, 3, 2
    blk = mm.v2.SwinTransformerBlock(C, grid, nH, window_size=4, shift_size=2)
    cross = mm.fu.Cross_SwinTransformerBlock(C, grid, nH, window_size=4, shift_size=2)
    enc = mm.cm.TransformerEncoderLayer(84, num_heads_mult=12, attn_dropout=0.0, relu_dropout=0.0, res_dropout=0.0, attn_mask=True)
    for i, m in enumerate((blk, cross, enc)):
        _randomise(m, 20 + i)
    with torch.no_grad():
        # SwinV2's own initial logit scale (swin_v2_module.py:89).  The clamp range up to 100 is covered in fp32 (test_swinv2_block_3d);
        # at g = 100 a 16-bit q/k operand moves a logit by ~0.1 whatever the kernel (tools/debug_fp16.py: same error under
        # bf16 autocast), which says nothing about the fp16 / GradScaler plumbing this test is about.
        blk.attn.logit_scale.fill_(math.log(10.0))
    g = torch.Generator().manual_seed(2)
    x, y = torch.randn(B, math.prod(grid), C, generator=g), torch.randn(B, math.prod(grid), C, generator=g)
    s = torch.randn(40, B, 84, generator=g)
    xo, yo, so = (t.double().requires_grad_(True) for t in (x, y, s))
    w1 = R.swin_v2_block(xo, _sd64(blk), grid, 4, 2, nH)
    w2a, w2b = R.cross_block(xo, yo, _sd64(cross), grid, grid, 4, 2, nH)
    w3 = R.encoder_layer(so, _sd64(enc), 12, True)
    gw = torch.autograd.grad(w1.mean() + w2a.mean() + w2b.mean() + w3.mean(), (xo, yo, so))   # a mean, like the trainer's BCE

    blk, cross, enc = blk.cuda(), cross.cuda(), enc.cuda().eval()
    xc, yc, sc_ = (t.cuda().requires_grad_(True) for t in (x, y, s))
    params = [p for m in (blk, cross, enc) for p in m.parameters()]
    opt = torch.optim.SGD(params, lr=0.0)
    scaler = torch.amp.GradScaler('cuda')
    opt.zero_grad()
    with torch.amp.autocast('cuda'):                      # fp16 (torch.cuda.amp.autocast() of trainer.py:378), the reference's mode (main.py:88: --amp defaults to on)
        o1 = blk(xc)
        o2a, o2b = cross(xc, yc, grid)
        o3 = enc(sc_)
        assert o1.dtype == torch.float32 or o1.dtype == torch.float16
        loss = o1.float().mean() + o2a.float().mean() + o2b.float().mean() + o3.float().mean()
    scale = scaler.get_scale()
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    scaler.step(opt)
    scaler.update()
    assert scaler.get_scale() >= 65536.0, "GradScaler saw inf/nan gradients"
    check(o1, w1, 3e-2, "fp16-autocast swinv2 block")
    check(o2a, w2a, 3e-2, "fp16-autocast cross block A")
    check(o2b, w2b, 3e-2, "fp16-autocast cross block B")
    check(o3, w3, 3e-2, "fp16-autocast encoder layer")
    for got, want, name in zip((xc.grad, yc.grad, sc_.grad), gw, ("dx", "dy", "ds")):     # inputs are not optimizer params: still scaled
        check(got / scale, want, 5e-2, "fp16-autocast " + name)
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in params)


# ------------------------------------------------------------------------------------------
# colsum for cols > 256; CPB tables whose backward needs > 48 KB of shared memory
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cols", [96, 192, 384, 768, 1536, 2048])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_colsum(mm, cols, dtype):
    for rows in (1, 37, 4096 + 5):
        x = torch.randn(rows, cols, generator=torch.Generator().manual_seed(rows + cols)).to(dtype).cuda()
        got = torch.ops.mmn_b200.colsum(x)
        want = x.double().sum(0)
        assert (got.double() - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item()) * (1 if dtype == torch.float32 else 4)
    wide = torch.randn(300, cols + 64, generator=torch.Generator().manual_seed(cols)).to(dtype).cuda()
    got = torch.ops.mmn_b200.colsum(wide[:, :cols])                      # strided rows
    assert (got.double() - wide[:, :cols].double().sum(0)).abs().max().item() <= 1e-3


@pytest.mark.parametrize("window,nH", [((16, 16), 6), ((6, 6, 6), 3), ((7, 7, 7), 3), ((24, 24), 4)],
                         ids=["2d_w16_T961", "3d_w6_T1331", "3d_w7_T2197", "2d_w24_T2209"])
def test_cpb_bias_large_tables(mm, window, nH):
    from multimodal_neuroimage_b200 import geometry
    g = torch.Generator().manual_seed(nH + window[0])
    n = len(window)
    coords = geometry.cpb_coords_table(window).reshape(-1, n).float()
    index = geometry.relative_position_index(window).reshape(-1)
    w1, b1, w2 = torch.randn(512, n, generator=g) * 0.7, torch.randn(512, generator=g) * 0.3, torch.randn(nH, 512, generator=g) * 0.1
    # a sparse cotangent keeps the fp64 reference cheap for N*N up to 3.3e5 entries
    cot = torch.randn(nH, index.numel(), generator=g)
    pd = [t.double().requires_grad_(True) for t in (w1, b1, w2)]
    tab = torch.relu(coords.double() @ pd[0].t() + pd[1]) @ pd[2].t()
    want = (16 * torch.sigmoid(tab))[index].t()
    gw = torch.autograd.grad((want * cot.double()).sum(), pd)
    pc = [t.cuda().requires_grad_(True) for t in (w1, b1, w2)]
    assert mm.ops.cpb_bias_supported(coords.cuda(), pc[0], pc[2])
    got, _ = torch.ops.mmn_b200.cpb_bias_fwd(coords.cuda(), pc[0], pc[1], pc[2], index.cuda())
    gg = torch.autograd.grad((got * cot.cuda()).sum(), pc)
    check(got, want, FP32_TOL, "cpb bias")
    for a, b, nm in zip(gg, gw, ("dw1", "db1", "dw2")):
        check(a, b, 5e-5, "cpb " + nm)       # float-atomic scatter of up to 3e5 entries per table row


def test_position_bias_cached_outside_training(mm):
    wa = mm.v2.WindowAttention(96, (4, 4, 4), 3).cuda()
    from multimodal_neuroimage_b200 import _lib
    wa.train()
    n0 = _lib.launch_count()
    b_train = wa.position_bias()
    assert b_train.requires_grad and _lib.launch_count() > n0           # training: recomputed, differentiable
    wa.eval()
    with torch.no_grad():
        b1 = wa.position_bias()
        n1 = _lib.launch_count()
        b2 = wa.position_bias()
        assert b2 is b1 and _lib.launch_count() == n1                   # cached: no kernel
        assert torch.equal(b1, b_train.detach())
        wa.cpb_mlp[2].weight.mul_(1.5)                                  # in-place update (optimizer step / load_state_dict)
        b3 = wa.position_bias()
        assert b3 is not b1 and not torch.equal(b3, b1)
        sd = {k: v.clone() for k, v in wa.state_dict().items()}
        sd["cpb_mlp.0.bias"] += 0.1
        wa.load_state_dict(sd)
        assert not torch.equal(wa.position_bias(), b3)
    # eval mode but gradients wanted (parity tests do this): not served from the cache
    assert wa.position_bias().requires_grad


# ------------------------------------------------------------------------------------------
# a11: the stacks around the blocks -- BasicLayer (+PatchMerging), BasicLayer_fusion, Cross_BasicLayer, RSTB, CRSTB --
# against the n-D oracle composed the way the reference composes them (swin_v2_module.py:376-451,
# swinfusion_module.py:609-939)
# ------------------------------------------------------------------------------------------
def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


@pytest.mark.parametrize("grid", [(12, 12), (8, 8, 8)], ids=["2d", "3d"])
def test_basic_layer_swinv2(mm, grid):
    C, nH, depth, ws = 32, 2, 2, (6 if len(grid) == 2 else 4)
    layer = mm.v2.BasicLayer(C, grid, depth, nH, ws, downsample=mm.v2.PatchMerging)
    _randomise(layer, 3)
    x = torch.randn(2, math.prod(grid), C, generator=torch.Generator().manual_seed(1))
    sd = _sd64(layer)
    xo = x.double().requires_grad_(True)
    h = xo
    for i in range(depth):                                              # shift 0 / ws//2 alternate (swin_v2_module.py:410)
        h = R.swin_v2_block(h, _sub(sd, f"blocks.{i}."), grid, ws, 0 if i % 2 == 0 else ws // 2, nH)
    n = len(grid)
    hv = h.view(2, *grid, C)
    parts = [hv[tuple([slice(None)] + [slice((code >> a) & 1, None, 2) for a in range(n)] + [slice(None)])] for code in range(2 ** n)]
    hm = torch.cat(parts, -1).view(2, -1, (2 ** n) * C)                 # swin_v2_module.py:347-354 (x0..x3 order)
    hm = torch.nn.functional.linear(hm, sd["downsample.reduction.weight"])
    want = torch.nn.functional.layer_norm(hm, (2 * C,), sd["downsample.norm.weight"], sd["downsample.norm.bias"])
    gwant, = torch.autograd.grad(want.sum(), xo)
    layer = layer.cuda()
    xc = x.cuda().requires_grad_(True)
    got = layer(xc)
    ggot, = torch.autograd.grad(got.sum(), xc)
    check(got, want, FP32_TOL * 2, "BasicLayer out")
    check(ggot, gwant, FP32_TOL * 10, "BasicLayer dx")


@pytest.mark.parametrize("grid", [(12, 12), (8, 8, 8)], ids=["2d", "3d"])
def test_rstb_and_crstb(mm, grid):
    C, nH, depth, ws = 24, 3, 2, (6 if len(grid) == 2 else 4)
    rstb = mm.fu.RSTB(C, grid, depth, nH, ws, img_size=grid[0], patch_size=1)
    crstb = mm.fu.CRSTB(C, grid, depth, nH, ws, img_size=grid[0], patch_size=1)
    _randomise(rstb, 5)
    _randomise(crstb, 6)
    g = torch.Generator().manual_seed(8)
    x, y = torch.randn(2, math.prod(grid), C, generator=g), torch.randn(2, math.prod(grid), C, generator=g)
    xo, yo = x.double().requires_grad_(True), y.double().requires_grad_(True)

    def group(h, sd, prefix):                                            # BasicLayer_fusion (swinfusion_module.py:664-676)
        for i in range(depth):
            h = R.fusion_block(h, _sub(sd, f"{prefix}blocks.{i}."), grid, grid, ws, 0 if i % 2 == 0 else ws // 2, nH)
        return h

    sd = _sd64(rstb)
    want_r = group(xo, sd, "residual_group.") + xo                      # RSTB.forward (:814)
    sd = _sd64(crstb)
    xa = group(xo, sd, "residual_group_A.") + xo                        # CRSTB.forward (:916-928)
    yb = group(yo, sd, "residual_group_B.") + yo
    hx, hy = xa, yb
    for i in range(depth):
        hx, hy = R.cross_block(hx, hy, _sub(sd, f"residual_group.blocks.{i}."), grid, grid, ws, 0 if i % 2 == 0 else ws // 2, nH)
    want_cx, want_cy = hx + xa, hy + yb
    gw = torch.autograd.grad(want_r.sum() + want_cx.sum() + 2 * want_cy.sum(), (xo, yo))
    rstb, crstb = rstb.cuda(), crstb.cuda()
    xc, yc = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    got_r = rstb(xc, grid)
    got_cx, got_cy = crstb(xc, yc, grid)
    gg = torch.autograd.grad(got_r.sum() + got_cx.sum() + 2 * got_cy.sum(), (xc, yc))
    check(got_r, want_r, FP32_TOL * 2, "RSTB")
    check(got_cx, want_cx, FP32_TOL * 2, "CRSTB x")
    check(got_cy, want_cy, FP32_TOL * 2, "CRSTB y")
    check(gg[0], gw[0], FP32_TOL * 10, "RSTB/CRSTB dx")
    check(gg[1], gw[1], FP32_TOL * 10, "CRSTB dy")


# ------------------------------------------------------------------------------------------
# Tensor-core projections (csrc/gemm_tc.cu): forward with bias / activation epilogues, dgrad with the activation
# derivative as epilogue, split-token wgrad -- against fp64 on the same bf16 operands
# ------------------------------------------------------------------------------------------
LINEAR_SHAPES = [(1000, 96, 288), (4096 + 37, 96, 96), (777, 192, 576), (300, 384, 1536), (129, 1536, 384), (64, 768, 768),
                 (50, 32, 32), (200, 160, 224), (3000, 96, 384), (70000, 192, 192), (5, 64, 2048)]


@pytest.mark.parametrize("rows,n_in,n_out", LINEAR_SHAPES)
@pytest.mark.parametrize("act", ["none", "gelu", "relu"])
def test_linear_fwd_tensor_core(mm, rows, n_in, n_out, act):
    g = torch.Generator().manual_seed(rows + n_in + n_out)
    x = torch.randn(rows, n_in, generator=g).bfloat16()
    w = (torch.randn(n_out, n_in, generator=g) * n_in ** -0.5).bfloat16()
    b = torch.randn(n_out, generator=g) * 0.5 if act != "relu" else None
    assert mm.ops.linear_supported(x.cuda(), w.cuda())
    code = mm.ops._ACT[act]
    y, pre = torch.ops.mmn_b200.linear_fwd(x.cuda(), w.cuda(), None if b is None else b.cuda(), code, act != "none")
    want_pre = x.double() @ w.double().t() + (0 if b is None else b.double())
    want = {"none": lambda t: t, "gelu": torch.nn.functional.gelu, "relu": torch.relu}[act](want_pre)
    check(y, want, 1e-2, f"linear_fwd {act}")
    if act != "none":                                      # second output: act'(pre), what the backward multiplies with
        p64 = want_pre.clone().requires_grad_(True)
        want_d, = torch.autograd.grad(({"gelu": torch.nn.functional.gelu, "relu": torch.relu}[act](p64)).sum(), p64)
        away = want_pre.abs() > 0.05                       # relu' jumps at 0: compare where bf16 rounding cannot flip the sign
        check(pre.double().cpu() * away, want_d * away, 1e-2, "linear_fwd act'(pre)")
    # strided input rows (a channel slice of a wider tensor)
    wide = torch.randn(rows, n_in + 64, generator=g).bfloat16().cuda()
    y2, _ = torch.ops.mmn_b200.linear_fwd(wide[:, :n_in], w.cuda(), None, 0, False)
    check(y2, wide[:, :n_in].double().cpu() @ w.double().t(), 1e-2, "linear_fwd strided x")


@pytest.mark.parametrize("rows,n_in,n_out", LINEAR_SHAPES)
def test_linear_bwd_tensor_core(mm, rows, n_in, n_out):
    g = torch.Generator().manual_seed(rows + 3 * n_in + n_out)
    dy = torch.randn(rows, n_out, generator=g).bfloat16()
    x = torch.randn(rows, n_in, generator=g).bfloat16()
    w = (torch.randn(n_out, n_in, generator=g) * n_in ** -0.5).bfloat16()
    pre = torch.randn(rows, n_in, generator=g).bfloat16()
    assert mm.ops.linear_bwd_supported(dy.cuda(), x.cuda(), w.cuda())
    dyd, xd, wd = dy.double(), x.double(), w.double()
    dx, dw, db = torch.ops.mmn_b200.linear_bwd(dy.cuda(), x.cuda(), w.cuda())
    check(dx, dyd @ wd, BF16_TOL, "linear_bwd dx")
    check(dw, dyd.t() @ xd, 2e-3, "linear_bwd dw")
    check(db, dyd.sum(0), 2e-3, "linear_bwd db")
    for act in ("gelu", "relu"):                           # act_aux = act'(pre) as linear_fwd writes it
        dx2, dw2, _ = torch.ops.mmn_b200.linear_bwd(dy.cuda(), x.cuda(), w.cuda(), pre.cuda(), mm.ops._ACT[act], True, True)
        check(dx2, (dyd @ wd) * pre.double(), BF16_TOL, f"linear_bwd dx through {act}'")
        check(dw2, dyd.t() @ xd, 2e-3, "linear_bwd dw (general path)")
    only_dx = torch.ops.mmn_b200.linear_bwd(dy.cuda(), x.cuda(), w.cuda(), None, 0, True, False)
    assert only_dx[1].numel() == 0 and rel_err(only_dx[0], dyd @ wd) < BF16_TOL
    only_dw = torch.ops.mmn_b200.linear_bwd(dy.cuda(), x.cuda(), w.cuda(), None, 0, False, True)
    assert only_dw[0].numel() == 0 and rel_err(only_dw[1], dyd.t() @ xd) < 2e-3


@pytest.mark.parametrize("rep", range(int(os.environ.get("MMN_REPEAT", "1"))))     # MMN_REPEAT=n: n draws of the weights
@pytest.mark.parametrize("C,act", [(96, "gelu"), (192, "gelu"), (768, "relu")])
def test_fused_mlp_matches_pytorch(mm, C, act, rep):
    """fused.mlp (two GEMMs + epilogues, three-pass backward) against the same Mlp evaluated by PyTorch in fp64.

    relu'(pre) is discontinuous at 0: of the 6.4 M pre-activations of the C = 768 case a handful lie within fp32 round-off
    of 0, and there the kernel's sign may legitimately differ from the fp64 one.  One such flip moves one row of dW1 by a
    single token's contribution, 2-4 % of max |dW1| -- a sporadic failure of the max-norm check (seen in about one run in
    five) that is no kernel error.  So the reference takes the kernel's own act'(pre) wherever the two disagree, and the
    test asserts that they disagree only within round-off of 0."""
    from multimodal_neuroimage_b200 import fused
    F = torch.nn.functional
    g = torch.Generator().manual_seed(C)
    fc1, fc2 = torch.nn.Linear(C, 4 * C), torch.nn.Linear(4 * C, C)
    x = torch.randn(3, 700, C, generator=g).bfloat16().float()
    cot = torch.randn(3, 700, C, generator=g)
    ps = [p.detach().bfloat16().double().requires_grad_(True) if p.dim() == 2 else p.detach().double().requires_grad_(True)
          for p in (fc1.weight, fc1.bias, fc2.weight, fc2.bias)]
    xo = x.double().requires_grad_(True)
    fc1, fc2 = fc1.cuda(), fc2.cuda()
    xc = x.cuda().requires_grad_(True)
    pre = F.linear(xo, ps[0], ps[1])
    if act == "gelu":
        hid = F.gelu(pre)
    else:
        with torch.no_grad():
            _, dact = torch.ops.mmn_b200.linear_fwd(xc.detach().bfloat16().reshape(-1, C), fc1.weight.detach().bfloat16(),
                                                    fc1.bias.detach().float(), mm.lib.ACT_RELU, True)
        on = dact.double().cpu().view_as(pre)
        assert ((on == 0) | (on == 1)).all()
        flipped = on != (pre.detach() > 0)
        assert flipped.sum() <= 16 and (pre.detach().abs()[flipped] < 1e-5).all(), "relu'(pre) differs away from pre = 0"
        hid = pre * on
    want = F.linear(hid, ps[2], ps[3])
    gw = torch.autograd.grad((want * cot.double()).sum(), [xo] + ps)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got = fused.mlp(xc, fc1, fc2, act)
    assert got.dtype == torch.bfloat16
    gg = torch.autograd.grad((got.float() * cot.cuda()).sum(), [xc, fc1.weight, fc1.bias, fc2.weight, fc2.bias])
    check(got, want, BF16_TOL, "mlp out")
    for a, b, n in zip(gg, gw, ("dx", "dw1", "db1", "dw2", "db2")):
        check(a, b, BF16_TOL, "mlp " + n)


# ------------------------------------------------------------------------------------------
# LayerNorm fused with the residual add (csrc/layernorm.cu) against torch.nn.functional.layer_norm in fp64
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cols", [12, 84, 96, 168, 192, 384, 768, 1536])
@pytest.mark.parametrize("stream_dtype,act_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)],
                         ids=["f32_f32", "f32_bf16", "bf16_bf16"])
def test_fused_layernorm(mm, cols, stream_dtype, act_dtype):
    from multimodal_neuroimage_b200 import fused
    F = torch.nn.functional
    rows = (3, 211)
    g = torch.Generator().manual_seed(cols)
    ln = torch.nn.LayerNorm(cols)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5, generator=g)
        ln.bias.normal_(0, 0.3, generator=g)
    resid = torch.randn(*rows, cols, generator=g).to(stream_dtype)
    delta = (torch.randn(*rows, cols, generator=g) * 2 + 0.5).to(act_dtype)
    c1, c2 = torch.randn(*rows, cols, generator=g), torch.randn(*rows, cols, generator=g)
    tol = FP32_TOL * 2 if act_dtype == torch.float32 and stream_dtype == torch.float32 else BF16_TOL
    W, Bv = ln.weight.detach().double(), ln.bias.detach().double()

    def ref(mode):
        r, d = resid.double().requires_grad_(True), delta.double().requires_grad_(True)
        w, b = W.clone().requires_grad_(True), Bv.clone().requires_grad_(True)
        if mode == "ln":
            outs = (F.layer_norm(r, (cols,), w, b, ln.eps),)
            loss = (outs[0] * c1.double()).sum()
        elif mode == "pre":
            s = r + d
            outs = (s, F.layer_norm(s, (cols,), w, b, ln.eps))
            loss = (outs[0] * c1.double()).sum() + (outs[1] * c2.double()).sum()
        else:
            s = r + F.layer_norm(d, (cols,), w, b, ln.eps)
            outs = (s, s)
            loss = (s * c1.double()).sum() + (s * c2.double()).sum()
        return outs, torch.autograd.grad(loss, (r, d, w, b), allow_unused=True)

    lnc = ln.cuda()
    for mode in ("ln", "pre", "post"):
        rc, dc = resid.cuda().requires_grad_(True), delta.cuda().requires_grad_(True)
        lnc.zero_grad()
        if mode == "ln":
            outs = (fused.layer_norm(rc, lnc, act_dtype),)
            loss = (outs[0].float() * c1.cuda()).sum()
        elif mode == "pre":
            outs = fused.add_layer_norm(rc, dc, lnc, act_dtype)
            loss = (outs[0].float() * c1.cuda()).sum() + (outs[1].float() * c2.cuda()).sum()
        else:
            outs = fused.post_norm_add(rc, dc, lnc, act_dtype)
            loss = (outs[0].float() * c1.cuda()).sum() + (outs[1].float() * c2.cuda()).sum()
        loss.backward()
        wo, wg = ref(mode)
        for o, w_ in zip(outs, wo):
            check(o, w_, tol, f"layernorm {mode} out")
        assert outs[0].dtype == (act_dtype if mode == "ln" else stream_dtype)
        check(rc.grad, wg[0], tol * 3, f"layernorm {mode} d resid")
        if mode != "ln":
            check(dc.grad, wg[1], tol * 3, f"layernorm {mode} d delta")
        check(lnc.weight.grad, wg[2], tol * 3, f"layernorm {mode} d gamma")
        check(lnc.bias.grad, wg[3], tol * 3, f"layernorm {mode} d beta")


# ------------------------------------------------------------------------------------------
# Tensor-core multi-head attention (csrc/mha_tc.cu): the op itself against the fp64 oracle core, every mask kind,
# ragged lengths (T, S not multiples of the 128-row tiles), packed projections (column slices as q / k / v)
# ------------------------------------------------------------------------------------------
def _mha_core_oracle(q, k, v, nH, scale, mask):
    T, B, E = q.shape
    S = k.shape[0]
    d = E // nH
    qh = (q * scale).reshape(T, B * nH, d).transpose(0, 1)
    kh = k.reshape(S, B * nH, d).transpose(0, 1)
    vh = v.reshape(S, B * nH, d).transpose(0, 1)
    s = qh @ kh.transpose(1, 2)
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, -1)
    return (p @ vh).transpose(0, 1).reshape(T, B, E), torch.logsumexp(s, -1)


@pytest.mark.parametrize("d,nH,T,S,B,mask_kind,packed", [
    (64, 2, 128, 128, 1, "none", False), (64, 3, 200, 333, 2, "none", False), (64, 2, 300, 300, 2, "future", True),
    (64, 2, 130, 257, 1, "tensor", False), (32, 4, 128, 128, 2, "none", False), (32, 3, 368, 368, 2, "future", True),
    (32, 2, 100, 500, 1, "future", False), (32, 2, 513, 140, 2, "tensor", False), (64, 12, 512, 512, 2, "future", True)],
    ids=lambda v: str(v))
def test_mha_tensor_core_vs_oracle(mm, d, nH, T, S, B, mask_kind, packed):
    E = d * nH
    g = torch.Generator().manual_seed(T * 7 + S)
    if packed and T == S:
        qkv = torch.randn(T, B, 3 * E, generator=g).bfloat16()
        q, k, v = qkv.chunk(3, -1)
    else:
        q, k, v = (torch.randn(n, B, E, generator=g).bfloat16() for n in (T, S, S))
    cot = torch.randn(T, B, E, generator=g).bfloat16()
    scale = d ** -0.5
    diag = 1 + abs(S - T)
    mask = None
    if mask_kind == "future":
        mask = R.future_mask(T, S, torch.float64)
    elif mask_kind == "tensor":
        mask = torch.randn(T, S, generator=g).double() * 2
        mask[torch.rand(T, S, generator=g) < 0.2] = float("-inf")
        mask[:, 0] = 0.0                                            # no fully masked row
    ins = [t.double().requires_grad_(True) for t in (q, k, v)]
    want, lse_w = _mha_core_oracle(*ins, nH, scale, mask)
    gw = torch.autograd.grad((want * cot.double()).sum(), ins)
    kind = {"none": mm.lib.MASK_NONE, "future": mm.lib.MASK_FUTURE, "tensor": mm.lib.MASK_TENSOR}[mask_kind]
    if packed and T == S:
        qkvc = qkv.cuda().requires_grad_(True)
        qc, kc, vc = qkvc.chunk(3, -1)
        leaves = [qkvc]
    else:
        qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q, k, v))
        leaves = [qc, kc, vc]
    dsc = mm.lib.MhaDesc()
    dsc.tgt_len, dsc.src_len, dsc.batch, dsc.num_heads, dsc.head_dim, dsc.io_dtype = T, S, B, nH, d, mm.lib.DT_BF16
    dsc.q_stride_t, dsc.q_stride_b = qc.stride(0), qc.stride(1)
    assert mm.lib.load().mmn_mha_path(dsc).decode() == "tcgen05"
    mk = mask.float().cuda().contiguous() if mask_kind == "tensor" else None
    out, lse = torch.ops.mmn_b200.mha_fwd(qc, kc, vc, mk, nH, kind, diag if mask_kind == "future" else 0, scale, 0.0, 0, 0)
    gg = torch.autograd.grad((out.float() * cot.cuda().float()).sum(), leaves)
    check(out, want, BF16_TOL, "mha tc out")
    assert rel_err(lse.reshape(B * nH, T), lse_w) < 1e-2
    if len(leaves) == 1:
        gg = gg[0].chunk(3, -1)
    for a, b, n in zip(gg, gw, "qkv"):
        check(a, b, BF16_TOL, "mha tc d" + n)


def test_mha_tensor_core_matches_generic_at_scale(mm):
    """E=768, 12 heads x 64, T=S=1024, batch 4, causal: the tcgen05 kernels against the generic fp32-arithmetic kernels on
    the same bf16 inputs (the oracle is compared at the smaller sizes above)."""
    T = S = 1024
    B, nH, d = 4, 12, 64
    E = nH * d
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn(T, B, 3 * E, device="cuda", generator=g).bfloat16().requires_grad_(True)
    cot = torch.randn(T, B, E, device="cuda", generator=g).bfloat16()
    res = {}
    for name, dt in (("tc", torch.bfloat16), ("gen", torch.float32)):
        x = qkv.detach().to(dt).requires_grad_(True)
        q, k, v = x.chunk(3, -1)
        out, lse = torch.ops.mmn_b200.mha_fwd(q, k, v, None, nH, mm.lib.MASK_FUTURE, 1, d ** -0.5, 0.0, 0, 0)
        gx, = torch.autograd.grad((out.float() * cot.float()).sum(), x)
        res[name] = (out.float(), lse, gx.float())
    for i, n in enumerate(("out", "lse", "dqkv")):
        assert rel_err(res["tc"][i], res["gen"][i]) < BF16_TOL, n


@pytest.mark.parametrize("d,nH,T,S,B,mask_kind", [(64, 3, 256, 256, 2, "none"), (64, 2, 300, 421, 1, "future"),
                                                    (32, 4, 257, 130, 2, "tensor"), (32, 2, 128, 512, 3, "future")], ids=lambda v: str(v))
def test_mha_tensor_core_dropout_matches_generic(mm, d, nH, T, S, B, mask_kind):
    """Attention dropout on the tcgen05 path (multihead_attention.py:123; main.py's default attn_dropout is 0.1): the mask is a
    pure function of (seed, offset, item, t, s) shared with the generic kernels (dropout_rng.cuh), so the tensor-core forward and
    both backward kernels must reproduce the generic fp32-arithmetic kernels' outputs and gradients on the same bf16 inputs
    -- element for element the same probabilities dropped -- and the keep rate must be 1 - p."""
    E, p, seed, off = d * nH, 0.3, 1234, 77
    g = torch.Generator(device="cuda").manual_seed(T + S)
    q, k, v = (torch.randn(n, B, E, device="cuda", generator=g).bfloat16() for n in (T, S, S))
    cot = torch.randn(T, B, E, device="cuda", generator=g).bfloat16()
    kind = {"none": mm.lib.MASK_NONE, "future": mm.lib.MASK_FUTURE, "tensor": mm.lib.MASK_TENSOR}[mask_kind]
    diag = 1 + abs(S - T) if mask_kind == "future" else 0
    mk = None
    if mask_kind == "tensor":
        mk = torch.randn(T, S, device="cuda", generator=g) * 2
        mk[torch.rand(T, S, device="cuda", generator=g) < 0.2] = float("-inf")
        mk[:, 0] = 0.0
    dsc = mm.lib.MhaDesc()
    dsc.tgt_len, dsc.src_len, dsc.batch, dsc.num_heads, dsc.head_dim, dsc.io_dtype = T, S, B, nH, d, mm.lib.DT_BF16
    dsc.dropout_p = p
    dsc.q_stride_t, dsc.q_stride_b = q.stride(0), q.stride(1)
    assert mm.lib.load().mmn_mha_path(dsc).decode() == "tcgen05"
    res = {}
    for name, dt in (("tc", torch.bfloat16), ("gen", torch.float32)):
        ins = [t.to(dt).requires_grad_(True) for t in (q, k, v)]
        out, lse = torch.ops.mmn_b200.mha_fwd(*ins, mk, nH, kind, diag, d ** -0.5, p, seed, off)
        gs = torch.autograd.grad((out.float() * cot.float()).sum(), ins)
        res[name] = [out.float(), lse] + [x.float() for x in gs]
    for i, n in enumerate(("out", "lse", "dq", "dk", "dv")):
        assert rel_err(res["tc"][i], res["gen"][i]) < BF16_TOL, n
    # a different offset is a different mask; p = 0 differs from both
    out2, _ = torch.ops.mmn_b200.mha_fwd(q, k, v, mk, nH, kind, diag, d ** -0.5, p, seed, off + 1)
    out0, _ = torch.ops.mmn_b200.mha_fwd(q, k, v, mk, nH, kind, diag, d ** -0.5, 0.0, 0, 0)
    assert rel_err(out2.float(), res["tc"][0]) > 0.05 and rel_err(out0.float(), res["tc"][0]) > 0.05
    if mask_kind == "none":                                # keep rate through the head-averaged weights (generic kernel, same mask)
        avg = torch.ops.mmn_b200.mha_avg_weights(q.float(), k.float(), None, res["gen"][1], nH, kind, diag, d ** -0.5, p, seed, off)
        assert abs(avg.sum(-1).mean().item() - 1.0) < 0.02


def test_parallel_branches_match_serial(mm):
    """fused.parallel: the two modalities' independent groups on two streams (eager and under CUDA-graph capture) give the
    serial result -- outputs and every gradient."""
    from multimodal_neuroimage_b200 import fused
    grid, C, nH, B = (8, 8, 8), 96, 3, 2
    blk = mm.fu.CRSTB(dim=C, input_resolution=grid, depth=2, num_heads=nH, window_size=4, img_size=grid, patch_size=1).cuda()
    _randomise(blk, 3)
    params = list(blk.parameters())                       # the convolutions CRSTB declares but never calls get no gradient
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, math.prod(grid), C, device="cuda", generator=g).requires_grad_(True)
    y = torch.randn(B, math.prod(grid), C, device="cuda", generator=g).requires_grad_(True)

    def run():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ox, oy = blk(x, y, grid)
        gr = torch.autograd.grad((ox.float() ** 2).mean() + oy.float().mean(), [x, y] + params, allow_unused=True)
        return [ox.detach().clone(), oy.detach().clone()] + [t.detach().clone() for t in gr if t is not None]

    try:
        fused.PARALLEL_BRANCHES = False
        serial = run()
        fused.PARALLEL_BRANCHES = True
        par = run()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            captured = run()
        for t in captured:
            t.zero_()
        graph.replay()
        torch.cuda.synchronize()
    finally:
        fused.PARALLEL_BRANCHES = False
    for a, b, c in zip(serial, par, captured):
        assert rel_err(b, a) < 1e-5 and rel_err(c, a) < 1e-5       # LayerNorm's dgamma / dbeta sums use atomics: order varies


def test_cfg4_workload_steps(mm):
    """The cfg4 model (fMRI cross-modal transformers -> SwinFusion trunk -> SwinV2 classifier) at a small size: one graph-
    replayed training step equals the eager step, every trainable parameter gets a finite gradient, the loss moves."""
    from multimodal_neuroimage_b200 import train_step as TS
    from multimodal_neuroimage_b200 import workloads as W
    torch.manual_seed(0)
    model = W.FuncStructCross3D(img_size=32, fmri_layers=1, seq_len=48)
    W.randomise_norms(model)
    model = model.cuda()
    x_l, x_u, struct, y = W.synthetic_batch_cfg4(2, 32, "cuda", seq_len=48)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x_l, x_u, struct)
    assert out.shape == (2, 1)
    # (no eager backward on the default stream before the capture: AccumulateGrad nodes remember their first stream, and
    # syncing with the legacy default stream invalidates a capture -- PyTorch's own rule for graphed training steps)
    ts = TS.TrainStep(model, torch.nn.functional.binary_cross_entropy_with_logits, (x_l, x_u, struct), y, lr=1e-3, warmup=2)
    assert ts.g_fb is not None
    losses = [float(ts().item()) for _ in range(6)]
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0]
    assert torch.isfinite(ts.flat).all() and (ts.flat != 0).float().mean() > 0.5      # every parameter's gradient arrived in the flat buffer


@pytest.mark.parametrize("window,nH", [((4, 4, 4), 3), ((8, 8), 6), ((7, 7), 4), ((4, 4, 2), 48)], ids=["3d_w4", "2d_w8", "2d_w7", "3d_442_48h"])
def test_table_bias_gather_scatter(mm, window, nH):
    """The learned relative-position bias of the fusion blocks (swinfusion_module.py:127-130) through the one-launch
    gather / scatter: forward bit-exact against table[index] (pure data movement), backward against index_add in fp64."""
    from multimodal_neuroimage_b200 import geometry
    N = math.prod(window)
    index = geometry.relative_position_index(window).view(-1).cuda()
    T = math.prod(2 * w - 1 for w in window)
    assert int(index.max()) == T - 1
    g = torch.Generator().manual_seed(N + nH)
    table = torch.randn(T, nH, generator=g).cuda().requires_grad_(True)
    cot = torch.randn(nH, N, N, generator=g).cuda()
    got = torch.ops.mmn_b200.table_bias_fwd(table, index).view(nH, N, N)
    want = table.detach()[index].view(N, N, nH).permute(2, 0, 1)
    assert torch.equal(got, want)
    dgot, = torch.autograd.grad((got * cot).sum(), table)
    dwant = torch.zeros(T, nH, dtype=torch.float64).index_add_(0, index.cpu(), cot.double().cpu().permute(1, 2, 0).reshape(N * N, nH))
    check(dgot, dwant, FP32_TOL, "table bias backward")
    # and through the module: the attention class hands the kernel the same tensor the PyTorch expression gives
    attn = mm.fu.WindowAttention_fusion(32 * nH, window, nH).cuda()
    with torch.no_grad():
        attn.relative_position_bias_table.normal_(0, 0.5)
    b = attn.position_bias()
    assert torch.equal(b, attn.relative_position_bias_table.detach()[attn.relative_position_index.view(-1)].view(N, N, nH).permute(2, 0, 1))


def test_train_step_bf16_weight_copies_change_nothing(mm):
    """TrainStep keeps bf16 copies of the weights and refreshes them after the optimizer step (ops.Bf16Shadows) instead of
    casting every weight inside the forward: the loss trajectory must be the one of the casting step, bit for bit in the
    forward (the copies hold exactly what the casts produce), graph-replayed and eager."""
    from multimodal_neuroimage_b200 import train_step as TS
    from multimodal_neuroimage_b200 import workloads as W

    def run(shadows, use_graph):
        torch.manual_seed(0)
        model = W.SwinFusion3D(img_size=32, Ex_depths=(2,), Fusion_depths=(2,), Re_depths=(2,))
        W.randomise_norms(model)
        model = model.cuda()
        A, B, y = W.synthetic_batch(2, 32, "cuda")
        ts = TS.TrainStep(model, torch.nn.functional.binary_cross_entropy_with_logits, (A, B), y, lr=1e-3, warmup=2, use_graph=use_graph,
                          bf16_weight_copies=shadows)
        big = [p for p in ts.params if p.dim() >= 2]
        if shadows:
            assert len(ts.shadows.src) > 20 and all(mm.ops.weight_bf16(p) is d for p, d in zip(ts.shadows.src, ts.shadows.dst))
        else:
            assert ts.shadows is None and all(mm.ops.weight_bf16(p) is not mm.ops.weight_bf16(p) for p in big)    # fresh casts
        out = [float(ts().item()) for _ in range(4)]
        if shadows:                       # still the current weights after four optimizer steps
            assert all(mm.ops.weight_bf16(p) is d and torch.equal(d, p.detach().bfloat16()) for p, d in zip(ts.shadows.src, ts.shadows.dst))
            ts.shadows.close()
        return out

    for use_graph in (False, True):      # (a graphed TrainStep takes its warm-up steps before the capture: compare like with like)
        cast, copy = run(False, use_graph), run(True, use_graph)
        assert cast[-1] < cast[0]
        assert all(abs(a - b) <= 2e-3 * abs(a) for a, b in zip(cast, copy)), (use_graph, cast, copy)


def test_patch_embed_3d_projection_matches_conv(mm):
    """PatchEmbed3D as gather + tensor-core projection against the Conv3d it replaces (outputs and parameter gradients)."""
    pe = mm.v2.PatchEmbed3D(32, 4, 1, 96, torch.nn.LayerNorm).cuda()
    x = torch.randn(2, 1, 32, 32, 32, device="cuda").half()
    cot = torch.randn(2, 512, 96, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got = pe(x)
    gg = torch.autograd.grad((got * cot).sum(), [pe.proj.weight, pe.proj.bias, pe.norm.weight])
    w, b = pe.proj.weight.detach().double().requires_grad_(True), pe.proj.bias.detach().double().requires_grad_(True)
    nw = pe.norm.weight.detach().double().requires_grad_(True)
    ref = torch.nn.functional.conv3d(x.double(), w, b, stride=4).flatten(2).transpose(1, 2)
    ref = torch.nn.functional.layer_norm(ref, (96,), nw, pe.norm.bias.detach().double(), pe.norm.eps)
    gw = torch.autograd.grad((ref * cot.double()).sum(), [w, b, nw])
    check(got, ref, BF16_TOL, "patch embed")
    for a, b_, n in zip(gg, gw, ("dW", "db", "dgamma")):
        check(a, b_, BF16_TOL, "patch embed " + n)
