"""Pins oracle/ref_nd.py against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Integer maps / masks: bit-exact.  Float outputs and
gradients: fp32 oracle vs fp32 reference, rtol 1e-5 (same torch ops, same order)."""
import math

import numpy as np
import pytest
import torch

from oracle import ref_nd as R

RT, AT = 1e-5, 1e-6


def close(a, b, rt=RT, at=AT):
    torch.testing.assert_close(a, b, rtol=rt, atol=at)


def parse(key):            # "12x12_w6_s3" -> (12,12),6,3
    g, w, s = key.split("_")
    return tuple(int(v) for v in g.split("x")), int(w[1:]), int(s[1:])


def test_gather_maps_bit_exact(golden):
    G = golden("index_maps")
    for k in G.keys("gather/"):
        (H, W), ws, s = parse(k.split("/")[1])
        ids = torch.arange(H * W, dtype=torch.int64).view(1, H, W, 1)
        win = R.window_partition_nd(R.cyclic_shift_nd(ids, (s, s)), (ws, ws)).view(-1, ws * ws)
        assert np.array_equal(win.numpy(), G.arr(k)), k
        back = R.cyclic_shift_nd(R.window_reverse_nd(win.view(-1, ws, ws, 1), (ws, ws), (H, W)), (s, s), inverse=True)
        assert torch.equal(back, ids)


def test_shift_masks_bit_exact(golden):
    G = golden("index_maps")
    for fam in ("mask_v2/", "mask_fusion/", "mask_cross/"):
        for k in G.keys(fam):
            grid, ws, s = parse(k.split("/")[1])
            w_eff, s_eff = R.effective_window(grid, ws, s)
            if fam == "mask_v2/":
                assert [w_eff, s_eff] == G.arr("mask_v2_eff/" + k.split("/")[1]).tolist()
            ref = G.arr(k)
            got = R.shift_mask_nd(grid, (w_eff,) * 2, (s_eff,) * 2) if s_eff > 0 else None
            if got is None:
                assert ref.size == 0, k
            else:
                assert np.array_equal(got.numpy(), ref), k
    got = R.shift_mask_nd((18, 24), (6, 6), (3, 3))
    assert np.array_equal(got.numpy(), G.arr("mask_fusion_xsize/18x24_w6_s3"))


def test_relative_position_tables(golden):
    G = golden("index_maps")
    for k in G.keys("rpi/"):
        ws = tuple(int(v) for v in k.split("/")[1].split("x"))
        assert np.array_equal(R.relative_position_index_nd(ws).numpy(), G.arr(k)), k
        np.testing.assert_array_equal(R.cpb_coords_table_nd(ws).numpy(), G.arr("coords/" + k.split("/")[1]))
    np.testing.assert_array_equal(R.cpb_coords_table_nd((6, 6), (4, 4)).numpy(), G.arr("coords_pretrained4/6x6"))


def test_future_mask_bit_exact(golden):
    G = golden("index_maps")
    for k in G.keys("future/"):
        T, S = (int(v) for v in k.split("/")[1].split("x"))
        m = R.future_mask(T, S)
        assert np.array_equal(torch.isinf(m).numpy().astype(np.uint8), G.arr(k)), k
        assert ((m == 0) | (m == float("-inf"))).all()


def test_positions(golden):
    G = golden("index_maps")
    tok = G.t("pos/tokens")
    assert np.array_equal(R.token_positions(tok).numpy(), G.arr("pos/positions"))
    for dim in (28, 7):
        close(R.sinusoidal_positions(tok, dim), G.t(f"pos/emb{dim}"))


def _grads(outs, cots, wrt):
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    return torch.autograd.grad(loss, wrt, allow_unused=True)


def _check_case(G, case, fn, in_names):
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in G.group(f"{case}/sd/").items()}
    ins = {k: v.clone().requires_grad_(k in in_names) for k, v in G.group(f"{case}/in/").items()}
    outs = fn(sd, ins, G.arr(f"{case}/cfg").tolist())
    outs = outs if isinstance(outs, tuple) else (outs,)
    cots = [G.t(f"{case}/cot/{i}") for i in range(len(outs))]
    for i, o in enumerate(outs):
        close(o, G.t(f"{case}/out/{i}"))
    gsd = G.group(f"{case}/gsd/")
    wrt = [ins[k] for k in in_names] + [sd[k] for k in gsd]
    used = G.arr(f"{case}/loss_outs").tolist()
    got = _grads([outs[i] for i in used], [cots[i] for i in used], wrt)
    for k, g in zip(in_names, got[:len(in_names)]):
        close(g, G.t(f"{case}/gin/{k}"), 1e-4, 1e-5)
    for k, g in zip(gsd, got[len(in_names):]):
        ref = gsd[k]
        if g is None:
            assert float(ref.abs().max()) == 0.0, k
        else:
            close(g, ref, 1e-4, 2e-5)


def test_swinv2_window_attention(golden):
    G = golden("swinv2")
    for case in ("wa_c12", "wa_c96", "wa_c48_nomask"):
        def fn(sd, ins, cfg):
            C, nH, w0, w1 = cfg
            return R.window_attention_cosine(ins["x"], sd, (w0, w1), nH, ins.get("mask"))
        _check_case(G, case, fn, ["x"])


def test_swinv2_block(golden):
    G = golden("swinv2")
    for case in ("blk_shift", "blk_noshift", "blk_clamp", "blk_rect"):
        def fn(sd, ins, cfg):
            C, nH, H, W, ws, s = cfg
            return R.swin_v2_block(ins["x"], sd, (H, W), ws, s, nH)
        _check_case(G, case, fn, ["x"])


def test_fusion_window_attention(golden):
    G = golden("swinfusion")
    for c in ("wa_c12", "wa_c64"):
        def fs(sd, ins, cfg):
            return R.window_attention_scaled(ins["x"], sd, (cfg[2], cfg[3]), cfg[1], ins["mask"])
        _check_case(G, "self_" + c, fs, ["x"])

        def fc(sd, ins, cfg):
            return R.window_attention_scaled(ins["x"], sd, (cfg[2], cfg[3]), cfg[1], ins["mask"], y=ins["y"])
        _check_case(G, "cross_" + c, fc, ["x", "y"])


def test_fusion_blocks(golden):
    G = golden("swinfusion")
    for c in ("blk_shift", "blk_xsize", "blk_noshift"):
        def fs(sd, ins, cfg):
            C, nH, r0, r1, x0, x1, ws, s = cfg
            return R.fusion_block(ins["x"], sd, (x0, x1), (r0, r1), ws, s, nH)
        _check_case(G, "self_" + c, fs, ["x"])

        def fc(sd, ins, cfg):
            C, nH, r0, r1, x0, x1, ws, s = cfg
            return R.cross_block(ins["x"], ins["y"], sd, (x0, x1), (r0, r1), ws, s, nH)
        _check_case(G, "cross_" + c, fc, ["x", "y"])


def test_multihead_attention(golden):
    G = golden("crossmodal")
    for case in ("mha_self_d7", "mha_cross_d7", "mha_cross_TneS", "mha_self_d14_nomask"):
        names = ["q"] if "self" in case else ["q", "k", "v"]

        def fn(sd, ins, cfg):
            E, nH, T, S, B, has_mask = cfg
            mask = R.future_mask(T, S) if has_mask else None
            q = ins["q"]
            k, v = (q, q) if "k" not in ins else (ins["k"], ins["v"])
            return R.multihead_attention(q, k, v, sd, nH, mask)
        _check_case(G, case, fn, names)


def test_transformer_encoder(golden):
    G = golden("crossmodal")
    for case in ("enc_self", "enc_cross", "enc_cross_nomask"):
        names = ["x"] if case == "enc_self" else ["x", "xk", "xv"]

        def fn(sd, ins, cfg):
            E, nH, L, T, B, use_mask, cross = cfg
            return R.transformer_encoder(ins["x"], sd, nH, L, bool(use_mask), ins.get("xk"), ins.get("xv"))
        _check_case(G, case, fn, names)


def test_core_matches_module_path():
    """window_attention_core (what the kernels compute) + projections == a6, n in {2,3}."""
    torch.manual_seed(0)
    for grid, w, s in [((8, 8), (4, 4), (2, 2)), ((4, 8, 4), (2, 4, 2), (1, 2, 1))]:
        n, C, nH, B = len(grid), 16, 2, 2
        N = math.prod(w)
        x = torch.randn(B, math.prod(grid), C, dtype=torch.float64)
        Wqkv = torch.randn(3 * C, C, dtype=torch.float64) * 0.3
        hs = torch.rand(nH, dtype=torch.float64) + 0.5
        bias = torch.randn(nH, N, N, dtype=torch.float64)
        mask = R.shift_mask_nd(grid, w, s, torch.float64)
        qkv = (x @ Wqkv.t()).view(B, *grid, 3 * C)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        out, lse = R.window_attention_core(q, k, v, grid, w, s, nH, cosine=True, head_scale=hs, bias=bias, mask=mask)
        # the long way round
        xw = R.window_partition_nd(R.cyclic_shift_nd(x.view(B, *grid, C), s), w).view(-1, N, C)
        qkvw = (xw @ Wqkv.t()).view(-1, N, 3, nH, C // nH).permute(2, 0, 3, 1, 4)
        a = torch.nn.functional.normalize(qkvw[0], dim=-1) @ torch.nn.functional.normalize(qkvw[1], dim=-1).transpose(-1, -2)
        a = a * hs.view(1, nH, 1, 1) + bias
        a = R._softmax_with_mask(a, mask, nH)
        o = (a @ qkvw[2]).transpose(1, 2).reshape(-1, *w, C)
        o = R.cyclic_shift_nd(R.window_reverse_nd(o, w, grid), s, inverse=True)
        close(out, o, 1e-10, 1e-12)
        assert lse.shape == (xw.shape[0], nH, N)
