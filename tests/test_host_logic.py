"""CPU tests of the host side: integer maps (bit-exact vs oracle and golden), state_dict
compatibility with the reference, the C-ABI library's exports, loud failure without CUDA,
and batch sharding."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from multimodal_neuroimage_b200 import _lib, geometry
from multimodal_neuroimage_b200.modules import crossmodal_transformer as cm
from multimodal_neuroimage_b200.modules import multihead_attention as mh
from multimodal_neuroimage_b200.modules import position_embedding as pe
from multimodal_neuroimage_b200.modules import swin_v2_module as v2
from multimodal_neuroimage_b200.modules import swinfusion_module as fu
from oracle import ref_nd as R

GEOMS = [((12, 12), (6, 6), (3, 3)), ((8, 16), (4, 4), (2, 2)), ((6, 6), (3, 3), (1, 1)), ((16, 8), (8, 8), (4, 4)),
         ((12, 12), (6, 6), (0, 0)), ((8, 8, 8), (4, 4, 4), (2, 2, 2)), ((4, 8, 12), (2, 4, 4), (1, 2, 2)),
         ((8, 8, 8), (4, 4, 4), (0, 0, 0)), ((24,), (6,), (3,)), ((8, 8, 4), (4, 4, 4), (2, 2, 0))]


@pytest.mark.parametrize("grid,window,shift", GEOMS)
def test_gather_map_and_mask_match_oracle(grid, window, shift):
    L = math.prod(grid)
    ids = torch.arange(L, dtype=torch.int64).view(1, *grid, 1)
    want = R.window_partition_nd(R.cyclic_shift_nd(ids, shift), window).view(-1, math.prod(window))
    got = geometry.window_gather_map(grid, window, shift)
    assert torch.equal(got, want)
    m_want = R.shift_mask_nd(grid, window, shift)
    m_got = geometry.shift_attention_mask(grid, window, shift)
    if m_want is None:
        assert m_got is None
    else:
        assert torch.equal(m_got, m_want)
    if all(s > 0 for s in shift):
        rid_want = R.window_partition_nd(R.shift_region_ids_nd(grid, window, shift).view(1, *grid, 1), window)
        assert torch.equal(geometry.shift_region_ids(grid, window, shift), rid_want.view(-1, math.prod(window)))


@pytest.mark.parametrize("window", [(3, 3), (6, 6), (4, 8), (4, 4, 4), (2, 3, 4), (7,)])
def test_relative_tables_match_oracle(window):
    assert torch.equal(geometry.relative_position_index(window), R.relative_position_index_nd(window))
    assert torch.equal(geometry.cpb_coords_table(window), R.cpb_coords_table_nd(window))
    pre = tuple(max(2, w - 1) for w in window)
    assert torch.equal(geometry.cpb_coords_table(window, pre), R.cpb_coords_table_nd(window, pre))


def test_maps_match_reference_golden(golden):
    G = golden("index_maps")
    for k in G.keys("gather/"):
        g, w, s = k.split("/")[1].split("_")
        grid = tuple(int(v) for v in g.split("x"))
        ws, sh = int(w[1:]), int(s[1:])
        assert np.array_equal(geometry.window_gather_map(grid, (ws, ws), (sh, sh)).numpy(), G.arr(k)), k
    for k in G.keys("mask_v2/"):
        g, w, s = k.split("/")[1].split("_")
        grid = tuple(int(v) for v in g.split("x"))
        ws, sh = geometry.clamp_window(grid, int(w[1:]), int(s[1:]))
        assert [ws, sh] == G.arr("mask_v2_eff/" + k.split("/")[1]).tolist()
        got = geometry.shift_attention_mask(grid, (ws, ws), (sh, sh))
        assert (got is None and G.arr(k).size == 0) or np.array_equal(got.numpy(), G.arr(k)), k
    for k in G.keys("rpi/"):
        ws = tuple(int(v) for v in k.split("/")[1].split("x"))
        assert np.array_equal(geometry.relative_position_index(ws).numpy(), G.arr(k))
        assert np.array_equal(geometry.cpb_coords_table(ws).numpy(), G.arr("coords/" + k.split("/")[1]))
    assert np.array_equal(geometry.cpb_coords_table((6, 6), (4, 4)).numpy(), G.arr("coords_pretrained4/6x6"))
    for k in G.keys("future/"):
        T, S = (int(v) for v in k.split("/")[1].split("x"))
        assert np.array_equal(torch.isinf(geometry.future_mask(T, S)).numpy().astype(np.uint8), G.arr(k)), k
    tok = G.t("pos/tokens")
    assert np.array_equal(pe.make_positions(tok, 0, 0).numpy(), G.arr("pos/positions"))
    for dim in (28, 7):
        torch.testing.assert_close(pe.SinusoidalPositionalEmbedding(dim)(tok), G.t(f"pos/emb{dim}"), rtol=1e-6, atol=1e-6)


def test_window_partition_reverse_api():
    x = torch.randn(2, 8, 12, 5)
    w = v2.window_partition(x, 4)
    assert torch.equal(w, R.window_partition_nd(x, (4, 4)))
    assert torch.equal(v2.window_reverse(w, 4, 8, 12), x)
    x3 = torch.randn(1, 4, 8, 4, 3)
    w3 = fu.window_partition_fusion(x3, (2, 4, 2))
    assert torch.equal(w3, R.window_partition_nd(x3, (2, 4, 2)))
    assert torch.equal(fu.window_reverse_fusion(w3, (2, 4, 2), 4, 8, 4), x3)


def _sd_matches(module, G, case):
    ref = G.group(f"{case}/sd/")
    ours = module.state_dict()
    assert set(ours.keys()) == set(ref.keys()), (sorted(set(ours) ^ set(ref)))
    for k in ref:
        assert tuple(ours[k].shape) == tuple(ref[k].shape), k
        assert ours[k].dtype == ref[k].dtype, k
    module.load_state_dict(ref, strict=True)


def test_state_dicts_match_reference(golden):
    G = golden("swinv2")
    C, nH, w0, w1 = G.arr("wa_c12/cfg").tolist()
    _sd_matches(v2.WindowAttention(C, (w0, w1), nH), G, "wa_c12")
    for case in ("blk_shift", "blk_noshift", "blk_clamp", "blk_rect"):
        C, nH, H, W, ws, s = G.arr(f"{case}/cfg").tolist()
        _sd_matches(v2.SwinTransformerBlock(C, (H, W), nH, window_size=ws, shift_size=s), G, case)
    G = golden("swinfusion")
    C, nH, w0, w1 = G.arr("self_wa_c12/cfg").tolist()
    _sd_matches(fu.WindowAttention_fusion(C, (w0, w1), nH), G, "self_wa_c12")
    _sd_matches(fu.Cross_WindowAttention(C, (w0, w1), nH), G, "cross_wa_c12")
    for case in ("blk_shift", "blk_xsize", "blk_noshift"):
        C, nH, r0, r1, x0, x1, ws, s = G.arr(f"self_{case}/cfg").tolist()
        _sd_matches(fu.SwinTransformerBlock_fusion(C, (r0, r1), nH, window_size=ws, shift_size=s), G, "self_" + case)
        _sd_matches(fu.Cross_SwinTransformerBlock(C, (r0, r1), nH, window_size=ws, shift_size=s), G, "cross_" + case)
    G = golden("crossmodal")
    E, nH = G.arr("mha_self_d7/cfg").tolist()[:2]
    _sd_matches(mh.MultiheadAttention(E, nH), G, "mha_self_d7")
    E, nH, L, T, B, use_mask, cross = G.arr("enc_cross/cfg").tolist()
    _sd_matches(cm.TransformerEncoder(E, nH, L, attn_mask=bool(use_mask)), G, "enc_cross")


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(_lib.ROOT, "include", "mmn_b200.h")).read()
    declared = set(re.findall(r"\b(mmn_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mmn_abi_version() == _lib.ABI_VERSION


def test_ctypes_structs_match_header(tmp_path):
    """Compile a C probe against include/mmn_b200.h and compare sizeof/offsetof with ctypes."""
    import subprocess
    probe = tmp_path / "probe.c"
    probe.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "mmn_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(mmn_winattn_desc), offsetof(mmn_winattn_desc, scale),
         offsetof(mmn_winattn_desc, seed), offsetof(mmn_winattn_desc, q_row_stride), offsetof(mmn_winattn_desc, dv_row_stride));
  printf("%zu %zu %zu %zu %zu\n", sizeof(mmn_mha_desc), offsetof(mmn_mha_desc, scale), offsetof(mmn_mha_desc, seed),
         offsetof(mmn_mha_desc, q_stride_t), offsetof(mmn_mha_desc, dv_stride_b));
  return 0;
}
""")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(_lib.ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    W, M = _lib.WinAttnDesc, _lib.MhaDesc
    assert [int(v) for v in out[:5]] == [ctypes.sizeof(W), W.scale.offset, W.seed.offset, W.q_row_stride.offset,
                                         W.dv_row_stride.offset]
    assert [int(v) for v in out[5:]] == [ctypes.sizeof(M), M.scale.offset, M.seed.offset, M.q_stride_t.offset,
                                         M.dv_stride_b.offset]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    blk = v2.SwinTransformerBlock(12, (6, 6), 3, window_size=3, shift_size=1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        blk(torch.randn(1, 36, 12))
    m = mh.MultiheadAttention(8, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(4, 1, 8), torch.randn(4, 1, 8), torch.randn(4, 1, 8))
    # and the C entry point itself refuses without a device
    d = _lib.WinAttnDesc()
    d.ndim, d.batch, d.num_heads, d.head_dim = 1, 1, 1, 4
    d.grid[0] = d.window[0] = 4
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = _lib.load().mmn_winattn_fwd(ctypes.byref(d), p, p, p, None, None, None, p, p, p, 0, None)
    assert rc == -3 and b"no CPU path" in _lib.load().mmn_last_error()


def test_descriptor_validation():
    lib = _lib.load()
    d = _lib.WinAttnDesc()
    d.ndim, d.batch, d.num_heads, d.head_dim = 2, 1, 1, 4
    d.grid[0], d.grid[1], d.window[0], d.window[1] = 6, 6, 4, 3
    buf = (ctypes.c_float * 8)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.mmn_winattn_fwd(ctypes.byref(d), p, p, p, None, None, None, p, p, p, 0, None) == -1
    assert b"does not divide" in lib.mmn_last_error()
    d.window[0] = 3
    d.shift[0] = 3
    assert lib.mmn_winattn_fwd(ctypes.byref(d), p, p, p, None, None, None, p, p, p, 0, None) == -1
    assert lib.mmn_winattn_path(ctypes.byref(d)) == b"invalid"


def test_shard_range():
    for total in (0, 1, 7, 8, 64, 1000):
        for ws in (1, 2, 3, 8):
            spans = [geometry.shard_range(total, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_mha_head_padding_is_exact():
    """The zero-padded projections MultiheadAttention uses to put head dims 7 / 14 on the tensor-core kernels (16-bit
    autocast on CUDA): padded q.k logits and the padded out-projection equal the unpadded ones, bit for bit in fp32."""
    import torch
    import torch.nn.functional as F
    from multimodal_neuroimage_b200.modules.multihead_attention import MultiheadAttention
    torch.manual_seed(3)
    for E, nH in ((84, 12), (168, 12)):
        m = MultiheadAttention(E, nH, add_bias_kv=True)
        d, pad = E // nH, 32
        x = torch.randn(5, 2, E)
        assert m._padded_head_dim(x) == 0                       # CPU / fp32 tensors never take the padded route
        for kw in ({}, {"end": E}, {"start": E}, {"start": E, "end": 2 * E}, {"start": 2 * E}):
            plain, padded = m._in_proj(x, **kw), m._in_proj(x, pad=pad, **kw)
            assert padded.shape[-1] == plain.shape[-1] // d * pad
            back = padded.view(5, 2, -1, pad)
            assert torch.equal(back[..., :d].reshape(plain.shape), plain) and not back[..., d:].any()
        assert torch.equal(m._pad_heads(m.bias_k, pad).view(1, 1, nH, pad)[..., :d].reshape(1, 1, E), m.bias_k)
        a = torch.randn(5, 2, nH, pad)
        a[..., d:] = 7.0                                          # whatever sits in the pad channels must not reach the output
        w_out = F.pad(m.out_proj.weight.view(E, nH, d), (0, pad - d)).reshape(E, nH * pad)
        got = F.linear(a.reshape(5, 2, nH * pad), w_out, m.out_proj.bias)
        want = m.out_proj(a[..., :d].reshape(5, 2, E))
        assert torch.allclose(got, want, atol=1e-5)


def test_bf16_weight_shadows_follow_the_version_counter():
    """ops.weight_bf16 serves a registered bf16 copy only while the weight is the one last refreshed: any in-place update
    from outside (optimizer step without refresh, load_state_dict) falls back to a fresh cast."""
    import torch
    from multimodal_neuroimage_b200 import ops
    lin = torch.nn.Linear(8, 4)
    assert ops.weight_bf16(lin.weight).data_ptr() != ops.weight_bf16(lin.weight).data_ptr()      # no copy registered: casts
    sh = ops.Bf16Shadows(lin.parameters())
    assert len(sh.src) == 1                                          # the bias (1-D) keeps its own dtype
    a = ops.weight_bf16(lin.weight)
    assert a is sh.dst[0] and torch.equal(a, lin.weight.detach().bfloat16())
    with torch.no_grad():
        lin.weight.mul_(2.0)                                         # update without refresh: the copy is stale and not served
    b = ops.weight_bf16(lin.weight)
    assert b is not sh.dst[0] and torch.equal(b, lin.weight.detach().bfloat16())
    sh.refresh()
    assert ops.weight_bf16(lin.weight) is sh.dst[0] and torch.equal(sh.dst[0], lin.weight.detach().bfloat16())
    lin.load_state_dict({"weight": torch.ones(4, 8), "bias": torch.zeros(4)})
    assert torch.equal(ops.weight_bf16(lin.weight), torch.ones(4, 8, dtype=torch.bfloat16))
    sh.close()
    assert ops.weight_bf16(lin.weight) is not ops.weight_bf16(lin.weight)
    w16 = torch.zeros(2, 2, dtype=torch.bfloat16)
    assert ops.weight_bf16(w16) is w16


def test_in_proj_row_split_gradients_match_slicing():
    """MultiheadAttention hands the q / kv (or q / k / v) row blocks of in_proj_weight and in_proj_bias to the projections
    through _SplitRows, whose backward is one torch.cat instead of autograd's zero-fill + memcpy + add per slice: values are
    the slices themselves and the parameter gradients equal those of plain slicing, including an unused block."""
    import torch
    from multimodal_neuroimage_b200.modules.multihead_attention import MultiheadAttention
    torch.manual_seed(5)
    E = 24
    m = MultiheadAttention(E, 4)
    with torch.no_grad():
        m.in_proj_bias.normal_()
    x, y, z = torch.randn(5, 2, E), torch.randn(7, 2, E), torch.randn(7, 2, E)

    def grads(fn):
        m.zero_grad()
        fn().backward()
        return m.in_proj_weight.grad.clone(), m.in_proj_bias.grad.clone()

    def split2():
        (wq, wkv), (bq, bkv) = m._split_in_proj(E, 2 * E)
        assert torch.equal(wq, m.in_proj_weight[:E]) and torch.equal(wkv, m.in_proj_weight[E:]) and torch.equal(bkv, m.in_proj_bias[E:])
        return m._project(x, wq, bq).sin().sum() + m._project(y, wkv, bkv).pow(2).sum()

    def slice2():
        return m._in_proj(x, end=E).sin().sum() + m._in_proj(y, start=E).pow(2).sum()

    def split3():
        (wq, wk, wv), (bq, bk, bv) = m._split_in_proj(E, E, E)
        return m._project(x, wq, bq).sin().sum() + m._project(y, wk, bk).pow(2).sum() + m._project(z, wv, bv).cos().sum()

    def slice3():
        return m._in_proj(x, end=E).sin().sum() + m._in_proj(y, start=E, end=2 * E).pow(2).sum() + m._in_proj(z, start=2 * E).cos().sum()

    def split_unused():
        (wq, _), (bq, _) = m._split_in_proj(E, 2 * E)
        return m._project(x, wq, bq).sin().sum()

    def slice_unused():
        return m._in_proj(x, end=E).sin().sum()

    for a, b in ((split2, slice2), (split3, slice3), (split_unused, slice_unused)):
        (gw, gb), (hw, hb) = grads(a), grads(b)
        assert torch.allclose(gw, hw, atol=1e-6) and torch.allclose(gb, hb, atol=1e-6)
    m2 = MultiheadAttention(E, 4, bias=False)
    (wq, wkv), (bq, bkv) = m2._split_in_proj(E, 2 * E)
    assert bq is None and bkv is None and wkv.shape == (2 * E, E)
