"""The window-attention MODULE as one autograd node: projection(s) -> fused shifted-window
attention kernel -> output projection, with a hand-written backward.

Why not leave it to autograd: F.linear's backward reduces its (tokens x channels) grad_output
over tokens to get the bias gradient with a generic reduction kernel -- 20 % of the step at
BASELINE cfg2.  Here the q/k/v bias gradients come out of the attention backward kernel for
free (column sums of dq, dk, dv, `dcolsum`), the output-projection bias gradient from a
memory-speed column-sum kernel, and nothing but the GEMMs (cuBLAS: plain library GEMMs) is left
to PyTorch.  Mirrors, for the self and the cross variant:
  reference swin_v2_module.py:147-176, swinfusion_module.py:121-143, 221-244.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib, ops  # noqa: F401  (ops registers torch.ops.mmn_b200.*)


def _mm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b with fp32 output (weight gradients are accumulated over ~1e6 tokens)."""
    if a.dtype == torch.float32:
        return a @ b
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except TypeError:
        return (a @ b).float()


class WindowAttentionModuleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, w_a, b_a, w_b, b_b, w_proj, b_proj, bias, head_scale, grid, window, shift, num_heads,
                score_kind, mask_kind, scale, dropout, path):
        """x (B, L, C) [queries; also keys/values when y is None]; y (B, L, C) or None.
        self:  w_a (3C, C), b_a (3C) | None;  w_b, b_b = None.
        cross: w_a (C, C) for q from x;  w_b (2C, C), b_b for kv from y."""
        odt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        cdt = torch.bfloat16 if odt == torch.float16 else odt     # fp16 autocast (trainer.py:378) runs as bf16 (ops.kernel_io)
        B, L, C = x.shape
        with torch.autocast("cuda", enabled=False):
            cast = lambda t: None if t is None else t.to(cdt)
            xc, yc = cast(x).reshape(B * L, C), (cast(y).reshape(B * L, C) if y is not None else None)
            wa, ba, wb, bb, wp, bp = cast(w_a), cast(b_a), cast(w_b), cast(b_b), cast(w_proj), cast(b_proj)
            a = F.linear(xc, wa, ba).view(B, *grid, -1)
            b = F.linear(yc, wb, bb).view(B, *grid, -1) if y is not None else None
            p, seed, off = dropout
            out, lse = torch.ops.mmn_b200.winattn_fwd(a, b, bias, head_scale, None, list(grid), list(window), list(shift),
                                                      num_heads, score_kind, mask_kind, scale, p, seed, off, path)
            res = F.linear(out.view(B * L, C), wp, bp).view(B, L, C).to(odt)
        ctx.save_for_backward(xc, yc, a, b, out, lse, wa, wb, wp, bias, head_scale)
        ctx.cfg = (list(grid), list(window), list(shift), num_heads, score_kind, mask_kind, scale, p, seed, off, path)
        ctx.meta = (x.dtype, None if y is None else y.dtype, w_a.dtype, None if b_a is None else b_a.dtype,
                    None if w_b is None else w_b.dtype, None if b_b is None else b_b.dtype, w_proj.dtype,
                    None if b_proj is None else b_proj.dtype, (B, L, C))
        return res

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dres):
        xc, yc, a, b, out, lse, wa, wb, wp, bias, head_scale = ctx.saved_tensors
        xdt, ydt, wadt, badt, wbdt, bbdt, wpdt, bpdt, (B, L, C) = ctx.meta
        with torch.autocast("cuda", enabled=False):
            dy = dres.to(xc.dtype).reshape(B * L, C)
            if not dy.is_contiguous():
                dy = dy.contiguous()
            out2 = out.view(B * L, C)
            if ops.linear_bwd_supported(dy, out2, wp):
                # one pass over dy: dout = dy Wp, dWp = dy^T out, dbp = colsum(dy)  (libmmn_b200: linbwd_tc.cu)
                dout2, d_wp, d_bp = torch.ops.mmn_b200.linear_bwd(dy, out2, wp)
                dout = dout2.view(out.shape)
                d_wp = d_wp.to(wpdt)
                d_bp = d_bp.to(bpdt) if bpdt is not None else None
            else:
                d_bp = None
                if bpdt is not None:      # the column-sum kernel moves 8 columns per thread; odd widths are the reference's tiny stages
                    d_bp = (torch.ops.mmn_b200.colsum(dy) if C % 8 == 0 and C <= 2048 else dy.float().sum(0)).to(bpdt)
                d_wp = _mm_f32(dy.t(), out2).to(wpdt)
                dout = (dy @ wp).view(out.shape)
            da, db, dbias, dhs, dcs = torch.ops.mmn_b200.winattn_bwd(dout, a, b, bias, head_scale, None, out, lse, *ctx.cfg,
                                                                     True)
            da2 = da.view(B * L, -1)
            if ops.linear_bwd_supported(da2, xc, wa):
                dx, d_wa, _ = torch.ops.mmn_b200.linear_bwd(da2, xc, wa)
                dx, d_wa = dx.view(B, L, C).to(xdt), d_wa.to(wadt)
            else:
                d_wa = _mm_f32(da2.t(), xc).to(wadt)
                dx = (da2 @ wa).view(B, L, C).to(xdt)
            if b is None:
                d_ba = dcs.reshape(-1).to(badt) if badt is not None else None
                d_y = d_wb = d_bb = None
            else:
                db2 = db.view(B * L, -1)
                d_ba = dcs[0].to(badt) if badt is not None else None
                d_bb = dcs[1:].reshape(-1).to(bbdt) if bbdt is not None else None
                if ops.linear_bwd_supported(db2, yc, wb):
                    d_y, d_wb, _ = torch.ops.mmn_b200.linear_bwd(db2, yc, wb)
                    d_y, d_wb = d_y.view(B, L, C).to(ydt), d_wb.to(wbdt)
                else:
                    d_wb = _mm_f32(db2.t(), yc).to(wbdt)
                    d_y = (db2 @ wb).view(B, L, C).to(ydt)
        return (dx, d_y, d_wa, d_ba, d_wb, d_bb, d_wp, d_bp, dbias if bias is not None else None,
                dhs if head_scale is not None else None) + (None,) * 9


def window_attention_module(x: torch.Tensor, y: Optional[torch.Tensor], w_a, b_a, w_b, b_b, w_proj, b_proj, bias, head_scale,
                            grid, window, shift, num_heads: int, score_kind: int, mask_kind: int, scale: float,
                            dropout=(0.0, 0, 0), path: int = 0) -> torch.Tensor:
    return WindowAttentionModuleFn.apply(x, y, w_a, b_a, w_b, b_b, w_proj, b_proj, bias, head_scale, tuple(grid), tuple(window),
                                         tuple(shift), num_heads, score_kind, mask_kind, float(scale), dropout, path)
