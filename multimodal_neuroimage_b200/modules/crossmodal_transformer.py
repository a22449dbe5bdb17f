"""Drop-in replacement for the reference's modules/crossmodal_transformer.py (MulT-style
encoder: Q from one stream, K/V from another, optional causal mask).

Same class names, constructor arguments, forward signatures and state_dict keys
(`version`, `embed_positions._float_tensor`, `layers.{i}.self_attn.*`, `layers.{i}.fc1/fc2.*`,
`layers.{i}.layer_norms.{0,1}.*`, `layer_norm.*`).  The per-layer, per-forward CPU build and
`.cuda()` copy of the future mask (crossmodal_transformer.py:179-186, 541 KB H2D x 36
layers per step in cfg1) is gone: the mask is cached on the device and the attention
kernel generates the causal pattern from indices.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import fused, geometry
from .multihead_attention import MultiheadAttention, cached_future_mask
from .position_embedding import SinusoidalPositionalEmbedding


class TransformerEncoder(nn.Module):
    def __init__(self, embed_dim, num_heads_mult, layers, attn_dropout=0.0, relu_dropout=0.0, res_dropout=0.0,
                 embed_dropout=0.0, attn_mask=False):
        super().__init__()
        self.dropout = embed_dropout
        self.attn_dropout = attn_dropout
        self.embed_dim = embed_dim
        self.embed_scale = math.sqrt(embed_dim)
        self.embed_positions = SinusoidalPositionalEmbedding(embed_dim)
        self.attn_mask = attn_mask
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(embed_dim, num_heads_mult=num_heads_mult, attn_dropout=attn_dropout,
                                    relu_dropout=relu_dropout, res_dropout=res_dropout, attn_mask=attn_mask)
            for _ in range(layers)])
        self.register_buffer('version', torch.Tensor([2]))
        self.normalize = True
        if self.normalize:
            self.layer_norm = LayerNorm(embed_dim)

    def _embed(self, t):
        x = self.embed_scale * t
        if self.embed_positions is not None:
            x = x + self.embed_positions(t.transpose(0, 1)[:, :, 0]).transpose(0, 1)
        return F.dropout(x, p=self.dropout, training=self.training)

    def forward(self, x_in, x_in_k=None, x_in_v=None):
        """x_in (T,B,E); optional key/value streams (S,B,E) -> (T,B,E)
        (crossmodal_transformer.py:49-90)."""
        x = self._embed(x_in)
        cross = x_in_k is not None and x_in_v is not None
        if cross:
            x_k, x_v = self._embed(x_in_k), self._embed(x_in_v)
        for layer in self.layers:
            x = layer(x, x_k, x_v) if cross else layer(x)
        if self.normalize:
            x = self.layer_norm(x)
        return x

    def max_positions(self):
        if self.embed_positions is None:
            return self.max_source_positions
        return min(self.max_source_positions, self.embed_positions.max_positions())


class TransformerEncoderLayer(nn.Module):
    """Pre-LN layer; the SAME layer_norms[0] is applied to x, x_k and x_v
    (crossmodal_transformer.py:133-165)."""

    def __init__(self, embed_dim, num_heads_mult=4, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.1, attn_mask=False):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads_mult = num_heads_mult
        self.self_attn = MultiheadAttention(embed_dim=self.embed_dim, num_heads_mult=self.num_heads_mult,
                                            attn_dropout=attn_dropout)
        self.self_attn.need_weights = False          # the reference discards them (:148,:152)
        self.attn_mask = attn_mask
        self.relu_dropout = relu_dropout
        self.res_dropout = res_dropout
        self.normalize_before = True
        self.fc1 = Linear(self.embed_dim, 4 * self.embed_dim)
        self.fc2 = Linear(4 * self.embed_dim, self.embed_dim)
        self.layer_norms = nn.ModuleList([LayerNorm(self.embed_dim) for _ in range(2)])

    def forward(self, x, x_k=None, x_v=None):
        residual = x
        x = self.maybe_layer_norm(0, x, before=True)
        mask = buffered_future_mask(x, x_k) if self.attn_mask else None
        if x_k is None and x_v is None:
            x, _ = self.self_attn(query=x, key=x, value=x, attn_mask=mask)
        else:
            x_k = self.maybe_layer_norm(0, x_k, before=True)
            x_v = self.maybe_layer_norm(0, x_v, before=True)
            x, _ = self.self_attn(query=x, key=x_k, value=x_v, attn_mask=mask)
        x = F.dropout(x, p=self.res_dropout, training=self.training)
        x = residual + x
        x = self.maybe_layer_norm(0, x, after=True)

        residual = x
        x = self.maybe_layer_norm(1, x, before=True)
        x = fused.mlp(x, self.fc1, self.fc2, "relu", self.relu_dropout, self.training, 0.0)     # fc1 -> relu -> dropout -> fc2 (:158-160)
        x = F.dropout(x, p=self.res_dropout, training=self.training)
        x = residual + x
        return self.maybe_layer_norm(1, x, after=True)

    def maybe_layer_norm(self, i, x, before=False, after=False):
        assert before ^ after
        return fused.layer_norm(x, self.layer_norms[i], x.dtype if not torch.is_autocast_enabled("cuda") else None) \
            if after ^ self.normalize_before else x


def fill_with_neg_inf(t):
    return t.float().fill_(float('-inf')).type_as(t)


def buffered_future_mask(tensor, tensor2=None):
    """(T,S) mask, -inf strictly above diagonal 1+|S-T| (crossmodal_transformer.py:179-186).
    Built once per (T,S,device) on the device and cached; MultiheadAttention recognises the
    cached object and lets the kernel regenerate the pattern from indices."""
    T = tensor.size(0)
    S = tensor2.size(0) if tensor2 is not None else T
    return cached_future_mask(T, S, tensor.device)


def Linear(in_features, out_features, bias=True):
    m = nn.Linear(in_features, out_features, bias)
    nn.init.xavier_uniform_(m.weight)
    if bias:
        nn.init.constant_(m.bias, 0.)
    return m


def LayerNorm(embedding_dim):
    return nn.LayerNorm(embedding_dim)
