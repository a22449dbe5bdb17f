"""Sinusoidal positional embedding of the cross-modal transformer
(reference: modules/position_embedding.py).  Elementwise input prep, not a contraction:
stays in PyTorch, but without the reference's per-device Python caches --
the table is a function of (length, dim) and is rebuilt on the input's device when needed.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def make_positions(tensor, padding_idx, left_pad=False):
    """Position t+padding_idx+1 where the 'token' differs from padding_idx, else the token
    itself (position_embedding.py:8-27)."""
    T = tensor.size(1)
    pos = torch.arange(padding_idx + 1, padding_idx + 1 + T, device=tensor.device, dtype=tensor.dtype).expand_as(tensor)
    keep = tensor.ne(padding_idx)
    if left_pad:
        pos = pos - T + keep.long().sum(dim=1, keepdim=True)
    return torch.where(keep, pos, tensor).long()


class SinusoidalPositionalEmbedding(nn.Module):
    def __init__(self, embedding_dim, padding_idx=0, left_pad=0, init_size=128):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.padding_idx = padding_idx
        self.left_pad = left_pad
        self.weights = dict()
        self.register_buffer('_float_tensor', torch.FloatTensor(1))

    @staticmethod
    def get_embedding(num_embeddings, embedding_dim, padding_idx=None, device=None):
        """cat(sin, cos) of position * exp(-ln(1e4)/(half-1) * i); odd dims zero-padded;
        the padding row zeroed (position_embedding.py:41-60)."""
        half = embedding_dim // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float, device=device) * -(math.log(10000) / (half - 1)))
        ang = torch.arange(num_embeddings, dtype=torch.float, device=device).unsqueeze(1) * freq.unsqueeze(0)
        emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).view(num_embeddings, -1)
        if embedding_dim % 2 == 1:
            emb = torch.cat([emb, emb.new_zeros(num_embeddings, 1)], dim=1)
        if padding_idx is not None:
            emb[padding_idx, :] = 0
        return emb

    def forward(self, input):
        """input: (bsz, seqlen) float 'tokens' -> (bsz, seqlen, dim), detached."""
        bsz, seq_len = input.size()
        max_pos = self.padding_idx + 1 + seq_len
        key = str(input.device)
        tab = self.weights.get(key)
        if tab is None or tab.size(0) < max_pos:
            tab = self.get_embedding(max_pos, self.embedding_dim, self.padding_idx, device=input.device)
            self.weights[key] = tab
        tab = tab.type_as(self._float_tensor)
        positions = make_positions(input, self.padding_idx, self.left_pad)
        return tab.index_select(0, positions.flatten()).view(bsz, seq_len, -1).detach()

    def max_positions(self):
        return int(1e5)
