"""Per-conformer attention over packed rows for the encoder (``models/encoder.py:125-140``, ``:188-197``).

``self_attention`` is multi-head self-attention of every conformer over its own valid residues -- what
``nn.MultiheadAttention(key_padding_mask=~mask)`` computes for the valid rows; ``pooled_attention`` is the single-query
attention pooling of ``HierLatent``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _dest(pk):
    """Row of packed residue n in a ``[B, lmax]`` padded layout."""
    d = pk.__dict__.get("_dest")
    if d is None:
        n = torch.arange(pk.N, device=pk.cu.device)
        d = pk.conf * pk.lmax + (n - pk.cu[:-1].long().index_select(0, pk.conf))
        pk._dest = d
        pk._keymask = (torch.arange(pk.lmax, device=pk.cu.device)[None, :]
                       < torch.tensor(pk.lengths, device=pk.cu.device)[:, None])
    return d


def self_attention(qkv, pk, nheads: int, dropout_p: float = 0.0, precise: bool = False):
    """``qkv [N, 3d]`` (packed rows, ``[q | k | v]``) -> attention output ``[N, d]`` (heads concatenated, before the output
    projection)."""
    N, d3 = qkv.shape
    d = d3 // 3
    hd = d // nheads
    B, lmax = pk.B, pk.lmax
    uniform = all(n == lmax for n in pk.lengths)
    dest = None if uniform else _dest(pk)
    padded = qkv if uniform else torch.zeros(B * lmax, d3, device=qkv.device, dtype=qkv.dtype).index_copy(0, dest, qkv)
    q, k, v = padded.view(B, lmax, 3, nheads, hd).permute(2, 0, 3, 1, 4)
    mask = None if uniform else pk._keymask[:, None, None, :]
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=dropout_p)
    o = o.permute(0, 2, 1, 3).reshape(B * lmax, d)
    return o if uniform else o.index_select(0, dest)


def pooled_attention(q, kv, pk, nheads: int, dropout_p: float = 0.0):
    """One query ``q [1,d]`` (already projected) against every conformer's keys / values ``kv [N, 2d]`` -> ``[B, d]``
    (before the output projection); an empty conformer pools to zeros."""
    N, d2 = kv.shape
    d = d2 // 2
    hd = d // nheads
    B, lmax = pk.B, pk.lmax
    dest = _dest(pk)
    scores = (kv[:, :d].reshape(N, nheads, hd) * q.reshape(1, nheads, hd)).sum(-1) / math.sqrt(hd)        # [N, h]
    S = torch.full((B * lmax, nheads), float("-inf"), device=kv.device, dtype=kv.dtype).index_copy(0, dest, scores)
    w = torch.nan_to_num(torch.softmax(S.view(B, lmax, nheads), dim=1), nan=0.0)
    if dropout_p > 0:
        w = F.dropout(w, dropout_p)
    Vp = torch.zeros(B * lmax, d, device=kv.device, dtype=kv.dtype).index_copy(0, dest, kv[:, d:])
    return (w.unsqueeze(-1) * Vp.view(B, lmax, nheads, hd)).sum(1).reshape(B, d)
