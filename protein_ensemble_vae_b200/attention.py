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


NT, NN, TN = 0, 1, 2
_COUNTER = [0]


def _tables(pk):
    """Padded-row bookkeeping of the score buffer: conformer b owns rows cup[b] .. cup[b+1] (multiples of 128)."""
    t = pk.__dict__.get("_attn_tables")
    if t is None:
        dev = pk.cu.device
        cup, conf = [0], []
        for b, n in enumerate(pk.lengths):
            tiles = (n + 127) // 128
            conf += [b] * tiles
            cup.append(cup[-1] + 128 * tiles)
        t = pk._attn_tables = (torch.tensor(conf, dtype=torch.int32, device=dev), torch.tensor(cup, dtype=torch.int32, device=dev),
                               len(conf), max(32, (pk.lmax + 31) // 32 * 32))
        n = torch.arange(pk.N, device=dev)
        cu64, cup64 = pk.cu.long(), t[1].long()
        pk._row_pad = (cup64.index_select(0, pk.conf) + (n - cu64.index_select(0, pk.conf))).to(torch.int32)   # padded row of row i
    return t


def _gemm(form, A, a_col0, Bm, b_col0, pk, H, hd, scale, out, out_col0=0):
    from . import _lib
    from ._lib import ptr, stream
    tile_conf, cup, m_tiles, Lpad = _tables(pk)
    _lib.lib().call("pev_attn_gemm", form, ptr(A), A.shape[-1], a_col0, ptr(Bm), Bm.shape[-1], b_col0, ptr(tile_conf), ptr(pk.cu),
                    ptr(cup), m_tiles, H, hd, Lpad, pk.N, float(scale), ptr(out), out.shape[-1], out_col0, stream(out))


def _scores(mode, A, a_col0, Bm, b_col0, pk, H, hd, scale, out, out2=None, P=None, delta=None, p_drop=0.0, seed=0):
    """NT GEMM with the softmax (mode 1) or its backward (mode 2) fused into the epilogue (``pev_attn_scores``)."""
    from . import _lib
    from ._lib import ptr, stream
    tile_conf, cup, m_tiles, Lpad = _tables(pk)
    _lib.lib().call("pev_attn_scores", mode, ptr(A), A.shape[-1], a_col0, ptr(Bm), Bm.shape[-1], b_col0, ptr(tile_conf), ptr(pk.cu),
                    ptr(cup), m_tiles, H, hd, Lpad, pk.N, float(scale), ptr(out), ptr(out2), ptr(P), ptr(delta), float(p_drop),
                    int(seed), stream(out))


def _softmax(backward, S, G, pk, H, p_drop, seed):
    from . import _lib
    from ._lib import ptr, stream
    tile_conf, cup, m_tiles, Lpad = _tables(pk)
    _lib.lib().call("pev_attn_softmax", int(backward), ptr(S), ptr(G), ptr(tile_conf), ptr(pk.cu), ptr(cup), m_tiles, H, Lpad,
                    float(p_drop), int(seed), stream(S))


class PackedSelfAttention(torch.autograd.Function):
    """Multi-head self-attention of every conformer over its own rows on the tcgen05 kernels of ``csrc/attn_kernels.cu``
    (``pev_attn_gemm`` / ``pev_attn_softmax``): scores, softmax (+ dropout), ``P V`` forward; ``dV``, ``dP``, softmax
    backward, ``dQ``, ``dK`` backward.  The probabilities are kept for the backward pass (``[H, Np, Lpad]`` fp32)."""

    @staticmethod
    def forward(ctx, qkv, pk, nheads, p_drop):
        qkv = qkv.float().contiguous()
        N, d3 = qkv.shape
        d = d3 // 3
        hd = d // nheads
        _, _, m_tiles, Lpad = _tables(pk)
        _COUNTER[0] += 1
        seed = (torch.initial_seed() * 2654435761 + _COUNTER[0] * 40503) & 0xFFFFFFFF
        with torch.cuda.device_of(qkv):
            P = torch.empty(nheads, m_tiles * 128, Lpad, dtype=torch.float32, device=qkv.device)
            Pd = torch.empty_like(P) if p_drop > 0 else None
            if Lpad <= 256:          # whole score rows fit one key tile: softmax in the GEMM epilogue, S never reaches HBM
                _scores(1, qkv, 0, qkv, d, pk, nheads, hd, 1.0 / math.sqrt(hd), P, out2=Pd, p_drop=p_drop, seed=seed)
            else:
                _gemm(NT, qkv, 0, qkv, d, pk, nheads, hd, 1.0 / math.sqrt(hd), P)
                _softmax(0, P, Pd, pk, nheads, p_drop, seed)
            out = torch.empty(N, d, dtype=torch.float32, device=qkv.device)
            _gemm(NN, Pd if Pd is not None else P, 0, qkv, 2 * d, pk, nheads, hd, 1.0, out)
        ctx.pk, ctx.nheads, ctx.p_drop, ctx.seed = pk, nheads, p_drop, seed
        ctx.save_for_backward(qkv, P, Pd, out)
        return out

    @staticmethod
    def backward(ctx, go):
        qkv, P, Pd, out = ctx.saved_tensors
        pk, H = ctx.pk, ctx.nheads
        go = go.float().contiguous()
        N, d3 = qkv.shape
        d = d3 // 3
        hd = d // H
        sc = 1.0 / math.sqrt(hd)
        with torch.cuda.device_of(qkv):
            gqkv = torch.empty_like(qkv)
            _gemm(TN, Pd if Pd is not None else P, 0, go, 0, pk, H, hd, 1.0, gqkv, 2 * d)          # dV = P^T dO
            G = torch.empty_like(P)
            # dS = P (keep dP / (1 - p) - delta) in the epilogue of dP = dO V^T, delta_i = <dO_i, O_i> per head
            from . import _lib
            from ._lib import ptr, stream
            delta = torch.empty(H, P.shape[1], dtype=torch.float32, device=qkv.device)
            _lib.lib().call("pev_attn_delta", ptr(go), ptr(out), N, H, hd, ptr(pk._row_pad), P.shape[1], ptr(delta), stream(go))
            _scores(2, go, 0, qkv, 2 * d, pk, H, hd, 1.0, G, P=P, delta=delta, p_drop=ctx.p_drop, seed=ctx.seed)
            _gemm(NN, G, 0, qkv, d, pk, H, hd, sc, gqkv, 0)                                        # dQ = scale dS K
            _gemm(TN, G, 0, qkv, 0, pk, H, hd, sc, gqkv, d)                                        # dK = scale dS^T Q
        return gqkv, None, None, None


def self_attention(qkv, pk, nheads: int, dropout_p: float = 0.0, precise: bool = False):
    """``qkv [N, 3d]`` (packed rows, ``[q | k | v]``) -> attention output ``[N, d]`` (heads concatenated, before the output
    projection).  TF32 path: the repo's own tensor-core kernels (:class:`PackedSelfAttention`); the exact path
    (``precise``) evaluates the same attention in plain fp32 through torch's scaled_dot_product_attention."""
    N, d3 = qkv.shape
    if not precise and PackedSelfAttention is not None and qkv.is_cuda and (d3 // 3) // nheads in (64, 128) and N > 0:
        return PackedSelfAttention.apply(qkv, pk, nheads, float(dropout_p))
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
