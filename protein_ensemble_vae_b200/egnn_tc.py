"""bf16 tensor-core (tcgen05) path of one EGNN layer (K1, ``csrc/edge_tc_kernels.cu``).

Forward per layer over the packed batch::

    AB   = [h Wa^T + b1 | h Wb^T]              node-level fp32 GEMM (cuBLAS), stored as bf16 [N,512]
    v,agg = pev_edge_mlp1_fwd_bf16(AB, x, ...)  gather + SiLU -> tcgen05 GEMM W2 -> bias, SiLU, segment sum
    w    = pev_edge_mlp2_fwd_bf16(v, ...)       SiLU -> tcgen05 GEMM W5 -> bias, SiLU, dot w6
    x'   = pev_scatter_coord_fwd(w, x, dinv)    exact-order coordinate update (K2)
    h'   = LayerNorm(h + phi_h([h, agg]))       node-level fp32 (cuBLAS + torch)

The per-edge pre-activations ``v`` and ``s`` are kept in HBM as bf16 ``[E,256]`` for the backward
pass (2.47 GB each per layer at L=256, B=256 -- sized for the 180 GB of a B200).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import f32c, ptr, stream
from .graph import PackedGraph

H = 256


def supports(layer) -> bool:
    return (layer.node_dim == H and layer.hidden_dim == H
            and all(isinstance(layer.phi_e[i], nn.SiLU) for i in (1, 3)) and isinstance(layer.phi_x[1], nn.SiLU))


def packed_weight(W: torch.Tensor, transpose: bool = False, cache: dict | None = None) -> torch.Tensor:
    """bf16 swizzled image of a 256x256 weight (the resident tcgen05 B operand).

    ``cache`` (a dict owned by the module that owns ``W``) avoids repacking while the parameter is
    unchanged; it is keyed on the parameter's in-place version counter and storage pointer.
    """
    key = ("T" if transpose else "N", W.data_ptr(), W._version)
    if cache is not None and cache.get("key" + key[0]) == key:
        return cache["img" + key[0]]
    Wc = f32c(W.detach())
    with torch.cuda.device_of(Wc):
        out = torch.empty(H * H, dtype=torch.bfloat16, device=W.device)
        _lib.lib().call("pev_pack_weight_bf16", ptr(Wc), int(transpose), ptr(out), stream(Wc))
    if cache is not None:
        cache["key" + key[0]], cache["img" + key[0]] = key, out
    return out


def _silu_and_grad(z):
    sg = torch.sigmoid(z)
    return z * sg, sg * (1.0 + z * (1.0 - sg))


class FusedEdgeBF16(torch.autograd.Function):
    """(AB, x, wd, W2, b2, W5, b5, w6, b6, dinv, graph) -> (agg[N,256], x'[N,3])."""

    @staticmethod
    def forward(ctx, AB, x, wd, W2, b2, W5, b5, w6, b6, dinv, g: PackedGraph, keep: bool, caches):
        L = _lib.lib()
        x, wd, b2, b5 = f32c(x), f32c(wd), f32c(b2), f32c(b5)
        w6v, b6v = f32c(w6).reshape(-1), f32c(b6).reshape(-1)
        dinv = f32c(dinv)
        N, E = g.num_nodes, g.num_edges
        with torch.cuda.device_of(x):
            dev = x.device
            ABh = AB.detach().to(torch.bfloat16).contiguous()
            W2p, W5p = packed_weight(W2, cache=caches[0]), packed_weight(W5, cache=caches[1])
            v = torch.empty(E, H, dtype=torch.bfloat16, device=dev)
            s = torch.empty(E, H, dtype=torch.bfloat16, device=dev) if keep else None
            agg = torch.empty(N, H, dtype=torch.float32, device=dev)
            w = torch.empty(E, dtype=torch.float32, device=dev)
            x_out = torch.empty_like(x)
            st = stream(x)
            with _lib.profiled("edge_mlp1"):
                L.call("pev_edge_mlp1_fwd_bf16", ptr(ABh), ptr(x), ptr(wd), ptr(W2p), ptr(b2), ptr(g.row),
                       ptr(g.col), N, E, ptr(v), ptr(agg), st)
            with _lib.profiled("edge_mlp2"):
                L.call("pev_edge_mlp2_fwd_bf16", ptr(v), ptr(W5p), ptr(b5), ptr(w6v), ptr(b6v), E, ptr(w), ptr(s), st)
            L.call("pev_scatter_coord_fwd", None, ptr(w), ptr(x), ptr(dinv), ptr(g.row_ptr), ptr(g.col), N, H,
                   None, ptr(x_out), st)
        ctx.g = g
        ctx.save_for_backward(ABh, x, wd, W2, W5, w6v, dinv, v, s, w)
        return agg, x_out

    @staticmethod
    def backward(ctx, gagg, gxo):
        ABh, x, wd, W2, W5, w6v, dinv, v, s, w = ctx.saved_tensors
        if s is None:
            raise RuntimeError("FusedEdgeBF16 was run without keep=True; backward is unavailable")
        g = ctx.g
        N, E = g.num_nodes, g.num_edges
        gagg, gxo = f32c(gagg), f32c(gxo)
        bf = torch.bfloat16
        with torch.cuda.device_of(x):
            dev = x.device
            gw = torch.empty(E, dtype=torch.float32, device=dev)
            gx = torch.empty(N, 3, dtype=torch.float32, device=dev)
            _lib.lib().call("pev_scatter_coord_bwd", None, ptr(gxo), ptr(w), ptr(x), ptr(dinv), ptr(g.row_ptr),
                            ptr(g.row), ptr(g.col), ptr(g.col_ptr), ptr(g.csc_perm), N, E, H, None, ptr(gw),
                            ptr(gx), stream(x))
            # Interim dense backward (torch elementwise + cuBLAS bf16 GEMMs), chunked over edges.
            W2b, W5b = W2.detach().to(bf), W5.detach().to(bf)
            gAB = torch.zeros(N, 2 * H, dtype=torch.float32, device=dev)
            gW2 = torch.zeros(H, H, dtype=torch.float32, device=dev)
            gW5 = torch.zeros(H, H, dtype=torch.float32, device=dev)
            gb2 = torch.zeros(H, dtype=torch.float32, device=dev)
            gb5 = torch.zeros(H, dtype=torch.float32, device=dev)
            gw6 = torch.zeros(H, dtype=torch.float32, device=dev)
            gwd = torch.zeros(H, dtype=torch.float32, device=dev)
            row, col = g.row.long(), g.col.long()
            chunk = 1 << 19
            for e0 in range(0, E, chunk):
                e1 = min(E, e0 + chunk)
                r, c = row[e0:e1], col[e0:e1]
                t, dt = _silu_and_grad(s[e0:e1].float())
                gs = gw[e0:e1, None] * w6v[None, :] * dt
                gw6 += gw[e0:e1] @ t
                gb5 += gs.sum(0)
                m, dm = _silu_and_grad(v[e0:e1].float())
                gsb = gs.to(bf)
                gW5 += (gsb.t() @ m.to(bf)).float()
                gv = ((gsb @ W5b).float() + gagg[r]) * dm
                gb2 += gv.sum(0)
                rel = x[r] - x[c]
                d2 = (rel * rel).sum(-1, keepdim=True)
                u = ABh[r, :H].float() + ABh[c, H:].float() + wd[None, :] * d2
                a, da = _silu_and_grad(u)
                gvb = gv.to(bf)
                gW2 += (gvb.t() @ a.to(bf)).float()
                gu = (gvb @ W2b).float() * da
                gAB[:, :H].index_add_(0, r, gu)
                gAB[:, H:].index_add_(0, c, gu)
                gwd += (gu * d2).sum(0)
                grel = (2.0 * (gu @ wd))[:, None] * rel
                gx.index_add_(0, r, grel)
                gx.index_add_(0, c, -grel)
            gb6 = gw.sum().reshape(1)
        return (gAB, gx, gwd, gW2, gb2, gW5, gb5, gw6.reshape(1, H), gb6, None, None, None, None)


def egn_layer_bf16(layer, h, x, g: PackedGraph, dinv):
    """One EGNN layer, edge MLP on the tensor cores; ``layer`` is an ``EGNLayer`` (parameter holder)."""
    W1 = layer.phi_e[0].weight                                            # [256, 513] = [Wa | Wb | wd]
    Wcat = torch.cat([W1[:, :H], W1[:, H:2 * H]], 0)                      # [512, 256]
    bias = torch.cat([layer.phi_e[0].bias, torch.zeros_like(layer.phi_e[0].bias)])
    AB = torch.addmm(bias, h, Wcat.t())                                   # [N, 512]
    caches = layer.__dict__.setdefault("_pev_packed", ({}, {}))
    keep = torch.is_grad_enabled() and any(
        t.requires_grad for t in (h, x, W1, layer.phi_e[2].weight, layer.phi_x[0].weight))
    agg, x_new = FusedEdgeBF16.apply(AB, x, W1[:, 2 * H], layer.phi_e[2].weight, layer.phi_e[2].bias,
                                     layer.phi_x[0].weight, layer.phi_x[0].bias, layer.phi_x[2].weight,
                                     layer.phi_x[2].bias, dinv, g, keep, caches)
    h_new = layer.norm_h(h + layer.phi_h(torch.cat([h, agg], -1)))
    return h_new, x_new
