"""Autograd operators of one EGNN layer over a :class:`~.graph.PackedGraph`.

Layer math (``EGNLayer.forward``, ``models/en_gnn_decoder.py:53-87``; derivation in SURVEY.md 8a)::

    rel = x_i - x_j;  d2 = |rel|^2
    u = Wa h_i + Wb h_j + wd d2 + b1      (first edge Linear, factored into node-level GEMMs)
    a = silu(u);  v = W2 a + b2;  m = silu(v)
    agg_i = sum_j m_ij;   s = W5 m + b5;  t = silu(s);  w = W6 t + b6
    x'_i = x_i + 0.2 dinv_i sum_j w_ij rel_ij

Two precisions:

* ``"fp32"`` -- exact-order path: :class:`EdgePrologue` (K1, fp32 form) -> dense fp32 edge MLP ->
  :class:`ScatterCoord` (K2, sequential ascending-edge sums, bit-identical to CPU ``index_add_``).
* ``"bf16"`` -- fused tcgen05 edge MLP (``egnn_tc.py`` / ``egnn_tc2.py``), H = 256.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import f32c, ptr, stream
from .graph import PackedGraph


class EdgePrologue(torch.autograd.Function):
    """``u[e] = A[row e] + B[col e] + wd |x_row - x_col|^2 + b1`` with ``AB = [A | B]`` (``[N, 2H]``)."""

    @staticmethod
    def forward(ctx, AB, x, wd, b1, g: PackedGraph):
        AB, x, wd, b1 = f32c(AB), f32c(x), f32c(wd), f32c(b1)
        H = wd.numel()
        with torch.cuda.device_of(AB):
            u = torch.empty(g.num_edges, H, dtype=torch.float32, device=AB.device)
            _lib.lib().call("pev_edge_prologue_fwd", ptr(AB), ptr(x), ptr(wd), ptr(b1), ptr(g.row), ptr(g.col),
                            g.num_edges, H, ptr(u), stream(AB))
        ctx.g = g
        ctx.save_for_backward(x, wd)
        return u

    @staticmethod
    def backward(ctx, gu):
        x, wd = ctx.saved_tensors
        g = ctx.g
        H = wd.numel()
        gu = f32c(gu)
        with torch.cuda.device_of(gu):
            dev = gu.device
            gAB = torch.empty(g.num_nodes, 2 * H, dtype=torch.float32, device=dev)
            gx = torch.empty(g.num_nodes, 3, dtype=torch.float32, device=dev)
            part = torch.empty(g.num_nodes, H, dtype=torch.float32, device=dev)
            gd2 = torch.empty(max(g.num_edges, 1), dtype=torch.float32, device=dev)
            _lib.lib().call("pev_edge_prologue_bwd", ptr(gu), ptr(x), ptr(wd), ptr(g.row_ptr), ptr(g.row),
                            ptr(g.col), ptr(g.col_ptr), ptr(g.csc_perm), g.num_nodes, g.num_edges, H,
                            ptr(gAB), ptr(gx), ptr(part), ptr(gd2), stream(gu))
        gwd = part.sum(0)
        gb1 = gAB[:, :H].sum(0)
        return gAB, gx, gwd, gb1, None


class ScatterCoord(torch.autograd.Function):
    """K2: ``agg = segment_sum(m)`` and ``x' = x + 0.2 dinv segment_sum(w rel)`` in edge order."""

    @staticmethod
    def forward(ctx, m, w, x, dinv, g: PackedGraph):
        m, w, x = f32c(m), f32c(w), f32c(x)
        dinv = f32c(dinv)
        H = m.shape[1]
        with torch.cuda.device_of(m):
            agg = torch.empty(g.num_nodes, H, dtype=torch.float32, device=m.device)
            x_out = torch.empty_like(x)
            _lib.lib().call("pev_scatter_coord_fwd", ptr(m), ptr(w), ptr(x), ptr(dinv), ptr(g.row_ptr),
                            ptr(g.col), g.num_nodes, H, ptr(agg), ptr(x_out), stream(m))
        ctx.g, ctx.H = g, H
        ctx.save_for_backward(w, x, dinv)
        return agg, x_out

    @staticmethod
    def backward(ctx, gagg, gxo):
        w, x, dinv = ctx.saved_tensors
        g, H = ctx.g, ctx.H
        gagg, gxo = f32c(gagg), f32c(gxo)
        with torch.cuda.device_of(gagg):
            dev = gagg.device
            gm = torch.empty(g.num_edges, H, dtype=torch.float32, device=dev)
            gw = torch.empty(g.num_edges, dtype=torch.float32, device=dev)
            gx = torch.empty(g.num_nodes, 3, dtype=torch.float32, device=dev)
            _lib.lib().call("pev_scatter_coord_bwd", ptr(gagg), ptr(gxo), ptr(w), ptr(x), ptr(dinv),
                            ptr(g.row_ptr), ptr(g.row), ptr(g.col), ptr(g.col_ptr), ptr(g.csc_perm),
                            g.num_nodes, g.num_edges, H, ptr(gm), ptr(gw), ptr(gx), stream(gagg))
        return gm, gw, gx, None, None


def _linear(lin, *xs):
    """``lin([x1 | x2 ...])``: 3xTF32 tensor-core GEMMs on the device (``egnn_tc.Linear3x``), plain torch elsewhere."""
    from . import egnn_tc
    if all(x.shape[1] == 256 for x in xs) and egnn_tc.linear3x_supported(xs[0], lin.weight[:, :256]) \
            and lin.weight.shape[1] == 256 * len(xs):
        return egnn_tc.Linear3x.apply(lin.weight, lin.bias, *xs)
    return lin(xs[0] if len(xs) == 1 else torch.cat(xs, -1))


def egn_layer_fp32(layer, h, x, g: PackedGraph, dinv):
    """One EGNN layer on the exact-order fp32 path; ``layer`` is an ``EGNLayer`` (parameter holder).  Every 256-wide
    linear (``phi_e[0]``'s node halves, ``phi_e[2]``, ``phi_x[0]``, ``phi_h``) runs on the tensor cores as a 3xTF32 product
    with fp32-level accuracy; ``phi_x[2]`` (256 -> 1) is a row dot product and stays a GEMV."""
    D = layer.node_dim
    W1 = layer.phi_e[0].weight                                    # [H, 2D+1] = [Wa | Wb | wd]
    Wab = torch.cat([W1[:, :D], W1[:, D:2 * D]], 0)               # [2H, D]: AB = h [Wa ; Wb]^T
    from . import egnn_tc
    if egnn_tc.linear3x_supported(h, Wab):
        AB = egnn_tc.Linear3x.apply(Wab, None, h)
    else:
        AB = h @ Wab.t()
    u = EdgePrologue.apply(AB, x, W1[:, 2 * D], layer.phi_e[0].bias, g)
    m = layer.phi_e[3](_linear(layer.phi_e[2], layer.phi_e[1](u)))   # silu -> Linear -> silu
    w = layer.phi_x[2](layer.phi_x[1](_linear(layer.phi_x[0], m))).squeeze(-1)   # [E]
    agg, x_new = ScatterCoord.apply(m, w, x, dinv, g)
    h_new = layer.norm_h(h + _linear(layer.phi_h[2], layer.phi_h[1](_linear(layer.phi_h[0], h, agg))))
    return h_new, x_new
