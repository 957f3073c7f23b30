"""bf16 tensor-core (tcgen05) path of one EGNN layer (K1, ``csrc/edge_tc_kernels.cu``).

Forward per layer over the packed batch::

    AB   = [h Wa^T + b1 | h Wb^T]              node-level fp32 GEMM (cuBLAS), stored as bf16 [N,512]
    v,agg = pev_edge_mlp1_fwd_bf16(AB, x, ...)  gather + SiLU -> tcgen05 GEMM W2 -> bias, SiLU, segment sum
    w    = pev_edge_mlp2_fwd_bf16(v, ...)       SiLU -> tcgen05 GEMM W5 -> bias, SiLU, dot w6
    x'   = pev_scatter_coord_fwd(w, x, dinv)    exact-order coordinate update (K2)
    h'   = LayerNorm(h + phi_h([h, agg]))       node-level fp32 (cuBLAS + torch)

For training the per-edge tensors ``a, silu'(u), m, silu'(v), s`` are kept in HBM as bf16 ``[E,256]`` for
the backward pass (2.47 GB each per layer at L=256, B=256: 74 GB for 6 layers -- sized for the 180 GB of a
B200); storing the SiLU derivatives keeps every transcendental out of the backward epilogues.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import f32c, ptr, stream
from .graph import PackedGraph

H = 256
USE_V2 = os.environ.get("PEV_EDGE_V2", "1") != "0"      # v2 edge kernels (csrc/edge_tc2_kernels.cu)
NODE_TF32_FORWARD = os.environ.get("PEV_NODE_TF32", "1") != "0"


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


class _tf32_matmul:
    """Context: allow TF32 tensor-core GEMMs for plain fp32 ``torch.matmul`` calls (restores the flag on exit)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


def column_sum(g: torch.Tensor) -> torch.Tensor:
    """``g.sum(0)`` of a contiguous fp32 ``[N,D]`` (bias gradients) on ``pev_column_sum``."""
    N, D = g.shape
    if D % 4 or g.dtype != torch.float32:
        return g.sum(0)
    with torch.cuda.device_of(g):
        out = torch.empty(D, dtype=torch.float32, device=g.device)
        _lib.lib().call("pev_column_sum", ptr(g), N, D, ptr(out), stream(g))
    return out


class AddLayerNorm(torch.autograd.Function):
    """``LayerNorm(x + res)`` (``res`` may be None) on ``pev_add_layernorm_fwd`` / ``pev_layernorm_bwd``."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps):
        x = f32c(x)
        res = f32c(res)
        N, D = x.shape
        with torch.cuda.device_of(x):
            r = torch.empty_like(x) if res is not None else x
            y = torch.empty_like(x)
            mean = torch.empty(N, dtype=torch.float32, device=x.device)
            rstd = torch.empty(N, dtype=torch.float32, device=x.device)
            g_, b_ = f32c(gamma.detach()), f32c(beta.detach())
            _lib.lib().call("pev_add_layernorm_fwd", ptr(x), ptr(res), ptr(g_), ptr(b_), float(eps), N, D,
                            ptr(r) if res is not None else None, ptr(y), ptr(mean), ptr(rstd), stream(x))
        ctx.save_for_backward(r, g_, mean, rstd)
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        r, gamma, mean, rstd = ctx.saved_tensors
        gy = f32c(gy)
        N, D = r.shape
        with torch.cuda.device_of(r):
            gr = torch.empty_like(r)
            dg = torch.empty(D, dtype=torch.float32, device=r.device)
            db = torch.empty(D, dtype=torch.float32, device=r.device)
            _lib.lib().call("pev_layernorm_bwd", ptr(gy), ptr(r), ptr(gamma), ptr(mean), ptr(rstd), N, D, ptr(gr),
                            ptr(dg), ptr(db), stream(r))
        return gr, (gr if ctx.has_res else None), dg, db, None


def layer_norm(module: nn.LayerNorm, x, res=None):
    """``module(x + res)`` through :class:`AddLayerNorm` when the shape is supported, else plain torch."""
    D = x.shape[-1]
    if (x.is_cuda and x.dim() == 2 and D in (256, 512) and module.elementwise_affine and module.bias is not None
            and tuple(module.normalized_shape) == (D,)):
        return AddLayerNorm.apply(x, res, module.weight, module.bias, module.eps)
    return module(x if res is None else x + res)


class NodeLinear(torch.autograd.Function):
    """``x W^T + b`` for the node-level linears of the bf16 path (plain library GEMMs).

    All three products (forward, ``g W`` and ``g^T x``) run on the tensor cores in TF32 with fp32 accumulation:
    TF32's 5e-4 relative rounding is an order of magnitude below the bf16 edge MLP's own (4e-3) and well inside the
    1e-2 output budget of the bf16 path; the fp32 path (``precision="fp32"``) never comes here.
    """

    @staticmethod
    def forward(ctx, x, W, b, fp32_forward=False):
        ctx.save_for_backward(x, W)
        if fp32_forward or not NODE_TF32_FORWARD:   # ill-conditioned consumers (N / C direction heads)
            return torch.addmm(b, x, W.t())
        with _tf32_matmul():
            return torch.addmm(b, x, W.t())

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = g.contiguous()
        with _tf32_matmul():
            gx = g @ W if ctx.needs_input_grad[0] else None
            gW = g.t() @ x if ctx.needs_input_grad[1] else None
        gb = column_sum(g) if ctx.needs_input_grad[2] else None
        return gx, gW, gb, None


def apply_tf32(module, x, fp32_forward=False):
    """Run an ``nn.Linear`` / ``nn.Sequential`` of the decoder with its linears on :class:`NodeLinear` (bf16 path)."""
    if isinstance(module, nn.Linear):
        return NodeLinear.apply(x, module.weight, module.bias, fp32_forward)
    if isinstance(module, nn.Sequential):
        for m in module:
            x = apply_tf32(m, x, fp32_forward)
        return x
    if isinstance(module, nn.LayerNorm):
        return layer_norm(module, x)
    return module(x)


def _wgrad(g_bf16: torch.Tensor, act_bf16: torch.Tensor) -> torch.Tensor:
    """``g^T @ act`` over the edge dimension ([256,E] x [E,256]); plain library GEMM (cuBLAS bf16, fp32 out)."""
    try:
        return torch.mm(g_bf16.t(), act_bf16, out_dtype=torch.float32)
    except TypeError:                                    # torch without mm(out_dtype=)
        return torch.mm(g_bf16.t(), act_bf16).float()


class FusedEdgeBF16(torch.autograd.Function):
    """(AB, x, wd, W2, b2, W5, b5, w6, b6, dinv, graph) -> (agg[N,256], x'[N,3]).

    Forward: stage-1 and stage-2 tcgen05 kernels + the exact-order coordinate update.  When a backward
    pass will follow, the per-edge tensors a, v, m, s are kept in HBM as bf16 [E,256].
    Backward: K2 coordinate backward -> stage-3 kernel (gs, gv) -> stage-4 kernel (gu, gd2)
    -> segmented row/column sums; the two weight gradients dW5 = gs^T m, dW2 = gv^T a are library GEMMs.
    """

    @staticmethod
    def forward(ctx, AB, x, wd, W2, b2, W5, b5, w6, b6, dinv, g: PackedGraph, keep: bool, caches):
        L = _lib.lib()
        AB, x, wd, b2, b5 = f32c(AB.detach()), f32c(x), f32c(wd), f32c(b2), f32c(b5)
        w6v, b6v = f32c(w6).reshape(-1), f32c(b6).reshape(-1)
        dinv = f32c(dinv)
        N, E = g.num_nodes, g.num_edges
        bf = torch.bfloat16
        with torch.cuda.device_of(x):
            dev = x.device
            W2p, W5p = packed_weight(W2, cache=caches[0]), packed_weight(W5, cache=caches[1])
            v = torch.empty(E, H, dtype=bf, device=dev)
            a, da, m, dm, s = (torch.empty(E, H, dtype=bf, device=dev) if keep else None for _ in range(5))
            agg = torch.empty(N, H, dtype=torch.float32, device=dev)
            w = torch.empty(E, dtype=torch.float32, device=dev)
            x_out = torch.empty_like(x)
            st = stream(x)
            with _lib.profiled("edge_mlp1"):
                L.call("pev_edge_mlp1_fwd_bf16", ptr(AB), ptr(x), ptr(wd), ptr(W2p), ptr(b2), ptr(g.row),
                       ptr(g.col), N, E, ptr(v), ptr(a), ptr(da), ptr(agg), st)
            with _lib.profiled("edge_mlp2"):
                L.call("pev_edge_mlp2_fwd_bf16", ptr(v), ptr(W5p), ptr(b5), ptr(w6v), ptr(b6v), E, ptr(w), ptr(s),
                       ptr(m), ptr(dm), st)
            L.call("pev_scatter_coord_fwd", None, ptr(w), ptr(x), ptr(dinv), ptr(g.row_ptr), ptr(g.col), N, H,
                   None, ptr(x_out), st)
        ctx.g, ctx.caches = g, caches
        ctx.save_for_backward(x, wd, W2, W5, w6v, dinv, a, da, m, dm, s, w)
        return agg, x_out

    @staticmethod
    def backward(ctx, gagg, gxo):
        x, wd, W2, W5, w6v, dinv, a, da, m, dm, s, w = ctx.saved_tensors
        if s is None:
            raise RuntimeError("FusedEdgeBF16 ran with keep=False (no_grad); backward is unavailable")
        g = ctx.g
        N, E = g.num_nodes, g.num_edges
        gagg, gxo = f32c(gagg), f32c(gxo)
        bf, f32 = torch.bfloat16, torch.float32
        L = _lib.lib()
        with torch.cuda.device_of(x):
            dev, st = x.device, stream(x)
            W5tp = packed_weight(W5, transpose=True, cache=ctx.caches[1])
            W2tp = packed_weight(W2, transpose=True, cache=ctx.caches[0])
            gw = torch.empty(E, dtype=f32, device=dev)
            gx = torch.empty(N, 3, dtype=f32, device=dev)
            L.call("pev_scatter_coord_bwd", None, ptr(gxo), ptr(w), ptr(x), ptr(dinv), ptr(g.row_ptr), ptr(g.row),
                   ptr(g.col), ptr(g.col_ptr), ptr(g.csc_perm), N, E, H, None, ptr(gw), ptr(gx), st)
            gs = torch.empty(E, H, dtype=bf, device=dev)
            gv = torch.empty(E, H, dtype=bf, device=dev)
            gb5, gw6 = torch.empty(H, dtype=f32, device=dev), torch.empty(H, dtype=f32, device=dev)
            with _lib.profiled("edge_mlp2_bwd"):
                L.call("pev_edge_mlp2_bwd_bf16", ptr(s), ptr(dm), ptr(gw), ptr(w6v), ptr(W5tp), ptr(gagg), ptr(g.row),
                       E, ptr(gs), ptr(gv), ptr(gb5), ptr(gw6), st)
            gW5 = _wgrad(gs, m)
            del gs
            gu = torch.empty(E, H, dtype=bf, device=dev)
            gd2 = torch.empty(max(E, 1), dtype=f32, device=dev)
            gb2 = torch.empty(H, dtype=f32, device=dev)
            with _lib.profiled("edge_mlp1_bwd"):
                L.call("pev_edge_mlp1_bwd_bf16", ptr(gv), ptr(da), ptr(W2tp), ptr(wd), E, ptr(gu), ptr(gd2), ptr(gb2), st)
            gW2 = _wgrad(gv, a)
            del gv
            gAB = torch.empty(N, 2 * H, dtype=f32, device=dev)
            part = torch.empty(N, H, dtype=f32, device=dev)
            with _lib.profiled("edge_prologue_bwd"):
                L.call("pev_edge_prologue_bwd_bf16", ptr(gu), ptr(gd2), ptr(x), ptr(g.row_ptr), ptr(g.row), ptr(g.col),
                       ptr(g.col_ptr), ptr(g.csc_perm), N, E, ptr(gAB), ptr(gx), ptr(part), st)
            gwd = part.sum(0)
            gb6 = gw.sum().reshape(1)
        return (gAB, gx, gwd, gW2, gb2, gW5, gb5, gw6.reshape(1, H), gb6, None, None, None, None)


def egn_layer_bf16(layer, h, x, g: PackedGraph, dinv):
    """One EGNN layer, edge MLP on the tensor cores; ``layer`` is an ``EGNLayer`` (parameter holder)."""
    W1 = layer.phi_e[0].weight                                            # [256, 513] = [Wa | Wb | wd]
    caches = layer.__dict__.setdefault("_pev_packed", ({}, {}))
    keep = torch.is_grad_enabled() and any(
        t.requires_grad for t in (h, x, W1, layer.phi_e[2].weight, layer.phi_x[0].weight))
    if USE_V2:
        from . import egnn_tc2
        return egnn_tc2.egn_layer_v2(layer, h, x, g, dinv)
    Wcat = torch.cat([W1[:, :H], W1[:, H:2 * H]], 0)                      # [512, 256]
    bias = torch.cat([layer.phi_e[0].bias, torch.zeros_like(layer.phi_e[0].bias)])
    AB = NodeLinear.apply(h, Wcat, bias)                                  # [N, 512]
    agg, x_new = FusedEdgeBF16.apply(AB, x, W1[:, 2 * H], layer.phi_e[2].weight, layer.phi_e[2].bias,
                                     layer.phi_x[0].weight, layer.phi_x[0].bias, layer.phi_x[2].weight,
                                     layer.phi_x[2].bias, dinv, g, keep, caches)
    q = layer.phi_h[1](NodeLinear.apply(torch.cat([h, agg], -1), layer.phi_h[0].weight, layer.phi_h[0].bias))
    h_new = layer.norm_h(h + NodeLinear.apply(q, layer.phi_h[2].weight, layer.phi_h[2].bias))
    return h_new, x_new
