"""bf16 path of one EGNN layer: node-level operators and the layer entry point.

Per layer over the packed batch (``egnn_tc2.egn_layer_v2``)::

    ABh  = 0.5 [h Wa^T + b1 | h Wb^T]           node-level TF32 GEMM (cuBLAS), staged as fp16 [N,512]
    agg, w = tcgen05 edge kernels               csrc/edge_tc2_kernels.cu through egnn_tc2.FusedEdgeV2
    x'   = pev_scatter_coord_fwd(w, x, dinv)    exact-order coordinate update (K2)
    h'   = LayerNorm(h + phi_h([h, agg]))       node-level TF32 GEMMs + pev_add_layernorm_fwd

This module holds the node-level pieces: :class:`NodeLinear` (TF32 tensor-core linears with custom bias-gradient
column sums), :class:`AddLayerNorm` (fused residual + LayerNorm, one-pass backward) and helpers.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib
from ctypes import c_void_p

from ._lib import f32c, ptr, stream
from .graph import PackedGraph

H = 256
NODE_TF32_FORWARD = os.environ.get("PEV_NODE_TF32", "1") != "0"


def supports(layer) -> bool:
    return (layer.node_dim == H and layer.hidden_dim == H
            and all(isinstance(layer.phi_e[i], nn.SiLU) for i in (1, 3)) and isinstance(layer.phi_x[1], nn.SiLU)
            and isinstance(layer.phi_h[1], nn.SiLU) and layer.norm_h.elementwise_affine
            and layer.norm_h.bias is not None)


class _tf32_matmul:
    """Context: allow TF32 tensor-core GEMMs for plain fp32 ``torch.matmul`` calls (restores the flag on exit)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


_NODE_WS: dict = {}


def node_workspace(device) -> torch.Tensor:
    """Per-device scratch for the fixed-order column-sum reductions of the node-level kernels."""
    ws = _NODE_WS.get(device)
    if ws is None:
        ws = torch.empty(_lib.lib().cdll.pev_node_workspace_bytes() // 4, dtype=torch.float32, device=device)
        _NODE_WS[device] = ws
    return ws


def column_sum(g: torch.Tensor) -> torch.Tensor:
    """``g.sum(0)`` of a contiguous fp32 ``[N,D]`` (bias gradients) on ``pev_column_sum`` (bit-reproducible)."""
    N, D = g.shape
    if D % 4 or D > 4096 or g.dtype != torch.float32:
        return g.sum(0)
    with torch.cuda.device_of(g):
        out = torch.empty(D, dtype=torch.float32, device=g.device)
        _lib.lib().call("pev_column_sum", ptr(g), N, D, ptr(node_workspace(g.device)), ptr(out), stream(g))
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
            dgb = torch.empty(2 * D, dtype=torch.float32, device=r.device)     # adjacent: one reduction launch
            dg, db = dgb[:D], dgb[D:]
            _lib.lib().call("pev_layernorm_bwd", ptr(gy), ptr(r), ptr(gamma), ptr(mean), ptr(rstd), N, D,
                            ptr(node_workspace(r.device)), ptr(gr), ptr(dg), ptr(db), stream(r))
        return gr, (gr if ctx.has_res else None), dg, db, None


def layer_norm(module: nn.LayerNorm, x, res=None):
    """``module(x + res)`` through :class:`AddLayerNorm` when the shape is supported, else plain torch."""
    D = x.shape[-1]
    if (x.is_cuda and x.dim() == 2 and D in (128, 256, 512) and module.elementwise_affine and module.bias is not None
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


class NodeLinear2(torch.autograd.Function):
    """``[x1 | x2] W^T + b`` without materialising the concatenation (``phi_h[0]`` on ``[h, agg]``,
    ``models/en_gnn_decoder.py:71``): two accumulating TF32 GEMMs forward, contiguous gradients backward."""

    @staticmethod
    def forward(ctx, x1, x2, W, b):
        D1 = x1.shape[1]
        ctx.save_for_backward(x1, x2, W)
        with _tf32_matmul():
            y = torch.addmm(b, x1, W[:, :D1].t())
            return y.addmm_(x2, W[:, D1:].t())

    @staticmethod
    def backward(ctx, g):
        x1, x2, W = ctx.saved_tensors
        D1 = x1.shape[1]
        g = g.contiguous()
        with _tf32_matmul():
            g1 = g @ W[:, :D1] if ctx.needs_input_grad[0] else None
            g2 = g @ W[:, D1:] if ctx.needs_input_grad[1] else None
            gW = torch.cat([g.t() @ x1, g.t() @ x2], 1) if ctx.needs_input_grad[2] else None
        gb = column_sum(g) if ctx.needs_input_grad[3] else None
        return g1, g2, gW, gb


EPI_ABH, EPI_SILU, EPI_RES_LN, EPI_PLAIN, EPI_DSILU = range(5)


def transposed(W, scale: float = 1.0):
    """``scale * W^T`` as a contiguous tensor (``pev_transpose``: tiled, coalesced both ways) -- the weight operand of the
    data-gradient GEMMs."""
    W = f32c(W.detach())
    if not W.is_cuda or W.dim() != 2:
        return (scale * W).t().contiguous()
    rows, cols = W.shape
    with torch.cuda.device_of(W):
        out = torch.empty(cols, rows, dtype=torch.float32, device=W.device)
        _lib.lib().call("pev_transpose", ptr(W), rows, cols, float(scale), ptr(out), stream(W))
    return out


def node_gemm(epi, A1, W, bias=None, A2=None, scale=1.0, aux=None, gamma=None, beta=None, eps=1e-5, out2=False,
              stats=False):
    """``pev_node_gemm``: ``epilogue([A1 | A2] W^T)`` on the tensor cores (TF32 operands, fp32 accumulation), see
    ``include/pev_b200.h``.  Returns ``(out, out2, mean, rstd)`` (unused ones ``None``)."""
    M, K1 = A1.shape
    K2 = 0 if A2 is None else A2.shape[1]
    Nout = W.shape[0]
    dev = A1.device
    with torch.cuda.device_of(A1):
        out = torch.empty(M, Nout, dtype=torch.float16 if epi == EPI_ABH else torch.float32, device=dev)
        o2 = torch.empty(M, Nout, dtype=torch.float32, device=dev) if out2 else None
        mean = torch.empty(M, dtype=torch.float32, device=dev) if stats else None
        rstd = torch.empty(M, dtype=torch.float32, device=dev) if stats else None
        _lib.lib().call("pev_node_gemm", epi, ptr(A1), K1, ptr(A2), K2, ptr(W), ptr(bias), M, Nout, float(scale), ptr(aux),
                        ptr(gamma), ptr(beta), float(eps), ptr(out), ptr(o2), ptr(mean), ptr(rstd), stream(A1))
    return out, o2, mean, rstd


_WGRAD_WS: dict = {}


def node_wgrad(G, X, scale=1.0, out=None):
    """``scale * G^T X`` (``G[N,Mo]``, ``Mo`` in {256, 512}; ``X[N,256]``) on ``pev_node_wgrad`` (tcgen05 TF32 split-K,
    fixed-order reduction); ``out`` may be a ``[Mo,256]`` column block of a wider row-major matrix."""
    N, Mo = G.shape
    dev = G.device
    ws = _WGRAD_WS.get(dev)
    with torch.cuda.device_of(G):
        if ws is None:
            ws = _WGRAD_WS[dev] = torch.empty(_lib.lib().cdll.pev_node_wgrad_workspace_bytes() // 4, dtype=torch.float32,
                                              device=dev)
        if out is None:
            out = torch.empty(Mo, X.shape[1], dtype=torch.float32, device=dev)
        _lib.lib().call("pev_node_wgrad", ptr(G), Mo, ptr(X), N, float(scale), ptr(ws), c_void_p(out.data_ptr()),
                        out.stride(0), stream(G))
    return out


def node_abh(h, W1, b1):
    """fp16 half-domain node projection ``0.5 [h Wa^T + b1 | h Wb^T]`` ([N,512]) of ``phi_e[0]``'s factored form
    (``models/en_gnn_decoder.py:65-66``): one GEMM, scaling / bias / fp16 staging in its epilogue.  Not an autograd
    function: :class:`egnn_tc2.FusedEdgeV2` calls it and owns the backward (:func:`node_abh_backward`)."""
    Wcat = torch.cat([W1[:, :H], W1[:, H:2 * H]], 0).contiguous()               # [512, 256]
    bias = torch.cat([b1, torch.zeros_like(b1)]).contiguous()
    return node_gemm(EPI_ABH, f32c(h), Wcat, bias, scale=0.5)[0]


def node_abh_backward(gAB, h, W1, need_h=True):
    """``(gh, gWa|gWb as [256,512], gb1)`` from ``gAB = dL/dABh`` (fp32 [N,512])."""
    Wcat_t = transposed(torch.cat([W1[:, :H], W1[:, H:2 * H]], 0), 0.5)           # [256, 512]: gh = gAB (0.5 Wcat)
    gh = node_gemm(EPI_PLAIN, gAB, Wcat_t)[0] if need_h else None
    gWab = node_wgrad(gAB, h, 0.5)                                               # [512, 256] = [gWa ; gWb]
    gb1 = 0.5 * column_sum(gAB)[:H]
    return gh, gWab, gb1


class NodePhiH(torch.autograd.Function):
    """``LayerNorm(h + phi_h([h, agg]))`` (``models/en_gnn_decoder.py:70-73``) as two tensor-core GEMMs whose epilogues
    hold the SiLU and the residual + LayerNorm; backward: LayerNorm backward (one pass), two data-gradient GEMMs with
    the SiLU derivative in the first one's epilogue, weight gradients as library TF32 GEMMs."""

    @staticmethod
    def forward(ctx, h, agg, W3, b3, W4, b4, gamma, beta, eps):
        h, agg = f32c(h), f32c(agg)
        train = any(ctx.needs_input_grad)          # (grad mode is off inside forward: ask the context)
        W3c, W4c = f32c(W3.detach()), f32c(W4.detach())
        q, p, _, _ = node_gemm(EPI_SILU, h, W3c, f32c(b3.detach()), A2=agg, out2=train)
        y, r, mean, rstd = node_gemm(EPI_RES_LN, q, W4c, f32c(b4.detach()), aux=h, gamma=f32c(gamma.detach()),
                                     beta=f32c(beta.detach()), eps=eps, out2=train, stats=train)
        if train:
            ctx.save_for_backward(h, agg, W3c, W4c, f32c(gamma.detach()), p, q, r, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        h, agg, W3, W4, gamma, p, q, r, mean, rstd = ctx.saved_tensors
        gy = f32c(gy)
        N, D = r.shape
        with torch.cuda.device_of(r):
            gr = torch.empty_like(r)
            dgb = torch.empty(2 * D, dtype=torch.float32, device=r.device)
            _lib.lib().call("pev_layernorm_bwd", ptr(gy), ptr(r), ptr(gamma), ptr(mean), ptr(rstd), N, D,
                            ptr(node_workspace(r.device)), ptr(gr), ptr(dgb), ptr(dgb[D:]), stream(r))
        gp = node_gemm(EPI_DSILU, gr, transposed(W4), aux=p)[0]                   # (gr W4) * silu'(p)
        W3t = transposed(W3)                                                      # [512, 256]: rows = [h part ; agg part]
        gh = node_gemm(EPI_PLAIN, gp, W3t[:D], aux=gr)[0]                          # dL/dh = gr + gp W3[:, :D]  (residual in the epilogue)
        gagg = node_gemm(EPI_PLAIN, gp, W3t[D:])[0]                               # dL/dagg, contiguous for the edge kernels
        gW4 = node_wgrad(gr, q)
        gW3 = torch.empty(D, 2 * D, dtype=torch.float32, device=r.device)
        node_wgrad(gp, h, out=gW3[:, :D])
        node_wgrad(gp, agg, out=gW3[:, D:])
        return gh, gagg, gW3, column_sum(gp), gW4, column_sum(gr), dgb[:D], dgb[D:], None


# ------------------------------------------------------------------------------------------------- 3xTF32 (exact path)
def split_weight(W, transpose=False):
    """Split image ``[hi ; lo]`` (``[2R, C]``) of ``W`` (or of ``W^T``) for :func:`node_gemm3` (``pev_split_tf32``)."""
    W = f32c(W.detach())
    rows, cols = W.shape
    R, C = (cols, rows) if transpose else (rows, cols)
    with torch.cuda.device_of(W):
        out = torch.empty(2 * R, C, dtype=torch.float32, device=W.device)
        _lib.lib().call("pev_split_tf32", ptr(W), rows, cols, int(transpose), ptr(out), stream(W))
    return out


def node_gemm3(A, W3, bias=None, res=None):
    """``A W^T (+ bias) (+ res)`` with fp32-level accuracy on the TF32 tensor cores (``pev_node_gemm3``); ``W3`` is the
    split image of ``W [Nout,K]``."""
    M, K = A.shape
    Nout = W3.shape[0] // 2
    with torch.cuda.device_of(A):
        out = torch.empty(M, Nout, dtype=torch.float32, device=A.device)
        _lib.lib().call("pev_node_gemm3", ptr(A), K, ptr(W3), ptr(bias), M, Nout, ptr(res), ptr(out), stream(A))
    return out


def node_wgrad3(G, X, scale=1.0, out=None):
    """:func:`node_wgrad` with 3xTF32 products (``pev_node_wgrad3``)."""
    N, Mo = G.shape
    dev = G.device
    ws = _WGRAD_WS.get(dev)
    with torch.cuda.device_of(G):
        if ws is None:
            ws = _WGRAD_WS[dev] = torch.empty(_lib.lib().cdll.pev_node_wgrad_workspace_bytes() // 4, dtype=torch.float32,
                                              device=dev)
        if out is None:
            out = torch.empty(Mo, X.shape[1], dtype=torch.float32, device=dev)
        _lib.lib().call("pev_node_wgrad3", ptr(G), Mo, ptr(X), N, float(scale), ptr(ws), c_void_p(out.data_ptr()),
                        out.stride(0), stream(G))
    return out


def linear3x_supported(x, W) -> bool:
    return (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and W.shape[0] in (256, 512)
            and W.shape[1] == 256 and x.shape[1] == 256)


class Linear3x(torch.autograd.Function):
    """``[x1 | x2 ...] W^T + b`` (256-column operands ``xs``, never concatenated) on the tensor cores at fp32 accuracy:
    forward, ``g W`` and ``g^T x`` are all 3xTF32 (``csrc/node_gemm_kernels.cu``).  The exact path's replacement for
    ``nn.Linear`` in ``EGNLayer.forward`` (``models/en_gnn_decoder.py:61-79``)."""

    @staticmethod
    def forward(ctx, W, b, *xs):
        xs = tuple(f32c(x) for x in xs)
        Wd = f32c(W.detach())
        y = None
        for i, x in enumerate(xs):
            y = node_gemm3(x, split_weight(Wd[:, 256 * i:256 * (i + 1)]),
                           f32c(b.detach()) if (i == 0 and b is not None) else None, y)
        ctx.save_for_backward(Wd, *xs)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        Wd, *xs = ctx.saved_tensors
        g = f32c(g)
        Nout, K = Wd.shape
        gW = torch.empty(Nout, K, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[0] else None
        gxs = []
        for i, x in enumerate(xs):                                           # every operand is 256 columns wide
            Wk = Wd[:, 256 * i:256 * (i + 1)]
            gxs.append(node_gemm3(g, split_weight(Wk, transpose=True)) if ctx.needs_input_grad[2 + i] else None)
            if gW is not None:
                node_wgrad3(g, x, out=gW[:, 256 * i:256 * (i + 1)])
        gb = column_sum(g) if (ctx.has_bias and ctx.needs_input_grad[1]) else None
        return (gW, gb, *gxs)


def apply_tf32(module, x, fp32_forward=False):
    """Run an ``nn.Linear`` / ``nn.Sequential`` of the decoder with its linears on :class:`NodeLinear` (bf16 path)."""
    if isinstance(module, nn.Linear):
        from . import tc_linear
        if module.bias is not None and tc_linear.supported(x, module.weight):
            # own tcgen05 GEMMs: TF32, or 3xTF32 where the forward product must be fp32-accurate (N / C direction heads)
            return tc_linear.linear(x, module.weight, module.bias, precise=fp32_forward)
        return NodeLinear.apply(x, module.weight, module.bias, fp32_forward)
    if isinstance(module, nn.Sequential):
        for m in module:
            x = apply_tf32(m, x, fp32_forward)
        return x
    if isinstance(module, nn.LayerNorm):
        return layer_norm(module, x)
    return module(x)


def egn_layer_bf16(layer, h, x, g: PackedGraph, dinv):
    """One EGNN layer, edge MLP on the tensor cores (``egnn_tc2.py``); ``layer`` is an ``EGNLayer`` (parameter holder)."""
    from . import egnn_tc2
    return egnn_tc2.egn_layer_v2(layer, h, x, g, dinv)
