"""General linear layers on the tcgen05 node-GEMM kernels (``pev_linear`` / ``pev_linear_wgrad``,
``csrc/node_gemm_kernels.cu``): forward, data gradient and weight gradient of ``y = act(x W^T + b (+ res))`` for row-major
fp32 activations of any width that is a multiple of 256, either single-pass TF32 or fp32-accurate 3xTF32.  Used by the
encoder (``encoder.py``); the EGNN layer has its own specialised epilogues (``egnn_tc.py``)."""
from __future__ import annotations

from ctypes import c_void_p

import torch

from . import _lib
from ._lib import stream
from .egnn_tc import _WGRAD_WS, column_sum, split_weight, transposed


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _rows(t):
    """2-D fp32 CUDA tensor whose rows are contiguous (column blocks of wider tensors are fine)."""
    if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % 4 or t.data_ptr() % 16:
        t = t.float().contiguous()
    return t


def supported(x, W) -> bool:
    """Output width a multiple of 128 (128 mod 256 runs on the 3xTF32 kernel, whose tiles are 128 wide), input width of 256."""
    return x.is_cuda and x.dim() == 2 and W.shape[0] % 128 == 0 and W.shape[1] % 256 == 0 and x.shape[1] == W.shape[1]


_COUNTER = [0]


def _next_seed() -> int:
    """Host-side dropout seed (no device sync): torch's seed mixed with a call counter."""
    _COUNTER[0] += 1
    return (torch.initial_seed() * 2654435761 + _COUNTER[0] * 40503 + 12345) & 0xFFFFFFFF


def mask_grad(g, y, p_drop, seed):
    """Backward of the dropout / ReLU epilogues: ``g / (1 - p)`` where the element was kept (and ``y != 0``), else 0."""
    g = g.float().contiguous()
    with torch.cuda.device_of(g):
        out = torch.empty_like(g)
        _lib.lib().call("pev_mask_grad", _p(g), _p(y), g.numel(), float(p_drop), int(seed), _p(out), stream(g))
    return out


def linear_fwd(x, W, bias=None, relu=False, res=None, precise=False, out=None, p_drop=0.0, seed=0):
    """``dropout(act(x W^T + bias)) + res``; ``W`` is the fp32 weight ``[Nout,K]`` (TF32) or its split image (``precise``)."""
    x = _rows(x)
    M, K = x.shape
    Nout = W.shape[0] // 2 if precise else W.shape[0]
    res = None if res is None else _rows(res)
    with torch.cuda.device_of(x):
        if out is None:
            out = torch.empty(M, Nout, dtype=torch.float32, device=x.device)
        _lib.lib().call("pev_linear", int(precise), _p(x), x.stride(0), K, _p(W), _p(bias), M, Nout, int(relu), float(p_drop),
                        int(seed), _p(res), 0 if res is None else res.stride(0), _p(out), out.stride(0), stream(x))
    return out


def linear_wgrad(g, x, precise=False):
    """``g^T x`` (``[Nout,K]``): every 256 x 256 block of the result in one launch (``pev_linear_wgrad``)."""
    g, x = _rows(g), _rows(x)
    N, Nout = g.shape
    K = x.shape[1]
    dev = g.device
    with torch.cuda.device_of(g):
        ws = _WGRAD_WS.get(dev)
        if ws is None:
            ws = _WGRAD_WS[dev] = torch.empty(_lib.lib().cdll.pev_node_wgrad_workspace_bytes() // 4, dtype=torch.float32,
                                              device=dev)
        gW = torch.empty(Nout, K, dtype=torch.float32, device=dev)
        _lib.lib().call("pev_linear_wgrad", int(precise), _p(g), g.stride(0), Nout, _p(x), x.stride(0), K, N, 1.0, _p(ws),
                        _p(gW), K, stream(g))
    return gW


class TCLinear(torch.autograd.Function):
    """``dropout(relu?(x W^T + b)) + res`` with all three GEMMs on the tensor cores; the dropout mask is a counter hash of
    (seed, element) applied in the GEMM epilogue and re-derived in the backward pass."""

    @staticmethod
    def forward(ctx, x, W, b, relu, precise, res, p_drop):
        assert not (relu and res is not None)
        Wd = W.detach().float().contiguous()
        bd = None if b is None else b.detach().float().contiguous()
        seed = _next_seed() if p_drop > 0 else 0
        y = linear_fwd(x, split_weight(Wd) if precise else Wd, bd, relu=relu, res=res, precise=precise, p_drop=p_drop, seed=seed)
        ctx.relu, ctx.precise, ctx.has_res, ctx.p_drop, ctx.seed = relu, precise, res is not None, p_drop, seed
        ctx.save_for_backward(x, Wd, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, Wd, y = ctx.saved_tensors
        g_res = g if ctx.has_res else None
        if ctx.relu or ctx.p_drop > 0:
            g = mask_grad(g, y, ctx.p_drop, ctx.seed)
        g = _rows(g)
        gx = gW = gb = None
        if ctx.needs_input_grad[0]:
            gx = linear_fwd(g, split_weight(transposed(Wd)) if ctx.precise else transposed(Wd), precise=ctx.precise)
        if ctx.needs_input_grad[1]:
            gW = linear_wgrad(g, x, ctx.precise)
        if ctx.needs_input_grad[2]:
            gb = column_sum(g.contiguous())
        return gx, gW, gb, None, None, g_res, None


def linear(x, W, b=None, relu=False, precise=False, res=None, p_drop=0.0):
    """``dropout(relu?(x W^T + b)) + res`` -- tensor-core path for supported shapes, else plain torch (tiny / odd-shaped
    linears)."""
    if supported(x, W):
        return TCLinear.apply(x, W, b, relu, precise or W.shape[0] % 256 != 0, res, float(p_drop))
    y = torch.nn.functional.linear(x, W, b)
    if relu:
        y = torch.relu(y)
    if p_drop > 0:
        y = torch.nn.functional.dropout(y, p_drop)
    return y if res is None else y + res
