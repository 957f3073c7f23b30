"""Node-level dense layers of the bf16 path.

The node-level linears (``AB = h [Wa;Wb]^T``, ``phi_h``, ``input_embedding`` ...) are plain library
GEMMs.  cuBLAS' true-fp32 path runs them on the SIMT pipe (~35 TFLOP/s on a B200); here the same product
is evaluated on the bf16 tensor cores with an error-compensated split,

    x = x_hi + x_lo,  W = W_hi + W_lo   (bf16 + bf16, exact to 2^-16)
    x W^T ~= [x_hi | x_hi | x_lo] [W_hi | W_lo | W_hi]^T      (one bf16 GEMM, K' = 3K, fp32 accumulate/output)

which drops only ``x_lo W_lo`` (relative 2^-16 ~ 1.5e-5 per product, ~1e-6 after accumulation) -- fp32-grade
results at ~8x the SIMT throughput.  Forward and both backward products use the same scheme.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_SUPPORTED: bool | None = None


def _split(t: torch.Tensor):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def _mm3(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fp32 ``a @ b`` (``a[M,K]``, ``b[K,N]``, both fp32) through one K'=3K bf16 GEMM."""
    ah, al = _split(a)
    bh, bl = _split(b)
    a3 = torch.cat([ah, ah, al], 1)
    b3 = torch.cat([bh, bl, bh], 0)
    return torch.mm(a3, b3, out_dtype=torch.float32)


def supported(device) -> bool:
    """``torch.mm(bf16, bf16, out_dtype=float32)`` is available on this build/device."""
    global _SUPPORTED
    if _SUPPORTED is None:
        try:
            a = torch.zeros(16, 16, dtype=torch.bfloat16, device=device)
            torch.mm(a, a, out_dtype=torch.float32)
            _SUPPORTED = True
        except (TypeError, RuntimeError):
            _SUPPORTED = False
    return _SUPPORTED


class LinearX3(torch.autograd.Function):
    """``y = x W^T + b`` with fp32-grade accuracy on the bf16 tensor cores (see module docstring)."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        y = _mm3(x, W.t())
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = g.contiguous()
        gx = _mm3(g, W) if ctx.needs_input_grad[0] else None
        gW = _mm3(g.t(), x) if ctx.needs_input_grad[1] else None
        gb = g.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gW, gb


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, fast: bool) -> torch.Tensor:
    """``F.linear`` for 2-D ``x``; ``fast`` selects the split-bf16 tensor-core evaluation on CUDA."""
    if fast and x.is_cuda and x.dim() == 2 and x.shape[0] >= 1024 and supported(x.device):
        return LinearX3.apply(x, weight, bias)
    return F.linear(x, weight, bias)
