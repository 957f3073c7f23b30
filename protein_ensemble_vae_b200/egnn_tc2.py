"""bf16 tensor-core (tcgen05) path of one EGNN layer, second generation (``csrc/edge_tc2_kernels.cu``).

Conventions of the v2 kernels (see the header of the .cu file):

* *half domain*: pre-activations are carried as ``h = z/2``; the factor is folded into the packed weight
  images (``0.5 W``), the biases and the node-level projection ``ABh = 0.5 [h Wa^T + b1 | h Wb^T]``;
* *tile images*: per-edge tensors written by a feature-lane epilogue (``hv``, ``ghv``) live in HBM as one 64 KB
  block per 128-edge tile, laid out as the SWIZZLE_128B shared-memory image the tensor core consumes
  (:func:`tile_image_to_rows` / :func:`rows_to_tile_image` convert to and from plain ``[E,256]`` rows --
  tests and tools only).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import f32c, ptr, stream

H = 256
TILE = 128
TILE_IMG_BYTES = TILE * H * 2


def packed_weight_scaled(W: torch.Tensor, scale: float, transpose: bool = False, cache: dict | None = None) -> torch.Tensor:
    """bf16 swizzled image of ``scale * W`` (or its transpose): the resident tcgen05 weight operand.

    ``cache`` (a dict owned by the module that owns ``W``) avoids repacking while the parameter is unchanged; it is
    keyed on the parameter's in-place version counter and storage pointer.
    """
    tag = ("T" if transpose else "N") + repr(float(scale))
    key = (W.data_ptr(), W._version)
    if cache is not None and cache.get("key" + tag) == key:
        return cache["img" + tag]
    Wc = f32c(W.detach())
    with torch.cuda.device_of(Wc):
        out = torch.empty(H * H, dtype=torch.bfloat16, device=W.device)
        _lib.lib().call("pev_pack_weight_bf16_scaled", ptr(Wc), int(transpose), float(scale), ptr(out), stream(Wc))
    if cache is not None:
        cache["key" + tag], cache["img" + tag] = key, out
    return out


def num_tiles(E: int) -> int:
    return (E + TILE - 1) // TILE


def alloc_tile_image(E: int, device) -> torch.Tensor:
    """Uninitialised tile-image buffer for ``E`` edges (bf16, ``[tiles, 128*256]``)."""
    return torch.empty(num_tiles(E), TILE * H, dtype=torch.bfloat16, device=device)


def _swz_index(device):
    r = torch.arange(8, device=device).view(8, 1)
    c = torch.arange(8, device=device).view(1, 8)
    return (c ^ r)                                            # [r, c] -> stored chunk position


def tile_image_to_rows(img: torch.Tensor, E: int) -> torch.Tensor:
    """Tile images ``[T, 128*256]`` -> rows ``[E, 256]`` (bf16).  Test / tool helper."""
    T = img.shape[0]
    v = img.view(T, 4, 2, 8, 8, 8, 8)                          # t, fq, eh, fg, r, cpos, i
    idx = _swz_index(img.device).view(1, 1, 1, 1, 8, 8, 1).expand(T, 4, 2, 8, 8, 8, 8)
    u = torch.gather(v, 5, idx)                               # t, fq, eh, fg, r, c, i
    rows = u.permute(0, 2, 5, 6, 1, 3, 4).reshape(T * TILE, H)  # t, (eh, c, i) = edge, (fq, fg, r) = feature
    return rows[:E].contiguous()


def rows_to_tile_image(rows: torch.Tensor) -> torch.Tensor:
    """Rows ``[E, 256]`` (bf16) -> tile images ``[T, 128*256]`` (zero padded).  Test / tool helper."""
    E = rows.shape[0]
    T = num_tiles(E)
    full = torch.zeros(T * TILE, H, dtype=rows.dtype, device=rows.device)
    full[:E] = rows
    u = full.view(T, 2, 8, 8, 4, 8, 8).permute(0, 4, 1, 5, 6, 2, 3).contiguous()   # t, fq, eh, fg, r, c, i
    idx = _swz_index(rows.device).view(1, 1, 1, 1, 8, 8, 1).expand(T, 4, 2, 8, 8, 8, 8)
    v = torch.empty_like(u)
    v.scatter_(5, idx, u)
    return v.view(T, TILE * H)


# ------------------------------------------------------------------------------------------------ layer forward
def edge_forward(ABh: torch.Tensor, x: torch.Tensor, wd, W2, b2, W5, b5, w6, b6, dinv, g, keep: bool, caches):
    """Edge half of one EGNN layer on the v2 kernels.

    ``ABh`` is the bf16 ``[N,512]`` half-domain node projection.  Returns ``(agg[N,256], x'[N,3], saved)`` where
    ``saved`` holds what the backward kernels need (``hvT`` tile images, ``hs`` rows, ``w``, ``d2``) when ``keep``.
    """
    L = _lib.lib()
    N, E = g.num_nodes, g.num_edges
    dev = x.device
    x = f32c(x)
    wd, b2, b5 = f32c(wd.detach()), f32c(b2.detach()), f32c(b5.detach())
    w6v, b6v = f32c(w6.detach()).reshape(-1), f32c(b6.detach()).reshape(-1)
    dinv = f32c(dinv)
    with torch.cuda.device_of(x):
        st = stream(x)
        W2hp = packed_weight_scaled(W2, 0.5, cache=caches[0])
        W5hp = packed_weight_scaled(W5, 0.5, cache=caches[1])
        d2 = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
        hvT = alloc_tile_image(E, dev)
        hs = torch.empty(E, H, dtype=torch.bfloat16, device=dev) if keep else None
        agg = torch.empty(N, H, dtype=torch.float32, device=dev)
        w = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
        x_out = torch.empty_like(x)
        L.call("pev_edge_d2", ptr(x), ptr(g.row), ptr(g.col), E, ptr(d2), st)
        with _lib.profiled("edge2_fwd1"):
            L.call("pev_edge2_fwd1", ptr(ABh), ptr(d2), ptr(wd), ptr(W2hp), ptr(b2), ptr(g.row), ptr(g.col), N, E,
                   ptr(hvT), ptr(agg), st)
        with _lib.profiled("edge2_fwd2"):
            L.call("pev_edge2_fwd2", ptr(hvT), ptr(W5hp), ptr(b5), ptr(w6v), ptr(b6v), E, ptr(w), ptr(hs), st)
        L.call("pev_scatter_coord_fwd", None, ptr(w), ptr(x), ptr(dinv), ptr(g.row_ptr), ptr(g.col), N, H,
               None, ptr(x_out), st)
    saved = (hvT, hs, w, d2) if keep else None
    return agg, x_out, saved


def node_projection_half(layer, h: torch.Tensor) -> torch.Tensor:
    """``ABh = 0.5 [h Wa^T + b1 | h Wb^T]`` as bf16 ``[N,512]`` (node-level library GEMM, TF32 tensor cores)."""
    from .egnn_tc import NodeLinear
    W1 = layer.phi_e[0].weight
    Wcat = 0.5 * torch.cat([W1[:, :H], W1[:, H:2 * H]], 0)
    bias = 0.5 * torch.cat([layer.phi_e[0].bias, torch.zeros_like(layer.phi_e[0].bias)])
    return NodeLinear.apply(h, Wcat, bias).to(torch.bfloat16)
