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
