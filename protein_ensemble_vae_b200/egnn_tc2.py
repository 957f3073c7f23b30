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
    if W.is_inference():                                        # inference tensors keep no version counter: never cache
        cache = None
    key = (W.data_ptr(), W._version) if cache is not None else None
    capturing = W.is_cuda and torch.cuda.is_current_stream_capturing()   # a CUDA graph must contain the pack kernel
    if cache is not None and cache.get("key" + tag) == key and not capturing:
        return cache["img" + tag]
    Wc = f32c(W.detach())
    with torch.cuda.device_of(Wc):
        out = torch.empty(H * H, dtype=torch.bfloat16, device=W.device)
        _lib.lib().call("pev_pack_weight_bf16_scaled", ptr(Wc), int(transpose), float(scale), ptr(out), stream(Wc))
    if cache is not None and not capturing:
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


# ------------------------------------------------------------------------------------------------ one layer
_WORKSPACE: dict = {}


def _wgrad_workspace(device) -> torch.Tensor:
    ws = _WORKSPACE.get(device)
    if ws is None:
        ws = torch.empty(_lib.lib().cdll.pev_edge2_wgrad_workspace_bytes() // 4, dtype=torch.float32, device=device)
        _WORKSPACE[device] = ws
    return ws


_SCRATCH: dict = {}


def _scratch(device, E: int, N: int = 0):
    """One set of per-edge buffers (hv / m tile images, hs rows) + agg / w dummies for the recompute mode, shared by
    every layer of every decoder on ``device`` and grown on demand."""
    key = str(device)
    cur = _SCRATCH.get(key)
    if cur is None or cur[0].shape[0] < num_tiles(E) or cur[3].shape[0] < N:
        Ea, Na = max(E, cur[2].shape[0] if cur else 0), max(N, cur[3].shape[0] if cur else 0)
        cur = (alloc_tile_image(Ea, device), alloc_tile_image(Ea, device),
               torch.empty(max(Ea, 1), H, dtype=torch.bfloat16, device=device),
               torch.empty(max(Na, 1), H, dtype=torch.float32, device=device),
               torch.empty(max(Ea, 1), dtype=torch.float32, device=device))
        _SCRATCH[key] = cur
    return cur


class FusedEdgeV2(torch.autograd.Function):
    """(h, W1, b1, x, W2, b2, W5, b5, w6, b6, dinv, graph) -> (agg[N,256], x'[N,3]) on the v2 kernels.

    The fp16 half-domain node projection ``ABh = 0.5 [h Wa^T + b1 | h Wb^T]`` comes from one tensor-core GEMM with the
    scaling / bias / fp16 staging in its epilogue (``egnn_tc.node_abh``); its backward is part of this function.
    Forward: ``pev_edge_d2`` -> ``pev_edge2_fwd1`` -> ``pev_edge2_fwd2`` -> exact-order coordinate update (K2).
    Kept for the backward pass: the ``hv`` and ``m`` tile images and ``hs`` rows (bf16, 3 x 2.47 GB per layer at
    config 2), ``w``, ``d2`` and the fp16 ``ABh``.  With ``recompute`` the three per-edge streams are NOT kept: the
    backward pass re-runs ``fwd1`` / ``fwd2`` of the layer into one scratch set shared by all layers (+1 forward of
    edge work, 1/num_layers of the activation memory: what lets L=1024 dense graphs train).  Backward: K2 backward -> ``bwd2`` (ghv) -> ``wgrad5`` ->
    ``bwd1`` (ghu) -> ``wgrad2`` -> segmented row / column sums of ``ghu``; the activations a, m and the SiLU
    derivatives are rebuilt inside the kernels (one tanh each) instead of being stored.
    """

    @staticmethod
    def forward(ctx, h, W1, b1, x, W2, b2, W5, b5, w6, b6, dinv, g, keep: bool, caches, recompute: bool = False):
        from .egnn_tc import node_abh
        L = _lib.lib()
        N, E = g.num_nodes, g.num_edges
        dev = x.device
        x, h = f32c(x), f32c(h)
        W1d, b1d = W1.detach(), b1.detach()
        wd = f32c(W1d[:, 2 * H])
        b2, b5 = f32c(b2), f32c(b5)
        w6v, b6v = f32c(w6).reshape(-1), f32c(b6).reshape(-1)
        dinv = f32c(dinv)
        with torch.cuda.device_of(x):
            st = stream(x)
            ABb = node_abh(h.detach(), W1d, b1d)                 # fp16 [N,512] (see edge_tc2_kernels.cu)
            W2hp = packed_weight_scaled(W2, 0.5, cache=caches[0])
            W5hp = packed_weight_scaled(W5, 0.5, cache=caches[1])
            d2 = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
            store = keep and not recompute
            hvT = alloc_tile_image(E, dev) if store else None
            mT = alloc_tile_image(E, dev) if not (keep and recompute) else _scratch(dev, E)[1]
            hs = torch.empty(E, H, dtype=torch.bfloat16, device=dev) if store else None
            agg = torch.empty(N, H, dtype=torch.float32, device=dev)
            w = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
            x_out = torch.empty_like(x)
            L.call("pev_edge_d2", ptr(x), ptr(g.row), ptr(g.col), E, ptr(d2), st)
            with _lib.profiled("edge2_fwd1"):
                L.call("pev_edge2_fwd1", ptr(ABb), ptr(d2), ptr(wd), ptr(W2hp), ptr(b2), ptr(g.row), ptr(g.col), N, E,
                       ptr(hvT), ptr(mT), ptr(agg), st)
            with _lib.profiled("edge2_fwd2"):
                L.call("pev_edge2_fwd2", ptr(mT), ptr(W5hp), ptr(b5), ptr(w6v), ptr(b6v), E, ptr(w), ptr(hs), st)
            L.call("pev_scatter_coord_fwd", None, ptr(w), ptr(x), ptr(dinv), ptr(g.row_ptr), ptr(g.col), N, H,
                   None, ptr(x_out), st)
        ctx.g, ctx.caches = g, caches
        ctx.recompute = bool(keep and recompute)
        if ctx.recompute:
            ctx.save_for_backward(x, wd, W2, W5, w6v, dinv, ABb, None, None, None, w, d2, b2, b5, b6v, h, W1d)
        elif keep:
            ctx.save_for_backward(x, wd, W2, W5, w6v, dinv, ABb, hvT, mT, hs, w, d2, None, None, None, h, W1d)
        else:
            ctx.save_for_backward(x, wd, W2, W5, w6v, dinv, None, None, None, None, None, None, None, None, None, None,
                                  None)
        return agg, x_out

    @staticmethod
    def backward(ctx, gagg, gxo):
        x, wd, W2, W5, w6v, dinv, ABb, hvT, mT, hs, w, d2, b2, b5, b6v, h, W1d = ctx.saved_tensors
        if hs is None and not ctx.recompute:
            raise RuntimeError("FusedEdgeV2 ran with keep=False (no_grad); backward is unavailable")
        g = ctx.g
        N, E = g.num_nodes, g.num_edges
        if ctx.recompute:
            # rebuild hv, m, hs of this layer into the shared scratch set (same kernels, same bits as the forward pass)
            with torch.cuda.device_of(x):
                L, st = _lib.lib(), stream(x)
                hvT, mT, hs, agg_s, w_s = _scratch(x.device, E, N)
                W2hp = packed_weight_scaled(W2, 0.5, cache=ctx.caches[0])
                W5hp = packed_weight_scaled(W5, 0.5, cache=ctx.caches[1])
                with _lib.profiled("edge2_fwd1"):
                    L.call("pev_edge2_fwd1", ptr(ABb), ptr(d2), ptr(wd), ptr(W2hp), ptr(b2), ptr(g.row), ptr(g.col), N, E,
                           ptr(hvT), ptr(mT), ptr(agg_s), st)
                with _lib.profiled("edge2_fwd2"):
                    L.call("pev_edge2_fwd2", ptr(mT), ptr(W5hp), ptr(b5), ptr(w6v), ptr(b6v), E, ptr(w_s), ptr(hs), st)
        gagg, gxo = f32c(gagg), f32c(gxo)
        bf, f32 = torch.bfloat16, torch.float32
        L = _lib.lib()
        with torch.cuda.device_of(x):
            dev, st = x.device, stream(x)
            W5thp = packed_weight_scaled(W5, 0.5, transpose=True, cache=ctx.caches[1])
            W2thp = packed_weight_scaled(W2, 0.5, transpose=True, cache=ctx.caches[0])
            ws = _wgrad_workspace(dev)
            gw = torch.empty(max(E, 1), dtype=f32, device=dev)
            gx = torch.empty(N, 3, dtype=f32, device=dev)
            L.call("pev_scatter_coord_bwd", None, ptr(gxo), ptr(w), ptr(x), ptr(dinv), ptr(g.row_ptr), ptr(g.row),
                   ptr(g.col), ptr(g.col_ptr), ptr(g.csc_perm), N, E, H, None, ptr(gw), ptr(gx), st)
            ghvT = alloc_tile_image(E, dev)
            db2h = torch.empty(H, dtype=f32, device=dev)
            with _lib.profiled("edge2_bwd2"):
                L.call("pev_edge2_bwd2", ptr(hs), ptr(gw), ptr(w6v), ptr(W5thp), ptr(gagg), ptr(g.row), ptr(hvT), E,
                       ptr(ws), ptr(ghvT), ptr(db2h), st)
            gW5 = torch.empty(H, H, dtype=f32, device=dev)
            db5h, gw6 = torch.empty(H, dtype=f32, device=dev), torch.empty(H, dtype=f32, device=dev)
            with _lib.profiled("edge2_wgrad5"):
                L.call("pev_edge2_wgrad5", ptr(hs), ptr(gw), ptr(w6v), ptr(mT), E, ptr(ws), ptr(gW5), ptr(db5h),
                       ptr(gw6), st)
            ghu = torch.empty(E, H, dtype=bf, device=dev)
            gd2 = torch.empty(max(E, 1), dtype=f32, device=dev)
            gd2p = torch.empty(4 * max(E, 1), dtype=f32, device=dev)
            with _lib.profiled("edge2_bwd1"):
                L.call("pev_edge2_bwd1", ptr(ghvT), ptr(W2thp), ptr(ABb), ptr(d2), ptr(g.row), ptr(g.col), ptr(wd), E,
                       ptr(ghu), ptr(gd2), ptr(gd2p), st)
            del gd2p
            gW2 = torch.empty(H, H, dtype=f32, device=dev)
            with _lib.profiled("edge2_wgrad2"):
                L.call("pev_edge2_wgrad2", ptr(ghvT), ptr(ABb), ptr(d2), ptr(g.row), ptr(g.col), ptr(wd), E, ptr(ws),
                       ptr(gW2), st)
            del ghvT
            gAB = torch.empty(N, 2 * H, dtype=f32, device=dev)
            gwdh = torch.empty(H, dtype=f32, device=dev)
            with _lib.profiled("edge2_sums"):
                L.call("pev_edge2_sums", ptr(ghu), ptr(d2), ptr(g.row_ptr), ptr(g.col_ptr), ptr(g.csc_perm), N, E,
                       ptr(ws), ptr(gAB), ptr(gwdh), st)
            L.call("pev_edge_coord_bwd_accum", ptr(gd2), ptr(x), ptr(g.row_ptr), ptr(g.row), ptr(g.col), ptr(g.col_ptr),
                   ptr(g.csc_perm), N, E, ptr(gx), st)
            gwd = 0.5 * gwdh                            # hu = ... + (wd/2) d2
            gb6 = gw[:E].sum().reshape(1)
            from .egnn_tc import node_abh_backward
            gh, gWab, gb1 = node_abh_backward(gAB, h, W1d, ctx.needs_input_grad[0])
            gW1 = torch.cat([gWab[:H], gWab[H:], gwd.unsqueeze(1)], 1)            # [256, 513] = [gWa | gWb | gwd]
        return (gh, gW1, gb1, gx, gW2, 0.5 * db2h, gW5, 0.5 * db5h, gw6.reshape(1, H), gb6, None, None, None, None, None)


def egn_layer_v2(layer, h, x, g, dinv):
    """One EGNN layer with the edge MLP on the v2 kernels; ``layer`` is an ``EGNLayer`` (parameter holder)."""
    from .egnn_tc import NodePhiH
    W1 = layer.phi_e[0].weight                                            # [256, 513] = [Wa | Wb | wd]
    keep = torch.is_grad_enabled() and any(
        t.requires_grad for t in (h, x, W1, layer.phi_e[2].weight, layer.phi_x[0].weight))
    caches = layer.__dict__.setdefault("_pev_packed2", ({}, {}))
    agg, x_new = FusedEdgeV2.apply(h, W1, layer.phi_e[0].bias, x, layer.phi_e[2].weight, layer.phi_e[2].bias,
                                   layer.phi_x[0].weight, layer.phi_x[0].bias, layer.phi_x[2].weight,
                                   layer.phi_x[2].bias, dinv, g, keep, caches,
                                   bool(getattr(layer, "recompute_edges", False)))
    ln = layer.norm_h
    h_new = NodePhiH.apply(h, agg, layer.phi_h[0].weight, layer.phi_h[0].bias, layer.phi_h[2].weight,
                           layer.phi_h[2].bias, ln.weight, ln.bias, ln.eps)
    return h_new, x_new
