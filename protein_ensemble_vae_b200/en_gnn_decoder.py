"""E(n)-equivariant GNN decoder with the reference's module API (``models/en_gnn_decoder.py``).

Same classes, constructor arguments, ``forward`` signatures, parameter names / shapes (so the
reference's checkpoints load unchanged) and output conventions (exact zeros at
padded residues).  What differs is the execution: the reference loops over conformers in Python
and launches ~25 small ops per layer per conformer; here the batch's valid residues are packed
into one ragged node array, the banded graph is built once on the device, and each layer runs
as a handful of kernels over *all* conformers (they never interact).

``precision`` (constructor keyword, default from ``PEV_PRECISION`` or ``"bf16"``):
  ``"bf16"`` -- fused tcgen05 edge MLP with fp32 accumulation (needs ``hidden_dim == 256``),
  ``"fp32"`` -- exact-order fp32 path (any ``hidden_dim``).
CUDA tensors only; there is no CPU fallback.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import egnn_ops
from .graph import PackedGraph, band_graph, graph_from_edge_index

N_CA_LENGTH = 1.46      # models/en_gnn_decoder.py:274
CA_C_LENGTH = 1.52      # :275
PEPTIDE_LENGTH = 1.33   # :304


def _default_precision() -> str:
    return os.environ.get("PEV_PRECISION", "bf16")


def _default_recompute() -> bool:
    return os.environ.get("PEV_RECOMPUTE_EDGES", "0") not in ("0", "")


def _layer_forward(layer, h, x, g, dinv, precision):
    if precision == "bf16":
        from . import egnn_tc
        if egnn_tc.supports(layer):
            return egnn_tc.egn_layer_bf16(layer, h, x, g, dinv)
        raise RuntimeError("precision='bf16' needs node_dim == hidden_dim == 256 and SiLU activations; "
                           "construct the module with precision='fp32' for other shapes")
    if precision != "fp32":
        raise ValueError(f"unknown precision {precision!r}")
    return egnn_ops.egn_layer_fp32(layer, h, x, g, dinv)


class EGNLayer(nn.Module):
    """One EGNN layer (``models/en_gnn_decoder.py:15-87``).

    ``forward(h[N,D], x[N,3], edge_index[2,E], degree_inv[N]|None) -> (h', x')``; ``edge_index`` may be
    any edge list (rows are destinations) or an already built :class:`PackedGraph`.
    """

    def __init__(self, node_dim: int, hidden_dim: int, activation: nn.Module = nn.SiLU(),
                 precision: str | None = None, recompute_edges: bool | None = None):
        super().__init__()
        self.node_dim = node_dim
        self.hidden_dim = hidden_dim
        self.precision = precision or _default_precision()
        # bf16 path: rebuild the per-edge activations in backward instead of keeping them (egnn_tc2.FusedEdgeV2)
        self.recompute_edges = _default_recompute() if recompute_edges is None else bool(recompute_edges)
        self.phi_e = nn.Sequential(nn.Linear(2 * node_dim + 1, hidden_dim), activation,
                                   nn.Linear(hidden_dim, hidden_dim), activation)
        self.phi_h = nn.Sequential(nn.Linear(node_dim + hidden_dim, hidden_dim), activation,
                                   nn.Linear(hidden_dim, node_dim))
        self.phi_x = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), activation, nn.Linear(hidden_dim, 1))
        self.norm_h = nn.LayerNorm(node_dim)

    def invalidate_packed_weights(self) -> None:
        """Drop the cached bf16 tensor-core images of ``phi_e[2]`` / ``phi_x[0]``.  They are keyed on the parameters'
        version counters, which in-place writes through ``.data`` (e.g. ``p.data.copy_(ema)``) do not bump: call this
        after such a write (``load_state_dict`` and optimizer steps bump the counter and need nothing)."""
        self.__dict__.pop("_pev_packed2", None)

    def forward(self, h, x, edge_index, degree_inv=None):
        g = edge_index if isinstance(edge_index, PackedGraph) else graph_from_edge_index(edge_index, h.shape[0])
        return _layer_forward(self, h, x, g, degree_inv, self.precision)


class _Backbone(torch.autograd.Function):
    """``x_n = pull3(x_ca + 1.46 normalize(n_dir))``, ``x_c = x_ca + 1.52 normalize(c_dir)`` (``models/en_gnn_decoder.py:260-310``;
    a residue is pulled towards C of its predecessor unless it starts a conformer)."""

    @staticmethod
    def forward(ctx, n_dir, c_dir, x_ca, starts):
        from ctypes import c_void_p
        from . import _lib
        from ._lib import f32c, ptr, stream
        n_dir, c_dir = n_dir.float(), c_dir.float()
        if n_dir.stride(1) != 1:
            n_dir = n_dir.contiguous()
        if c_dir.stride(1) != 1:
            c_dir = c_dir.contiguous()
        x_ca = f32c(x_ca)
        st = starts.contiguous().view(torch.uint8)
        N = x_ca.shape[0]
        with torch.cuda.device_of(x_ca):
            x_n, x_c = torch.empty_like(x_ca), torch.empty_like(x_ca)
            _lib.lib().call("pev_backbone_fwd", c_void_p(n_dir.data_ptr()), n_dir.stride(0), c_void_p(c_dir.data_ptr()),
                            c_dir.stride(0), ptr(x_ca), ptr(st), N, ptr(x_n), ptr(x_c), stream(x_ca))
        ctx.save_for_backward(n_dir, c_dir, x_ca, st)
        return x_n, x_c

    @staticmethod
    def backward(ctx, g_xn, g_xc):
        from ctypes import c_void_p
        from . import _lib
        from ._lib import f32c, ptr, stream
        n_dir, c_dir, x_ca, st = ctx.saved_tensors
        N = x_ca.shape[0]
        with torch.cuda.device_of(x_ca):
            g_xn = f32c(g_xn) if g_xn is not None else torch.zeros_like(x_ca)
            g_xc = f32c(g_xc) if g_xc is not None else torch.zeros_like(x_ca)
            g_n, g_c, g_ca = torch.empty_like(x_ca), torch.empty_like(x_ca), torch.empty_like(x_ca)
            _lib.lib().call("pev_backbone_bwd", c_void_p(n_dir.data_ptr()), n_dir.stride(0), c_void_p(c_dir.data_ptr()),
                            c_dir.stride(0), ptr(x_ca), ptr(st), N, ptr(g_xn), ptr(g_xc), ptr(g_n), ptr(g_c), ptr(g_ca),
                            stream(x_ca))
        return g_n, g_c, g_ca, None


class EGNNDecoder(nn.Module):
    """Latents -> backbone (``models/en_gnn_decoder.py:90-333``)."""

    def __init__(self, z_g: int, z_l: int, hidden_dim: int = 256, num_layers: int = 8, max_neighbors: int = 20,
                 dropout: float = 0.1, degree_normalize: bool = True, precision: str | None = None,
                 recompute_edges: bool | None = None, cache_mask: bool = True):
        super().__init__()
        self.cache_mask = cache_mask
        self.z_g, self.z_l = z_g, z_l
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.max_neighbors = max_neighbors
        self.degree_normalize = degree_normalize
        self.precision = precision or _default_precision()
        self.dropout = nn.Dropout(dropout)
        self.input_embedding = nn.Linear(z_g + z_l, hidden_dim)
        self.layers = nn.ModuleList([EGNLayer(hidden_dim, hidden_dim, precision=self.precision,
                                              recompute_edges=recompute_edges) for _ in range(num_layers)])
        self.latent_to_coords = nn.Sequential(
            nn.Linear(z_g + z_l, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(), nn.Dropout(dropout * 0.5),
            nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(), nn.Linear(hidden_dim // 2, 3))
        with torch.no_grad():                                   # :135-137
            self.latent_to_coords[-1].weight.mul_(0.1)
            self.latent_to_coords[-1].bias.zero_()
        self.n_offset_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(),
                                           nn.Linear(hidden_dim // 2, 4))
        self.c_offset_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(),
                                           nn.Linear(hidden_dim // 2, 4))
        self.sequence_head = nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim * 2), nn.LayerNorm(hidden_dim * 2), nn.ReLU(), nn.Dropout(dropout * 0.5),
            nn.Linear(hidden_dim * 2, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(), nn.Dropout(dropout * 0.5),
            nn.Linear(hidden_dim, 20))

    # ------------------------------------------------------------------ graph helpers (static API)
    @staticmethod
    def build_edge_index(L: int, device, max_neighbors: int) -> torch.Tensor:
        """int64 ``[2,E]``: edges ``i <- j``, ``0 < |i-j| <= max_neighbors``, sorted by ``(i,j)``
        (``models/en_gnn_decoder.py:174-189``), built by ``pev_band_graph_build`` on ``device``."""
        L, W = int(L), int(max_neighbors)
        if W <= 0 and L >= 2:       # the reference's fallback chain list (:186-187), not row-sorted
            f = torch.arange(L - 1, device=device)
            return torch.stack([torch.cat([f, f + 1]), torch.cat([f + 1, f])]).long()
        return band_graph((L,), W, device).edge_index()

    @staticmethod
    def degrees(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
        """In-degree per destination (``models/en_gnn_decoder.py:191-198``)."""
        return torch.bincount(edge_index[0], minlength=num_nodes)

    # ------------------------------------------------------------------ forward
    def _graph(self, lengths, device) -> tuple[PackedGraph, torch.Tensor | None]:
        if self.max_neighbors <= 0:
            # reference fallback graph (chain), built per conformer and packed through the generic path
            rows, cols, off = [], [], 0
            for Lb in lengths:
                if Lb >= 2:
                    f = torch.arange(Lb - 1, device=device) + off
                    rows += [f, f + 1]
                    cols += [f + 1, f]
                off += Lb
            ei = torch.stack([torch.cat(rows), torch.cat(cols)]) if rows else torch.zeros(2, 0, dtype=torch.long,
                                                                                          device=device)
            g = graph_from_edge_index(ei, off)
            g.starts = torch.zeros(off, dtype=torch.bool, device=device)
            first, o2 = [], 0
            for Lb in lengths:
                if Lb > 0:
                    first.append(o2)
                o2 += Lb
            if first:
                g.starts[torch.tensor(first, device=device)] = True
            deg = torch.bincount(ei[0], minlength=off).float()
            dinv = torch.where(deg > 0, 1.0 / deg, torch.zeros_like(deg)) if self.degree_normalize else None
            return g, dinv
        g = band_graph(lengths, self.max_neighbors, device)
        return g, (g.dinv if self.degree_normalize else None)

    def invalidate_packed_weights(self) -> None:
        """See :meth:`EGNLayer.invalidate_packed_weights` (all layers)."""
        for layer in self.layers:
            layer.invalidate_packed_weights()

    def invalidate_mask_cache(self) -> None:
        """Forget the cached lengths / packing of the last mask (needed after out-of-band writes to it)."""
        self.__dict__.pop("_pev_mask_cache", None)

    def _mask_version(self, mask):
        if not self.__dict__.get("cache_mask", True):
            return None
        try:
            if mask.is_inference():
                return None
            return mask._version
        except RuntimeError:
            return None

    def _run(self, module, x, fp32_forward=False):
        """Node-level heads: plain modules on the fp32 path, TF32 tensor-core linears on the bf16 path
        (``fp32_forward`` keeps the forward product in fp32 for the ill-conditioned N / C direction heads)."""
        if self.precision == "bf16" and x.is_cuda:
            from .egnn_tc import apply_tf32
            return apply_tf32(module, x, fp32_forward)
        return module(x)

    def forward(self, z_g: torch.Tensor, z_l: torch.Tensor, mask: torch.Tensor | None = None):
        """``z_g[B,zg], z_l[B,L,zl], mask[B,L]|None -> (N, CA, C)[B,L,3], seq_logits[B,L,20]``."""
        B, L, _ = z_l.shape
        device = z_l.device
        if mask is not None:
            # The packed size needs the valid lengths on the host: one sync per forward -- unless this very mask
            # tensor (same object, same in-place version; the cache keeps it alive so its storage cannot be recycled)
            # was resolved on the previous call, as in a training loop over one resident batch.
            # The cache is keyed on the tensor object and its autograd version counter, so it cannot see writes that
            # bypass the counter (``mask.data.*``, numpy-shared storage, a CUDA graph refilling a static buffer):
            # after such a write call :meth:`invalidate_mask_cache`, or construct with ``cache_mask=False``.
            # Inference-mode tensors have no version counter and are never cached.
            ver = self._mask_version(mask)
            c = self.__dict__.get("_pev_mask_cache") if ver is not None else None
            if c is not None and c[0] is mask and c[1] == ver:
                lengths, flat_idx = c[2], c[3]
            else:
                mb = mask.bool()
                lengths = mb.sum(1).tolist()
                flat_idx = torch.nonzero(mb.reshape(-1)).squeeze(-1) if sum(lengths) != B * L else None
                if ver is not None:
                    self.__dict__["_pev_mask_cache"] = (mask, ver, lengths, flat_idx)
        else:
            lengths, flat_idx = [L] * B, None
        N = int(sum(lengths))
        out_shape3, out_shape20 = (B, L, 3), (B, L, 20)
        if N == 0:
            z3 = torch.zeros(out_shape3, device=device, dtype=z_l.dtype)
            return z3, z3.clone(), z3.clone(), torch.zeros(out_shape20, device=device, dtype=z_l.dtype)

        # pack valid residues: z = [z_g (replicated) | z_l]   (:233-234)
        zl_flat = z_l.reshape(B * L, -1)
        zl_p = zl_flat if flat_idx is None else zl_flat.index_select(0, flat_idx)
        g, dinv = self._graph(lengths, device)
        conf_of = g.conf_of                                     # cached with the band graph: no per-step H2D copy
        if conf_of is None:
            conf_of = torch.repeat_interleave(torch.arange(B, device=device),
                                              torch.tensor(lengths, device=device), output_size=N)
        z = torch.cat([z_g.index_select(0, conf_of), zl_p], -1)
        x = self._run(self.latent_to_coords, z)                 # :237
        h = self._run(self.input_embedding, z)                  # :240
        for layer in self.layers:                               # :248-250
            h, x = _layer_forward(layer, h, x, g, dinv, self.precision)
            h = self.dropout(h)
        logits = self._run(self.sequence_head, h)               # :253
        x_n, x_c = self._backbone(h, x, g)

        def unpack(t, width):
            if flat_idx is None:
                return t.reshape(B, L, width)
            full = torch.zeros(B * L, width, device=device, dtype=t.dtype)
            return full.index_put((flat_idx,), t).reshape(B, L, width)     # :313-328

        return unpack(x_n, 3), unpack(x, 3), unpack(x_c, 3), unpack(logits, 20)

    def _backbone(self, h, x_ca, g):
        """N / C placement and the 3-step peptide pull (``:260-310``) over the packed batch: the two direction heads, then one
        kernel per direction (``pev_backbone_fwd`` / ``pev_backbone_bwd``)."""
        n_out = self._run(self.n_offset_head, h, True)               # [N,4]: 4th channel unused in the reference too
        c_out = self._run(self.c_offset_head, h, True)
        return _Backbone.apply(n_out[:, :3], c_out[:, :3], x_ca, g.starts)


class SE3EquivariantDecoder(nn.Module):
    """``models/en_gnn_decoder.py:336-357``: hard-codes hidden 256 / 8 layers / 40 neighbours (F1)."""

    def __init__(self, z_g: int, z_l: int, hidden: int = 256, dropout: float = 0.2, equivariant: bool = True,
                 precision: str | None = None):
        super().__init__()
        self.decoder = EGNNDecoder(z_g=z_g, z_l=z_l, hidden_dim=256, num_layers=8, max_neighbors=40,
                                   dropout=dropout, degree_normalize=True, precision=precision)

    def forward(self, z_g, z_l, mask=None):
        return self.decoder(z_g, z_l, mask=mask)


class ResidueDecoder(nn.Module):
    """``models/en_gnn_decoder.py:359-392``.  ``equivariant=False`` is broken in the reference
    (undefined ``mlp_n``, F10); here it raises ``NotImplementedError`` instead of ``AttributeError``."""

    def __init__(self, z_g: int, z_l: int, hidden: int = 256, dropout: float = 0.2, equivariant: bool = True,
                 precision: str | None = None):
        super().__init__()
        self.equivariant = equivariant
        self.decoder = SE3EquivariantDecoder(z_g=z_g, z_l=z_l, hidden=hidden, dropout=dropout, equivariant=True,
                                             precision=precision)

    def forward(self, z_g, z_l, mask=None):
        if not self.equivariant:
            raise NotImplementedError("the reference's non-equivariant branch references undefined modules")
        return self.decoder(z_g, z_l, mask=mask)
