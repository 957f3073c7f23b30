"""Packed residue graphs on the device.

The reference rebuilds a banded edge list with a Python double loop for every conformer of every
forward (``models/en_gnn_decoder.py:174-189``, ``:243``).  Here a whole batch of conformers is one
packed graph: nodes of conformer ``b`` are ``[cu_seqlens[b], cu_seqlens[b+1])`` and its edges are
the band ``0 < |i-j| <= W`` on those (compacted) residues, sorted by ``(i, j)`` -- so the in-edges
of a node are one contiguous CSR segment.  ``pev_band_graph_build`` writes CSR, the edge list, the
CSC permutation used by the backward column sums, and ``1/deg``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream


def band_edge_count(L: int, W: int) -> int:
    """Edges of one conformer (closed form of ``build_edge_index(L, ., W).shape[1]``)."""
    if L < 2 or W <= 0:
        return 0
    if W >= L - 1:
        return L * (L - 1)
    return 2 * W * L - W * (W + 1)


@dataclass
class PackedGraph:
    """CSR / CSC view of a packed batch of conformer graphs (int32 device tensors)."""
    num_nodes: int
    num_edges: int
    row_ptr: torch.Tensor            # [N+1]
    row: torch.Tensor                # [E]  destination i of edge e (sorted ascending)
    col: torch.Tensor                # [E]  source j
    col_ptr: torch.Tensor            # [N+1]
    csc_perm: torch.Tensor           # [E]  edge ids sorted by (col, position)
    dinv: torch.Tensor | None        # [N]  1/deg (float32) or None
    cu_seqlens: torch.Tensor | None = None
    lengths: tuple | None = None
    sort_perm: torch.Tensor | None = None   # generic graphs: positions of the sorted edges in the caller's order
    starts: torch.Tensor | None = None      # [N] bool, True at the first node of each conformer
    conf_of: torch.Tensor | None = None     # [N] int64, conformer of each packed node (band graphs; cached with them)

    def edge_index(self) -> torch.Tensor:
        """int64 ``[2, E]`` in the reference's layout."""
        return torch.stack([self.row.long(), self.col.long()], 0)


_BAND_CACHE: dict = {}


def band_graph(lengths, max_neighbors: int, device, cache: bool = True) -> PackedGraph:
    """Packed band graph for conformers of the given valid lengths (host ints)."""
    lengths = tuple(int(v) for v in lengths)
    W = int(max_neighbors)
    key = (lengths, W, str(device))
    if cache and key in _BAND_CACHE:
        return _BAND_CACHE[key]
    B = len(lengths)
    cu = np.zeros(B + 1, np.int64)
    np.cumsum(np.asarray(lengths, np.int64), out=cu[1:])
    eb = np.zeros(B + 1, np.int64)
    np.cumsum(np.asarray([band_edge_count(v, W) for v in lengths], np.int64), out=eb[1:])
    N, E = int(cu[-1]), int(eb[-1])
    if N >= 2 ** 31 or E >= 2 ** 31:
        raise ValueError("packed graph exceeds int32 indexing")
    dev = torch.device(device)
    with torch.cuda.device_of(torch.empty(0, device=dev)):
        cu_d = torch.from_numpy(cu.astype(np.int32)).to(dev)
        eb_d = torch.from_numpy(eb).to(dev)
        row_ptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
        row = torch.empty(E, dtype=torch.int32, device=dev)
        col = torch.empty(E, dtype=torch.int32, device=dev)
        csc = torch.empty(E, dtype=torch.int32, device=dev)
        dinv = torch.empty(N, dtype=torch.float32, device=dev)
        if N > 0:
            _lib.lib().call("pev_band_graph_build", ptr(cu_d), ptr(eb_d), B, W, N, ptr(row_ptr), ptr(row),
                            ptr(col), ptr(csc), ptr(dinv), stream(row))
        starts = torch.zeros(N, dtype=torch.bool, device=dev)
        first = [int(c) for c, v in zip(cu[:-1], lengths) if v > 0]
        if first:
            starts[torch.tensor(first, device=dev)] = True
        conf_of = torch.repeat_interleave(torch.arange(B, device=dev), torch.tensor(lengths, device=dev), output_size=N)
    g = PackedGraph(N, E, row_ptr, row, col, row_ptr, csc, dinv, cu_d, lengths, starts=starts, conf_of=conf_of)
    if cache:
        if len(_BAND_CACHE) > 64:
            _BAND_CACHE.clear()
        _BAND_CACHE[key] = g
    return g


def graph_from_edge_index(edge_index: torch.Tensor, num_nodes: int) -> PackedGraph:
    """CSR/CSC of an arbitrary ``[2,E]`` edge list (``EGNLayer.forward``'s public argument).

    Edges are stably sorted by destination, which keeps each node's neighbours in the caller's
    order -- the order CPU ``index_add_`` sums them in.
    """
    row, col = edge_index[0].long(), edge_index[1].long()
    E = row.numel()
    perm = torch.argsort(row, stable=True)
    row_s, col_s = row[perm], col[perm]
    row_ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=row.device)
    row_ptr[1:] = torch.cumsum(torch.bincount(row_s, minlength=num_nodes), 0)
    cperm = torch.argsort(col_s, stable=True)
    col_ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=row.device)
    col_ptr[1:] = torch.cumsum(torch.bincount(col_s, minlength=num_nodes), 0)
    i32 = lambda t: t.to(torch.int32).contiguous()  # noqa: E731
    return PackedGraph(num_nodes, E, i32(row_ptr), i32(row_s), i32(col_s), i32(col_ptr), i32(cperm), None,
                       sort_perm=perm)
