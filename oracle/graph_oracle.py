"""Integer oracle: banded residue graph (numpy).  Test infrastructure only.

Follows ``models/en_gnn_decoder.py:174-198`` of the reference.
"""
from __future__ import annotations

import numpy as np


def build_edge_index(L: int, max_neighbors: int) -> np.ndarray:
    """Edges ``i <- j`` for ``0 < |i-j| <= max_neighbors`` sorted by ``(i, j)``.

    Reference: ``EGNNDecoder.build_edge_index`` (``models/en_gnn_decoder.py:174-189``),
    including its fallback chain list when the window yields no pair
    (``:186-187``).  Returns int64 ``[2, E]``; ``[2, 0]`` when ``L < 2`` (the
    reference returns a shape-``[0]`` tensor there and then crashes, F10).
    """
    W = int(max_neighbors)
    i = np.arange(L, dtype=np.int64)[:, None]
    j = np.arange(L, dtype=np.int64)[None, :]
    keep = (np.abs(i - j) <= W) & (i != j)
    rows, cols = np.nonzero(keep)            # row-major order == sorted by (i, j)
    if rows.size == 0 and L >= 2:            # reference fallback, en_gnn_decoder.py:186-187
        fwd = np.arange(L - 1, dtype=np.int64)
        rows = np.concatenate([fwd, fwd + 1])
        cols = np.concatenate([fwd + 1, fwd])
    return np.stack([rows.astype(np.int64), cols.astype(np.int64)], axis=0)


def num_band_edges(L: int, W: int) -> int:
    """Closed form of ``build_edge_index(L, W).shape[1]``."""
    if L < 2:
        return 0
    if W >= L - 1:
        return L * (L - 1)
    if W == 0:
        return 2 * (L - 1)
    return 2 * W * L - W * (W + 1)


def degrees(edge_index: np.ndarray, num_nodes: int) -> np.ndarray:
    """In-degree per destination node.  Reference: ``models/en_gnn_decoder.py:191-198``."""
    return np.bincount(edge_index[0], minlength=num_nodes).astype(np.int64)


def packed_band_graph(lengths, W: int):
    """Band graphs of several conformers concatenated with global node ids.

    Returns ``(row_ptr[N+1], row[E], col[E])`` int64 -- what the CUDA graph
    builder must reproduce bit for bit.
    """
    rows, cols, off = [], [], 0
    for Lb in lengths:
        ei = build_edge_index(int(Lb), W)
        rows.append(ei[0] + off)
        cols.append(ei[1] + off)
        off += int(Lb)
    row = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    col = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    deg = np.bincount(row, minlength=off)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    return row_ptr, row, col
