"""Oracle: centring + zero-padding collate (numpy).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Restatement of ``models/data.py``: ``_process_conformer``'s centring
(:166-172) and ``_collate_single_batch`` (:219-266); pinned on ``tests/golden/data.npz``, which
``tests/golden/make_golden.py::gen_data`` produces by calling those reference functions."""
from __future__ import annotations

import numpy as np


def center(n, ca, c, mask):
    """:166-172 (float32 like the reference)."""
    valid = ca[mask.astype(bool)]
    if len(valid) > 0:
        cen = valid.mean(axis=0, dtype=np.float32)
        return n - cen, ca - cen, c - cen
    return n, ca, c


def collate(batch):
    """:219-266: list of (n, ca, c, mask, emb|None, dih, labels) -> padded arrays."""
    B, Lmax = len(batch), max(b[0].shape[0] for b in batch)
    D = next((b[4].shape[-1] for b in batch if b[4] is not None), 0)
    out = [np.zeros((B, Lmax, 3), np.float32), np.zeros((B, Lmax, 3), np.float32), np.zeros((B, Lmax, 3), np.float32),
           np.zeros((B, Lmax), np.float32), np.zeros((B, Lmax, D), np.float32) if D else None,
           np.zeros((B, Lmax, 6), np.float32), np.zeros((B, Lmax), np.int64)]
    for i, (n, ca, c, m, emb, dih, lbl) in enumerate(batch):
        L = n.shape[0]
        out[0][i, :L], out[1][i, :L], out[2][i, :L], out[3][i, :L], out[5][i, :L], out[6][i, :L] = n, ca, c, m, dih, lbl
        if D and emb is not None:
            out[4][i, :L] = emb
    return out
