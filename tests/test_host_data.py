"""Ragged packed batches: ``collate_packed`` + ``unpack_batch`` (one kernel: centre + pad on the device)
against the reference's own ``_process_conformer`` + ``_collate_single_batch`` outputs (tests/golden/data.npz, produced by
calling models/data.py) and against the numpy restatement."""
import os

import numpy as np
import torch

import cases
from oracle import data_oracle as do

G = os.path.join(os.path.dirname(__file__), "golden")
AA = {a: i for i, a in enumerate("ARNDCQEGHILKMFPSTWYV")}


def _items(center):
    out = []
    for c in cases.data_conformers():
        n, ca, cc, m = c["n"], c["ca"], c["c"], c["mask"]
        if center:
            n, ca, cc = do.center(n, ca, cc, m)
        lbl = np.array([AA.get(a, 0) for a in c["sequence"][:len(m)]], np.int64)
        out.append((n, ca, cc, m, c["seq_emb"], c["dihedrals"], lbl))
    return out


def test_data_oracle_matches_reference():
    gold = np.load(os.path.join(G, "data.npz"))
    n, ca, c, mask, emb, dih, lbl = do.collate(_items(center=True))
    for got, key in ((n, "n"), (ca, "ca"), (c, "c")):
        assert np.allclose(got, gold[key], atol=4e-6)            # float32 centroid: summation order differs
    for got, key in ((mask, "mask"), (emb, "emb"), (dih, "dih"), (lbl, "labels")):
        assert np.array_equal(got, gold[key])


def test_unpack_center_matches_reference(bk):
    from protein_ensemble_vae_b200 import data as pd
    gold = np.load(os.path.join(G, "data.npz"))
    items = [tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in it) for it in _items(center=False)]
    packed = pd.collate_packed(items, pin=False)
    assert packed["n"].shape == (37 + 64 + 5 + 50, 3) and packed["lmax"] == 64
    assert packed["cu_seqlens"].tolist() == [0, 37, 101, 106, 156]
    with bk.ctx():
        n, ca, c, mask, emb, dih, lbl = pd.unpack_batch({k: (v.to(bk.dev) if torch.is_tensor(v) else v) for k, v in packed.items()})
    for got, key in ((n, "n"), (ca, "ca"), (c, "c")):
        assert np.allclose(got.cpu().numpy(), gold[key], atol=4e-6), key
    for got, key in ((mask, "mask"), (emb, "emb"), (dih, "dih"), (lbl, "labels")):
        assert np.array_equal(got.cpu().numpy(), gold[key]), key
    assert float(n[2].abs().max()) > 1.0                          # fully masked conformer: not centred (:169)
    with bk.ctx():
        raw = pd.unpack_batch({k: (v.to(bk.dev) if torch.is_tensor(v) else v) for k, v in packed.items()}, center=False)
    assert torch.equal(raw[1][0, :37].cpu(), items[0][1]) and float(raw[1][0, 37:].abs().max()) == 0.0
    no_emb = dict(packed, emb=None)
    with bk.ctx():
        assert pd.unpack_batch({k: (v.to(bk.dev) if torch.is_tensor(v) else v) for k, v in no_emb.items()})[4] is None
