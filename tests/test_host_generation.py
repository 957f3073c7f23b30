"""Generation driver pieces: the geometry validity filter against the oracle -- and, where the
reference is present, against the reference function itself (lifted by ast: its module needs h5py) -- plus the batched
driver on the host backend."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import generation_oracle

REF = "/root/reference/generate_ensemble_pdbs.py"


def _structures():
    rng = np.random.default_rng(0)
    L = 40
    helix = np.stack([[2.3 * np.cos(1.745 * i), 2.3 * np.sin(1.745 * i), 1.5 * i] for i in range(L)])   # 3.8 A steps
    out = [helix, helix * 1.8, helix * 0.5, np.cumsum(rng.standard_normal((L, 3)) * 2.2, 0),
           np.stack([[3.8 * i, 0.0, 0.0] for i in range(L)]), helix + 0.3 * rng.standard_normal((L, 3))]
    zig = np.zeros((L, 3))
    zig[1::2, 0] = 3.8
    zig[:, 1] = 0.5 * np.arange(L)                      # sharp CA-CA-CA angles (< 60 degrees)
    out.append(zig)
    far = helix.copy()
    far[20:] += 9.0                                     # one 6+ A jump
    out.append(far)
    return np.stack(out).astype(np.float32)


def test_geometry_filter_matches_oracle_and_reference(bk):
    from protein_ensemble_vae_b200 import validate_geometry_batch, validate_protein_geometry
    ca = _structures()
    S, L, _ = ca.shape
    masks = np.ones((S, L), np.float32)
    masks[1, 30:] = 0
    masks[3, 5:9] = 0                                   # interior gap: bridged
    masks[5] = 0
    masks[5, 7] = 1                                     # a single residue: valid by the reference's rules
    with bk.ctx():
        status, stats = validate_geometry_batch(bk.t32(ca), bk.t32(masks))
        one = validate_protein_geometry(bk.t32(ca[0]), bk.t32(masks[0]))
        none = validate_protein_geometry(bk.t32(ca[0]), bk.t32(np.zeros(L, np.float32)))
        shared, _ = validate_geometry_batch(bk.t32(ca), bk.t32(masks[0]))
    assert one == (True, "Valid geometry") and none == (False, "No valid residues")
    want = [generation_oracle.validate_protein_geometry(ca[s], masks[s]) for s in range(S)]
    assert status.cpu().tolist() == [w[0] for w in want]
    assert {0, 2, 3, 4} <= set(status.cpu().tolist())               # every branch is exercised
    assert np.allclose(stats.cpu().numpy(), np.array([w[1:] for w in want]), rtol=1e-4, atol=1e-4)
    assert shared.cpu().tolist() == [generation_oracle.validate_protein_geometry(ca[s], masks[0])[0] for s in range(S)]
    if os.path.exists(REF):                                          # the reference function itself, where it exists
        src = open(REF).read()
        fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "validate_protein_geometry")
        ns = {"torch": torch, "np": np}
        exec(compile(ast.Module([fn], []), REF, "exec"), ns)
        for s in range(S):
            ok, reason = ns["validate_protein_geometry"](torch.tensor(ca[s]), torch.tensor(masks[s]))
            with bk.ctx():
                mine = validate_protein_geometry(bk.t32(ca[s]), bk.t32(masks[s]))
            assert mine == (ok, reason), (s, mine, ok, reason)


def test_generate_ensemble_driver(bk):
    from protein_ensemble_vae_b200 import EGNNDecoder, generate_ensemble, kabsch_rmsd_batch
    torch.manual_seed(0)
    dec = EGNNDecoder(8, 4, hidden_dim=32, num_layers=1, max_neighbors=3, dropout=0.0, precision="fp32").to(bk.dev).eval()
    S, L = 7, 12
    zg, zl = torch.randn(S, 8, device=bk.dev), torch.randn(S, L, 4, device=bk.dev)
    mask = torch.ones(L, device=bk.dev)
    mask[10:] = 0
    ref = torch.cumsum(torch.randn(L, 3, device=bk.dev) * 2.2, 0)
    with bk.ctx():
        res = generate_ensemble(dec, zg, zl, mask, ref, chunk=3)
        with torch.no_grad():
            ca = dec(zg, zl, mask.unsqueeze(0).expand(S, -1))[1]
            want = kabsch_rmsd_batch(ca, ref, mask)
    assert res["CA"].shape == (S, L, 3) and torch.allclose(res["CA"], ca, atol=1e-6)
    assert torch.allclose(res["rmsd"], want, atol=1e-5) and res["valid"].dtype == torch.bool
    assert res["status"].shape == (S,) and res["stats"].shape == (S, 3) and float(res["diversity"]) >= 0.0
