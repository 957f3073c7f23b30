"""Fused CTA-pair forward kernel (csrc/edge_tc3_kernels.cu, ``pev_edge3_fwd``; experimental -- see
profiles/r02_fused_pair_fwd_ablation.md) against a torch emulation with the same rounding points, through the C ABI."""
import pytest
import torch

from conftest import rel_err
from test_gpu_tc2 import BF, H, _setup, bf, silu2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lengths,W", [((100,), 40), ((7, 130, 64), 40), ((256,) * 3, 40), ((33,), 5), ((3, 2), 1), ((90, 41), 3)])
def test_fused_pair_forward_matches_emulation(lengths, W):
    from protein_ensemble_vae_b200 import _lib, egnn_tc2 as T2
    from protein_ensemble_vae_b200._lib import ptr, stream
    c = _setup(lengths, W, 5)
    g, N, E = c["g"], c["N"], c["E"]
    row, col = g.row.long(), g.col.long()
    L, st = _lib.lib(), stream(c["x"])
    d2 = ((c["x"][row] - c["x"][col]) ** 2).sum(-1).contiguous()
    W2p, W5p = T2.packed_weight_scaled(c["W2"], 0.5), T2.packed_weight_scaled(c["W5"], 0.5)     # both stay alive
    outs = []
    for train in (True, False):
        agg = torch.full((N, H), 7.0, device="cuda")                                            # zeroed inside
        w = torch.full((E,), 7.0, device="cuda")
        hv = torch.zeros(E, H, dtype=BF, device="cuda") if train else None
        hs = torch.zeros(E, H, dtype=BF, device="cuda") if train else None
        L.call("pev_edge3_fwd", ptr(c["ABh"]), ptr(d2), ptr(c["wd"]), ptr(W2p), ptr(c["b2"]), ptr(W5p), ptr(c["b5"]),
               ptr(c["w6"]), ptr(c["b6"]), ptr(g.row), ptr(g.col), N, E, ptr(hv), ptr(hs), ptr(agg), ptr(w), st)
        outs.append((agg, w, hv, hs))
    (agg, w, hv, hs), (agg2, w2, _, _) = outs
    hu = (c["ABh"][row, :H] + c["ABh"][col, H:]).float() + 0.5 * c["wd"] * d2[:, None]
    hv_ref = bf(silu2(hu)) @ bf(0.5 * c["W2"]).t() + 0.5 * c["b2"]
    m_ref = silu2(hv_ref)
    hs_ref = bf(m_ref) @ bf(0.5 * c["W5"]).t() + 0.5 * c["b5"]
    assert rel_err(hv.float(), hv_ref) < 6e-3 and rel_err(hs.float(), hs_ref) < 8e-3
    assert rel_err(agg, torch.zeros(N, H, device="cuda").index_add_(0, row, m_ref)) < 4e-3      # sums the bf16 operand
    assert rel_err(w, silu2(hs_ref) @ c["w6"] + c["b6"]) < 8e-3
    assert rel_err(agg2, agg) < 1e-5 and rel_err(w2, w) < 1e-5
