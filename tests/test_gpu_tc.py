"""tcgen05 edge-MLP kernels (K1 bf16 form) on a real B200: layout/descriptor validation against a torch
emulation with the same rounding points, then the whole bf16 decoder against the float64 oracle."""
import os

import numpy as np
import pytest
import torch

import cases
import synth
from conftest import assert_named_close, rel_err
from oracle import egnn_oracle, graph_oracle

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
H = 256


def _silu(z):
    return z * torch.sigmoid(z)


def _setup(lengths, W, seed):
    from protein_ensemble_vae_b200.graph import band_graph
    rng = np.random.default_rng(seed)
    dev = "cuda"
    g = band_graph(lengths, W, dev, cache=False)
    N = g.num_nodes
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device=dev)  # noqa: E731
    AB = t(rng.standard_normal((N, 2 * H)))
    x = t(rng.standard_normal((N, 3)) * 1.5)
    wd = t(rng.standard_normal(H) * 0.05)
    W2 = t(rng.standard_normal((H, H)) / 16)
    b2 = t(rng.standard_normal(H) * 0.1)
    W5 = t(rng.standard_normal((H, H)) / 16)
    b5 = t(rng.standard_normal(H) * 0.1)
    w6 = t(rng.standard_normal(H) / 16)
    b6 = t([0.3])
    return g, AB, x, wd, W2, b2, W5, b5, w6, b6


@pytest.mark.parametrize("lengths,W", [((100,), 40), ((7, 130, 64), 40), ((256,) * 3, 40), ((33,), 5), ((3, 2), 1)])
def test_edge_mlp_kernels_match_bf16_emulation(lengths, W):
    from protein_ensemble_vae_b200 import _lib
    from protein_ensemble_vae_b200._lib import ptr, stream
    from protein_ensemble_vae_b200.egnn_tc import packed_weight
    g, AB, x, wd, W2, b2, W5, b5, w6, b6 = _setup(lengths, W, 5)
    N, E = g.num_nodes, g.num_edges
    bf = torch.bfloat16
    ABh = AB.to(bf).contiguous()
    v = torch.empty(E, H, dtype=bf, device="cuda")
    s = torch.empty(E, H, dtype=bf, device="cuda")
    agg = torch.full((N, H), 7.0, device="cuda")        # must be zeroed inside
    w = torch.full((E,), 7.0, device="cuda")
    L = _lib.lib()
    L.call("pev_edge_mlp1_fwd_bf16", ptr(ABh), ptr(x), ptr(wd), ptr(packed_weight(W2)), ptr(b2), ptr(g.row), ptr(g.col),
           N, E, ptr(v), ptr(agg), stream(x))
    L.call("pev_edge_mlp2_fwd_bf16", ptr(v), ptr(packed_weight(W5)), ptr(b5), ptr(w6), ptr(b6), E, ptr(w), ptr(s), stream(x))
    torch.cuda.synchronize()
    # emulation with the kernel's rounding points (bf16 operands, fp32 accumulation)
    row, col = g.row.long(), g.col.long()
    rel = x[row] - x[col]
    d2 = (rel * rel).sum(-1, keepdim=True)
    u = ABh[row, :H].float() + ABh[col, H:].float() + wd[None] * d2
    a = _silu(u).to(bf).float()
    v_ref = a @ W2.to(bf).float().t() + b2
    assert rel_err(v.float(), v_ref) < 6e-3                      # bf16 output rounding + tanh.approx
    m = _silu(v_ref)
    agg_ref = torch.zeros(N, H, device="cuda").index_add_(0, row, m)
    assert rel_err(agg, agg_ref) < 6e-3
    m2 = _silu(v.float()).to(bf).float()                         # stage 2 starts from the stored bf16 v
    s_ref = m2 @ W5.to(bf).float().t() + b5
    assert rel_err(s.float(), s_ref) < 6e-3
    w_ref = _silu(s_ref) @ w6 + b6
    assert rel_err(w, w_ref) < 6e-3


def test_packed_weight_image():
    from protein_ensemble_vae_b200.egnn_tc import packed_weight
    W = torch.arange(H * H, dtype=torch.float32, device="cuda").reshape(H, H) / 4096.0
    img = packed_weight(W).cpu().float().numpy()
    imgT = packed_weight(W, transpose=True).cpu().float().numpy()
    Wb = W.to(torch.bfloat16).float().cpu().numpy()
    for n, k in ((0, 0), (1, 8), (9, 63), (200, 64), (255, 255), (77, 130)):
        kb, kl = divmod(k, 64)
        idx = (kb * 32768 + (n // 8) * 1024 + (n % 8) * 128 + (((kl // 8) ^ (n % 8)) << 4) + (kl % 8) * 2) // 2
        assert img[idx] == Wb[n, k]
        assert imgT[idx] == Wb[k, n]


@pytest.mark.parametrize("tag", ["h256_gaps", "refdims"])
def test_decoder_bf16_vs_oracle(tag):
    from protein_ensemble_vae_b200 import EGNNDecoder
    gold = np.load(os.path.join(G, "decoders.npz"))
    case = cases.DECODER_CASES[tag]
    z_g, z_l, Hd, nl, W, B, L, mkind, pseed, dseed = case
    dec = EGNNDecoder(z_g, z_l, hidden_dim=Hd, num_layers=nl, max_neighbors=W, dropout=0.0, precision="bf16").cuda().eval()
    params = synth.make_params(synth.decoder_param_shapes(z_g, z_l, Hd, nl), pseed)
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device="cuda")  # noqa: E731
    dec.load_state_dict({k: t(v) for k, v in params.items()})
    zg, zl, mask, coef = cases.decoder_inputs(case)
    zg_t, zl_t = t(zg).requires_grad_(), t(zl).requires_grad_()
    outs = dec(zg_t, zl_t, None if mask is None else t(mask))
    for name, o in zip(("N", "CA", "C", "logits"), outs):
        assert rel_err(o.detach(), gold[f"{tag}.{name}"]) < 1e-2, name          # north_star: bf16 edge MLP 1e-2
        if mask is not None and (mask == 0).any():
            assert float(o.detach()[t(mask) == 0].abs().max()) == 0.0
    sum((o * t(c)).sum() for o, c in zip(outs, coef)).backward()
    T64 = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))  # noqa: E731
    sd = {k: T64(v).requires_grad_() for k, v in params.items()}
    zg64, zl64 = T64(zg).requires_grad_(), T64(zl).requires_grad_()
    o64 = egnn_oracle.egnn_decoder(sd, zg64, zl64, None if mask is None else T64(mask), max_neighbors=W)
    sum((o * T64(c)).sum() for o, c in zip(o64, coef)).backward()
    ref = {"z_g": zg64.grad, "z_l": zl64.grad}
    ref.update({k: v.grad for k, v in sd.items() if v.grad is not None})
    grads = {"z_g": zg_t.grad, "z_l": zl_t.grad}
    grads.update({k: p.grad for k, p in dec.named_parameters() if p.grad is not None})
    assert_named_close(grads, ref, tol=5e-2, max_outliers=8, outlier_tol=0.5)


def test_no_grad_decode_keeps_nothing():
    from protein_ensemble_vae_b200 import EGNNDecoder
    dec = EGNNDecoder(32, 16, hidden_dim=256, num_layers=2, max_neighbors=40, dropout=0.0, precision="bf16").cuda().eval()
    with torch.no_grad():
        n, ca, c, lg = dec(torch.randn(5, 32, device="cuda"), torch.randn(5, 70, 16, device="cuda"))
    assert ca.shape == (5, 70, 3) and lg.shape == (5, 70, 20) and torch.isfinite(ca).all()
    # N-CA and CA-C bonds are fixed-length offsets (models/en_gnn_decoder.py:289-293); C is never pulled
    assert torch.allclose((c - ca).norm(dim=-1), torch.full((5, 70), 1.52, device="cuda"), atol=1e-4)
