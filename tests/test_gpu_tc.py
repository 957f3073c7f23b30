"""tcgen05 edge-MLP kernels (K1 bf16 form) on a real B200: layout/descriptor validation against a torch
emulation with the same rounding points, then the whole bf16 decoder against the float64 oracle."""
import os

import numpy as np
import pytest
import torch

import cases
import synth
from conftest import assert_named_close_l2, rel_err
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
    v = torch.empty(E, H, dtype=bf, device="cuda")
    s = torch.empty(E, H, dtype=bf, device="cuda")
    a_k = torch.empty(E, H, dtype=bf, device="cuda")
    m_k = torch.empty(E, H, dtype=bf, device="cuda")
    da_k = torch.empty(E, H, dtype=bf, device="cuda")
    dm_k = torch.empty(E, H, dtype=bf, device="cuda")
    agg = torch.full((N, H), 7.0, device="cuda")        # must be zeroed inside
    w = torch.full((E,), 7.0, device="cuda")
    L = _lib.lib()
    L.call("pev_edge_mlp1_fwd_bf16", ptr(AB), ptr(x), ptr(wd), ptr(packed_weight(W2)), ptr(b2), ptr(g.row), ptr(g.col),
           N, E, ptr(v), ptr(a_k), ptr(da_k), ptr(agg), stream(x))
    L.call("pev_edge_mlp2_fwd_bf16", ptr(v), ptr(packed_weight(W5)), ptr(b5), ptr(w6), ptr(b6), E, ptr(w), ptr(s),
           ptr(m_k), ptr(dm_k), stream(x))
    torch.cuda.synchronize()
    # emulation with the kernel's rounding points (bf16 operands, fp32 accumulation)
    row, col = g.row.long(), g.col.long()
    rel = x[row] - x[col]
    d2 = (rel * rel).sum(-1, keepdim=True)
    u = AB[row, :H] + AB[col, H:] + wd[None] * d2
    a = _silu(u).to(bf).float()
    assert rel_err(a_k.float(), a) < 6e-3
    v_ref = a @ W2.to(bf).float().t() + b2
    assert rel_err(v.float(), v_ref) < 6e-3                      # bf16 output rounding + tanh.approx
    m = _silu(v_ref)
    agg_ref = torch.zeros(N, H, device="cuda").index_add_(0, row, m)
    assert rel_err(agg, agg_ref) < 6e-3
    m2 = _silu(v.float()).to(bf).float()                         # stage 2 starts from the stored bf16 v
    assert rel_err(m_k.float(), m2) < 6e-3
    s_ref = m2 @ W5.to(bf).float().t() + b5
    assert rel_err(s.float(), s_ref) < 6e-3
    w_ref = _silu(s_ref) @ w6 + b6
    assert rel_err(w, w_ref) < 6e-3
    # ---- backward kernels (stage 3 / stage 4) against the same kind of emulation
    rng = np.random.default_rng(11)
    t = lambda arr: torch.tensor(np.asarray(arr, dtype=np.float32), device="cuda")  # noqa: E731
    gw = t(rng.standard_normal(E))
    gagg = t(rng.standard_normal((N, H)))
    gs_k = torch.empty(E, H, dtype=bf, device="cuda")
    gv_k = torch.empty(E, H, dtype=bf, device="cuda")
    gu_k = torch.empty(E, H, dtype=bf, device="cuda")
    db5, dw6, db2 = (torch.full((H,), 3.0, device="cuda") for _ in range(3))
    gd2 = torch.full((E,), 3.0, device="cuda")
    L.call("pev_edge_mlp2_bwd_bf16", ptr(s), ptr(dm_k), ptr(gw), ptr(w6), ptr(packed_weight(W5, transpose=True)), ptr(gagg),
           ptr(g.row), E, ptr(gs_k), ptr(gv_k), ptr(db5), ptr(dw6), stream(x))
    L.call("pev_edge_mlp1_bwd_bf16", ptr(gv_k), ptr(da_k), ptr(packed_weight(W2, transpose=True)), ptr(wd), E,
           ptr(gu_k), ptr(gd2), ptr(db2), stream(x))
    gAB = torch.empty(N, 2 * H, device="cuda")
    part = torch.empty(N, H, device="cuda")
    gx = torch.zeros(N, 3, device="cuda")
    L.call("pev_edge_prologue_bwd_bf16", ptr(gu_k), ptr(gd2), ptr(x), ptr(g.row_ptr), ptr(g.row), ptr(g.col),
           ptr(g.col_ptr), ptr(g.csc_perm), N, E, ptr(gAB), ptr(gx), ptr(part), stream(x))
    torch.cuda.synchronize()

    def dsilu(z):
        sg = torch.sigmoid(z)
        return sg * (1 + z * (1 - sg))
    sf, vf = s.float(), v.float()
    assert rel_err(da_k.float(), dsilu(u)) < 6e-3 and rel_err(dm_k.float(), dsilu(vf)) < 6e-3
    gs_ref = gw[:, None] * w6[None] * dsilu(sf)
    assert rel_err(gs_k.float(), gs_ref) < 6e-3
    assert rel_err(db5, gs_ref.sum(0)) < 6e-3
    assert rel_err(dw6, gw @ _silu(sf)) < 6e-3
    gv_ref = (gs_k.float() @ W5.to(bf).float() + gagg[row]) * dm_k.float()
    assert rel_err(gv_k.float(), gv_ref) < 6e-3
    assert rel_err(db2, gv_k.float().sum(0)) < 6e-3
    gu_ref = (gv_k.float() @ W2.to(bf).float()) * da_k.float()
    assert rel_err(gu_k.float(), gu_ref) < 6e-3
    assert rel_err(gd2, gu_ref @ wd) < 1e-2
    guf = gu_k.float()
    gA_ref = torch.zeros(N, H, device="cuda").index_add_(0, row, guf)
    gB_ref = torch.zeros(N, H, device="cuda").index_add_(0, col, guf)
    assert rel_err(gAB[:, :H], gA_ref) < 1e-5 and rel_err(gAB[:, H:], gB_ref) < 1e-5
    assert rel_err(part.sum(0), (guf * d2).sum(0)) < 1e-4
    grel = (2.0 * gd2)[:, None] * rel
    gx_ref = torch.zeros(N, 3, device="cuda").index_add_(0, row, grel).index_add_(0, col, -grel)
    assert rel_err(gx, gx_ref) < 1e-4


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
    # 1e-1: at random init the gradients are ~10x more sensitive than the outputs -- sequence_head.0.weight,
    # whose backward is plain fp32 torch in both paths, already differs by 4.5% (relative L2) when its input h
    # carries the bf16 path's 0.5% forward error (tools/grad_diag.py).  The backward kernels themselves are
    # held to 6e-3 against a same-rounding emulation in test_edge_mlp_kernels_match_bf16_emulation.
    assert_named_close_l2(grads, ref, tol=1e-1)


def test_no_grad_decode_keeps_nothing():
    from protein_ensemble_vae_b200 import EGNNDecoder
    dec = EGNNDecoder(32, 16, hidden_dim=256, num_layers=2, max_neighbors=40, dropout=0.0, precision="bf16").cuda().eval()
    with torch.no_grad():
        n, ca, c, lg = dec(torch.randn(5, 32, device="cuda"), torch.randn(5, 70, 16, device="cuda"))
    assert ca.shape == (5, 70, 3) and lg.shape == (5, 70, 20) and torch.isfinite(ca).all()
    # N-CA and CA-C bonds are fixed-length offsets (models/en_gnn_decoder.py:289-293); C is never pulled
    assert torch.allclose((c - ca).norm(dim=-1), torch.full((5, 70), 1.52, device="cuda"), atol=1e-4)



def test_add_layernorm_and_column_sum_match_torch():
    """Node-level kernels of the bf16 path (csrc/node_kernels.cu) against torch fp32 (1e-5)."""
    from protein_ensemble_vae_b200.egnn_tc import AddLayerNorm, column_sum
    torch.manual_seed(3)
    for D, N in ((256, 1000), (512, 333)):
        x = torch.randn(N, D, device="cuda", requires_grad=True)
        res = torch.randn(N, D, device="cuda", requires_grad=True)
        ln = torch.nn.LayerNorm(D).cuda()
        with torch.no_grad():
            ln.weight.uniform_(0.5, 1.5)
            ln.bias.uniform_(-0.5, 0.5)
        coef = torch.randn(N, D, device="cuda")
        y = AddLayerNorm.apply(x, res, ln.weight, ln.bias, ln.eps)
        (y * coef).sum().backward()
        got = (y.detach(), x.grad.clone(), res.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone())
        x.grad = res.grad = ln.weight.grad = ln.bias.grad = None
        y2 = ln(x + res)
        (y2 * coef).sum().backward()
        ref = (y2.detach(), x.grad, res.grad, ln.weight.grad, ln.bias.grad)
        for a, b in zip(got, ref):
            assert rel_err(a, b) < 1e-5
        y3 = AddLayerNorm.apply(x.detach(), None, ln.weight, ln.bias, ln.eps)
        assert rel_err(y3, ln(x.detach())) < 1e-5
        assert rel_err(column_sum(coef), coef.sum(0)) < 1e-5


def test_edge2_sums_match_index_add():
    """v2 segment-sum kernel (one warp per node) against torch index_add_ on a banded and a generic graph."""
    from protein_ensemble_vae_b200 import _lib
    from protein_ensemble_vae_b200._lib import ptr, stream
    from protein_ensemble_vae_b200.graph import band_graph, graph_from_edge_index
    torch.manual_seed(11)
    ei = torch.randint(0, 50, (2, 900), device="cuda")
    for g in (band_graph((70, 33, 1, 100), 40, "cuda"), graph_from_edge_index(ei, 50)):
        N, E = g.num_nodes, g.num_edges
        ghu = (torch.randn(E, H, device="cuda") * 0.3).to(torch.bfloat16)
        d2 = torch.rand(E, device="cuda") * 9
        gAB = torch.empty(N, 2 * H, device="cuda")
        gwdh = torch.empty(H, device="cuda")
        _lib.lib().call("pev_edge2_sums", ptr(ghu), ptr(d2), ptr(g.row_ptr), ptr(g.col_ptr), ptr(g.csc_perm), N, E,
                        ptr(gAB), ptr(gwdh), stream(ghu))
        gf = ghu.float()
        gA = torch.zeros(N, H, device="cuda").index_add_(0, g.row.long(), gf)
        gB = torch.zeros(N, H, device="cuda").index_add_(0, g.col.long(), gf)
        assert rel_err(gAB[:, :H], gA) < 1e-5 and rel_err(gAB[:, H:], gB) < 1e-5
        assert rel_err(gwdh, (gf * d2[:, None]).sum(0)) < 1e-4
