"""v2 tcgen05 edge kernels (csrc/edge_tc2_kernels.cu) on a real B200: every kernel against a torch emulation with the
same rounding points (bf16 GEMM operands, bf16 stored streams, fp16 staged node projection, fp32 accumulation).
The layouts these tests pin: transposed / MN-major UMMA descriptors, the tile-image format, the TMA bulk and tensor
stores (including clipping of the last partial tile), segment handling on banded and short-segment graphs."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
H = 256
BF = torch.bfloat16


def bf(t):
    return t.to(BF).float()


def one_plus_r(h):          # d silu(2h) / dh
    t = torch.tanh(h)
    return 1 + t + h * (1 - t * t)


def silu2(h):               # silu(2h) in the half domain
    return h + h * torch.tanh(h)


def _setup(lengths, W, seed):
    from protein_ensemble_vae_b200.graph import band_graph
    g = band_graph(lengths, W, "cuda", cache=False)
    gen = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s, k=1.0: torch.randn(*s, device="cuda", generator=gen) * k  # noqa: E731
    N = g.num_nodes
    return dict(g=g, N=N, E=g.num_edges, x=r(N, 3, k=2.0), ABh=r(N, 2 * H, k=0.5).to(torch.float16), wd=r(H, k=0.02),
                W2=r(H, H, k=1 / 16), W5=r(H, H, k=1 / 16), b2=r(H, k=0.1), b5=r(H, k=0.1), w6=r(H, k=0.1), b6=r(1))


CASES = [((100,), 40), ((7, 130, 64), 40), ((256,) * 3, 40), ((33,), 5), ((3, 2), 1), ((90, 41), 3)]


@pytest.mark.parametrize("lengths,W", CASES)
def test_forward_kernels_match_emulation(lengths, W):
    from protein_ensemble_vae_b200 import _lib, egnn_tc2 as T2
    from protein_ensemble_vae_b200._lib import ptr, stream
    c = _setup(lengths, W, 5)
    g, N, E = c["g"], c["N"], c["E"]
    row, col = g.row.long(), g.col.long()
    L, st = _lib.lib(), stream(c["x"])
    d2 = torch.empty(E, device="cuda")
    L.call("pev_edge_d2", ptr(c["x"]), ptr(g.row), ptr(g.col), E, ptr(d2), st)
    d2_ref = ((c["x"][row] - c["x"][col]) ** 2).sum(-1)
    assert rel_err(d2, d2_ref) < 1e-6
    hvT, mT = T2.alloc_tile_image(E, "cuda"), T2.alloc_tile_image(E, "cuda")
    agg = torch.full((N, H), 7.0, device="cuda")            # zeroed inside
    w = torch.full((E,), 7.0, device="cuda")
    hs = torch.empty(E, H, dtype=BF, device="cuda")
    L.call("pev_edge2_fwd1", ptr(c["ABh"]), ptr(d2), ptr(c["wd"]), ptr(T2.packed_weight_scaled(c["W2"], 0.5)), ptr(c["b2"]),
           ptr(g.row), ptr(g.col), N, E, ptr(hvT), ptr(mT), ptr(agg), st)
    L.call("pev_edge2_fwd2", ptr(mT), ptr(T2.packed_weight_scaled(c["W5"], 0.5)), ptr(c["b5"]), ptr(c["w6"]), ptr(c["b6"]),
           E, ptr(w), ptr(hs), st)
    hu = (c["ABh"][row, :H] + c["ABh"][col, H:]).float() + 0.5 * c["wd"] * d2_ref[:, None]
    hv = bf(silu2(hu)) @ bf(0.5 * c["W2"]).t() + 0.5 * c["b2"]
    agg_ref = torch.zeros(N, H, device="cuda").index_add_(0, row, silu2(hv))
    hv_k = T2.tile_image_to_rows(hvT, E).float()
    assert rel_err(hv_k, hv) < 6e-3
    assert rel_err(agg, agg_ref) < 1e-3
    m_k = T2.tile_image_to_rows(mT, E).float()
    assert rel_err(m_k, silu2(hv)) < 6e-3
    hs_ref = m_k @ bf(0.5 * c["W5"]).t() + 0.5 * c["b5"]                       # from the kernel's own stored m
    assert rel_err(hs.float(), hs_ref) < 6e-3
    assert rel_err(w, silu2(hs_ref) @ c["w6"] + c["b6"]) < 6e-3
    agg2 = torch.full((N, H), 7.0, device="cuda")          # inference form: neither hv nor hs written
    mT2 = T2.alloc_tile_image(E, "cuda")
    L.call("pev_edge2_fwd1", ptr(c["ABh"]), ptr(d2), ptr(c["wd"]), ptr(T2.packed_weight_scaled(c["W2"], 0.5)), ptr(c["b2"]),
           ptr(g.row), ptr(g.col), N, E, None, ptr(mT2), ptr(agg2), st)
    assert torch.equal(T2.tile_image_to_rows(mT2, E), T2.tile_image_to_rows(mT, E)) and rel_err(agg2, agg) < 1e-5
    w2 = torch.full((E,), 7.0, device="cuda")
    L.call("pev_edge2_fwd2", ptr(mT), ptr(T2.packed_weight_scaled(c["W5"], 0.5)), ptr(c["b5"]), ptr(c["w6"]), ptr(c["b6"]),
           E, ptr(w2), None, st)
    assert rel_err(w2, w) < 1e-5


@pytest.mark.parametrize("lengths,W", CASES)
def test_backward_kernels_match_emulation(lengths, W):
    from protein_ensemble_vae_b200 import _lib, egnn_tc2 as T2
    from protein_ensemble_vae_b200._lib import ptr, stream
    c = _setup(lengths, W, 9)
    g, N, E = c["g"], c["N"], c["E"]
    row, col = g.row.long(), g.col.long()
    gen = torch.Generator(device="cuda").manual_seed(1)
    hs = (torch.randn(E, H, device="cuda", generator=gen) * 0.7).to(BF)
    hv = (torch.randn(E, H, device="cuda", generator=gen) * 0.7).to(BF)
    gw = torch.randn(E, device="cuda", generator=gen)
    gagg = torch.randn(N, H, device="cuda", generator=gen) * 0.3
    L, st = _lib.lib(), stream(c["x"])
    d2 = ((c["x"][row] - c["x"][col]) ** 2).sum(-1).contiguous()
    hvT = T2.rows_to_tile_image(hv)
    mT = T2.rows_to_tile_image(silu2(hv.float()).to(BF))
    ghvT = T2.alloc_tile_image(E, "cuda")
    db2h = torch.empty(H, device="cuda")
    ghu = torch.empty(E, H, dtype=BF, device="cuda")
    gd2 = torch.full((E,), 7.0, device="cuda")
    ws = torch.empty(L.cdll.pev_edge2_wgrad_workspace_bytes() // 4, device="cuda")
    dW5, dW2 = torch.empty(H, H, device="cuda"), torch.empty(H, H, device="cuda")
    db5h, dw6 = torch.empty(H, device="cuda"), torch.empty(H, device="cuda")
    P = lambda Wt: T2.packed_weight_scaled(Wt, 0.5, transpose=True)  # noqa: E731
    L.call("pev_edge2_bwd2", ptr(hs), ptr(gw), ptr(c["w6"]), ptr(P(c["W5"])), ptr(gagg), ptr(g.row), ptr(hvT), E, ptr(ws), ptr(ghvT),
           ptr(db2h), st)
    L.call("pev_edge2_bwd1", ptr(ghvT), ptr(P(c["W2"])), ptr(c["ABh"]), ptr(d2), ptr(g.row), ptr(g.col), ptr(c["wd"]), E,
           ptr(ghu), ptr(gd2), ptr(torch.empty(4 * E, device="cuda")), st)
    L.call("pev_edge2_wgrad5", ptr(hs), ptr(gw), ptr(c["w6"]), ptr(mT), E, ptr(ws), ptr(dW5), ptr(db5h), ptr(dw6), st)
    L.call("pev_edge2_wgrad2", ptr(ghvT), ptr(c["ABh"]), ptr(d2), ptr(g.row), ptr(g.col), ptr(c["wd"]), E, ptr(ws), ptr(dW2), st)
    ghs = gw[:, None] * c["w6"] * one_plus_r(hs.float())
    ghv = (bf(ghs) @ bf(0.5 * c["W5"]) + gagg[row]) * one_plus_r(hv.float())
    ghv_k = T2.tile_image_to_rows(ghvT, E).float()
    assert rel_err(ghv_k, ghv) < 6e-3
    assert rel_err(db2h, ghv.sum(0)) < 1e-3
    hu = (c["ABh"][row, :H] + c["ABh"][col, H:]).float() + 0.5 * c["wd"] * d2[:, None]
    ghu_ref = (ghv_k @ bf(0.5 * c["W2"])) * one_plus_r(hu)                      # from the kernel's own stored ghv
    assert rel_err(ghu.float(), ghu_ref) < 6e-3
    assert rel_err(gd2, ghu_ref @ (0.5 * c["wd"])) < 1e-3
    m, t, a = silu2(hv.float()), silu2(hs.float()), silu2(hu)
    assert rel_err(dW5, 0.5 * bf(ghs).t().double() @ bf(m).double()) < 2e-3
    assert rel_err(db5h, ghs.double().sum(0)) < 1e-3
    assert rel_err(dw6, (gw[:, None].double() * t.double()).sum(0)) < 1e-3
    assert rel_err(dW2, 0.5 * ghv_k.t().double() @ bf(a).double()) < 2e-3


def test_packed_weight_image_scaled():
    from protein_ensemble_vae_b200.egnn_tc2 import packed_weight_scaled
    W = torch.arange(H * H, dtype=torch.float32, device="cuda").reshape(H, H) / 4096.0
    img = packed_weight_scaled(W, 0.5).cpu().float().numpy()
    imgT = packed_weight_scaled(W, 0.5, transpose=True).cpu().float().numpy()
    Wb = (0.5 * W).to(BF).float().cpu().numpy()
    for n, k in ((0, 0), (1, 8), (9, 63), (200, 64), (255, 255), (77, 130)):
        kb, kl = divmod(k, 64)
        idx = (kb * 32768 + (n // 8) * 1024 + (n % 8) * 128 + (((kl // 8) ^ (n % 8)) << 4) + (kl % 8) * 2) // 2
        assert img[idx] == Wb[n, k]
        assert imgT[idx] == Wb[k, n]
