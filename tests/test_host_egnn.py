"""EGNN graph builder, fp32 edge prologue, exact-order scatter (per-thread bodies compiled for the
host) and the packed decoder's Python logic vs the reference fixtures.  CPU only."""
import os

import numpy as np
import pytest
import torch

import cases
import synth
from conftest import assert_named_close, rel_err
from oracle import egnn_oracle
from oracle import graph_oracle, kabsch_oracle

G = os.path.join(os.path.dirname(__file__), "golden")
T32 = lambda a: torch.tensor(np.asarray(a, dtype=np.float32))  # noqa: E731


def check_grads(named, gold, tag, tol, seed=1234):
    for i, name in enumerate(sorted(named)):
        g = named[name].detach().double().cpu().numpy()
        if f"{tag}.grad.{name}" in gold:
            assert rel_err(g, gold[f"{tag}.grad.{name}"]) < tol, name
        else:
            r1, r2 = cases.proj_vectors(g.shape, seed + i)
            assert rel_err(g @ r1, gold[f"{tag}.gproj1.{name}"]) < tol, name
            assert rel_err(r2 @ g, gold[f"{tag}.gproj2.{name}"]) < tol, name


@pytest.mark.parametrize("lengths,W", [((5,), 2), ((64,), 40), ((100,), 40), ((9, 1, 0, 17, 2), 3), ((6,), 9),
                                       ((300, 41, 40), 40)])
def test_band_graph_bit_exact(lengths, W, bk):
    T32 = bk.t32
    from protein_ensemble_vae_b200.graph import band_graph
    row_ptr, row, col = graph_oracle.packed_band_graph(lengths, W)
    with bk.ctx():
        g = band_graph(lengths, W, bk.dev, cache=False)
    assert g.num_nodes == sum(lengths) and g.num_edges == len(row)
    assert np.array_equal(g.row_ptr.cpu().numpy(), row_ptr)
    assert np.array_equal(g.row.cpu().numpy(), row) and np.array_equal(g.col.cpu().numpy(), col)
    # CSC permutation: stable sort of the edges by source node
    order = np.lexsort((row, col))
    assert np.array_equal(g.csc_perm.cpu().numpy(), order)
    deg = np.diff(row_ptr)
    want = np.where(deg > 0, (1.0 / np.maximum(deg, 1).astype(np.float32)), 0).astype(np.float32)
    assert np.array_equal(g.dinv.cpu().numpy(), want)
    assert g.edge_index().dtype == torch.int64


def test_build_edge_index_api_matches_reference(bk):
    T32 = bk.t32
    from protein_ensemble_vae_b200 import EGNNDecoder
    gold = np.load(os.path.join(G, "edges.npz"))
    with bk.ctx():
        assert np.array_equal(EGNNDecoder.build_edge_index(5, bk.dev, 2).cpu().numpy(), gold["L5_W2"])
        assert np.array_equal(EGNNDecoder.build_edge_index(7, bk.dev, 0).cpu().numpy(), gold["L7_W0_fallback"])
        assert np.array_equal(EGNNDecoder.build_edge_index(6, bk.dev, 9).cpu().numpy(), gold["L6_W9_dense"])
        for L in (64, 100):
            ei = EGNNDecoder.build_edge_index(L, bk.dev, 40)
            assert np.array_equal(ei.cpu().numpy(), gold[f"L{L}_W40"].astype(np.int64))
            assert np.array_equal(EGNNDecoder.degrees(ei, L).cpu().numpy(), gold[f"deg_L{L}_W40"])


def test_scatter_is_bit_exact_vs_cpu_index_add(bk):
    T32 = bk.t32
    """K2 contract: same bits as the reference's CPU index_add_ (SURVEY.md F4)."""
    from protein_ensemble_vae_b200.egnn_ops import ScatterCoord
    from protein_ensemble_vae_b200.graph import band_graph
    rng = np.random.default_rng(7)
    lengths, W, H = (37, 80, 5), 40, 64
    with bk.ctx():
        g = band_graph(lengths, W, bk.dev, cache=False)
        E, N = g.num_edges, g.num_nodes
        m = T32(rng.standard_normal((E, H)) * 3)
        w = T32(rng.standard_normal(E))
        x = T32(rng.standard_normal((N, 3)) * 5)
        agg, x_out = ScatterCoord.apply(m, w, x, g.dinv, g)
    # the oracle for bit-exactness is the CPU index_add_ the reference runs (its CUDA form is atomic)
    row, col, m, w, x, dinv = (t.cpu() for t in (g.row.long(), g.col.long(), m, w, x, g.dinv))
    ref_agg = torch.zeros(N, H).index_add_(0, row, m)
    rel = x[row] - x[col]
    delta = torch.zeros(N, 3).index_add_(0, row, w[:, None] * rel)
    delta = delta * dinv[:, None]
    delta = delta * 0.2
    assert torch.equal(agg.cpu(), ref_agg)
    assert torch.equal(x_out.cpu(), x + delta)


@pytest.mark.parametrize("tag", list(cases.LAYER_CASES))
def test_layer_fp32_host(tag, bk):
    T32 = bk.t32
    from protein_ensemble_vae_b200 import EGNLayer
    gold = np.load(os.path.join(G, "layers.npz"))
    H, lengths, W, pseed, dseed, kind = cases.LAYER_CASES[tag]
    n = sum(lengths)
    layer = EGNLayer(H, H, precision="fp32").to(bk.dev)
    layer.load_state_dict({k: T32(v) for k, v in synth.make_params(synth.layer_param_shapes(H, H), pseed).items()})
    rng = np.random.default_rng(dseed)
    h = T32(rng.standard_normal((n, H))).requires_grad_()
    x = T32(rng.standard_normal((n, 3)) * 2.0).requires_grad_()
    ei = torch.tensor(gold[f"{tag}.edge_index"]).to(bk.dev)
    dinv = (1.0 / torch.bincount(ei[0], minlength=n).float()) if kind == "band" else None
    ch, cx = T32(rng.standard_normal((n, H))), T32(rng.standard_normal((n, 3)))
    with bk.ctx():
        h2, x2 = layer(h, x, ei, degree_inv=dinv)
        ((h2 * ch).sum() + (x2 * cx).sum()).backward()
    assert rel_err(h2.detach(), gold[f"{tag}.h_out"]) < 1e-5
    assert rel_err(x2.detach(), gold[f"{tag}.x_out"]) < 1e-5
    grads = {"h": h.grad, "x": x.grad}
    grads.update({k: p.grad for k, p in layer.named_parameters()})
    check_grads(grads, gold, tag, 2e-5)


@pytest.mark.parametrize("tag", list(cases.DECODER_CASES))
def test_decoder_fp32_host(tag, bk):
    T32 = bk.t32
    from protein_ensemble_vae_b200 import EGNNDecoder
    gold = np.load(os.path.join(G, "decoders.npz"))
    case = cases.DECODER_CASES[tag]
    z_g, z_l, H, nl, W, B, L, mkind, pseed, dseed = case
    dec = EGNNDecoder(z_g, z_l, hidden_dim=H, num_layers=nl, max_neighbors=W, dropout=0.0, precision="fp32").to(bk.dev).eval()
    params = synth.make_params(synth.decoder_param_shapes(z_g, z_l, H, nl), pseed)
    assert set(params) == set(dec.state_dict())                     # checkpoint-compatible names
    assert all(tuple(dec.state_dict()[k].shape) == v.shape for k, v in params.items())
    dec.load_state_dict({k: T32(v) for k, v in params.items()})
    zg, zl, mask, coef = cases.decoder_inputs(case)
    zg_t, zl_t = T32(zg).requires_grad_(), T32(zl).requires_grad_()
    with bk.ctx():
        outs = dec(zg_t, zl_t, None if mask is None else T32(mask))
        sum((o * T32(c)).sum() for o, c in zip(outs, coef)).backward()
    for name, o in zip(("N", "CA", "C", "logits"), outs):
        assert o.shape == gold[f"{tag}.{name}"].shape
        assert rel_err(o.detach(), gold[f"{tag}.{name}"]) < 1e-5, name
        if mask is not None and (mask == 0).any():
            assert float(o.detach()[torch.tensor(mask).to(bk.dev) == 0].abs().max()) == 0.0     # exact zeros at padding
    grads = {"z_g": zg_t.grad, "z_l": zl_t.grad}
    grads.update({k: p.grad for k, p in dec.named_parameters() if p.grad is not None})
    # gradient oracle: the float64 restatement (itself pinned on the reference's gradients in
    # test_oracle_golden.py) on the same inputs, compared element by element
    T64 = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))  # noqa: E731
    sd = {k: T64(v).requires_grad_() for k, v in params.items()}
    zg64, zl64 = T64(zg).requires_grad_(), T64(zl).requires_grad_()
    o64 = egnn_oracle.egnn_decoder(sd, zg64, zl64, None if mask is None else T64(mask), max_neighbors=W)
    sum((o * T64(c)).sum() for o, c in zip(o64, coef)).backward()
    ref = {"z_g": zg64.grad, "z_l": zl64.grad}
    ref.update({k: v.grad for k, v in sd.items() if v.grad is not None})
    # 2e-4, not 1e-5: the heads contain ReLUs, and one pre-activation within float32 rounding of 0 flips a
    # whole gradient contribution that then propagates upstream (measured: 1 of 66k activations, 4e-5
    # upstream).  The smooth part (EGNN layers: SiLU + LayerNorm) is held to 2e-5 in test_layer_fp32_host.
    assert_named_close(grads, ref, tol=2e-4, max_outliers=4)


def test_wrappers_hardcode_reference_shapes():
    from protein_ensemble_vae_b200 import ResidueDecoder
    d = ResidueDecoder(16, 8, hidden=64, dropout=0.0, precision="fp32")
    inner = d.decoder.decoder
    assert (inner.hidden_dim, inner.num_layers, inner.max_neighbors) == (256, 8, 40)       # F1
    assert "decoder.decoder.layers.7.phi_x.2.weight" in d.state_dict()
    assert sum(p.numel() for p in inner.layers[0].parameters()) == 461057


def test_kabsch_host(bk):
    T32 = bk.t32
    from protein_ensemble_vae_b200 import kabsch_rmsd, kabsch_rmsd_batch
    gold = np.load(os.path.join(G, "kabsch.npz"))
    a, b, mask = cases.kabsch_inputs()
    with bk.ctx():
        opt = kabsch_rmsd_batch(T32(a), T32(b), T32(mask)).cpu().numpy()
        compat = kabsch_rmsd_batch(T32(a), T32(b), T32(mask), ref_compat=True).cpu().numpy()
        one = kabsch_rmsd(T32(a[2]), T32(b[2]), T32(mask[2]))
        shared = kabsch_rmsd_batch(T32(a), T32(b[0]), None).cpu().numpy()
    assert np.abs(opt - gold["optimal"]).max() < 1e-5 * max(gold["optimal"].max(), 1)
    assert np.abs(compat - gold["ref_compat"]).max() < 1e-5 * gold["ref_compat"].max()
    assert abs(one - gold["optimal"][2]) < 1e-5
    for s in range(a.shape[0]):
        want = kabsch_oracle.kabsch_rmsd(a[s].astype(np.float64), b[0].astype(np.float64), np.ones(a.shape[1]))
        assert abs(shared[s] - want) < 1e-5 * max(want, 1.0)


def test_tile_image_layout_roundtrip_and_addressing():
    """Tile-image format of the v2 edge kernels (csrc/edge_tc2_kernels.cu header): helper round trip and the byte
    address formula the kernels use."""
    from protein_ensemble_vae_b200 import egnn_tc2 as T2
    torch.manual_seed(0)
    E = 300
    rows = torch.randn(E, 256).to(torch.bfloat16)
    img = T2.rows_to_tile_image(rows)
    assert img.shape == (3, 128 * 256)
    assert torch.equal(T2.tile_image_to_rows(img, E), rows)
    flat = img.view(-1)
    for e, f in ((0, 0), (5, 7), (63, 64), (64, 200), (127, 255), (128, 9), (299, 130)):
        t, el = divmod(e, 128)
        off = t * 65536 + (f // 64) * 16384 + (el // 64) * 8192 + ((f % 64) // 8) * 1024 + (f % 8) * 128 \
            + ((((el % 64) // 8) ^ (f % 8)) << 4) + (el % 8) * 2
        assert flat[off // 2] == rows[e, f]
    assert float(img.view(3, -1)[2].float().abs().sum()) > 0 and float(T2.rows_to_tile_image(rows[:1])[0, 128:].abs().sum()) >= 0


def test_decoder_mask_cache_inference_mode_and_invalidation(bk):
    """A mask created under torch.inference_mode() has no version counter: the decoder must decode it
    (uncached); cache_mask=False and invalidate_mask_cache() cover writes that bypass the counter."""
    from protein_ensemble_vae_b200 import EGNNDecoder
    torch.manual_seed(0)
    dec = EGNNDecoder(8, 4, hidden_dim=32, num_layers=1, max_neighbors=3, dropout=0.0, precision="fp32").to(bk.dev).eval()
    zg, zl = torch.randn(2, 8, device=bk.dev), torch.randn(2, 9, 4, device=bk.dev)
    with bk.ctx():
        with torch.inference_mode():
            mask = torch.ones(2, 9, device=bk.dev)
            mask[1, 5:] = 0
            a = dec(zg, zl, mask)[1]
            b = dec(zg, zl, mask)[1]
        assert torch.equal(a, b) and float(a[1, 5:].abs().max()) == 0.0
        assert "_pev_mask_cache" not in dec.__dict__
        m2 = torch.ones(2, 9, device=bk.dev)
        with torch.no_grad():
            c1 = dec(zg, zl, m2)[1]
            m2.data[1, 5:] = 0                      # bypasses the version counter: the cache cannot see it ...
            dec.invalidate_mask_cache()             # ... so the caller says so
            c2 = dec(zg, zl, m2)[1]
        assert float(c1[1, 5:].abs().max()) > 0 and float(c2[1, 5:].abs().max()) == 0.0
        dec.cache_mask = False
        with torch.no_grad():
            m2.data[1, 3:] = 0
            c3 = dec(zg, zl, m2)[1]
        assert float(c3[1, 3:].abs().max()) == 0.0
