"""The oracle restatement vs the reference's own outputs (tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

import synth
import cases as mg_cases
from conftest import rel_err
from oracle import egnn_oracle, graph_oracle, kabsch_oracle, losses_oracle

G = os.path.join(os.path.dirname(__file__), "golden")
T = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))  # noqa: E731


def load(name):
    return np.load(os.path.join(G, name))


def check_grads(named, gold, tag, tol, seed=1234):
    names = sorted(named)
    for i, name in enumerate(names):
        g = named[name].detach().numpy()
        if f"{tag}.grad.{name}" in gold:
            assert rel_err(g, gold[f"{tag}.grad.{name}"]) < tol, name
        else:
            r1, r2 = mg_cases.proj_vectors(g.shape, seed + i)
            assert rel_err(g @ r1, gold[f"{tag}.gproj1.{name}"]) < tol, name
            assert rel_err(r2 @ g, gold[f"{tag}.gproj2.{name}"]) < tol, name


# ------------------------------------------------------------------ integer graph work: bit exact
def test_edge_index_known_answers():
    g = load("edges.npz")
    assert np.array_equal(graph_oracle.build_edge_index(5, 2), g["L5_W2"])
    assert g["L5_W2"][0].tolist() == [0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4]
    assert g["L5_W2"][1].tolist() == [1, 2, 0, 2, 3, 0, 1, 3, 4, 1, 2, 4, 2, 3]
    for L, E in ((64, 3480), (100, 6360), (256, 18840), (512, 39320), (1024, 80280)):
        assert int(g[f"E_L{L}_W40"]) == E == graph_oracle.num_band_edges(L, 40)
    for L in (64, 100):
        ei = graph_oracle.build_edge_index(L, 40)
        assert np.array_equal(ei, g[f"L{L}_W40"].astype(np.int64))
        assert np.array_equal(graph_oracle.degrees(ei, L), g[f"deg_L{L}_W40"])
    assert np.array_equal(graph_oracle.build_edge_index(7, 0), g["L7_W0_fallback"])
    assert np.array_equal(graph_oracle.build_edge_index(6, 9), g["L6_W9_dense"])
    assert np.array_equal(graph_oracle.build_edge_index(2, 1), g["L2_W1"])
    assert graph_oracle.num_band_edges(6, 9) == 30 and graph_oracle.num_band_edges(7, 0) == 12
    assert graph_oracle.build_edge_index(1, 4).shape == (2, 0)


def test_packed_band_graph_is_csr():
    row_ptr, row, col = graph_oracle.packed_band_graph([5, 1, 9], 3)
    assert row_ptr[-1] == len(row) == graph_oracle.num_band_edges(5, 3) + 0 + graph_oracle.num_band_edges(9, 3)
    assert np.all(np.diff(row) >= 0)
    assert np.array_equal(np.repeat(np.arange(15), np.diff(row_ptr)), row)
    assert col[row >= 6].min() >= 6 and col[row < 5].max() <= 4


# ------------------------------------------------------------------ EGNN layer
@pytest.mark.parametrize("tag", list(mg_cases.LAYER_CASES))
def test_layer_matches_reference(tag):
    gold = load("layers.npz")
    H, lengths, W, pseed, dseed, kind = mg_cases.LAYER_CASES[tag]
    n = sum(lengths)
    sd = {k: T(v).requires_grad_() for k, v in synth.make_params(synth.layer_param_shapes(H, H), pseed).items()}
    rng = np.random.default_rng(dseed)
    h = T(synth.f32(rng.standard_normal((n, H)))).requires_grad_()
    x = T(synth.f32(rng.standard_normal((n, 3)) * 2.0)).requires_grad_()
    ei = torch.tensor(gold[f"{tag}.edge_index"])
    if kind == "band":
        assert np.array_equal(np.stack(graph_oracle.packed_band_graph(lengths, W)[1:]), ei.numpy())
        dinv = (1.0 / torch.bincount(ei[0], minlength=n).float()).double()
    else:
        dinv = None
    ch = T(synth.f32(rng.standard_normal((n, H))))
    cx = T(synth.f32(rng.standard_normal((n, 3))))
    h2, x2 = egnn_oracle.egn_layer(sd, "", h, x, ei, dinv)
    assert rel_err(h2.detach(), gold[f"{tag}.h_out"]) < 1e-12
    assert rel_err(x2.detach(), gold[f"{tag}.x_out"]) < 1e-12
    ((h2 * ch).sum() + (x2 * cx).sum()).backward()
    grads = {"h": h.grad, "x": x.grad}
    grads.update({k: v.grad for k, v in sd.items()})
    check_grads(grads, gold, tag, 1e-10)


def test_layer_equivariance_fp64():
    H = 32
    sd = {k: T(v) for k, v in synth.make_params(synth.layer_param_shapes(H, H), 5).items()}
    rng = np.random.default_rng(6)
    h, x = T(rng.standard_normal((11, H))), T(rng.standard_normal((11, 3)))
    ei = torch.from_numpy(graph_oracle.build_edge_index(11, 4))
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    Q, t = T(q), T(rng.standard_normal(3))
    h1, x1 = egnn_oracle.egn_layer(sd, "", h, x, ei)
    h2, x2 = egnn_oracle.egn_layer(sd, "", h, x @ Q + t, ei)
    assert rel_err(h2, h1) < 1e-13 and rel_err(x2, x1 @ Q + t) < 1e-13


# ------------------------------------------------------------------ decoder
@pytest.mark.parametrize("tag", list(mg_cases.DECODER_CASES))
def test_decoder_matches_reference(tag):
    gold = load("decoders.npz")
    case = mg_cases.DECODER_CASES[tag]
    z_g, z_l, H, nl, W, B, L, mkind, pseed, dseed = case
    sd = {k: T(v).requires_grad_() for k, v in
          synth.make_params(synth.decoder_param_shapes(z_g, z_l, H, nl), pseed).items()}
    zg, zl, mask, coef = mg_cases.decoder_inputs(case)
    zg_t, zl_t = T(zg).requires_grad_(), T(zl).requires_grad_()
    outs = egnn_oracle.egnn_decoder(sd, zg_t, zl_t, None if mask is None else T(mask), max_neighbors=W)
    for name, o in zip(("N", "CA", "C", "logits"), outs):
        assert rel_err(o.detach(), gold[f"{tag}.{name}"]) < 1e-11, name
        if mask is not None:
            assert float(o.detach()[torch.tensor(mask) == 0].abs().max() if (mask == 0).any() else 0.0) == 0.0
    sum((o * T(c)).sum() for o, c in zip(outs, coef)).backward()
    grads = {"z_g": zg_t.grad, "z_l": zl_t.grad}
    grads.update({k: v.grad for k, v in sd.items() if v.grad is not None})
    gold_names = {k.split(".", 2)[2] for k in gold.files if k.startswith(tag + ".g")}
    assert gold_names == set(grads), gold_names ^ set(grads)
    check_grads(grads, gold, tag, 1e-9)


@pytest.mark.parametrize("tag", list(mg_cases.BIG_DECODER_CASES))
def test_big_decoder_matches_reference(tag):
    """BASELINE-config shapes (configs[1]: 6 layers x L=256; configs[0]/[3]: 8 layers x L=100 through the reference's
    ResidueDecoder; configs[2]: ragged 512/64/300) and the train-mode path with injected dropout masks."""
    gold = load("decoders_big.npz")
    case = mg_cases.BIG_DECODER_CASES[tag]
    z_g, z_l, H, nl, W, B, L, mkind, pseed, dseed, pdrop = case
    sd = {k: T(v).requires_grad_() for k, v in
          synth.make_params(synth.decoder_param_shapes(z_g, z_l, H, nl), pseed).items()}
    zg, zl, mask, coef = mg_cases.big_decoder_inputs(case)
    zg_t, zl_t = T(zg).requires_grad_(), T(zl).requires_grad_()
    dropout = None
    if pdrop > 0:
        dropout = (pdrop, lambda site, b, rows, dim, p: mg_cases.dropout_keep(dseed, site, b, rows, dim, p))
    outs = egnn_oracle.egnn_decoder(sd, zg_t, zl_t, T(mask), max_neighbors=W, dropout=dropout)
    for name, o in zip(("N", "CA", "C", "logits"), outs):
        assert rel_err(o.detach(), gold[f"{tag}.{name}"]) < 1e-11, name
    sum((o * T(c)).sum() for o, c in zip(outs, coef)).backward()
    grads = {"z_g": zg_t.grad, "z_l": zl_t.grad}
    grads.update({k: v.grad for k, v in sd.items() if v.grad is not None})
    gold_names = {k.split(".", 2)[2] for k in gold.files if k.startswith(tag + ".g")}
    assert gold_names == set(grads), gold_names ^ set(grads)
    check_grads({k: torch.as_tensor(mg_cases.flat2d(v.detach().numpy())) for k, v in grads.items()}, gold, tag, 1e-9,
                seed=4321)


# ------------------------------------------------------------------ losses
@pytest.mark.parametrize("tag", list(mg_cases.LOSS_CASES))
def test_losses_match_reference(tag):
    gold = load("losses.npz")
    case = mg_cases.LOSS_CASES[tag]
    d = mg_cases.loss_inputs(case)
    mask = T(d["mask"])
    tgt = [T(d[k]) for k in ("target_N", "target_CA", "target_C")]
    tdih = losses_oracle.compute_dihedrals_from_coords(*tgt, mask)
    assert rel_err(tdih, gold[f"{tag}.target_dihedrals"]) < 1e-12
    pdih = losses_oracle.compute_dihedrals_from_coords(T(d["pred_N"]), T(d["pred_CA"]), T(d["pred_C"]), mask)
    assert rel_err(pdih, gold[f"{tag}.pred_dihedrals"]) < 1e-12
    for stride in case[5]:
        leaves = {k: T(d[k]).requires_grad_() for k in mg_cases.GRAD_INPUTS}
        res = losses_oracle.compute_total_loss(
            leaves["pred_N"], leaves["pred_CA"], leaves["pred_C"], leaves["pred_seq"], tgt[0], tgt[1], tgt[2],
            torch.tensor(d["labels"]), mask, leaves["mu_g"], leaves["lv_g"], leaves["mu_l"], leaves["lv_l"],
            tdih, pair_stride=stride, **mg_cases.LOSS_WEIGHTS)
        assert tuple(res) == losses_oracle.LOSS_KEYS
        for k, v in res.items():
            ref = float(gold[f"{tag}.s{stride}.{k}"])
            assert abs(float(v) - ref) <= 2e-7 * max(abs(ref), 1e-3), (k, float(v), ref)   # cdist mm path, F7
        res["total"].backward()
        for k, v in leaves.items():
            assert rel_err(v.grad, gold[f"{tag}.s{stride}.grad.{k}"]) < 1e-6, k


def test_losses_exercise_every_branch():
    """The fixtures must actually hit clashes, both Huber branches and the forbidden quadrant."""
    gold = load("losses.npz")
    assert float(gold["compact.s4.clash"]) > 1e-3 and float(gold["walk.s4.clash"]) >= 0.0
    assert float(gold["walk.s4.ramachandran"]) > 1.0       # 5.0 forbidden-quadrant penalties present
    assert float(gold["walk.s4.bond_length"]) > 0.01       # linear Huber branch


def test_ideal_backbone_known_answers():
    gold = load("losses.npz")
    N, CA, C = (T(a)[None] for a in synth.nerf_backbone(32))
    ones = torch.ones(1, 32, dtype=torch.float64)
    dih = losses_oracle.compute_dihedrals_from_coords(N, CA, C, ones)
    assert rel_err(dih, gold["nerf.dihedrals"]) < 1e-12
    deg = lambda s, c: np.degrees(np.arctan2(s, c))  # noqa: E731
    d = dih[0].numpy()
    assert np.allclose(deg(d[1:, 0], d[1:, 1]), -60.0, atol=1e-3)
    assert np.allclose(deg(d[:-1, 2], d[:-1, 3]), -45.0, atol=1e-3)
    assert np.allclose(np.abs(deg(d[1:, 4], d[1:, 5])), 180.0, atol=2e-2)
    assert float(losses_oracle.bond_length_loss(N, CA, C, ones)) < 1e-10
    assert float(losses_oracle.bond_angle_loss(N, CA, C, ones)) < 1e-10
    assert abs(float(losses_oracle.omega_trans_loss(dih, ones)) - float(gold["nerf.omega_trans"])) < 1e-9
    assert abs(float(gold["nerf.omega_trans"]) - 7.0 / 32) < 1e-6       # residue 0 contributes 7.0 (F10)
    assert float(losses_oracle.clash_loss(N, CA, C, ones)) == float(gold["nerf.clash"]) == 0.0
    assert float(losses_oracle.pair_distance_loss(CA, CA, ones, stride=4)) == 0.0
    assert abs(float(losses_oracle.pair_distance_loss(1.1 * CA, CA, ones, stride=4))
               - float(gold["nerf.pair_scaled"])) < 1e-7
    assert abs(float(losses_oracle.ramachandran_loss(dih, ones)) - float(gold["nerf.ramachandran"])) < 1e-10


# ------------------------------------------------------------------ Kabsch
def test_kabsch_matches_reference():
    gold = load("kabsch.npz")
    a, b, mask = mg_cases.kabsch_inputs()
    for s in range(a.shape[0]):
        x, y = a[s].astype(np.float64), b[s].astype(np.float64)
        assert abs(kabsch_oracle.kabsch_rmsd_ref_compat(x, y, mask[s]) - gold["ref_compat"][s]) < 1e-9
        opt = kabsch_oracle.kabsch_rmsd(x, y, mask[s])
        assert abs(opt - gold["optimal"][s]) < 1e-9
        assert abs(kabsch_oracle.kabsch_rmsd_closed_form(x, y, mask[s]) - opt) < 1e-7
    assert gold["ref_compat"][5] == 0.0 and gold["optimal"][5] == 0.0
    assert gold["ref_compat"][0] > 10 * gold["optimal"][0]      # the inverse-rotation defect (F6)
