"""BASELINE.json's configurations on a real B200, checked through size-independent properties where the
oracle would take too long: E(3) equivariance, padding invariance, shard invariance,
Kabsch invariances, loss consistency between the fused path and the per-function API."""
import math

import numpy as np
import pytest
import torch

import cases
from conftest import rel_err
from oracle import kabsch_oracle, losses_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rot(seed):
    q, _ = np.linalg.qr(np.random.default_rng(seed).standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    return torch.tensor(q, dtype=torch.float32, device=DEV)


def _decoder(layers, precision, seed=0, W=40):
    from protein_ensemble_vae_b200 import EGNNDecoder
    torch.manual_seed(seed)
    return EGNNDecoder(64, 32, hidden_dim=256, num_layers=layers, max_neighbors=W, dropout=0.0,
                       precision=precision).to(DEV).eval()


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_layer_equivariance_config2_shape(precision, tol):
    """EGNLayer(h, xQ+t) == (h', x'Q+t) at L=256, W=40."""
    from protein_ensemble_vae_b200 import EGNLayer
    from protein_ensemble_vae_b200.graph import band_graph
    torch.manual_seed(1)
    layer = EGNLayer(256, 256, precision=precision).to(DEV)
    g = band_graph((256,) * 4, 40, DEV)
    h = torch.randn(g.num_nodes, 256, device=DEV)
    x = torch.randn(g.num_nodes, 3, device=DEV) * 3
    Q, t = _rot(2), torch.tensor([1.0, -2.0, 0.5], device=DEV)
    with torch.no_grad():
        h1, x1 = layer(h, x, g, g.dinv)
        h2, x2 = layer(h, x @ Q + t, g, g.dinv)
    assert rel_err(h2, h1) < tol and rel_err(x2, x1 @ Q + t) < tol


def test_mixed_lengths_config3_padding_and_batch_invariance():
    """config 3: ragged lengths 64..512 with interior gaps; a conformer's output must not depend on its batch
    mates or on where the padding sits, and padded rows are exact zeros."""
    dec = _decoder(3, "fp32")
    rng = np.random.default_rng(3)
    B, L = 6, 512
    lengths = [512, 64, 300, 129, 257, 400]
    mask = torch.zeros(B, L, device=DEV)
    for b, n in enumerate(lengths):
        mask[b, :n] = 1
    mask[2, 100:105] = 0
    mask[4, 7:9] = 0
    zg = torch.randn(B, 64, device=DEV)
    zl = torch.randn(B, L, 32, device=DEV)
    with torch.no_grad():
        full = dec(zg, zl, mask)
        for b in (1, 2, 5):
            solo = dec(zg[b:b + 1], zl[b:b + 1], mask[b:b + 1])
            for o_full, o_solo in zip(full, solo):
                assert rel_err(o_full[b], o_solo[0]) < 1e-5
        # compacting the valid residues to the front gives the same per-residue outputs
        b = 2
        idx = torch.nonzero(mask[b]).squeeze(-1)
        zl_c = torch.zeros(1, L, 32, device=DEV)
        zl_c[0, :idx.numel()] = zl[b, idx]
        m_c = torch.zeros(1, L, device=DEV)
        m_c[0, :idx.numel()] = 1
        comp = dec(zg[b:b + 1], zl_c, m_c)
        for o_full, o_c in zip(full, comp):
            assert rel_err(o_full[b, idx], o_c[0, :idx.numel()]) < 1e-5
    for o in full:
        assert float(o[mask == 0].abs().max()) == 0.0


def test_decoder_bf16_tracks_fp32_at_config2_depth():
    """6 layers, L=256.  CA coordinates and logits of the tcgen05 path stay within 1e-2 of the fp32 exact-order
    path.  N / C = CA + 1.46 normalize(head(h)) are ill-conditioned at random init (normalising near-zero
    direction vectors): there the yardstick is an ideal bf16 edge MLP -- the fp32 path with the three edge linears
    rounded the way torch.autocast would (tests/bf16_yardstick.py); the tensor-core path must not drift more
    than 1.5x that (tools/bf16_drift.py prints the table for 1-8 layers)."""
    from bf16_yardstick import emulate_autocast_edge_mlp
    d32, dac, d16 = (_decoder(6, p, seed=5) for p in ("fp32", "fp32", "bf16"))
    dac.load_state_dict(d32.state_dict())
    d16.load_state_dict(d32.state_dict())
    dac.precision = "autocast-emu"
    zg, zl = torch.randn(4, 64, device=DEV), torch.randn(4, 256, 32, device=DEV)
    with torch.no_grad(), emulate_autocast_edge_mlp():
        ref, emu, got = d32(zg, zl), dac(zg, zl), d16(zg, zl)
    assert rel_err(got[1], ref[1]) < 1e-2 and rel_err(got[3], ref[3]) < 1e-2
    for r, e, g in zip(ref, emu, got):
        assert rel_err(g, r) < 1.5 * rel_err(e, r) + 2e-3


def test_stress_config5_l1024_band_and_dense():
    """config 5: L=1024 with the W=40 band (E=80 280) and the dense graph W=1023 (E=1 047 552) through the
    generic CSR path; full pairwise losses with pair_stride=1 and the 4.7 M-pair clash tile."""
    from protein_ensemble_vae_b200 import EGNNDecoder, compute_total_loss
    from protein_ensemble_vae_b200 import losses as pl
    from protein_ensemble_vae_b200.graph import band_edge_count, band_graph
    assert band_edge_count(1024, 40) == 80280 and band_edge_count(1024, 1023) == 1047552
    g = band_graph((1024,), 1023, DEV)
    assert g.num_edges == 1047552 and int(g.row_ptr[-1]) == 1047552
    B, L = 2, 1024
    for W, prec in ((40, "bf16"), (1023, "bf16"), (1023, "fp32")):
        torch.manual_seed(0)
        dec = EGNNDecoder(64, 32, hidden_dim=256, num_layers=2, max_neighbors=W, dropout=0.0, precision=prec).to(DEV)
        zg = torch.randn(B, 64, device=DEV)
        zl = torch.randn(B, L, 32, device=DEV, requires_grad=True)
        mask = torch.ones(B, L, device=DEV)
        n, ca, c, lg = dec(zg, zl, mask)
        d = cases.loss_inputs((B, L, "full", 9, 1.0, (1,)))
        t = lambda a: torch.tensor(a, device=DEV)  # noqa: E731
        tn, tca, tc = t(d["target_N"]), t(d["target_CA"]), t(d["target_C"])
        tdih = pl.compute_dihedrals_from_coords(tn, tca, tc, mask)
        res = compute_total_loss(n, ca, c, lg, tn, tca, tc, t(d["labels"]), mask, t(d["mu_g"]), t(d["lv_g"]),
                                 t(d["mu_l"]), t(d["lv_l"]), tdih, pair_stride=1, **cases.LOSS_WEIGHTS)
        res["total"].backward()
        assert all(torch.isfinite(v) for v in res.values()) and torch.isfinite(zl.grad).all()
        if W == 40:      # loss values of the fused kernels vs the float64 oracle on the same predictions
            T64 = lambda a: a.detach().cpu().double()  # noqa: E731
            chk = losses_oracle.compute_total_loss(
                T64(n), T64(ca), T64(c), T64(lg), T64(tn), T64(tca), T64(tc), torch.tensor(d["labels"]), T64(mask),
                T64(t(d["mu_g"])), T64(t(d["lv_g"])), T64(t(d["mu_l"])), T64(t(d["lv_l"])), T64(tdih), pair_stride=1,
                **cases.LOSS_WEIGHTS)
            for k in chk:
                assert abs(float(res[k]) - float(chk[k])) <= 2e-5 * max(abs(float(chk[k])), 1e-3), k


def test_generation_config4_sharded_decode_and_kabsch():
    """config 4 shape (L=100, 8 layers): chunked decode == one-shot decode; Kabsch RMSD is invariant to rigid
    motion of either structure and matches the numpy oracle."""
    from protein_ensemble_vae_b200 import ResidueDecoder, kabsch_rmsd_batch
    from protein_ensemble_vae_b200 import distributed as pd
    torch.manual_seed(0)
    dec = ResidueDecoder(64, 32, dropout=0.1, precision="bf16").to(DEV).eval()
    S, L = 300, 100
    zg, zl = torch.randn(S, 64, device=DEV), torch.randn(S, L, 32, device=DEV)
    mask = torch.ones(L, device=DEV)
    ref = torch.cumsum(torch.randn(L, 3, device=DEV) * 2.2, 0)
    with torch.no_grad():
        one = dec(zg, zl, mask.expand(S, L))[1]
    r1, coords = pd.decode_ensemble(dec, zg, zl, mask, ref, chunk=77, return_coords=True)
    assert rel_err(coords, one) < 5e-3          # fp32 atomics in the aggregation are order-dependent across batches
    r2 = kabsch_rmsd_batch(one, ref, mask)
    assert rel_err(r1, r2) < 5e-3
    Q, t = _rot(4), torch.tensor([3.0, 1.0, -2.0], device=DEV)
    assert rel_err(kabsch_rmsd_batch(one @ Q + t, ref, mask), r2) < 1e-5
    assert rel_err(kabsch_rmsd_batch(one, ref @ Q + t, mask), r2) < 1e-5
    for s in (0, 17, 299):
        want = kabsch_oracle.kabsch_rmsd(one[s].cpu().double().numpy(), ref.cpu().double().numpy(), np.ones(L))
        assert abs(float(r2[s]) - want) < 1e-5 * max(want, 1.0)
    assert float(kabsch_rmsd_batch(ref[None] @ Q + t, ref, mask)[0]) < 1e-4        # rigid copy -> ~0


def test_single_residue_and_empty_conformers():
    """Lb == 1 crashes the reference; here it decodes (no edges) and empty rows give zeros."""
    dec = _decoder(2, "bf16")
    mask = torch.zeros(3, 9, device=DEV)
    mask[0, 4] = 1
    mask[2, :5] = 1
    with torch.no_grad():
        n, ca, c, lg = dec(torch.randn(3, 64, device=DEV), torch.randn(3, 9, 32, device=DEV), mask)
    assert torch.isfinite(ca).all() and float(ca[1].abs().max()) == 0.0 and float(lg[1].abs().max()) == 0.0
    assert float(ca[0, 4].abs().max()) > 0 and float(ca[0, :4].abs().max()) == 0.0


def test_device_prefetcher_yields_batches_in_order():
    """Double-buffered H2D staging (protein_ensemble_vae_b200/data.py): every batch arrives intact and in order,
    also when the consumer is slow or the iterable is empty / a single batch."""
    from protein_ensemble_vae_b200 import DevicePrefetcher
    host = [{"a": torch.full((1000, 64), float(i)).pin_memory(), "b": torch.arange(5) + i} for i in range(7)]
    seen = []
    for d in DevicePrefetcher(host, DEV):
        junk = torch.randn(2048, 2048, device=DEV) @ torch.randn(2048, 2048, device=DEV)   # keep the stream busy
        seen.append((float(d["a"].max()), int(d["b"][0]), float(d["a"].min()), float(junk.sum() * 0)))
    assert [(s[0], s[1], s[2]) for s in seen] == [(float(i), i, float(i)) for i in range(7)]
    assert list(DevicePrefetcher([], DEV)) == []
    one = list(DevicePrefetcher(host[:1], DEV))
    assert len(one) == 1 and float(one[0]["a"].max()) == 0.0


@pytest.mark.parametrize("S,L", [(1, 20), (2, 33), (7, 100), (40, 64)])
def test_kabsch_rmsd_pairs_matches_oracle_pair_by_pair(S, L):
    """All-pairs RMSD matrix of one ensemble (generate_ensemble_pdbs.py:591-595 in one launch) against the numpy oracle
    for every pair, in both conventions; symmetric, zero diagonal; the mean is the reference's avg_diversity."""
    from protein_ensemble_vae_b200 import ensemble_diversity, kabsch_rmsd_pairs
    rng = np.random.default_rng(S * 131 + L)
    base = np.cumsum(rng.standard_normal((L, 3)) * 2.0, 0)
    ens = np.stack([base + 0.7 * rng.standard_normal((L, 3)) for _ in range(S)]).astype(np.float32)
    mask = np.ones(L, dtype=np.float32)
    mask[rng.integers(0, L, 3)] = 0
    c, m = torch.tensor(ens, device=DEV), torch.tensor(mask, device=DEV)
    for compat, fn in ((False, kabsch_oracle.kabsch_rmsd), (True, kabsch_oracle.kabsch_rmsd_ref_compat)):
        M = kabsch_rmsd_pairs(c, m, ref_compat=compat).cpu().numpy()
        assert M.shape == (S, S) and np.all(np.diag(M) == 0) and np.array_equal(M, M.T)
        vals = []
        for i in range(S):
            for j in range(i + 1, S):
                want = fn(ens[i].astype(np.float64), ens[j].astype(np.float64), mask)
                assert abs(M[i, j] - want) <= 1e-5 * max(want, 1e-3) + 2e-6, (i, j, M[i, j], want)
                vals.append(want)
        div = float(ensemble_diversity(c, m, ref_compat=compat))
        assert abs(div - (np.mean(vals) if vals else 0.0)) <= 1e-5 * max(np.mean(vals) if vals else 0.0, 1e-3) + 2e-6
    M0 = kabsch_rmsd_pairs(c, None).cpu().numpy()                           # no mask
    if S >= 2:
        want = kabsch_oracle.kabsch_rmsd(ens[0].astype(np.float64), ens[S - 1].astype(np.float64), np.ones(L))
        assert abs(M0[0, S - 1] - want) <= 1e-5 * want + 2e-6
