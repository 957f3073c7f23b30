"""Evaluation metrics: oracle restatement vs the reference's own outputs (tests/golden/metrics.npz,
produced by importing scripts/validation_metrics.py), and the kernels (host-compiled bodies / CUDA) vs both."""
import os

import numpy as np
import torch

import cases
from oracle import metrics_oracle as mo

G = os.path.join(os.path.dirname(__file__), "golden")


def test_metrics_oracle_matches_reference():
    gold = np.load(os.path.join(G, "metrics.npz"))
    pred, true, mask, ens = cases.metrics_inputs()
    for s in range(pred.shape[0]):
        p, t, m = pred[s].astype(np.float64), true[s].astype(np.float64), mask[s]
        assert abs(mo.tm_score(p, t) - gold["tm"][s]) < 1e-12
        assert np.allclose(mo.gdt(p, t), (gold["gdt_ts"][s], gold["gdt_ha"][s]), atol=1e-12)
        assert np.allclose(mo.gdt(p, t, m), (gold["gdt_ts_masked"][s], gold["gdt_ha_masked"][s]), atol=1e-12)
        g, r = mo.lddt(p, t)
        assert abs(g - gold["lddt"][s]) < 1e-12 and np.allclose(r, gold["lddt_res"][s], atol=1e-12)
        g, r = mo.lddt(p, t, m)
        assert abs(g - gold["lddt_masked"][s]) < 1e-12 and np.allclose(r, gold["lddt_res_masked"][s], atol=1e-12)
    assert np.allclose(mo.rmsf(ens.astype(np.float64)), gold["rmsf"], atol=1e-10)
    assert np.allclose(mo.kabsch_align(pred[0].astype(np.float64), true[0].astype(np.float64)), gold["aligned0"], atol=1e-10)
    assert gold["tm"][0] > 0.99 and gold["tm"][5] < 0.5 and 0 < gold["lddt"][3] < 1          # the cases span the range


def test_metrics_kernels_match_reference(bk):
    from protein_ensemble_vae_b200 import metrics as pm
    gold = np.load(os.path.join(G, "metrics.npz"))
    pred, true, mask, ens = cases.metrics_inputs()
    P, T, M = bk.t32(pred), bk.t32(true), bk.t32(mask)
    with bk.ctx():
        sp = pm.superpose(P, T)
        gts, gha = pm.compute_gdt(P, T, M)
        lg, lr = pm.compute_lddt(P, T)
        lgm, lrm = pm.compute_lddt(P, T, M)
        one_tm = pm.compute_tm_score(P[2], T[2])
        one_l, one_lr = pm.compute_lddt(P[3], T[3], M[3] > 0)
        shared = pm.compute_tm_score(P, T[0])
        rm = pm.compute_rmsf(bk.t32(ens))
        rm1 = pm.compute_rmsf(bk.t32(ens[:1]))
    c = lambda t: t.cpu().numpy().astype(np.float64)  # noqa: E731
    assert np.allclose(c(sp["tm"]), gold["tm"], atol=1e-5) and np.allclose(c(sp["gdt_ts"]), gold["gdt_ts"], atol=1e-4)
    assert np.allclose(c(sp["gdt_ha"]), gold["gdt_ha"], atol=1e-4)
    assert np.allclose(c(gts), gold["gdt_ts_masked"], atol=1e-4) and np.allclose(c(gha), gold["gdt_ha_masked"], atol=1e-4)
    assert np.allclose(c(sp["aligned"][0]), gold["aligned0"], atol=2e-4)
    assert np.allclose(c(lg), gold["lddt"], atol=1e-5) and np.allclose(c(lr), gold["lddt_res"], atol=1e-5)
    assert np.allclose(c(lgm), gold["lddt_masked"], atol=1e-5) and np.allclose(c(lrm), gold["lddt_res_masked"], atol=1e-5)
    assert abs(float(one_tm) - gold["tm"][2]) < 1e-5 and abs(float(one_l) - gold["lddt_masked"][3]) < 1e-5
    assert np.allclose(c(one_lr), gold["lddt_res_masked"][3], atol=1e-5)
    assert abs(float(shared[0]) - gold["tm"][0]) < 1e-5 and shared.shape == (pred.shape[0],)
    assert np.allclose(c(rm), gold["rmsf"], atol=2e-5) and float(rm1.abs().max()) == 0.0
