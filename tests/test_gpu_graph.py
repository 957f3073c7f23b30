"""GraphedStep: CUDA-graph replay of forward + backward equals the eager step and follows parameter updates."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _setup(dropout):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    from protein_ensemble_vae_b200 import EGNNDecoder, compute_total_loss
    from protein_ensemble_vae_b200 import losses as pl
    torch.manual_seed(3)
    B, L = 6, 48
    dec = EGNNDecoder(bench.CFG["z_g"], bench.CFG["z_l"], hidden_dim=256, num_layers=2, max_neighbors=40, dropout=dropout,
                      precision="bf16").cuda().train()
    batches = [bench.synth_batch(B, L, bench.CFG["z_g"], bench.CFG["z_l"], seed=s, device="cuda") for s in (0, 1)]
    tdih = pl.compute_dihedrals_from_coords(batches[0]["target_N"], batches[0]["target_CA"], batches[0]["target_C"],
                                            batches[0]["mask"])

    def step(d):
        o = dec(d["z_g"], d["z_l"], d["mask"])
        r = compute_total_loss(o[0], o[1], o[2], o[3], d["target_N"], d["target_CA"], d["target_C"], d["labels"], d["mask"],
                               d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih, **bench.LOSS_W)
        r["total"].backward()
        return r["total"].detach()
    return dec, batches, step


def test_graphed_step_matches_eager_and_tracks_parameter_updates():
    from protein_ensemble_vae_b200 import GraphedStep
    dec, batches, step = _setup(dropout=0.0)
    params = list(dec.parameters())
    g = GraphedStep(step, batches[0], params)
    opt = torch.optim.SGD(params, lr=1e-3)
    for it, d in enumerate((batches[1], batches[0], batches[1])):
        loss_g = float(g(d))
        grads_g = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        loss_e = float(step(d))
        assert abs(loss_g - loss_e) <= 2e-4 * abs(loss_e), (it, loss_g, loss_e)
        for p, gg in zip(params, grads_g):
            assert rel_err(gg, p.grad) < 2e-3                 # atomics in the segment sums: not bit-reproducible
        g.attach_grads()
        for p, gg in zip(g.params, grads_g):
            p.grad.copy_(gg)
        opt.step()                                            # the next replay must see the updated weights (repacked)


def test_graphed_step_host_batch_check():
    from protein_ensemble_vae_b200 import GraphedStep
    dec, batches, step = _setup(dropout=0.1)
    g = GraphedStep(step, batches[0], list(dec.parameters()))
    host = {k: v.cpu() for k, v in batches[1].items()}
    assert g.matches(host)
    host["mask"] = host["mask"].clone()
    host["mask"][0, -1] = 0
    assert not g.matches(host)
    assert torch.isfinite(g(batches[1]))
