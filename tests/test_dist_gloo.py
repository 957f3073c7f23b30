"""N > 1 host logic on CPU: world_size-2 gloo processes (no GPU).  The product kernels are CUDA-only, so
each process runs the package's Python layer against tests/hostcheck (the kernels' per-thread bodies built
for the host) -- what is under test here is the sharding and the gradient exchange, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_decoder():
    import synth
    from protein_ensemble_vae_b200 import EGNNDecoder
    dec = EGNNDecoder(12, 6, hidden_dim=32, num_layers=2, max_neighbors=3, dropout=0.0, precision="fp32").eval()
    params = synth.make_params(synth.decoder_param_shapes(12, 6, 32, 2), 21)
    dec.load_state_dict({k: torch.tensor(v) for k, v in params.items()})
    return dec


def _batch(B=4, L=10):
    import cases
    rng = np.random.default_rng(3)
    f = lambda *s: torch.tensor(rng.standard_normal(s).astype(np.float32))  # noqa: E731
    d = cases.loss_inputs((B, L, "full", 5, 1.0, (4,)))
    d = {k: torch.tensor(v) for k, v in d.items()}
    d["z_g"], d["z_l"] = f(B, 12), f(B, L, 6)
    return d


def _loss(dec, d, sl):
    import cases
    from protein_ensemble_vae_b200 import compute_total_loss
    from protein_ensemble_vae_b200 import losses as pl
    n, ca, c, lg = dec(d["z_g"][sl], d["z_l"][sl], d["mask"][sl])
    tdih = pl.compute_dihedrals_from_coords(d["target_N"][sl], d["target_CA"][sl], d["target_C"][sl], d["mask"][sl])
    res = compute_total_loss(n, ca, c, lg, d["target_N"][sl], d["target_CA"][sl], d["target_C"][sl], d["labels"][sl],
                             d["mask"][sl], d["mu_g"][sl], d["lv_g"][sl], d["mu_l"][sl], d["lv_l"][sl], tdih,
                             pair_stride=4, **cases.LOSS_WEIGHTS)
    return res["total"]


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from hostlib import host_backend
        from protein_ensemble_vae_b200 import distributed as pd
        torch.set_num_threads(1)
        with host_backend():
            dec = _make_decoder()
            d = _batch()
            lo, hi = pd.shard_range(d["z_l"].shape[0], rank, ws)
            _loss(dec, d, slice(lo, hi)).backward()
            params = list(dec.parameters())
            pd.allreduce_gradients(params)                               # the step's only collective
            flat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])
            ref_ca = d["target_CA"][0]
            rm = pd.decode_ensemble(dec, d["z_g"], d["z_l"], d["mask"][0], ref_ca, chunk=1)
            rm_local = pd.decode_ensemble(dec, d["z_g"], d["z_l"], d["mask"][0], ref_ca, chunk=3, gather=False)
        if rank == 0:
            torch.save({"grad": flat, "rmsd": rm, "n_local": rm_local.numel()}, out)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from protein_ensemble_vae_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 100000):
        for ws in (1, 2, 3, 8):
            r = [shard_range(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gradients_and_sharded_decode(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    from hostlib import host_backend
    from protein_ensemble_vae_b200 import distributed as pd
    with host_backend():
        dec = _make_decoder()
        d = _batch()
        _loss(dec, d, slice(0, 4)).backward()                            # single process, global batch
        want = torch.cat([p.grad.reshape(-1) for p in dec.parameters() if p.grad is not None])
        rm = pd.decode_ensemble(dec, d["z_g"], d["z_l"], d["mask"][0], d["target_CA"][0])
    # mean-of-means == global mean for equal shards with full masks (distributed.py docstring)
    assert float((got["grad"] - want).abs().max() / want.abs().max()) < 2e-5
    assert got["n_local"] == 2
    assert torch.allclose(got["rmsd"], rm, atol=1e-6) and rm.numel() == 4


# ------------------------------------------------------------------ ragged shards (BASELINE config 3)
def _ragged_batch():
    """5 conformers with different numbers of valid residues (one with an interior gap); rank 0 gets 2, rank 1 gets 3."""
    d = _batch(B=5, L=12)
    m = torch.ones(5, 12)
    for b, n in enumerate((12, 7, 10, 5, 9)):
        m[b, n:] = 0
    m[2, 3:5] = 0
    d["mask"] = m
    return d


def _loss_dp(dec, d, idx, dp):
    import cases
    from protein_ensemble_vae_b200 import compute_total_loss
    from protein_ensemble_vae_b200 import losses as pl
    sl = torch.tensor(idx)
    n, ca, c, lg = dec(d["z_g"][sl], d["z_l"][sl], d["mask"][sl])
    tdih = pl.compute_dihedrals_from_coords(d["target_N"][sl], d["target_CA"][sl], d["target_C"][sl], d["mask"][sl])
    res = compute_total_loss(n, ca, c, lg, d["target_N"][sl], d["target_CA"][sl], d["target_C"][sl], d["labels"][sl],
                             d["mask"][sl], d["mu_g"][sl], d["lv_g"][sl], d["mu_l"][sl], d["lv_l"][sl], tdih,
                             pair_stride=4, dp_normalize=dp, **cases.LOSS_WEIGHTS)
    return res


def _ragged_worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from hostlib import host_backend
        from protein_ensemble_vae_b200 import distributed as pd
        torch.set_num_threads(1)
        with host_backend():
            dec = _make_decoder()
            d = _ragged_batch()
            buckets = pd.GradBuckets(pd.decoder_buckets(dec))            # gradients accumulate into flat buckets,
            idx = [0, 1] if rank == 0 else [2, 3, 4]                     # all-reduced while backward is still running
            res = _loss_dp(dec, d, idx, True)
            res["total"].backward()
            buckets.finish()
            tot = res["total"].detach().clone()
            dist.all_reduce(tot)
            grads = {k: p.grad.clone() for k, p in dec.named_parameters() if p.grad is not None}
            buckets.zero()
            assert all(float(p.grad.abs().max()) == 0.0 for p in dec.parameters() if p.grad is not None)
        if rank == 0:
            torch.save({"grads": grads, "total": tot / ws}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_ragged_shards_equal_the_global_batch(tmp_path):
    """Unequal conformer and valid-residue counts per rank: with dp_normalize the rank-averaged loss and gradients
    equal the single-process global batch (numerators and denominators are combined, not per-rank means)."""
    out = str(tmp_path / "r0.pt")
    mp.spawn(_ragged_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    from hostlib import host_backend
    with host_backend():
        dec = _make_decoder()
        d = _ragged_batch()
        res = _loss_dp(dec, d, [0, 1, 2, 3, 4], False)
        res["total"].backward()
        want = {k: p.grad for k, p in dec.named_parameters() if p.grad is not None}
        # and the naive mean of per-rank means is NOT the global batch (what dp_normalize repairs)
        naive = 0.5 * (_loss_dp(dec, d, [0, 1], False)["total"] + _loss_dp(dec, d, [2, 3, 4], False)["total"])
    assert abs(float(got["total"]) - float(res["total"])) < 2e-5 * abs(float(res["total"]))
    assert abs(float(naive) - float(res["total"])) > 1e-3 * abs(float(res["total"]))
    assert set(got["grads"]) == set(want)
    for k in want:
        assert float((got["grads"][k] - want[k]).abs().max()) <= 2e-5 * float(want[k].abs().max()) + 1e-7, k


def test_balanced_shards_equalise_edge_work():
    from protein_ensemble_vae_b200.distributed import balanced_shards, band_edges
    assert [band_edges(L, 40) for L in (1, 2, 41, 64, 256, 512)] == [0, 2, 1640, 3480, 18840, 39320]
    rng = np.random.default_rng(0)
    lengths = rng.integers(64, 513, 256).tolist()
    for ws in (1, 2, 4, 8):
        shards = balanced_shards(lengths, ws)
        assert sorted(i for s in shards for i in s) == list(range(256))
        loads = [sum(band_edges(lengths[i], 40) for i in s) for s in shards]
        assert (max(loads) - min(loads)) <= 0.01 * max(loads)           # contiguous equal-count shards are off by ~10 %
