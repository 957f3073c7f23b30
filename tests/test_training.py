"""``run_epoch`` (models/training.py:22-199, the caller of the hot path): same signature and result keys; on the device it
must give the numbers of the reference's per-batch loop (written out below with its ``.item()`` read-backs) from the same
initial state."""
import copy
import inspect

import numpy as np
import pytest
import torch

import cases

REF_ARGS = ["model", "loader", "opt", "device", "klw_g", "klw_l", "w_pair", "pair_stride", "train", "w_dihedral", "w_rama",
            "w_bond", "w_angle", "w_rec", "w_seq", "w_clash", "epoch"]                     # models/training.py:22-23
REF_KEYS = ["loss", "rec", "pair", "klg", "kll", "dihedral", "rama", "bond", "angle", "seq", "seq_acc", "clash"]  # :186-199
W = dict(klw_g=1.0, klw_l=0.5, w_pair=10.0, pair_stride=4, w_dihedral=20.0, w_rama=400.0, w_bond=500.0, w_angle=500.0,
         w_rec=10.0, w_seq=50.0, w_clash=300.0)


def test_run_epoch_signature_and_host_side_plumbing():
    from protein_ensemble_vae_b200 import training as tr
    sig = inspect.signature(tr.run_epoch)
    positional = [n for n, p in sig.parameters.items() if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert positional == REF_ARGS
    assert list(tr.STAT_KEYS) == REF_KEYS
    assert set(tr._LOSS_OF_STAT) == set(REF_KEYS) - {"seq_acc"}
    # loader items: the reference's pair of padded 7-tuples, or a pair of packed dicts
    t7 = lambda: (torch.zeros(2, 5, 3), torch.zeros(2, 5, 3), torch.zeros(2, 5, 3), torch.ones(2, 5), None,   # noqa: E731
                  torch.zeros(2, 5, 6), torch.zeros(2, 5, dtype=torch.int64))
    d = tr._as_host_dict((t7(), t7()))
    assert set(d) == {f"{s}.{k}" for s in ("in", "tgt") for k in ("n", "ca", "c", "mask", "dih", "labels")}
    side = tr._side(d, "tgt", "cpu")
    assert side[4] is None and side[6].dtype == torch.int64 and side[3].shape == (2, 5)
    packed = {"n": torch.zeros(7, 3), "cu_seqlens": torch.tensor([0, 3, 7], dtype=torch.int32), "lmax": 4, "emb": None}
    d = tr._as_host_dict((packed, packed))
    tensors, scalars = tr._HostScalars.split(d)
    assert scalars == {"in.lmax": 4, "tgt.lmax": 4} and "in.emb" not in d and "tgt.cu_seqlens" in tensors


def _conformers(n_items, seed, emb_dim):
    """Centred conformers of unequal lengths as the reference dataset yields them (models/data.py:153-194)."""
    rng = np.random.default_rng(seed)
    items = []
    for L in rng.integers(14, 25, n_items):
        n, ca, c = (torch.tensor(a[0]) for a in cases.synth.make_backbone(1, int(L), int(rng.integers(1 << 30))))
        cen = ca.mean(0, keepdim=True)
        ang = rng.uniform(-np.pi, np.pi, (int(L), 3))
        dih = np.stack([np.sin(ang[:, 0]), np.cos(ang[:, 0]), np.sin(ang[:, 1]), np.cos(ang[:, 1]), np.sin(ang[:, 2]),
                        np.cos(ang[:, 2])], -1)
        items.append((n - cen, ca - cen, c - cen, torch.ones(int(L)), torch.tensor(rng.standard_normal((int(L), emb_dim)),
                                                                                    dtype=torch.float32),
                      torch.tensor(dih, dtype=torch.float32), torch.tensor(rng.integers(0, 20, int(L)))))
    return items


def _collate_pad(items):
    """models/data.py:219-266: zero padding to the batch maximum."""
    Lm = max(it[0].shape[0] for it in items)

    def pad(k, tail, dtype=torch.float32):
        out = torch.zeros(len(items), Lm, *tail, dtype=dtype)
        for b, it in enumerate(items):
            out[b, :it[k].shape[0]] = it[k]
        return out
    return (pad(0, (3,)), pad(1, (3,)), pad(2, (3,)), pad(3, ()), pad(4, (items[0][4].shape[1],)), pad(5, (6,)),
            pad(6, (), torch.int64))


def _reference_loop(model, loader, opt, train, dev):
    """The arithmetic of models/training.py:57-199, read-backs included."""
    from protein_ensemble_vae_b200 import compute_total_loss
    model.train(train)
    tot = {k: 0.0 for k in REF_KEYS}
    n = 0
    for inp, tgt in loader:
        n_in, ca_in, c_in, _, emb_in, dih_in, _ = (t.to(dev) for t in inp)
        n_t, ca_t, c_t, mask, _, dih_t, lbl = (t.to(dev) for t in tgt)
        with torch.set_grad_enabled(train):
            pN, pCA, pC, pS, mu_g, lv_g, mu_l, lv_l = model(emb_in, n_in, ca_in, c_in, dih_in, mask)
            ld = compute_total_loss(pred_N=pN, pred_CA=pCA, pred_C=pC, pred_seq=pS, target_N=n_t, target_CA=ca_t, target_C=c_t,
                                    target_seq_labels=lbl, mask=mask, mu_g=mu_g, lv_g=lv_g, mu_l=mu_l, lv_l=lv_l,
                                    target_dihedrals=dih_t, **W)
            acc = ((pS.argmax(-1) == lbl) & mask.bool()).sum().float() / mask.sum().float()
            if train:
                opt.zero_grad()
                ld["total"].backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10.0)
                opt.step()
        bs = ca_t.size(0)
        for k, src in (("loss", "total"), ("rec", "reconstruction"), ("pair", "pair_distance"), ("klg", "kl_global"),
                       ("kll", "kl_local"), ("dihedral", "dihedral_total"), ("rama", "ramachandran"), ("bond", "bond_length"),
                       ("angle", "bond_angle"), ("seq", "sequence"), ("clash", "clash")):
            tot[k] += ld[src].item() * bs
        tot["seq_acc"] += acc.item() * bs
        n += bs
    return {k: v / n for k, v in tot.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("train", [False, True])
def test_run_epoch_matches_the_reference_loop(train):
    from protein_ensemble_vae_b200 import HierCVAE
    from protein_ensemble_vae_b200.data import collate_packed
    from protein_ensemble_vae_b200.training import run_epoch
    dev = torch.device("cuda")
    torch.manual_seed(0)
    base = HierCVAE(seqemb_dim=32, nlayers=1, dropout=0.0, precision_encoder="fp32", precision_decoder="fp32").to(dev)
    base.encoder.latent.global_attention.dropout = 0.0
    items = _conformers(6, 11, 32)
    batches = [items[0:2], items[2:5], items[5:6]]                      # batch sizes 2 / 3 / 1: the means are size-weighted
    padded = [(_collate_pad(b), _collate_pad(b)) for b in batches]
    results = []
    for variant in ("reference loop", "run_epoch", "run_epoch packed"):
        model = copy.deepcopy(base)
        opt = torch.optim.SGD(model.parameters(), lr=1e-2) if train else None   # plain SGD: update differences stay proportional
        torch.manual_seed(123)                                          # the reparameterisation noise of the encoder
        if variant == "reference loop":
            res = _reference_loop(model, padded, opt, train, dev)
        else:
            loader = padded if variant == "run_epoch" else [(collate_packed(b), collate_packed(b)) for b in batches]
            res = run_epoch(model, loader, opt, dev, W["klw_g"], W["klw_l"], W["w_pair"], W["pair_stride"], train,
                            W["w_dihedral"], W["w_rama"], W["w_bond"], W["w_angle"], W["w_rec"], W["w_seq"], W["w_clash"], 0)
        assert list(res) == REF_KEYS and all(np.isfinite(v) for v in res.values())
        results.append((res, [p.detach().clone() for p in model.parameters()]))
    ref, ref_params = results[0]
    for (res, params), tol in zip(results[1:], (1e-5, 1e-3)):           # packed: centring re-done on the device in float32
        for k in REF_KEYS:
            assert abs(res[k] - ref[k]) <= tol * max(abs(ref[k]), 1e-3), (k, res[k], ref[k])
        if train:
            moved = max(float((p - q.detach()).abs().max()) for p, q in zip(ref_params, base.parameters()))
            assert moved > 1e-5                                         # three optimizer steps happened
            if tol == 1e-5:
                worst = max(float((p - q).abs().max()) for p, q in zip(params, ref_params))
                assert worst < 1e-5, worst


@pytest.mark.gpu
def test_run_epoch_raises_on_collapse_and_on_an_empty_loader():
    from protein_ensemble_vae_b200 import HierCVAE
    from protein_ensemble_vae_b200.training import run_epoch
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = HierCVAE(seqemb_dim=32, nlayers=1, dropout=0.0).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    args = (W["klw_g"], W["klw_l"], W["w_pair"], W["pair_stride"], True, W["w_dihedral"], W["w_rama"], W["w_bond"],
            W["w_angle"], W["w_rec"], W["w_seq"], W["w_clash"], 3)
    with pytest.raises(ValueError, match="no batch"):
        run_epoch(model, [], opt, dev, *args)
    b = _collate_pad(_conformers(2, 5, 32))
    bad = tuple(t.clone() for t in b)
    bad[1][0, 0, 0] = float("nan")                                      # a NaN target coordinate poisons the total
    with pytest.raises(ValueError, match="NaN"):
        run_epoch(model, [(b, bad)], opt, dev, *args)
    with pytest.raises(ValueError, match="by batch 0"):
        run_epoch(model, [(b, bad), (b, b)], opt, dev, *args, check_finite_every=1)
