"""``HierCVAE`` (models/model.py:15-116; the caller of the hot path named in SURVEY.md 8b): checkpoint layout, the oracle
composition and the device module against outputs of the reference's own ``HierCVAE`` (tests/golden/hiercvae.npz, produced by
``make_golden.py::gen_hiercvae`` in float64 with the reparameterisation noise injected)."""
import json
import os

import numpy as np
import pytest
import torch

import cases
from conftest import rel_err

G = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("N", "CA", "C", "logits", "mu_g", "lv_g", "mu_l", "lv_l")


def _model(sd_dim, nl, **kw):
    from protein_ensemble_vae_b200 import HierCVAE
    vae = HierCVAE(seqemb_dim=sd_dim, nlayers=nl, dropout=0.0, **kw)
    return vae, {k: tuple(v.shape) for k, v in vae.state_dict().items() if k != "encoder.enc.pe.pe"}


def test_hiercvae_state_dict_is_the_reference_layout():
    """Names AND shapes of every entry of the reference's ``HierCVAE(seqemb_dim=1280).state_dict()`` (what
    ``models/training.py:456`` saves): a reference checkpoint loads with ``strict=True``."""
    from protein_ensemble_vae_b200 import HierCVAE
    with open(os.path.join(G, "hiercvae_keys.json")) as f:
        ref = {k: tuple(v) for k, v in json.load(f).items()}
    vae = HierCVAE(seqemb_dim=1280)
    mine = {k: tuple(v.shape) for k, v in vae.state_dict().items()}
    assert mine == ref
    assert sum(p.numel() for p in vae.parameters()) == 20423592
    vae.load_state_dict({k: torch.zeros(s) for k, s in ref.items()}, strict=True)


def test_hiercvae_is_cuda_only_and_keeps_the_reference_signatures():
    import inspect
    from protein_ensemble_vae_b200 import HierCVAE
    sig = inspect.signature(HierCVAE.__init__)
    assert list(sig.parameters)[1:12] == ["seqemb_dim", "d_model", "nhead", "ff", "nlayers", "z_g", "z_l", "dropout",
                                          "equivariant", "decoder_hidden", "use_dihedrals"]
    assert list(inspect.signature(HierCVAE.forward).parameters)[1:7] == ["seqemb_or_none", "n_coords", "ca_coords", "c_coords",
                                                                         "dihedrals", "mask"]
    assert list(inspect.signature(HierCVAE.sample).parameters)[1:4] == ["mask", "seqemb_or_none", "num_samples"]
    vae, _ = _model(256, 1)
    z = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError):
        vae(torch.zeros(1, 4, 256), z, z, z, torch.zeros(1, 4, 6), torch.ones(1, 4))       # no CPU fallback
    with pytest.raises(RuntimeError):
        vae.sample(torch.ones(1, 4))


def test_oracle_composition_matches_reference_hiercvae():
    """encoder oracle -> z = mu + eps * exp(lv / 2) (models/encoder.py:231-236) -> decoder oracle, against the reference's
    end-to-end outputs: pins the oracle pair on the composed module, not only on the two halves."""
    from oracle import egnn_oracle
    from oracle import encoder_oracle as eo
    gold = np.load(os.path.join(G, "hiercvae.npz"))
    pe = np.load(os.path.join(G, "encoders.npz"))["pe"]
    case = cases.HIERCVAE_CASE
    _, shapes = _model(case[0], case[1])
    sd = cases.hiercvae_params(shapes, case[5])
    xs, mask, eps, _ = cases.hiercvae_inputs(case)
    enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    dec_sd = {k[len("decoder.decoder.decoder."):]: torch.tensor(v, dtype=torch.float64) for k, v in sd.items()
              if k.startswith("decoder.")}
    _, mu_g, lv_g, mu_l, lv_l = eo.encoder(enc_sd, *[a.astype(np.float64) for a in xs], mask, pe=pe)
    mb = mask.astype(bool)
    for got, key in ((mu_g, "mu_g"), (lv_g, "lv_g")):
        assert np.abs(got - gold[f"vae.{key}"]).max() < 1e-10
    for got, key in ((mu_l, "mu_l"), (lv_l, "lv_l")):
        assert np.abs(got[mb] - gold[f"vae.{key}"][mb]).max() < 1e-10
    z_g = mu_g + eps[0] * np.exp(0.5 * lv_g)
    z_l = mu_l + eps[1] * np.exp(0.5 * lv_l)
    outs = egnn_oracle.egnn_decoder(dec_sd, torch.tensor(z_g), torch.tensor(z_l), torch.tensor(mask, dtype=torch.float64),
                                    max_neighbors=40)
    for name, o in zip(NAMES[:4], outs):
        assert rel_err(o, gold[f"vae.{name}"]) < 1e-9, name


def _run_device(precision_encoder, precision_decoder):
    case = cases.HIERCVAE_CASE
    vae, shapes = _model(case[0], case[1], precision_encoder=precision_encoder, precision_decoder=precision_decoder)
    sd = cases.hiercvae_params(shapes, case[5])
    vae.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}, strict=False)
    vae = vae.cuda().train()
    vae.encoder.latent.global_attention.dropout = 0.0
    xs, mask, eps, coef = cases.hiercvae_inputs(case)
    xs = [torch.tensor(a, device="cuda") for a in xs]
    m = torch.tensor(mask, device="cuda")
    res = vae(*xs, m, eps_g=torch.tensor(eps[0], device="cuda"), eps_l=torch.tensor(eps[1], device="cuda"))
    sum((r * torch.tensor(c, device="cuda")).sum() for r, c in zip(res, coef)).backward()
    grads = {k: p.grad for k, p in vae.named_parameters() if p.grad is not None}
    return vae, mask, res, grads


def _output_errors(mask, res, gold):
    mb = torch.tensor(mask.astype(bool))
    errs = {}
    for name, r in zip(NAMES, res):
        got, ref = r.detach().cpu().double(), torch.tensor(gold[f"vae.{name}"])
        assert got.shape == ref.shape
        if got.dim() == 3:
            assert float(got[~mb].abs().max()) == 0.0                                    # exact zeros at padding
            got, ref = got[mb], ref[mb]
        errs[name] = rel_err(got, ref)
    return errs


@pytest.mark.gpu
def test_hiercvae_fp32_matches_reference():
    """Exact paths of both halves (3xTF32 linears, fp32 edge MLP): outputs 2e-5 (measured 0.9 - 5.2e-6), every parameter
    gradient 2e-4 relative L2 against the reference's autograd (measured worst 5.2e-5, ``layers.5.phi_x.2.bias``)."""
    import test_gpu_parity_big as tb
    gold = np.load(os.path.join(G, "hiercvae.npz"))
    vae, mask, res, grads = _run_device("fp32", "fp32")
    errs = _output_errors(mask, res, gold)
    print("hiercvae fp32 outputs", {k: f"{v:.1e}" for k, v in errs.items()})
    assert max(errs.values()) < 2e-5, errs
    gerrs = tb._grad_errors(grads, gold, "vae")
    assert {k.split(".", 2)[2] for k in gold.files if k.startswith("vae.g")} == set(gerrs)
    worst = max(gerrs.items(), key=lambda kv: kv[1][1])
    print("hiercvae fp32 grad l2 worst", worst)
    assert worst[1][1] < 2e-4, worst


@pytest.mark.gpu
def test_hiercvae_default_precision_and_sample():
    """The default (throughput) arithmetic -- TF32 encoder, bf16 edge MLP: latents 5e-3 (measured 2.0 - 2.5e-3), CA 1e-2
    (1.6e-3), N / C / logits 4e-2 (1.3e-2 / 9.5e-3 / 8.7e-3; the bf16 yardstick of tests/bf16_yardstick.py); every gradient finite.  ``sample``: shapes, ``repeat_interleave`` order of the
    masks (models/model.py:111), exact zeros at padding, reproducible under a generator."""
    gold = np.load(os.path.join(G, "hiercvae.npz"))
    vae, mask, res, grads = _run_device("tf32", "bf16")
    errs = _output_errors(mask, res, gold)
    print("hiercvae tf32/bf16 outputs", {k: f"{v:.1e}" for k, v in errs.items()})
    assert max(errs[k] for k in ("mu_g", "lv_g", "mu_l", "lv_l")) < 5e-3, errs
    assert errs["CA"] < 1e-2 and max(errs[k] for k in ("N", "C", "logits")) < 4e-2, errs
    assert all(torch.isfinite(g).all() for g in grads.values())
    vae.eval()
    m = torch.tensor(mask, device="cuda")
    S = 3
    with torch.no_grad():
        outs = vae.sample(m, num_samples=S, generator=torch.Generator(device="cuda").manual_seed(5))
        again = vae.sample(m, num_samples=S, generator=torch.Generator(device="cuda").manual_seed(5))
    B, L = mask.shape
    assert [tuple(o.shape) for o in outs] == [(B * S, L, 3)] * 3 + [(B * S, L, 20)]
    pad = (m.repeat_interleave(S, dim=0) == 0)
    for o, a in zip(outs, again):
        assert torch.isfinite(o).all() and torch.equal(o, a)
        assert float(o[pad].abs().max()) == 0.0 and float(o[~pad].abs().max()) > 0.0
    assert not torch.equal(outs[1][0], outs[1][1])                                       # different draws per sample
