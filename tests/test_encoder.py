"""ProteinEncoder: oracle restatement vs the reference's outputs (CPU), device path vs the reference's
outputs and gradients (GPU; tests/golden/encoders.npz is produced by models/encoder.py in float64)."""
import os

import numpy as np
import pytest
import torch

import cases
from conftest import rel_err

G = os.path.join(os.path.dirname(__file__), "golden")


def _shapes(sd_dim, nl):
    from protein_ensemble_vae_b200.encoder import ProteinEncoder
    enc = ProteinEncoder(seqemb_dim=sd_dim, nlayers=nl, dropout=0.0)
    return enc, {k: tuple(v.shape) for k, v in enc.state_dict().items() if k != "enc.pe.pe"}


@pytest.mark.parametrize("tag", list(cases.ENCODER_CASES))
def test_encoder_oracle_matches_reference(tag):
    from oracle import encoder_oracle as eo
    gold = np.load(os.path.join(G, "encoders.npz"))
    case = cases.ENCODER_CASES[tag]
    _, shapes = _shapes(case[0], case[1])
    sd = cases.encoder_params(shapes, case[5])
    xs, mask, _ = cases.encoder_inputs(case)
    H, mu_g, lv_g, mu_l, lv_l = eo.encoder(sd, *[a.astype(np.float64) for a in xs], mask, pe=gold["pe"])
    mb = mask.astype(bool)
    assert np.abs(H[mb] - gold[f"{tag}.H"][mb]).max() < 2e-7 * np.abs(H[mb]).max()   # H is stored as float32
    for got, key in ((mu_g, "mu_g"), (lv_g, "lv_g")):
        assert np.abs(got - gold[f"{tag}.{key}"]).max() < 1e-11
    for got, key in ((mu_l, "mu_l"), (lv_l, "lv_l")):
        assert np.abs(got[mb] - gold[f"{tag}.{key}"][mb]).max() < 1e-11


def test_encoder_state_dict_keys_match_reference_layout():
    enc, shapes = _shapes(1280, 6)
    assert sum(int(np.prod(s)) for s in shapes.values()) == 18068353 - 4096 * 512     # reference count minus the PE buffer
    assert "enc.transformer_layers.5.self_attn.in_proj_weight" in shapes and shapes["latent.global_query"] == (1, 1, 512)
    with pytest.raises(RuntimeError):
        z = torch.zeros(1, 4, 3)
        enc(torch.zeros(1, 4, 1280), z, z, z, torch.zeros(1, 4, 6), torch.ones(1, 4))    # CUDA only, no CPU fallback


def _run(tag, precision, yardstick=False):
    """``yardstick``: the same module with every linear / attention on torch's library kernels with TF32 allowed -- what an
    ideal single-pass TF32 implementation gives against the float64 reference (gradients through the ReLUs and LayerNorms of
    the encoder are ill-conditioned, as in the decoder heads: tests/bf16_yardstick.py)."""
    import test_gpu_parity_big as tb
    from protein_ensemble_vae_b200 import attention, tc_linear
    if yardstick:
        saved = (tc_linear.supported, attention.PackedSelfAttention, torch.backends.cuda.matmul.allow_tf32)
        tc_linear.supported = lambda x, W: False
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            attention.PackedSelfAttention = None
            return _run_inner(tag, "fp32", tb)                         # "fp32" selects the torch attention path
        finally:
            tc_linear.supported, attention.PackedSelfAttention, torch.backends.cuda.matmul.allow_tf32 = saved
    return _run_inner(tag, precision, tb)


def _run_inner(tag, precision, tb):
    case = cases.ENCODER_CASES[tag]
    enc, shapes = _shapes(case[0], case[1])
    enc.enc.precision = enc.latent.precision = precision
    sd = cases.encoder_params(shapes, case[5])
    enc.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}, strict=False)
    enc = enc.cuda().train()
    enc.latent.global_attention.dropout = 0.0
    xs, mask, coef = cases.encoder_inputs(case)
    xs = [torch.tensor(a, device="cuda") for a in xs]
    m = torch.tensor(mask, device="cuda")
    z_g, z_l, mu_g, lv_g, mu_l, lv_l = enc(*xs, m, eps_g=torch.zeros(mask.shape[0], 512, device="cuda"),
                                           eps_l=torch.zeros(*mask.shape, 256, device="cuda"))
    res = (mu_g, lv_g, mu_l, lv_l)
    sum((r * torch.tensor(c, device="cuda")).sum() for r, c in zip(res, coef)).backward()
    grads = {k: p.grad for k, p in enc.named_parameters() if p.grad is not None}
    return mask, res, (z_g, z_l), grads, tb


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol_out", [("fp32", 1e-5), ("tf32", 5e-3)])
@pytest.mark.parametrize("tag", list(cases.ENCODER_CASES))
def test_encoder_matches_reference(tag, precision, tol_out):
    """fp32 path (3xTF32 linears): outputs 1e-5, parameter gradients 1e-4.  TF32 path (own GEMM + attention kernels):
    outputs 5e-3 (measured 1.5 - 2.7e-3); gradients within 6x (worst parameter: a single flipped unit) / 3x (median over parameters) of the library-TF32 yardstick ( measured: mine
    worst 0.10 - 0.35 / median 0.03, yardstick worst 0.15 - 0.20 / median 0.014 - 0.026: single-pass TF32 flips ~1e-4 of the
    ReLU units of the feed-forward blocks, which moves weight gradients by percents whoever does the rounding)."""
    gold = np.load(os.path.join(G, "encoders.npz"))
    mask, res, (z_g, z_l), grads, tb = _run(tag, precision)
    mb = torch.tensor(mask.astype(bool))
    errs = {}
    for name, r in zip(("mu_g", "lv_g", "mu_l", "lv_l"), res):
        got, ref = r.detach().cpu().double(), torch.tensor(gold[f"{tag}.{name}"])
        if got.dim() == 3:
            assert (not (~mb).any()) or float(got[~mb].abs().max()) == 0.0              # exact zeros at padding
            got, ref = got[mb], ref[mb]
        errs[name] = rel_err(got, ref)
    assert max(errs.values()) < tol_out, errs
    assert torch.equal(z_g, res[0]) and torch.equal(z_l, res[2])                          # eps = 0: z = mu
    gerrs = tb._grad_errors(grads, gold, tag)
    assert {k.split(".", 2)[2] for k in gold.files if k.startswith(tag + ".g")} == set(gerrs)
    l2 = sorted(v[1] for v in gerrs.values())
    worst, median = l2[-1], l2[len(l2) // 2]
    if precision == "fp32":
        assert worst < 1e-4, max(gerrs.items(), key=lambda kv: kv[1][1])
    else:
        yard = sorted(v[1] for v in tb._grad_errors(_run(tag, precision, yardstick=True)[3], gold, tag).values())
        assert worst < max(2e-2, 6.0 * yard[-1]) and median < max(1e-2, 3.0 * yard[len(yard) // 2]), (worst, median, yard[-1], yard[len(yard) // 2])
        print(tag, "yardstick worst %.1e median %.1e" % (yard[-1], yard[len(yard) // 2]))
    print(tag, precision, "outputs", {k: f"{v:.1e}" for k, v in errs.items()}, "grad l2 worst %.1e median %.1e" % (worst, median))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_encoder_edge_cases(precision):
    """mask=None equals an all-ones mask; a conformer without valid residues and a single-residue conformer run (the reference
    returns NaN for the former: softmax over an empty key set) and do not disturb the other conformers of the batch."""
    from protein_ensemble_vae_b200.encoder import ProteinEncoder
    torch.manual_seed(7)
    enc = ProteinEncoder(seqemb_dim=256, nlayers=2, dropout=0.0, precision=precision).cuda().eval()
    B, L = 4, 40
    g = torch.Generator(device="cuda").manual_seed(8)
    emb = torch.randn(B, L, 256, device="cuda", generator=g)
    ca = torch.cumsum(torch.randn(B, L, 3, device="cuda", generator=g) * 2.2, 1)
    n, c = ca + 0.5, ca - 0.5
    dih = torch.randn(B, L, 6, device="cuda", generator=g).clamp(-1, 1)
    eg, el = torch.zeros(B, 512, device="cuda"), torch.zeros(B, L, 256, device="cuda")
    with torch.no_grad():
        full = enc(emb, n, ca, c, dih, None, eps_g=eg, eps_l=el)
        ones = enc(emb, n, ca, c, dih, torch.ones(B, L, device="cuda"), eps_g=eg, eps_l=el)
        for a, b in zip(full, ones):
            assert torch.equal(a, b)
        mask = torch.ones(B, L, device="cuda")
        mask[1] = 0                                           # empty conformer
        mask[2, 1:] = 0                                       # one residue
        mask[3, 25:] = 0
        out = enc(emb, n, ca, c, dih, mask, eps_g=eg, eps_l=el)
        assert all(torch.isfinite(t).all() for t in out)
        assert float(out[4][1].abs().max()) == 0.0 and float(out[4][2, 1:].abs().max()) == 0.0      # mu_l zeros at padding
        # conformer 0 (full length) is unaffected by what happens to its batch neighbours
        tol = 1e-5 if precision == "fp32" else 2e-3
        assert float((out[4][0] - full[4][0]).abs().max()) < tol * float(full[4][0].abs().max())
        assert float((out[2][0] - full[2][0]).abs().max()) < tol * float(full[2][0].abs().max())
        # conformer 3 alone (cropped to its 25 residues) gives the same latents as inside the ragged batch
        solo = enc(emb[3:4, :25], n[3:4, :25], ca[3:4, :25], c[3:4, :25], dih[3:4, :25], None, eps_g=eg[:1], eps_l=el[:1, :25])
        assert float((solo[4][0] - out[4][3, :25]).abs().max()) < tol * float(solo[4].abs().max())
        assert float((solo[2][0] - out[2][3]).abs().max()) < tol * float(solo[2].abs().max())
