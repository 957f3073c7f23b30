"""Case tables and seeded input builders shared by ``make_golden.py`` (which needs the
reference) and the tests (which must not: ``/root/reference`` is absent on the GPU box)."""
from __future__ import annotations

import numpy as np

import synth


def proj_vectors(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(shape[1]), rng.standard_normal(shape[0])



LAYER_CASES = {
    # tag: (node_dim/hidden, lengths of the packed band graph, W, param seed, data seed, graph kind)
    "h32_band": (32, (9, 5), 3, 11, 12, "band"),
    "h32_random": (32, (12,), 0, 13, 14, "random"),
    "h256_band": (256, (24, 16), 5, 15, 16, "band"),
}



DECODER_CASES = {
    # tag: (z_g, z_l, hidden, layers, W, B, L, mask kind, param seed, data seed)
    "small": (12, 6, 32, 2, 3, 4, 10, "empty_row", 21, 22),
    "h256_gaps": (64, 32, 256, 2, 8, 3, 48, "gaps", 23, 24),
    "refdims": (512, 256, 256, 2, 40, 2, 100, "ragged", 25, 26),
    "nomask": (12, 6, 32, 1, 20, 2, 7, "none", 27, 28),
}


def decoder_inputs(case):
    z_g, z_l, H, nl, W, B, L, mkind, pseed, dseed = case
    rng = np.random.default_rng(dseed)
    zg = synth.f32(rng.standard_normal((B, z_g)))
    zl = synth.f32(rng.standard_normal((B, L, z_l)))
    mask = None if mkind == "none" else synth.make_masks(B, L, dseed + 1, mkind)
    coef = [synth.f32(rng.standard_normal((B, L, 3))) for _ in range(3)]
    coef.append(synth.f32(rng.standard_normal((B, L, 20))))
    return zg, zl, mask, coef


# ------------------------------------------------------------------ BASELINE-config shapes (round 2)
# The shapes BASELINE.json's configs name, at a batch the float64 reference finishes in a minute on CPU.
BIG_DECODER_CASES = {
    # tag: (z_g, z_l, hidden, layers, W, B, L, mask kind, param seed, data seed, dropout p)
    "config2": (512, 256, 256, 6, 40, 2, 256, "ragged", 51, 52, 0.0),   # configs[1]: L=256, 6 layers
    "config14": (512, 256, 256, 8, 40, 2, 100, "full", 53, 54, 0.0),    # configs[0]/[3]: ResidueDecoder (8 layers), L=100
    "ragged3": (64, 32, 256, 3, 40, 3, 512, "mixed", 55, 56, 0.0),      # configs[2]: L in {512, 64, 300 with a gap}
    "dropout": (64, 32, 256, 2, 40, 3, 60, "ragged", 57, 60, 0.1),      # train mode, injected dropout masks (data seed chosen so that no ReLU argument is within float32 rounding of 0)
}
BIG_WRAPPED = ("config14",)        # generated through the reference's ResidueDecoder wrapper (F1)


def big_mask(case):
    z_g, z_l, H, nl, W, B, L, mkind, pseed, dseed, p = case
    if mkind == "mixed":
        m = np.zeros((B, L), np.float32)
        for b, n in enumerate((512, 64, 300)[:B]):
            m[b, :n] = 1
        m[2, 100:105] = 0                                                # interior gap (bridged by the graph)
        return m
    return synth.make_masks(B, L, dseed + 1, mkind)


def big_decoder_inputs(case):
    z_g, z_l, H, nl, W, B, L, mkind, pseed, dseed, p = case
    rng = np.random.default_rng(dseed)
    zg = synth.f32(rng.standard_normal((B, z_g)))
    zl = synth.f32(rng.standard_normal((B, L, z_l)))
    mask = big_mask(case)
    coef = [synth.f32(rng.standard_normal((B, L, 3))) for _ in range(3)]
    coef.append(synth.f32(rng.standard_normal((B, L, 20))))
    return zg, zl, mask, coef


def dropout_keep(dseed, site, b, rows, dim, p):
    """float32 0/1 keep mask of dropout site `site` (call order inside one conformer's forward: latent_to_coords,
    one per EGNN layer, two in sequence_head) for conformer `b`: [rows, dim]."""
    rng = np.random.default_rng(dseed * 1000 + site * 17 + b)
    return (rng.random((rows, dim)) >= p).astype(np.float32)


def flat2d(g):
    """Gradients with more than two axes are projected as [prod(leading), last]."""
    g = np.asarray(g)
    return g.reshape(-1, g.shape[-1]) if g.ndim > 2 else g


LOSS_WEIGHTS = dict(klw_g=1.0, klw_l=0.5, w_pair=10.0, w_dihedral=20.0, w_rama=400.0, w_bond=500.0,
                    w_angle=500.0, w_rec=10.0, w_seq=50.0, w_clash=300.0)   # models/vae.py:39-50
LOSS_CASES = {
    # tag: (B, L, mask kind, seed, coordinate scale, pair strides)
    "walk": (4, 24, "gaps", 31, 1.0, (1, 4, 8)),
    "compact": (3, 17, "ragged", 33, 0.35, (4,)),
    "full": (2, 33, "full", 35, 0.6, (8,)),
}


def loss_inputs(case):
    B, L, mkind, seed, scale, _ = case
    rng = np.random.default_rng(seed)
    tn, tca, tc = (a * np.float32(scale) for a in synth.make_backbone(B, L, seed + 1))
    pn, pca, pc = (synth.f32(a + 0.5 * scale * rng.standard_normal(a.shape)) for a in (tn, tca, tc))
    d = dict(pred_N=pn, pred_CA=pca, pred_C=pc, target_N=tn, target_CA=tca, target_C=tc,
             pred_seq=synth.f32(rng.standard_normal((B, L, 20)) * 2.0),
             labels=rng.integers(0, 20, (B, L)).astype(np.int64),
             mask=synth.make_masks(B, L, seed + 2, mkind),
             mu_g=synth.f32(rng.standard_normal((B, 16))), lv_g=synth.f32(0.3 * rng.standard_normal((B, 16))),
             mu_l=synth.f32(rng.standard_normal((B, L, 8))), lv_l=synth.f32(0.3 * rng.standard_normal((B, L, 8))))
    return d


GRAD_INPUTS = ("pred_N", "pred_CA", "pred_C", "pred_seq", "mu_g", "lv_g", "mu_l", "lv_l")



def kabsch_inputs():
    rng = np.random.default_rng(41)
    S, L = 8, 30
    a = synth.f32(np.cumsum(rng.standard_normal((S, L, 3)) * 2.2, axis=1))
    q, _ = np.linalg.qr(rng.standard_normal((S, 3, 3)))
    q[:, :, 0] *= np.sign(np.linalg.det(q))[:, None]              # proper rotations
    b = np.einsum("slk,sjk->slj", a, q) + 3.0 * rng.standard_normal((S, 1, 3))
    b = b + np.array([0.05, 0.1, 0.3, 1.0, 2.0, 0.0, 0.2, 5.0])[:, None, None] * rng.standard_normal((S, L, 3))
    b[6] = b[6] * np.array([1.0, 1.0, -1.0])                     # a reflected copy: d = -1 branch
    b = synth.f32(b)
    mask = synth.make_masks(S, L, 43, "gaps")
    mask[5] = 0                                                   # empty selection -> 0.0
    return a, b, mask




def metrics_inputs():
    """Structure pairs and an ensemble for the evaluation metrics (scripts/validation_metrics.py): S pairs of length L with
    growing noise (one rigidly moved, one reflected), float32; masks with gaps."""
    rng = np.random.default_rng(71)
    S, L = 6, 48
    true = np.cumsum(rng.standard_normal((S, L, 3)) * 2.2, axis=1)
    q, _ = np.linalg.qr(rng.standard_normal((S, 3, 3)))
    q[:, :, 0] *= np.sign(np.linalg.det(q))[:, None]
    pred = np.einsum("slk,sjk->slj", true, q) + 4.0 * rng.standard_normal((S, 1, 3))
    pred = pred + np.array([0.0, 0.2, 0.6, 1.5, 3.0, 8.0])[:, None, None] * rng.standard_normal((S, L, 3))
    pred[4] = pred[4] * np.array([1.0, 1.0, -1.0])
    mask = synth.make_masks(S, L, 73, "gaps")
    ens = true[0][None] + 0.8 * rng.standard_normal((7, L, 3)) * np.linspace(0.2, 2.0, L)[None, :, None]
    return synth.f32(pred), synth.f32(true), mask, synth.f32(ens)


def data_conformers():
    """Raw conformer dicts as ``EnsembleDataset._load`` builds them (models/data.py:120-131): lengths 37 / 64 / 5 / 50, one
    with interior mask gaps, one fully masked, uncentred coordinates, 24-wide sequence embeddings."""
    rng = np.random.default_rng(81)
    aa = "ARNDCQEGHILKMFPSTWYV"
    confs = []
    for i, L in enumerate((37, 64, 5, 50)):
        ca = np.cumsum(rng.standard_normal((L, 3)) * 2.2, axis=0) + 30.0 * rng.standard_normal(3)
        mask = np.ones(L)
        if i == 1:
            mask[10:15] = 0
        if i == 2:
            mask[:] = 0
        confs.append({"n": synth.f32(ca + 0.8 * rng.standard_normal((L, 3))), "ca": synth.f32(ca),
                      "c": synth.f32(ca + 0.8 * rng.standard_normal((L, 3))), "mask": synth.f32(mask),
                      "seq_emb": synth.f32(rng.standard_normal((L, 24))), "dihedrals": synth.f32(rng.uniform(-1, 1, (L, 6))),
                      "sequence": "".join(aa[k] for k in rng.integers(0, 20, L)) + ("X" if i == 0 else "")})
    confs[0]["sequence"] = "X" + confs[0]["sequence"][1:37]        # unknown letter -> label 0 (:183)
    return confs


# --------------------------------------------------------------------------------------------- encoder (models/encoder.py)
ENCODER_CASES = {
    # tag: (seqemb_dim, nlayers, B, L, mask kind, param seed, data seed)
    "enc2_gaps": (256, 2, 3, 48, "gaps", 91, 92),
    "enc6_ragged": (1280, 6, 2, 64, "ragged", 93, 94),
}
ENC_NORM_WEIGHTS = ("coord_norm.weight", "dihedral_norm.weight", "feature_fusion.1.weight", "norm1.weight", "norm2.weight",
                    "enc.ln.weight")


def encoder_params(shapes, seed):
    """name -> float32 ndarray for every parameter of ProteinEncoder (``shapes``: name -> shape, the ``pe`` buffer excluded)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        g = rng.standard_normal(shp)
        if len(shp) == 2:
            g = g / np.sqrt(shp[1])
        elif name.endswith(ENC_NORM_WEIGHTS):
            g = 1.0 + 0.1 * g
        elif name.endswith("geom_res_scale"):
            g = np.asarray(0.3)
        else:
            g = 0.1 * g
        out[name] = synth.f32(g)
    return out


def encoder_inputs(case):
    """(seq_emb, n, ca, c, dihedrals, mask, coefficients of the scalar test loss over mu_g, lv_g, mu_l, lv_l)."""
    sd, nl, B, L, mkind, pseed, dseed = case
    rng = np.random.default_rng(dseed)
    ca = np.cumsum(rng.standard_normal((B, L, 3)) * 2.2, axis=1)
    ca = ca - ca.mean(1, keepdims=True)
    n, c = ca + 0.8 * rng.standard_normal((B, L, 3)), ca + 0.8 * rng.standard_normal((B, L, 3))
    ang = rng.uniform(-np.pi, np.pi, (B, L, 3))
    dih = np.stack([np.sin(ang[..., 0]), np.cos(ang[..., 0]), np.sin(ang[..., 1]), np.cos(ang[..., 1]), np.sin(ang[..., 2]),
                    np.cos(ang[..., 2])], -1)
    emb = rng.standard_normal((B, L, sd))
    mask = synth.make_masks(B, L, dseed + 1, mkind)
    coef = [rng.standard_normal((B, 512)), rng.standard_normal((B, 512)), rng.standard_normal((B, L, 256)) * mask[..., None],
            rng.standard_normal((B, L, 256)) * mask[..., None]]
    return [synth.f32(a) for a in (emb, n, ca, c, dih)], mask, [synth.f32(a) for a in coef]


# --------------------------------------------------------------------------------------------- HierCVAE (models/model.py)
# (seqemb_dim, encoder layers, B, L, mask kind, param seed, data seed); the decoder half is the wrapper's 8 layers / 256 / W=40
HIERCVAE_CASE = (256, 2, 2, 48, "gaps", 95, 96)


def hiercvae_params(shapes, seed):
    """name -> float32 ndarray for every parameter of HierCVAE (``shapes``: name -> shape, the ``pe`` buffer excluded):
    matrices ~ N(0, 1/fan_in), LayerNorm weights ~ 1 + 0.1 N(0,1), biases / vectors ~ 0.1 N(0,1)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        g = rng.standard_normal(shp)
        if len(shp) >= 2:
            g = g / np.sqrt(shp[-1])
        elif name.endswith("geom_res_scale"):
            g = np.asarray(0.3)
        elif name.endswith(".weight"):                     # every 1-D weight of the model is a LayerNorm scale
            g = 1.0 + 0.1 * g
        else:
            g = 0.1 * g
        out[name] = synth.f32(g)
    return out


def hiercvae_inputs(case=HIERCVAE_CASE):
    """(seq_emb, n, ca, c, dihedrals), mask, (eps_g, eps_l) reparameterisation noise, coefficients of the scalar test loss
    over the eight outputs of ``HierCVAE.forward``."""
    sd, nl, B, L, mkind, pseed, dseed = case
    xs, mask, _ = encoder_inputs(case)
    rng = np.random.default_rng(dseed + 7)
    eps = [synth.f32(rng.standard_normal((B, 512))), synth.f32(rng.standard_normal((B, L, 256)))]
    m3 = mask[..., None]
    coef = [rng.standard_normal((B, L, 3)) * m3, rng.standard_normal((B, L, 3)) * m3, rng.standard_normal((B, L, 3)) * m3,
            rng.standard_normal((B, L, 20)) * m3, rng.standard_normal((B, 512)), rng.standard_normal((B, 512)),
            rng.standard_normal((B, L, 256)) * m3, rng.standard_normal((B, L, 256)) * m3]
    return xs, mask, eps, [synth.f32(a) for a in coef]


def pdb_inputs():
    """Three models of a 14-residue backbone (float32) with an interior gap and a masked first residue, coordinates that
    exercise the %8.3f field (negative zero, a rounding tie j/16, four integer digits), a sequence with an unknown letter."""
    rng = np.random.default_rng(101)
    S, L = 3, 14
    ca = np.cumsum(rng.standard_normal((S, L, 3)) * 2.2, axis=1)
    n, c = ca + 0.8 * rng.standard_normal((S, L, 3)), ca + 0.8 * rng.standard_normal((S, L, 3))
    n[0, 3] = [-0.0004, 0.0625, 1234.5678]
    ca[1, 5] = [-99.9996, 0.1875, -0.0]
    c[2, 7] = [9999.9, -12.3125, 0.0005]
    mask = np.ones(L, np.float32)
    mask[0] = 0
    mask[8:10] = 0
    return synth.f32(n), synth.f32(ca), synth.f32(c), mask, "MKXLVAGGHWYRTS"
