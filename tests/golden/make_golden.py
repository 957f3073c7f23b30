#!/usr/bin/env python3
"""Generate the golden fixtures by RUNNING THE REFERENCE in this container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference (mohit03031999/Protein-Ensemble-VAE) has no tests and no golden
vectors, so parity is pinned on its own outputs: its modules are imported from
``/root/reference/models`` (flat imports, as its scripts do), cast to float64,
put in ``eval()`` and fed the deterministic inputs of ``synth.py``.  Only
inputs that cannot be regenerated from a seed, and the reference's outputs /
autograd gradients, are stored (``*.npz`` next to this file).  Nothing here is
needed at test time on the GPU box -- the fixtures travel, the reference does not.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PEV_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REF, "models"))

import synth  # noqa: E402
from cases import *  # noqa: E402,F401,F403
import cases  # noqa: E402
import en_gnn_decoder as ref_dec  # noqa: E402  (reference)
import losses as ref_losses  # noqa: E402  (reference)

torch.set_default_dtype(torch.float64)
T = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))  # noqa: E731


def lift_function(path, name, namespace):
    """exec one top-level function of a reference file whose module cannot be imported here."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            code = compile(ast.Module([node], []), path, "exec")
            exec(code, namespace)
            return namespace[name]
    raise KeyError(name)


def pack_grads(named_grads, out, tag, seed=1234):
    """Small grads are stored whole, matrices as two random projections (G @ r1, r2 @ G)."""
    for i, (name, g) in enumerate(sorted(named_grads.items())):
        g = g.detach().numpy()
        if g.ndim == 2 and g.size > 4096:
            r1, r2 = proj_vectors(g.shape, seed + i)
            out[f"{tag}.gproj1.{name}"] = g @ r1
            out[f"{tag}.gproj2.{name}"] = r2 @ g
        else:
            out[f"{tag}.grad.{name}"] = g


# --------------------------------------------------------------------------- edges
def gen_edges():
    out = {}
    ei = ref_dec.EGNNDecoder.build_edge_index(5, torch.device("cpu"), 2)
    out["L5_W2"] = ei.numpy()
    for L in (64, 100, 256, 512, 1024):
        ei = ref_dec.EGNNDecoder.build_edge_index(L, torch.device("cpu"), 40)
        out[f"E_L{L}_W40"] = np.array(ei.shape[1])
        if L in (64, 100):
            out[f"L{L}_W40"] = ei.numpy().astype(np.int32)
            out[f"deg_L{L}_W40"] = ref_dec.EGNNDecoder.degrees(ei, L).numpy()
    out["L7_W0_fallback"] = ref_dec.EGNNDecoder.build_edge_index(7, torch.device("cpu"), 0).numpy()
    out["L6_W9_dense"] = ref_dec.EGNNDecoder.build_edge_index(6, torch.device("cpu"), 9).numpy()
    out["L2_W1"] = ref_dec.EGNNDecoder.build_edge_index(2, torch.device("cpu"), 1).numpy()
    np.savez_compressed(os.path.join(HERE, "edges.npz"), **out)


# --------------------------------------------------------------------------- one layer
def layer_graph(lengths, W, kind, seed):
    if kind == "band":
        rows, cols, off = [], [], 0
        for Lb in lengths:
            ei = ref_dec.EGNNDecoder.build_edge_index(Lb, torch.device("cpu"), W)
            rows.append(ei[0] + off), cols.append(ei[1] + off)
            off += Lb
        return torch.stack([torch.cat(rows), torch.cat(cols)])
    rng = np.random.default_rng(seed)          # unsorted multigraph with an isolated node
    n = sum(lengths)
    E = 5 * n
    row = rng.integers(0, n - 1, E)
    col = rng.integers(0, n, E)
    return torch.tensor(np.stack([row, col]), dtype=torch.long)


def gen_layers():
    out = {}
    for tag, (H, lengths, W, pseed, dseed, kind) in LAYER_CASES.items():
        n = sum(lengths)
        params = synth.make_params(synth.layer_param_shapes(H, H), pseed)
        layer = ref_dec.EGNLayer(H, H).double().eval()
        layer.load_state_dict({k: T(v) for k, v in params.items()})
        rng = np.random.default_rng(dseed)
        h = T(synth.f32(rng.standard_normal((n, H)))).requires_grad_()
        x = T(synth.f32(rng.standard_normal((n, 3)) * 2.0)).requires_grad_()
        ei = layer_graph(lengths, W, kind, dseed + 100)
        deg = ref_dec.EGNNDecoder.degrees(ei, n)
        dinv = (1.0 / deg.float()).double() if kind == "band" else None
        ch = T(synth.f32(rng.standard_normal((n, H))))
        cx = T(synth.f32(rng.standard_normal((n, 3))))
        h2, x2 = layer(h, x, ei, degree_inv=dinv)
        ((h2 * ch).sum() + (x2 * cx).sum()).backward()
        out[f"{tag}.edge_index"] = ei.numpy()
        out[f"{tag}.h_out"] = h2.detach().numpy()
        out[f"{tag}.x_out"] = x2.detach().numpy()
        grads = {"h": h.grad, "x": x.grad}
        grads.update({k: p.grad for k, p in layer.named_parameters()})
        pack_grads(grads, out, tag)
    np.savez_compressed(os.path.join(HERE, "layers.npz"), **out)


# --------------------------------------------------------------------------- decoder
def gen_decoders():
    out = {}
    for tag, case in DECODER_CASES.items():
        z_g, z_l, H, nl, W, B, L, mkind, pseed, dseed = case
        params = synth.make_params(synth.decoder_param_shapes(z_g, z_l, H, nl), pseed)
        dec = ref_dec.EGNNDecoder(z_g, z_l, hidden_dim=H, num_layers=nl, max_neighbors=W,
                                  dropout=0.0).double().eval()
        dec.load_state_dict({k: T(v) for k, v in params.items()})
        zg, zl, mask, coef = decoder_inputs(case)
        zg_t, zl_t = T(zg).requires_grad_(), T(zl).requires_grad_()
        outs = dec(zg_t, zl_t, mask=None if mask is None else T(mask))
        loss = sum((o * T(c)).sum() for o, c in zip(outs, coef))
        loss.backward()
        for name, o in zip(("N", "CA", "C", "logits"), outs):
            out[f"{tag}.{name}"] = o.detach().numpy()
        grads = {"z_g": zg_t.grad, "z_l": zl_t.grad}
        grads.update({k: p.grad for k, p in dec.named_parameters() if p.grad is not None})
        pack_grads(grads, out, tag)
    np.savez_compressed(os.path.join(HERE, "decoders.npz"), **out)



# --------------------------------------------------------------------------- decoder, BASELINE-config shapes
class InjectedDropout:
    """Replaces ``nn.Dropout.forward`` while the reference runs in train mode: the k-th call inside one forward is
    site ``k % sites`` of conformer ``k // sites`` (the reference loops over conformers, models/en_gnn_decoder.py:216,
    and every conformer visits the same dropout modules in the same order: latent_to_coords :128, one per layer :250,
    two in sequence_head :166/:170); the keep mask comes from ``cases.dropout_keep`` instead of torch's generator."""

    def __init__(self, dseed, sites):
        self.dseed, self.sites, self.k = dseed, sites, 0

    def __enter__(self):
        self.orig = torch.nn.Dropout.forward
        me = self

        def forward(mod, x):
            if not mod.training or mod.p == 0.0:
                return x
            b, site = divmod(me.k, me.sites)
            me.k += 1
            keep = T(cases.dropout_keep(me.dseed, site, b, x.shape[0], x.shape[1], mod.p))
            return x * keep / (1.0 - mod.p)

        torch.nn.Dropout.forward = forward
        return self

    def __exit__(self, *exc):
        torch.nn.Dropout.forward = self.orig
        return False


def pack_grads_big(named_grads, out, tag, seed=4321):
    for i, (name, g) in enumerate(sorted(named_grads.items())):
        g = cases.flat2d(g.detach().numpy())
        if g.ndim == 2 and g.size > 4096:
            r1, r2 = proj_vectors(g.shape, seed + i)
            out[f"{tag}.gproj1.{name}"] = g @ r1
            out[f"{tag}.gproj2.{name}"] = r2 @ g
        else:
            out[f"{tag}.grad.{name}"] = g


def gen_big_decoders():
    out = {}
    for tag, case in cases.BIG_DECODER_CASES.items():
        z_g, z_l, H, nl, W, B, L, mkind, pseed, dseed, pdrop = case
        params = synth.make_params(synth.decoder_param_shapes(z_g, z_l, H, nl), pseed)
        if tag in cases.BIG_WRAPPED:       # the wrapper hard-codes hidden 256 / 8 layers / W 40 (F1)
            assert (H, nl, W) == (256, 8, 40)
            top = ref_dec.ResidueDecoder(z_g, z_l, hidden=64, dropout=pdrop).double()
            dec = top.decoder.decoder
        else:
            top = dec = ref_dec.EGNNDecoder(z_g, z_l, hidden_dim=H, num_layers=nl, max_neighbors=W,
                                            dropout=pdrop).double()
        dec.load_state_dict({k: T(v) for k, v in params.items()})
        top.train(pdrop > 0.0)
        zg, zl, mask, coef = cases.big_decoder_inputs(case)
        zg_t, zl_t = T(zg).requires_grad_(), T(zl).requires_grad_()
        with InjectedDropout(dseed, nl + 3):
            outs = top(zg_t, zl_t, mask=T(mask))
        # second gradient set, "smooth": the CA term alone -- between this loss and the EGNN layers there is no ReLU,
        # so the gradients of the layer parameters are differentiable functions of the forward activations
        (outs[1] * T(coef[1])).sum().backward(retain_graph=True)
        grads = {"z_g": zg_t.grad, "z_l": zl_t.grad}
        grads.update({k: p.grad for k, p in dec.named_parameters() if p.grad is not None})
        pack_grads_big(grads, out, tag + ".sm")
        zg_t.grad = zl_t.grad = None
        dec.zero_grad(set_to_none=True)
        loss = sum((o * T(c)).sum() for o, c in zip(outs, coef))
        loss.backward()
        for name, o in zip(("N", "CA", "C", "logits"), outs):
            out[f"{tag}.{name}"] = o.detach().numpy()
        grads = {"z_g": zg_t.grad, "z_l": zl_t.grad}
        grads.update({k: p.grad for k, p in dec.named_parameters() if p.grad is not None})
        pack_grads_big(grads, out, tag)
        print("big decoder case", tag, "done", flush=True)
    np.savez_compressed(os.path.join(HERE, "decoders_big.npz"), **out)

# --------------------------------------------------------------------------- losses
def gen_losses():
    out = {}
    for tag, case in LOSS_CASES.items():
        d = loss_inputs(case)
        mask = T(d["mask"])
        tgt = [T(d[k]) for k in ("target_N", "target_CA", "target_C")]
        tdih = ref_losses.compute_dihedrals_from_coords(*tgt, mask)
        out[f"{tag}.target_dihedrals"] = tdih.numpy()
        for stride in case[5]:
            leaves = {k: T(d[k]).requires_grad_() for k in GRAD_INPUTS}
            res = ref_losses.compute_total_loss(
                leaves["pred_N"], leaves["pred_CA"], leaves["pred_C"], leaves["pred_seq"],
                tgt[0], tgt[1], tgt[2], torch.tensor(d["labels"]), mask,
                leaves["mu_g"], leaves["lv_g"], leaves["mu_l"], leaves["lv_l"], tdih,
                pair_stride=stride, **LOSS_WEIGHTS)
            res["total"].backward()
            for k, v in res.items():
                out[f"{tag}.s{stride}.{k}"] = v.detach().numpy()
            for k, v in leaves.items():
                out[f"{tag}.s{stride}.grad.{k}"] = v.grad.numpy()
        pdih = ref_losses.compute_dihedrals_from_coords(T(d["pred_N"]), T(d["pred_CA"]), T(d["pred_C"]), mask)
        out[f"{tag}.pred_dihedrals"] = pdih.numpy()
    # ideal backbone known answers
    N, CA, C = (T(a)[None] for a in synth.nerf_backbone(32))
    ones = torch.ones(1, 32)
    dih = ref_losses.compute_dihedrals_from_coords(N, CA, C, ones)
    out["nerf.dihedrals"] = dih.numpy()
    out["nerf.bond_length"] = ref_losses.bond_length_loss(N, CA, C, ones).numpy()
    out["nerf.bond_angle"] = ref_losses.bond_angle_loss(N, CA, C, ones).numpy()
    out["nerf.omega_trans"] = ref_losses.omega_trans_loss(dih, ones).numpy()
    out["nerf.ramachandran"] = ref_losses.ramachandran_loss(dih, ones).numpy()
    out["nerf.clash"] = ref_losses.clash_loss(N, CA, C, ones).numpy()
    out["nerf.pair_self"] = ref_losses.pair_distance_loss(CA, CA, ones, stride=4).numpy()
    out["nerf.pair_scaled"] = ref_losses.pair_distance_loss(1.1 * CA, CA, ones, stride=4).numpy()
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)


# --------------------------------------------------------------------------- Kabsch
def gen_kabsch():
    ns = {"np": np, "torch": torch}
    kabsch_rmsd = lift_function(os.path.join(REF, "generate_ensemble_pdbs.py"), "kabsch_rmsd", ns)
    ns2 = {"np": np}
    kabsch_align = lift_function(os.path.join(REF, "scripts", "validation_metrics.py"), "kabsch_align", ns2)
    a, b, mask = kabsch_inputs()
    compat, correct = [], []
    for s in range(a.shape[0]):
        compat.append(kabsch_rmsd(T(a[s]), T(b[s]), torch.tensor(mask[s])))
        sel = mask[s].astype(bool)
        if sel.sum() == 0:
            correct.append(0.0)
            continue
        x, y = a[s][sel].astype(np.float64), b[s][sel].astype(np.float64)
        al = kabsch_align(x, y)
        correct.append(float(np.sqrt(((al - y) ** 2).sum(-1).mean())))
    np.savez_compressed(os.path.join(HERE, "kabsch.npz"), ref_compat=np.array(compat, np.float64),
                        optimal=np.array(correct, np.float64))


# --------------------------------------------------------------------------- evaluation metrics
def gen_metrics():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_validation_metrics", os.path.join(REF, "scripts", "validation_metrics.py"))
    vm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vm)
    pred, true, mask, ens = cases.metrics_inputs()
    out = {"tm": [], "gdt_ts": [], "gdt_ha": [], "gdt_ts_masked": [], "gdt_ha_masked": [], "lddt": [], "lddt_res": [],
           "lddt_masked": [], "lddt_res_masked": []}
    for s in range(pred.shape[0]):
        p, t, m = pred[s].astype(np.float64), true[s].astype(np.float64), mask[s].astype(bool)
        out["tm"].append(vm.compute_tm_score_python(p, t))
        a, b = vm.compute_gdt(p, t)
        out["gdt_ts"].append(a), out["gdt_ha"].append(b)
        a, b = vm.compute_gdt(p, t, m)
        out["gdt_ts_masked"].append(a), out["gdt_ha_masked"].append(b)
        g, r = vm.compute_lddt(p, t)
        out["lddt"].append(g), out["lddt_res"].append(r)
        g, r = vm.compute_lddt(p, t, m)
        out["lddt_masked"].append(g), out["lddt_res_masked"].append(r)
    out = {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}
    out["rmsf"] = vm.compute_rmsf(ens.astype(np.float64))
    out["aligned0"] = vm.kabsch_align(pred[0].astype(np.float64), true[0].astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)


# --------------------------------------------------------------------------- PDB writer
def gen_pdb():
    ns = {"np": np}
    src = os.path.join(REF, "generate_ensemble_pdbs.py")
    ns["compute_backbone_oxygen"] = lift_function(src, "compute_backbone_oxygen", ns)
    write_pdb = lift_function(src, "write_pdb", ns)
    n, ca, c, mask, seq = cases.pdb_inputs()
    path = os.path.join(HERE, "ensemble.pdb")
    for m in range(n.shape[0]):
        write_pdb(n[m], ca[m], c[m], mask, path, model_num=m + 1, sequence=seq, pdb_id="1abc", chain_id="B",
                  title="synthetic ensemble", num_models=n.shape[0])
    path2 = os.path.join(HERE, "ensemble_plain.pdb")      # no sequence / id / title, all residues valid
    for m in range(2):
        write_pdb(n[m], ca[m], c[m], np.ones_like(mask), path2, model_num=m + 1, num_models=2)


# --------------------------------------------------------------------------- encoder
def gen_encoders():
    sys.path.insert(0, os.path.join(REF, "models"))
    import encoder as ref_enc
    out = {}
    for tag, case in cases.ENCODER_CASES.items():
        sd, nl, B, L, mkind, pseed, dseed = case
        enc = ref_enc.ProteinEncoder(seqemb_dim=sd, nlayers=nl, dropout=0.0).double()
        shapes = {k: v.shape for k, v in enc.state_dict().items() if k != "enc.pe.pe"}
        enc.load_state_dict({k: T(v) for k, v in cases.encoder_params(shapes, pseed).items()}, strict=False)
        enc.train()                                   # dropout 0: train mode keeps nn.TransformerEncoderLayer off its fused path
        latent_dropout = enc.latent.global_attention.dropout
        enc.latent.global_attention.dropout = 0.0     # (hard-coded 0.1 in the reference, :159)
        xs, mask, coef = cases.encoder_inputs(case)
        H = enc.enc(*[T(a) for a in xs], T(mask))
        res = enc.latent(H, T(mask))
        enc.latent.global_attention.dropout = latent_dropout
        loss = sum((r * T(c)).sum() for r, c in zip(res, coef))
        loss.backward()
        for name, r in zip(("mu_g", "lv_g", "mu_l", "lv_l"), res):
            out[f"{tag}.{name}"] = r.detach().numpy()
        out[f"{tag}.H"] = (H.detach() * T(mask)[..., None]).numpy().astype(np.float32)
        pack_grads_big({k: p.grad for k, p in enc.named_parameters() if p.grad is not None}, out, tag)
        print("encoder case", tag, "done", flush=True)
    out["pe"] = enc.enc.pe.pe[:64].numpy().astype(np.float32)      # the reference builds this buffer in float32 (:17-23)
    np.savez_compressed(os.path.join(HERE, "encoders.npz"), **out)


# --------------------------------------------------------------------------- HierCVAE (models/model.py:15-116)
def gen_hiercvae():
    """The reference's top-level module end to end (encoder -> reparameterisation with INJECTED noise -> ResidueDecoder), in
    float64, dropout 0; plus its state_dict layout (names and shapes at the default constructor arguments)."""
    sys.path.insert(0, os.path.join(REF, "models"))
    import json
    import model as ref_model
    case = cases.HIERCVAE_CASE
    sd, nl, B, L, mkind, pseed, dseed = case
    vae = ref_model.HierCVAE(seqemb_dim=sd, nlayers=nl, dropout=0.0).double()
    shapes = {k: v.shape for k, v in vae.state_dict().items() if k != "encoder.enc.pe.pe"}
    vae.load_state_dict({k: T(v) for k, v in cases.hiercvae_params(shapes, pseed).items()}, strict=False)
    vae.train()                                       # dropout 0: keeps nn.TransformerEncoderLayer off its fused path
    vae.encoder.latent.global_attention.dropout = 0.0   # (hard-coded 0.1 in the reference, models/encoder.py:159)
    xs, mask, eps, coef = cases.hiercvae_inputs(case)
    noise = [T(eps[0]), T(eps[1])]
    vae.encoder.reparam = lambda mu, lv: mu + noise.pop(0) * torch.exp(0.5 * lv)      # models/encoder.py:231-236, eps fixed
    res = vae(*[T(a) for a in xs], T(mask))
    assert not noise
    loss = sum((r * T(c)).sum() for r, c in zip(res, coef))
    loss.backward()
    out = {}
    for name, r in zip(("N", "CA", "C", "logits", "mu_g", "lv_g", "mu_l", "lv_l"), res):
        out[f"vae.{name}"] = r.detach().numpy()
    pack_grads_big({k: p.grad for k, p in vae.named_parameters() if p.grad is not None}, out, "vae")
    np.savez_compressed(os.path.join(HERE, "hiercvae.npz"), **out)
    full = ref_model.HierCVAE(seqemb_dim=1280)        # default widths: the layout a reference checkpoint has
    with open(os.path.join(HERE, "hiercvae_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in full.state_dict().items()}, f, indent=0, sort_keys=True)
    print("hiercvae done", flush=True)


# --------------------------------------------------------------------------- data path (centring + padding collate)
def gen_data():
    import importlib.util, types
    for name in ("h5py",):                       # models/data.py imports h5py at the top; nothing used here touches it
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_data", os.path.join(REF, "models", "data.py"))
    rd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rd)
    confs = cases.data_conformers()
    items = [rd.EnsembleDataset._process_conformer(None, c) for c in confs]          # centred 7-tuples (:153-194)
    n, ca, c, mask, emb, dih, lbl = rd._collate_single_batch(items)                  # (:219-266)
    np.savez_compressed(os.path.join(HERE, "data.npz"), n=n.numpy(), ca=ca.numpy(), c=c.numpy(), mask=mask.numpy(),
                        emb=emb.numpy(), dih=dih.numpy(), labels=lbl.numpy(),
                        item_labels=np.concatenate([it[6].numpy() for it in items]))


if __name__ == "__main__":
    gen_edges()
    gen_layers()
    gen_decoders()
    gen_big_decoders()
    gen_losses()
    gen_kabsch()
    gen_metrics()
    gen_data()
    gen_pdb()
    gen_encoders()
    gen_hiercvae()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
