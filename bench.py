#!/usr/bin/env python3
"""Benchmark of the hot path: EGNN decoder fwd+bwd + compute_total_loss fwd+bwd (train) and decoder fwd + Kabsch
RMSD (decode), on synthetic backbone ensembles.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config train|decode|mixed|stress]

One JSON line on stdout (rank 0).  Default ``--config train`` = BASELINE.json configs[1] (L=256, 6 EGNN layers, 256
conformers per GPU, bf16 edge MLP): ``value`` = train conformers/s with the batch resident in HBM, ``e2e`` = the same
step fed from pinned host buffers through the public module / loss API with the H2D copies and the loss read-back inside
the timed region, ``roofline`` = the dominant tcgen05 kernel and the whole step against the measured bf16 tensor peak
(SURVEY.md 8d assigns K1 to the tensor roofline) plus one entry per kernel class K1..K4, ``decode`` = configs[3] in
full (100 000 latent samples through ``distributed.decode_ensemble``, final gather inside the timed region).
``--config decode | mixed | stress`` put configs[3] / configs[2] / configs[4] on the headline instead.
``--impl reference`` times the reference's own CPU implementation of the path on the host cores: the unmodified
reference sources from ``oracle/_ref`` (copied there by ``oracle/make_ref.py`` where /root/reference exists; kind
"reference"), else the CPU restatement in ``oracle/`` (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

LOSS_W = dict(klw_g=1.0, klw_l=0.5, w_pair=10.0, pair_stride=8, w_dihedral=20.0, w_rama=400.0, w_bond=500.0,
              w_angle=500.0, w_rec=10.0, w_seq=50.0, w_clash=300.0)             # models/vae.py:39-50
CFG = dict(L=256, layers=6, batch_per_gpu=256, z_g=512, z_l=256, hidden=256, max_neighbors=40, dropout=0.1)
DECODE = dict(L=100, layers=8, chunk=2048, z_g=512, z_l=256, samples=100000)
MIXED = dict(Lmin=64, Lmax=512, layers=6, batch_per_gpu=256, z_g=512, z_l=256, hidden=256, max_neighbors=40, dropout=0.1)
STRESS = dict(L=1024, layers=6, batch_per_gpu=8, z_g=512, z_l=256, hidden=256, dropout=0.1, pair_stride=1)


def synth_batch(B, L, z_g, z_l, seed, device="cpu", pin=False, lengths=None):
    """Synthetic inputs of SURVEY.md 8(d): latents, random-walk targets, labels, posterior stats."""
    g = torch.Generator().manual_seed(seed)
    ca = torch.cumsum(torch.randn(B, L, 3, generator=g) * 2.2, 1)
    ca = ca - ca.mean(1, keepdim=True)
    mask = torch.ones(B, L)
    if lengths is not None:
        for b, n in enumerate(lengths):
            mask[b, n:] = 0
    d = dict(
        z_g=torch.randn(B, z_g, generator=g), z_l=torch.randn(B, L, z_l, generator=g),
        target_CA=ca, target_N=ca + 0.8 * torch.randn(B, L, 3, generator=g),
        target_C=ca + 0.8 * torch.randn(B, L, 3, generator=g),
        labels=torch.randint(0, 20, (B, L), generator=g), mask=mask,
        mu_g=torch.randn(B, z_g, generator=g), lv_g=0.1 * torch.randn(B, z_g, generator=g),
        mu_l=torch.randn(B, L, z_l, generator=g), lv_l=0.1 * torch.randn(B, L, z_l, generator=g))
    if pin:
        d = {k: v.pin_memory() for k, v in d.items()}
    if device != "cpu":
        d = {k: v.to(device) for k, v in d.items()}
    return d


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): ONE `nvidia-smi -lms 200`
    process started before the warm-up (its start-up cost stays outside the timed region; re-spawning nvidia-smi for
    every sample stalls kernel launches for milliseconds), rows kept only if they arrive inside [mark_start, mark_stop]."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.all_rows, self.rows, self.proc = index, [], [], None
        self.t0 = self.t1 = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.all_rows.append((time.time(), [c.strip() for c in line.strip().split(",")]))
        except Exception:
            pass

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()
        time.sleep(0.25)                           # let the sample in flight arrive
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.all_rows if self.t0 is not None and self.t0 <= t <= self.t1 + 0.25]
        self.rows = rows or [r for _, r in self.all_rows[-2:]]

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), float(p["bf16_tflops_sustained"]), "measured"
    return 6650.0, 1400.0, "fallback"


def band_edges(L, W=40):
    if L < 2:
        return 0
    return L * (L - 1) if W >= L - 1 else 2 * W * L - W * (W + 1)


def flops_per_conformer(L, layers, W=40, H=256, train=True):
    """Algorithmic flops of SURVEY.md 8(d), factored form: F_fwd per conformer-layer = 4 E H^2 (W2, W5) + 2 E H (w6)
    + 2 N H 2H (Wa, Wb) + 2 N (2H H + H H) (phi_h); F_train = 3 F_fwd (dgrad + wgrad; recompute is not counted)."""
    E, N = band_edges(L, W), L
    f = 4.0 * E * H * H + 2.0 * E * H + 2.0 * N * H * 2 * H + 2.0 * N * (2 * H * H + H * H)
    return (3.0 if train else 1.0) * f * layers


def mixed_lengths(n, seed, lo, hi):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, hi + 1, (n,), generator=g).tolist()


# ------------------------------------------------------------------------------------------ CPU arm
def _ref_modules():
    """The unmodified reference modules from oracle/_ref (None when the copy was never made)."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not (os.path.exists(os.path.join(ref, "en_gnn_decoder.py")) and os.path.exists(os.path.join(ref, "losses.py"))):
        return None
    import importlib.util
    mods = []
    for name in ("en_gnn_decoder", "losses", "kabsch_ref"):
        spec = importlib.util.spec_from_file_location(f"pev_ref_{name}", os.path.join(ref, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


def cpu_step_fn(B, cfg, config="train"):
    """One step of the reference's CPU path on a bounded sample: returns (step, kind, description).
    train: decoder fwd + compute_total_loss + backward, fp32; decode: decoder fwd (no grad) + kabsch_rmsd per sample."""
    import synth
    torch.manual_seed(0)
    ref = _ref_modules()
    d = synth_batch(B, cfg["L"], cfg["z_g"], cfg["z_l"], seed=0)
    W = cfg.get("max_neighbors", 40)
    if ref is not None:
        rd, rl, rk = ref
        dec = rd.EGNNDecoder(cfg["z_g"], cfg["z_l"], hidden_dim=256, num_layers=cfg["layers"], max_neighbors=W,
                             dropout=cfg.get("dropout", 0.1))
        kind = "reference"
        what = "unmodified reference (oracle/_ref: models/en_gnn_decoder.py + models/losses.py), fp32, torch CPU"
        if config == "decode":
            dec.eval()
            ref_ca, ones = d["target_CA"][0], torch.ones(cfg["L"])

            def step():                  # generate_ensemble_pdbs.py:554 + :594: decode, then kabsch_rmsd per sample
                with torch.no_grad():
                    n, ca, c, lg = dec(d["z_g"], d["z_l"], mask=d["mask"])
                return sum(rk.kabsch_rmsd(ca[s], ref_ca, ones) for s in range(B))
            return step, kind, what + " + generate_ensemble_pdbs.py:kabsch_rmsd"
        dec.train()
        tdih = rl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])

        def step():
            dec.zero_grad(set_to_none=True)
            n, ca, c, lg = dec(d["z_g"], d["z_l"], mask=d["mask"])
            res = rl.compute_total_loss(n, ca, c, lg, d["target_N"], d["target_CA"], d["target_C"], d["labels"], d["mask"],
                                        d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih, **LOSS_W)
            res["total"].backward()
            return float(res["total"].detach())
        return step, kind, what
    from oracle import egnn_oracle, kabsch_oracle, losses_oracle
    params = synth.make_params(synth.decoder_param_shapes(cfg["z_g"], cfg["z_l"], 256, cfg["layers"]), 0)
    sd = {k: torch.tensor(v).requires_grad_() for k, v in params.items()}
    cache = {}
    kind, what = "port", "oracle/ restatement (edge lists cached between conformers: faster than the reference), fp32"
    if config == "decode":
        ref_ca = d["target_CA"][0].double().numpy()

        def step():
            with torch.no_grad():
                n, ca, c, lg = egnn_oracle.egnn_decoder(sd, d["z_g"], d["z_l"], d["mask"], max_neighbors=W, edge_cache=cache)
            return sum(kabsch_oracle.kabsch_rmsd(ca[s].double().numpy(), ref_ca, torch.ones(cfg["L"]).numpy())
                       for s in range(B))
        return step, kind, what
    tdih = losses_oracle.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])

    def step():
        for v in sd.values():
            v.grad = None
        n, ca, c, lg = egnn_oracle.egnn_decoder(sd, d["z_g"], d["z_l"], d["mask"], max_neighbors=W, edge_cache=cache)
        res = losses_oracle.compute_total_loss(n, ca, c, lg, d["target_N"], d["target_CA"], d["target_C"], d["labels"],
                                               d["mask"], d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih, **LOSS_W)
        res["total"].backward()
        return float(res["total"].detach())
    return step, kind, what


def cpu_config(config):
    if config == "decode":
        return dict(DECODE, max_neighbors=40), 4, "decoded_conformers_per_s"
    if config == "mixed":
        return dict(MIXED, L=288), 2, "train_conformers_per_s"          # mean of U{64..512}
    if config == "stress":
        return dict(STRESS, max_neighbors=40), 1, "train_conformers_per_s"
    return CFG, 2, "train_conformers_per_s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, B, metric = cpu_config(args.config)
    step, kind, what = cpu_step_fn(B, cfg, args.config)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    sample = f"{B} conformers per step, L={cfg['L']}, {cfg['layers']} layers; {what}"
    out = {"impl": "reference", "metric": metric, "value": val, "unit": "conformers/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "strong" if args.config == "decode" else "weak", "vs_baseline": None,
           "dtype": "fp32", "data": "synthetic", "config": workload_config(args.gpus, args.config),
           "cpu_baseline": {"value": val, "unit": "conformers/s", "cores": torch.get_num_threads(), "kind": kind,
                            "sample": sample},
           "e2e": {"value": val, "unit": "conformers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(n_gpus, config="train"):
    if config == "decode":
        return {"workload": f"configs[3]: ensemble generation, {DECODE['samples']} latent samples -> ResidueDecoder-shape "
                            f"decoder (8 EGNN layers, hidden 256, W=40) at L={DECODE['L']} + Kabsch RMSD against one reference "
                            "structure, samples sharded across GPUs, one final gather of the RMSDs",
                "L": DECODE["L"], "layers": DECODE["layers"], "samples": DECODE["samples"], "chunk": DECODE["chunk"],
                "parallelism": f"dp{n_gpus}", "cache": "inputs_larger_than_l2 (10.2 GB of latents / n_gpus per rank)"}
    if config == "mixed":
        return {"workload": f"configs[2]: multi-protein training, lengths U{{{MIXED['Lmin']}..{MIXED['Lmax']}}} padded to "
                            f"{MIXED['Lmax']}, 6 EGNN layers, {MIXED['batch_per_gpu']} conformers per GPU assigned to ranks "
                            "by edge count, decoder fwd+bwd + compute_total_loss(dp_normalize) fwd+bwd + Adam step, "
                            "bucketed gradient all-reduce overlapped with backward",
                "Lmax": MIXED["Lmax"], "layers": MIXED["layers"], "batch_per_gpu": MIXED["batch_per_gpu"],
                "global_batch": MIXED["batch_per_gpu"] * n_gpus, "parallelism": f"dp{n_gpus}",
                "cache": "inputs_larger_than_l2"}
    if config == "stress":
        return {"workload": f"configs[4]: L={STRESS['L']}, dense residue graph (W=1023, E=1 047 552 per conformer) and the W=40 "
                            f"band, {STRESS['batch_per_gpu']} conformers per GPU, 6 EGNN layers with per-edge activations "
                            "recomputed in backward, pair_stride=1 (1024^2 CA pairs) and the 4.7 M-pair clash tile",
                "L": STRESS["L"], "layers": STRESS["layers"], "batch_per_gpu": STRESS["batch_per_gpu"],
                "global_batch": STRESS["batch_per_gpu"] * n_gpus, "parallelism": f"dp{n_gpus}",
                "cache": "inputs_larger_than_l2"}
    return {"workload": "configs[1]: single protein L=256, 6 EGNN layers (hidden 256, W=40), 256 conformers per GPU, "
                        "decoder fwd+bwd + compute_total_loss fwd+bwd + Adam step, bf16 edge MLP",
            "L": CFG["L"], "layers": CFG["layers"], "batch_per_gpu": CFG["batch_per_gpu"],
            "global_batch": CFG["batch_per_gpu"] * n_gpus, "parallelism": f"dp{n_gpus}",
            "cache": "inputs_larger_than_l2 (603 MB of latents/posteriors + 7.4 GB of per-edge bf16 streams per layer)"}


# ------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    """Process-group / device context and the timing helper shared by the four configurations."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def timed(self, fn, steps):
        """max over ranks of the CUDA-event time of `steps` calls, barrier + synchronize on both sides."""
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_train_step(ctx, dec, loss_w, dp_normalize=False):
    """(step(batch, tdih) -> loss tensor).  Data parallel: gradients live in flat buckets that are all-reduced while
    backward is still running (distributed.GradBuckets); single GPU: plain optimizer.zero_grad."""
    from protein_ensemble_vae_b200 import compute_total_loss
    from protein_ensemble_vae_b200 import distributed as pdist
    params = list(dec.parameters())
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    buckets = pdist.GradBuckets(pdist.decoder_buckets(dec)) if ctx.world > 1 else None

    def step(d, tdih):
        outs = dec(d["z_g"], d["z_l"], d["mask"])
        res = compute_total_loss(outs[0], outs[1], outs[2], outs[3], d["target_N"], d["target_CA"], d["target_C"],
                                 d["labels"], d["mask"], d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih,
                                 dp_normalize=dp_normalize, **loss_w)
        res["total"].backward()
        if buckets is not None:
            buckets.finish()
            opt.step()
            buckets.zero()
        else:
            opt.step()
            opt.zero_grad(set_to_none=True)
        return res["total"].detach()
    return step


def decode_full(ctx, dd, e2e=False):
    """configs[3] in full: S latent samples (this rank's contiguous share resident in HBM, or -- e2e -- streamed chunk by
    chunk from pinned host memory) -> decoder forward -> Kabsch RMSD against one reference, final gather of the [S]
    RMSDs inside the timed region.  Returns (total ms of ONE pass, setup dict)."""
    from protein_ensemble_vae_b200 import EGNNDecoder
    from protein_ensemble_vae_b200 import distributed as pdist
    S, L = dd["samples"], dd["L"]
    torch.manual_seed(1)
    dec8 = EGNNDecoder(dd["z_g"], dd["z_l"], hidden_dim=256, num_layers=dd["layers"], max_neighbors=40, dropout=0.1,
                       precision="bf16").to(ctx.dev).eval()
    lo, hi = pdist.shard_range(S, ctx.rank, ctx.world)
    gen = torch.Generator(device=ctx.dev).manual_seed(100 + ctx.rank)
    mask1 = torch.ones(L, device=ctx.dev)
    ref_ca = torch.cumsum(torch.randn(L, 3, device=ctx.dev, generator=gen) * 2.2, 0)
    n_loc = hi - lo
    if not e2e:
        # full-size tensors would be 10 GB per rank even for the ranks' foreign shares: allocate the local share only
        zg = torch.randn(n_loc, dd["z_g"], device=ctx.dev, generator=gen)
        zl = torch.randn(n_loc, L, dd["z_l"], device=ctx.dev, generator=gen)

        class Shard:        # decode_ensemble indexes [lo:hi) of "global" tensors: present the local share at that offset
            def __init__(self, t):
                self.t, self.shape = t, (S,) + tuple(t.shape[1:])

            def __getitem__(self, sl):
                return self.t[sl.start - lo:sl.stop - lo]

            @property
            def device(self):
                return self.t.device

        def one_pass():
            return pdist.decode_ensemble(dec8, Shard(zg), Shard(zl), mask1, ref_ca, chunk=dd["chunk"])
        return one_pass, {"decoder": dec8}
    # end to end: pinned host latents of the local share, chunk by chunk through DevicePrefetcher, RMSDs back to the host
    from protein_ensemble_vae_b200 import DevicePrefetcher, kabsch_rmsd_batch
    chunk = dd["chunk"]
    hostg = torch.randn(min(n_loc, 4 * chunk), dd["z_g"]).pin_memory()          # a ring of 4 distinct host chunks
    hostl = torch.randn(min(n_loc, 4 * chunk), L, dd["z_l"]).pin_memory()
    nchunks = (n_loc + chunk - 1) // chunk

    def batches():
        for c in range(nchunks):
            n = min(chunk, n_loc - c * chunk)
            o = (c % 4) * chunk
            o = 0 if o + n > hostg.shape[0] else o
            yield {"z_g": hostg[o:o + n], "z_l": hostl[o:o + n]}

    def one_pass():
        rm = []
        with torch.no_grad():
            for d in DevicePrefetcher(batches(), ctx.dev):
                n, ca, c, lg = dec8(d["z_g"], d["z_l"], mask1.unsqueeze(0).expand(d["z_l"].shape[0], -1))
                rm.append(kabsch_rmsd_batch(ca, ref_ca, mask1))
        local = torch.cat(rm)
        if ctx.world > 1:
            parts = [torch.empty(pdist.shard_range(S, r, ctx.world)[1] - pdist.shard_range(S, r, ctx.world)[0],
                                 device=ctx.dev) for r in range(ctx.world)]
            width = max(p.numel() for p in parts)
            pad = torch.zeros(width, device=ctx.dev)
            pad[:local.numel()] = local
            outs = [torch.empty_like(pad) for _ in range(ctx.world)]
            ctx.dist.all_gather(outs, pad)
            local = torch.cat([o[:p.numel()] for o, p in zip(outs, parts)])
        return local.cpu()
    bytes_h2d = n_loc * (dd["z_g"] + L * dd["z_l"]) * 4
    return one_pass, {"decoder": dec8, "h2d_bytes": bytes_h2d, "d2h_bytes": S * 4}


def micro_rooflines(dev, hbm, sm_mhz):
    """K2, K3(a), K3(b), K4 timed ALONE with CUDA events at config-2 sizes (inputs larger than L2 where the kernel is
    HBM-bound), against the roofline SURVEY.md 8(d) assigns to each."""
    from protein_ensemble_vae_b200 import _lib, kabsch_rmsd_batch
    from protein_ensemble_vae_b200 import losses as pl
    from protein_ensemble_vae_b200._lib import ptr, stream
    from protein_ensemble_vae_b200.graph import band_graph
    out = {}
    f_sm = (sm_mhz or 1700.0) * 1e6

    def t_ms(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    H, L, B = 256, CFG["L"], CFG["batch_per_gpu"]
    # ---- K2: exact-order segmented scatter-sum + coordinate update, stand-alone fp32 form (64 conformers: 1.2 GB of m)
    Bk = 64
    g = band_graph((L,) * Bk, 40, dev)
    N, E = g.num_nodes, g.num_edges
    m = torch.randn(E, H, device=dev)
    w = torch.randn(E, device=dev)
    x = torch.randn(N, 3, device=dev)
    agg, xo = torch.empty(N, H, device=dev), torch.empty(N, 3, device=dev)
    lib = _lib.lib()
    ms = t_ms(lambda: lib.call("pev_scatter_coord_fwd", ptr(m), ptr(w), ptr(x), ptr(g.dinv), ptr(g.row_ptr), ptr(g.col), N, H,
                               ptr(agg), ptr(xo), stream(m)))
    nbytes = E * H * 4 + E * 12 + N * H * 4 + N * 12 * 2 + N * 4
    out["K2_scatter_coord_fwd"] = {"kernel": "scatter_agg_kernel + coord_update_kernel (pev_scatter_coord_fwd)", "bound": "hbm",
                                   "ms_per_launch": ms, "algorithmic_gb": nbytes / 1e9, "achieved_gbs": nbytes / ms / 1e6,
                                   "frac": nbytes / ms / 1e6 / hbm, "sample": f"{Bk} conformers x L={L}, fp32 m [E,256]"}
    gagg, gxo = torch.randn(N, H, device=dev), torch.randn(N, 3, device=dev)
    gm, gw, gx = torch.empty(E, H, device=dev), torch.empty(E, device=dev), torch.empty(N, 3, device=dev)
    ms = t_ms(lambda: lib.call("pev_scatter_coord_bwd", ptr(gagg), ptr(gxo), ptr(w), ptr(x), ptr(g.dinv), ptr(g.row_ptr),
                               ptr(g.row), ptr(g.col), ptr(g.col_ptr), ptr(g.csc_perm), N, E, H, ptr(gm), ptr(gw), ptr(gx),
                               stream(m)))
    nbytes = E * H * 4 + E * 16 + N * H * 4 + N * 12 * 3
    out["K2_scatter_coord_bwd"] = {"kernel": "scatter_bwd kernels (pev_scatter_coord_bwd)", "bound": "hbm", "ms_per_launch": ms,
                                   "algorithmic_gb": nbytes / 1e9, "achieved_gbs": nbytes / ms / 1e6,
                                   "frac": nbytes / ms / 1e6 / hbm, "sample": f"{Bk} conformers x L={L}"}
    del m, gm, agg, gagg
    # ---- K3: loss kernels at config 2 (256 x 256 residues), by differences of pev_loss_fwd / bwd configurations
    d = synth_batch(B, L, CFG["z_g"], CFG["z_l"], seed=7, device=dev)
    pred = {k: (d["target_" + k] + 0.5 * torch.randn_like(d["target_" + k])).requires_grad_() for k in ("N", "CA", "C")}
    logits = torch.randn(B, L, 20, device=dev, requires_grad=True)
    mu_l, lv_l = d["mu_l"].clone().requires_grad_(), d["lv_l"].clone().requires_grad_()
    mu_g, lv_g = d["mu_g"].clone().requires_grad_(), d["lv_g"].clone().requires_grad_()
    tdih = pl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])

    def loss_fb(cfg, kl=True, reps=20):
        kw = dict(pred_N=pred["N"], pred_CA=pred["CA"], pred_C=pred["C"], logits=logits, target_N=d["target_N"],
                  target_CA=d["target_CA"], target_C=d["target_C"], target_dih=tdih, labels=d["labels"])
        if kl:
            kw.update(mu_l=mu_l, lv_l=lv_l, mu_g=mu_g, lv_g=lv_g)
        coef = torch.ones(17, device=dev)
        box = {}

        def fwd():
            box["t"] = pl._terms(cfg, d["mask"], **kw)

        def bwd():
            torch.autograd.grad(box["t"], [v for v in kw.values() if v.requires_grad], coef, retain_graph=True,
                                allow_unused=True)
        f = t_ms(fwd, reps)
        fwd()
        return f, t_ms(bwd, reps)
    f_all, b_all = loss_fb({"pair_stride": 8, "clash": True, "geometry": True})
    f_noclash, b_noclash = loss_fb({"pair_stride": 8, "clash": False, "geometry": True})
    f_nopair, b_nopair = loss_fb({"pair_stride": 0, "clash": False, "geometry": True})
    res = B * L
    fb = res * 2236 + B * 2 * CFG["z_g"] * 4
    bb = res * (2236 + 36 + 80 + 2048) + B * 4 * CFG["z_g"] * 4
    out["K3a_residue_stream_fwd"] = {"kernel": "loss_residue_fwd_kernel + loss_kl_fwd_kernel x2 + finalize", "bound": "hbm",
                                     "ms_per_launch": f_nopair, "algorithmic_gb": fb / 1e9, "achieved_gbs": fb / f_nopair / 1e6,
                                     "frac": fb / f_nopair / 1e6 / hbm,
                                     "note": "four launch-latency-bound kernels over 146 MB; includes the autograd wrapper"}
    out["K3a_residue_stream_bwd"] = {"kernel": "loss_residue_bwd_kernel + loss_kl_bwd_kernel x2", "bound": "hbm",
                                     "ms_per_launch": b_nopair, "algorithmic_gb": bb / 1e9, "achieved_gbs": bb / b_nopair / 1e6,
                                     "frac": bb / b_nopair / 1e6 / hbm}
    # K3(b) clash: the C entry points called directly with preallocated buffers, the clash term switched on and off with
    # everything else but the (then nearly empty) residue kernel disabled -- GPU-bound on both sides of the difference.
    # (Differences of the autograd-level passes above are launch-latency-bound on the host and hid most of the kernel.)
    import ctypes
    tens = {k: None for k in pl._DIFF + pl._CONST}
    tens.update(pred_N=pred["N"].detach().contiguous(), pred_CA=pred["CA"].detach().contiguous(),
                pred_C=pred["C"].detach().contiguous(), mask=d["mask"].float().contiguous())
    acc_g = torch.zeros(2 * pl.NUM_TERMS, dtype=torch.float64, device=dev)
    acc_s = torch.zeros(B * 8, dtype=torch.float64, device=dev)
    terms = torch.empty(pl.NUM_TERMS, dtype=torch.float32, device=dev)
    inv_den = torch.empty(pl.NUM_TERMS + 2 * B, dtype=torch.float32, device=dev)
    coef17 = torch.ones(pl.NUM_TERMS, device=dev)
    gbuf = [torch.zeros_like(tens[k]) for k in ("pred_N", "pred_CA", "pred_C")]
    st = stream(tens["mask"])
    clash_ms = {}
    for on in (True, False):
        a = pl._make_args(tens, {"pair_stride": 0, "clash": on, "geometry": False})
        lib.call("pev_loss_fwd", ctypes.byref(a), ptr(acc_g), ptr(acc_s), st)
        lib.call("pev_loss_finalize", ptr(acc_g), ptr(acc_s), B, ptr(terms), ptr(inv_den), st)
        clash_ms["fwd", on] = t_ms(lambda: lib.call("pev_loss_fwd", ctypes.byref(a), ptr(acc_g), ptr(acc_s), st), reps=50)
        clash_ms["bwd", on] = t_ms(lambda: lib.call("pev_loss_bwd", ctypes.byref(a), ptr(coef17), ptr(inv_den), ptr(gbuf[0]),
                                                    ptr(gbuf[1]), ptr(gbuf[2]), None, None, None, None, None, st), reps=50)
    # K3(a) the same way: the residue stream + both KL kernels (+ finalize) through the C ABI, no pair / clash term
    full = {k: None for k in pl._DIFF + pl._CONST}
    full.update(tens)
    full.update(logits=logits.detach().contiguous(), mu_l=mu_l.detach().contiguous(), lv_l=lv_l.detach().contiguous(),
                mu_g=mu_g.detach().contiguous(), lv_g=lv_g.detach().contiguous(), target_N=d["target_N"].contiguous(),
                target_CA=d["target_CA"].contiguous(), target_C=d["target_C"].contiguous(), target_dih=tdih.contiguous(),
                labels=d["labels"].long().contiguous())
    a_full = pl._make_args(full, {"pair_stride": 0, "clash": False, "geometry": True})
    gfull = {k: torch.zeros_like(full[k]) for k in pl._DIFF}

    def k3a_fwd():
        lib.call("pev_loss_fwd", ctypes.byref(a_full), ptr(acc_g), ptr(acc_s), st)
        lib.call("pev_loss_finalize", ptr(acc_g), ptr(acc_s), B, ptr(terms), ptr(inv_den), st)
    k3a_fwd()
    f_abi = t_ms(k3a_fwd, reps=50)
    b_abi = t_ms(lambda: lib.call("pev_loss_bwd", ctypes.byref(a_full), ptr(coef17), ptr(inv_den),
                                  *[ptr(gfull[k]) for k in pl._DIFF], st), reps=50)
    for tag, t_abi, t_auto, nb in (("fwd", f_abi, f_nopair, fb), ("bwd", b_abi, b_nopair, bb)):
        e = out["K3a_residue_stream_" + tag]
        e.update(ms_per_launch=t_abi, achieved_gbs=nb / t_abi / 1e6, frac=nb / t_abi / 1e6 / hbm, ms_with_autograd_wrapper=t_auto)
        e["note"] = "C ABI calls back to back (GPU-bound); ms_with_autograd_wrapper: the same through losses._terms / autograd.grad"
    pairs = B * (3 * L) * (3 * L - 1) / 2.0
    for tag, evals in (("fwd", 2.0), ("bwd", 2.0)):
        tms = max(clash_ms[tag, True] - clash_ms[tag, False], 1e-4)
        # the kernel walks the full row (both triangles: atomic-free gradients), i.e. 2 evaluations per unordered pair
        t_alu = pairs * evals * 7 / (148 * 128 * f_sm) * 1e3          # 3 sub + 3 fma + 1 compare per evaluation
        t_mufu = pairs * 0.02 / (148 * 16 * f_sm) * 1e3               # sqrt only for the ~1 % close pairs
        t_hbm = B * 3 * L * 16 / (hbm * 1e9) * 1e3
        roof = max(t_alu, t_mufu, t_hbm)
        out["K3b_clash_" + tag] = {"kernel": f"loss_clash_kernel<{'true' if tag == 'bwd' else 'false'}>", "bound": "fp32-alu",
                                   "ms_per_launch": tms, "pairs": pairs, "pair_evals_per_s": pairs * evals / tms * 1e3,
                                   "roof_ms": roof, "frac": roof / tms,
                                   "ms_with_and_without": [clash_ms[tag, True], clash_ms[tag, False]],
                                   "note": "pev_loss_fwd / pev_loss_bwd through the C ABI with only the clash term on, minus the "
                                           "same call with it off; target-like (spread-out) coordinates: ~1 % close pairs"}
    M = (L + 7) // 8
    out["K3b_pair_distance_fwd"] = {"kernel": "loss_pair_kernel", "bound": "launch-latency", "ms_per_launch": max(f_noclash - f_nopair, 1e-4),
                                    "pairs": float(B * M * M), "note": "32 x 32 strided CA pairs per conformer at pair_stride=8"}
    # ---- K4: batched Kabsch RMSD, 100 000 x L=100 against one reference
    S, Lk = DECODE["samples"], DECODE["L"]
    ca = torch.randn(S, Lk, 3, device=dev)
    refc = torch.randn(Lk, 3, device=dev)
    ms = t_ms(lambda: kabsch_rmsd_batch(ca, refc))
    nbytes = S * (12 * Lk + 4)
    out["K4_kabsch_rmsd"] = {"kernel": "kabsch_rmsd_kernel", "bound": "hbm", "ms_per_launch": ms, "algorithmic_gb": nbytes / 1e9,
                             "achieved_gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / hbm,
                             "sample": f"{S} conformers x L={Lk}, shared reference, no mask"}
    return out


# dram__bytes_read.sum + dram__bytes_write.sum per edge of each v2 edge kernel (ncu --set full, profiles/)
NCU_TRAFFIC_PER_EDGE = {"edge2_fwd1": 980.4, "edge2_fwd2": 956.6, "edge2_bwd2": 1470.9, "edge2_wgrad5": 1035.8,
                        "edge2_bwd1": 974.9, "edge2_wgrad2": 544.8, "edge2_sums": 774.0}
KSPEC = {  # tag: (kernel, bf16 [E,256] streams read+written, extra bytes per edge, tcgen05 GEMMs, tanh per edge-feature)
    "edge2_fwd1": ("fwd1_kernel", 2, 16, 1, 2),
    "edge2_fwd2": ("fwd2_kernel", 2, 4, 1, 1),
    "edge2_bwd2": ("bwd2_kernel", 3, 8, 1, 2),
    "edge2_bwd1": ("bwd1_kernel", 2, 20, 1, 1),
    "edge2_wgrad5": ("wgrad_kernel<5>", 2, 4, 1, 1),
    "edge2_wgrad2": ("wgrad_kernel<2>", 1, 12, 1, 1),
    "edge2_sums": ("edge_sums_kernel", 2, 8, 0, 0),
}


def k1_rooflines(prof, ms_total, E, N, hbm, tf, layers):
    """Per-kernel rooflines of the tcgen05 edge kernels from the CUDA-event pairs recorded inside the timed region.
    K1 is assigned to the TENSOR roofline: frac = algorithmic flops (2 E 256^2 per GEMM) / time / measured
    sustained bf16 peak.  The HBM view (bytes of the bf16 [E,256] streams the six-kernel design moves) and
    traffic / compulsory bytes (3 N H 4 per layer) say why it is where it is."""
    compulsory = 3.0 * N * 256 * 4
    kernels = {}
    for tag, (kname, streams, extra, gemms, tanhs) in KSPEC.items():
        ev = prof.get(tag, []) if prof else []
        if not ev:
            continue
        tot = sum(s.elapsed_time(e) for s, e in ev)
        per = tot / len(ev)
        nbytes = (512.0 * streams + extra) * E
        flops = gemms * 2.0 * E * 256 * 256
        traffic = NCU_TRAFFIC_PER_EDGE.get(tag, 0.0) * E
        kernels[tag] = {"kernel": kname, "bound": "tensor" if gemms else "hbm", "ms_per_launch": per,
                        "launches_timed": len(ev), "share_of_step": tot / ms_total,
                        "tensor_tflops": flops / (per * 1e-3) / 1e12, "tensor_frac": flops / (per * 1e-3) / 1e12 / tf,
                        "algorithmic_gb": nbytes / 1e9, "hbm_gbs": nbytes / (per * 1e-3) / 1e9,
                        "hbm_frac": nbytes / (per * 1e-3) / 1e9 / hbm,
                        "mufu_frac": tanhs * 256.0 * E / (per * 1e-3) / (148 * 16 * 1.965e9),
                        "traffic_gb": traffic / 1e9, "wasted_traffic_ratio": traffic / compulsory,
                        "frac": (flops / (per * 1e-3) / 1e12 / tf) if gemms else nbytes / (per * 1e-3) / 1e9 / hbm}
    return kernels


def run_train(ctx, args):
    from protein_ensemble_vae_b200 import DevicePrefetcher, EGNNDecoder, _lib
    from protein_ensemble_vae_b200 import losses as pl
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    lib = _lib.lib()
    B, L = CFG["batch_per_gpu"], CFG["L"]
    torch.manual_seed(0)
    dec = EGNNDecoder(CFG["z_g"], CFG["z_l"], hidden_dim=CFG["hidden"], num_layers=CFG["layers"],
                      max_neighbors=CFG["max_neighbors"], dropout=CFG["dropout"], precision="bf16").to(dev).train()
    host = synth_batch(B, L, CFG["z_g"], CFG["z_l"], seed=rank, pin=True)
    resident = {k: v.to(dev) for k, v in host.items()}
    tdih = pl.compute_dihedrals_from_coords(resident["target_N"], resident["target_CA"], resident["target_C"],
                                            resident["mask"])
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    step = make_train_step(ctx, dec, LOSS_W)

    sampler = ClockSampler(ctx.local)
    if rank == 0 and not os.environ.get("PEV_BENCH_NO_CLOCKS"):   # set under ncu: it would follow the child process
        sampler.start()
    for _ in range(args.warmup):
        step(resident, tdih)
    torch.cuda.synchronize()
    sampler.mark_start()
    _lib.PROFILE = {}
    n0 = lib.launch_count()
    ms = ctx.timed(lambda: step(resident, tdih), args.steps)
    launches = lib.launch_count() - n0
    prof, _lib.PROFILE = _lib.PROFILE, None
    sampler.mark_stop()
    value = B * world * args.steps / (ms / 1e3)

    if args.train_only:
        if rank == 0:
            print(json.dumps({"metric": "train_conformers_per_s", "value": value, "ms_per_step": ms / args.steps,
                              "gpu_launches": int(launches), "note": "--train-only profiling run"}), flush=True)
        return

    # end to end: pinned host buffers -> device every step (DevicePrefetcher: the copy of step i+1 runs on a side
    # stream under step i's kernels), loss read back to the host every step; K copies and K read-backs per K steps
    def e2e_steps(k):
        for d in DevicePrefetcher((host for _ in range(k)), dev):
            float(step(d, tdih))
    e2e_steps(2)
    ms_e2e = ctx.timed(lambda: e2e_steps(args.steps), 1)
    e2e = B * world * args.steps / (ms_e2e / 1e3)

    # decode: configs[3] in full, final gather inside the timed region; and its end-to-end form
    one_pass, _ = decode_full(ctx, DECODE)
    one_pass()
    ms_dec = ctx.timed(one_pass, 1)
    one_e2e, info = decode_full(ctx, DECODE, e2e=True)
    one_e2e()
    ms_dec_e2e = ctx.timed(one_e2e, 1)
    del one_pass, one_e2e
    torch.cuda.empty_cache()

    if rank == 0:
        hbm, tf, which = peaks()
        clocks = sampler.summary()
        E, N = band_edges(L) * B, L * B
        kernels = k1_rooflines(prof, ms, E, N, hbm, tf, CFG["layers"])
        dom = max(kernels, key=lambda t: kernels[t]["share_of_step"]) if kernels else None
        step_flops = flops_per_conformer(L, CFG["layers"]) * B
        step_tf = step_flops / (ms / args.steps * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": None, "achieved": None, "peak": tf, "unit": "TFLOP/s", "frac": None,
                "traffic": None, "peak_source": which + " (bf16_tflops_sustained: kernel timed inside a long step)",
                "step": {"algorithmic_tflop": step_flops / 1e12, "achieved": step_tf, "frac": step_tf / tf,
                         "formula": "SURVEY.md 8(d): 3 x layers x (4 E H^2 + 2 E H + 4 N H^2 + 6 N H^2) per conformer"}}
        if dom:
            k = kernels[dom]
            roof.update({"kernel": f"{k['kernel']} ({dom})", "achieved": k["tensor_tflops"], "frac": k["tensor_frac"],
                         "ms_per_launch": k["ms_per_launch"], "launches_timed": k["launches_timed"],
                         "share_of_step": k["share_of_step"], "hbm_frac_of_design_bytes": k["hbm_frac"],
                         "traffic": k["traffic_gb"] * 1e9, "wasted_traffic_ratio": k["wasted_traffic_ratio"],
                         "compulsory_bytes": 3.0 * N * 256 * 4,
                         "traffic_source": "profiles/r01_edge2_kernels_ncu.md (ncu --set full, per launch, scaled by E)"})
        try:
            kernels.update(micro_rooflines(dev, hbm, clocks.get("sm_mhz")))
        except Exception as e:      # the headline must not depend on the side measurements
            kernels["micro_error"] = repr(e)
        roof["kernels"] = kernels
        cpu = None
        if not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            Bc = 2
            cstep, kind, what = cpu_step_fn(Bc, CFG)
            cstep()
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 12):
                cstep()
                reps += 1
            dtc = time.perf_counter() - t0
            cpu = {"value": Bc * reps / dtc, "unit": "conformers/s", "cores": torch.get_num_threads(), "kind": kind,
                   "sample": f"{reps} steps of {Bc} conformers (of {B}), L={L}, {CFG['layers']} layers; {what}"}
        S = DECODE["samples"]
        out = {"metric": "train_conformers_per_s", "value": value, "unit": "conformers/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": workload_config(world), "clocks": clocks,
               "e2e": {"value": e2e, "unit": "conformers/s", "h2d_bytes_per_step": h2d_bytes * world,
                       "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
               "decode": {"metric": "decoded_conformers_per_s", "value": S / (ms_dec / 1e3), "unit": "conformers/s",
                          "scaling": "strong", "ms_total": ms_dec, "config": workload_config(world, "decode"),
                          "tensor_frac": flops_per_conformer(DECODE["L"], DECODE["layers"], train=False) * S
                          / (ms_dec * 1e-3) / 1e12 / tf / world,
                          "e2e": {"value": S / (ms_dec_e2e / 1e3), "unit": "conformers/s", "ms_total": ms_dec_e2e,
                                  "h2d_bytes": info["h2d_bytes"] * world, "d2h_bytes": info["d2h_bytes"]}}}
        print(json.dumps(out), flush=True)


def run_decode(ctx, args):
    from protein_ensemble_vae_b200 import _lib
    lib = _lib.lib()
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0 and not os.environ.get("PEV_BENCH_NO_CLOCKS"):
        sampler.start()
    one_pass, _ = decode_full(ctx, DECODE)
    for _ in range(max(1, min(args.warmup, 2))):
        one_pass()
    torch.cuda.synchronize()
    sampler.mark_start()
    n0 = lib.launch_count()
    steps = max(1, min(args.steps, 5))
    ms = ctx.timed(one_pass, steps)
    launches = lib.launch_count() - n0
    sampler.mark_stop()
    one_e2e, info = decode_full(ctx, DECODE, e2e=True)
    one_e2e()
    ms_e2e = ctx.timed(one_e2e, 1)
    if ctx.rank == 0:
        hbm, tf, which = peaks()
        S = DECODE["samples"]
        flops = flops_per_conformer(DECODE["L"], DECODE["layers"], train=False) * S
        tfs = flops / (ms / steps * 1e-3) / 1e12 / ctx.world
        cpu = None
        if not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            cfg, Bc, _ = cpu_config("decode")
            cstep, kind, what = cpu_step_fn(Bc, cfg, "decode")
            cstep()
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 20):
                cstep()
                reps += 1
            dtc = time.perf_counter() - t0
            cpu = {"value": Bc * reps / dtc, "unit": "conformers/s", "cores": torch.get_num_threads(), "kind": kind,
                   "sample": f"{reps} x {Bc} conformers, L={cfg['L']}, 8 layers, decoder fwd + kabsch_rmsd; {what}"}
        out = {"metric": "decoded_conformers_per_s", "value": S * steps / (ms / 1e3), "unit": "conformers/s",
               "n_gpus": ctx.world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": workload_config(ctx.world, "decode"), "clocks": sampler.summary(),
               "e2e": {"value": S / (ms_e2e / 1e3), "unit": "conformers/s", "h2d_bytes_per_step": info["h2d_bytes"] * ctx.world,
                       "d2h_bytes_per_step": info["d2h_bytes"], "ms_per_step": ms_e2e},
               "gpu_launches": int(launches),
               "roofline": {"bound": "tensor", "kernel": "decode pass (fwd1 + fwd2 per layer)", "achieved": tfs, "peak": tf,
                            "unit": "TFLOP/s", "frac": tfs / tf, "traffic": None, "peak_source": which},
               "cpu_baseline": cpu}
        print(json.dumps(out), flush=True)


def run_mixed(ctx, args):
    """configs[2]: ragged lengths, conformers assigned to ranks by edge count, dp_normalize, overlapped bucketed all-reduce."""
    from protein_ensemble_vae_b200 import EGNNDecoder, _lib
    from protein_ensemble_vae_b200 import distributed as pdist
    from protein_ensemble_vae_b200 import losses as pl
    lib = _lib.lib()
    c = MIXED
    Bg = c["batch_per_gpu"] * ctx.world
    lengths = mixed_lengths(Bg, 11, c["Lmin"], c["Lmax"])
    mine = pdist.balanced_shards(lengths, ctx.world, c["max_neighbors"])[ctx.rank]
    my_len = [lengths[i] for i in mine]
    torch.manual_seed(0)
    dec = EGNNDecoder(c["z_g"], c["z_l"], hidden_dim=c["hidden"], num_layers=c["layers"], max_neighbors=c["max_neighbors"],
                      dropout=c["dropout"], precision="bf16").to(ctx.dev).train()
    host = synth_batch(len(mine), c["Lmax"], c["z_g"], c["z_l"], seed=ctx.rank, pin=True, lengths=my_len)
    resident = {k: v.to(ctx.dev) for k, v in host.items()}
    tdih = pl.compute_dihedrals_from_coords(resident["target_N"], resident["target_CA"], resident["target_C"], resident["mask"])
    step = make_train_step(ctx, dec, LOSS_W, dp_normalize=True)
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0 and not os.environ.get("PEV_BENCH_NO_CLOCKS"):
        sampler.start()
    for _ in range(args.warmup):
        step(resident, tdih)
    torch.cuda.synchronize()
    sampler.mark_start()
    n0 = lib.launch_count()
    ms = ctx.timed(lambda: step(resident, tdih), args.steps)
    launches = lib.launch_count() - n0
    sampler.mark_stop()
    edges = torch.tensor([float(sum(band_edges(n, c["max_neighbors"]) for n in my_len))], device=ctx.dev)
    emax, emin = edges.clone(), edges.clone()
    if ctx.world > 1:
        ctx.dist.all_reduce(emax, op=ctx.dist.ReduceOp.MAX)
        ctx.dist.all_reduce(emin, op=ctx.dist.ReduceOp.MIN)
    if ctx.rank == 0:
        hbm, tf, which = peaks()
        flops = sum(flops_per_conformer(n, c["layers"]) for n in lengths)
        tfs = flops / (ms / args.steps * 1e-3) / 1e12 / ctx.world
        out = {"metric": "train_conformers_per_s", "value": Bg * args.steps / (ms / 1e3), "unit": "conformers/s",
               "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": workload_config(ctx.world, "mixed"), "clocks": sampler.summary(), "gpu_launches": int(launches),
               "valid_residues_global": int(sum(lengths)), "edge_balance_max_over_min": float(emax / emin),
               "roofline": {"bound": "tensor", "kernel": "training step", "achieved": tfs, "peak": tf, "unit": "TFLOP/s",
                            "frac": tfs / tf, "traffic": None, "peak_source": which},
               "e2e": None, "cpu_baseline": None}
        print(json.dumps(out), flush=True)


def run_stress(ctx, args):
    """configs[4]: L=1024, dense (W=1023) and banded (W=40) graphs, pair_stride=1, full clash, recompute mode."""
    from protein_ensemble_vae_b200 import EGNNDecoder, _lib
    from protein_ensemble_vae_b200 import losses as pl
    lib = _lib.lib()
    c = STRESS
    B, L = c["batch_per_gpu"], c["L"]
    host = synth_batch(B, L, c["z_g"], c["z_l"], seed=ctx.rank, pin=True)
    resident = {k: v.to(ctx.dev) for k, v in host.items()}
    tdih = pl.compute_dihedrals_from_coords(resident["target_N"], resident["target_CA"], resident["target_C"], resident["mask"])
    lw = dict(LOSS_W, pair_stride=c["pair_stride"])
    res = {}
    launches = 0
    steps = max(2, min(args.steps, 5))
    for tag, W in (("dense_W1023", 1023), ("band_W40", 40)):
        torch.manual_seed(0)
        dec = EGNNDecoder(c["z_g"], c["z_l"], hidden_dim=c["hidden"], num_layers=c["layers"], max_neighbors=W,
                          dropout=c["dropout"], precision="bf16", recompute_edges=True).to(ctx.dev).train()
        step = make_train_step(ctx, dec, lw)
        for _ in range(max(1, min(args.warmup, 2))):
            step(resident, tdih)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        n0 = lib.launch_count()
        ms = ctx.timed(lambda: step(resident, tdih), steps)
        launches += lib.launch_count() - n0
        res[tag] = {"ms_per_step": ms / steps, "conformers_per_s": B * ctx.world * steps / (ms / 1e3),
                    "edges_per_conformer": band_edges(L, W), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                    "tensor_tflops": flops_per_conformer(L, c["layers"], W) * B / (ms / steps * 1e-3) / 1e12}
        del dec, step
        torch.cuda.empty_cache()
    if ctx.rank == 0:
        hbm, tf, which = peaks()
        d = res["dense_W1023"]
        out = {"metric": "train_conformers_per_s", "value": d["conformers_per_s"], "unit": "conformers/s", "n_gpus": ctx.world,
               "steps": steps, "warmup": args.warmup, "ms_per_step": d["ms_per_step"], "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": workload_config(ctx.world, "stress"), "gpu_launches": int(launches), "variants": res,
               "roofline": {"bound": "tensor", "kernel": "training step (dense graph)", "achieved": d["tensor_tflops"], "peak": tf,
                            "unit": "TFLOP/s", "frac": d["tensor_tflops"] / tf, "traffic": None, "peak_source": which},
               "e2e": None, "cpu_baseline": None}
        print(json.dumps(out), flush=True)


def run_vae(ctx, args):
    """The whole HierCVAE training step (models/model.py:60-68 + models/training.py:89-102) at configs[1] sizes: encoder
    (TF32 tensor-core linears + attention kernels) -> decoder (bf16 edge MLP) -> compute_total_loss -> backward -> Adam on all
    20.4 M parameters.  SURVEY.md 8(f) N1; not the headline metric (BASELINE's is the decoder + loss step)."""
    from protein_ensemble_vae_b200 import DevicePrefetcher, EGNNDecoder, compute_total_loss
    from protein_ensemble_vae_b200 import losses as pl
    from protein_ensemble_vae_b200.encoder import ProteinEncoder
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    B, L, SD = CFG["batch_per_gpu"], CFG["L"], 1280
    torch.manual_seed(0)
    enc = ProteinEncoder(SD, z_g=CFG["z_g"], z_l=CFG["z_l"], dropout=CFG["dropout"]).to(dev).train()
    dec = EGNNDecoder(CFG["z_g"], CFG["z_l"], hidden_dim=CFG["hidden"], num_layers=CFG["layers"],
                      max_neighbors=CFG["max_neighbors"], dropout=CFG["dropout"], precision="bf16").to(dev).train()
    host = synth_batch(B, L, CFG["z_g"], CFG["z_l"], seed=rank, pin=True)
    for k in ("z_g", "z_l", "mu_g", "lv_g", "mu_l", "lv_l"):
        host.pop(k)                                             # the encoder produces them
    g = torch.Generator().manual_seed(1000 + rank)
    host["seq_emb"] = torch.randn(B, L, SD, generator=g).pin_memory()
    resident = {k: v.to(dev) for k, v in host.items()}
    params = list(enc.parameters()) + list(dec.parameters())
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    if world > 1:
        from protein_ensemble_vae_b200 import distributed as pdist

    def step(d):
        tdih = pl.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])
        z_g, z_l, mu_g, lv_g, mu_l, lv_l = enc(d["seq_emb"], d["target_N"], d["target_CA"], d["target_C"], tdih, d["mask"])
        outs = dec(z_g, z_l, d["mask"])
        res = compute_total_loss(outs[0], outs[1], outs[2], outs[3], d["target_N"], d["target_CA"], d["target_C"],
                                 d["labels"], d["mask"], mu_g, lv_g, mu_l, lv_l, tdih, **LOSS_W)
        res["total"].backward()
        if world > 1:
            pdist.allreduce_gradients(params)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return res["total"].detach()

    sampler = ClockSampler(ctx.local)
    if rank == 0 and not os.environ.get("PEV_BENCH_NO_CLOCKS"):
        sampler.start()
    for _ in range(args.warmup):
        step(resident)
    torch.cuda.synchronize()
    sampler.mark_start()
    from protein_ensemble_vae_b200 import _lib
    n0 = _lib.lib().launch_count()
    ms = ctx.timed(lambda: step(resident), args.steps)
    launches = _lib.lib().launch_count() - n0
    sampler.mark_stop()
    tdih = pl.compute_dihedrals_from_coords(resident["target_N"], resident["target_CA"], resident["target_C"], resident["mask"])
    enc_ms = [0.0, 0.0]
    for it in range(4):                                         # encoder alone, forward / backward (first pass = warm-up)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        lat = enc(resident["seq_emb"], resident["target_N"], resident["target_CA"], resident["target_C"], tdih, resident["mask"])
        ev[1].record()
        (lat[0].square().mean() + lat[1].square().mean() + lat[3].mean() + lat[5].mean()).backward()
        ev[2].record()
        torch.cuda.synchronize()
        enc.zero_grad(set_to_none=True)
        if it:
            enc_ms[0] += ev[0].elapsed_time(ev[1]) / 3
            enc_ms[1] += ev[1].elapsed_time(ev[2]) / 3
    del lat

    def e2e_steps(k):
        for d in DevicePrefetcher((host for _ in range(k)), dev):
            float(step(d))
    e2e_steps(2)
    ms_e2e = ctx.timed(lambda: e2e_steps(args.steps), 1)
    if rank == 0:
        _, tf, which = peaks()
        N = B * L
        # encoder GEMM flops per residue (forward): seq_proj, fusion, 7 x (QKV + out), 6 x feed-forward, K / V of the pooling,
        # local head; attention 4 L d per residue and attention layer
        lin = 2 * (SD * 256 + 512 * 512 + 7 * (512 * 1536 + 512 * 512) + 6 * 2 * 512 * 1024 + 512 * 1024 + 512 * 256 + 256 * 512)
        enc_flops = 3.0 * N * (lin + 7 * 4 * L * 512)
        dec_flops = flops_per_conformer(L, CFG["layers"]) * B
        step_tf = (enc_flops + dec_flops) / (ms / args.steps * 1e-3) / 1e12
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        out = {"metric": "train_conformers_per_s", "value": B * world * args.steps / (ms / 1e3), "unit": "conformers/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 edge MLP / tf32 encoder",
               "data": "synthetic",
               "config": {"workload": f"whole HierCVAE step: ProteinEncoder (6 transformer layers, d_model 512, ESM width {SD}) "
                                      f"+ EGNNDecoder (6 layers) + compute_total_loss + Adam, L={L}, {B} conformers per GPU",
                          "L": L, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                          "cache": "inputs_larger_than_l2 (335 MB of sequence embeddings per step)"},
               "clocks": sampler.summary(),
               "e2e": {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": "conformers/s",
                       "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches),
               "encoder": {"forward_ms": enc_ms[0], "backward_ms": enc_ms[1], "algorithmic_tflop": enc_flops / 1e12,
                           "tensor_frac": enc_flops / ((enc_ms[0] + enc_ms[1]) * 1e-3) / 1e12 / tf},
               "roofline": {"bound": "tensor", "kernel": "whole step", "achieved": step_tf, "peak": tf, "unit": "TFLOP/s",
                            "frac": step_tf / tf, "traffic": None, "peak_source": which},
               "cpu_baseline": None}
        print(json.dumps(out), flush=True)


def run_gpu(args):
    ctx = Ctx()
    try:
        {"train": run_train, "decode": run_decode, "mixed": run_mixed, "stress": run_stress, "vae": run_vae}[args.config](ctx, args)
    finally:
        ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="train", choices=["train", "decode", "mixed", "stress", "vae"],
                    help="train = BASELINE configs[1] (the headline); decode / mixed / stress = configs[3] / [2] / [4]; "
                         "vae = encoder + decoder + loss (SURVEY.md 8f N1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--train-only", action="store_true",
                    help="profiling runs: only the timed training steps (no e2e / decode / per-kernel side measurements)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
