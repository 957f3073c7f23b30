#!/usr/bin/env python3
"""Benchmark of the hot path: EGNN decoder fwd+bwd + compute_total_loss fwd+bwd (train) and
decoder fwd + Kabsch RMSD (decode), on synthetic backbone ensembles.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  ``value`` = train conformers/s with the batch resident in HBM
(BASELINE.json metric, configs[1]: L=256, 6 EGNN layers, 256 conformers per GPU, bf16 edge MLP);
``e2e`` = the same step fed from pinned host buffers through the public module / loss API with the
H2D copies and the loss read-back inside the timed region; ``decode`` = decoded conformers/s
(configs[3] shape: L=100, 8 layers, + Kabsch RMSD against one reference structure).
``--impl reference`` times the CPU restatement of the reference (``oracle/``) on the host cores for
the same metric/config on a bounded sample (the reference itself is Python and cannot travel).
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
DECODE = dict(L=100, layers=8, chunk=2048, z_g=512, z_l=256)


def synth_batch(B, L, z_g, z_l, seed, device="cpu", pin=False):
    """Synthetic inputs of SURVEY.md 8(d): latents, random-walk targets, labels, posterior stats."""
    g = torch.Generator().manual_seed(seed)
    ca = torch.cumsum(torch.randn(B, L, 3, generator=g) * 2.2, 1)
    ca = ca - ca.mean(1, keepdim=True)
    d = dict(
        z_g=torch.randn(B, z_g, generator=g), z_l=torch.randn(B, L, z_l, generator=g),
        target_CA=ca, target_N=ca + 0.8 * torch.randn(B, L, 3, generator=g),
        target_C=ca + 0.8 * torch.randn(B, L, 3, generator=g),
        labels=torch.randint(0, 20, (B, L), generator=g), mask=torch.ones(B, L),
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


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_step_fn(B, cfg):
    """One train step of the CPU restatement (oracle port, fp32): decoder fwd + total loss + backward."""
    import synth
    from oracle import egnn_oracle, losses_oracle
    params = synth.make_params(synth.decoder_param_shapes(cfg["z_g"], cfg["z_l"], cfg["hidden"], cfg["layers"]), 0)
    sd = {k: torch.tensor(v).requires_grad_() for k, v in params.items()}
    d = synth_batch(B, cfg["L"], cfg["z_g"], cfg["z_l"], seed=0)
    tdih = losses_oracle.compute_dihedrals_from_coords(d["target_N"], d["target_CA"], d["target_C"], d["mask"])
    cache = {}

    def step():
        for v in sd.values():
            v.grad = None
        n, ca, c, lg = egnn_oracle.egnn_decoder(sd, d["z_g"], d["z_l"], d["mask"], max_neighbors=cfg["max_neighbors"],
                                                edge_cache=cache)
        res = losses_oracle.compute_total_loss(n, ca, c, lg, d["target_N"], d["target_CA"], d["target_C"], d["labels"],
                                               d["mask"], d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih, **LOSS_W)
        res["total"].backward()
        return float(res["total"].detach())
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    step = cpu_step_fn(B, CFG)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    sample = f"{B} conformers per step (of {CFG['batch_per_gpu']}), L={CFG['L']}, {CFG['layers']} layers, fp32, oracle port"
    out = {"impl": "reference", "metric": "train_conformers_per_s", "value": val, "unit": "conformers/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
           "config": workload_config(args.gpus),
           "cpu_baseline": {"value": val, "unit": "conformers/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": val, "unit": "conformers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(n_gpus):
    return {"workload": "configs[1]: single protein L=256, 6 EGNN layers (hidden 256, W=40), 256 conformers per GPU, "
                        "decoder fwd+bwd + compute_total_loss fwd+bwd + Adam step, bf16 edge MLP",
            "L": CFG["L"], "layers": CFG["layers"], "batch_per_gpu": CFG["batch_per_gpu"],
            "global_batch": CFG["batch_per_gpu"] * n_gpus, "parallelism": f"dp{n_gpus}",
            "cache": "inputs_larger_than_l2 (603 MB of latents/posteriors + 7.4 GB of per-edge bf16 streams per layer)"}


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch.distributed as dist
    from protein_ensemble_vae_b200 import EGNNDecoder, _lib, compute_total_loss, kabsch_rmsd_batch
    from protein_ensemble_vae_b200 import distributed as pdist
    from protein_ensemble_vae_b200 import losses as pl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    B, L = CFG["batch_per_gpu"], CFG["L"]

    torch.manual_seed(0)
    dec = EGNNDecoder(CFG["z_g"], CFG["z_l"], hidden_dim=CFG["hidden"], num_layers=CFG["layers"],
                      max_neighbors=CFG["max_neighbors"], dropout=CFG["dropout"], precision="bf16").to(dev).train()
    params = [p for p in dec.parameters()]
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    host = synth_batch(B, L, CFG["z_g"], CFG["z_l"], seed=rank, pin=True)
    resident = {k: v.to(dev) for k, v in host.items()}
    tdih = pl.compute_dihedrals_from_coords(resident["target_N"], resident["target_CA"], resident["target_C"],
                                            resident["mask"])
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def step(d):
        outs = dec(d["z_g"], d["z_l"], d["mask"])
        res = compute_total_loss(outs[0], outs[1], outs[2], outs[3], d["target_N"], d["target_CA"], d["target_C"],
                                 d["labels"], d["mask"], d["mu_g"], d["lv_g"], d["mu_l"], d["lv_l"], tdih, **LOSS_W)
        res["total"].backward()
        if world > 1:       # data parallel over conformers: average the decoder gradients (17.8 MB) over NVLink
            pdist.allreduce_gradients(params)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return res["total"].detach()

    def timed(fn, steps):
        """max over ranks of the CUDA-event time of `steps` calls, barrier + synchronize on both sides."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("PEV_BENCH_NO_CLOCKS"):   # set under ncu: it would follow the child process
        sampler.start()
    for _ in range(args.warmup):
        step(resident)
    torch.cuda.synchronize()
    sampler.mark_start()
    _lib.PROFILE = {}
    n0 = lib.launch_count()
    ms = timed(lambda: step(resident), args.steps)
    launches = lib.launch_count() - n0
    prof, _lib.PROFILE = _lib.PROFILE, None
    sampler.mark_stop()
    value = B * world * args.steps / (ms / 1e3)

    # end to end: pinned host buffers -> device every step (DevicePrefetcher: the copy of step i+1 runs on a side
    # stream under step i's kernels), loss read back to the host every step; K copies and K read-backs per K steps
    from protein_ensemble_vae_b200 import DevicePrefetcher

    def e2e_steps(k):
        for d in DevicePrefetcher((host for _ in range(k)), dev):
            float(step(d))
    e2e_steps(2)

    def timed_e2e(k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e2e_steps(k)
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    ms_e2e = timed_e2e(args.steps)
    e2e = B * world * args.steps / (ms_e2e / 1e3)

    # decode: decoder forward (no grad) + Kabsch RMSD of every sample against one reference CA trace
    dd = DECODE
    dec8 = EGNNDecoder(dd["z_g"], dd["z_l"], hidden_dim=256, num_layers=dd["layers"], max_neighbors=40, dropout=0.1,
                       precision="bf16").to(dev).eval()
    S = dd["chunk"]
    zg = torch.randn(S, dd["z_g"], device=dev)
    zl = torch.randn(S, dd["L"], dd["z_l"], device=dev)
    mask1 = torch.ones(S, dd["L"], device=dev)
    ref_ca = resident["target_CA"][0, :dd["L"]].contiguous()

    def decode_step():
        with torch.no_grad():
            n, ca, c, lg = dec8(zg, zl, mask1)
            return kabsch_rmsd_batch(ca, ref_ca)
    for _ in range(2):
        decode_step()
    dsteps = max(3, args.steps // 2)
    ms_dec = timed(decode_step, dsteps)
    decode = S * world * dsteps / (ms_dec / 1e3)

    if rank == 0:
        hbm, tf, which = peaks()
        E = lib_edges(L) * B
        # Per-kernel rooflines of the tcgen05 edge kernels, timed live with CUDA events inside the timed region
        # (_lib.profiled).  Algorithmic bytes per edge = the bf16 [E,256] streams a kernel must read / write (512 B
        # each) + the per-edge scalars; algorithmic flops per edge = 2 * 256 * 256 per GEMM.
        # "traffic" of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum per launch from the
        # ncu --set full capture in profiles/ (scaled from its B=32 launch by the edge count).
        KSPEC = {  # tag: (kernel, streams read+written, extra bytes per edge, tcgen05 GEMMs, tanh per edge-feature)
            "edge2_fwd1": ("fwd1_kernel", 2, 16, 1, 2),
            "edge2_fwd2": ("fwd2_kernel", 2, 4, 1, 1),
            "edge2_bwd2": ("bwd2_kernel", 3, 8, 1, 2),
            "edge2_bwd1": ("bwd1_kernel", 2, 20, 1, 1),
            "edge2_wgrad5": ("wgrad_kernel<5>", 2, 4, 1, 1),
            "edge2_wgrad2": ("wgrad_kernel<2>", 1, 12, 1, 1),
            "edge2_sums": ("edge_sums_kernel", 2, 8, 0, 0),
        }
        kernels = {}
        for tag, (kname, streams, extra, gemms, tanhs) in KSPEC.items():
            ev = prof.get(tag, []) if prof else []
            if not ev:
                continue
            tot = sum(s.elapsed_time(e) for s, e in ev)
            per = tot / len(ev)
            nbytes = (512.0 * streams + extra) * E
            kernels[tag] = {"kernel": kname, "ms_per_launch": per, "launches_timed": len(ev), "share_of_step": tot / ms,
                            "algorithmic_gb": nbytes / 1e9, "hbm_gbs": nbytes / (per * 1e-3) / 1e9,
                            "hbm_frac": nbytes / (per * 1e-3) / 1e9 / hbm,
                            "tensor_tflops": gemms * 2.0 * E * 256 * 256 / (per * 1e-3) / 1e12,
                            "tensor_frac": gemms * 2.0 * E * 256 * 256 / (per * 1e-3) / 1e12 / tf,
                            "mufu_frac": tanhs * 256.0 * E / (per * 1e-3) / (148 * 16 * 1.965e9)}
        dom = max(kernels, key=lambda t: kernels[t]["share_of_step"]) if kernels else None
        roof = {"bound": "hbm", "kernel": None, "achieved": None, "peak": hbm, "unit": "GB/s", "frac": None,
                "traffic": None, "peak_source": which, "kernels": kernels}
        if dom:
            k = kernels[dom]
            roof.update({"kernel": f"{k['kernel']} ({dom})", "achieved": k["hbm_gbs"], "frac": k["hbm_frac"],
                         "ms_per_launch": k["ms_per_launch"], "launches_timed": k["launches_timed"],
                         "share_of_step": k["share_of_step"], "tensor_frac": k["tensor_frac"],
                         "traffic": NCU_TRAFFIC_PER_EDGE.get(dom, 0.0) * E or None,
                         "traffic_source": "profiles/r01_edge2_kernels_ncu.md"})
        cpu = None
        if not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            Bc = 2
            cstep = cpu_step_fn(Bc, CFG)
            cstep()
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 12):
                cstep()
                reps += 1
            dtc = time.perf_counter() - t0
            cpu = {"value": Bc * reps / dtc, "unit": "conformers/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{reps} steps of {Bc} conformers (of {B}), L={L}, {CFG['layers']} layers, fp32 oracle port"}
        out = {"metric": "train_conformers_per_s", "value": value, "unit": "conformers/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": workload_config(world), "clocks": sampler.summary(),
               "e2e": {"value": e2e, "unit": "conformers/s", "h2d_bytes_per_step": h2d_bytes * world,
                       "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
               "decode": {"metric": "decoded_conformers_per_s", "value": decode, "unit": "conformers/s",
                          "config": {"workload": "configs[3] shape: L=100, 8 EGNN layers, decoder fwd + Kabsch RMSD vs one "
                                                 "reference, latent samples sharded across GPUs", "chunk_per_gpu": S},
                          "ms_per_chunk": ms_dec / dsteps}}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per edge of each v2 edge kernel (ncu --set full, profiles/)
NCU_TRAFFIC_PER_EDGE = {"edge2_fwd1": 980.4, "edge2_fwd2": 956.6, "edge2_bwd2": 1470.9, "edge2_wgrad5": 1035.8,
                        "edge2_bwd1": 974.9, "edge2_wgrad2": 544.8, "edge2_sums": 774.0}


def lib_edges(L, W=40):
    from protein_ensemble_vae_b200.graph import band_edge_count
    return band_edge_count(L, W)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
