#!/usr/bin/env python3
"""Fused CTA-pair forward (pev_edge3_fwd) against a torch emulation with the same rounding points, and timing against
the two-kernel form (pev_edge2_fwd1 + fwd2).  Run under `timeout` on the GPU box."""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from protein_ensemble_vae_b200 import _lib, egnn_tc2 as T2  # noqa: E402
from protein_ensemble_vae_b200._lib import ptr, stream  # noqa: E402
from protein_ensemble_vae_b200.graph import band_graph  # noqa: E402

H, BF = 256, torch.bfloat16
bf = lambda t: t.to(BF).float()  # noqa: E731
silu2 = lambda h: h + h * torch.tanh(h)  # noqa: E731


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def setup(lengths, W, seed):
    g = band_graph(lengths, W, "cuda", cache=False)
    gen = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s, k=1.0: torch.randn(*s, device="cuda", generator=gen) * k  # noqa: E731
    N = g.num_nodes
    return dict(g=g, N=N, E=g.num_edges, x=r(N, 3, k=2.0), ABh=r(N, 2 * H, k=0.5).to(torch.float16), wd=r(H, k=0.02),
                W2=r(H, H, k=1 / 16), W5=r(H, H, k=1 / 16), b2=r(H, k=0.1), b5=r(H, k=0.1), w6=r(H, k=0.1), b6=r(1))


def run3(c, train):
    g, N, E = c["g"], c["N"], c["E"]
    L, st = _lib.lib(), stream(c["x"])
    d2 = torch.empty(max(E, 1), device="cuda")
    L.call("pev_edge_d2", ptr(c["x"]), ptr(g.row), ptr(g.col), E, ptr(d2), st)
    agg = torch.full((N, H), 7.0, device="cuda")
    w = torch.full((max(E, 1),), 7.0, device="cuda")
    hv = torch.zeros(E, H, dtype=BF, device="cuda") if train else None
    hs = torch.zeros(E, H, dtype=BF, device="cuda") if train else None
    W2p, W5p = T2.packed_weight_scaled(c["W2"], 0.5), T2.packed_weight_scaled(c["W5"], 0.5)   # keep both alive
    L.call("pev_edge3_fwd", ptr(c["ABh"]), ptr(d2), ptr(c["wd"]), ptr(W2p), ptr(c["b2"]), ptr(W5p), ptr(c["b5"]),
           ptr(c["w6"]), ptr(c["b6"]), ptr(g.row), ptr(g.col), N, E, ptr(hv), ptr(hs), ptr(agg), ptr(w), st)
    torch.cuda.synchronize()
    return d2, agg, w, hv, hs


def check(lengths, W):
    c = setup(lengths, W, 5)
    g, N, E = c["g"], c["N"], c["E"]
    row, col = g.row.long(), g.col.long()
    d2, agg, w, hv, hs = run3(c, True)
    d2 = d2[:E]
    hu = (c["ABh"][row, :H] + c["ABh"][col, H:]).float() + 0.5 * c["wd"] * d2[:, None]
    hv_ref = bf(silu2(hu)) @ bf(0.5 * c["W2"]).t() + 0.5 * c["b2"]
    m_ref = silu2(hv_ref)
    agg_ref = torch.zeros(N, H, device="cuda").index_add_(0, row, m_ref)
    hs_ref = bf(m_ref) @ bf(0.5 * c["W5"]).t() + 0.5 * c["b5"]
    w_ref = silu2(hs_ref) @ c["w6"] + c["b6"]
    errs = dict(hv=rel(hv.float(), hv_ref), agg=rel(agg, agg_ref), hs=rel(hs.float(), hs_ref), w=rel(w[:E], w_ref))
    _, agg2, w2, _, _ = run3(c, False)
    errs["agg_inf"] = rel(agg2, agg)
    errs["w_inf"] = rel(w2[:E], w[:E])
    ok = errs["hv"] < 6e-3 and errs["agg"] < 2e-3 and errs["hs"] < 8e-3 and errs["w"] < 8e-3 and errs["agg_inf"] < 1e-5 \
        and errs["w_inf"] < 1e-5
    print("OK  " if ok else "FAIL", lengths, W, "E=%d" % E, {k: "%.2e" % v for k, v in errs.items()}, flush=True)
    return ok


def bench(B=64):
    c = setup((256,) * B, 40, 1)
    g, N, E = c["g"], c["N"], c["E"]
    L, st = _lib.lib(), stream(c["x"])
    d2 = torch.empty(E, device="cuda")
    L.call("pev_edge_d2", ptr(c["x"]), ptr(g.row), ptr(g.col), E, ptr(d2), st)
    W2p, W5p = T2.packed_weight_scaled(c["W2"], 0.5), T2.packed_weight_scaled(c["W5"], 0.5)
    agg = torch.empty(N, H, device="cuda")
    w = torch.empty(E, device="cuda")
    hv = torch.empty(E, H, dtype=BF, device="cuda")
    hs = torch.empty(E, H, dtype=BF, device="cuda")
    hvT, mT = T2.alloc_tile_image(E, "cuda"), T2.alloc_tile_image(E, "cuda")

    def v3(train):
        L.call("pev_edge3_fwd", ptr(c["ABh"]), ptr(d2), ptr(c["wd"]), ptr(W2p), ptr(c["b2"]), ptr(W5p), ptr(c["b5"]),
               ptr(c["w6"]), ptr(c["b6"]), ptr(g.row), ptr(g.col), N, E, ptr(hv) if train else None,
               ptr(hs) if train else None, ptr(agg), ptr(w), st)

    def v2(train):
        L.call("pev_edge2_fwd1", ptr(c["ABh"]), ptr(d2), ptr(c["wd"]), ptr(W2p), ptr(c["b2"]), ptr(g.row), ptr(g.col), N, E,
               ptr(hvT) if train else None, ptr(mT), ptr(agg), st)
        L.call("pev_edge2_fwd2", ptr(mT), ptr(W5p), ptr(c["b5"]), ptr(c["w6"]), ptr(c["b6"]), E, ptr(w),
               ptr(hs) if train else None, st)

    for name, fn in (("v3 train", lambda: v3(True)), ("v3 infer", lambda: v3(False)), ("v2 train", lambda: v2(True)),
                     ("v2 infer", lambda: v2(False))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"{name}: {ms:.3f} ms for E={E} ({ms * 4823040 / E:.3f} ms at config 2)", flush=True)


def ablate():
    """Role ablation of the inference kernel (library built with -DPEV_TC3_ABLATE; PEV_TC3_ABL selects the variant)."""
    import os
    c = setup((256,) * 64, 40, 1)
    g, N, E = c["g"], c["N"], c["E"]
    L, st = _lib.lib(), stream(c["x"])
    d2 = torch.empty(E, device="cuda")
    L.call("pev_edge_d2", ptr(c["x"]), ptr(g.row), ptr(g.col), E, ptr(d2), st)
    W2p, W5p = T2.packed_weight_scaled(c["W2"], 0.5), T2.packed_weight_scaled(c["W5"], 0.5)
    agg = torch.empty(N, H, device="cuda")
    w = torch.empty(E, device="cuda")
    names = {0: "full", 1: "E1 no segsum", 2: "E1 no tanh", 3: "E1 no segsum/tanh", 4: "P no gathers", 12: "P no gathers/tanh",
             16: "E2 no tanh", 19: "E1+E2 light", 32: "no MMA", 31: "all CUDA-core roles light", 63: "handshakes only",
             15: "P+E1 light", 28: "P+E2 light"}
    for abl, name in names.items():
        os.environ["PEV_TC3_ABL"] = str(abl)

        def fn():
            L.call("pev_edge3_fwd", ptr(c["ABh"]), ptr(d2), ptr(c["wd"]), ptr(W2p), ptr(c["b2"]), ptr(W5p), ptr(c["b5"]),
                   ptr(c["w6"]), ptr(c["b6"]), ptr(g.row), ptr(g.col), N, E, None, None, ptr(agg), ptr(w), st)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"abl {abl:3d} {name:28s}: {ms * 4823040 / E:.3f} ms at config 2", flush=True)


if __name__ == "__main__":
    if "--ablate" in sys.argv:
        ablate()
        sys.exit(0)
    cases = [((100,), 40), ((7, 130, 64), 40), ((256,) * 3, 40), ((33,), 5), ((3, 2), 1), ((90, 41), 3), ((256,) * 40, 40)]
    good = all([check(*c) for c in cases])
    if good and "--bench" in sys.argv:
        bench()
    sys.exit(0 if good else 1)
