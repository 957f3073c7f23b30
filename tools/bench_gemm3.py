#!/usr/bin/env python3
"""3xTF32 tensor-core GEMMs of the exact path against the library fp32 GEMMs they replace, at config-2 sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_ensemble_vae_b200 import egnn_tc

torch.backends.cuda.matmul.allow_tf32 = False


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for M in (65536, int(sys.argv[1]) if len(sys.argv) > 1 else 1205760):
    A = torch.randn(M, 256, device="cuda")
    W = torch.randn(256, 256, device="cuda") / 16
    G = torch.randn(M, 256, device="cuda")
    W3 = egnn_tc.split_weight(W)
    ms3 = t(lambda: egnn_tc.node_gemm3(A, W3))
    ms1 = t(lambda: egnn_tc.node_gemm(egnn_tc.EPI_PLAIN, A, W))
    msf = t(lambda: A @ W.t())
    mw3 = t(lambda: egnn_tc.node_wgrad3(G, A))
    mw1 = t(lambda: egnn_tc.node_wgrad(G, A))
    mwf = t(lambda: G.t() @ A)
    gf = 2 * M * 256 * 256 / 1e9
    print(f"M={M}: gemm 3xTF32 {ms3:.3f} ms ({gf / ms3:.0f} TF/s eff, {M * 2048 / ms3 / 1e6:.0f} GB/s) | TF32 {ms1:.3f} | "
          f"library fp32 {msf:.3f} ms;  wgrad 3xTF32 {mw3:.3f} | TF32 {mw1:.3f} | library fp32 {mwf:.3f} ms")
