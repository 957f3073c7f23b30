#!/usr/bin/env python3
"""Accuracy table of the 3xTF32 kernels (max |err| / max |ref| against float64) beside plain TF32 and the library fp32 GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_ensemble_vae_b200 import egnn_tc

torch.backends.cuda.matmul.allow_tf32 = False
err = lambda a, r: float((a.double() - r).abs().max() / r.abs().max())  # noqa: E731
print("gemm   M      K    Nout | 3xTF32    TF32      library fp32")
for M, K, Nout in ((1000, 256, 256), (77, 256, 512), (4101, 512, 256), (300, 768, 512), (1205760, 256, 256)):
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g) * torch.exp(torch.randn(M, 1, device="cuda", generator=g))
    W = torch.randn(Nout, K, device="cuda", generator=g) / K ** 0.5
    ref = A.double() @ W.double().t()
    print(f"{M:9d} {K:5d} {Nout:5d} | {err(egnn_tc.node_gemm3(A, egnn_tc.split_weight(W)), ref):.2e}  "
          f"{err(egnn_tc.node_gemm(egnn_tc.EPI_PLAIN, A, W)[0], ref):.2e}  {err(A @ W.t(), ref):.2e}")
    del A, ref
print("wgrad  N      Mo        | 3xTF32    TF32      library fp32")
for N, Mo in ((5000, 256), (33, 512), (65553, 512), (1205760, 256), (4823040, 256)):
    g = torch.Generator(device="cuda").manual_seed(N)
    G = torch.randn(N, Mo, device="cuda", generator=g)
    X = torch.randn(N, 256, device="cuda", generator=g) + 0.5
    ref = G.double().t() @ X.double()
    print(f"{N:9d} {Mo:5d}       | {err(egnn_tc.node_wgrad3(G, X), ref):.2e}  {err(egnn_tc.node_wgrad(G, X), ref):.2e}  "
          f"{err(G.t() @ X, ref):.2e}")
    del G, X, ref
