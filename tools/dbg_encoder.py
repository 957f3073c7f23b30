#!/usr/bin/env python3
"""Per-parameter gradient errors of the encoder (TF32 path, and a library-TF32 yardstick) against the reference fixtures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import cases, test_encoder as te
from protein_ensemble_vae_b200 import tc_linear

gold = np.load(os.path.join(ROOT, "tests", "golden", "encoders.npz"))
tag = sys.argv[1] if len(sys.argv) > 1 else "enc2_gaps"
for mode in ("tf32", "lib-tf32", "lib-fp32"):
    if mode.startswith("lib"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "lib-tf32"
        orig = tc_linear.supported
        tc_linear.supported = lambda x, W: False
    mask, res, z, grads, tb = te._run(tag, "tf32")
    if mode.startswith("lib"):
        tc_linear.supported = orig
        torch.backends.cuda.matmul.allow_tf32 = False
    gerrs = tb._grad_errors(grads, gold, tag)
    top = sorted(gerrs.items(), key=lambda kv: -kv[1][1])[:8]
    print(mode, [(k, "%.1e" % v[1]) for k, v in top])
