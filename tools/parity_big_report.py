#!/usr/bin/env python3
"""Error table of the BASELINE-shape parity cases (tests/test_gpu_parity_big.py) on the GPU: outputs and gradients of the
fp32 path, the tcgen05 bf16 path and the autocast-style yardstick (tests/bf16_yardstick.py) against the
reference-generated fixtures; `smooth` = gradients of the CA-only loss (no ReLU between the loss and the EGNN layers)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import cases  # noqa: E402
import test_gpu_parity_big as tb  # noqa: E402
from bf16_yardstick import emulate_autocast_edge_mlp  # noqa: E402
from conftest import Backend, rel_err  # noqa: E402

gold = np.load(os.path.join(ROOT, "tests", "golden", "decoders_big.npz"))
bk = Backend("cuda")
only = sys.argv[1:] or ["fp32", "bf16", "autocast-emu"]
for prec in only:
    for tag in cases.BIG_DECODER_CASES:
        for smooth in (False, True):
            if smooth and f"{tag}.sm.grad.z_g" not in gold and f"{tag}.sm.gproj1.z_g" not in gold:
                continue
            with emulate_autocast_edge_mlp():
                case, mask, outs, grads = tb._run(tag, prec, bk, smooth=smooth)
            oe = {n: round(rel_err(o.detach(), gold[f"{tag}.{n}"]), 7) for n, o in zip(("N", "CA", "C", "logits"), outs)}
            errs = tb._grad_errors(grads, gold, tag + (".sm" if smooth else ""))
            top = sorted(errs.items(), key=lambda kv: -kv[1][1])[:4]
            print(prec, tag, "smooth" if smooth else "full", oe)
            for k, v in top:
                print("     grad", k, "max %.2e  l2 %.2e" % v)
            lay = max(v[1] for k, v in errs.items() if k.startswith("layers."))
            oth = max(v[1] for k, v in errs.items() if not k.startswith("layers."))
            print("     worst layer l2 %.3e, worst other l2 %.3e" % (lay, oth), flush=True)
