import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_err(a, b):
    """max|a-b| / max|b|  -- the tolerance metric used everywhere."""
    import numpy as np
    a = np.asarray(a.detach().cpu() if hasattr(a, "detach") else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if hasattr(b, "detach") else b, dtype=np.float64)
    den = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
    return float(np.abs(a - b).max()) / den if b.size else 0.0


def assert_close_cond(mine, gold64, oracle32=None, tol=1e-5, slack=8.0, what=""):
    """|mine - gold| <= tol * max|gold| element-wise, widened where float32 itself is ill-conditioned:
    `oracle32` is the SAME formula evaluated by the oracle in float32; where that deviates from the
    float64 value (e.g. d sin/d cos = -c / sqrt(1 - c^2 + 1e-8) near |c| = 1) the bound grows to
    `slack` times that deviation."""
    import numpy as np
    to_np = lambda a: np.asarray(a.detach().cpu() if hasattr(a, "detach") else a, dtype=np.float64)  # noqa: E731
    mine, gold64 = to_np(mine), to_np(gold64)
    bound = tol * max(float(np.abs(gold64).max()) if gold64.size else 0.0, 1e-30) * np.ones_like(gold64)
    if oracle32 is not None:
        bound = bound + slack * np.abs(to_np(oracle32) - gold64)
    bad = np.abs(mine - gold64) > bound
    assert not bad.any(), (what, int(bad.sum()), float(np.abs(mine - gold64).max()), float(bound.min()))


def assert_named_close(mine, ref, tol=1e-5, max_outliers=2, outlier_tol=1e-2):
    """Compare dicts of arrays element-wise at `tol` * max|ref|.  Up to `max_outliers` elements per
    tensor may miss `tol` (but not `outlier_tol`): a ReLU / Huber / clamp kink whose argument is within
    float32 rounding of its threshold flips a whole gradient contribution (seen: 1 element of 512)."""
    import numpy as np
    assert set(mine) == set(ref), set(mine) ^ set(ref)
    for k in sorted(ref):
        a = np.asarray(mine[k].detach().cpu() if hasattr(mine[k], "detach") else mine[k], dtype=np.float64)
        b = np.asarray(ref[k].detach().cpu() if hasattr(ref[k], "detach") else ref[k], dtype=np.float64)
        assert a.shape == b.shape, k
        scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
        err = np.abs(a - b) / scale
        allowed = max_outliers + int(0.002 * err.size)      # one flipped unit touches a whole weight row
        assert int((err > tol).sum()) <= allowed, (k, int((err > tol).sum()), float(err.max()))
        assert float(err.max()) if err.size else 0.0 <= outlier_tol, (k, float(err.max()))


def assert_named_close_l2(mine, ref, tol):
    """Per-tensor relative L2 error -- the meaningful metric for the bf16 path, where single ReLU / rounding
    flips move individual elements by O(1) of their value but not the gradient as a whole."""
    import numpy as np
    assert set(mine) == set(ref), set(mine) ^ set(ref)
    worst = ("", 0.0)
    for k in sorted(ref):
        a = np.asarray(mine[k].detach().cpu() if hasattr(mine[k], "detach") else mine[k], dtype=np.float64)
        b = np.asarray(ref[k].detach().cpu() if hasattr(ref[k], "detach") else ref[k], dtype=np.float64)
        err = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < tol, worst


class Backend:
    """'host': CPU tensors against tests/hostcheck (kernel bodies compiled with g++);
    'cuda': CUDA tensors against libpev_b200.so -- the product path."""

    def __init__(self, kind):
        self.kind = kind
        self.dev = "cpu" if kind == "host" else "cuda"

    def ctx(self):
        import contextlib
        if self.kind == "host":
            from hostlib import host_backend
            return host_backend()
        return contextlib.nullcontext()

    def t32(self, a):
        import numpy as np
        import torch
        return torch.tensor(np.asarray(a, dtype=np.float32), device=self.dev)

    def ti(self, a):
        import torch
        return torch.as_tensor(a).to(self.dev)


@pytest.fixture(params=["host", pytest.param("cuda", marks=pytest.mark.gpu)])
def bk(request):
    return Backend(request.param)
