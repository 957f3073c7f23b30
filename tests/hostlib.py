"""Test seam: run the package's Python host logic on CPU tensors against tests/hostcheck
(the kernels' per-thread bodies compiled with g++).  TESTS ONLY -- the product never does this."""
import contextlib
import os
import subprocess

from protein_ensemble_vae_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
SO = os.path.join(HERE, "hostcheck", "libpev_hostcheck.so")


def build_hostcheck(force=False):
    csrc = os.path.join(os.path.dirname(HERE), "protein_ensemble_vae_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17",
                               "-o", SO, SRC])
    return SO


@contextlib.contextmanager
def host_backend():
    so = build_hostcheck()
    old = (_lib._LIB, _lib._REQUIRE_CUDA)
    _lib._LIB, _lib._REQUIRE_CUDA = _lib.Lib(so, require_tc=False), False
    try:
        yield
    finally:
        _lib._LIB, _lib._REQUIRE_CUDA = old
