"""The C-ABI library loads on a machine without a GPU and exports every symbol include/pev_b200.h declares."""
import ctypes
import os
import re
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pev_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pev_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def so_path():
    from protein_ensemble_vae_b200 import build
    if not os.path.exists(build.OUT) or build.needs_build():
        if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
            pytest.skip("nvcc not available and libpev_b200.so not prebuilt")
        build.build()
    return build.OUT


def test_header_symbols_are_exported(so_path):
    cdll = ctypes.CDLL(so_path)
    names = declared_symbols()
    assert len(names) >= 19
    missing = [n for n in names if not hasattr(cdll, n)]
    assert not missing, missing
    cdll.pev_abi_version.restype = ctypes.c_int
    assert cdll.pev_abi_version() == 1


def test_python_binding_matches_header(so_path):
    from protein_ensemble_vae_b200 import _lib
    bound = set(_lib._PROTOS) | set(_lib._PROTOS_TC)
    assert bound == set(declared_symbols()), bound ^ set(declared_symbols())
    lib = _lib.Lib(so_path)            # resolves every prototype; no compute call without a GPU
    assert lib.launch_count() == 0


def test_product_fails_loudly_without_library(monkeypatch):
    from protein_ensemble_vae_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpev_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()
