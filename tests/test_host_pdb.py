"""Streaming ensemble PDB writer: the file written from device tensors must equal, byte for byte, what
the reference's ``write_pdb`` appends model by model (tests/golden/ensemble*.pdb, produced by the reference function)."""
import os

import numpy as np
import pytest
import torch

import cases

G = os.path.join(os.path.dirname(__file__), "golden")


def _write(bk, tmp_path, name, **kw):
    from protein_ensemble_vae_b200 import generation as pg
    path = str(tmp_path / name)
    with bk.ctx():
        nbytes = pg.write_ensemble_pdb(path, **kw)
    data = open(path, "rb").read()
    assert nbytes == len(data)
    return data


def test_pdb_writer_is_byte_identical_to_the_reference(bk, tmp_path):
    n, ca, c, mask, seq = cases.pdb_inputs()
    t = lambda a: bk.t32(a)  # noqa: E731
    got = _write(bk, tmp_path, "a.pdb", coords_n=t(n), coords_ca=t(ca), coords_c=t(c), mask=t(mask), sequence=seq, pdb_id="1abc",
                 chain_id="B", title="synthetic ensemble", chunk=2)             # 3 models in chunks of 2 + 1
    ref = open(os.path.join(G, "ensemble.pdb"), "rb").read()
    assert got == ref
    got2 = _write(bk, tmp_path, "b.pdb", coords_n=t(n[:2]), coords_ca=t(ca[:2]), coords_c=t(c[:2]), mask=t(np.ones_like(mask)))
    assert got2 == open(os.path.join(G, "ensemble_plain.pdb"), "rb").read()


def test_pdb_writer_edge_cases(bk, tmp_path):
    n, ca, c, mask, seq = cases.pdb_inputs()
    t = lambda a: bk.t32(a)  # noqa: E731
    empty = _write(bk, tmp_path, "e.pdb", coords_n=t(n[:1]), coords_ca=t(ca[:1]), coords_c=t(c[:1]), mask=t(np.zeros_like(mask)))
    assert empty.endswith(b"MODEL        1\n\nTER\nENDMDL\n")
    big = ca.copy()
    big[0, 2, 1] = 12345.0
    from protein_ensemble_vae_b200 import generation as pg
    with pytest.raises(ValueError), bk.ctx():
        pg.write_ensemble_pdb(str(tmp_path / "x.pdb"), t(n), t(big), t(c), t(mask))


def test_pdb_block_offsets_follow_the_model_number_width():
    """MODEL lines widen past model 9999 (Python's %4d): offsets of 12 000 models against a direct count."""
    import ctypes
    import hostlib
    lib = ctypes.CDLL(hostlib.build_hostcheck())
    lib.pev_pdb_models_bytes.restype = ctypes.c_int64
    lib.pev_pdb_models_bytes.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
    nv = 7
    block = 15 + 4 * 81 * nv + 1 + 17 * (4 * nv - 1) + 11
    want = sum(block + max(0, len(str(m)) - 4) for m in range(1, 12001))
    assert lib.pev_pdb_models_bytes(1, 12000, nv) == want
    assert lib.pev_pdb_models_bytes(9990, 30, nv) == sum(block + max(0, len(str(m)) - 4) for m in range(9990, 10020))
