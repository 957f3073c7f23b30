"""ctypes binding of the C ABI in ``include/pev_b200.h`` (``csrc/`` -> ``libpev_b200.so``).

The product path is CUDA-only: :func:`lib` raises if the shared library cannot be
loaded and :func:`ptr` raises on a non-CUDA tensor.  There is no CPU fallback.
(``tests/`` may point :data:`_LIB` at ``tests/hostcheck`` -- the kernels' per-thread
bodies compiled for the host -- to exercise the Python host logic without a GPU;
nothing in this package does.)
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_uint32, POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PEV_B200_LIB") or os.path.join(HERE, "libpev_b200.so")   # override: A/B builds of csrc/
NUM_TERMS = 17

# enum pev_term
(T_REC_CA, T_REC_N, T_REC_C, T_PAIR, T_KL_G, T_KL_L, T_DIH_CONS, T_OMEGA, T_RAMA, T_BOND_NCA,
 T_BOND_CAC, T_BOND_CN, T_ANG_NCAC, T_ANG_CNCA, T_ANG_CACN, T_SEQ, T_CLASH) = range(NUM_TERMS)


class LossArgs(ctypes.Structure):
    """``struct pev_loss_args``."""
    _fields_ = [(n, c_void_p) for n in (
        "pred_N", "pred_CA", "pred_C", "target_N", "target_CA", "target_C", "mask", "target_dih",
        "logits", "labels", "mu_l", "lv_l", "mu_g", "lv_g")] + [
        (n, c_int32) for n in ("B", "L", "C", "D", "G", "pair_stride", "enable_clash", "enable_geometry")
    ] + [("clash_dist", c_float), ("soft_margin", c_float)]


_P, _I, _L = c_void_p, c_int32, c_int64
_PROTOS = {
    "pev_abi_version": (c_int32, []),
    "pev_last_error": (c_char_p, []),
    "pev_launch_count": (c_int64, []),
    "pev_band_graph_build": (c_int32, [_P, _P, _I, _I, _L, _P, _P, _P, _P, _P, _P]),
    "pev_edge_prologue_fwd": (c_int32, [_P, _P, _P, _P, _P, _P, _L, _I, _P, _P]),
    "pev_edge_prologue_bwd": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _I, _P, _P, _P, _P, _P]),
    "pev_scatter_coord_fwd": (c_int32, [_P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P]),
    "pev_scatter_coord_bwd": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _I, _P, _P, _P, _P]),
    "pev_loss_fwd": (c_int32, [POINTER(LossArgs), _P, _P, _P]),
    "pev_loss_finalize": (c_int32, [_P, _P, _I, _P, _P, _P]),
    "pev_loss_bwd": (c_int32, [POINTER(LossArgs), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pev_dihedrals_fwd": (c_int32, [_P, _P, _P, _P, _I, _I, _P, _P]),
    "pev_dihedrals_bwd": (c_int32, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "pev_dihedral_terms_fwd": (c_int32, [_P, _P, _P, _I, _I, _P, _P]),
    "pev_dihedral_terms_bwd": (c_int32, [_P, _P, _P, _P, _I, _I, _P, _P]),
    "pev_kabsch_rmsd": (c_int32, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "pev_validate_geometry": (c_int32, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "pev_superpose_scores": (c_int32, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pev_lddt": (c_int32, [_P, _P, _P, _I, _I, _I, _I, c_float, _P, _P, _P]),
    "pev_rmsf": (c_int32, [_P, _I, _I, _P, _P]),
    "pev_backbone_fwd": (c_int32, [_P, _I, _P, _I, _P, _P, _L, _P, _P, _P]),
    "pev_backbone_bwd": (c_int32, [_P, _I, _P, _I, _P, _P, _L, _P, _P, _P, _P, _P, _P]),
    "pev_pdb_models_bytes": (c_int64, [_L, _I, _I]),
    "pev_pdb_format_models": (c_int32, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _P, _P, _P]),
    "pev_unpack_center": (c_int32, [_P] * 8 + [_I, _I, _I, _I] + [_P] * 8),
}
# tensor-core entry points: present in libpev_b200.so only (no host restatement)
_PROTOS_TC = {
    # v2 edge pipeline (csrc/edge_tc2_kernels.cu)
    "pev_pack_weight_bf16_scaled": (c_int32, [_P, _I, c_float, _P, _P]),
    "pev_edge2_tile_image_bytes": (c_int64, [_L]),
    "pev_edge_d2": (c_int32, [_P, _P, _P, _L, _P, _P]),
    "pev_edge2_fwd1": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _L, _L, _P, _P, _P, _P]),
    "pev_edge2_fwd2": (c_int32, [_P, _P, _P, _P, _P, _L, _P, _P, _P]),
    "pev_edge2_bwd2": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P]),
    "pev_edge2_bwd1": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P]),
    "pev_kabsch_rmsd_pairs": (c_int32, [_P, _P, _I, _I, _I, _P, _P]),
    # node-level kernels (csrc/node_kernels.cu)
    "pev_add_layernorm_fwd": (c_int32, [_P, _P, _P, _P, c_float, _L, _I, _P, _P, _P, _P, _P]),
    "pev_node_workspace_bytes": (c_int64, []),
    "pev_layernorm_bwd": (c_int32, [_P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _P]),
    "pev_column_sum": (c_int32, [_P, _L, _I, _P, _P, _P]),
    "pev_edge2_sums": (c_int32, [_P, _P, _P, _P, _P, _L, _L, _P, _P, _P, _P]),
    "pev_edge_coord_bwd_accum": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _L, _L, _P, _P]),
    "pev_edge2_wgrad_workspace_bytes": (c_int64, []),
    "pev_edge2_wgrad5": (c_int32, [_P, _P, _P, _P, _L, _P, _P, _P, _P, _P]),
    "pev_edge2_wgrad2": (c_int32, [_P, _P, _P, _P, _P, _P, _L, _P, _P, _P]),
    # node-level tcgen05 TF32 GEMMs with fused epilogues (csrc/node_gemm_kernels.cu)
    "pev_node_wgrad_workspace_bytes": (c_int64, []),
    "pev_node_wgrad": (c_int32, [_P, _I, _P, _L, c_float, _P, _P, _I, _P]),
    "pev_attn_gemm": (c_int32, [_I, _P, _L, _I, _P, _L, _I, _P, _P, _P, _I, _I, _I, _I, _L, c_float, _P, _L, _I, _P]),
    "pev_attn_scores": (c_int32, [_I, _P, _L, _I, _P, _L, _I, _P, _P, _P, _I, _I, _I, _I, _L, c_float, _P, _P, _P, _P, c_float,
                                  c_uint32, _P]),
    "pev_attn_delta": (c_int32, [_P, _P, _L, _I, _I, _P, _L, _P, _P]),
    "pev_attn_softmax": (c_int32, [_I, _P, _P, _P, _P, _P, _I, _I, _I, c_float, c_uint32, _P]),
    "pev_linear": (c_int32, [_I, _P, _L, _I, _P, _P, _L, _I, _I, c_float, c_uint32, _P, _L, _P, _L, _P]),
    "pev_transpose": (c_int32, [_P, _I, _I, c_float, _P, _P]),
    "pev_mask_grad": (c_int32, [_P, _P, _L, c_float, c_uint32, _P, _P]),
    "pev_linear_wgrad": (c_int32, [_I, _P, _L, _I, _P, _L, _I, _L, c_float, _P, _P, _L, _P]),
    "pev_split_tf32": (c_int32, [_P, _I, _I, _I, _P, _P]),
    "pev_node_gemm3": (c_int32, [_P, _I, _P, _P, _L, _I, _P, _P, _P]),
    "pev_node_wgrad3": (c_int32, [_P, _I, _P, _L, c_float, _P, _P, _I, _P]),
    "pev_node_gemm": (c_int32, [_I, _P, _I, _P, _I, _P, _P, _L, _I, c_float, _P, _P, _P, c_float, _P, _P, _P, _P, _P]),
    # fused CTA-pair edge kernels (csrc/edge_tc3_kernels.cu)
    "pev_edge3_fwd": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _P, _P, _P, _P, _P]),
}


class Lib:
    """A loaded shared library exporting (a subset of) the C ABI."""

    def __init__(self, path: str, require_tc: bool = True):
        self.path = path
        self.cdll = ctypes.CDLL(path)
        protos = dict(_PROTOS)
        if require_tc:
            protos.update(_PROTOS_TC)
        for name, (res, args) in protos.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError as e:
                raise RuntimeError(f"{path} does not export {name}; rebuild it "
                                   "(python -m protein_ensemble_vae_b200.build --force)") from e
            fn.restype, fn.argtypes = res, args
        if self.cdll.pev_abi_version() != 1:
            raise RuntimeError(f"{path}: ABI version mismatch")

    def call(self, name: str, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            msg = self.cdll.pev_last_error()
            raise RuntimeError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")

    def launch_count(self) -> int:
        return int(self.cdll.pev_launch_count())


_LIB: Lib | None = None
_REQUIRE_CUDA = True      # flipped only by tests that run the host restatement


def lib() -> Lib:
    """The CUDA library; fails loudly when it has not been built (``python -m protein_ensemble_vae_b200.build``)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the sm_100a kernels first "
                "(python -m protein_ensemble_vae_b200.build); there is no CPU fallback")
        _LIB = Lib(LIB_PATH)
    return _LIB


def ptr(t: torch.Tensor | None, dtype=None):
    """Device pointer of a contiguous tensor (``None`` -> NULL)."""
    if t is None:
        return None
    if _REQUIRE_CUDA and not t.is_cuda:
        raise RuntimeError("protein_ensemble_vae_b200 runs on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("internal error: non-contiguous tensor passed to the C ABI")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"internal error: expected {dtype}, got {t.dtype}")
    return c_void_p(t.data_ptr())


def stream(t: torch.Tensor | None = None):
    """``cudaStream_t`` of torch's current stream on the tensor's device."""
    if t is not None and t.is_cuda:
        return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return None


def f32c(t: torch.Tensor | None) -> torch.Tensor | None:
    """float32, contiguous view/copy (``None`` passes through)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# Optional per-kernel CUDA-event timing (bench.py sets PROFILE = {} around its timed region).
PROFILE: dict | None = None


class profiled:
    """``with profiled(tag): <launch>`` records a CUDA-event pair on the current stream when PROFILE is a dict."""

    def __init__(self, tag: str):
        self.tag = tag

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record()
            PROFILE.setdefault(self.tag, []).append((self.a, self.b))
        return False
