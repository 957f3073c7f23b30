"""Deterministic synthetic inputs shared by ``make_golden.py`` and the tests.

Everything is drawn from ``numpy.random.default_rng`` (PCG64, stream-stable) and
rounded to float32 so the same bits feed the float64 reference, the oracle and
the CUDA kernels.
"""
from __future__ import annotations

import numpy as np


def f32(a):
    return np.asarray(a, dtype=np.float32)


def decoder_param_shapes(z_g, z_l, hidden, num_layers):
    """Parameter names/shapes of the reference ``EGNNDecoder``."""
    H = hidden
    s = {
        "input_embedding.weight": (H, z_g + z_l), "input_embedding.bias": (H,),
        "latent_to_coords.0.weight": (H, z_g + z_l), "latent_to_coords.0.bias": (H,),
        "latent_to_coords.1.weight": (H,), "latent_to_coords.1.bias": (H,),
        "latent_to_coords.4.weight": (H // 2, H), "latent_to_coords.4.bias": (H // 2,),
        "latent_to_coords.6.weight": (3, H // 2), "latent_to_coords.6.bias": (3,),
        "n_offset_head.0.weight": (H // 2, H), "n_offset_head.0.bias": (H // 2,),
        "n_offset_head.2.weight": (4, H // 2), "n_offset_head.2.bias": (4,),
        "c_offset_head.0.weight": (H // 2, H), "c_offset_head.0.bias": (H // 2,),
        "c_offset_head.2.weight": (4, H // 2), "c_offset_head.2.bias": (4,),
        "sequence_head.0.weight": (2 * H, H), "sequence_head.0.bias": (2 * H,),
        "sequence_head.1.weight": (2 * H,), "sequence_head.1.bias": (2 * H,),
        "sequence_head.4.weight": (H, 2 * H), "sequence_head.4.bias": (H,),
        "sequence_head.5.weight": (H,), "sequence_head.5.bias": (H,),
        "sequence_head.8.weight": (20, H), "sequence_head.8.bias": (20,),
    }
    for l in range(num_layers):
        s.update(layer_param_shapes(H, H, prefix=f"layers.{l}."))
    return s


def layer_param_shapes(node_dim, hidden, prefix=""):
    D, H, p = node_dim, hidden, prefix
    return {
        p + "phi_e.0.weight": (H, 2 * D + 1), p + "phi_e.0.bias": (H,),
        p + "phi_e.2.weight": (H, H), p + "phi_e.2.bias": (H,),
        p + "phi_h.0.weight": (H, D + H), p + "phi_h.0.bias": (H,),
        p + "phi_h.2.weight": (D, H), p + "phi_h.2.bias": (D,),
        p + "phi_x.0.weight": (H, H), p + "phi_x.0.bias": (H,),
        p + "phi_x.2.weight": (1, H), p + "phi_x.2.bias": (1,),
        p + "norm_h.weight": (D,), p + "norm_h.bias": (D,),
    }


_NORM_WEIGHTS = ("norm_h.weight", "latent_to_coords.1.weight", "sequence_head.1.weight",
                 "sequence_head.5.weight")


def make_params(shapes, seed):
    """name -> float32 ndarray.  Matrices ~ N(0, 1/fan_in), vectors ~ 0.1 N(0,1), LN gains ~ 1 + 0.1 N."""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        g = rng.standard_normal(shp)
        if len(shp) == 2:
            g = g / np.sqrt(shp[1])
        elif name.endswith(_NORM_WEIGHTS):
            g = 1.0 + 0.1 * g
        else:
            g = 0.1 * g
        out[name] = f32(g)
    return out


def make_masks(B, L, seed, kind="ragged"):
    """float32 [B,L] masks: 'full', 'ragged' (padded tails), 'gaps' (interior holes too), 'empty_row'."""
    rng = np.random.default_rng(seed)
    m = np.ones((B, L), np.float32)
    if kind == "full":
        return m
    for b in range(B):
        keep = int(rng.integers(max(2, L // 2), L + 1)) if b > 0 else L
        m[b, keep:] = 0
    if kind in ("gaps", "empty_row"):
        for b in range(1, B):
            start = int(rng.integers(1, max(2, L // 2)))
            m[b, start:start + 2] = 0
    if kind == "empty_row" and B > 1:
        m[B - 1] = 0
    return m


def make_backbone(B, L, seed, noise=0.8):
    """Random-walk CA trace (step N(0,2.2^2)) centred per conformer; N,C = CA + N(0,noise^2)."""
    rng = np.random.default_rng(seed)
    ca = np.cumsum(rng.standard_normal((B, L, 3)) * 2.2, axis=1)
    ca = ca - ca.mean(axis=1, keepdims=True)
    n = ca + noise * rng.standard_normal((B, L, 3))
    c = ca + noise * rng.standard_normal((B, L, 3))
    return f32(n), f32(ca), f32(c)


def nerf_backbone(L, phi=-60.0, psi=-45.0, omega=180.0):
    """Ideal backbone by NeRF placement with ideal bonds/angles; float64 [L,3] x3 (N, CA, C)."""
    bond = {"N_CA": 1.46, "CA_C": 1.52, "C_N": 1.33}
    ang = {"N_CA_C": 110.0, "CA_C_N": 116.0, "C_N_CA": 121.0}
    rad = np.pi / 180.0

    def place(a, b, c, length, theta, tors):
        bc = c - b
        bc /= np.linalg.norm(bc)
        n = np.cross(b - a, bc)
        n /= np.linalg.norm(n)
        m = np.cross(n, bc)
        d2 = np.array([-length * np.cos(theta), length * np.sin(theta) * np.cos(tors),
                       length * np.sin(theta) * np.sin(tors)])
        return c + d2[0] * bc + d2[1] * m + d2[2] * n

    N = [np.array([0.0, 0.0, 0.0])]
    CA = [np.array([bond["N_CA"], 0.0, 0.0])]
    th = ang["N_CA_C"] * rad
    C = [CA[0] + bond["CA_C"] * np.array([-np.cos(th), np.sin(th), 0.0])]
    for i in range(1, L):
        n_i = place(N[i - 1], CA[i - 1], C[i - 1], bond["C_N"], ang["CA_C_N"] * rad, psi * rad)
        ca_i = place(CA[i - 1], C[i - 1], n_i, bond["N_CA"], ang["C_N_CA"] * rad, omega * rad)
        c_i = place(C[i - 1], n_i, ca_i, bond["CA_C"], ang["N_CA_C"] * rad, phi * rad)
        N.append(n_i), CA.append(ca_i), C.append(c_i)
    return np.stack(N), np.stack(CA), np.stack(C)
