"""``HierCVAE`` -- ``models/model.py:15-116`` over the device encoder and decoder of this package.

The reference's top-level module is glue: ``ProteinEncoder`` -> ``ResidueDecoder`` (``:60-68``), plus ``encode`` (``:70-72``),
``decode`` (``:74-76``) and prior sampling (``:78-116``).  Same constructor arguments, sub-module names (``encoder`` /
``decoder``, hence the checkpoint keys ``encoder.*`` and ``decoder.decoder.decoder.*`` that ``models/training.py:456`` saves)
and return tuples, so a reference checkpoint loads with ``load_state_dict`` and the callers ``run_epoch``
(``models/training.py:89-102``) and ``generate_ensembles`` (``generate_ensemble_pdbs.py:462``, ``:554``) see the interface they
were written against.  CUDA only: both halves raise on host tensors.

Additions (keyword-only, defaults reproduce the reference): ``precision_encoder`` / ``precision_decoder`` choose the
arithmetic of the two halves (``"tf32"`` | ``"fp32"`` and ``"bf16"`` | ``"fp32"``), ``forward`` / ``encode`` accept the
reparameterisation noise (``eps_g``, ``eps_l``) so tests can pin the sampled latents, and ``sample`` accepts a
``torch.Generator``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .en_gnn_decoder import ResidueDecoder
from .encoder import ProteinEncoder


class HierCVAE(nn.Module):
    """``models/model.py:15-40``."""

    def __init__(self, seqemb_dim: Optional[int], d_model: int = 512, nhead: int = 8, ff: int = 1024, nlayers: int = 6,
                 z_g: int = 512, z_l: int = 256, dropout: float = 0.1, equivariant: bool = True, decoder_hidden: int = 256,
                 use_dihedrals: bool = True, *, precision_encoder: str = "tf32", precision_decoder: str | None = None):
        super().__init__()
        self.seqemb_dim = seqemb_dim
        self.equivariant = equivariant
        self.use_dihedrals = use_dihedrals
        self.z_g, self.z_l = z_g, z_l
        self.encoder = ProteinEncoder(seqemb_dim=seqemb_dim, d_model=d_model, nhead=nhead, ff=ff, nlayers=nlayers, z_g=z_g,
                                      z_l=z_l, dropout=dropout, use_dihedrals=use_dihedrals, precision=precision_encoder)
        self.decoder = ResidueDecoder(z_g=z_g, z_l=z_l, hidden=decoder_hidden, dropout=dropout, equivariant=equivariant,
                                      precision=precision_decoder)

    def forward(self, seqemb_or_none, n_coords, ca_coords, c_coords, dihedrals, mask, eps_g=None, eps_l=None):
        """``:42-68`` -> ``(pred_N, pred_CA, pred_C [B,L,3], pred_seq [B,L,20], mu_g, lv_g [B,zg], mu_l, lv_l [B,L,zl])``."""
        z_g, z_l, mu_g, lv_g, mu_l, lv_l = self.encoder(seqemb_or_none, n_coords, ca_coords, c_coords, dihedrals, mask,
                                                        eps_g=eps_g, eps_l=eps_l)
        pred_N, pred_CA, pred_C, pred_seq = self.decoder(z_g, z_l, mask=mask)
        return pred_N, pred_CA, pred_C, pred_seq, mu_g, lv_g, mu_l, lv_l

    def encode(self, seqemb_or_none, n_coords, ca_coords, c_coords, dihedrals, mask, eps_g=None, eps_l=None):
        """``:70-72`` -> ``(z_g, z_l, mu_g, lv_g, mu_l, lv_l)``."""
        return self.encoder(seqemb_or_none, n_coords, ca_coords, c_coords, dihedrals, mask, eps_g=eps_g, eps_l=eps_l)

    def decode(self, z_g, z_l, mask=None):
        """``:74-76`` -> ``(N, CA, C, seq_logits)``."""
        return self.decoder(z_g, z_l, mask=mask)

    def sample(self, mask, seqemb_or_none=None, num_samples: int = 1, generator: torch.Generator | None = None):
        """``:78-116``: ``num_samples`` prior draws per row of ``mask`` (sample ``s`` of protein ``b`` is output row
        ``b * num_samples + s``, the ``repeat_interleave`` order of ``:111``) -> ``(N, CA, C [B*S,L,3], seq_logits [B*S,L,20])``.
        The reference reads the latent widths off the last linear of the latent heads (``:101-102``); they are the constructor's
        ``z_g`` / ``z_l``."""
        B, L = mask.shape
        device = mask.device
        z_g = torch.randn(B * num_samples, self.z_g, device=device, generator=generator)
        z_l = torch.randn(B * num_samples, L, self.z_l, device=device, generator=generator)
        mask_expanded = mask.repeat_interleave(num_samples, dim=0)
        return self.decode(z_g, z_l, mask=mask_expanded)
