"""B200-native (sm_100a) EGNN decoder, geometric losses and Kabsch RMSD for Protein-Ensemble-VAE.

Drop-in for ``models/en_gnn_decoder.py``, ``models/losses.py``, ``generate_ensemble_pdbs.py::kabsch_rmsd`` and -- the
callers either side of that path -- ``models/encoder.py`` and ``models/model.py`` of the reference: same class / function names,
signatures, parameter names and return conventions; the work is done by hand-written CUDA
kernels reached through the C ABI of ``include/pev_b200.h``.  CUDA-only, no CPU fallback.
"""
from . import data, en_gnn_decoder, encoder, generation, graph, graphs, kabsch, losses, metrics, model, training  # noqa: F401
from .data import DevicePrefetcher  # noqa: F401
from .generation import (generate_ensemble, validate_geometry_batch, validate_protein_geometry,  # noqa: F401
                         write_ensemble_pdb)
from .graphs import GraphedStep  # noqa: F401
from .metrics import compute_gdt, compute_lddt, compute_rmsf, compute_tm_score  # noqa: F401
from .en_gnn_decoder import EGNLayer, EGNNDecoder, ResidueDecoder, SE3EquivariantDecoder  # noqa: F401
from .encoder import ProteinEncoder  # noqa: F401
from .model import HierCVAE  # noqa: F401
from .training import run_epoch  # noqa: F401
from .kabsch import ensemble_diversity, kabsch_rmsd, kabsch_rmsd_batch, kabsch_rmsd_pairs  # noqa: F401
from .losses import compute_total_loss  # noqa: F401

__all__ = ["EGNLayer", "EGNNDecoder", "SE3EquivariantDecoder", "ResidueDecoder", "ProteinEncoder", "HierCVAE", "run_epoch",
           "compute_total_loss",
           "kabsch_rmsd", "kabsch_rmsd_batch", "kabsch_rmsd_pairs", "ensemble_diversity", "DevicePrefetcher",
           "GraphedStep", "generate_ensemble", "validate_geometry_batch", "validate_protein_geometry", "write_ensemble_pdb", "compute_tm_score",
           "compute_lddt", "compute_gdt", "compute_rmsf", "metrics", "data", "losses",
           "en_gnn_decoder", "encoder", "model", "training", "generation", "graph", "graphs", "kabsch"]
