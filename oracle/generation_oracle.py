"""Oracle: geometry validity filter of the generation driver.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Restatement of ``validate_protein_geometry``
(``generate_ensemble_pdbs.py:290-340``; its module imports h5py and cannot be imported here, so
``tests/test_host_generation.py`` also pins this restatement on the function's own source, lifted by ``ast`` where
/root/reference exists).  Parity unpinned by the reference's tests (it has none)."""
from __future__ import annotations

import numpy as np


def validate_protein_geometry(coords_ca: np.ndarray, mask: np.ndarray):
    """``(status, max_dist, avg_dist, avg_angle)``; status 0 valid, 1 no residues (:302-303), 2 extreme distance
    (:314-315), 3 abnormal average distance (:317-318), 4 abnormal average angle (:334-336)."""
    v = np.asarray(coords_ca, dtype=np.float64)[np.asarray(mask) != 0]      # :301-305, gaps are bridged
    if len(v) == 0:
        return 1, 0.0, 0.0, 0.0
    if len(v) < 2:
        return 0, 0.0, 0.0, 0.0
    d = np.linalg.norm(v[1:] - v[:-1], axis=1)                              # :309-311
    mx, avg = float(d.max()), float(d.mean())
    ang = 0.0
    if len(v) > 2:
        v1, v2 = v[:-2] - v[1:-1], v[2:] - v[1:-1]                          # :324-325
        c = (v1 * v2).sum(1) / (np.linalg.norm(v1, axis=1) * np.linalg.norm(v2, axis=1) + 1e-8)
        ang = float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))).mean())    # :328-333
    if mx > 6.0:
        return 2, mx, avg, ang
    if avg < 2.5 or avg > 5.0:
        return 3, mx, avg, ang
    if len(v) > 2 and (ang < 60 or ang > 180):
        return 4, mx, avg, ang
    return 0, mx, avg, ang
