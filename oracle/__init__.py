"""CPU oracle for the EGNN-decoder / geometric-loss / Kabsch hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``protein_ensemble_vae_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU arm, never as the product.

The reference (mohit03031999/Protein-Ensemble-VAE) is pure Python on top of
PyTorch and numpy, so this restatement uses the same two libraries on the CPU:
``torch`` (so that float64 autograd supplies the gradient oracle) and ``numpy``
(integer graph construction, Kabsch SVD).  Every function cites the reference
``file:line`` it follows (paths relative to ``/root/reference``).

Parity pinning: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so the pins are outputs of the reference itself, run in
this container in float64 and committed under ``tests/golden/`` together with
the generating script ``tests/golden/make_golden.py``.  ``tests/test_oracle_*``
check this restatement against those fixtures.
"""
from . import graph_oracle, egnn_oracle, losses_oracle, kabsch_oracle  # noqa: F401
