#!/usr/bin/env python3
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference sources of the hot path, copied from where they lie under
``/root/reference`` so that they travel to the GPU box (``oracle/_ref/`` is git-ignored, not gpurun-ignored).

    python oracle/make_ref.py            # no-op when /root/reference is absent (e.g. on the GPU box)

Copied byte for byte: ``models/en_gnn_decoder.py`` and ``models/losses.py`` (both import only torch).  ``kabsch_rmsd`` lives
in ``generate_ensemble_pdbs.py``, whose module imports h5py (absent here), so that one function's source segment is
written to ``_ref/kabsch_ref.py`` verbatim behind the two imports it needs.  TEST / BENCH INFRASTRUCTURE ONLY:
``bench.py --impl reference`` and the ``cpu_baseline`` leg time these files on the host cores (kind "reference");
nothing in the product imports them.
"""
from __future__ import annotations

import ast
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PEV_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "models")):
        print(f"{REF} not present: oracle/_ref left as it is")
        return 0
    os.makedirs(OUT, exist_ok=True)
    for name in ("en_gnn_decoder.py", "losses.py"):
        shutil.copyfile(os.path.join(REF, "models", name), os.path.join(OUT, name))
    path = os.path.join(REF, "generate_ensemble_pdbs.py")
    src = open(path).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "kabsch_rmsd")
    with open(os.path.join(OUT, "kabsch_ref.py"), "w") as f:
        f.write("# source segment of generate_ensemble_pdbs.py:kabsch_rmsd, verbatim (oracle/make_ref.py)\n"
                "import numpy as np\nimport torch\n\n\n" + ast.get_source_segment(src, fn) + "\n")
    print("oracle/_ref:", sorted(os.listdir(OUT)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
