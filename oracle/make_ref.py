"""TEST INFRASTRUCTURE ONLY -- recipe that ships the UNMODIFIED reference to the GPU box.

    python -m oracle.make_ref            # /root/reference -> oracle/_ref/

The reference (GeophyAI/seistorch) is pure Python; ``/root/reference`` exists only in the
authoring container.  This script copies its ``seistorch`` package and the driver scripts
byte for byte into the git-ignored directory ``oracle/_ref/`` (never committed, but shipped
with the gpurun snapshot like a built ``.so``), so that on the GPU box

  * ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's own code
    (``kind: "reference"``), and
  * the ``-m gpu`` overlay tests run the reference's ``build_model`` / drivers' call
    sequence on top of the sm_100a kernels.

``__graft_entry__.build()`` runs it whenever ``/root/reference`` is present.  A manifest
with the sha256 of every copied file is written next to the copy so a test can verify that
nothing in ``oracle/_ref`` was edited.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SEISTORCH_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
DRIVERS = ["fwi.py", "fwi_torchrun.py", "seistorch_dist.py", "seistorch_dist_lbfgs.py", "codingfwi.py",
           "template.yml", "requirements.txt"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def make_ref(verbose=True) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "seistorch")):
        return DST if os.path.isdir(os.path.join(DST, "seistorch")) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(SRC, "seistorch"), os.path.join(DST, "seistorch"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name in DRIVERS:
        if os.path.exists(os.path.join(SRC, name)):
            shutil.copy2(os.path.join(SRC, name), os.path.join(DST, name))
    manifest = {}
    for root, _dirs, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, DST)] = _sha(p)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    if verbose:
        print(f"[oracle.make_ref] copied {len(manifest)} files {SRC} -> {DST}")
    return DST


def verify() -> bool:
    """True if every file of oracle/_ref still has the sha256 recorded at copy time."""
    mf = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    manifest = json.load(open(mf))
    return all(os.path.exists(os.path.join(DST, k)) and _sha(os.path.join(DST, k)) == v for k, v in manifest.items())


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
