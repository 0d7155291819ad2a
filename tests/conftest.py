import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load_golden(name):
    from oracle.make_golden import unpack_case
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, unpack_case(z)


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def cat_records(recs):
    return np.concatenate([np.asarray(r, dtype=np.float64).ravel() for r in recs])


def golden_records(z, tag):
    n = len([k for k in z.files if k.startswith(tag + "_rec_")])
    return [z[f"{tag}_rec_{i}"] for i in range(n)]
