"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/smooth.npz by calling the REAL reference's
``seistorch.signal.gaussian_filter`` (CPU) in the order of ``PostProcess.smooth_gradient`` (process.py:66-112):
a random 61 x 83 "gradient", two smoothing configurations (counts, sigma_z, sigma_x, radius_z, radius_x).

    python -m oracle.make_smooth_golden
"""
import os

import numpy as np
import torch

from . import ref_shim


def main():
    ref_shim.import_reference()
    from seistorch.signal import gaussian_filter as ref_gf
    rng = np.random.default_rng(3)
    g = (rng.standard_normal((61, 83)) * np.linspace(1, 5, 83)[None, :]).astype(np.float32)
    out = {"g": g}
    cfgs = [(1, {"z": 2.0, "x": 3.5}, {"z": 4, "x": 6}), (3, {"z": 1.0, "x": 1.0}, {"z": 2, "x": 2})]
    for k, (counts, sigma, radius) in enumerate(cfgs):
        t = torch.from_numpy(g)
        for _ in range(counts):
            t = ref_gf(t, sigma["z"], radius["z"], axis=0)
            t = ref_gf(t, sigma["x"], radius["x"], axis=1)
        out[f"y_{k}"] = t.numpy()
        out[f"cfg_{k}"] = np.array([counts, sigma["z"], sigma["x"], radius["z"], radius["x"]], dtype=np.float64)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "smooth.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
