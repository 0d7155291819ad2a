"""CPU, only where the reference tree is mounted (/root/reference): the overlay makes the
reference's OWN build_model construct our WaveCell / WaveRNN / equation modules."""
import os
import pickle

import numpy as np
import pytest
import yaml

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present (GPU box)")


def test_reference_build_model_uses_overlay(tmp_path):
    import sys
    ref_shim._install_stubs()
    ref_shim._patch_tensor_to()
    import seistorch_b200.overlay as ov
    saved = {k: v for k, v in sys.modules.items() if k == "seistorch" or k.startswith("seistorch.")}
    for k in list(saved):
        del sys.modules[k]
    try:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
        names = ov.install()
        assert "seistorch.equations2d.acoustic_habc" in names
        # class-level patches inside the reference's own modules: fused misfits, device filter, device smoothing
        assert {"seistorch.loss.L2", "seistorch.loss.Envelope", "seistorch.signal.SeisSignal.filter",
                "seistorch.process.PostProcess.smooth_gradient"} <= set(names)
        import seistorch.loss as rl
        import seistorch_b200.loss as ol
        crit = rl.Loss("l2").loss({})
        assert type(crit) is ol.L2 and type(rl.Loss("envelope").loss({})) is ol.Envelope
        assert type(rl.Loss("phase").loss({})).__module__ == "seistorch.loss"      # not accelerated: stays the reference's
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
        from seistorch.model import build_model          # the reference's builder
        import seistorch.compile as sc
        sc.force_compile = False
        vp = np.full((20, 30), 2000.0, np.float32)
        np.save(tmp_path / "vp.npy", vp)
        pickle.dump([[5, 1]], open(tmp_path / "s.pkl", "wb"))
        pickle.dump([[[1, 2, 3], [1, 1, 1]]], open(tmp_path / "r.pkl", "wb"))
        paths = {k: None for k in ["vp", "vs", "rho", "Q", "epsilon", "delta", "theta", "m", "rx", "rz"]}
        paths["vp"] = str(tmp_path / "vp.npy")
        cfg = {"seed": 1, "name": "t", "dtype": "float32", "equation": "acoustic_habc",
               "training": {"implicit": {"use": False, "pretrained": None}, "minibatch": True, "batch_size": 1,
                            "N_epochs": 1, "lr": None, "scale_decay": 1.0, "lr_decay": 1.0, "filter_ord": 3},
               "geom": {"obsPath": None, "truePath": dict(paths), "initPath": dict(paths),
                        "sources": str(tmp_path / "s.pkl"), "receivers": str(tmp_path / "r.pkl"), "wavelet": None,
                        "multiple": False, "boundary_saving": True, "wavelet_delay": 0, "wavelet_inverse": False,
                        "source_type": ["h1"], "receiver_type": ["h1"],
                        "invlist": {k: k == "vp" for k in paths}, "inv_savePath": None, "multiscale": ["all"],
                        "dt": 1e-3, "nt": 10, "fm": 10.0, "h": 10.0, "Nshots": 1,
                        "boundary": {"type": "habc", "width": 50}}}
        yaml.safe_dump(cfg, open(tmp_path / "c.yml", "w"))
        cfg2, model = build_model(str(tmp_path / "c.yml"), device="cpu", mode="inversion")
        import seistorch_b200.cell
        import seistorch_b200.rnn
        assert type(model) is seistorch_b200.rnn.WaveRNN
        assert type(model.cell) is seistorch_b200.cell.WaveCell
        assert model.cell.forward_func.__module__ == "seistorch_b200.equations2d.acoustic_habc"
        assert model.cell.geom.vp.requires_grad and tuple(model.cell.geom.domain_shape) == (120, 130)
        model.reset_geom([0], [[5, 1]], [[[1, 2, 3], [1, 1, 1]]], cfg2)
        assert int(model.sources[0].x) == 55
    finally:
        for k in [k for k in sys.modules if k == "seistorch" or k.startswith("seistorch.")]:
            del sys.modules[k]
        sys.modules.update(saved)
