"""TEST INFRASTRUCTURE ONLY -- runs the REFERENCE's own code (oracle/_ref, or /root/reference in the authoring
container) on top of the sm_100a kernels through ``seistorch_b200.overlay`` and prints one JSON line.  Started as a
subprocess by tests/test_gpu_overlay.py so that the overlaid ``seistorch`` package never mixes with the plain
reference imported elsewhere in the test session.

    python tests/ref_on_kernels.py golden <name> [device]     build_model -> reset_geom -> model(x) -> Loss -> backward
                                                               (the call sequence of seistorch_dist.py:92-258)
    python tests/ref_on_kernels.py forward <name> [device]    the worker body of fwi.py:146-162 (forward modelling)
    torchrun --nproc-per-node N tests/ref_on_kernels.py dist <workdir>
                                                               runs the UNMODIFIED seistorch_dist.py (runpy, __main__)
    python tests/ref_on_kernels.py prepare-dist <workdir> <name>   writes config / models / geometry / observed data for it
"""
import json
import os
import pickle
import runpy
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def setup_overlay(use_overlay=True):
    """stand-ins for the absent third-party modules -> overlay -> reference package."""
    from oracle import ref_shim
    import standins.h5py as fake_h5py
    sys.modules["h5py"] = fake_h5py                     # before the generic attribute-sink stubs
    ref_shim._install_stubs()
    ref_shim._patch_tensor_to()
    if ref_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    names = []
    if use_overlay:
        import seistorch_b200.overlay as ov
        names = ov.install()
    import seistorch.compile as sc
    sc.force_compile = False
    _torch_compat()
    return names


def _torch_compat():
    """The reference targets torch ~2.0; torch 2.11 dropped ``verbose=`` from the LR schedulers
    (seistorch/setup.py:181 passes it).  Accept and ignore the keyword -- an environment shim like the
    h5py stand-in, not a change of the reference."""
    import torch
    sched = torch.optim.lr_scheduler
    if getattr(sched.ExponentialLR, "_accepts_verbose", False):
        return
    base = sched.ExponentialLR

    class ExponentialLR(base):
        _accepts_verbose = True

        def __init__(self, *a, verbose=None, **k):
            super().__init__(*a, **k)

    sched.ExponentialLR = ExponentialLR


def _check_ours(model):
    import seistorch_b200.cell
    import seistorch_b200.rnn
    m = getattr(model, "module", model)
    assert type(m) is seistorch_b200.rnn.WaveRNN, type(m)
    assert type(m.cell) is seistorch_b200.cell.WaveCell, type(m.cell)
    assert m.cell.forward_func.__module__.startswith("seistorch_b200.equations"), m.cell.forward_func.__module__


def run_golden(name, device="cuda", forward_only=False):
    import torch
    from conftest import cat_records, golden_records, load_golden, rel
    from oracle import ref_runner
    names = setup_overlay()
    z, case = load_golden(name)
    if device == "cuda:last":            # a device that is NOT the current one (ADVICE: no set_device in the drivers)
        device = f"cuda:{torch.cuda.device_count() - 1}"
    cfg, model, x = ref_runner.build_reference(case, "float32", want_grad=not forward_only, device=device)
    model.to(device)                                     # seistorch_dist.py:95 `model.to(rank)`, fwi.py:112 `model.to(args.dev)`
    _check_ours(model)
    from seistorch.loss import Loss                      # the reference's wrapper; classes patched by the overlay
    out = {"name": name, "device": device, "overlaid": len(names), "current_device": torch.cuda.current_device()}
    x = x.to(device)
    if forward_only:
        # fwi.py:146-162 worker body
        model.train()
        with torch.no_grad():
            shots = list(range(len(case["sources"])))
            model.reset_geom(shots, case["sources"], case["receivers"], cfg)
            y = model(x)
            record = y.numpy()
        out["rec_err"] = rel(cat_records(list(record)), cat_records(golden_records(z, "f64")))
        print(json.dumps(out))
        return
    model.train()
    syn = model(x)
    loss_name = bytes(z["loss_name"]).decode() if "loss_name" in z.files else "l2"
    crit = Loss(loss_name).loss(cfg)
    out["loss_class"] = type(crit).__module__ + "." + type(crit).__name__
    if case.get("obs") is not None:
        obs = [torch.as_tensor(o, device=device) for o in case["obs"]]
    else:
        obs = [torch.zeros_like(s) for s in syn]
    if len({tuple(t.shape) for t in syn}) == 1:
        loss = crit(torch.stack(list(syn), 0), torch.stack(obs, 0))
    else:
        loss = crit(list(syn), obs)
    loss.backward()
    out["rec_err"] = rel(cat_records([s.detach().cpu().numpy() for s in syn]), cat_records(golden_records(z, "f64")))
    out["loss_err"] = abs(float(loss) - float(z["f64_loss"])) / abs(float(z["f64_loss"]))
    out["grad_err"] = {}
    for k in model.cell.geom.pars_need_invert:
        g = getattr(model.cell.geom, k).grad
        key = f"f64_grad_{k}"
        if g is not None and key in z.files:
            out["grad_err"][k] = rel(g.cpu().numpy(), z[key])
    print(json.dumps(out))


# ------------------------------------------------------------------ the unmodified torchrun driver
PARAM_KEYS = ["vp", "vs", "rho", "Q", "epsilon", "delta", "theta", "m", "rx", "rz"]


def dist_case():
    """4 shots on a 40 x 64 model, acoustic_habc; the observed data come from a perturbed model."""
    from oracle import cases
    case = cases.make_case("acoustic_habc", nz=40, nx=64, nshots=4, nt=160, rec_step=3)
    true = dict(case, models={"vp": (case["models"]["vp"] * 1.04).astype(np.float32)})
    return case, true


def prepare_dist(work):
    """Files the driver reads: YAML config, .npy models, pickled geometry, hdf5 observed data (stand-in)."""
    import torch
    import yaml
    from oracle import loop
    import standins.h5py as h5
    os.makedirs(work, exist_ok=True)
    case, true = dist_case()
    obs, _ = loop.simulate(true, dtype=torch.float32)
    with h5.File(os.path.join(work, "obs.hdf5"), "w") as f:
        for i, o in enumerate(obs):
            f.create_dataset(f"shot_{i}", data=o.numpy())
    paths = {k: None for k in PARAM_KEYS}
    np.save(os.path.join(work, "vp.npy"), case["models"]["vp"])
    paths["vp"] = os.path.join(work, "vp.npy")
    pickle.dump(case["sources"], open(os.path.join(work, "sources.pkl"), "wb"))
    pickle.dump(case["receivers"], open(os.path.join(work, "receivers.pkl"), "wb"))
    cfg = {
        "seed": 20230503, "name": "dist", "dtype": "float32", "equation": "acoustic_habc",
        "training": {"implicit": {"use": False, "pretrained": None}, "minibatch": True, "batch_size": 4,
                     "N_epochs": 1, "lr": {"vp": 10.0}, "scale_decay": 1.0, "lr_decay": 1.0, "filter_ord": 3,
                     "optimizer": "adam"},
        "geom": {"obsPath": os.path.join(work, "obs.hdf5"), "truePath": dict(paths), "initPath": dict(paths),
                 "sources": os.path.join(work, "sources.pkl"), "receivers": os.path.join(work, "receivers.pkl"),
                 "wavelet": None, "multiple": False, "boundary_saving": True, "wavelet_delay": 60,
                 "wavelet_inverse": False, "source_type": ["h1"], "receiver_type": ["h1"],
                 "invlist": {k: k == "vp" for k in PARAM_KEYS}, "inv_savePath": os.path.join(work, "results"),
                 "multiscale": [[8.0]], "dt": float(case["dt"]), "nt": int(case["nt"]), "fm": 10.0,
                 "h": float(case["h"]), "Nshots": 4, "boundary": {"type": "habc", "width": 50}},
    }
    with open(os.path.join(work, "config.yml"), "w") as f:
        yaml.safe_dump(cfg, f)
    return cfg


def run_dist(work):
    """runpy the reference's seistorch_dist.py byte for byte, as __main__, under torchrun."""
    from oracle import ref_shim
    setup_overlay()
    script = os.path.join(ref_shim.REFERENCE_ROOT, "seistorch_dist.py")
    sys.argv = [script, os.path.join(work, "config.yml"), "--opt", "adam", "--loss", "vp=l2", "--lr", "vp=10.0",
                "--mode", "inversion", "--save-path", os.path.join(work, "results"), "--use-cuda"]
    runpy.run_path(script, run_name="__main__")


def expected_dist_gradient(world):
    """What the driver must have saved in grad_vp_nosm_0.pt: per rank the sum over its shots of the gradient of
    L2(filter(syn), filter(obs)), then DDP's mean over ranks -- from the float64 oracle."""
    import torch
    from oracle import loop, sigproc
    case, true = dist_case()
    obs, _ = loop.simulate(true, dtype=torch.float32)
    b, a = sigproc.butter(3, [8.0], float(case["dt"]))
    syn, params = loop.simulate(case, dtype=torch.float64, requires_grad=["vp"])

    class F(torch.autograd.Function):                   # the filter is linear and self-adjoint
        @staticmethod
        def forward(ctx, x):
            y = sigproc.filtfilt(x.detach().numpy().astype(np.float64), b, a)
            return torch.from_numpy(y.astype(np.float64))

        @staticmethod
        def backward(ctx, g):
            return torch.from_numpy(sigproc.filtfilt(g.numpy().astype(np.float64), b, a).astype(np.float64))

    loss = 0.0
    for s, o in zip(syn, obs):
        fo = torch.from_numpy(sigproc.filtfilt(o.numpy(), b, a).astype(np.float64))
        loss = loss + ((F.apply(s) - fo) ** 2).sum()
    loss.backward()
    return params["vp"].grad.numpy() / world, float(loss)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "golden":
        run_golden(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "cuda")
    elif mode == "forward":
        run_golden(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "cuda", forward_only=True)
    elif mode == "prepare-dist":
        prepare_dist(sys.argv[2])
    elif mode == "dist":
        run_dist(sys.argv[2])
    else:
        raise SystemExit(f"unknown mode {mode}")
