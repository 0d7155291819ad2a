"""CPU: the per-cell kernel arithmetic (csrc/st_wave2d_math.cuh, st_elastic2d.cuh), run cell
by cell on the host by tests/hostcheck, against the reference's golden vectors: forward
seismograms and the exact-adjoint gradients of every 2D equation on the path."""
import numpy as np
import pytest

import hostcheck_driver as hd
from conftest import cat_records, golden_records, load_golden, rel

CASES = ["acoustic", "acoustic_habc", "elastic", "vti_habc2", "tti_habc", "acoustic_lsrtm_habc", "acoustic_rho_habc", "acoustic_vti_lsrtm_habc",
         "acoustic_tti_lsrtm_habc", "acoustic_fwim_habc", "acoustic_multiple", "acoustic_habc_multiple",
         "acoustic_habc_ragged"]


@pytest.mark.parametrize("name", CASES)
def test_kernel_math_matches_reference(name):
    z, case = load_golden(name)
    recs, grads = hd.run_case(case)
    ref64 = cat_records(golden_records(z, "f64"))
    ref32 = cat_records(golden_records(z, "f32"))
    e_new, e_ref = rel(cat_records(recs), ref64), rel(ref32, ref64)
    assert e_new <= 1e-5, (e_new, e_ref)          # north-star seismogram tolerance
    for k, g in grads.items():
        if g is None:
            continue
        assert rel(g, z[f"f64_grad_{k}"]) <= 1e-4, k      # north-star gradient tolerance


def test_side_weights_tile_the_frame():
    """Every frame cell is owned (weights sum to 1), interior cells have none; agrees with
    the oracle's closed form, which is pinned bit-exactly to the reference."""
    import ctypes as C
    from oracle.equations import habc_side_weights
    for nz, nx, mult in [(130, 144, False), (101, 103, False), (80, 144, True), (151, 101, True)]:
        f = np.zeros((4, nz, nx), np.float32)
        fr = np.zeros((nz, nx), np.uint8)
        hd.lib().hc_side_weights(nz, nx, 50, int(mult), f.ctypes.data_as(C.c_void_p), fr.ctypes.data_as(C.c_void_p))
        assert np.array_equal(f, habc_side_weights(nz, nx, 50, mult).astype(np.float32))
        assert np.array_equal(f.sum(0) == 1.0, fr.astype(bool))
        assert np.all(f.sum(0)[fr == 0] == 0)


def test_wrap_cells_are_the_four_closed_form_candidates():
    """The one-way blend reads a wrapped neighbour only at depth bw-1; with the reference's HABC
    weights that carries a non-zero weight for at most the four cells st_wrap_candidate()
    enumerates (csrc/st_wave2d_band.cuh) -- checked here against the oracle's masks/weights."""
    import torch
    from oracle import boundary
    from oracle.equations import habc_side_weights
    for nz, nx, mult in [(130, 144, False), (851, 2401, False), (101, 103, False), (80, 144, True)]:
        f = habc_side_weights(nz, nx, 50, mult)
        b = boundary.habc_coefficients_2d((nz, nx), 50, mult, torch.float64).numpy()
        zz, xx = np.meshgrid(np.arange(nz), np.arange(nx), indexing="ij")
        depth = [zz, nz - 1 - zz, xx, nx - 1 - xx]
        found = set()
        for s in range(4):
            sel = (depth[s] == 49) & (f[s] != 0) & (b != 0)
            found |= {(int(z), int(x), s) for z, x in zip(zz[sel], xx[sel])}
        cand = {(49, nx - 1, 0), (0, nx - 50, 3), (nz - 50, 0, 1), (nz - 1, 49, 2)}
        assert found <= cand, (nz, nx, mult, found - cand)
        if not mult:
            assert found == cand
