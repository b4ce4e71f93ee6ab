"""GPU: flux normalisation and modal overlaps with the finite-grid correction factors of ModeSolver._grid_correction
(mode_solver.py:847-904; b200ms_problem.grid_correction, csrc/post.cuh GC kernel variants) against the numpy restatement
(oracle/postprocess.py) applied to the raw device fields.  (The file name sorts last on purpose: the newest device code of
the round is exercised after everything else.)"""
import numpy as np
import pytest

from oracle import postprocess as OP
from tidy3d_b200 import compute_modes_batch
from tidy3d_b200 import postprocess as PP
from tidy3d_b200 import workloads as W
from tidy3d_b200.solver import get_handle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["lossless", "bend_single_minus"])
def test_flux_normalisation_and_overlaps_with_grid_correction(case):
    """A mode plane a third of the way between two grid boundaries of a coarse normal grid: primal and dual factors differ
    from 1 by a few per cent, per mode.  The bent case has complex n_eff (factors of modulus != 1, different for every mode),
    complex64 fields and backward propagation (k -> -k)."""
    nf = 3
    if case == "lossless":
        wl = W.c2(nf=nf, n=96)
        kw = {}
    else:
        wl = W.c4(n=96)
        wl.mode_spec.precision = "single"
        wl.freqs = [wl.freqs[0] * s for s in (0.99, 1.0, 1.01)]
        kw = dict(direction="-")
    bounds = np.array([-0.12, -0.04, 0.05, 0.13])
    centers = (bounds[:-1] + bounds[1:]) / 2
    pos = bounds[1] + (bounds[2] - bounds[1]) / 3
    table = PP.grid_correction_table(bounds, centers, pos)
    probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec, **kw) for f in wl.freqs]
    h = get_handle(tolerance="tight")
    raw = compute_modes_batch(probs, handle=h)
    out, info = compute_modes_batch([dict(p, grid_correction=table) for p in probs], handle=h, post=("gauge", "normalize", "flux", "overlaps"),
                                    return_info=True)
    plain, info_plain = compute_modes_batch(probs, handle=h, post=("normalize", "flux"), return_info=True)
    tol = 2e-5 if case != "lossless" else 1e-9
    prev = None
    for i, ((f_raw, n_raw, _), (f_post, n_post, _)) in enumerate(zip(raw, out)):
        assert np.array_equal(n_raw, n_post)
        corr = OP.grid_correction(n_raw, wl.freqs[i], bounds, centers, pos, 0.0, kw.get("direction", "+"))
        assert np.abs(np.abs(corr[1]) - 1).max() > 1e-3  # the correction is not a no-op
        g, _ = OP.gauge(f_raw.astype(complex))
        fn, fl = OP.normalize(g, wl.coords, correction=corr)
        assert np.abs(info[i]["flux"] - fl).max() < max(tol, 1e-10) * np.abs(fl).max()
        assert np.abs(info[i]["flux"] - info_plain[i]["flux"]).max() > 1e-4 * np.abs(fl).max()  # and it reaches the flux
        assert np.abs(f_post - fn).max() < max(tol, 1e-9) * np.abs(fn).max()
        assert np.allclose(np.abs(OP.flux(f_post.astype(complex), wl.coords, correction=corr)), 1.0, atol=max(tol, 1e-9))
        if prev is not None:
            ref = OP.dot(prev[0], fn, wl.coords, correction_a=prev[1], correction_b=corr)
            assert np.abs(info[i]["overlap_prev"] - ref).max() < max(10 * tol, 1e-9)
        prev = (fn, corr)


@pytest.mark.parametrize("name", ["angled_48_minus", "angled_phi_48"])
def test_te_fraction_of_an_angled_plane_is_taken_in_the_propagation_axes(name):
    """ModeData.pol_fraction rotates the colocated field by -phi and -theta first (monitor_data.py:1603-1607, 1625-1652)."""
    from tests.golden.cases import CASES

    fac, kw, _ = CASES[name]
    wl = fac()
    theta, phi = wl.mode_spec.angle_theta, getattr(wl.mode_spec, "angle_phi", 0.0)
    assert theta != 0.0
    probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)]
    raw = compute_modes_batch(probs)[0][0]
    _, info = compute_modes_batch(probs, post=("flux",), return_info=True)
    want = OP.pol_fraction(raw.astype(complex), wl.coords, angle_theta=theta, angle_phi=phi)
    tol = 1e-5 if raw.dtype == np.complex64 else 1e-10
    assert np.abs(info[0]["te_fraction"] - want).max() < tol
    assert np.abs(want - OP.pol_fraction(raw.astype(complex), wl.coords)).max() > 1e-4  # the rotation matters


def test_flux_and_overlaps_of_a_finite_plane():
    """b200ms_problem.plane_bounds (ABI v203): the plane ends inside the second / second-to-last cell on every side; flux,
    normalisation, TE fraction and overlaps on the device use the truncated integration cells of the reference's _diff_area
    (monitor_data.py:437-455; the restatement is pinned to it by tests/golden/post_finite_plane.npz)."""
    wl = W.c1()
    x, y = (np.asarray(c, float) for c in wl.coords)
    pb = (x[1] + 0.37 * (x[2] - x[1]), x[-2] - 0.61 * (x[-2] - x[-3]), y[1] + 0.2 * (y[2] - y[1]), y[-2] - 0.45 * (y[-2] - y[-3]))
    freqs = [wl.freqs[0] * s for s in (1.0, 1.01)]
    probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in freqs]
    h = get_handle(tolerance="tight")
    raw = compute_modes_batch(probs, handle=h)
    out, info = compute_modes_batch([dict(p, plane_bounds=pb) for p in probs], handle=h, post=("gauge", "normalize", "flux", "overlaps"), return_info=True)
    prev = None
    for i, ((f_raw, n_raw, _), (f_post, _, _)) in enumerate(zip(raw, out)):
        g, _ = OP.gauge(f_raw.astype(complex))
        fn, fl = OP.normalize(g, wl.coords, plane_bounds=pb)
        assert np.abs(info[i]["flux"] - fl).max() < 1e-9 * np.abs(fl).max()
        assert np.abs(f_post - fn).max() < 1e-9 * np.abs(fn).max()
        assert np.abs(info[i]["te_fraction"] - OP.pol_fraction(g, wl.coords, plane_bounds=pb)).max() < 1e-9
        if prev is not None:
            assert np.abs(info[i]["overlap_prev"] - OP.dot(prev, fn, wl.coords, plane_bounds=pb)).max() < 1e-9
        prev = fn
    # the truncation reaches the result: a mode that fills the plane loses the half cells at the rim
    da_full, da_cut = OP.diff_area(wl.coords), OP.diff_area(wl.coords, plane_bounds=pb)
    assert da_cut.sum() < da_full.sum() - 1e-6
