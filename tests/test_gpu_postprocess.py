"""GPU: on-device post-processing (csrc/post.cuh: gauge, colocated flux, normalisation, modal overlaps between adjacent
frequencies) against the numpy restatement oracle/postprocess.py applied to the raw device fields."""
import numpy as np
import pytest

from oracle import postprocess as OP
from tests.golden.cases import CASES
from tidy3d_b200 import compute_modes_batch
from tidy3d_b200 import postprocess as PP
from tidy3d_b200 import workloads as W
from tidy3d_b200.solver import get_handle

pytestmark = pytest.mark.gpu
ALL = ("gauge", "normalize", "flux", "overlaps")


def _sweep(n=96, nf=5, **spec):
    wl = W.c2(nf=nf, n=n)
    for k, v in spec.items():
        setattr(wl.mode_spec, k, v)
    return wl, [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in wl.freqs]


def test_gauge_flux_normalisation_and_overlaps_match_the_restatement():
    wl, probs = _sweep()
    h = get_handle(tolerance="tight")
    raw = compute_modes_batch(probs, handle=h)
    out, info = compute_modes_batch(probs, handle=h, post=ALL, return_info=True)
    prev = None
    for i, ((f_raw, n_raw, _), (f_post, n_post, _)) in enumerate(zip(raw, out)):
        assert np.array_equal(n_raw, n_post)
        g, _ = OP.gauge(f_raw)
        fn, fl = OP.normalize(g, wl.coords)
        assert np.abs(info[i]["flux"] - fl).max() < 1e-10 * np.abs(fl).max()
        te = OP.pol_fraction(f_raw, wl.coords)
        assert np.abs(info[i]["te_fraction"] - te).max() < 1e-10
        assert list(PP.filter_polarization(info[i]["te_fraction"], "tm")) == list(np.concatenate((np.where(te <= 0.5)[0], np.where(te > 0.5)[0])))
        assert np.abs(f_post - fn).max() < 1e-9 * np.abs(fn).max()
        for m in range(fn.shape[-1]):  # gauge: the largest in-plane E entry is real positive (mode_solver.py:806-810)
            e = f_post[0, :2, ..., m]
            v = e.ravel()[np.argmax(np.abs(e))]
            assert abs(v.imag) < 1e-12 * abs(v) and v.real > 0
        assert np.allclose(np.abs(OP.flux(f_post, wl.coords)), 1.0, atol=1e-9)
        if prev is not None:
            ref = OP.dot(prev, fn, wl.coords)
            assert np.abs(info[i]["overlap_prev"] - ref).max() < 1e-9
            assert (np.abs(np.diag(ref)) > 0.99).all()  # adjacent sweep points: same physical modes
        else:
            assert not info[i]["overlap_prev"].any()
        prev = fn
    # mode tracking from the device overlaps: a smooth sweep keeps its ordering (monitor_data.py:1295-1375)
    sorting, phase, ov = PP.overlap_sort([inf["overlap_prev"] for inf in info])
    assert (sorting == np.arange(4)).all() and (ov > 0.99).all()
    s0, p0, o0 = OP.overlap_sort([inf["overlap_prev"] for inf in info][1:], len(info), 4)
    assert np.array_equal(sorting, s0) and np.allclose(phase, p0)


def test_small_results_only_mode():
    """want_fields=False with post-processing: only n_complex, flux and the overlap matrices leave the GPU."""
    wl, probs = _sweep(n=64, nf=3)
    full, info_full = compute_modes_batch(probs, post=ALL, return_info=True)
    out, info = compute_modes_batch(probs, post=ALL, return_info=True, want_fields=False)
    for a, b, c in zip(info_full, info, out):
        assert c[0] is None
        assert np.allclose(a["flux"], b["flux"], rtol=1e-12) and np.allclose(a["overlap_prev"], b["overlap_prev"], atol=1e-12)


@pytest.mark.parametrize("name", ["c1_64_sym_pmc_pec", "slab1d_x1", "nonuniform_56", "c1_64_single"])
def test_flux_edge_cases(name):
    """Symmetry planes (first boundary kept, integral doubled), a one-cell axis (no interpolation, unit size), a graded
    grid, complex64 fields."""
    fac, kw, _ = CASES[name]
    wl = fac()
    sym = kw.get("symmetry", (0, 0))
    probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)]
    raw = compute_modes_batch(probs)[0][0]
    out, info = compute_modes_batch(probs, post=("normalize", "flux"), return_info=True)
    fl = OP.flux(raw.astype(complex), wl.coords, sym)
    tol = 1e-5 if raw.dtype == np.complex64 else 1e-10
    assert np.abs(info[0]["flux"] - fl).max() < tol * np.abs(fl).max()
    fn = raw.astype(complex) / np.sqrt(np.abs(fl))
    assert np.abs(out[0][0] - fn).max() < max(tol, 1e-9) * np.abs(fn).max()
