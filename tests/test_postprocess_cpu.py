"""CPU: the post-processing restatement (oracle/postprocess.py) against the third-party ground truths that exist here, and
the product's host-side mode tracking (tidy3d_b200/postprocess.py) against the restatement."""
import numpy as np
from scipy.interpolate import interp1d

from oracle import postprocess as OP
from tidy3d_b200 import postprocess as PP


def _fields(nx, ny, m, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((2, 3, nx, ny, 1, m)) + 1j * rng.standard_normal((2, 3, nx, ny, 1, m))


def test_colocation_matches_scipy_interp1d():
    """xarray's DataArray.interp (mode_solver.py:507, monitor_data.py:538) is scipy interp1d per axis, NaN outside."""
    nx, ny, m = 9, 7, 2
    f = _fields(nx, ny, m)
    x = np.cumsum(np.r_[0.0, np.random.default_rng(1).uniform(0.5, 1.5, nx)])
    y = np.cumsum(np.r_[0.0, np.random.default_rng(2).uniform(0.5, 1.5, ny)])
    col = OP.colocate(f, [x, y])
    px, py = OP.colocation_points([x, y])
    for name, (kx, ky) in OP.SITES.items():
        a = f[OP.COMP[name][0], OP.COMP[name][1], :, :, 0, :]
        sx = 0.5 * (x[:-1] + x[1:]) if kx == "c" else x[:-1]
        sy = 0.5 * (y[:-1] + y[1:]) if ky == "c" else y[:-1]
        ref = interp1d(sx, a, axis=0, bounds_error=False)(px)
        ref = interp1d(sy, ref, axis=1, bounds_error=False)(py)
        ref = np.nan_to_num(ref, nan=0.0)  # the reference's sums skip NaN
        assert np.abs(col[name] - ref).max() < 1e-13, name


def test_diff_area_is_the_trapezoid_rule():
    x = np.array([0.0, 0.4, 1.0, 1.1, 2.0, 2.7])
    y = np.array([-1.0, -0.2, 0.1, 0.9])
    da = OP.diff_area([x, y])
    px, py = OP.colocation_points([x, y])
    g = np.outer(np.sin(px), np.cos(py))
    assert abs((g * da).sum() - np.trapz(np.trapz(g, py, axis=1), px)) < 1e-14
    assert OP.diff_area([np.array([0.0, 1.0]), y]).shape == (1, 2)  # one-cell axis: size 1


def test_gauge_and_normalisation_properties():
    f = _fields(8, 6, 3, seed=3)
    coords = [np.linspace(0, 1, 9), np.linspace(0, 2, 7)]
    g, phi = OP.gauge(f)
    for m in range(3):
        e = g[0, :2, ..., m]
        v = e.ravel()[np.argmax(np.abs(e))]
        assert abs(v.imag) < 1e-14 and v.real > 0
    fn, fl = OP.normalize(g, coords)
    assert np.allclose(np.abs(OP.flux(fn, coords)), 1.0)
    d = OP.dot(fn, fn, coords)
    assert np.allclose(np.diag(d).real, np.sign(fl))  # dot(mode, itself) = its flux / |flux| (monitor_data.py:640-697)
    assert np.allclose(OP.dot(fn, g, coords), np.conj(OP.dot(g, fn, coords)).T)  # dot(b, a) = conj(dot(a, b))^T


def test_product_overlap_sort_matches_restatement():
    rng = np.random.default_rng(5)
    nf, m = 7, 4
    mats = [None]
    for i in range(1, nf):
        perm = rng.permutation(m) if i in (2, 5) else np.arange(m)
        a = 0.05 * (rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
        a[np.arange(m), perm] += np.exp(1j * rng.uniform(-np.pi, np.pi, m)) * 0.97
        mats.append(a)
    for track in ("central", "lowest", "highest"):
        s0, p0, o0 = OP.overlap_sort(mats[1:], nf, m, track_freq=track)
        s1, p1, o1 = PP.overlap_sort(mats, track_freq=track)
        assert np.array_equal(s0, s1) and np.allclose(p0, p1) and np.allclose(o0, o1)
    # closest pairs on a crafted matrix (monitor_data.py:1421-1440)
    pairs, vals = OP.find_closest_pairs(np.array([[0.1, 0.9, 0.0], [0.8, 0.2, 0.1], [0.0, 0.1, 0.7]]))
    assert list(pairs) == [1, 0, 2] and np.allclose(vals, [0.9, 0.8, 0.7])
    n_sorted, f_sorted = PP.apply_sorting([np.arange(m) + 10 * i for i in range(nf)], [None] * nf, s1, p1)
    assert all(np.array_equal(n_sorted[i], (np.arange(m) + 10 * i)[s1[i]]) for i in range(nf))


def test_filter_polarization_order():
    """mode_solver.py:523-549: requested polarisation first, NaN fractions last."""
    te = np.array([0.9, 0.2, np.nan, 0.5, 0.7])
    assert list(PP.filter_polarization(te, "te")) == [0, 3, 4, 1, 2]
    assert list(PP.filter_polarization(te, "tm")) == [1, 3, 0, 4, 2]
