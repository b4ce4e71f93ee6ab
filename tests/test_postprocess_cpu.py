"""CPU: the post-processing restatement (oracle/postprocess.py) against the third-party ground truths that exist here, and
the product's host-side mode tracking (tidy3d_b200/postprocess.py) against the restatement."""
import numpy as np
import pytest
from scipy.interpolate import interp1d

_trapz = getattr(np, "trapezoid", None) or np.trapz  # numpy >= 2.0 renamed it

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
    assert abs((g * da).sum() - _trapz(_trapz(g, py, axis=1), px)) < 1e-14
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


def _mirror_full(fields, coords, symmetry):
    """Half-domain Yee data -> the full-domain Yee data the reference's ``symmetry_expanded`` represents (monitor_data.py:237-282),
    laid out on the full grid's own Yee sites (the lower-boundary site on the far mirrored wall is never read: 0)."""
    f = np.asarray(fields)
    coords = [np.asarray(c, float) for c in coords]
    for ax in (0, 1):
        s0 = symmetry[ax]
        if s0 == 0:
            continue
        c = coords[ax]
        n = c.size - 1
        out = np.zeros(f.shape[:2 + ax] + (2 * n,) + f.shape[3 + ax:], dtype=complex)
        for name, (fi, ci) in OP.COMP.items():
            kind = OP.SITES[name][ax]
            sign = s0 * OP.symmetry_eigenvalue(name, ax)
            half = f[fi, ci]
            if kind == "c":
                full = np.concatenate([sign * np.flip(half, axis=ax), half], axis=ax)
            else:
                inner = np.flip(np.take(half, range(1, n), axis=ax), axis=ax)
                full = np.concatenate([np.zeros_like(np.take(half, [0], axis=ax)), sign * inner, half], axis=ax)
            out[fi, ci] = full
        f = out
        coords[ax] = np.concatenate([2 * c[0] - c[:0:-1], c])
    return f, coords


@pytest.mark.parametrize("symmetry", [(-1, 0), (1, 0), (0, -1), (0, 1), (-1, 1), (1, -1), (-1, -1)])
def test_symmetry_plane_equals_the_mirrored_full_domain(symmetry):
    """The reference colocates and integrates the SYMMETRY-EXPANDED data (mode_solver.py:504-507, monitor_data.py:517, 237-282):
    flux, TE fraction and modal overlaps of a half domain with symmetry planes must equal those of the mirrored full
    domain without.  At a PEC plane (-1) Ex, Hy, Hz are even across the plane and keep their value there; dropping that
    point (an earlier reading of the code) loses half a cell of an even mode's power."""
    rng = np.random.default_rng(4)
    nx, ny, m = 9, 7, 3
    x = np.cumsum(np.r_[0.3, rng.uniform(0.05, 0.12, nx)])
    y = np.cumsum(np.r_[-0.2, rng.uniform(0.05, 0.12, ny)])
    fa = rng.standard_normal((2, 3, nx, ny, 1, m)) + 1j * rng.standard_normal((2, 3, nx, ny, 1, m))
    fb = rng.standard_normal((2, 3, nx, ny, 1, m)) + 1j * rng.standard_normal((2, 3, nx, ny, 1, m))
    full_a, full_coords = _mirror_full(fa, [x, y], symmetry)
    full_b, _ = _mirror_full(fb, [x, y], symmetry)
    assert np.allclose(OP.flux(fa, [x, y], symmetry), OP.flux(full_a, full_coords), rtol=1e-12, atol=1e-14)
    assert np.allclose(OP.pol_fraction(fa, [x, y], symmetry), OP.pol_fraction(full_a, full_coords), rtol=1e-12)
    assert np.allclose(OP.dot(fa, fb, [x, y], symmetry), OP.dot(full_a, full_b, full_coords), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("symmetry", [(0, 0), (-1, 0), (1, 0), (0, -1), (-1, 1), (1, -1)])
def test_library_post_tables_equal_the_symmetry_expanded_restatement(built_lib, symmetry):
    """The 1-D interpolation / integration tables the library uploads for its post-processing kernels (host code of
    csrc/api.cu, b200ms_debug_post_tables) applied in numpy: colocated fields, flux and overlaps equal the restatement of
    the reference's symmetry-expanded colocation, including the point on a PEC / PMC symmetry plane."""
    import ctypes as C

    rng = np.random.default_rng(9)
    nx, ny, m = 11, 8, 2
    x = np.cumsum(np.r_[0.0, rng.uniform(0.05, 0.12, nx)])
    y = np.cumsum(np.r_[-0.4, rng.uniform(0.05, 0.12, ny)])
    f = _fields(nx, ny, m, seed=2)
    tabs = []
    for c, n, s in ((x, nx, symmetry[0]), (y, ny, symmetry[1])):
        idx, wgt, area = np.zeros(4 * (n + 1), np.int32), np.zeros(4 * (n + 1)), np.zeros(n + 1)
        P = built_lib.lib().b200ms_debug_post_tables(built_lib._ptr(np.ascontiguousarray(c)), n, s, n + 1, idx.ctypes.data_as(built_lib._ip),
                                                     built_lib._ptr(wgt), built_lib._ptr(area))
        assert P == (n - 1 if s == 0 else n)
        tabs.append((idx[: 4 * P].reshape(P, 4), wgt[: 4 * P].reshape(P, 4), area[:P]))
    want = OP.colocate(f, [x, y], symmetry)
    got = {}
    for name, kinds in OP.SITES.items():
        g = f[OP.COMP[name][0], OP.COMP[name][1], :, :, 0, :]
        for ax, kind in enumerate(kinds):
            idx, wgt, _ = tabs[ax]
            o = 0 if kind == "c" else 2
            shape = [1, 1, 1]
            shape[ax] = -1
            g = np.take(g, idx[:, o], axis=ax) * wgt[:, o].reshape(shape) + np.take(g, idx[:, o + 1], axis=ax) * wgt[:, o + 1].reshape(shape)
        got[name] = g
        assert np.abs(g - want[name]).max() < 1e-13, name
    da = np.outer(tabs[0][2], tabs[1][2])
    assert np.abs(da - OP.diff_area([x, y], symmetry)).max() < 1e-15
    mult = 2 ** sum(1 for q in symmetry if q != 0)
    fl = mult * np.einsum("xym,xy->m", 0.5 * np.real(got["Ex"] * np.conj(got["Hy"]) - got["Ey"] * np.conj(got["Hx"])), da)
    assert np.allclose(fl, OP.flux(f, [x, y], symmetry), rtol=1e-12)


def test_grid_correction_factors_host_helper_and_library_equal_the_restatement(built_lib):
    """ModeSolver._grid_correction (mode_solver.py:847-904) three ways: the restatement (exp(i k r) sampled on the whole normal
    grid, then np.interp like DataArray.interp), the Python helper (bracketing points + weights) and the library's host
    routine behind b200ms_problem.grid_correction -- lossy and backward modes, an angled plane, a plane exactly on a grid
    boundary (primal factor exactly 1) and a one-point grid (squeeze instead of interp)."""
    import ctypes as C

    from tidy3d_b200 import workloads as W

    rng = np.random.default_rng(2)
    bounds = np.cumsum(np.r_[-0.3, rng.uniform(0.02, 0.05, 12)])
    centers = (bounds[:-1] + bounds[1:]) / 2
    n = np.array([2.45 + 0.0j, 1.86 + 3.3e-3j, 1.43 + 0.12j])
    freq = W.C_0 / 1.55
    wl = W.c1()
    for pos, theta, direction, primal_pts, dual_pts in (
        (0.5 * (bounds[4] + centers[4]), 0.0, "+", bounds, centers),
        (bounds[6], 0.0, "-", bounds, centers),
        (0.3 * bounds[7] + 0.7 * bounds[8], 0.2, "+", bounds, centers),
        (0.013, 0.0, "+", np.array([0.0]), np.array([0.025])),
    ):
        want_p, want_d = OP.grid_correction(n, freq, primal_pts, dual_pts, pos, theta, direction)
        table = PP.grid_correction_table(primal_pts, dual_pts, pos)
        got_p, got_d = PP.grid_correction_factors(n, freq, table, theta, direction)
        assert np.abs(got_p - want_p).max() < 1e-14 and np.abs(got_d - want_d).max() < 1e-14
        if pos == bounds[6]:
            assert np.abs(want_p - 1).max() < 1e-15 and np.abs(want_d).max() < 1 - 1e-4  # dual ~ cos(k dl / 2)
        spec = W.ModeSpecLike(num_modes=3, angle_theta=theta)
        pk = built_lib.PackedProblem(wl.eps_cross, wl.coords, freq, spec, direction=direction, grid_correction=table)
        lp, ld = np.zeros(3, complex), np.zeros(3, complex)
        rc = built_lib.lib().b200ms_debug_grid_factors(C.byref(pk.struct), built_lib._ptr(n.view(float)), built_lib._ptr(lp.view(float)),
                                                       built_lib._ptr(ld.view(float)))
        assert rc == 0 and np.abs(lp - want_p).max() < 1e-14 and np.abs(ld - want_d).max() < 1e-14
    # without a table the factors are 1
    pk = built_lib.PackedProblem(wl.eps_cross, wl.coords, freq, W.ModeSpecLike(num_modes=3))
    assert built_lib.lib().b200ms_debug_grid_factors(C.byref(pk.struct), built_lib._ptr(n.view(float)), built_lib._ptr(lp.view(float)),
                                                     built_lib._ptr(ld.view(float))) == 0
    assert np.array_equal(lp, np.ones(3)) and np.array_equal(ld, np.ones(3))


def test_pol_fraction_in_propagation_axes_closed_form():
    """The two rotations of monitor_data.py:1603-1607 written out (what csrc/post.cuh post_scan_kernel evaluates per point):
    E1 = cos(theta) (cos(phi) Ex + sin(phi) Ey) - sin(theta) Ez,  E2 = cos(phi) Ey - sin(phi) Ex."""
    f = _fields(9, 8, 3, seed=5)
    x, y = np.linspace(0, 1, 10), np.linspace(-0.4, 0.4, 9)
    for theta, phi in ((0.2, 0.0), (0.0, 0.7), (-0.3, 1.1)):
        c = OP.colocate(f, [x, y])
        da = OP.diff_area([x, y])
        e1 = np.cos(theta) * (np.cos(phi) * c["Ex"] + np.sin(phi) * c["Ey"]) - np.sin(theta) * c["Ez"]
        e2 = np.cos(phi) * c["Ey"] - np.sin(phi) * c["Ex"]
        te, tm = np.einsum("xym,xy->m", np.abs(e1) ** 2, da), np.einsum("xym,xy->m", np.abs(e2) ** 2, da)
        assert np.allclose(OP.pol_fraction(f, [x, y], angle_theta=theta, angle_phi=phi), te / (te + tm), rtol=1e-13)
    assert np.allclose(OP.pol_fraction(f, [x, y]), OP.pol_fraction(f, [x, y], angle_theta=0.0, angle_phi=0.0))


def test_library_post_tables_on_random_axes(built_lib):
    """Seeded sweep: the library's interpolation / integration tables (and ``postprocess.colocate`` built on them) against the
    restatement for random graded axes incl. one-cell axes, both symmetry kinds and finite planes (400 such cases were run
    when this test was written)."""
    rng = np.random.default_rng(5)
    for _ in range(60):
        nx, ny = int(rng.choice([1, 3, 4, 5, 7, 11])), int(rng.choice([1, 3, 4, 6, 9]))
        if nx == 1 and ny == 1:
            ny = 4
        x = np.cumsum(np.r_[rng.uniform(-1, 1), rng.uniform(0.03, 0.2, nx)])
        y = np.cumsum(np.r_[rng.uniform(-1, 1), rng.uniform(0.03, 0.2, ny)])
        sym = (int(rng.choice([0, 1, -1])) if nx > 1 else 0, int(rng.choice([0, 1, -1])) if ny > 1 else 0)
        f = _fields(nx, ny, 2, seed=int(rng.integers(1 << 30)))
        pb = None
        if sym == (0, 0) and nx >= 4 and ny >= 4 and rng.random() < 0.5:
            a = rng.uniform(0.05, 0.95, 4)
            pb = (x[1] + a[0] * (x[2] - x[1]), x[-2] - a[1] * (x[-2] - x[-3]), y[1] + a[2] * (y[2] - y[1]), y[-2] - a[3] * (y[-2] - y[-3]))
        got, _ = PP.colocate(f, [x, y], sym)
        want = OP.colocate(f, [x, y], sym)
        for k, name in enumerate(["Ex", "Ey", "Ez", "Hx", "Hy", "Hz"]):
            assert np.abs(got[k // 3, k % 3, :, :, 0, :] - want[name]).max() < 1e-13, (nx, ny, sym, name)
        areas = []
        for ax, (c, n, s) in enumerate(((x, nx, sym[0]), (y, ny, sym[1]))):
            idx, wgt, area = np.zeros(4 * (n + 1), np.int32), np.zeros(4 * (n + 1)), np.zeros(n + 1)
            lo, hi = (pb[2 * ax], pb[2 * ax + 1]) if pb else (-np.inf, np.inf)
            P = built_lib.lib().b200ms_debug_post_tables_bounded(built_lib._ptr(np.ascontiguousarray(c)), n, s, lo, hi, n + 1,
                                                                 idx.ctypes.data_as(built_lib._ip), built_lib._ptr(wgt), built_lib._ptr(area))
            areas.append(area[:P].copy())
        assert np.abs(np.outer(*areas) - OP.diff_area([x, y], sym, pb)).max() < 1e-15, (nx, ny, sym, pb)
