"""Post-processing pinning cases: seeded inputs, the quantities the UNMODIFIED reference produces for them
(``reference_results``: oracle/ref_post.py runs the reference's own method bodies; build container only) and the same
quantities from the restatement (``oracle_results``: oracle/postprocess.py).  ``tests/golden/make_post_golden.py`` stores the
former as ``tests/golden/post_<case>.npz``; ``tests/test_postprocess_pinning.py`` compares.

Input of every case = what ``compute_modes`` returns per frequency, in solver-plane axes: fields (2,3,Nx,Ny,1,M), n_complex (M,).
The fields are random (every term of every formula is exercised: no component vanishes, nothing is symmetric by accident)
and, for the tracking cases, random perturbations / permutations of the previous frequency's modes.
"""
from __future__ import annotations

import numpy as np

from oracle import postprocess as OP

CASES = {
    # name: nx, ny, M, F, symmetry, direction, theta, phi, track_freq, filter_pol, normal grid (primal, dual, pos) or None, seed
    "plain": dict(nx=9, ny=7, m=3, nf=3, seed=1),
    "sym_pmc_x": dict(nx=8, ny=6, m=3, nf=2, symmetry=(1, 0), seed=2),
    "sym_pec_y": dict(nx=7, ny=8, m=2, nf=2, symmetry=(0, -1), seed=3),
    "sym_both": dict(nx=6, ny=6, m=3, nf=2, symmetry=(-1, 1), seed=4),
    "grid_corr_minus": dict(nx=8, ny=7, m=3, nf=3, direction="-", normal=([-0.043, 0.061], [-0.095, 0.009], 0.002), seed=5),
    "angled": dict(nx=7, ny=9, m=3, nf=2, theta=0.21, phi=0.37, normal=([-0.05, 0.03], [-0.01, 0.07], 0.0), seed=6),
    "one_cell_y": dict(nx=12, ny=1, m=2, nf=2, seed=7),
    "track_lowest": dict(nx=9, ny=8, m=4, nf=5, track="lowest", coherent=True, seed=8),
    "track_central_swap": dict(nx=9, ny=8, m=4, nf=6, track="central", coherent=True, swaps=(1, 4), seed=9),
    "track_highest_sym": dict(nx=7, ny=7, m=3, nf=4, track="highest", coherent=True, swaps=(2,), symmetry=(1, -1), direction="-", seed=10),
    "filter_te": dict(nx=8, ny=8, m=4, nf=2, filter_pol="te", seed=11),
    "filter_tm_track": dict(nx=8, ny=7, m=4, nf=3, filter_pol="tm", track="central", coherent=True, swaps=(2,), seed=12),
    "finite_plane": dict(nx=9, ny=8, m=2, nf=2, plane_cut=(0.37, 0.61, 0.2, 0.45), seed=13),
    # an angled plane WITH symmetry walls (unphysical, but the reference computes something definite: found by fuzzing)
    "angled_sym": dict(nx=6, ny=7, m=3, nf=2, symmetry=(1, -1), theta=0.19, phi=-0.53, direction="-", seed=14),
    "angled_sym_x": dict(nx=5, ny=6, m=2, nf=2, symmetry=(-1, 0), theta=-0.22, phi=0.31, filter_pol="te", seed=15),
}


def inputs(name):
    """Deterministic inputs of a case: dict(coords, fields [F x (2,3,nx,ny,1,M)], n_complex [F x (M,)], freqs, + options)."""
    c = dict(symmetry=(0, 0), direction="+", theta=0.0, phi=0.0, track=None, filter_pol=None, normal=None, coherent=False, swaps=(),
             plane_cut=None)
    c.update(CASES[name])
    rng = np.random.default_rng(c["seed"])
    nx, ny, m, nf = c["nx"], c["ny"], c["m"], c["nf"]
    x = np.cumsum(np.r_[rng.uniform(-0.2, 0.2), rng.uniform(0.04, 0.11, nx)])
    y = np.cumsum(np.r_[rng.uniform(-0.2, 0.2), rng.uniform(0.04, 0.11, ny)])

    def rnd():
        return rng.standard_normal((2, 3, nx, ny, 1, m)) + 1j * rng.standard_normal((2, 3, nx, ny, 1, m))

    def mode_like():
        """Modes as a waveguide has them: well separated (here: supported on different x-blocks), H tied to E by the
        plane-wave relation (positive flux), a little noise on top -- so that overlaps are close to a permutation matrix."""
        f = 0.03 * rnd()
        blocks = np.array_split(np.arange(nx), m)
        for k, b in enumerate(blocks):
            f[0, :, b[0]:b[-1] + 1, :, :, k] += (rng.standard_normal((3, b.size, ny, 1)) + 1j * rng.standard_normal((3, b.size, ny, 1)))
        f[1] *= 2.6e-3
        f[1, 1] += 2.0 * 2.6e-3 * f[0, 0]  # Hy = n Ex / eta0
        f[1, 0] -= 2.0 * 2.6e-3 * f[0, 1]  # Hx = -n Ey / eta0
        return f

    fields = [mode_like() if c["coherent"] else rnd()]
    if not c["coherent"]:
        fields[0][1] *= 2.6e-3  # H ~ E / eta0
    for i in range(1, nf):
        if c["coherent"]:  # the next frequency's modes: the previous ones, perturbed, re-phased and (at `swaps`) permuted
            f = fields[-1] + 0.03 * rnd() * np.array([1.0, 2.6e-3]).reshape(2, 1, 1, 1, 1, 1)
            f = f * np.exp(1j * rng.uniform(-np.pi, np.pi, m))
            if i in c["swaps"]:
                f = f[..., np.roll(np.arange(m), 1)]
        else:
            f = rnd()
            f[1] *= 2.6e-3
        fields.append(f)
    n_complex = [np.sort(rng.uniform(1.5, 2.5, m))[::-1] + 1e-3j * rng.uniform(0, 1, m) for _ in range(nf)]
    freqs = list(np.linspace(1.9e14, 2.0e14, nf))
    c.update(coords=[x, y], fields=fields, n_complex=n_complex, freqs=freqs)
    if c["plane_cut"] is not None:  # a finite mode plane whose edges cut through cells: (fractions of the first / last cell kept)
        ax, bx, ay, by = c["plane_cut"]
        lo = (x[1] + ax * (x[2] - x[1]), y[1] + ay * (y[2] - y[1]))
        hi = (x[-2] - bx * (x[-2] - x[-3]), y[-2] - by * (y[-2] - y[-3]))
        c["plane_bounds"] = (lo[0], hi[0], lo[1], hi[1])
    else:
        c["plane_bounds"] = None
    return c


def compare(ref, got, skip=()):
    """{key: max |ref - got| / max |ref|} over ``KEYS`` (NaN == NaN; shapes must agree)."""
    err = {}
    for k in KEYS:
        if k in skip:
            continue
        a, b = np.asarray(ref[k]), np.asarray(got[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        a, b = np.nan_to_num(a), np.nan_to_num(b)
        err[k] = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)) if a.size else 0.0
    return err


def skipped_keys(c):
    """``outer_dot`` interpolates the other data set along EVERY tangential axis (monitor_data.py:709-718), also along a
    one-point axis of a 2-D simulation, where the result is whatever xarray / scipy make of a one-point interpolation (NaN
    from scipy's interp1d): not a property of the reference's arithmetic, not pinned.  ``dot``, which ``overlap_sort``
    evaluates first (:1386), is."""
    return ("outer_next",) if min(c["nx"], c["ny"]) == 1 else ()


KEYS = ("primal", "dual", "flux_yee", "te_fraction", "normalized_yee", "colocated", "dot_next", "outer_next", "sorting", "phase", "final_yee",
        "final_n_complex")


ARRAYS = ("x", "y", "fields", "n_complex", "freqs")


def to_arrays(c):
    """The array-valued inputs of a case, as stored in the fixture next to the reference's results."""
    return dict(x=c["coords"][0], y=c["coords"][1], fields=np.array(c["fields"]), n_complex=np.array(c["n_complex"]), freqs=np.array(c["freqs"]))


def from_arrays(name, z):
    """Case dict from the options of ``CASES[name]`` and the arrays of a fixture (independent of the RNG stream of this numpy)."""
    c = inputs(name)
    c.update(coords=[np.array(z["x"]), np.array(z["y"])], fields=list(np.array(z["fields"])), n_complex=list(np.array(z["n_complex"])),
             freqs=list(np.array(z["freqs"])))
    return c


def reference_results(c):
    """The reference's own code (oracle/ref_post.py) on the inputs ``c``."""
    from oracle import ref_post as RP

    pb = c["plane_bounds"]
    extra = {}
    if pb is not None:
        extra = dict(plane_center=(0.5 * (pb[0] + pb[1]), 0.5 * (pb[2] + pb[3])), plane_size=(pb[1] - pb[0], pb[3] - pb[2]))
    if c["normal"] is not None:
        extra.update(normal_primal=c["normal"][0], normal_dual=c["normal"][1], normal_pos=c["normal"][2])

    def make(colocate, track, filter_pol):
        return RP.solver(c["fields"], c["n_complex"], c["coords"], c["freqs"], symmetry=c["symmetry"], direction=c["direction"],
                         colocate=colocate, angle_theta=c["theta"], angle_phi=c["phi"], filter_pol=filter_pol, track_freq=track, **extra)

    out = {}
    yee = make(False, None, None)._data_on_yee_grid()  # gauge only (mode_solver.py:340-415)
    out["primal"], out["dual"] = yee.grid_primal_correction.values, yee.grid_dual_correction.values
    out["flux_yee"] = yee.flux.values
    out["te_fraction"] = yee.pol_fraction.te.values
    norm = make(False, None, None).data_raw  # + normalisation (mode_solver.py:517-521)
    out["normalized_yee"] = RP.packed(norm)
    out["colocated"] = RP.packed(make(True, None, None).data_raw)  # + colocation before the normalisation (:490-515)
    nf = len(c["freqs"])
    dots, outers = [], []
    for i in range(nf - 1):  # what overlap_sort evaluates between neighbouring frequencies (monitor_data.py:1337-1347)
        a = norm._isel(f=[i])
        b = norm._isel(f=[i + 1])._assign_coords(f=[c["freqs"][i]])
        dots.append(a.dot(b).values.ravel())
        outers.append(a.outer_dot(b).to_numpy()[0])
    out["dot_next"], out["outer_next"] = np.array(dots).reshape(nf - 1, c["m"]), np.array(outers).reshape(nf - 1, c["m"], c["m"])
    ms = make(False, c["track"], c["filter_pol"])
    final = ms.data_raw  # the whole of data_raw: + polarisation filter (:523-549) + mode tracking (monitor_data.py:1295-1505)
    rec = RP.last_reorder()
    out["sorting"] = rec[0] if (c["track"] and nf > 1) else np.tile(np.arange(c["m"]), (nf, 1))
    out["phase"] = rec[1] if (c["track"] and nf > 1) else np.zeros((nf, c["m"]))
    out["final_yee"] = RP.packed(final)
    out["final_n_complex"] = final.n_complex.values
    return out


def oracle_results(c):
    """oracle/postprocess.py on the same inputs, step by step as ``ModeSolver.data_raw`` orders them."""
    coords, sym, nf, m = c["coords"], c["symmetry"], len(c["freqs"]), c["m"]
    pb = c["plane_bounds"]
    gauged = [OP.gauge(f)[0] for f in c["fields"]]
    corr = []
    for i in range(nf):
        if c["normal"] is None:
            corr.append((np.ones(m, complex), np.ones(m, complex)))
        else:
            corr.append(OP.grid_correction(c["n_complex"][i], c["freqs"][i], c["normal"][0], c["normal"][1], c["normal"][2], c["theta"], c["direction"]))
    out = dict(primal=np.array([p for p, _ in corr]), dual=np.array([d for _, d in corr]))
    flux = [OP.flux(g, coords, sym, cr, plane_bounds=pb) for g, cr in zip(gauged, corr)]
    out["flux_yee"] = np.array(flux)
    out["te_fraction"] = np.array([OP.pol_fraction(g, coords, sym, c["theta"], c["phi"], plane_bounds=pb) for g in gauged])
    norm = [g / np.sqrt(np.abs(fl)) for g, fl in zip(gauged, flux)]

    def pack(per_freq):  # [F x (2,3,Px,Py,M)] -> (2,3,Px,Py,F,M)
        return np.stack(per_freq, axis=-2)

    out["normalized_yee"] = pack([n[:, :, :, :, 0, :] for n in norm])
    col = []
    for g, cr in zip(gauged, corr):  # colocate first, then normalise the colocated data by ITS flux (no second interpolation)
        cc = OP.colocate(g, coords, sym)
        arr = np.array([[cc["Ex"], cc["Ey"], cc["Ez"]], [cc["Hx"], cc["Hy"], cc["Hz"]]])
        col.append(arr / np.sqrt(np.abs(OP.flux(g, coords, sym, cr, plane_bounds=pb))))
    out["colocated"] = pack(col)
    outers = [OP.dot(norm[i], norm[i + 1], coords, sym, True, corr[i], corr[i + 1], plane_bounds=pb) for i in range(nf - 1)]
    out["outer_next"] = np.array(outers).reshape(nf - 1, m, m)
    out["dot_next"] = np.array([np.diag(o) for o in outers]).reshape(nf - 1, m)
    # data_raw: filter first, then track (mode_solver.py:331-337)
    n_c = [np.array(n) for n in c["n_complex"]]
    cur, cur_corr = list(norm), list(corr)
    if c["filter_pol"] is not None:
        from tidy3d_b200 import postprocess as PP  # host half of the product (pinned to this restatement in test_postprocess_cpu.py)

        for i in range(nf):
            order = PP.filter_polarization(out["te_fraction"][i], c["filter_pol"])
            cur[i], n_c[i] = cur[i][..., order], n_c[i][order]
            cur_corr[i] = (cur_corr[i][0][order], cur_corr[i][1][order])
    sorting, phase = np.tile(np.arange(m), (nf, 1)), np.zeros((nf, m))
    if c["track"] and nf > 1:
        ov = [OP.dot(cur[i], cur[i + 1], coords, sym, True, cur_corr[i], cur_corr[i + 1], plane_bounds=pb) for i in range(nf - 1)]
        sorting, phase, _ = OP.overlap_sort(ov, nf, m, c["track"], 0.9, c["direction"])
    out["sorting"], out["phase"] = sorting, phase
    out["final_yee"] = pack([(cur[i][..., sorting[i]] * np.exp(-1j * phase[i]))[:, :, :, :, 0, :] for i in range(nf)])
    out["final_n_complex"] = np.array([n_c[i][sorting[i]] for i in range(nf)])
    return out


def random_case(rng):
    """A random case description (for ``CASES``-style use) covering the option space: sizes incl. one-cell axes, both symmetry
    kinds, angles, one- / two-point normal grids, direction, tracking, polarisation filters, finite planes.  Excluded, because the
    reference itself cannot do them: a two-cell axis (its ``_diff_area`` and its colocation disagree about the number of points)
    and mode tracking on a one-cell axis (``outer_dot`` interpolates along the one-point axis, see ``skipped_keys``)."""
    nx, ny = (int(rng.choice([1, 3, 4, 5, 6, 7, 8])) for _ in range(2))
    if nx == 1 and ny == 1:
        nx = 3
    sym = (int(rng.choice([0, 0, 1, -1])) if nx > 1 else 0, int(rng.choice([0, 0, 1, -1])) if ny > 1 else 0)
    m, nf = int(rng.integers(1, 5)), int(rng.integers(1, 4))
    case = dict(nx=nx, ny=ny, m=m, nf=nf, symmetry=sym, direction=str(rng.choice(["+", "-"])), seed=int(rng.integers(1 << 30)))
    if rng.random() < 0.5:
        case.update(theta=float(rng.uniform(-0.4, 0.4)), phi=float(rng.uniform(-1, 1)))
    r = rng.random()
    if r < 0.3:
        case["normal"] = ([-0.04, 0.05], [-0.09, 0.01], float(rng.uniform(-0.03, 0.0)))
    elif r < 0.5:
        case["normal"] = ([0.013], [0.04], 0.013)
    if rng.random() < 0.5 and nf > 1 and min(nx, ny) > 1:
        case.update(track=str(rng.choice(["central", "lowest", "highest"])), coherent=bool(rng.random() < 0.7), swaps=(1,) if rng.random() < 0.5 else ())
    if rng.random() < 0.3 and m > 1:
        case["filter_pol"] = str(rng.choice(["te", "tm"]))
    if sym == (0, 0) and nx >= 4 and ny >= 4 and rng.random() < 0.3:
        case["plane_cut"] = tuple(float(v) for v in rng.uniform(0.05, 0.95, 4))
    return case
