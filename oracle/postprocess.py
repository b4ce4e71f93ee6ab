"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the post-processing ``ModeSolver.data_raw`` applies to the output of
``compute_modes``, in solver-plane coordinates (propagation along z; the rotation to the simulation's
axes, mode_solver.py:788-792, is a relabelling of arrays and is not restated).

PARITY PINNED (since round 2, last session): ``ModeSolverData`` cannot be imported here (xarray / shapely / h5py are not in
the image), but the reference's own METHOD BODIES can be executed: oracle/ref_post.py cuts them out of the
reference's files at run time and runs them over a small labelled-array stand-in for xarray (oracle/mini_xarray.py).  What
they produce for the seeded cases of tests/post_cases.py -- the whole of ``ModeSolver.data_raw`` (gauge, colocation, flux
normalisation with finite-grid correction, polarisation filter, ``overlap_sort``), ``flux``, ``pol_fraction``, ``dot``,
``outer_dot``, with symmetry planes, a one-cell axis, an angled plane, direction "-" and a finite plane cutting through
cells -- is committed as tests/golden/post_*.npz (generator tests/golden/make_post_golden.py) and reproduced by this file to
1e-12 (measured 4e-14) in tests/test_postprocess_pinning.py; the ``reference``-marked tests there repeat it live.  The
stand-in's own semantics (broadcast by name, inner join, NaN-skipping sums, interp = scipy ``interp1d``) are tested in the
same file; the older third-party pins (tests/test_postprocess_cpu.py: colocation against ``scipy.interpolate.interp1d``,
integration weights against ``numpy.trapz``) stay.

Yee sites of the six components in the solver plane (tidy3d/components/grid/grid.py ``Grid.yee``; c = cell centre,
b = lower cell boundary):  Ex (c,b)  Ey (b,c)  Ez (b,b)  Hx (b,c)  Hy (c,b)  Hz (c,c).
"""
from __future__ import annotations

import numpy as np

SITES = {"Ex": ("c", "b"), "Ey": ("b", "c"), "Ez": ("b", "b"), "Hx": ("b", "c"), "Hy": ("c", "b"), "Hz": ("c", "c")}
COMP = {"Ex": (0, 0), "Ey": (0, 1), "Ez": (0, 2), "Hx": (1, 0), "Hy": (1, 1), "Hz": (1, 2)}


def gauge(fields):
    """mode_solver.py:802-810: per mode, rotate the phase so that the largest-|.| in-plane E entry is real positive.
    ``fields`` (2,3,Nx,Ny,1,M) -> (gauged copy, phi[M]).  ``np.argmax`` takes the FIRST maximum in C order of E[:2]."""
    out = np.array(fields, dtype=complex, copy=True)
    phis = np.zeros(out.shape[-1])
    for m in range(out.shape[-1]):
        e = out[0, :2, ..., m]
        ind = np.argmax(np.abs(e))
        phi = np.angle(e.ravel()[ind])
        out[..., m] *= np.exp(-1j * phi)
        phis[m] = phi
    return out, phis


def colocation_points(coords, symmetry=(0, 0)):
    """mode_solver.py:494-502: interior cell boundaries; with a symmetry plane at the min side the first boundary is kept."""
    pts = []
    for c, s in zip(coords, symmetry):
        c = np.asarray(c, float)
        if c.size > 2:
            pts.append(c[1:-1] if s == 0 else c[:-1])
        else:
            pts.append(None)  # a one-cell axis is not interpolated
    return pts


def _site_coords(coords, kind):
    c = np.asarray(coords, float)
    return 0.5 * (c[:-1] + c[1:]) if kind == "c" else c[:-1]


def interp_weights(src, dst):
    """Linear interpolation src -> dst as (i0, w0, i1, w1) per destination point (what ``DataArray.interp`` does along one
    axis); points outside [src[0], src[-1]] get zero weights (NaN in xarray / scipy; does not happen on this path: with a
    symmetry plane the data is mirrored first, ``expand_symmetry``, and covers the plane)."""
    src, dst = np.asarray(src, float), np.asarray(dst, float)
    i1 = np.searchsorted(src, dst, side="left")
    i0 = np.clip(i1 - 1, 0, src.size - 1)
    i1 = np.clip(i1, 0, src.size - 1)
    w1 = np.where(src[i1] > src[i0], (dst - src[i0]) / np.where(src[i1] > src[i0], src[i1] - src[i0], 1.0), 0.0)
    exact = dst == src[i1]
    w0 = np.where(exact, 0.0, 1.0 - w1)
    w1 = np.where(exact, 1.0, w1)
    outside = (dst < src[0]) | (dst > src[-1])
    return i0, np.where(outside, 0.0, w0), i1, np.where(outside, 0.0, w1)


def symmetry_eigenvalue(name, axis):
    """components/data/dataset.py:210-220: E is a vector (the component along the mirrored axis flips), H a pseudovector."""
    k = "xyz".index(name[1])
    return (-1 if k == axis else +1) if name[0] == "E" else (+1 if k == axis else -1)


def expand_symmetry(f, sites, axis, name, sym_val, center):
    """monitor_data.py:237-282 (``_symmetry_update_dict``) along one axis: the half-domain data (sites >= center) is mirrored
    to the sites ``2 center - site`` on the other side of the symmetry plane and multiplied there by
    ``sym_val * symmetry_eigenvalue``.  A site exactly on the plane is not duplicated.  Returns (values, sites) of the
    full domain."""
    sites = np.asarray(sites, float)
    left = sites > center  # these have a mirror image strictly left of the plane
    mirrored = 2 * center - sites[left][::-1]
    vals = np.flip(np.compress(left, f, axis=axis), axis=axis) * (sym_val * symmetry_eigenvalue(name, axis))
    return np.concatenate([vals, f], axis=axis), np.concatenate([mirrored, sites])


def colocate(fields, coords, symmetry=(0, 0)):
    """mode_solver.py:490-507 (``_colocate_data``): every component of the SYMMETRY-EXPANDED data linearly interpolated from
    its Yee sites to the colocation points (with a symmetry plane the first of them is the plane itself, where a
    centre-site component is the mean of its first value and that value's mirror image: the value for an even component,
    zero for an odd one).  Returns dict name -> (Px, Py, M) arrays."""
    pts = colocation_points(coords, symmetry)
    out = {}
    for name, (kx, ky) in SITES.items():
        f = np.asarray(fields)[COMP[name][0], COMP[name][1], :, :, 0, :]
        for ax, (kind, p) in enumerate(zip((kx, ky), pts)):
            if p is None:
                continue
            src = _site_coords(coords[ax], kind)
            if symmetry[ax] != 0:
                f, src = expand_symmetry(f, src, ax, name, symmetry[ax], float(np.asarray(coords[ax])[0]))
            i0, w0, i1, w1 = interp_weights(src, p)
            shape = [1, 1, 1]
            shape[ax] = -1
            f = np.take(f, i0, axis=ax) * w0.reshape(shape) + np.take(f, i1, axis=ax) * w1.reshape(shape)
        out[name] = f
    return out


C_0 = 2.99792458e14  # um / s (tidy3d/constants.py:16)


def grid_correction(n_complex, freq, normal_primal, normal_dual, normal_pos, angle_theta=0.0, direction="+"):
    """mode_solver.py:847-904 (``ModeSolver._grid_correction``): the factors (primal[M], dual[M]) by which the tangential E /
    H fields of every mode are multiplied in flux and dot products (monitor_data.py:492-503): the mode is taken to propagate
    as exp(i k r) on the simulation grid along the normal and is linearly interpolated from the primal (cell boundaries:
    tangential E) and dual (cell centres: tangential H) grid points to the exact position of the mode plane."""
    n_complex = np.asarray(n_complex, complex)
    k_vec = 2 * np.pi * n_complex * freq / C_0 / np.cos(angle_theta)
    if direction == "-":
        k_vec = -k_vec
    out = []
    for pts in (np.atleast_1d(np.asarray(normal_primal, float)), np.atleast_1d(np.asarray(normal_dual, float))):
        phase = np.exp(1j * k_vec[:, None] * (pts[None, :] - normal_pos))  # (M, points)
        if pts.size > 1:  # DataArray.interp(normal_dim=normal_pos): linear
            out.append(np.array([np.interp(normal_pos, pts, ph.real) + 1j * np.interp(normal_pos, pts, ph.imag) for ph in phase]))
        else:  # .squeeze(dim=normal_dim)
            out.append(phase[:, 0])
    return out[0], out[1]


def diff_area(coords, symmetry=(0, 0), plane_bounds=None):
    """monitor_data.py:425-467 for data colocated at the points of ``colocation_points``: cell sizes from the mid-points
    between neighbouring points, closed with the first and last point (trapezoid weights); a one-cell axis has size 1.
    ``plane_bounds = (xmin, xmax, ymin, ymax)`` of a finite mode plane: the mid-points / end points are clipped to it
    (:450-455: "for pixels intersected by the monitor edge, the size is truncated to the part covered by the monitor");
    with a symmetry plane the bounds are those of the full, mirrored plane and only the max side can cut the half domain."""
    sizes = []
    for ax, p in enumerate(colocation_points(coords, symmetry)):
        if p is None or p.size == 1:
            sizes.append(np.array([1.0]))
            continue
        ctr = 0.5 * (p[1:] + p[:-1])
        ext = np.concatenate(([p[0]], ctr, [p[-1]]))
        if plane_bounds is not None:
            ext = np.clip(ext, plane_bounds[2 * ax], plane_bounds[2 * ax + 1])
        sizes.append(ext[1:] - ext[:-1])
    return np.outer(sizes[0], sizes[1])


def _corrected(c, correction):
    """monitor_data.py:488-503 (``_tangential_corrected``) in plane coordinates: tangential E (eigenvalue +1 under a
    reflection of the normal axis) times the primal factor, tangential H (eigenvalue -1) times the dual factor."""
    if correction is None:
        return c
    primal, dual = (np.asarray(q, complex) for q in correction)
    out = dict(c)
    for k in ("Ex", "Ey"):
        out[k] = c[k] * primal
    for k in ("Hx", "Hy"):
        out[k] = c[k] * dual
    return out


def flux(fields, coords, symmetry=(0, 0), correction=None, plane_bounds=None):
    """monitor_data.py:582-618: 0.5 Re(E1 H2* - E2 H1*) of the colocated tangential fields (times their grid-correction
    factors ``correction = (primal[M], dual[M])``, if any) integrated with ``diff_area``; a symmetry plane doubles the
    integral (symmetry_expanded mirrors the half domain)."""
    c = _corrected(colocate(fields, coords, symmetry), correction)
    s = 0.5 * np.real(c["Ex"] * np.conj(c["Hy"]) - c["Ey"] * np.conj(c["Hx"]))
    mult = 2 ** sum(1 for q in symmetry if q != 0)
    return mult * np.einsum("xym,xy->m", s, diff_area(coords, symmetry, plane_bounds))


def rotation_matrix(axis, angle):
    """components/transformation.py:112-131 (``RotationAroundAxis.matrix``), counter-clockwise by ``angle`` around ``axis``."""
    n = np.asarray(axis, float) / np.linalg.norm(axis)
    c, s = np.cos(angle), np.sin(angle)
    rot = np.zeros((3, 3))
    tan_dim = [[1, 2], [2, 0], [0, 1]]
    for dim in range(3):
        rot[dim, dim] = c + n[dim] ** 2 * (1 - c)
        rot[dim, tan_dim[dim][0]] = n[dim] * n[tan_dim[dim][0]] * (1 - c) - n[tan_dim[dim][1]] * s
        rot[dim, tan_dim[dim][1]] = n[dim] * n[tan_dim[dim][1]] * (1 - c) + n[tan_dim[dim][0]] * s
    return rot


def _expand_colocated(c, da, symmetry):
    """The colocated half-domain data ``c`` (name -> (Px, Py, M)) and its cell sizes ``da`` on the FULL, mirrored plane, as the
    reference's methods see them (``_colocated_fields`` works on ``symmetry_expanded`` data, monitor_data.py:527-542): the
    points right of a symmetry plane are mirrored with the factor ``sym_val * symmetry_eigenvalue``; the point ON the plane
    (the first one) is not duplicated and its cell is twice the half-domain one."""
    c = dict(c)
    for ax, s0 in enumerate(symmetry):
        if s0 == 0:
            continue
        for name in c:
            f = c[name]
            mirrored = np.flip(np.delete(f, 0, axis=ax), axis=ax) * (s0 * symmetry_eigenvalue(name, ax))
            c[name] = np.concatenate([mirrored, f], axis=ax)
        first = np.take(da, [0], axis=ax)
        da = np.concatenate([np.flip(np.delete(da, 0, axis=ax), axis=ax), 2 * first, np.delete(da, 0, axis=ax)], axis=ax)
    return c, da


def pol_fraction(fields, coords, symmetry=(0, 0), angle_theta=0.0, angle_phi=0.0, plane_bounds=None):
    """monitor_data.py:1625-1652: TE fraction = int |E1|^2 dS / int (|E1|^2 + |E2|^2) dS of the colocated field, E1 / E2 its
    first two components in the propagation axes (:1584-1614: [tangential 1, tangential 2, normal] rotated by -phi around the
    normal and then by -theta around the second axis).  The integrals run over the symmetry-EXPANDED plane: for an angled
    plane with a symmetry wall the rotation mixes components of opposite parity, whose products cancel between the two halves
    (pinned by the fuzzed comparison with the reference, tests/golden/post_angled_sym.npz) -- the half-domain integral times
    two is NOT the same number there."""
    c, da = _expand_colocated(colocate(fields, coords, symmetry), diff_area(coords, symmetry, plane_bounds), symmetry)
    field = np.array([c["Ex"], c["Ey"], c["Ez"]])
    if angle_phi != 0:
        field = np.tensordot(rotation_matrix([0, 0, 1], -angle_phi), field, axes=1)
    if angle_theta != 0:
        field = np.tensordot(rotation_matrix([0, 1, 0], -angle_theta), field, axes=1)
    te = np.einsum("xym,xy->m", np.abs(field[0]) ** 2, da)
    tm = np.einsum("xym,xy->m", np.abs(field[1]) ** 2, da)
    return te / (te + tm)


def normalize(fields, coords, symmetry=(0, 0), correction=None, plane_bounds=None):
    """mode_solver.py:517-521: all six components divided by sqrt(|flux|)."""
    fl = flux(fields, coords, symmetry, correction, plane_bounds)
    return np.asarray(fields) / np.sqrt(np.abs(fl)), fl


def dot(fields_a, fields_b, coords, symmetry=(0, 0), conjugate=True, correction_a=None, correction_b=None, plane_bounds=None):
    """monitor_data.py:640-697: M_a x M_b matrix  1/4 sum (E_a* x H_b + H_a* ... ) dS  of the colocated tangential fields
    (``outer_dot`` for all pairs), each data set with its own grid-correction factors."""
    a = _corrected(colocate(fields_a, coords, symmetry), correction_a)
    b = _corrected(colocate(fields_b, coords, symmetry), correction_b)
    if conjugate:
        a = {k: np.conj(v) for k, v in a.items()}
    da = diff_area(coords, symmetry, plane_bounds)
    mult = 2 ** sum(1 for q in symmetry if q != 0)

    def integ(x, y):
        return np.einsum("xym,xyn,xy->mn", x, y, da)

    e_x_h = integ(a["Ex"], b["Hy"]) - integ(a["Ey"], b["Hx"])
    h_x_e = integ(a["Hx"], b["Ey"]) - integ(a["Hy"], b["Ex"])
    return 0.25 * mult * (e_x_h - h_x_e)


def find_closest_pairs(arr):
    """monitor_data.py:1421-1440 (``_find_closest_pairs``): greedy pairing by largest |overlap|."""
    arr = np.asarray(arr)
    n = arr.shape[0]
    arr_abs = np.abs(arr).astype(float)
    pairs = -np.ones(n, dtype=int)
    values = np.zeros(n, dtype=complex)
    for _ in range(n):
        imax, jmax = np.unravel_index(np.argmax(arr_abs, axis=None), arr_abs.shape)
        pairs[imax] = jmax
        values[imax] = arr[imax, jmax]
        arr_abs[imax, :] = -1
        arr_abs[:, jmax] = -1
    return pairs, values


def find_ordering_one_freq(amps_matrix, overlap_thresh=0.9, direction="+"):
    """monitor_data.py:1378-1419 given the full M x M overlap matrix template -> to_sort: modes whose diagonal overlap is
    already above the threshold stay, the rest are re-paired."""
    amps_matrix = np.array(amps_matrix, dtype=complex)
    if direction == "-":
        amps_matrix = -amps_matrix
    m = amps_matrix.shape[0]
    pairs = np.arange(m)
    complex_amps = np.diag(amps_matrix).copy()
    to_sort = np.where(np.abs(complex_amps) < overlap_thresh)[0]
    if len(to_sort) <= 1:
        return pairs, complex_amps
    pr, vals = find_closest_pairs(amps_matrix[np.ix_(to_sort, to_sort)])
    complex_amps[to_sort] = vals
    pairs[to_sort] = to_sort[pr]
    return pairs, complex_amps


def overlap_sort(overlaps_next, num_freqs, num_modes, track_freq="central", overlap_thresh=0.9, direction="+"):
    """monitor_data.py:1295-1375.  ``overlaps_next[i]`` = dot(modes at f_i, modes at f_{i+1}) (M x M, already normalised
    fields), i < num_freqs - 1.  Returns (sorting[F,M], phase[F,M], overlap[F,M])."""
    f0 = {"lowest": 0, "highest": num_freqs - 1, "central": num_freqs // 2}[track_freq]
    sorting = -np.ones((num_freqs, num_modes), dtype=int)
    overlap = np.zeros((num_freqs, num_modes))
    phase = np.zeros((num_freqs, num_modes))
    sorting[f0] = np.arange(num_modes)
    overlap[f0] = 1.0
    for step, last in ((-1, -1), (1, num_freqs)):
        for fi in range(f0 + step, last, step):
            # template = previous frequency (fi - step), to_sort = fi;  dot(template, to_sort)
            if step == 1:
                mat = overlaps_next[fi - 1]
            else:  # dot(f_{i+1}, f_i) = conj(dot(f_i, f_{i+1}))^T term by term (monitor_data.py:680-697)
                mat = np.conj(overlaps_next[fi]).T
            one, amps = find_ordering_one_freq(mat, overlap_thresh, direction)
            sorting[fi] = one[sorting[fi - step]]
            overlap[fi] = np.abs(amps[sorting[fi - step]])
            phase[fi] = phase[fi - step] + np.angle(amps[sorting[fi - step]])
    return sorting, phase, overlap
