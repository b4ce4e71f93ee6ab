"""Host-side half of the post-processing seam.

The heavy per-cell work -- gauge, colocation, flux normalisation and the modal overlap matrices between adjacent
frequencies -- runs on the GPU while the fields are still in HBM (``compute_modes_batch(..., post=(...))``,
``csrc/post.cuh``).  What is left for the host is the bookkeeping of ``ModeData.overlap_sort``
(tidy3d/components/data/monitor_data.py:1295-1505) on F x M x M numbers: which mode of frequency i+1 continues which mode
of frequency i, and the phase that makes the continuation smooth.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def _closest_pairs(amps: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Greedy pairing of rows and columns by decreasing |overlap| (``_find_closest_pairs``, monitor_data.py:1421-1440)."""
    n = amps.shape[0]
    mag = np.abs(amps).astype(float)
    pairs = np.full(n, -1, dtype=int)
    values = np.zeros(n, dtype=complex)
    for _ in range(n):
        i, j = divmod(int(np.argmax(mag)), n)
        pairs[i], values[i] = j, amps[i, j]
        mag[i, :] = -1.0
        mag[:, j] = -1.0
    return pairs, values


def _ordering_one_step(amps: np.ndarray, thresh: float) -> Tuple[np.ndarray, np.ndarray]:
    """``_find_ordering_one_freq`` (monitor_data.py:1378-1419): modes already matching their own index stay put."""
    m = amps.shape[0]
    pairs = np.arange(m)
    diag = np.diag(amps).astype(complex).copy()
    loose = np.flatnonzero(np.abs(diag) < thresh)
    if loose.size > 1:
        sub_pairs, sub_vals = _closest_pairs(amps[np.ix_(loose, loose)])
        pairs[loose] = loose[sub_pairs]
        diag[loose] = sub_vals
    return pairs, diag


def overlap_sort(overlap_prev: Sequence[np.ndarray], track_freq: str = "central", overlap_thresh: float = 0.9, direction: str = "+"):
    """Mode tracking across a frequency sweep from the device-computed overlap matrices.

    ``overlap_prev[i]`` (i >= 1) = dot(modes at f_{i-1}, modes at f_i), what ``compute_modes_batch(post=("normalize",
    "overlaps"))`` returns per problem as ``info["overlap_prev"]`` (entry 0 is ignored).  Returns ``(sorting, phase,
    overlap)`` of shape (F, M) with the meaning of the reference's arrays: mode ``k`` of the sorted data at frequency i is
    mode ``sorting[i, k]`` of the unsorted data multiplied by ``exp(-1j * phase[i, k])`` (monitor_data.py:1442-1478).
    """
    nf = len(overlap_prev)
    m = np.asarray(overlap_prev[-1]).shape[0]
    base = {"lowest": 0, "highest": nf - 1, "central": nf // 2}[track_freq]
    sign = -1.0 if direction == "-" else 1.0  # store_fields_direction == "-" flips the overlaps (monitor_data.py:1391-1392)
    sorting = np.full((nf, m), -1, dtype=int)
    overlap = np.zeros((nf, m))
    phase = np.zeros((nf, m))
    sorting[base] = np.arange(m)
    overlap[base] = 1.0
    for step in (-1, 1):
        i = base + step
        while 0 <= i < nf:
            prev = i - step
            # dot(template = prev, to_sort = i).  Marching down in frequency needs dot(f_{i+1}, f_i) = conj(dot(f_i, f_{i+1}))^T
            amps = sign * (np.asarray(overlap_prev[i]) if step == 1 else np.conj(np.asarray(overlap_prev[i + 1])).T)
            one, vals = _ordering_one_step(amps, overlap_thresh)
            sorting[i] = one[sorting[prev]]
            overlap[i] = np.abs(vals[sorting[prev]])
            phase[i] = phase[prev] + np.angle(vals[sorting[prev]])
            i += step
    return sorting, phase, overlap


def apply_sorting(n_complex: Sequence[np.ndarray], fields: Sequence[np.ndarray], sorting: np.ndarray, phase: np.ndarray):
    """``_reorder_modes`` (monitor_data.py:1442-1478) on per-frequency arrays: returns re-ordered copies of ``n_complex``
    (list of (M,)) and ``fields`` (list of (2,3,Nx,Ny,1,M) or None entries)."""
    n_out: List[np.ndarray] = []
    f_out: List = []
    for i, (n, f) in enumerate(zip(n_complex, fields)):
        n_out.append(np.asarray(n)[sorting[i]])
        if f is None:
            f_out.append(None)
        else:
            f_out.append((np.asarray(f)[..., sorting[i]] * np.exp(-1j * phase[i])).astype(np.asarray(f).dtype))
    return n_out, f_out


def filter_polarization(te_fraction: np.ndarray, filter_pol: str) -> np.ndarray:
    """Mode order of ``ModeSolver._filter_polarization`` (mode_solver.py:523-549) for one frequency from the device-computed
    TE fractions (``info["te_fraction"]``): modes of the requested polarisation first, the others after, NaN last."""
    te = np.asarray(te_fraction, float)
    if filter_pol == "te":
        parts = (np.where(te >= 0.5)[0], np.where(te < 0.5)[0], np.where(np.isnan(te))[0])
    elif filter_pol == "tm":
        parts = (np.where(te <= 0.5)[0], np.where(te > 0.5)[0], np.where(np.isnan(te))[0])
    else:
        raise ValueError("filter_pol must be 'te' or 'tm'")
    return np.concatenate(parts)


def grid_correction_table(normal_primal, normal_dual, normal_pos: float) -> np.ndarray:
    """The 8 numbers of ``b200ms_problem.grid_correction`` for a mode plane at ``normal_pos`` in a simulation whose grid
    along the plane normal has the boundaries ``normal_primal`` and the centres ``normal_dual``
    (``simulation.grid.boundaries / centers`` along ``normal_axis``, mode_solver.py:879-883): for each of the two grids the
    offsets from the plane to the two grid points that bracket it and the weights of the linear interpolation the reference
    does with ``DataArray.interp`` (:893-900); a grid of one point contributes that point with weight 1 (``squeeze``).
    The library forms ``primal = w0 exp(i k d0) + w1 exp(i k d1)`` (and ``dual``) per mode once ``n_complex`` is known, and uses
    them in the flux normalisation and the modal overlaps (monitor_data.py:488-503) -- what ``ModeSolverData`` stores as
    ``grid_primal_correction`` / ``grid_dual_correction``."""
    out = []
    for pts in (normal_primal, normal_dual):
        pts = np.atleast_1d(np.asarray(pts, dtype=float))
        if pts.size == 1:
            out += [pts[0] - normal_pos, 1.0, 0.0, 0.0]
            continue
        if not (pts[0] <= normal_pos <= pts[-1]):
            raise ValueError("the mode plane lies outside the simulation grid along its normal")
        i = int(np.clip(np.searchsorted(pts, normal_pos, side="right") - 1, 0, pts.size - 2))
        w1 = (normal_pos - pts[i]) / (pts[i + 1] - pts[i])
        out += [pts[i] - normal_pos, 1.0 - w1, pts[i + 1] - normal_pos, w1]
    return np.array(out, dtype=float)


def grid_correction_factors(n_complex, freq: float, table, angle_theta: float = 0.0, direction: str = "+"):
    """``(primal[M], dual[M])`` for given ``n_complex`` (what the reference stores in ``ModeSolverData.grid_primal_correction`` /
    ``grid_dual_correction``, mode_solver.py:884-904), from the table of ``grid_correction_table``: cheap host arithmetic on
    M numbers for callers that assemble ``ModeSolverData`` themselves."""
    t = np.asarray(table, float)
    k = 2 * np.pi * np.asarray(n_complex, complex) * freq / 2.99792458e14 / np.cos(angle_theta) * (-1.0 if direction == "-" else 1.0)
    primal = t[1] * np.exp(1j * k * t[0]) + t[3] * np.exp(1j * k * t[2])
    dual = t[5] * np.exp(1j * k * t[4]) + t[7] * np.exp(1j * k * t[6])
    return primal, dual


# Yee sites of the six components in the solver plane (components/grid/grid.py Grid.yee): c = cell centre, b = lower boundary
_SITES = (("c", "b"), ("b", "c"), ("b", "b"), ("b", "c"), ("c", "b"), ("c", "c"))  # Ex Ey Ez Hx Hy Hz


def colocate(fields: np.ndarray, coords, symmetry=(0, 0)):
    """``ModeSolver._colocate_data`` (mode_solver.py:490-515) for delivered Yee-grid fields: every component linearly
    interpolated to the interior cell boundaries (plus the symmetry plane itself where the plane has a symmetry wall, the
    half-domain data being mirrored first as ``symmetry_expanded`` does, monitor_data.py:237-282) -- the format of
    ``ModeSolver(colocate=True)``, the reference's default.  ``fields``: (2,3,Nx,Ny,1,M) as ``compute_modes`` returns them
    (after ``post=("gauge", "normalize")`` they are the normalised modes, and colocating commutes with both).  Returns
    ``(colocated (2,3,Px,Py,1,M), (x_points, y_points))``.

    The interpolation tables are the library's own (the ones its flux / overlap kernels use on the device, host code of
    ``csrc/api.cu``), so the colocated fields integrate to the flux the device reports."""
    from . import _cabi

    f = np.asarray(fields)
    tabs, pts = [], []
    for c, s in zip(coords, symmetry):
        c = np.ascontiguousarray(c, dtype=np.float64)
        n = c.size - 1
        idx, wgt, area = np.zeros(4 * (n + 1), np.int32), np.zeros(4 * (n + 1)), np.zeros(n + 1)
        npt = _cabi.lib().b200ms_debug_post_tables(_cabi._ptr(c), n, int(s), n + 1, idx.ctypes.data_as(_cabi._ip), _cabi._ptr(wgt), _cabi._ptr(area))
        if npt < 1:
            raise ValueError("bad coordinates")
        tabs.append((idx[: 4 * npt].reshape(npt, 4), wgt[: 4 * npt].reshape(npt, 4)))
        pts.append((c[1:-1] if s == 0 else c[:-1]) if n > 1 else 0.5 * (c[:1] + c[1:]))
    out = np.empty(f.shape[:2] + (tabs[0][0].shape[0], tabs[1][0].shape[0]) + f.shape[4:], dtype=f.dtype)
    for k, kinds in enumerate(_SITES):
        g = f[k // 3, k % 3]
        for ax, kind in enumerate(kinds):
            idx, wgt = tabs[ax]
            o = 0 if kind == "c" else 2
            shape = [1] * g.ndim
            shape[ax] = -1
            g = np.take(g, idx[:, o], axis=ax) * wgt[:, o].reshape(shape) + np.take(g, idx[:, o + 1], axis=ax) * wgt[:, o + 1].reshape(shape)
        out[k // 3, k % 3] = g
    return out, tuple(pts)
