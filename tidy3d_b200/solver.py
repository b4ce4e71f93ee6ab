"""Host-side mirror of the reference mode-solver interface, backed by the CUDA library.

``compute_modes`` has the signature, return value and error behaviour of
``tidy3d.plugins.mode.solver.compute_modes`` (tidy3d/plugins/mode/solver.py:33-44, 941-943):

    fields, n_complex, eps_spec = compute_modes(eps_cross, coords, freq, mode_spec, symmetry=..., direction=...)

``fields`` has shape ``(2, 3, Nx, Ny, 1, num_modes)`` (E/H, component, x, y, 1, mode), ``n_complex`` is
``n_eff + 1j*k_eff`` sorted by descending ``n_eff`` and ``eps_spec`` is ``"diagonal"`` (tensorial cross-sections
raise ``NotImplementedError`` for now).  ``compute_modes_batch`` is the batched entry point that corresponds to the
frequency loop of ``ModeSolver._solve_all_freqs`` (tidy3d/plugins/mode/mode_solver.py:655-672): all problems are
solved concurrently on the GPU.
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi

_handles = {}
_hlock = threading.Lock()


def get_handle(device: int = -1) -> _cabi.Handle:
    with _hlock:
        if device not in _handles:
            _handles[device] = _cabi.Handle(device)
        return _handles[device]


def _raise_for(rc: int, handle, where=""):
    msg = handle.last_error()
    if rc == _cabi.ERR_SHAPE:
        raise ValueError("Mismatch between 'coords' and 'esp_cross' shapes.")  # solver.py:107
    if rc == _cabi.ERR_NO_MODES:
        raise RuntimeError("Could not find any eigenmodes for this waveguide.")  # solver.py:876
    if rc == _cabi.ERR_UNSUPPORTED:
        raise NotImplementedError(f"tidy3d_b200: {msg}")
    if rc == _cabi.ERR_NOCONV:
        raise RuntimeError(f"tidy3d_b200: eigen-iteration did not converge {where}: {msg}")  # ArpackNoConvergence
    if rc == _cabi.ERR_CUDA:
        raise RuntimeError(f"tidy3d_b200: CUDA failure (no CPU fallback): {msg}")
    if rc != _cabi.OK:
        raise RuntimeError(f"tidy3d_b200: error {rc}: {msg}")


def _check_unsupported(mu_cross, split_curl_scaling, solver_basis_fields):
    if mu_cross is not None or split_curl_scaling is not None:
        raise NotImplementedError("tidy3d_b200: mu_cross / split_curl_scaling (solver.py:93) are not built yet")


def compute_modes_batch(
    problems: Sequence[dict], device: int = -1, want_fields: bool = True, return_info: bool = False, handle=None
):
    """Solve many independent mode problems in one device call.

    Each problem is a dict with the keyword arguments of ``compute_modes`` (``eps_cross, coords, freq, mode_spec``
    and optionally ``symmetry, direction``).  Problems that share the same ``eps_cross`` object are packed once.
    Returns a list of ``(fields, n_complex, eps_spec)`` tuples (``fields`` is None when ``want_fields=False``),
    plus a list of per-problem info dicts when ``return_info``.
    """
    h = handle or get_handle(device)
    packed, cache = [], {}
    for p in problems:
        _check_unsupported(p.get("mu_cross"), p.get("split_curl_scaling"), p.get("solver_basis_fields"))
        key = id(p["eps_cross"])
        pk = _cabi.PackedProblem(
            p["eps_cross"], p["coords"], p["freq"], p["mode_spec"], p.get("symmetry", (0, 0)), p.get("direction", "+"),
            eps_packed=cache.get(key), basis_fields=p.get("solver_basis_fields"),
        )  # fmt: skip
        cache[key] = pk.eps
        if key in cache and len(packed) and packed[-1].eps is pk.eps:
            # share the coordinate arrays too so the library can detect identical cross-sections
            prev = packed[-1]
            if np.array_equal(prev.cx, pk.cx) and np.array_equal(prev.cy, pk.cy):
                pk.cx, pk.cy = prev.cx, prev.cy
                pk.struct.coords_x, pk.struct.coords_y = prev.struct.coords_x, prev.struct.coords_y
        packed.append(pk)
    rc, fields, ncs, results = h.solve_batch(packed, want_fields)
    if rc != _cabi.OK:
        bad = [i for i in range(len(packed)) if results[i].status != _cabi.OK]
        _raise_for(rc, h, f"(problems {bad[:8]})")
    out, infos = [], []
    for i, pk in enumerate(packed):
        f = fields[i] if want_fields else None
        if f is not None and pk.struct.precision == 1:
            f = f.astype(np.complex64)  # solver.py:265-267
        out.append((f, ncs[i], _cabi.SPEC_NAMES[results[i].eps_spec]))
        r = results[i]
        infos.append(
            dict(converged=r.converged, restarts=r.outer_iters, op_applies=r.op_applies, inner_iters=r.inner_iters,
                 stencil_applies=r.stencil_applies, is_complex=bool(r.is_complex), solve_ms=r.solve_ms, total_ms=r.total_ms,
                 max_residual=r.max_residual)
        )  # fmt: skip
    return (out, infos) if return_info else out


def compute_modes(
    eps_cross,
    coords,
    freq,
    mode_spec,
    mu_cross=None,
    split_curl_scaling=None,
    symmetry=(0, 0),
    direction="+",
    solver_basis_fields=None,
) -> Tuple[np.ndarray, np.ndarray, str]:
    """Drop-in for ``tidy3d.plugins.mode.solver.compute_modes`` (solver.py:941)."""
    _check_unsupported(mu_cross, split_curl_scaling, solver_basis_fields)
    return compute_modes_batch(
        [dict(eps_cross=eps_cross, coords=coords, freq=freq, mode_spec=mode_spec, symmetry=symmetry, direction=direction,
              solver_basis_fields=solver_basis_fields)]
    )[0]
