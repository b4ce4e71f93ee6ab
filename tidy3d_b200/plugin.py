"""Seams into a real ``tidy3d`` install (not importable in this image: xarray/shapely/... are absent).

Seam 1 (exact drop-in, one solve per call): ``tidy3d.plugins.mode.mode_solver`` imports the module-level name
``compute_modes`` (mode_solver.py:59-65) and calls it at :725; ``install()`` rebinds that name.

Seam 2 (batched, where the speed is): ``ModeSolver._solve_all_freqs(self, coords, symmetry)`` (mode_solver.py:655-672)
loops ``_solve_single_freq`` over ``self.freqs``; ``solve_all_freqs_batched`` gathers ``self._solver_eps(f)`` for every
frequency, runs ONE device call and reuses the reference's own ``_postprocess_solver_fields`` (:695).  ``run_batch``
mirrors ``tidy3d.web.api.mode.run_batch(mode_solvers)`` (web/api/mode.py:147-158) for many mode planes.
"""
from __future__ import annotations

from typing import List

from .solver import compute_modes, compute_modes_batch


def solve_all_freqs_batched(self, coords, symmetry):
    """Replacement body for ``ModeSolver._solve_all_freqs`` (same signature and return value)."""
    problems = [
        dict(eps_cross=self._solver_eps(freq), coords=coords, freq=freq, mode_spec=self.mode_spec, symmetry=symmetry,
             direction=self.direction)
        for freq in self.freqs
    ]  # fmt: skip
    n_complex, fields, eps_spec = [], [], []
    for solver_fields, n_freq, spec in compute_modes_batch(problems):
        fields.append(self._postprocess_solver_fields(solver_fields))
        n_complex.append(n_freq)
        eps_spec.append(spec)
    return n_complex, fields, eps_spec


def install(batched: bool = True):
    """Route ``tidy3d.plugins.mode.ModeSolver`` through the B200 library.  Raises ImportError without tidy3d."""
    import tidy3d.plugins.mode.mode_solver as ms  # noqa: PLC0415

    ms.compute_modes = compute_modes
    ms.LOCAL_SOLVER_IMPORTED = True
    if batched:
        ms.ModeSolver._solve_all_freqs = solve_all_freqs_batched
    return ms.ModeSolver


def run_batch(mode_solvers: List, **kwargs) -> List:
    """Local analogue of ``tidy3d.web.api.mode.run_batch``: ``[ms.data for ms in mode_solvers]`` with the solver
    routed through the GPU (each ``ModeSolver.data`` triggers one batched device call over its frequencies)."""
    install(batched=True)
    return [ms.data for ms in mode_solvers]
