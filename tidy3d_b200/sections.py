"""Geometric cross-sections rasterised on the GPU.

``ModeSolver._solver_eps(freq)`` samples the permittivity tensor of the simulation on the solver plane with nine
``Simulation.epsilon_on_grid`` calls per frequency (mode_solver.py:587-653, simulation.py:1135-1241: background value,
then every structure in order overwrites the Yee sites it contains) and hands a fresh (9,Nx,Ny) complex array to the
solver for every frequency.  For cross-sections made of axis-aligned rectangles (the cut of ``Box`` structures) that work
-- and the 144 N bytes of host-to-device traffic per frequency it implies -- is replaced by a list of rectangles plus 9
numbers per medium and frequency; ``csrc/medium.cuh::section_raster_kernel`` writes the same array straight into HBM.

    sec = Section(background=Medium(1.44**2), structures=[(Rect(center=(0, 0), size=(0.45, 0.22)), Medium(3.48**2))])
    compute_modes_batch([dict(section=sec, coords=coords, freq=f, mode_spec=spec) for f in freqs])
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Sequence, Tuple, Union

import numpy as np

from . import _cabi


@dataclass
class Rect:
    """Cut of a ``Box`` with the solver plane: ``center`` and ``size`` in plane coordinates (geometry/base.py:1799)."""

    center: Tuple[float, float]
    size: Tuple[float, float]


EpsLike = Union[complex, Sequence[complex], np.ndarray]


@dataclass
class Medium:
    """Relative permittivity of a medium: a scalar (isotropic), three numbers (diagonal, ``AnisotropicMedium``), a 3x3
    tensor (``FullyAnisotropicMedium``), or a callable ``freq -> any of those`` for dispersive media (the reference
    evaluates ``structure.eps_comp(row, col, frequency)``, simulation.py:1176-1185)."""

    eps: Union[EpsLike, Callable[[float], EpsLike]]

    def tensor(self, freq: float) -> np.ndarray:
        e = self.eps(freq) if callable(self.eps) else self.eps
        e = np.asarray(e, dtype=complex)
        if e.ndim == 0:
            return np.diag([e, e, e]).astype(complex)
        if e.shape == (3,):
            return np.diag(e).astype(complex)
        if e.shape == (3, 3):
            return e.astype(complex)
        raise ValueError("Medium.eps must be a scalar, 3 numbers or a 3x3 tensor")


@dataclass
class Section:
    background: Medium
    structures: List[Tuple[Rect, Medium]] = field(default_factory=list)

    def pack(self, freq: float):
        """ctypes ``b200ms_section`` for one frequency plus the arrays it points to (kept alive by the caller)."""
        media, ids = [self.background], []
        for _, m in self.structures:
            for k, known in enumerate(media):
                if known is m:
                    ids.append(k)
                    break
            else:
                media.append(m)
                ids.append(len(media) - 1)
        rects = np.ascontiguousarray([[r.center[0], r.center[1], r.size[0], r.size[1]] for r, _ in self.structures], dtype=np.float64).reshape(-1, 4)
        med = np.ascontiguousarray(ids, dtype=np.int32)
        table = np.ascontiguousarray([m.tensor(freq).ravel() for m in media], dtype=np.complex128)
        st = _cabi.SectionStruct()
        st.nrect = len(self.structures)
        st.rects = rects.ctypes.data_as(_cabi._dp)
        st.medium = med.ctypes.data_as(_cabi._ip)
        st.nmedia = len(media)
        st.eps_table = table.view(np.float64).ctypes.data_as(_cabi._dp)
        return st, (rects, med, table)
