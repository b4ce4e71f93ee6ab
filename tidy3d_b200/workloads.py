"""Deterministic synthetic cross-sections for the BASELINE.json configs.

These are *inputs* only (eps_cross, coords, freq, mode_spec) in exactly the form the reference's
``compute_modes`` consumes (``tidy3d/plugins/mode/solver.py:33-44``): nine ``(Nx, Ny)`` complex
arrays in the order xx,xy,xz,yx,yy,yz,zx,zy,zz, two coordinate arrays of length Nx+1 / Ny+1 and a
duck-typed mode spec.  Shared by the tests, ``bench.py`` and ``__graft_entry__.smoke``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

C_0 = 2.99792458e14  # um/s, tidy3d/constants.py:16


@dataclass
class ModeSpecLike:
    """Plain-attribute stand-in for ``tidy3d.components.mode.ModeSpec`` (mode.py:18-209).

    ``compute_modes`` only reads attributes (solver.py:86-90,198,204,247,265) so any object with
    these names works, including the real pydantic ``ModeSpec``.
    """

    num_modes: int = 1
    target_neff: Optional[float] = None
    num_pml: Tuple[int, int] = (0, 0)
    filter_pol: Optional[str] = None
    angle_theta: float = 0.0
    angle_phi: float = 0.0
    precision: str = "double"
    bend_radius: Optional[float] = None
    bend_axis: Optional[int] = None
    track_freq: Optional[str] = "central"
    group_index_step: float = 0.0


@dataclass
class Workload:
    name: str
    eps_cross: List[np.ndarray]  # 9 x (Nx, Ny) complex
    coords: List[np.ndarray]  # [x (Nx+1), y (Ny+1)]
    freqs: np.ndarray
    mode_spec: ModeSpecLike
    symmetry: Tuple[int, int] = (0, 0)
    direction: str = "+"
    note: str = ""
    extra: dict = field(default_factory=dict)


def _iso(eps2d):
    z = np.zeros_like(eps2d)
    return [eps2d.copy(), z.copy(), z.copy(), z.copy(), eps2d.copy(), z.copy(), z.copy(), z.copy(), eps2d.copy()]


def _grid(n, length):
    c = np.linspace(-length / 2, length / 2, n + 1)
    ctr = 0.5 * (c[:-1] + c[1:])
    return c, ctr


def strip_eps(nx, ny, lx=3.0, ly=3.0, w=0.45, h=0.22, n_core=3.48, n_clad=1.44):
    """Cell-centre staircased rectangular core, SURVEY 8(d) C1."""
    xc, xm = _grid(nx, lx)
    yc, ym = _grid(ny, ly)
    core = (np.abs(xm)[:, None] <= w / 2) & (np.abs(ym)[None, :] <= h / 2)
    eps = np.where(core, n_core**2, n_clad**2).astype(complex)
    return eps, [xc, yc]


def si_strip(n=64, num_modes=2, freqs=None, lam=1.55, **kw) -> Workload:
    """C1 / C2 / headline: Si strip 0.45x0.22 um in n=1.44, 3x3 um, no PML, target None."""
    eps, coords = strip_eps(n, n, **kw)
    if freqs is None:
        freqs = np.array([C_0 / lam])
    return Workload(
        name=f"si_strip_{n}x{n}_m{num_modes}_f{len(freqs)}",
        eps_cross=_iso(eps),
        coords=coords,
        freqs=np.asarray(freqs, float),
        mode_spec=ModeSpecLike(num_modes=num_modes, precision="double"),
    )


def sweep_freqs(nf=256, lam0=1.5, lam1=1.6):
    return C_0 / np.linspace(lam0, lam1, nf)


def c1() -> Workload:
    return si_strip(64, 2)


def c2(nf=256, n=256) -> Workload:
    return si_strip(n, 4, sweep_freqs(nf))


def headline(nf=256, n=512) -> Workload:
    w = si_strip(n, 4, sweep_freqs(nf))
    w.name = f"headline_{n}x{n}_m4_f{nf}"
    return w


def c3(n=512, num_modes=6, lam=1.55) -> Workload:
    """SiN rib on SiO2, diagonal-anisotropic eps + 12-cell PML, target 1.9."""
    length = 6.0
    xc, xm = _grid(n, length)
    yc, ym = _grid(n, length)
    rib_w, rib_h, slab_h = 1.2, 0.4, 0.2
    X, Y = np.meshgrid(xm, ym, indexing="ij")
    box = Y < 0.0
    slab = (Y >= 0.0) & (Y < slab_h)
    rib = (Y >= slab_h) & (Y < slab_h + rib_h) & (np.abs(X) <= rib_w / 2)
    core = slab | rib
    comps = []
    for val in (4.0, 4.2, 3.9):
        e = np.where(core, val, np.where(box, 2.0736, 1.0)).astype(complex)
        comps.append(e)
    z = np.zeros((n, n), complex)
    eps9 = [comps[0], z.copy(), z.copy(), z.copy(), comps[1], z.copy(), z.copy(), z.copy(), comps[2]]
    return Workload(
        name=f"sin_rib_aniso_pml_{n}x{n}_m{num_modes}",
        eps_cross=eps9,
        coords=[xc, yc],
        freqs=np.array([C_0 / lam]),
        mode_spec=ModeSpecLike(num_modes=num_modes, num_pml=(12, 12), target_neff=1.9),
    )


def c4(n=512, num_modes=4, lam=1.55, bend_radius=5.0) -> Workload:
    """Bent SOI strip, bend_radius 5 um, 12-cell PML, target 2.4 (complex n_eff)."""
    eps, coords = strip_eps(n, n, lx=4.0, ly=4.0)
    return Workload(
        name=f"soi_bend_R{bend_radius:g}_{n}x{n}_m{num_modes}",
        eps_cross=_iso(eps),
        coords=coords,
        freqs=np.array([C_0 / lam]),
        mode_spec=ModeSpecLike(
            num_modes=num_modes, num_pml=(12, 12), target_neff=2.4, bend_radius=bend_radius, bend_axis=1
        ),
    )


def c5_planes(n_planes=32, n=256, nf=128) -> List[Workload]:
    """C5: 32 mode planes = C2-type cross-sections with core width swept 0.40..0.71 um."""
    out = []
    for w in np.linspace(0.40, 0.71, n_planes):
        wl = si_strip(n, 4, sweep_freqs(nf), w=float(w))
        wl.name = f"plane_w{w:.3f}_{n}x{n}"
        out.append(wl)
    return out


def angled(n=96, theta=0.2, phi=0.0, num_modes=4, lam=1.55) -> Workload:
    """Angled strip (tensorial_real path, solver.py:594); SURVEY Appendix B row 5."""
    eps, coords = strip_eps(n, n)
    return Workload(
        name=f"si_strip_angled_{n}",
        eps_cross=_iso(eps),
        coords=coords,
        freqs=np.array([C_0 / lam]),
        mode_spec=ModeSpecLike(num_modes=num_modes, angle_theta=theta, angle_phi=phi),
    )
