"""GPU: geometric cross-sections rasterised on the device (csrc/medium.cuh::section_raster_kernel) against
the numpy restatement of Simulation.epsilon_on_grid for boxes (oracle/sections.py)."""
import numpy as np
import pytest

from oracle import sections as OS
from tidy3d_b200 import compute_modes_batch
from tidy3d_b200 import workloads as W
from tidy3d_b200.sections import Medium, Rect, Section

pytestmark = pytest.mark.gpu


def test_section_rasterisation_equals_host_sampling():
    """Same raw permittivity -> the two paths must agree to the last bit (the solver is deterministic)."""
    n = 96
    c = np.linspace(-1.5, 1.5, n + 1)
    si = Medium(lambda f: 3.48**2 + 0.02 * (f / 1.934e14 - 1.0))            # dispersive core (callable)
    aniso = Medium([4.0, 4.2, 3.9])                                          # diagonal-anisotropic slab
    sec = Section(background=Medium(1.44**2), structures=[
        (Rect(center=(0.0, -0.4), size=(3.0, 0.3)), aniso),
        (Rect(center=(0.0, 0.0), size=(0.45, 0.22)), si),
        (Rect(center=(0.2, 0.05), size=(0.1, 0.1)), Medium(1.0)),           # a later structure overrides earlier ones
    ])
    spec = W.ModeSpecLike(num_modes=3, precision="double")
    freqs = W.sweep_freqs(4)
    a = compute_modes_batch([dict(section=sec, coords=[c, c], freq=f, mode_spec=spec) for f in freqs])
    b = compute_modes_batch([dict(eps_cross=OS.eps_on_grid(sec, [c, c], f), coords=[c, c], freq=f, mode_spec=spec) for f in freqs])
    for (fa, na, sa), (fb, nb, sb) in zip(a, b):
        assert sa == sb == "diagonal" and np.array_equal(na, nb) and np.array_equal(fa, fb)
    assert np.abs(a[0][1] - a[-1][1]).max() > 1e-3  # the sweep really changes the problem


def test_section_with_tensor_medium_goes_tensorial():
    n = 48
    c = np.linspace(-1.5, 1.5, n + 1)
    t = np.diag([12.1, 12.1, 12.1]).astype(complex)
    t[0, 1] = t[1, 0] = 0.6
    sec = Section(background=Medium(2.0736), structures=[(Rect((0, 0), (0.45, 0.22)), Medium(t))])
    spec = W.ModeSpecLike(num_modes=2, precision="double")
    f = W.C_0 / 1.55
    (fa, na, sa), = compute_modes_batch([dict(section=sec, coords=[c, c], freq=f, mode_spec=spec)])
    (fb, nb, sb), = compute_modes_batch([dict(eps_cross=OS.eps_on_grid(sec, [c, c], f), coords=[c, c], freq=f, mode_spec=spec)])
    assert sa == sb == "tensorial_real" and np.array_equal(na, nb)


def shapes_section(n=96):
    """A rib-like cross-section that needs every way of describing geometry (b200ms_section v201): the slab comes as a site
    map (inside-masks evaluated by the caller), the rib is a slanted-wall trapezoid, a fibre-like disc sits next to it and a
    box punches a hole into the disc.  Also used by the CPU test of the host mirror (tests/test_host_logic.py)."""
    from tidy3d_b200.sections import Disc, Polygon, site_medium_from_masks

    c = np.linspace(-1.5, 1.5, n + 1)
    xc = (c[:-1] + c[1:]) / 2
    bg, slab = Medium(1.44**2), Medium([4.0, 4.2, 3.9])
    si = Medium(lambda f: 3.48**2 + 0.02 * (f / 1.934e14 - 1.0))
    slab_rect = Rect(center=(0.0, -0.4), size=(3.0, 0.3))
    masks = np.stack([OS.inside(slab_rect, sx, sy) for sx, sy in [(xc, c[:-1]), (c[:-1], xc), (c[:-1], c[:-1])]])
    sec = Section(background=bg, media=[bg, slab], site_medium=site_medium_from_masks((n, n), [(masks, 1)]), structures=[
        (Polygon([(-0.31, -0.252), (0.31, -0.252), (0.23, -0.03), (-0.23, -0.03)]), si),
        (Disc(center=(0.8, 0.3), radius=0.27), Medium(2.0**2 + 1e-4j)),  # a lossy nitride rod: complex arithmetic, no twin-core degeneracy
        (Rect(center=(0.8, 0.3), size=(0.11, 0.11)), bg),
    ])
    return sec, c


def test_section_shapes_and_site_map_equal_host_sampling():
    """Discs, polygons and a caller-made site map rasterised on the device: same raw permittivity as the numpy restatement
    of epsilon_on_grid -> the two paths agree to the last bit, at every frequency of a sweep."""
    sec, c = shapes_section()
    spec = W.ModeSpecLike(num_modes=3, precision="double")
    freqs = W.sweep_freqs(3)
    a = compute_modes_batch([dict(section=sec, coords=[c, c], freq=f, mode_spec=spec) for f in freqs])
    b = compute_modes_batch([dict(eps_cross=OS.eps_on_grid(sec, [c, c], f), coords=[c, c], freq=f, mode_spec=spec) for f in freqs])
    for (fa, na, sa), (fb, nb, sb) in zip(a, b):
        assert sa == sb == "diagonal" and np.array_equal(na, nb) and np.array_equal(fa, fb)
    assert na[0].real > 2.0  # a guided mode of the silicon rib, not of the cladding
