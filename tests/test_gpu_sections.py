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
