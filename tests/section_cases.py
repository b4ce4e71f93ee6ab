"""The 3-D scene of the f-2 pinning tests as plain data: consumed by the literal stand-in of
tests/test_plugin_seams.py, by oracle/ref_sections.py (the reference's own sampling code, build container) and by the fixture
generator tests/golden/make_sections_golden.py."""
import numpy as np

FULL = np.array([[4.0, 0.3, 0.1], [0.3, 4.4, 0.2j], [0.1, -0.2j, 3.7]])
SI = np.diag([12.0, 12.1, 12.2])
BACKGROUND = (2.1 * np.eye(3), 0.0)  # (tensor, dispersion slope)
# (kind, geometry parameters, tensor, slope); later structures override earlier ones (simulation.py:1191-1226)
STRUCTURES = [
    ("Box", dict(center=(0.05, -0.1, 0.0), size=(0.9, 0.5, 0.7)), FULL, -0.02),
    ("Sphere", dict(center=(0.2, 0.1, 0.05), radius=0.33), SI, 0.05),
    ("Box", dict(center=(-0.3, 0.2, -0.2), size=(0.2, 0.3, 0.25)), SI, 0.05),  # equal medium, another object
    ("Cylinder", dict(center=(-0.35, -0.3, 0.1), radius=0.21, length=0.5, axis=1), 1.3 * FULL, 0.01),
    # edges on grid lines: the inclusive `<=` of Box.inside decides (base.py:2070)
    ("Box", dict(center=(0.5, 0.35, 0.3), size=(0.4, 0.3, 0.6)), np.diag([6.0, 6.5, 7.0]), 0.0),
]
BOUNDS = [(-0.8, 0.9), (-0.7, 0.7), (-0.6, 0.75)]
CELLS = (17, 14, 15)
FREQS = (1.9e14, 2.1e14)


def edges(normal_axis):
    """Cell boundaries of the solver grid: one cell along the plane normal."""
    e = [np.linspace(lo, hi, k + 1) for (lo, hi), k in zip(BOUNDS, CELLS)]
    e[normal_axis] = np.array([-0.01, 0.01])
    return e


# primitive cuts of oracle/sections.py / tidy3d_b200/sections.py (plane z = 0) next to the reference geometry they stand for;
# sites chosen so that some fall exactly on an edge / on the circle (inclusive comparisons, base.py:2070, primitives.py:70, 631)
SX = np.round(np.linspace(-1.0, 1.0, 41), 12)
SY = np.round(np.linspace(-0.6, 0.9, 31), 12)
PRIMITIVES = {
    # name: (reference kind, reference parameters, plane cut (kind, parameters))
    "rect": ("Box", dict(center=(0.1, 0.15, 0.0), size=(0.8, 0.5, 1.0)), ("Rect", dict(center=(0.1, 0.15), size=(0.8, 0.5)))),
    "sphere": ("Sphere", dict(center=(-0.2, 0.1, 0.3), radius=0.5), ("Disc", dict(center=(-0.2, 0.1), radius=0.5, dz=0.3))),
    "sphere_touching": ("Sphere", dict(center=(0.0, 0.0, 0.0), radius=0.5), ("Disc", dict(center=(0.0, 0.0), radius=0.5, dz=0.0))),
    "cylinder": ("Cylinder", dict(center=(0.3, -0.1, 0.05), radius=0.35, length=0.4, axis=2), ("Disc", dict(center=(0.3, -0.1), radius=0.35, dz=0.0))),
}
