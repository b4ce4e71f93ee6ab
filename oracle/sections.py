"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the permittivity sampling ``ModeSolver._solver_eps`` performs for a
cross-section (mode_solver.py:587-653 -> Simulation.epsilon_on_grid, simulation.py:1135-1241) whose structures cut the
plane as boxes (Box.inside, components/geometry/base.py:2042-2068), discs (Cylinder.inside / Sphere.inside,
geometry/primitives.py:600-632, 44-70) or polygons (PolySlab.inside, geometry/polyslab.py:464-546), or whose inside-masks
were evaluated beforehand.

PARITY PINNED for boxes, spheres, cylinders, the override order, the Yee sites of the nine components and the rotation of
the tensor into plane axes: oracle/ref_sections.py runs the reference's own ``_solver_eps`` / ``epsilon_on_grid`` /
``inside_meshgrid`` / ``Box, Sphere, Cylinder.inside`` (method bodies cut out of the reference at run time), its results for
the scene of tests/section_cases.py are committed as tests/golden/sections_ref.npz and reproduced bit for bit
(tests/test_plugin_seams.py).  UNPINNED: ``Polygon`` (the reference asks matplotlib's ``Path.contains_points``, which is not
in this image; pinned against closed forms only)."""
import numpy as np


def inside(shape, sx, sy):
    """Boolean (len(sx), len(sy)) array: the sites (sx[i], sy[j]) the structure's cut contains."""
    kind = type(shape).__name__
    X, Y = np.meshgrid(np.asarray(sx, float), np.asarray(sy, float), indexing="ij")
    # Geometry.inside_meshgrid (base.py:195-204) evaluates ``inside`` only at the sites within the bounding box
    # (_inds_inside_bounds, :164-170: bounds[0] <= site <= bounds[1] with bounds = centre -/+ half size or radius as computed
    # in floating point, base.py:2116-2120, primitives.py:150-152, 644-648): a site on the rim can fall out by one rounding
    if kind == "Rect":  # base.py:2062-2070: dist <= size / 2 on every axis
        hx, hy = shape.size[0] / 2, shape.size[1] / 2
        box = (shape.center[0] - hx <= X) & (X <= shape.center[0] + hx) & (shape.center[1] - hy <= Y) & (Y <= shape.center[1] + hy)
        return box & (np.abs(X - shape.center[0]) <= hx) & (np.abs(Y - shape.center[1]) <= hy)
    if kind == "Disc":  # primitives.py:66-70 / :624-632
        r = shape.radius
        box = (shape.center[0] - r <= X) & (X <= shape.center[0] + r) & (shape.center[1] - r <= Y) & (Y <= shape.center[1] + r)
        dist_x, dist_y, dist_z = np.abs(X - shape.center[0]), np.abs(Y - shape.center[1]), np.abs(shape.dz)
        return box & ((dist_x**2 + dist_y**2 + dist_z**2) <= (shape.radius**2))
    if kind == "Polygon":
        # polyslab.py:509-516 asks matplotlib's Path.contains_points; restated as the winding number of the polygon around the
        # point being non-zero (equal to the even-odd rule for the simple polygons a PolySlab accepts); sites exactly on an
        # edge are unspecified there and here
        v = np.asarray(shape.vertices, float)
        wn = np.zeros(X.shape, int)
        for (x0, y0), (x1, y1) in zip(v, np.roll(v, -1, axis=0)):
            left = (x1 - x0) * (Y - y0) - (X - x0) * (y1 - y0)  # > 0: the point is left of the directed edge
            wn += ((y0 <= Y) & (y1 > Y) & (left > 0)).astype(int)
            wn -= ((y0 > Y) & (y1 <= Y) & (left < 0)).astype(int)
        return wn != 0
    raise TypeError(kind)


def eps_on_grid(section, coords, freq):
    """(9, Nx, Ny) complex array: background value, then every structure in order overwrites the Yee sites it contains
    (simulation.py:1191-1226); eps_x* is sampled at the Ex site (centre, lower boundary), eps_y* at the Ey site, eps_z* at
    the Ez site (simulation.py:1231-1236 with Grid.yee)."""
    x, y = np.asarray(coords[0], float), np.asarray(coords[1], float)
    xc, yc, xb, yb = (x[:-1] + x[1:]) / 2, (y[:-1] + y[1:]) / 2, x[:-1], y[:-1]
    sites = [(xc, yb), (xb, yc), (xb, yb)]
    out = np.zeros((9, xb.size, yb.size), complex)
    for row, (sx, sy) in enumerate(sites):
        for col in range(3):
            arr = np.full((sx.size, sy.size), section.background.tensor(freq)[row, col], complex)
            if getattr(section, "site_medium", None) is not None:  # inside-masks evaluated beforehand, one medium index per site
                for k, med in enumerate(section.media):
                    arr[np.asarray(section.site_medium)[row] == k] = med.tensor(freq)[row, col]
            for shape, med in section.structures:
                arr[inside(shape, sx, sy)] = med.tensor(freq)[row, col]
            out[3 * row + col] = arr
    return out
