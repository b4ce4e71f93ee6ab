"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the permittivity sampling ``ModeSolver._solver_eps`` performs for a
cross-section made of boxes (mode_solver.py:587-653 -> Simulation.epsilon_on_grid, simulation.py:1135-1241; Box.inside,
components/geometry/base.py:2042-2068).  PARITY UNPINNED: needs the full tidy3d package to run the reference itself."""
import numpy as np


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
            for rect, med in section.structures:
                inside = (np.abs(sx - rect.center[0]) <= rect.size[0] / 2)[:, None] & (np.abs(sy - rect.center[1]) <= rect.size[1] / 2)[None, :]
                arr[inside] = med.tensor(freq)[row, col]
            out[3 * row + col] = arr
    return out
