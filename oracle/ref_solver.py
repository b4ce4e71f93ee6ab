"""TEST INFRASTRUCTURE ONLY -- a ``ModeSolver`` made of the UNMODIFIED reference's own methods, end to end: permittivity sampling
(``_solver_eps`` .. ``epsilon_on_grid`` .. ``Box / Sphere / Cylinder.inside``), the solve loops (``_solve_all_freqs``,
``_solve_single_freq`` and their ``_relative`` twins, mode_solver.py:655-785) around the reference's ``compute_modes``
(oracle/ref_shim.py), the construction of ``ModeSolverData`` (``_data_on_yee_grid``, ``_data_on_yee_grid_relative``), its
post-processing (``data_raw``, ``data``: colocation, normalisation, polarisation filter, mode tracking, group index,
symmetry expansion) -- registered as the module ``tidy3d.plugins.mode.mode_solver`` so that ``tidy3d_b200.plugin.install()``
and ``plugin.run_batch()`` meet the very code they are written against (SURVEY 8(b) seams 1-3, rows a19 / f-3).

Technique and stand-ins as in oracle/ref_post.py and oracle/ref_sections.py (method bodies cut out of the reference's files at
run time; attribute holders for the pydantic models).  What is NOT the reference here: how the solver grid follows from the
simulation grid and the plane (``_get_solver_grid`` / ``Simulation._discretize_inds_monitor`` / ``_subgrid``: the caller
gives the cell boundaries of the solver grid directly), and the medium classes.

Only available where ``/root/reference`` exists.  Nothing under ``tidy3d_b200/`` may import this.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from oracle import ref_post as RP
from oracle import ref_sections as RS
from oracle import ref_shim

MODULE = "tidy3d.plugins.mode.mode_solver"

_PARTS = {
    "full_solver": [("plugins/mode/mode_solver.py", "ModeSolver", [
        "normal_axis", "solver_symmetry", "_freqs_for_group_index", "_get_data_with_group_index", "data_raw", "_data_on_yee_grid",
        "_data_on_yee_grid_relative", "_colocate_data", "_normalize_modes", "_filter_polarization", "data", "_get_epsilon",
        "_tensorial_material_profile_modal_plane_tranform", "_solver_eps", "_solve_all_freqs", "_solve_all_freqs_relative",
        "_postprocess_solver_fields", "_solve_single_freq", "_rotate_field_coords_inverse", "_postprocess_solver_fields_inverse",
        "_solve_single_freq_relative", "_rotate_field_coords", "_process_fields", "_grid_correction"])],
    "full_data": RP._PARTS["data"] + [
        ("components/data/monitor_data.py", "AbstractFieldData", ["symmetry_expanded_copy"]),
        ("components/data/monitor_data.py", "ModeData", ["_group_index_post_process"]),
        ("components/data/dataset.py", "ModeSolverDataset", ["n_eff"]),
    ],
}

available = RP.available
_MOD = None


class ModeSpec(RP._Model):
    """The attributes of ``ModeSpec`` (components/mode.py:18) the numerics and the caller read."""

    _fields = ("num_modes", "target_neff", "num_pml", "filter_pol", "angle_theta", "angle_phi", "precision", "bend_radius", "bend_axis",
               "track_freq", "group_index_step")
    _defaults = dict(target_neff=None, num_pml=(0, 0), filter_pol=None, angle_theta=0.0, angle_phi=0.0, precision="double", bend_radius=None,
                     bend_axis=None, track_freq="central", group_index_step=0)


def module():
    """The stand-in module ``tidy3d.plugins.mode.mode_solver`` (built once): ``ModeSolver``, ``compute_modes``,
    ``LOCAL_SOLVER_IMPORTED`` -- the three names ``plugin.install()`` touches (mode_solver.py:59-65)."""
    global _MOD
    if _MOD is not None:
        return _MOD
    ns = RS._namespace()
    ref = ref_shim.load()  # registers the stub packages tidy3d / tidy3d.plugins / tidy3d.plugins.mode and the real solver.py
    mod = types.ModuleType(MODULE)
    mod.__dict__.update(ns)
    g = mod.__dict__
    g.update(compute_modes=ref.compute_modes, LOCAL_SOLVER_IMPORTED=True, IMPORT_ERROR_MSG="local solver missing",
             GroupIndexDataArray=RP._typed_array, ModeDispersionDataArray=RP._typed_array, ModeIndexDataArray=RP._typed_array,
             FreqModeDataArray=RP._typed_array, ScalarModeFieldDataArray=RP._typed_array)
    RP._PARTS.update(_PARTS)

    class _DataFields(RP._Model):
        _fields = ("monitor", "symmetry", "symmetry_center", "grid_expanded", "grid_primal_correction", "grid_dual_correction",
                   "Ex", "Ey", "Ez", "Hx", "Hy", "Hz", "n_complex", "n_group_raw", "dispersion_raw", "eps_spec")
        _defaults = dict(symmetry=(0, 0, 0), grid_primal_correction=1.0, grid_dual_correction=1.0)

    RP._make_class("ModeSolverData", "full_data", [_DataFields], g)
    RP._make_class("ModeSolver", "full_solver", [_SolverFields], g)
    parent = sys.modules["tidy3d.plugins.mode"]
    sys.modules[MODULE] = mod
    parent.mode_solver = mod
    sys.modules["tidy3d.plugins"].mode = parent
    sys.modules["tidy3d"].plugins = sys.modules["tidy3d.plugins"]
    _MOD = mod
    return mod


class Simulation:
    """Attribute holder + the reference's ``epsilon_on_grid``.  ``edges``: cell boundaries of the SOLVER grid along x, y, z
    (two boundaries = one cell along the plane normal; the half domain when the plane has a symmetry wall)."""

    def __init__(self, edges, normal_axis, structures, background, symmetry=(0, 0, 0), normal_primal=None, normal_dual=None):
        ns = RS._namespace()
        self.normal_axis = normal_axis
        self.edges = [np.asarray(e, float) for e in edges]
        self.symmetry = tuple(symmetry)
        pos = 0.5 * (self.edges[normal_axis][0] + self.edges[normal_axis][1])
        self.normal_pos = pos
        # the symmetry plane is the first boundary of the half-domain solver grid; the simulation's own grid is the mirrored one
        self.full = [np.concatenate([2 * e[0] - e[:0:-1], e]) if (s != 0 and a != normal_axis) else e
                     for a, (e, s) in enumerate(zip(self.edges, self.symmetry))]
        self.center = tuple(float(e[0]) if (s != 0 and a != normal_axis) else (pos if a == normal_axis else 0.5 * (e[0] + e[-1]))
                            for a, (e, s) in enumerate(zip(self.edges, self.symmetry)))
        primal = np.asarray([pos] if normal_primal is None else normal_primal, float)
        dual = np.asarray([pos] if normal_dual is None else normal_dual, float)
        b = list(self.full)
        c = [None, None, None]
        b[normal_axis], c[normal_axis] = primal, dual
        self.grid = types.SimpleNamespace(boundaries=ns["Coords"](**dict(zip("xyz", b))), centers=ns["Coords"](**dict(zip("xyz", c))))
        self._eps = ns["Simulation"](structures=[ns["Structure"](geometry=geo, medium=m) for geo, m in structures], background=background)
        self.scene = self._eps.scene
        self.volumetric_structures = self._eps.volumetric_structures

    def epsilon_on_grid(self, grid, coord_key="centers", freq=None):
        return self._eps.epsilon_on_grid(grid, coord_key, freq)

    def _snapped(self, edges):
        ns = RS._namespace()
        e = list(edges)
        e[self.normal_axis] = np.array([self.normal_pos, self.normal_pos])  # snap_to_box_zero_dim: the normal axis collapses onto the plane
        return ns["Grid"](boundaries=ns["Coords"](**dict(zip("xyz", e))))

    def discretize_monitor(self, monitor):  # simulation.py:1041-1073
        lo = [1 if (monitor.colocate and a != self.normal_axis and e.size > 2) else 0 for a, e in enumerate(self.full)]
        return self._snapped([e[k:] for e, k in zip(self.full, lo)])


class _SolverFields(RP._Model):
    """Field storage of ``ModeSolver`` + what stands in for the grid machinery of ``Simulation``."""

    _fields = ("simulation", "plane", "mode_spec", "freqs", "direction", "colocate")
    _defaults = dict(direction="+", colocate=True)

    @property
    def _solver_grid(self):
        sim = self.simulation
        g = RS._namespace()["Grid"](boundaries=RS._namespace()["Coords"](**dict(zip("xyz", sim.edges))))
        object.__setattr__(g, "num_cells", [len(e) - 1 for e in sim.edges])
        return g

    @property
    def grid_snapped(self):
        return self.simulation._snapped(self.simulation.edges)

    @property
    def reduced_simulation_copy(self):
        return self

    def to_mode_solver_monitor(self, name, colocate=None):
        return RS._namespace()["Monitor"](center=self.plane.center, size=self.plane.size, freqs=list(self.freqs), mode_spec=self.mode_spec,
                                          colocate=self.colocate if colocate is None else colocate, store_fields_direction=self.direction, name=name)

    def _field_decay_warning(self, field_data):  # a log message only (mode_solver.py:820-845)
        return None


def mode_solver(edges, normal_axis, structures, background, freqs, mode_spec, symmetry=(0, 0, 0), direction="+", colocate=True,
                plane_size=None, normal_primal=None, normal_dual=None):
    """``ModeSolver`` of the stand-in module for a plane normal to ``normal_axis`` of a scene ``structures`` = [(geometry made by
    oracle.ref_sections.geometry, medium)]."""
    mod = module()
    sim = Simulation(edges, normal_axis, structures, background, symmetry, normal_primal, normal_dual)
    size = [np.inf, np.inf, np.inf] if plane_size is None else list(plane_size)
    size[normal_axis] = 0.0
    plane = mod.Box(center=sim.center, size=tuple(size))
    return mod.ModeSolver(simulation=sim, plane=plane, mode_spec=mode_spec, freqs=list(freqs), direction=direction, colocate=colocate)


def data_arrays(data):
    """{name: ndarray} of everything a ``ModeSolverData`` carries (for comparisons)."""
    out = {k: np.asarray(getattr(data, k).values) for k in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz", "n_complex")}
    for k in ("grid_primal_correction", "grid_dual_correction", "n_group_raw", "dispersion_raw"):
        v = getattr(data, k)
        if v is not None and not isinstance(v, float):
            out[k] = np.asarray(v.values)
    return out
