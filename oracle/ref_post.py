"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference's own post-processing methods to pin
``oracle/postprocess.py`` and to generate the fixtures ``tests/golden/post_*.npz`` (``tests/golden/make_post_golden.py``).

``ModeSolverData`` / ``ModeSolver`` cannot be imported here (xarray, shapely, h5py, autograd ... are not in the image,
SURVEY 8(c)).  What this module does instead: at run time it reads the reference's source files where they lie under
``/root/reference``, cuts the METHOD BODIES listed in ``_PARTS`` out of their classes with ``ast`` (verbatim text, decorators
included; nothing is written into the repository) and compiles them into classes whose bases are the small stand-ins below:

  * ``oracle/mini_xarray.DataArray`` for ``xarray.DataArray`` (labelled broadcasting, interp = scipy interp1d, NaN-skipping sum);
  * ``_Model``: attribute storage with ``dict() / parse_obj() / copy(update=)`` for the pydantic base class;
  * ``_Monitor`` / ``_Plane`` / ``_Simulation`` / ``_Solver``: the few ATTRIBUTES those methods read (``size``, ``center``,
    ``bounds``, ``zero_dims``, ``colocate``, ``interval_space``, ``freqs``, ``mode_spec``; ``discretize_monitor`` follows
    simulation.py:1041-1073: a non-colocating monitor is extended by one cell on both sides, a colocating one on the right).

Every arithmetic step of the path -- gauge (mode_solver.py:794-818), Yee sites (grid.py:419-491), symmetry expansion
(monitor_data.py:235-282), colocation (mode_solver.py:490-515, monitor_data.py:527-542), differential area (:425-467),
grid correction (mode_solver.py:847-904, monitor_data.py:469-503), Poynting / flux (:566-618), dot / outer_dot (:640-910),
TE fraction (:1594-1652), mode tracking (:1295-1505), normalisation / polarisation filter (mode_solver.py:517-549), the
construction of the data arrays (mode_solver.py:340-415) and the order of the steps (``data_raw``, :306-340) -- is the
reference's own code.

Only available where ``/root/reference`` exists (the build container).  Nothing under ``tidy3d_b200/`` may import this.
"""
from __future__ import annotations

import ast
import os
import textwrap
import types
from functools import cached_property

import numpy as np

from oracle import mini_xarray as mx
from oracle.ref_shim import REF_ROOT

_T3D = os.path.join(REF_ROOT, "tidy3d")

# file (relative to tidy3d/) -> class -> method names cut out of the reference
_PARTS = {
    "data": [
        ("components/data/monitor_data.py", "MonitorData", ["_updated"]),
        ("components/data/dataset.py", "ElectromagneticFieldDataset", ["field_components", "grid_locations", "symmetry_eigenvalues"]),
        ("components/data/monitor_data.py", "AbstractFieldData", ["symmetry_expanded", "_symmetry_update_dict"]),
        ("components/data/monitor_data.py", "ElectromagneticFieldData", [
            "_expanded_grid_field_coords", "_grid_correction_dict", "_tangential_dims", "colocation_boundaries",
            "_plane_grid_boundaries", "_plane_grid_centers", "_diff_area", "_tangential_corrected", "_tangential_fields",
            "_colocated_fields", "_colocated_tangential_fields", "poynting", "package_flux_results", "flux", "dot",
            "_interpolated_tangential_fields", "outer_dot", "_outer_fn_summation"]),
        ("components/data/monitor_data.py", "ModeData", [
            "overlap_sort", "_isel", "_assign_coords", "_find_ordering_one_freq", "_find_closest_pairs", "_reorder_modes",
            "_colocated_propagation_axes_field", "pol_fraction"]),
    ],
    "solver": [
        ("plugins/mode/mode_solver.py", "ModeSolver", [
            "data_raw", "_data_on_yee_grid", "_colocate_data", "_normalize_modes", "_filter_polarization", "_rotate_field_coords",
            "_process_fields", "_postprocess_solver_fields", "_grid_correction"]),
    ],
    "grid": [("components/grid/grid.py", "Grid", ["_avg", "_min", "centers", "sizes", "yee", "__getitem__", "_yee_e", "_yee_h"])],
    "coords": [("components/grid/grid.py", "Coords", ["to_dict", "to_list"])],
    "array": [("components/data/data_array.py", "DataArray", ["multiply_at", "abs"])],
    "geometry": [("components/geometry/base.py", "Geometry", ["pop_axis", "unpop_axis", "rotate_points"])],
    "monitor": [("components/base_sim/monitor.py", "AbstractMonitor", ["downsample"])],
    "rotation": [("components/transformation.py", "AbstractRotation", ["rotate_vector"]),
                 ("components/transformation.py", "RotationAroundAxis", ["isidentity", "matrix"])],
}


def available() -> bool:
    return os.path.isfile(os.path.join(_T3D, "components", "data", "monitor_data.py"))


_trees = {}


def _cut(rel, cls, names):
    """Verbatim source text of the methods ``names`` of class ``cls`` in the reference file ``rel`` (decorators included)."""
    path = os.path.join(_T3D, rel)
    if path not in _trees:
        with open(path) as f:
            src = f.read()
        _trees[path] = (src.splitlines(), ast.parse(src))
    lines, tree = _trees[path]
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    out, found = [], set()
    for item in node.body:
        if isinstance(item, (ast.FunctionDef, ast.AsyncFunctionDef)) and item.name in names:
            start = min([item.lineno] + [d.lineno for d in item.decorator_list]) - 1
            out.append(textwrap.dedent("\n".join(lines[start:item.end_lineno])))
            found.add(item.name)
    missing = set(names) - found
    if missing:
        raise RuntimeError(f"reference {rel}: class {cls} has no {sorted(missing)} (another version of the reference?)")
    return out


def _make_class(name, key, bases, ns):
    """class ``name``(*bases) whose body is the reference's method text listed under ``_PARTS[key]``, compiled with ``ns`` as
    its globals (the names those bodies use: numpy, the stand-ins, the other classes built the same way)."""
    body = []
    for rel, cls, names in _PARTS[key]:
        body += _cut(rel, cls, names)
    base_names = []
    for i, b in enumerate(bases):
        ns[f"_base_{name}_{i}"] = b
        base_names.append(f"_base_{name}_{i}")
    src = "from __future__ import annotations\n" + f"class {name}({', '.join(base_names)}):\n" + textwrap.indent("\n\n".join(body), "    ") + "\n"
    exec(compile(src, f"<reference methods: {key}>", "exec"), ns)
    return ns[name]


# ---------------------------------------------------------------------------------------------------------------------
# stand-ins (attribute holders; no arithmetic of the path lives here)
# ---------------------------------------------------------------------------------------------------------------------
class _Model:
    """What the cut methods use of ``Tidy3dBaseModel``: field storage, ``dict``, ``parse_obj``, ``copy(update=)``."""

    _fields: tuple = ()

    def __init__(self, **kw):
        unknown = set(kw) - set(self._fields)
        if unknown:
            raise TypeError(f"{type(self).__name__}: unknown fields {sorted(unknown)}")
        for k in self._fields:
            object.__setattr__(self, k, kw.get(k, getattr(type(self), "_defaults", {}).get(k)))

    def dict(self, **_):
        return {k: getattr(self, k) for k in self._fields}

    @classmethod
    def parse_obj(cls, d):
        return cls(**d)

    def copy(self, update=None, deep=False, **_):
        d = {k: (v.copy() if deep and hasattr(v, "copy") else v) for k, v in self.dict().items()}
        d.update(update or {})
        return type(self)(**d)

    def updated_copy(self, **kw):
        return self.copy(update=kw)


class _ModeSpec(_Model):
    _fields = ("num_modes", "angle_theta", "angle_phi", "filter_pol", "track_freq", "group_index_step")
    _defaults = dict(angle_theta=0.0, angle_phi=0.0, filter_pol=None, track_freq="central", group_index_step=0)


class _Log:
    def __init__(self):
        self.messages = []

    def warning(self, msg, *a, **k):
        self.messages.append(str(msg))

    info = debug = error = warning


def _typed_array(data, coords=None, dims=None, **kw):
    """``FreqModeDataArray(...)``, ``ModeIndexDataArray(...)``, ...: the reference's typed DataArray subclasses differ from
    the base class by validated dimension names only."""
    if isinstance(data, mx.DataArray) and coords is None and dims is None:
        return _NS["DataArray"](data)
    return _NS["DataArray"](data, coords=coords, dims=dims)


_NS = None
_LAST_REORDER = [None, None]


def _namespace():
    """Build (once) the classes made of the reference's method bodies."""
    global _NS
    if _NS is not None:
        return _NS
    if not available():
        raise RuntimeError(f"the reference tree is not under {REF_ROOT}")
    import importlib.util
    from typing import Any, Callable, Dict, List, Literal, Tuple, Union

    spec = importlib.util.spec_from_file_location("_b200_ref_constants", os.path.join(_T3D, "constants.py"))
    consts = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(consts)  # numpy only

    ns = dict(np=np, Any=Any, Callable=Callable, Dict=Dict, List=List, Literal=Literal, Tuple=Tuple, Union=Union,
              cached_property=cached_property, C_0=consts.C_0, fp_eps=consts.fp_eps, ETA_0=consts.ETA_0,
              DataError=RuntimeError, SetupError=RuntimeError, ValidationError=ValueError, Tidy3dNotImplementedError=NotImplementedError,
              log=_Log(), isbox=lambda x: False, MODE_MONITOR_NAME="<<<MODE_SOLVER_MONITOR>>>", pydantic=types.SimpleNamespace(NonNegativeInt=int))
    _NS = ns
    # DataArray = mini_xarray + the two helpers the reference adds in its own subclass (data_array.py:237-305)
    DataArray = _make_class("DataArray", "array", [mx.DataArray], ns)
    ns["xr"] = types.SimpleNamespace(DataArray=DataArray, Dataset=mx.Dataset)
    for nm in ("FluxDataArray", "FreqModeDataArray", "ModeAmpsDataArray", "MixedModeDataArray", "ModeIndexDataArray",
               "ScalarModeFieldDataArray", "ScalarFieldDataArray"):
        ns[nm] = _typed_array

    class _CoordsBase(_Model):
        _fields = ("x", "y", "z")

    _make_class("Coords", "coords", [_CoordsBase], ns)

    class _GridBase(_Model):
        _fields = ("boundaries",)

    class _FieldGrid(_Model):
        _fields = ("x", "y", "z")

    class _YeeGrid(_Model):
        _fields = ("E", "H")

    ns.update(FieldGrid=_FieldGrid, YeeGrid=_YeeGrid)
    _make_class("Grid", "grid", [_GridBase], ns)

    class _RotBase(_Model):
        _fields = ("axis", "angle")

    _make_class("RotationAroundAxis", "rotation", [_RotBase], ns)
    Geometry = _make_class("Geometry", "geometry", [object], ns)
    MonitorBase = _make_class("MonitorBase", "monitor", [_Model, Geometry], ns)

    class Monitor(MonitorBase):
        """Attributes of ``ModeSolverMonitor`` read by the data methods."""

        _fields = ("center", "size", "freqs", "mode_spec", "colocate", "interval_space", "store_fields_direction", "name")
        _defaults = dict(colocate=False, interval_space=(1, 1, 1), store_fields_direction="+", name="mode")

        @property
        def bounds(self):  # geometry/base.py Box.bounds
            c, s = np.asarray(self.center, float), np.asarray(self.size, float)
            return tuple(c - s / 2), tuple(c + s / 2)

        @property
        def zero_dims(self):  # geometry/base.py Box.zero_dims
            return [d for d, s in enumerate(self.size) if s == 0]

    class Plane(_Model, Geometry):
        _fields = ("center", "size")

    ns.update(Box=Plane, Monitor=Monitor)

    class _DataBase(_Model):
        _fields = ("monitor", "symmetry", "symmetry_center", "grid_expanded", "grid_primal_correction", "grid_dual_correction",
                   "Ex", "Ey", "Ez", "Hx", "Hy", "Hz", "n_complex", "eps_spec")
        _defaults = dict(symmetry=(0, 0, 0), grid_primal_correction=1.0, grid_dual_correction=1.0)

    RefData = _make_class("_RefModeSolverData", "data", [_DataBase], ns)

    class ModeSolverData(RefData):
        """Observation only: keeps the (sorting, phase) arrays ``overlap_sort`` hands to ``_reorder_modes``."""

        def _reorder_modes(self, sorting, phase, track_freq):
            _LAST_REORDER[:] = [np.array(sorting), np.array(phase)]
            return super()._reorder_modes(sorting=sorting, phase=phase, track_freq=track_freq)

    ns["ModeSolverData"] = ModeSolverData
    _make_class("ModeSolver", "solver", [_SolverBase], ns)
    return ns


class _Simulation:
    """``discretize_monitor`` (simulation.py:1041-1073) on the plane's own grid: ``full`` are the cell boundaries of the
    non-colocating monitor's discretisation (= the solver grid, mirrored at a symmetry plane); a colocating monitor starts
    one cell later on the left."""

    def __init__(self, full, normal_pos, normal_primal, normal_dual, symmetry3, center3):
        self.full, self.normal_pos = full, normal_pos
        self.symmetry, self.center = symmetry3, center3
        ns = _namespace()
        self.grid = types.SimpleNamespace(
            boundaries=ns["Coords"](x=full[0], y=full[1], z=np.asarray(normal_primal, float)),
            centers=ns["Coords"](x=None, y=None, z=np.asarray(normal_dual, float)))

    def discretize_monitor(self, monitor):
        ns = _namespace()
        lo = [1 if (monitor.colocate and c.size > 2) else 0 for c in self.full]
        z = np.array([self.normal_pos, self.normal_pos])  # snap_to_box_zero_dim: the normal axis collapses onto the plane
        return ns["Grid"](boundaries=ns["Coords"](x=self.full[0][lo[0]:], y=self.full[1][lo[1]:], z=z))


class _SolverBase:
    """Attributes of ``ModeSolver`` read by the cut methods; ``_solve_all_freqs`` hands back the given solver output."""

    def __init__(self, coords, symmetry, freqs, mode_spec, solver_out, direction="+", colocate=True, plane_size=None, plane_center=None,
                 normal_pos=0.0, normal_primal=None, normal_dual=None):
        ns = _namespace()
        self.normal_axis = 2
        self.freqs, self.mode_spec, self.direction, self.colocate = list(freqs), mode_spec, direction, colocate
        self.solver_symmetry = tuple(symmetry)
        cx, cy = (np.asarray(c, float) for c in coords)
        self._coords = (cx, cy)
        z = np.array([normal_pos, normal_pos])
        self._solver_grid = ns["Grid"](boundaries=ns["Coords"](x=cx, y=cy, z=np.array([normal_pos - 0.5, normal_pos + 0.5])))
        # a one-cell axis is a zero-size dimension of the simulation: its boundaries collapse onto the cell centre in the data
        # grids (simulation.py _snap_zero_dim) and it is never extended (mode_solver.py:222-225)
        sx, sy = (np.array([0.5 * (c[0] + c[1])] * 2) if c.size == 2 else c for c in (cx, cy))
        self.grid_snapped = ns["Grid"](boundaries=ns["Coords"](x=sx, y=sy, z=z))
        # the symmetry plane is the first boundary of the (half-domain) solver grid; the simulation's grid is the mirrored one
        full = [np.concatenate([2 * c[0] - c[:0:-1], c]) if s != 0 else c for c, s in zip((sx, sy), symmetry)]
        center3 = (float(cx[0]) if symmetry[0] else 0.5 * (cx[0] + cx[-1]), float(cy[0]) if symmetry[1] else 0.5 * (cy[0] + cy[-1]), normal_pos)
        if plane_size is None:  # the whole discretised region and beyond: nothing is truncated in _diff_area
            plane_size = (np.inf, np.inf)
        if plane_center is not None:
            assert not any(symmetry), "a symmetric plane is centred on the symmetry planes"
            center3 = (plane_center[0], plane_center[1], normal_pos)
        self.plane = ns["Box"](center=center3, size=(plane_size[0], plane_size[1], 0.0))
        normal_primal = [normal_pos] if normal_primal is None else normal_primal
        normal_dual = [normal_pos] if normal_dual is None else normal_dual
        self.simulation = _Simulation(full, normal_pos, normal_primal, normal_dual, (symmetry[0], symmetry[1], 0), center3)
        self._solver_out = solver_out
        self.reduced_simulation_copy = self

    def to_mode_solver_monitor(self, name, colocate=None):
        ns = _namespace()
        return ns["Monitor"](center=self.plane.center, size=self.plane.size, freqs=self.freqs, mode_spec=self.mode_spec,
                             colocate=self.colocate if colocate is None else colocate, store_fields_direction=self.direction, name=name)

    def _solve_all_freqs(self, coords, symmetry):
        """mode_solver.py:655-672 with ``compute_modes`` replaced by the given per-frequency output (fields, n_complex, eps_spec)."""
        assert all(np.array_equal(a, b) for a, b in zip(coords, self._coords)) and tuple(symmetry) == self.solver_symmetry
        n_complex, fields, eps_spec = [], [], []
        for solver_fields, n_c, spec in self._solver_out:
            n_complex.append(np.asarray(n_c))
            fields.append(self._postprocess_solver_fields(np.array(solver_fields, dtype=complex, copy=True)))
            eps_spec.append(spec)
        return n_complex, fields, eps_spec

    def _field_decay_warning(self, field_data):  # a log message only (mode_solver.py:820-845)
        return None


# ---------------------------------------------------------------------------------------------------------------------
# entry points used by the tests and the fixture generator
# ---------------------------------------------------------------------------------------------------------------------
def mode_spec(num_modes, angle_theta=0.0, angle_phi=0.0, filter_pol=None, track_freq=None):
    return _ModeSpec(num_modes=num_modes, angle_theta=angle_theta, angle_phi=angle_phi, filter_pol=filter_pol, track_freq=track_freq,
                     group_index_step=0)


def solver(fields_per_freq, n_complex, coords, freqs, symmetry=(0, 0), direction="+", colocate=True, angle_theta=0.0, angle_phi=0.0,
           filter_pol=None, track_freq=None, plane_size=None, plane_center=None, normal_pos=0.0, normal_primal=None, normal_dual=None):
    """A ``ModeSolver`` made of the reference's methods around the given ``compute_modes`` output.
    ``fields_per_freq[i]``: (2,3,Nx,Ny,1,M) in solver-plane axes, ``n_complex[i]``: (M,)."""
    ns = _namespace()
    m = int(np.asarray(n_complex[0]).size)
    spec = mode_spec(m, angle_theta, angle_phi, filter_pol, track_freq)
    out = [(f, n, "diagonal") for f, n in zip(fields_per_freq, n_complex)]
    return ns["ModeSolver"](coords, symmetry, freqs, spec, out, direction=direction, colocate=colocate, plane_size=plane_size,
                            plane_center=plane_center, normal_pos=normal_pos, normal_primal=normal_primal, normal_dual=normal_dual)


def packed(data, comps=("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")):
    """Field arrays of a ``ModeSolverData`` as one (2,3,Px,Py,F,M) array (the normal axis dropped)."""
    a = np.array([getattr(data, c).values[:, :, 0] for c in comps])
    return a.reshape((2, 3) + a.shape[1:])


def log_messages():
    return _namespace()["log"].messages


def last_reorder():
    """(sorting[F,M], phase[F,M]) of the latest ``overlap_sort`` (monitor_data.py:1360-1365)."""
    return tuple(_LAST_REORDER)
