"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference's own permittivity sampling:
``ModeSolver._solver_eps`` -> ``_get_epsilon`` -> ``Simulation.epsilon_on_grid`` -> ``Geometry.inside_meshgrid`` /
``Box.inside`` / ``Sphere.inside`` / ``Cylinder.inside`` -> ``Structure.eps_comp``, and the rotation of the tensor into plane
axes (``_tensorial_material_profile_modal_plane_tranform``), mode_solver.py:587-653, simulation.py:1135-1241,
geometry/base.py:145-204, 2043-2070, geometry/primitives.py:44-70, 600-633.

Same technique as oracle/ref_post.py (see there): the method bodies are cut out of the reference's files at run time and
compiled into classes over attribute-holder stand-ins.  Stand-ins: the pydantic field storage, the medium (a tensor-valued
function of frequency: the medium zoo is out of scope), ``Scene`` (background structure) and the list of
structures.  Used to pin the literal stand-in of tests/test_plugin_seams.py, ``plugin.section_of`` and oracle/sections.py,
and to generate tests/golden/sections_ref.npz (tests/golden/make_sections_golden.py).

Only available where ``/root/reference`` exists.  Nothing under ``tidy3d_b200/`` may import this.
"""
from __future__ import annotations

import types
from functools import cached_property
from math import isclose

import numpy as np

from oracle import ref_post as RP

_PARTS = {
    "geom": [("components/geometry/base.py", "Geometry", ["_inds_inside_bounds", "inside_meshgrid", "_ensure_equal_shape", "pop_axis", "unpop_axis"])],
    "box": [("components/geometry/base.py", "Box", ["inside", "bounds"])],
    "sphere": [("components/geometry/primitives.py", "Sphere", ["inside", "bounds"])],
    "cylinder": [("components/geometry/primitives.py", "Cylinder", ["inside", "bounds", "_radius_z"]),
                 ("components/geometry/base.py", "Planar", ["finite_length_axis"])],
    "structure": [("components/structure.py", "Structure", ["eps_comp"])],
    "simulation": [("components/simulation.py", "AbstractYeeGridSimulation", ["epsilon_on_grid"])],
    "eps_solver": [("plugins/mode/mode_solver.py", "ModeSolver", ["_get_epsilon", "_tensorial_material_profile_modal_plane_tranform", "_solver_eps"])],
}

available = RP.available
_NS = None


class TensorMedium:
    """Stand-in medium: ``eps_comp(row, col, frequency)`` of a (possibly dispersive, fully anisotropic) tensor."""

    def __init__(self, tensor, slope=0.0):
        self.t, self.slope = np.asarray(tensor, complex), slope
        self.nonlinear_spec = None

    def eps_comp(self, row, col, frequency):
        return self.t[row, col] * (1 + self.slope * (frequency / 2e14 - 1))

    def __eq__(self, other):
        return isinstance(other, TensorMedium) and np.array_equal(self.t, other.t) and self.slope == other.slope

    __hash__ = None


def _namespace():
    global _NS
    if _NS is not None:
        return _NS
    ns = dict(RP._namespace())  # numpy, typing names, Coords / Grid made of the reference's own methods, the log stand-in
    RP._PARTS.update({f"sec_{k}": v for k, v in _PARTS.items()})

    class _LogCtx(type(ns["log"])):
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    ns.update(log=_LogCtx(), isclose=isclose, cached_property=cached_property, LARGE_NUMBER=1e10, NUM_CELLS_WARN_EPSILON=10**18,
              NUM_STRUCTURES_WARN_EPSILON=10**9, AbstractCustomMedium=type("AbstractCustomMedium", (), {}),
              TriangleMesh=type("TriangleMesh", (), {}))
    Model = RP._Model
    Geometry = RP._make_class("Geometry", "sec_geom", [Model], ns)

    class _BoxFields(Geometry):
        _fields = ("center", "size")

    class _SphereFields(Geometry):
        _fields = ("center", "radius")

    class _CylFields(Geometry):
        _fields = ("center", "radius", "length", "axis", "sidewall_angle", "reference_plane")
        _defaults = dict(sidewall_angle=0.0, reference_plane="middle")
        length_axis = property(lambda self: self.length)      # geometry/primitives.py:353-356
        radius_max = property(lambda self: self.radius)       # vertical side walls (:700-713)
        center_axis = property(lambda self: self.center[self.axis])

    ns["Box"] = RP._make_class("Box", "sec_box", [_BoxFields], ns)
    ns["Sphere"] = RP._make_class("Sphere", "sec_sphere", [_SphereFields], ns)
    ns["Cylinder"] = RP._make_class("Cylinder", "sec_cylinder", [_CylFields], ns)

    class _StructFields(Model):
        _fields = ("geometry", "medium")

    ns["Structure"] = RP._make_class("Structure", "sec_structure", [_StructFields], ns)

    class _SimFields(Model):
        _fields = ("structures", "background")
        volumetric_structures = property(lambda self: self.structures)  # no 2-D materials here (simulation.py:1243-1290)

        @property
        def scene(self):  # components/scene.py background_structure: the medium of the simulation everywhere
            return types.SimpleNamespace(background_structure=ns["Structure"](geometry=None, medium=self.background))

    ns["Simulation"] = RP._make_class("Simulation", "sec_simulation", [_SimFields], ns)

    class _SolverFields(Model):
        _fields = ("simulation", "_solver_grid", "normal_axis")

    ns["EpsSolver"] = RP._make_class("EpsSolver", "sec_eps_solver", [_SolverFields], ns)
    _NS = ns
    return ns


def geometry(kind, **kw):
    """``Box(center, size)``, ``Sphere(center, radius)`` or ``Cylinder(center, radius, length, axis)`` made of the reference's methods."""
    return _namespace()[kind](**kw)


def solver(normal_axis, edges, structures, background):
    """The permittivity-sampling part of a ``ModeSolver``: ``edges`` = cell boundaries of the solver grid along x, y, z (one cell
    along the normal), ``structures`` = [(geometry, medium)], later ones override earlier ones."""
    ns = _namespace()
    grid = ns["Grid"](boundaries=ns["Coords"](x=np.asarray(edges[0], float), y=np.asarray(edges[1], float), z=np.asarray(edges[2], float)))
    num_cells = [len(e) - 1 for e in edges]
    object.__setattr__(grid, "num_cells", num_cells)
    sim = ns["Simulation"](structures=[ns["Structure"](geometry=g, medium=m) for g, m in structures], background=background)
    return ns["EpsSolver"](simulation=sim, _solver_grid=grid, normal_axis=normal_axis)
