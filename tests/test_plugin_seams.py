"""CPU: the two seams into tidy3d (tidy3d_b200/plugin.py) against a stand-in for tidy3d.plugins.mode.mode_solver
(the real package is not importable in this image).  The device call is replaced by a recorder."""
import os
import sys
import types

import numpy as np
import pytest


def _fake_tidy3d(monkeypatch):
    calls = {}

    class ModeSolver:  # the attributes/methods the seam touches (mode_solver.py:655-735)
        def __init__(self, freqs, num_modes):
            self.freqs = freqs
            self.mode_spec = types.SimpleNamespace(num_modes=num_modes)
            self.direction = "+"

        def _solver_eps(self, freq):
            return [np.full((4, 5), freq * 1e-14, complex) for _ in range(9)]

        def _postprocess_solver_fields(self, solver_fields):
            return {"Ex": solver_fields[0, 0], "Hz": solver_fields[1, 2]}

        def _postprocess_solver_fields_inverse(self, fields):
            return np.stack([fields["Ex"]] * 6)

        def _solve_all_freqs(self, coords, symmetry):
            raise AssertionError("reference loop should have been replaced")

        def _solve_all_freqs_relative(self, coords, symmetry, basis_fields):
            raise AssertionError("reference loop should have been replaced")

        @property
        def data(self):  # like data_raw: the group-index variant solves a copy with 3x the frequencies (mode_solver.py:283-299)
            solver = self
            if getattr(self, "group_index_step", 0):
                solver = ModeSolver(list(np.outer(self.freqs, (1 - self.group_index_step, 1, 1 + self.group_index_step)).flatten()),
                                    self.mode_spec.num_modes)
            return solver._solve_all_freqs(coords=[np.arange(5.0), np.arange(6.0)], symmetry=(0, 0))

    ms = types.ModuleType("tidy3d.plugins.mode.mode_solver")
    ms.ModeSolver = ModeSolver
    ms.compute_modes = lambda *a, **k: calls.setdefault("ref", True)
    ms.LOCAL_SOLVER_IMPORTED = False
    for name in ("tidy3d", "tidy3d.plugins", "tidy3d.plugins.mode"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.setitem(sys.modules, "tidy3d.plugins.mode.mode_solver", ms)
    sys.modules["tidy3d.plugins.mode"].mode_solver = ms
    sys.modules["tidy3d.plugins"].mode = sys.modules["tidy3d.plugins.mode"]
    sys.modules["tidy3d"].plugins = sys.modules["tidy3d.plugins"]
    return ms


def test_install_rebinds_both_seams(monkeypatch):
    ms = _fake_tidy3d(monkeypatch)
    import tidy3d_b200.plugin as plugin
    from tidy3d_b200 import compute_modes

    seen = []

    def fake_batch(problems):
        seen.append(list(problems))
        out = []
        for p in problems:
            m = p["mode_spec"].num_modes
            f = np.full((2, 3, 4, 5, 1, m), p["freq"] * 1e-14, complex)
            out.append((f, np.arange(m) + p["freq"] * 1e-14, "diagonal"))
        return out

    monkeypatch.setattr(plugin, "compute_modes_batch", fake_batch)
    cls = plugin.install(batched=True)
    assert ms.compute_modes is compute_modes and ms.LOCAL_SOLVER_IMPORTED is True
    solver = cls(freqs=[1.9e14, 2.0e14, 2.1e14], num_modes=2)
    n_complex, fields, eps_spec = solver._solve_all_freqs(coords=[np.arange(5.0), np.arange(6.0)], symmetry=(0, 1))
    assert len(seen) == 1 and len(seen[0]) == 3  # ONE device call for the whole frequency loop
    assert [p["freq"] for p in seen[0]] == [1.9e14, 2.0e14, 2.1e14]
    assert all(p["symmetry"] == (0, 1) and p["direction"] == "+" for p in seen[0])
    assert eps_spec == ["diagonal"] * 3 and len(n_complex) == 3
    assert fields[1]["Ex"].shape == (4, 5, 1, 2) and np.allclose(fields[1]["Ex"], 2.0)
    # relative seam (mode_solver.py:674-693): one device call, basis passed through _postprocess_solver_fields_inverse
    basis = [{"Ex": np.zeros((4, 5, 1, 2))}] * 3
    seen.clear()
    solver._solve_all_freqs_relative(coords=[np.arange(5.0), np.arange(6.0)], symmetry=(0, 0), basis_fields=basis)
    assert len(seen) == 1 and all(p["solver_basis_fields"].shape == (6, 4, 5, 1, 2) for p in seen[0])
    # run_batch mirrors web.api.mode.run_batch: a list of ModeSolver -> list of .data, with ONE device call for all of
    # them, including the 3x frequency copy of a group-index solver
    seen.clear()
    gi = cls(freqs=[1.5e14, 1.6e14], num_modes=3)
    gi.group_index_step = 0.01
    res = plugin.run_batch([cls(freqs=[2.0e14], num_modes=1), gi, cls(freqs=[1.9e14], num_modes=2)])
    assert len(seen) == 1 and [len(p) for p in seen] == [1 + 6 + 1]
    assert len(res) == 3 and len(res[0][0]) == 1 and len(res[1][0]) == 6 and len(res[2][0]) == 1
    assert np.allclose(res[1][1][4]["Ex"], 1.6)  # second base frequency, centre of its triple
    assert cls._solve_all_freqs is plugin.solve_all_freqs_batched  # the patch is restored


def test_group_index_formula():
    """monitor_data.py:1507-1548 on an analytic dispersion n(f) = a + b f + c f^2: n_g = n + f dn/df exactly recovered."""
    import tidy3d_b200.plugin as plugin

    a, b, c = 2.0, 1e-15, 3e-30
    f0 = np.array([1.9e14, 2.0e14])
    step = 0.005
    f = np.outer(f0, (1 - step, 1, 1 + step)).flatten()
    n = (a + b * f + c * f**2)[:, None] * np.ones((1, 2))
    fc, nc, ng, disp = plugin.group_index(n, f, step)
    assert np.allclose(fc, f0) and np.allclose(nc[:, 0], a + b * f0 + c * f0**2)
    assert np.allclose(ng[:, 0], a + 2 * b * f0 + 3 * c * f0**2, rtol=1e-12)
    d2 = 2 * c
    assert np.allclose(disp[:, 0], -(f0 / 2.99792458e14) ** 2 * (2 * (b + 2 * c * f0) + f0 * d2) * 1e18, rtol=1e-6)


class _Coords:
    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z


class _BoxGeom:
    """Stand-in geometry with the reference's Geometry.inside_meshgrid contract (geometry/base.py:172-201)."""

    def __init__(self, center, size):
        self.center, self.size = np.array(center, float), np.array(size, float)

    def inside_meshgrid(self, x, y, z):
        X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
        return ((np.abs(X - self.center[0]) <= self.size[0] / 2) & (np.abs(Y - self.center[1]) <= self.size[1] / 2)
                & (np.abs(Z - self.center[2]) <= self.size[2] / 2))


class _BallGeom:
    def __init__(self, center, radius):
        self.center, self.radius = np.array(center, float), radius

    def inside_meshgrid(self, x, y, z):
        X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
        return (X - self.center[0]) ** 2 + (Y - self.center[1]) ** 2 + (Z - self.center[2]) ** 2 <= self.radius**2


class _CylGeom:
    """Cylinder with vertical side walls (geometry/primitives.py:600-633)."""

    def __init__(self, center, radius, length, axis):
        self.center, self.radius, self.length, self.axis = np.array(center, float), radius, length, axis

    def inside_meshgrid(self, x, y, z):
        g = list(np.meshgrid(x, y, z, indexing="ij"))
        d = [np.abs(g[a] - self.center[a]) for a in range(3)]
        t = [a for a in range(3) if a != self.axis]
        return (d[t[0]] ** 2 + d[t[1]] ** 2 <= self.radius**2) & (d[self.axis] <= self.length / 2)


def _literal_geometry(kind, kw):
    return {"Box": _BoxGeom, "Sphere": _BallGeom, "Cylinder": _CylGeom}[kind](**kw)


class _TensorMedium:
    def __init__(self, t, slope=0.0):
        self.t, self.slope = np.asarray(t, complex), slope

    def eps_comp(self, row, col, frequency):
        return self.t[row, col] * (1 + self.slope * (frequency / 2e14 - 1))

    def __eq__(self, other):
        return isinstance(other, _TensorMedium) and np.array_equal(self.t, other.t) and self.slope == other.slope


class _Structure:
    def __init__(self, geometry, medium):
        self.geometry, self.medium = geometry, medium

    def eps_comp(self, row, col, frequency, coords):  # structure.py:279-300
        return self.medium.eps_comp(row, col, frequency)


class _PlaneSolver:
    """The part of ModeSolver that feeds the solver its permittivity, written out literally from the reference:
    _get_epsilon (mode_solver.py:587-593) -> Simulation.epsilon_on_grid (simulation.py:1179-1236) for the nine keys, then
    _tensorial_material_profile_modal_plane_tranform (:594-624)."""

    def __init__(self, normal_axis, bounds, n, structures, background):
        self.normal_axis = normal_axis
        edges = [np.linspace(lo, hi, k + 1) for (lo, hi), k in zip(bounds, n)]
        edges[normal_axis] = np.array([-0.01, 0.01])  # one cell along the normal
        centers = [(e[:-1] + e[1:]) / 2 for e in edges]
        lower = [e[:-1] for e in edges]
        self._grid = {}
        for a, key in enumerate(("Ex", "Ey", "Ez")):  # Yee E_a site: centre along a, lower cell boundary along the others
            self._grid[key] = _Coords(*[centers[d] if d == a else lower[d] for d in range(3)])
        self.edges = edges
        self.simulation = types.SimpleNamespace(scene=types.SimpleNamespace(background_structure=_Structure(None, background)),
                                                volumetric_structures=structures)
        self.freqs = [1.9e14, 2.1e14]
        self.mode_spec = types.SimpleNamespace(num_modes=1)
        self.direction = "+"

    @property
    def _solver_grid(self):
        return self._grid

    def _epsilon_on_grid(self, coord_key, freq):
        row = "xyz".index(coord_key[1])
        col = row if len(coord_key) == 2 else "xyz".index(coord_key[2])
        c = self._grid[coord_key[0:2]]
        arrays = (np.array(c.x), np.array(c.y), np.array(c.z))
        sim = self.simulation
        eps_array = sim.scene.background_structure.eps_comp(row, col, freq, None) * np.ones(tuple(len(a) for a in arrays), dtype=complex)
        for structure in sim.volumetric_structures:
            is_inside = structure.geometry.inside_meshgrid(*arrays)
            eps_array[is_inside] = structure.eps_comp(row, col, freq, None)
        return eps_array

    def _solver_eps(self, freq):
        keys = ["Ex", "Exy", "Exz", "Eyx", "Ey", "Eyz", "Ezx", "Ezy", "Ez"]
        mat_data = np.stack([self._epsilon_on_grid(k, freq) for k in keys], axis=0)
        mat_tensor = np.take(mat_data, indices=[0], axis=1 + self.normal_axis)
        mat_tensor = np.squeeze(mat_tensor, axis=1 + self.normal_axis)
        flat_shape = np.shape(mat_tensor)
        mat_tensor = mat_tensor.reshape([3, 3] + list(flat_shape[1:]))
        if self.normal_axis == 0:
            mat_tensor[[0, 1], :, ...] = mat_tensor[[1, 0], :, ...]
            mat_tensor[:, [0, 1], ...] = mat_tensor[:, [1, 0], ...]
        if self.normal_axis <= 1:
            mat_tensor[[1, 2], :, ...] = mat_tensor[[2, 1], :, ...]
            mat_tensor[:, [1, 2], ...] = mat_tensor[:, [2, 1], ...]
        return mat_tensor.reshape(flat_shape)

    def plane_coords(self):
        return [self.edges[a] for a in range(3) if a != self.normal_axis]


def test_section_of_reproduces_solver_eps_for_every_plane_orientation(built_lib):
    """Seam 2b: the per-plane description built by plugin.section_of + the rasteriser give exactly the array
    ModeSolver._solver_eps samples per frequency -- for the three plane normals (tensor rows / columns and Yee sites rotate
    together), fully anisotropic dispersive media, overlapping structures, media shared between structures."""
    import ctypes as C

    import tidy3d_b200.plugin as plugin
    from oracle import sections as OS

    from tests import section_cases as SC

    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sections_ref.npz"))
    for normal in (0, 1, 2):
        structures = [_Structure(_literal_geometry(kind, kw), _TensorMedium(t, slope=slope)) for kind, kw, t, slope in SC.STRUCTURES]
        ms = _PlaneSolver(normal, SC.BOUNDS, SC.CELLS, structures, _TensorMedium(*SC.BACKGROUND))
        ms.freqs = list(SC.FREQS)
        sec = plugin.section_of(ms)
        assert len(sec.media) == 5  # background + four distinct media (the third structure re-uses silicon)
        coords = ms.plane_coords()
        for k, freq in enumerate(ms.freqs):
            want = ms._solver_eps(freq)
            # the literal stand-in above IS what the unmodified reference samples (fixture made by oracle/ref_sections.py from
            # the reference's own epsilon_on_grid / inside_meshgrid / Box, Sphere, Cylinder.inside / plane transform)
            assert np.array_equal(want, ref[f"eps_n{normal}_f{k}"])
            got = OS.eps_on_grid(sec, coords, freq)
            assert got.shape == want.shape and np.array_equal(got, want)
            # ... and the library's own rasteriser (host mirror of section_raster_kernel) sets the same problem up
            outs = []
            for pk in (built_lib.PackedProblem(None, coords, freq, ms.mode_spec, section=sec), built_lib.PackedProblem(want, coords, freq, ms.mode_spec)):
                f = np.zeros((6, pk.nx * pk.ny), complex)
                flags, sigma = (C.c_int * 4)(), np.zeros(2)
                rc = built_lib.lib().b200ms_debug_setup(C.byref(pk.struct), built_lib._ptr(sigma), flags, None, None, None, None, built_lib._ptr(f.view(float)))
                assert rc == 0
                outs.append((list(flags), sigma.copy(), f))
            assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    # the batched seam hands the section to the device call instead of nine arrays per frequency
    plugin.DEVICE_EPS = True
    try:
        probs = plugin._problems(ms, coords, (0, 0))
    finally:
        plugin.DEVICE_EPS = False
    assert all("eps_cross" not in p and p["section"] is probs[0]["section"] for p in probs) and len(probs) == 2
    # custom (space-dependent) media keep the sampled-array path
    structures[0].medium.eps_comp_on_grid = lambda *a, **k: None
    assert plugin.section_of(ms) is None


@pytest.mark.reference
def test_section_of_on_a_solver_made_of_the_reference_code():
    """Live: ``plugin.section_of`` handed a ModeSolver whose sampling path IS the reference's code (oracle/ref_sections.py);
    section + restated rasteriser == the reference's ``_solver_eps`` for every plane normal, and the committed fixture is
    what the reference tree here produces."""
    from oracle import ref_sections as RS

    if not RS.available():
        pytest.skip("no reference tree here")
    import tidy3d_b200.plugin as plugin
    from oracle import sections as OS
    from tests import section_cases as SC

    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sections_ref.npz"))
    for normal in (0, 1, 2):
        structures = [(RS.geometry(kind, **kw), RS.TensorMedium(t, slope)) for kind, kw, t, slope in SC.STRUCTURES]
        ms = RS.solver(normal, SC.edges(normal), structures, RS.TensorMedium(*SC.BACKGROUND))
        sec = plugin.section_of(ms)
        coords = [e for a, e in enumerate(SC.edges(normal)) if a != normal]
        for k, freq in enumerate(SC.FREQS):
            want = np.array(ms._solver_eps(freq))
            assert np.array_equal(want, ref[f"eps_n{normal}_f{k}"])
            assert np.array_equal(OS.eps_on_grid(sec, coords, freq), want)


def test_primitive_cuts_equal_the_reference_geometry(built_lib):
    """Rect / Disc of the restatement (oracle/sections.py) and of the library's rasteriser (host mirror of the device code,
    csrc/medium.cuh section_cell) mark exactly the sites the reference's Box / Sphere / Cylinder.inside_meshgrid mark
    (fixture made by the reference's own methods, oracle/ref_sections.py) -- sites on an edge or on the circle included."""
    import tidy3d_b200.sections as S
    from oracle import sections as OS
    from tests import section_cases as SC

    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sections_ref.npz"))
    for name, (_, _, (kind, kw)) in SC.PRIMITIVES.items():
        shape = getattr(S, kind)(**kw)
        want = ref[f"mask_{name}"]
        assert want.any() and not want.all()
        assert np.array_equal(OS.inside(shape, SC.SX, SC.SY), want), name
        # the library: boundaries chosen so that the Ez sites (lower cell boundaries) are the fixture's sites, and the Ex / Ey
        # sites (cell centres along one axis) the same sites shifted by half a cell
        import ctypes as C

        coords = [np.r_[SC.SX, SC.SX[-1] + 0.05], np.r_[SC.SY, SC.SY[-1] + 0.05]]
        sec = S.Section(background=S.Medium(2.0), structures=[(shape, S.Medium((11.0, 12.0, 13.0)))])
        eps = OS.eps_on_grid(sec, coords, 2e14)
        assert np.array_equal(eps[8] == 13.0, want)
        spec = types.SimpleNamespace(num_modes=1)
        outs = []
        for pk in (built_lib.PackedProblem(None, coords, 2e14, spec, section=sec), built_lib.PackedProblem(eps, coords, 2e14, spec)):
            f = np.zeros((6, pk.nx * pk.ny), complex)
            flags, sigma = (C.c_int * 4)(), np.zeros(2)
            assert built_lib.lib().b200ms_debug_setup(C.byref(pk.struct), built_lib._ptr(sigma), flags, None, None, None, None, built_lib._ptr(f.view(float))) == 0
            outs.append((list(flags), sigma.copy(), f))
        assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2]), name


@pytest.mark.reference
def test_primitive_cut_fixture_is_current():
    from oracle import ref_sections as RS

    if not RS.available():
        pytest.skip("no reference tree here")
    from tests import section_cases as SC

    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sections_ref.npz"))
    for name, (kind, kw, _) in SC.PRIMITIVES.items():
        assert np.array_equal(RS.geometry(kind, **kw).inside_meshgrid(SC.SX, SC.SY, np.array([0.0]))[:, :, 0], ref[f"mask_{name}"])


def test_grid_correction_of_reads_the_normal_grid_of_the_simulation():
    import tidy3d_b200.plugin as plugin
    from oracle import postprocess as OP
    from tidy3d_b200 import postprocess as PP

    bounds = [np.linspace(-1, 1, 21), np.array([-0.3, -0.1, 0.05, 0.2, 0.4]), np.linspace(-1, 1, 11)]
    centers = [(b[:-1] + b[1:]) / 2 for b in bounds]
    ms = types.SimpleNamespace(
        normal_axis=1, plane=types.SimpleNamespace(center=(0.0, 0.08, 0.0)),
        simulation=types.SimpleNamespace(grid=types.SimpleNamespace(boundaries=types.SimpleNamespace(to_list=bounds),
                                                                    centers=types.SimpleNamespace(to_list=centers))))
    table = plugin.grid_correction_of(ms)
    n = np.array([2.4 + 0.01j, 1.7])
    got = PP.grid_correction_factors(n, 1.9e14, table)
    want = OP.grid_correction(n, 1.9e14, bounds[1], centers[1], 0.08)
    assert np.allclose(got[0], want[0], rtol=1e-14) and np.allclose(got[1], want[1], rtol=1e-14)
