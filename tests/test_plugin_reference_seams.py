"""CPU, build container only (``reference`` marker): the seams of tidy3d_b200/plugin.py installed on a ``ModeSolver`` made of the
UNMODIFIED reference's own methods (oracle/ref_solver.py registers it as ``tidy3d.plugins.mode.mode_solver``: solve loops,
permittivity sampling, data construction, colocation, normalisation, mode tracking, group index, symmetry expansion are the
reference's code; rows a19 / f-2 / f-3 of SURVEY 8).

There is no GPU here, so the device call behind the seams (``compute_modes_batch``) is replaced by a loop over the reference's
``compute_modes`` -- the numerics are not under test in this file (tests/test_gpu_parity.py), the SEAMS are: which problems the
batched replacements hand to the device call (permittivity per frequency, coordinates, symmetry, direction, basis fields, the
section of the f-2 seam) and what they make of its results must leave ``ModeSolver.data`` bit-identical to the reference's own
per-frequency loop.
"""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.reference
C0 = 2.99792458e14


def _available():
    from oracle import ref_post

    return ref_post.available()


def _make(normal=0, track=None, gi=0, sym=(0, 0, 0), colocate=True, nf=3, direction="+", num_modes=2, extra=False):
    from oracle import ref_sections as RS
    from oracle import ref_solver as RSV

    edges = [np.linspace(-1.0, 1.0, 25), np.linspace(-0.9, 0.9, 23), np.linspace(-0.8, 0.8, 21)]
    plane_axes = [a for a in range(3) if a != normal]
    csize = [10.0, 10.0, 10.0]
    csize[plane_axes[0]], csize[plane_axes[1]] = 0.5, 0.22
    edges[normal] = np.array([-0.02, 0.02])
    for a in range(3):
        if sym[a] != 0:
            edges[a] = edges[a][edges[a].size // 2:]
    structures = [(RS.geometry("Box", center=(0.0, 0.0, 0.0), size=tuple(csize)), RS.TensorMedium(3.48**2 * np.eye(3), 0.02))]
    if extra:  # a second, anisotropic structure next to the core and a sphere on top: the f-2 seam has something to draw
        c = [0.0, 0.0, 0.0]
        c[plane_axes[0]] = 0.45
        structures.append((RS.geometry("Box", center=tuple(c), size=(0.3, 0.3, 0.3)), RS.TensorMedium(np.diag([4.0, 4.2, 3.9]), -0.01)))
        structures.append((RS.geometry("Sphere", center=(0.0, 0.0, 0.0), radius=0.2), RS.TensorMedium(3.2**2 * np.eye(3))))
    spec = RSV.ModeSpec(num_modes=num_modes, track_freq=track, group_index_step=gi)
    freqs = C0 / np.linspace(1.5, 1.6, nf)
    return RSV.mode_solver(edges, normal, structures, RS.TensorMedium(1.44**2 * np.eye(3)), freqs, spec, symmetry=sym, direction=direction,
                           colocate=colocate, normal_primal=[-0.03, 0.05], normal_dual=[-0.07, 0.01])


class _CpuDevice:
    """Stands in for the device call: every problem solved by the reference's compute_modes, calls recorded."""

    def __init__(self):
        self.calls = []

    def batch(self, problems, **kw):
        from oracle import ref_shim
        from oracle import sections as OS

        ref = ref_shim.load()
        self.calls.append(list(problems))
        out = []
        for p in problems:
            eps = OS.eps_on_grid(p["section"], p["coords"], p["freq"]) if "section" in p else p["eps_cross"]
            out.append(ref.compute_modes(eps_cross=eps, coords=p["coords"], freq=p["freq"], mode_spec=p["mode_spec"], symmetry=p.get("symmetry", (0, 0)),
                                         direction=p.get("direction", "+"), solver_basis_fields=p.get("solver_basis_fields")))
        return out

    def single(self, **kw):
        return self.batch([kw])[0]


@pytest.fixture()
def seam(monkeypatch):
    if not _available():
        pytest.skip("no reference tree here")
    from oracle import ref_solver as RSV

    import tidy3d_b200.plugin as plugin

    mod = RSV.module()
    dev = _CpuDevice()
    monkeypatch.setattr(plugin, "compute_modes_batch", dev.batch)
    monkeypatch.setattr(plugin, "compute_modes", lambda *a, **k: dev.single(**k))
    saved = {k: getattr(mod.ModeSolver, k) for k in ("_solve_all_freqs", "_solve_all_freqs_relative")}
    saved_cm = mod.compute_modes
    warnings.simplefilter("ignore")
    yield plugin, mod, dev
    for k, v in saved.items():
        setattr(mod.ModeSolver, k, v)
    mod.compute_modes = saved_cm
    plugin.DEVICE_EPS = False


def _same(a, b):
    from oracle import ref_solver as RSV

    da, db = RSV.data_arrays(a), RSV.data_arrays(b)
    assert da.keys() == db.keys()
    for k in da:
        assert da[k].shape == db[k].shape and np.array_equal(da[k], db[k], equal_nan=True), k


@pytest.mark.parametrize("normal,sym,direction", [(0, (0, 0, 0), "+"), (1, (1, 0, -1), "-"), (2, (0, 0, 0), "+")])
def test_batched_seam_leaves_the_data_unchanged(seam, normal, sym, direction):
    """Seam 2: ``_solve_all_freqs`` replaced by ONE batched call; data_raw (colocated, normalised, tracked) identical to the
    reference's own loop, for every plane normal (field rotation, the H sign of a y-normal), with symmetry walls and for
    backward modes."""
    plugin, mod, dev = seam
    want = _make(normal, track="central", sym=sym, direction=direction).data_raw
    ModeSolver = plugin.install(batched=True)
    assert ModeSolver is mod.ModeSolver and mod.compute_modes is plugin.compute_modes and mod.LOCAL_SOLVER_IMPORTED is True
    ms = _make(normal, track="central", sym=sym, direction=direction)
    got = ms.data_raw
    assert len(dev.calls) == 1 and len(dev.calls[0]) == 3  # one device call for the three frequencies
    p = dev.calls[0][1]
    assert p["freq"] == ms.freqs[1] and p["direction"] == direction and tuple(p["symmetry"]) == tuple(ms.solver_symmetry)
    assert np.array_equal(p["eps_cross"], ms._solver_eps(ms.freqs[1]))
    _same(want, got)


def test_single_solve_seam(seam):
    """Seam 1: only the module-level name ``compute_modes`` is rebound (mode_solver.py:59-65, called at :725)."""
    plugin, mod, dev = seam
    want = _make(2).data_raw
    plugin.install(batched=False)
    got = _make(2).data_raw
    assert [len(c) for c in dev.calls] == [1, 1, 1]  # the reference's loop calls the drop-in once per frequency
    _same(want, got)


def test_device_eps_seam(seam):
    """Seam 2b: the plane is described once (``section_of``), the per-frequency arrays never exist on the host; what the
    rasteriser makes of the section (restated: oracle/sections.py, pinned to the reference) gives the same ModeSolverData."""
    plugin, mod, dev = seam
    want = _make(0, extra=True).data_raw
    plugin.install(batched=True, device_eps=True)
    got = _make(0, extra=True).data_raw
    probs = dev.calls[0]
    assert all("eps_cross" not in p and p["section"] is probs[0]["section"] for p in probs)
    assert len(probs[0]["section"].media) == 4
    _same(want, got)


def test_relative_seam(seam):
    """``_solve_all_freqs_relative`` (mode_solver.py:674-693): the basis fields of every frequency go through the reference's
    own ``_postprocess_solver_fields_inverse`` and reach the device call as ``solver_basis_fields``."""
    plugin, mod, dev = seam
    basis = _make(2, colocate=False, num_modes=3).data_raw
    want = _make(2, colocate=False, num_modes=3)._data_on_yee_grid_relative(basis)
    plugin.install(batched=True)
    ms = _make(2, colocate=False, num_modes=3)
    n0 = len(dev.calls)
    got = ms._data_on_yee_grid_relative(basis)
    assert len(dev.calls) == n0 + 1 and all(p["solver_basis_fields"].shape == (6, 24, 22, 1, 3) for p in dev.calls[-1])
    _same(want, got)


def test_run_batch_one_device_call_for_many_solvers(seam):
    """Seam 3: three planes -- one with group index (the reference solves a copy with 3x the frequencies,
    mode_solver.py:267-299), one with symmetry walls, one plain -- solved by ONE device call; every ``ModeSolver.data``
    assembled by the reference's own code equals the reference's own serial result, group index and dispersion included."""
    plugin, mod, dev = seam
    kinds = [dict(normal=0, gi=0.005, track="central"), dict(normal=1, sym=(1, 0, -1)), dict(normal=2, nf=2)]
    want = [_make(**k).data for k in kinds]
    got = plugin.run_batch([_make(**k) for k in kinds])
    assert len(dev.calls) == 1 and len(dev.calls[0]) == 9 + 3 + 2
    for w, g in zip(want, got):
        _same(w, g)
    assert got[0].n_group_raw is not None and got[0].n_complex.values.shape == (3, 2)
    # the formulas of plugin.group_index (for callers that drive compute_modes_batch themselves) against the reference's result
    ms = _make(normal=0, gi=0.005)
    f3 = ms._freqs_for_group_index()
    n3 = np.array([r[1] for r in dev.batch([dict(eps_cross=ms._solver_eps(f), coords=dev.calls[0][0]["coords"], freq=f, mode_spec=ms.mode_spec) for f in f3])])
    f0, n0, ng, disp = plugin.group_index(n3, f3, 0.005)
    assert np.allclose(ng, got[0].n_group_raw.values, rtol=1e-12) and np.allclose(disp, got[0].dispersion_raw.values, rtol=1e-9)


def test_grid_correction_and_plane_bounds_helpers_on_the_reference_solver(seam):
    """``plugin.grid_correction_of`` / ``plane_bounds_of`` read a ModeSolver made of the reference's methods; the factors the host
    helper forms from the table equal ``ModeSolverData.grid_primal_correction / grid_dual_correction`` as the reference's own
    ``_grid_correction`` computed them (mode_solver.py:847-904), for every plane normal and for backward modes."""
    plugin, mod, dev = seam
    from tidy3d_b200 import postprocess as PP

    for normal, direction in ((0, "+"), (1, "-"), (2, "+")):
        ms = _make(normal, direction=direction)
        data = ms._data_on_yee_grid()
        table = plugin.grid_correction_of(ms)
        for i, f in enumerate(ms.freqs):
            primal, dual = PP.grid_correction_factors(data.n_complex.values[i], f, table, 0.0, direction)
            assert np.allclose(primal, data.grid_primal_correction.values[i], rtol=1e-13, atol=1e-15)
            assert np.allclose(dual, data.grid_dual_correction.values[i], rtol=1e-13, atol=1e-15)
        assert np.abs(np.abs(data.grid_dual_correction.values) - 1).max() > 1e-4  # not a no-op in this set-up
        lo, hi = ms.plane.bounds
        axes = [a for a in range(3) if a != normal]
        assert list(plugin.plane_bounds_of(ms)) == [lo[axes[0]], hi[axes[0]], lo[axes[1]], hi[axes[1]]]
