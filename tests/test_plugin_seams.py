"""CPU: the two seams into tidy3d (tidy3d_b200/plugin.py) against a stand-in for tidy3d.plugins.mode.mode_solver
(the real package is not importable in this image).  The device call is replaced by a recorder."""
import sys
import types

import numpy as np


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
