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

        def _solve_all_freqs(self, coords, symmetry):
            raise AssertionError("reference loop should have been replaced")

        @property
        def data(self):
            return self._solve_all_freqs(coords=[np.arange(5.0), np.arange(6.0)], symmetry=(0, 0))

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
        seen.append(problems)
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
    # run_batch mirrors web.api.mode.run_batch: a list of ModeSolver -> list of .data
    res = plugin.run_batch([cls(freqs=[2.0e14], num_modes=1), cls(freqs=[1.5e14, 1.6e14], num_modes=3)])
    assert len(res) == 2 and len(res[1][0]) == 2
