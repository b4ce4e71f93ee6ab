"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference numerics.

Imports ``tidy3d/plugins/mode/{solver,derivatives,transforms}.py`` straight from
``/root/reference`` under a stub package (full ``import tidy3d`` needs xarray/shapely/... which
this image lacks; recipe from SURVEY.md Appendix C).  Only usable in the build container: the GPU
box has no ``/root/reference``.  Used to (a) pin ``oracle/restatement.py`` and (b) generate the
golden fixtures under ``tests/golden`` (``tests/golden/make_golden.py``).

Nothing under ``tidy3d_b200/`` may import this module.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

REF_ROOT = os.environ.get("B200MS_REFERENCE", "/root/reference")
_REF = os.path.join(REF_ROOT, "tidy3d")


def available() -> bool:
    return os.path.isfile(os.path.join(_REF, "plugins", "mode", "solver.py"))


def _mod(name, path=None):
    if path is None:
        m = types.ModuleType(name)
        m.__path__ = []
    else:
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    if path is not None:
        spec.loader.exec_module(m)
    return m


_loaded = None


def load():
    """Return the reference ``tidy3d.plugins.mode.solver`` module (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    if "tidy3d" in sys.modules and not getattr(sys.modules["tidy3d"], "_b200_shim", False):
        raise RuntimeError("a real 'tidy3d' is already imported; refusing to shadow it")
    for pkg in ("tidy3d", "tidy3d.components", "tidy3d.plugins", "tidy3d.plugins.mode"):
        _mod(pkg)._b200_shim = True
    _mod("tidy3d.constants", f"{_REF}/constants.py")
    base = _mod("tidy3d.components.base")
    base.Tidy3dBaseModel = type("Tidy3dBaseModel", (), {})
    typ = _mod("tidy3d.components.types")
    typ.EpsSpecType = typ.ModeSolverType = str
    typ.Numpy = np.ndarray
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _mod("tidy3d.plugins.mode.derivatives", f"{_REF}/plugins/mode/derivatives.py")
        _mod("tidy3d.plugins.mode.transforms", f"{_REF}/plugins/mode/transforms.py")
        _loaded = _mod("tidy3d.plugins.mode.solver", f"{_REF}/plugins/mode/solver.py")
    return _loaded


def compute_modes(eps_cross, coords, freq, mode_spec, **kw):
    """Run the reference ``compute_modes`` (solver.py:941) on copies of the inputs."""
    ref = load()
    eps = [np.array(e, dtype=complex, copy=True) for e in eps_cross]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref.compute_modes(eps, [np.array(c, float) for c in coords], freq, mode_spec, **kw)
