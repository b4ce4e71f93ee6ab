"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference numerics.

Imports ``tidy3d/plugins/mode/{solver,derivatives,transforms}.py`` straight from
``/root/reference`` under a stub package (full ``import tidy3d`` needs xarray/shapely/... which
this image lacks; recipe from SURVEY.md Appendix C).  Where ``/root/reference`` does not exist (the
GPU box) the same four files are loaded from their byte-compiled form under ``oracle/_ref``
(``oracle/build_ref.py``: sourceless ``.pyc``, nothing of the reference is copied into the repository).
Used to (a) pin ``oracle/restatement.py``, (b) generate the golden fixtures under ``tests/golden``
(``tests/golden/make_golden.py``) and (c) as the CPU arm of ``bench.py`` (``cpu_baseline.kind ==
"reference"``).

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
_BUILT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "tidy3d")


def source_available() -> bool:
    return os.path.isfile(os.path.join(_REF, "plugins", "mode", "solver.py"))


def built_available() -> bool:
    """``oracle/_ref`` holds byte code made by THIS interpreter version (oracle/build_ref.py)."""
    path = os.path.join(_BUILT, "plugins", "mode", "solver.pyc")
    try:
        with open(path, "rb") as f:
            return f.read(4) == importlib.util.MAGIC_NUMBER
    except OSError:
        return False


def available() -> bool:
    return source_available() or built_available()


def origin() -> str:
    """Where the reference code comes from: "source" (/root/reference), "built" (oracle/_ref) or "none"."""
    return "source" if source_available() else "built" if built_available() else "none"


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
        raise RuntimeError(f"reference tree not found under {REF_ROOT} and oracle/_ref is not built")
    root, ext = (_REF, ".py") if source_available() else (_BUILT, ".pyc")
    if "tidy3d" in sys.modules and not getattr(sys.modules["tidy3d"], "_b200_shim", False):
        raise RuntimeError("a real 'tidy3d' is already imported; refusing to shadow it")
    for pkg in ("tidy3d", "tidy3d.components", "tidy3d.plugins", "tidy3d.plugins.mode"):
        _mod(pkg)._b200_shim = True
    _mod("tidy3d.constants", f"{root}/constants{ext}")
    base = _mod("tidy3d.components.base")
    base.Tidy3dBaseModel = type("Tidy3dBaseModel", (), {})
    typ = _mod("tidy3d.components.types")
    typ.EpsSpecType = typ.ModeSolverType = str
    typ.Numpy = np.ndarray
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _mod("tidy3d.plugins.mode.derivatives", f"{root}/plugins/mode/derivatives{ext}")
        _mod("tidy3d.plugins.mode.transforms", f"{root}/plugins/mode/transforms{ext}")
        _loaded = _mod("tidy3d.plugins.mode.solver", f"{root}/plugins/mode/solver{ext}")
    return _loaded


def compute_modes(eps_cross, coords, freq, mode_spec, **kw):
    """Run the reference ``compute_modes`` (solver.py:941) on copies of the inputs."""
    ref = load()
    eps = [np.array(e, dtype=complex, copy=True) for e in eps_cross]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref.compute_modes(eps, [np.array(c, float) for c in coords], freq, mode_spec, **kw)
