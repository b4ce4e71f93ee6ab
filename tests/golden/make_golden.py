"""Generate golden fixtures by running the UNMODIFIED reference (oracle/ref_shim.py).

Build-container only (needs /root/reference).  Usage:
    python tests/golden/make_golden.py [case ...]        # default: all cases not yet on disk
Each fixture ``<case>.npz`` holds n_complex at the reference tolerance (TOL_EIGS = fp_eps) and at
TOL_EIGS = 1e-12 ("tight"), eps_spec, the tight mode fields when the grid is small, and always a
per-mode field signature (|E|,|H| component norms) that is phase independent.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from tests.golden.cases import CASES, resolve_kwargs  # noqa: E402


def signature(fields):
    """(M, 6) array of component 2-norms per mode after unit-normalising the in-plane E part."""
    f = fields.reshape(6, -1, fields.shape[-1])
    scale = np.sqrt((np.abs(f[:2]) ** 2).sum(axis=(0, 1)))
    return (np.sqrt((np.abs(f) ** 2).sum(axis=1)) / scale).T


def run(name):
    factory, kw, store = CASES[name]
    wl = factory()
    kw = resolve_kwargs(wl, kw)
    ref = ref_shim.load()
    out = {}
    single = getattr(wl.mode_spec, "precision", "double") == "single"
    for tag, tol in (("ref", None), ("tight", 1e-12)):
        if single and tag == "tight":
            # a single-precision ARPACK run cannot reach 1e-12: the "tight" companion of a single case is the reference
            # in DOUBLE precision at TOL_EIGS = 1e-12 (what the single result approximates)
            wl.mode_spec.precision = "double"
        old = ref.TOL_EIGS
        if tol is not None:
            ref.TOL_EIGS = tol
        t0 = time.time()
        try:
            fields, n_complex, spec = ref_shim.compute_modes(wl.eps_cross, wl.coords, wl.freqs[0], wl.mode_spec, **kw)
        except Exception as e:  # noqa: BLE001  (scipy ArpackNoConvergence on the pathological rand10 input at 1e-12)
            if tag != "tight":
                raise
            print(name, "tight run failed:", type(e).__name__, "-> n_tight := n_ref", flush=True)
            out["tight_failed"] = True
            out["n_tight"], out["sig_tight"], out["sec_tight"] = out["n_ref"], out["sig_ref"], 0.0
            ref.TOL_EIGS = old
            continue
        finally:
            ref.TOL_EIGS = old
        out[f"n_{tag}"] = n_complex
        out[f"sig_{tag}"] = signature(fields)
        out[f"sec_{tag}"] = time.time() - t0
        out["spec"] = spec
        if store and tag == "tight" and not single:
            out["fields_tight"] = fields if fields.size < 2e5 else fields.astype(np.complex64)  # keep the fixtures small
        if store and single and tag == "ref":
            out["fields_ref"] = fields  # complex64, the reference's own single-precision fields
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, out["spec"], out["n_ref"], "max|n_ref-n_tight|=%.2e" % np.abs(out["n_ref"] - out["n_tight"]).max(),
          "%.1fs/%.1fs" % (out["sec_ref"], out["sec_tight"]), flush=True)  # fmt: skip


def run_relative():
    """Fixture for the relative solver: basis = reference modes at 1.56 um, solve at 1.55 um in that span."""
    from tests.golden.cases import relative_case
    from tidy3d_b200 import workloads as W

    wl = relative_case()
    basis, nb, _ = ref_shim.compute_modes(wl.eps_cross, wl.coords, W.C_0 / wl.extra["basis_lam"], wl.mode_spec)
    sbf = np.concatenate((basis[0], basis[1]), axis=0)  # (6, Nx, Ny, 1, M) like _postprocess_solver_fields_inverse
    fields, n_complex, spec = ref_shim.compute_modes(wl.eps_cross, wl.coords, wl.freqs[0], wl.mode_spec, solver_basis_fields=sbf)
    np.savez_compressed(os.path.join(HERE, "relative_48.npz"), basis=sbf, n_ref=n_complex, n_tight=n_complex, spec=spec,
                        fields_tight=fields, sig_ref=signature(fields))
    print("relative_48", spec, n_complex, flush=True)


if __name__ == "__main__":
    if sys.argv[1:] == ["relative"]:
        run_relative()
        sys.exit(0)
    names = sys.argv[1:] or [n for n in CASES if not os.path.exists(os.path.join(HERE, n + ".npz"))]
    for n in names:
        run(n)
