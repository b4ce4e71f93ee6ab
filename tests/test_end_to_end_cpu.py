"""CPU: the product-side chain of tests/e2e_case.py with the device call EMULATED by the restatements of oracle/ (sections ->
numerics at tol 1e-12 -> gauge / grid correction / flux normalisation / overlaps): reproduces ``ModeSolver.data_raw`` of the
unmodified reference end to end (tests/golden/e2e_strip.npz).  The GPU twin (tests/test_gpu_zzz_end_to_end.py) runs the very
same ``check`` with ``compute_modes_batch`` as the device call; this file fixes the tolerances independently of a GPU:
the reference solves with ARPACK at tol = float32 eps (solver.py:745), so its n_eff carry ~1e-8 and its fields ~1e-5."""
import warnings

import numpy as np

from tests import e2e_case as E


def emulated_device(problems, post):
    from oracle import postprocess as OP
    from oracle import restatement as R
    from oracle import sections as OS
    from tidy3d_b200 import postprocess as PP

    assert set(post) == {"gauge", "normalize", "flux", "overlaps"}
    results, info, prev = [], [], None
    for p in problems:
        eps = OS.eps_on_grid(p["section"], p["coords"], p["freq"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f, n, spec = R.compute_modes(eps, p["coords"], p["freq"], p["mode_spec"], tol=1e-12)
        corr = PP.grid_correction_factors(n, p["freq"], p["grid_correction"], 0.0, "+")
        g, _ = OP.gauge(f)
        fn, fl = OP.normalize(g, p["coords"], correction=corr)
        d = dict(flux=fl, te_fraction=OP.pol_fraction(g, p["coords"]))
        if prev is not None:
            d["overlap_prev"] = OP.dot(prev[0], fn, p["coords"], correction_a=prev[1], correction_b=corr)
        prev = (fn, corr)
        results.append((fn, n, spec))
        info.append(d)
    return results, info


def test_restated_chain_reproduces_the_reference_mode_solver_data():
    worst = E.check(emulated_device)
    assert worst["n"] < 1e-7 and worst["field"] < 1e-4 and worst["overlap"] < 1e-4, worst
    print(worst)
