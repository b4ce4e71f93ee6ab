"""Generates tests/golden/e2e_strip.npz: ``ModeSolver.data_raw`` of the case in tests/e2e_case.py, computed end to end by the
UNMODIFIED reference's own code (oracle/ref_solver.py), plus the frequency-independent description of the plane that
``tidy3d_b200.plugin.section_of`` makes of the same ModeSolver.

Run in the build container (needs /root/reference):   python tests/golden/make_e2e_golden.py
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import e2e_case as E  # noqa: E402


def main():
    for symmetric in (False, True):
        one(symmetric)


def one(symmetric):
    from oracle import ref_post as RP
    from oracle import ref_solver as RSV

    import tidy3d_b200.plugin as plugin

    warnings.simplefilter("ignore")
    ms = E.reference_solver(track=None, symmetric=symmetric)
    sec = plugin.section_of(ms)
    freqs = np.array(ms.freqs)
    out = dict(x=ms.simulation.edges[0], y=ms.simulation.edges[1], freqs=freqs, site_medium=sec.site_medium,
               media=np.array([[m.tensor(f) for m in sec.media] for f in freqs]))
    yee = ms._data_on_yee_grid()
    out["n_raw"] = yee.n_complex.values
    norm = E.reference_solver(track=None, symmetric=symmetric).data_raw  # gauge + flux normalisation with the grid-correction factors, no tracking
    out["normalized_yee"] = RP.packed(norm)
    outers = []
    for i in range(len(freqs) - 1):
        a = norm._isel(f=[i])
        b = norm._isel(f=[i + 1])._assign_coords(f=[freqs[i]])
        outers.append(a.outer_dot(b).to_numpy()[0])
    out["outer_next"] = np.array(outers)
    mod = RSV.module()
    rec = {}
    orig = mod.ModeSolverData._reorder_modes

    def spy(self, sorting, phase, track_freq):
        rec.update(sorting=np.array(sorting), phase=np.array(phase))
        return orig(self, sorting=sorting, phase=phase, track_freq=track_freq)

    mod.ModeSolverData._reorder_modes = spy
    try:
        final = E.reference_solver(track="central", symmetric=symmetric).data_raw
    finally:
        mod.ModeSolverData._reorder_modes = orig
    out.update(sorting=rec["sorting"], phase=rec["phase"], final_yee=RP.packed(final), final_n_complex=final.n_complex.values)
    path = E.GOLDEN_SYM if symmetric else E.GOLDEN
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1024:.0f} KB", "n_raw", out["n_raw"][0], "sorting", out["sorting"].tolist())


if __name__ == "__main__":
    main()
