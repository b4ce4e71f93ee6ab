"""CPU: the product-side chain of tests/e2e_case.py with the device EMULATED by the restatements of oracle/ (sections ->
numerics at tol 1e-12 -> gauge / grid correction / flux normalisation / overlaps): reproduces ``ModeSolver.data_raw`` of the
unmodified reference end to end (tests/golden/e2e_strip.npz).  The GPU twin (tests/test_gpu_zzz_end_to_end.py) runs the very
same ``check`` with the real handle; here everything of ``compute_modes_batch`` runs for real -- argument handling, the ctypes
packing of the section / grid-correction / post flags, the assembly of results and info dicts -- except
``Handle.solve_batch``, which checks the packed structs (and lets the library's HOST mirror of the device rasteriser set every
problem up) and answers with the restated chain.  This fixes the tolerances independently of a GPU: the reference solves with
ARPACK at tol = float32 eps (solver.py:745), so its n_eff carry ~1e-8 and its fields ~1e-6."""
import ctypes as C
import threading
import warnings

import numpy as np

from tests import e2e_case as E


class EmulatedHandle:
    """``_cabi.Handle`` without a device: same ``solve_batch`` contract, results from oracle/."""

    def __init__(self, cabi, problems, nmedia=4):
        self.cabi, self.problems, self.lock, self.nmedia = cabi, problems, threading.Lock(), nmedia
        self.last_flux, self.last_te, self.last_overlaps = [], [], []
        self.setups = 0

    def last_error(self):
        return ""

    def solve_batch(self, packed, want_fields=True, fields_ptrs=None, want_flux=False, want_overlaps=False):
        from oracle import postprocess as OP
        from oracle import restatement as R
        from oracle import sections as OS
        from tidy3d_b200 import postprocess as PP

        assert want_fields and want_flux and want_overlaps and fields_ptrs is None and len(packed) == len(self.problems)
        results = (self.cabi.Result * len(packed))()
        fields, ncs, prev = [], [], None
        self.last_flux, self.last_te, self.last_overlaps = [], [], []
        for i, (pk, p) in enumerate(zip(packed, self.problems)):
            st = pk.struct
            assert st.post == 3 and not st.eps and st.section and st.section.contents.nrect == 0 and st.section.contents.nmedia == self.nmedia
            assert (st.symmetry[0], st.symmetry[1]) == tuple(p["symmetry"])
            assert [st.grid_correction[k] for k in range(8)] == list(p["grid_correction"]) and not st.plane_bounds
            eps = OS.eps_on_grid(p["section"], p["coords"], p["freq"])
            # the library's host mirror of the device rasteriser sets this very struct up like the sampled array
            outs = []
            for q in (pk, self.cabi.PackedProblem(eps, p["coords"], p["freq"], p["mode_spec"], p["symmetry"])):
                f = np.zeros((6, q.nx * q.ny), complex)
                flags, sigma = (C.c_int * 4)(), np.zeros(2)
                assert self.cabi.lib().b200ms_debug_setup(C.byref(q.struct), self.cabi._ptr(sigma), flags, None, None, None, None, self.cabi._ptr(f.view(float))) == 0
                outs.append((list(flags), sigma.copy(), f))
            assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
            self.setups += 1
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                f, n, spec = R.compute_modes(eps, p["coords"], p["freq"], p["mode_spec"], symmetry=p["symmetry"], tol=1e-12)
            corr = PP.grid_correction_factors(n, p["freq"], p["grid_correction"], 0.0, "+")
            g, _ = OP.gauge(f)
            fn, fl = OP.normalize(g, p["coords"], p["symmetry"], correction=corr)
            self.last_flux.append(fl)
            self.last_te.append(OP.pol_fraction(g, p["coords"], p["symmetry"]))
            self.last_overlaps.append(OP.dot(prev[0], fn, p["coords"], p["symmetry"], correction_a=prev[1], correction_b=corr) if prev
                                      else np.zeros((n.size, n.size), complex))
            prev = (fn, corr)
            fields.append(fn)
            ncs.append(n)
            results[i].status, results[i].converged, results[i].eps_spec = self.cabi.OK, 1, {v: k for k, v in self.cabi.SPEC_NAMES.items()}[spec]
        return self.cabi.OK, fields, ncs, results


def test_restated_chain_reproduces_the_reference_mode_solver_data(built_lib):
    from tidy3d_b200 import compute_modes_batch

    state = {}

    def emulated_device(problems, post):
        state["h"] = EmulatedHandle(built_lib, problems)
        return compute_modes_batch(problems, handle=state["h"], post=post, return_info=True)

    worst = E.check(emulated_device)
    assert state["h"].setups == E.NF
    assert worst["n"] < 1e-7 and worst["field"] < 1e-4 and worst["overlap"] < 1e-4, worst


def test_restated_chain_with_a_symmetry_wall(built_lib):
    """The same for the right half of a mirror-symmetric scene with a PMC wall at x = 0 (CPU twin only): half-domain solve,
    symmetry-expanded colocation, doubled integrals, tracking -- against the reference's own ModeSolver.data_raw."""
    from tidy3d_b200 import compute_modes_batch

    state = {}

    def emulated_device(problems, post):
        state["h"] = EmulatedHandle(built_lib, problems, nmedia=3)
        return compute_modes_batch(problems, handle=state["h"], post=post, return_info=True)

    worst = E.check(emulated_device, symmetric=True)
    assert worst["n"] < 1e-7 and worst["field"] < 1e-4 and worst["overlap"] < 1e-4, worst
