"""End-to-end case (SURVEY 8 rows f-2 -> a -> f-1): a 3-D scene, a mode plane, a short frequency sweep -- the whole of
``ModeSolver.data_raw`` as the UNMODIFIED reference computes it (oracle/ref_solver.py: permittivity sampling, compute_modes,
gauge, grid correction, flux normalisation, mode tracking all by the reference's own code) is stored in
tests/golden/e2e_strip.npz by tests/golden/make_e2e_golden.py, together with the frequency-independent description of the
plane the f-2 seam produces (``plugin.section_of``) and the media tensors per frequency.

``check(solve)`` runs the product-side chain on that description -- ``solve(problems, post)`` is the device call
(``compute_modes_batch`` with on-device post-processing) in tests/test_gpu_zzz_end_to_end.py and its CPU emulation from the
restatements (oracle/) in tests/test_end_to_end_cpu.py -- and compares with the reference's data.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "e2e_strip.npz")
GOLDEN_SYM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "e2e_strip_sym.npz")  # half domain, PMC wall at x = 0 (CPU twin only)
C0 = 2.99792458e14
NUM_MODES, NF = 2, 4
NORMAL_PRIMAL, NORMAL_DUAL = (-0.045, 0.055), (-0.095, 0.005)  # the plane (z = 0) sits between grid points of a coarse normal grid


def reference_solver(track, symmetric=False):
    """The reference-made ModeSolver of the case (build container only).  ``symmetric``: the right half of a mirror-symmetric
    variant of the scene with a symmetry wall (PMC, +1) at x = 0."""
    from oracle import ref_sections as RS
    from oracle import ref_solver as RSV

    rng = np.random.default_rng(3)
    x = np.cumsum(np.r_[0.0 if symmetric else -1.0, rng.uniform(0.045, 0.055, 20 if symmetric else 40)])  # slightly non-uniform
    y = np.cumsum(np.r_[-0.8, rng.uniform(0.045, 0.055, 32)])
    edges = [x, y, np.array([-0.02, 0.02])]
    structures = [
        (RS.geometry("Box", center=(0.0 if symmetric else 0.02, 0.01, 0.0), size=(0.5, 0.22, 10.0)), RS.TensorMedium(3.48**2 * np.eye(3), 0.02)),  # dispersive core
        (RS.geometry("Box", center=(0.0, -0.3, 0.0), size=(10.0, 0.36, 10.0)), RS.TensorMedium(np.diag([2.1, 2.15, 2.05]), -0.01)),  # anisotropic substrate
    ]
    if not symmetric:
        structures.append((RS.geometry("Sphere", center=(0.5, 0.2, 0.05), radius=0.17), RS.TensorMedium(2.0**2 * np.eye(3))))  # breaks every symmetry
    spec = RSV.ModeSpec(num_modes=NUM_MODES, track_freq=track, group_index_step=0, precision="double")
    freqs = C0 / np.linspace(1.5, 1.6, NF)
    return RSV.mode_solver(edges, 2, structures, RS.TensorMedium(1.44**2 * np.eye(3)), freqs, spec, colocate=False,
                           symmetry=(1, 0, 0) if symmetric else (0, 0, 0), normal_primal=NORMAL_PRIMAL, normal_dual=NORMAL_DUAL)


def check(solve, tol_n=1e-6, tol_field=2e-4, tol_overlap=2e-4, symmetric=False):
    """``solve(problems, post) -> (results, info)`` like ``compute_modes_batch(problems, post=post, return_info=True)``."""
    from tidy3d_b200 import postprocess as PP
    from tidy3d_b200.sections import Medium, Section

    z = np.load(GOLDEN_SYM if symmetric else GOLDEN)
    freqs = list(z["freqs"])
    tables = z["media"]  # (F, nmedia, 3, 3)

    def medium(k):
        return Medium(lambda f, k=k: tables[int(np.argmin(np.abs(np.array(freqs) - f))), k])

    media = [medium(k) for k in range(tables.shape[1])]
    sec = Section(background=media[0], media=media, site_medium=z["site_medium"])
    spec = PP_mode_spec()
    gc = PP.grid_correction_table(NORMAL_PRIMAL, NORMAL_DUAL, 0.0)
    coords = [z["x"], z["y"]]
    problems = [dict(section=sec, coords=coords, freq=float(f), mode_spec=spec, grid_correction=gc, symmetry=(1, 0) if symmetric else (0, 0)) for f in freqs]
    results, info = solve(problems, ("gauge", "normalize", "flux", "overlaps"))
    worst = dict(n=0.0, field=0.0, overlap=0.0)
    fields, n_c = [], []
    for i, (f, n, eps_spec) in enumerate(results):
        assert eps_spec == "diagonal"
        worst["n"] = max(worst["n"], float(np.abs(n - z["n_raw"][i]).max()))
        ref = z["normalized_yee"][:, :, :, :, i, :]
        got = np.asarray(f)[:, :, :, :, 0, :]
        worst["field"] = max(worst["field"], float(np.abs(got - ref).max() / np.abs(ref).max()))
        if i:
            worst["overlap"] = max(worst["overlap"], float(np.abs(np.asarray(info[i]["overlap_prev"]) - z["outer_next"][i - 1]).max()))
        fields.append(np.asarray(f))
        n_c.append(np.asarray(n))
    assert worst["n"] <= tol_n and worst["field"] <= tol_field and worst["overlap"] <= tol_overlap, worst
    sorting, phase, _ = PP.overlap_sort([None] + [np.asarray(info[i]["overlap_prev"]) for i in range(1, len(freqs))], track_freq="central")
    assert np.array_equal(sorting, z["sorting"]) and np.abs(phase - z["phase"]).max() <= 10 * tol_overlap
    n_sorted, f_sorted = PP.apply_sorting(n_c, fields, sorting, phase)
    final = np.stack([f[:, :, :, :, 0, :] for f in f_sorted], axis=-2)
    assert np.abs(final - z["final_yee"]).max() / np.abs(z["final_yee"]).max() <= 20 * tol_field
    assert np.abs(np.array(n_sorted) - z["final_n_complex"]).max() <= tol_n
    return worst


def PP_mode_spec():
    from tidy3d_b200 import workloads as W

    return W.ModeSpecLike(num_modes=NUM_MODES, precision="double")
