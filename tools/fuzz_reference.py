#!/usr/bin/env python
"""Fuzz the restatements of oracle/ -- the numerics (oracle/restatement.py), f-1 and f-2 (oracle/postprocess.py, oracle/sections.py + plugin.section_of + the library's host mirror
of the device rasteriser) against the UNMODIFIED reference's own code (oracle/ref_post.py, oracle/ref_sections.py).  Build
container only.     python tools/fuzz_reference.py [seed] [cases]"""
import ctypes as C
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fuzz_post(rng, n):
    from tests import post_cases as PC

    bad = 0
    for _ in range(n):
        case = PC.random_case(rng)
        PC.CASES["_fuzz"] = case
        c = PC.inputs("_fuzz")
        err = PC.compare(PC.reference_results(c), PC.oracle_results(c), skip=PC.skipped_keys(c))
        if max(err.values()) > 1e-10:
            bad += 1
            print("post MISMATCH", case, {k: v for k, v in err.items() if v > 1e-10})
    return bad


def fuzz_sections(rng, n):
    import tidy3d_b200.plugin as plugin
    import tidy3d_b200.sections as S
    from oracle import ref_sections as RS
    from oracle import sections as OS
    from tidy3d_b200 import _cabi as L

    bad = 0
    spec = types.SimpleNamespace(num_modes=1)
    for t in range(n):
        normal = int(rng.integers(0, 3))
        cells = [int(rng.integers(3, 14)) for _ in range(3)]
        edges = [np.unique(np.round(rng.uniform(-1, 1, k + 1), 3)) if rng.random() < 0.5 else np.linspace(-1, 1, k + 1) for k in cells]
        edges[normal] = np.array([-0.01, 0.01])
        structures = []
        for _ in range(int(rng.integers(0, 6))):
            kind = str(rng.choice(["Box", "Sphere", "Cylinder"]))
            c = tuple(np.round(rng.uniform(-0.6, 0.6, 3), 2))
            if kind == "Box":
                kw = dict(center=c, size=tuple(np.round(rng.uniform(0.1, 1.2, 3), 2)))
            elif kind == "Sphere":
                kw = dict(center=c, radius=float(np.round(rng.uniform(0.1, 0.7), 2)))
            else:
                kw = dict(center=c, radius=float(np.round(rng.uniform(0.1, 0.6), 2)), length=float(np.round(rng.uniform(0.1, 1.5), 2)), axis=int(rng.integers(0, 3)))
            structures.append((RS.geometry(kind, **kw), RS.TensorMedium(rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3)), float(rng.uniform(-0.05, 0.05)))))
        ms = RS.solver(normal, edges, structures, RS.TensorMedium(2.0 * np.eye(3)))
        sec = plugin.section_of(ms)
        coords = [e for a, e in enumerate(edges) if a != normal]
        for f in (1.9e14, 2.2e14):  # the f-2 seam: section_of + rasteriser == the reference's _solver_eps, bit for bit
            want, got = np.array(ms._solver_eps(f)), OS.eps_on_grid(sec, coords, f)
            if want.shape != got.shape or not np.array_equal(want, got):
                bad += 1
                print("sections MISMATCH scene", t, normal, cells)
        # primitive cuts against the reference geometry on a lattice with many rim points
        sx = sy = np.round(np.linspace(-1, 1, 41), 12)
        cx, cy = np.round(rng.uniform(-0.5, 0.5, 2), 2)
        r, dz = float(np.round(rng.uniform(0.05, 0.6), 2)), float(np.round(rng.uniform(0, 0.3), 2))
        pairs = [(RS.geometry("Box", center=(cx, cy, 0.0), size=(2 * r, r, 1.0)), S.Rect((cx, cy), (2 * r, r))),
                 (RS.geometry("Sphere", center=(cx, cy, dz), radius=r), S.Disc((cx, cy), r, dz)),
                 (RS.geometry("Cylinder", center=(cx, cy, 0.1), radius=r, length=0.5, axis=2), S.Disc((cx, cy), r, 0.0))]
        for g, shape in pairs:
            if not np.array_equal(g.inside_meshgrid(sx, sy, np.array([0.0]))[:, :, 0], OS.inside(shape, sx, sy)):
                bad += 1
                print("primitive MISMATCH", type(shape).__name__, cx, cy, r, dz)
        # the library's host mirror of the device rasteriser on the same primitives
        sec2 = S.Section(background=S.Medium(1.7), structures=[(shape, S.Medium(tuple(rng.uniform(2, 13, 3)))) for _, shape in pairs])
        x = np.round(np.linspace(-1, 1, int(rng.integers(5, 40))), 12)
        y = np.round(np.linspace(-1, 1, int(rng.integers(5, 40))), 12)
        eps = OS.eps_on_grid(sec2, [x, y], 2e14)
        outs = []
        for pk in (L.PackedProblem(None, [x, y], 2e14, spec, section=sec2), L.PackedProblem(eps, [x, y], 2e14, spec)):
            f = np.zeros((6, pk.nx * pk.ny), complex)
            flags, sigma = (C.c_int * 4)(), np.zeros(2)
            assert L.lib().b200ms_debug_setup(C.byref(pk.struct), L._ptr(sigma), flags, None, None, None, None, L._ptr(f.view(float))) == 0
            outs.append((list(flags), sigma.copy(), f))
        if not (outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])):
            bad += 1
            print("library rasteriser MISMATCH", t)
    return bad


def fuzz_numerics(rng, n):
    """oracle/restatement.py against the unmodified reference's compute_modes (oracle/ref_shim.py) on random small problems:
    n_eff to 1e-9 (double) / 2e-6 (single: float32 matrices), fields by separate E / H overlaps with a common phase."""
    from oracle import ref_shim
    from oracle import restatement as R
    from tests.helpers import mode_overlaps, well_separated
    from tidy3d_b200 import workloads as W

    ref = ref_shim.load()
    bad = 0
    for t in range(n):
        nx, ny = int(rng.choice([1, 12, 17, 24])), int(rng.choice([1, 10, 16, 21]))
        if nx == 1 and ny == 1:
            nx = 15
        x = np.cumsum(np.r_[rng.uniform(-1, 0), rng.uniform(0.04, 0.08, nx)])
        y = np.cumsum(np.r_[rng.uniform(-1, 0), rng.uniform(0.04, 0.08, ny)])
        xm, ym = 0.5 * (x[:-1] + x[1:]), 0.5 * (y[:-1] + y[1:])
        blob = np.exp(-((xm[:, None] - xm.mean()) / 0.25) ** 2 - ((ym[None, :] - ym.mean()) / 0.2) ** 2)
        base = 2.1 + 9 * blob + 0.1 * rng.random((nx, ny))
        kind = int(rng.integers(0, 3))
        eps = [np.zeros((nx, ny), complex) for _ in range(9)]
        for k, s in zip((0, 4, 8), (1.0, 1.05, 0.95)):
            eps[k] = base * s + 0j
        if kind == 1:
            for k in (0, 4, 8):
                eps[k] = eps[k] + 1j * 0.05 * blob
        if kind == 2:
            eps[1] = eps[3] = 0.05 * base + 0j
        npml = (int(rng.choice([0, 0, 3])) if nx > 12 else 0, int(rng.choice([0, 0, 3])) if ny > 12 else 0)
        sym = (int(rng.choice([0, 0, 1, -1])) if nx > 1 else 0, int(rng.choice([0, 0, 1, -1])) if ny > 1 else 0)
        spec = W.ModeSpecLike(num_modes=int(rng.integers(1, 3)), num_pml=npml, target_neff=None if rng.random() < 0.4 else float(rng.uniform(2.0, 3.0)),
                              precision=str(rng.choice(["single", "double"])))
        r = rng.random()
        if r < 0.25 and nx > 1 and ny > 1:
            spec.bend_radius, spec.bend_axis = float(rng.choice([-1, 1]) * rng.uniform(4, 10)), int(rng.integers(0, 2))
        elif r < 0.5:
            spec.angle_theta, spec.angle_phi = float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-1, 1))
        direction, freq = str(rng.choice(["+", "-"])), W.C_0 / float(rng.uniform(1.3, 1.6))
        try:
            f0, n0, s0 = ref.compute_modes(eps_cross=eps, coords=[x, y], freq=freq, mode_spec=spec, symmetry=sym, direction=direction)
        except Exception:  # noqa: BLE001  (ARPACK did not converge on this random problem: nothing to compare)
            continue
        f1, n1, s1 = R.compute_modes(eps, [x, y], freq, spec, symmetry=sym, direction=direction)
        ok = well_separated(n0, 1e-3)
        ov = mode_overlaps(f1.astype(complex), f0.astype(complex))[ok]
        if s0 != s1 or f0.dtype != f1.dtype or np.abs(n0 - n1).max() > (1e-9 if spec.precision == "double" else 2e-6) or (ov.size and ov.min() < 1 - 1e-4):
            bad += 1
            print("numerics MISMATCH", t, nx, ny, kind, npml, sym, vars(spec), direction)
    return bad


if __name__ == "__main__":
    warnings.simplefilter("ignore")
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    rng = np.random.default_rng(seed)
    b1 = fuzz_post(rng, n)
    b2 = fuzz_sections(rng, n)
    b3 = fuzz_numerics(rng, n)
    print(f"seed {seed}: {n} post-processing cases, {b1} mismatches; {n} scenes, {b2} mismatches; {n} eigenproblems, {b3} mismatches")
