"""DEV TOOL (round 2, session 2): the pair-marching stencil kernel (csrc/march2.cuh) against the one-column marching kernel.

  eq     kernel by kernel: every mode on levels 0..2 of several problems, stencil_pair = 1, 2 against stencil_pair = 0
         (same arithmetic expression for expression, so the difference should be rounding-order noise or exactly zero)
  time   b200ms_bench_stencil (fp32 stored-diagonal sweep) at B = 64 / 32 for stencil_pair = 0, 1, 2 and several rows-per-CTA
  solve  whole 64-problem batches of the headline sweep with the three settings (times, iteration counts, |dn|)
Run under gpurun; output archived in profiles/r02_pair_kernel.txt.
"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo")
from tests.golden.cases import CASES  # noqa: E402
from tidy3d_b200 import _cabi, compute_modes_batch  # noqa: E402
from tidy3d_b200 import workloads as W  # noqa: E402

L = _cabi.lib()
REF = dict(eig_tol=1.1920928955078125e-07, inner_tol=1e-8)
rng = np.random.default_rng(11)


def dev_apply(H, pk, level, mode, x, rhs):
    y = np.zeros(x.size, complex)
    xa = np.ascontiguousarray(x.ravel().astype(complex))
    ra = np.ascontiguousarray(rhs.ravel().astype(complex))
    rc = L.b200ms_debug_apply(H._h, C.byref(pk.struct), level, mode, _cabi._ptr(xa.view(float)), _cabi._ptr(ra.view(float)), _cabi._ptr(y.view(float)))
    assert rc == 0, (rc, H.last_error())
    return y


def shapes_of(pk):
    sh = (C.c_int * 40)()
    n = L.b200ms_debug_hierarchy(C.byref(pk.struct), None, 20, sh)
    return [(sh[2 * i], sh[2 * i + 1]) for i in range(n)]


def eq():
    def rect(nx, ny):
        eps, coords = W.strip_eps(nx, ny)
        return W.Workload(name=f"strip_{nx}x{ny}", eps_cross=W._iso(eps), coords=coords, freqs=np.array([W.C_0 / 1.55]),
                          mode_spec=W.ModeSpecLike(num_modes=2, precision="double"))

    cases = [("headline512", W.headline(nf=4, n=512), {}), ("headline256", W.headline(nf=4, n=256), {})]
    for name in ("nonuniform_56", "c1_64_sym_pmc_pec"):
        fac, kw, _ = CASES[name]
        cases.append((name, fac(), kw))
    # odd row counts, a width that is not a multiple of 64 columns, and one that needs two strips (600 columns = 300 pairs)
    cases += [("strip_150x202", rect(150, 202), {}), ("strip_100x600", rect(100, 600), {}), ("strip_70x1100", rect(70, 1100), {})]
    worst = 0.0
    for name, wl, kw in cases:
        pk = _cabi.PackedProblem(wl.eps_cross, wl.coords, wl.freqs[0], wl.mode_spec, kw.get("symmetry", (0, 0)), kw.get("direction", "+"))
        shp = shapes_of(pk)
        print(f"== {name}: levels {shp[:4]}", flush=True)
        hs = {sp: _cabi.Handle(**{**REF, "stencil_pair": sp}) for sp in VARIANTS}
        hs[9] = _cabi.Handle(**{**REF, "stencil_pair": VARIANTS[-1], "stencil_pair_rows": 9})
        for lvl, (nx, ny) in enumerate(shp[:3]):
            if nx < 32:
                continue
            x = rng.standard_normal((2, nx, ny))
            rhs = rng.standard_normal((2, nx, ny))
            for mode, nm in ((16, "apply"), (17, "resid"), (18, "jacobi_d"), (20, "jacobi_d0")):
                ys = {sp: dev_apply(h, pk, lvl, mode, x, rhs) for sp, h in hs.items()}
                sc = np.abs(ys[0]).max()
                d = {sp: np.abs(ys[sp] - ys[0]).max() / sc for sp in hs if sp != 0}
                worst = max(worst, *d.values())
                print(f"   level {lvl} ({nx}x{ny}) {nm:10s} max|y| {sc:.3e}  rel diff " + " ".join(f"[{sp}] {v:.1e}" for sp, v in d.items()), flush=True)
        for h in hs.values():
            h.close()
    print(f"EQ worst relative difference {worst:.2e}  ({'OK' if worst < 2e-5 else 'MISMATCH'})", flush=True)


def timing():
    wl = W.headline(nf=4, n=512)
    pk = _cabi.PackedProblem(wl.eps_cross, wl.coords, wl.freqs[0], wl.mode_spec)
    ms, byts = C.c_double(), C.c_double()
    for nb in (64, 32):
        for opts in [dict(stencil_pair=sp) for sp in VARIANTS] + [dict(stencil_pair=sp, stencil_pair_rows=r) for sp in VARIANTS[1:] for r in ROWS]:
            h = _cabi.Handle(**{**REF, **opts})
            rc = L.b200ms_bench_stencil(h._h, C.byref(pk.struct), nb, 1, 50, 0, None, None, C.byref(ms), C.byref(byts))
            assert rc == 0, h.last_error()
            print(f"## sweep B={nb} {opts}: {ms.value * 1e3:.1f} us  {byts.value / (ms.value * 1e-3) / 1e9:.0f} GB/s", flush=True)
            h.close()


def solve():
    g = np.load("/root/repo/tests/golden/headline_512_f0.npz")
    wl = W.headline(nf=256, n=512)
    for nb, opts in [(nb, dict(stencil_pair=sp)) for nb in (64, 16) for sp in VARIANTS]:
        h = _cabi.Handle(**{**REF, "max_batch": 64, **opts})
        probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in wl.freqs[:nb]]
        try:
            for rep in range(2):
                t0 = time.time()
                out, info = compute_modes_batch(probs, return_info=True, handle=h)
                dt = time.time() - t0
            st = h.last_stats()
            dn = np.abs(out[0][1] - g["n_tight"]).max()
            print(f"## solve B={nb} {opts}: |dn| {dn:.1e} op {info[0]['op_applies']} inner {info[0]['inner_iters']} dev_ms {st['device_ms']:.0f} "
                  f"({st['device_ms'] / max(1, info[0]['inner_iters']):.3f}/it) wall {dt:.2f} launches {st['launches']}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"## solve B={nb} {opts}: FAILED {e}", flush=True)
        h.close()


VARIANTS = [0] + [int(a[1:]) for a in sys.argv[1:] if a.startswith("v")] or [0, 1, 2]
if VARIANTS == [0]:
    VARIANTS = [0, 1, 2]
ROWS = [int(a[1:]) for a in sys.argv[1:] if a.startswith("r")] or [33, 45, 57, 69]
which = [a for a in sys.argv[1:] if a in ("eq", "time", "solve")] or ["eq", "time", "solve"]
if "eq" in which:
    eq()
if "time" in which:
    timing()
if "solve" in which:
    solve()
