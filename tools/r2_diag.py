"""DEV TOOL (round 2): A/B of the inner solvers and tolerance presets on the GPU (run under gpurun)."""
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo")
from tests.golden.cases import CASES, resolve_kwargs  # noqa: E402
from tidy3d_b200 import _cabi, compute_modes_batch  # noqa: E402
from tidy3d_b200 import workloads as W  # noqa: E402

H = {}


def handle(**opts):
    key = tuple(sorted(opts.items()))
    if key not in H:
        H[key] = _cabi.Handle(**opts)
    return H[key]


def run(name, nbatch=1, label="", **opts):
    fac, kw, _ = CASES[name]
    wl = fac()
    kw = resolve_kwargs(wl, kw)
    g = np.load(f"/root/repo/tests/golden/{name}.npz")
    h = handle(**opts)
    if nbatch == 1:
        probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)]
    else:
        fr = W.sweep_freqs(256)[:nbatch]
        fr[0] = wl.freqs[0]
        probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec, **kw) for f in fr]
    t0 = time.time()
    try:
        out, info = compute_modes_batch(probs, return_info=True, handle=h)
    except Exception as e:  # noqa: BLE001
        print(f"## {name} {label} {opts}: FAILED {type(e).__name__}: {e}", flush=True)
        return
    dt = time.time() - t0
    st = h.last_stats()
    n = out[0][1]
    print(f"## {name} B={nbatch} {label} {opts}: |dn_tight| {np.abs(n - g['n_tight']).max():.2e} |dn_ref| {np.abs(n - g['n_ref']).max():.2e} "
          f"op {info[0]['op_applies']} inner {info[0]['inner_iters']} restarts {info[0]['restarts']} res {info[0]['max_residual']:.1e} "
          f"dev_ms {st['device_ms']:.1f} wall {dt:.2f}s syncs {st['host_syncs']} launches {st['launches']}", flush=True)


REF = dict(eig_tol=1.1920928955078125e-07, inner_tol=1e-8)
TIGHT = dict(eig_tol=1e-9, inner_tol=1e-10)
which = sys.argv[1:] or ["small", "big"]
if "small" in which:
    for name in ["c1_64", "strip_128_m4", "c3_96", "c4_96", "lossy_48", "nonuniform_56", "slab1d_x1", "pec_block_40", "angled_64", "c3_128"]:
        run(name, label="legacy/tight", inner_mode=0, **TIGHT)
        run(name, label="new-fp64/tight", inner_mode=1, inner_ir=0, **TIGHT)
        run(name, label="new-ir/tight", inner_mode=1, inner_ir=1, **TIGHT)
        run(name, label="new-ir/ref", inner_mode=1, inner_ir=1, **REF)
if "big" in which:
    for nb in (16, 64):
        run("headline_512_f0", nb, label="legacy/tight", inner_mode=0, max_batch=64, **TIGHT)
        run("headline_512_f0", nb, label="legacy/ref", inner_mode=0, max_batch=64, **REF)
        run("headline_512_f0", nb, label="new-fp64/ref", inner_mode=1, inner_ir=0, max_batch=64, **REF)
        run("headline_512_f0", nb, label="new-ir/tight", inner_mode=1, inner_ir=1, max_batch=64, **TIGHT)
        run("headline_512_f0", nb, label="new-ir/ref", inner_mode=1, inner_ir=1, max_batch=64, **REF)
        run("headline_512_f0", nb, label="new-ir/ref-1e-7", inner_mode=1, inner_ir=1, max_batch=64, eig_tol=1.19e-7, inner_tol=1e-7)
    for name in ("c3_512", "c4_512", "c2_256_f0"):
        run(name, label="legacy/tight", inner_mode=0, **TIGHT)
        run(name, label="new-ir/tight", inner_mode=1, **TIGHT)
        run(name, label="new-ir/ref", inner_mode=1, **REF)
