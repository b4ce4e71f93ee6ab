"""DEV TOOL (round 2): parameter sweeps on the headline problem (run under gpurun)."""
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo")
from tidy3d_b200 import _cabi, compute_modes_batch  # noqa: E402
from tidy3d_b200 import workloads as W  # noqa: E402

REF = dict(eig_tol=1.1920928955078125e-07, inner_tol=1e-8)
g = np.load("/root/repo/tests/golden/headline_512_f0.npz")


def run(nb, label, n=512, **opts):
    wl = W.headline(nf=256, n=n)
    h = _cabi.Handle(**{**REF, "max_batch": 64, **opts})
    fr = wl.freqs[:nb]
    probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in fr]
    try:
        t0 = time.time()
        out, info = compute_modes_batch(probs, return_info=True, handle=h)
        dt = time.time() - t0
        st = h.last_stats()
        dn = np.abs(out[0][1] - g["n_tight"]).max() if n == 512 else float("nan")
        print(f"## B={nb} {label} {opts}: |dn| {dn:.1e} op {info[0]['op_applies']} inner {info[0]['inner_iters']} rst {info[0]['restarts']} "
              f"dev_ms {st['device_ms']:.0f} ({st['device_ms'] / max(1, info[0]['inner_iters']):.2f}/it) wall {dt:.2f} syncs {st['host_syncs']} launches {st['launches']}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"## B={nb} {label} {opts}: FAILED {e}", flush=True)
    h.close()


which = sys.argv[1:] or ["a"]
if "a" in which:
    for nb in (16, 64):
        run(nb, "legacy", inner_mode=0)
        run(nb, "new-ir", inner_mode=1)
        run(nb, "new-ir floor1e-4", inner_mode=1, ir_floor=1e-4)
        run(nb, "new-ir floor1e-3", inner_mode=1, ir_floor=1e-3)
        run(nb, "new-fp64", inner_mode=1, inner_ir=0)
    for ncv in (24, 28, 32, 40):
        run(16, "legacy ncv", inner_mode=0, ncv=ncv)
    for kw in (dict(mg_nu=1), dict(mg_nu=3), dict(mg_nu_growth=1), dict(mg_omega=0.7), dict(mg_cycles=2), dict(inner_relax=3.0), dict(inner_relax=10.0),
               dict(inner_relax_cap=1e-3), dict(gmres_cgs2=0), dict(gmres_cgs2=1), dict(use_graph=0)):
        run(16, "legacy", inner_mode=0, **kw)
if "b" in which:
    for nb in (64,):
        run(nb, "new default")
        run(nb, "new txr32", stencil_variant=2)
        run(nb, "new no-dgks", outer_dgks=0)
        run(nb, "new floor2e-5", ir_floor=2e-5)
        run(nb, "legacy", inner_mode=0)
    run(16, "new default")
    run(16, "new txr32", stencil_variant=2)
if "pec" in which:
    from tests.golden.cases import CASES
    fac, kw, _ = CASES["pec_block_40"]
    wl = fac()
    for opts in (dict(inner_mode=1), dict(inner_mode=0), dict(inner_mode=1, inner_ir=0)):
        h = _cabi.Handle(**{**REF, **opts, "verbose": 2})
        try:
            out = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)], handle=h)
            print("pec", opts, out[0][1], flush=True)
        except Exception as e:  # noqa: BLE001
            print("pec", opts, "FAILED", e, flush=True)
        h.close()
if "c" in which:
    for nb in (64, 16):
        run(nb, "async0", stencil_async=0)
        run(nb, "async1", stencil_async=1)
    run(64, "async1 txr32", stencil_async=1, stencil_variant=2)
    run(64, "kappa1 (ARPACK test)", kappa_cap=1.0)
    for name in ("c3_512", "c4_512"):
        from tests.golden.cases import CASES
        fac, kw, _ = CASES[name]
        wl = fac()
        gg = np.load(f"/root/repo/tests/golden/{name}.npz")
        for opts in (dict(stencil_async=0), dict(stencil_async=1)):
            h = _cabi.Handle(**{**REF, **opts})
            t0 = time.time()
            out, info = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)], return_info=True, handle=h)
            print(f"## {name} {opts}: |dn| {np.abs(out[0][1] - gg['n_tight']).max():.1e} inner {info[0]['inner_iters']} dev_ms {h.last_stats()['device_ms']:.0f} wall {time.time()-t0:.2f}", flush=True)
            h.close()
if "pml" in which:
    from tests.golden.cases import CASES
    fac, kw, _ = CASES["pml_none_128"]
    wl = fac()
    gg = np.load("/root/repo/tests/golden/pml_none_128.npz")
    for label, opts in (("ref", dict(REF)), ("tight", dict(eig_tol=1e-9, inner_tol=1e-10)), ("ref-norelax", dict(REF, inner_relax=0.0)), ("ref-kappa1", dict(REF, kappa_cap=1.0))):
        h = _cabi.Handle(**{**opts, "verbose": 2 if label == "ref" else 0})
        t0 = time.time()
        out, info = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)], return_info=True, handle=h)
        print(f"## pml_none_128 {label}: dn {np.abs(out[0][1] - gg['n_tight'])} op {info[0]['op_applies']} inner {info[0]['inner_iters']} rst {info[0]['restarts']} wall {time.time()-t0:.2f}", flush=True)
        h.close()
if "fused" in which:
    for nb in (64, 16, 1):
        run(nb, "fused1", mg_fused_tail=1)
        run(nb, "fused0", mg_fused_tail=0)
    from tests.golden.cases import CASES
    for name in ("c1_64", "strip_128_m4", "c3_96", "c4_96", "lossy_48", "nonuniform_56", "slab1d_x1", "c3_512"):
        fac, kw, _ = CASES[name]
        wl = fac()
        gg = np.load(f"/root/repo/tests/golden/{name}.npz")
        for opts in (dict(mg_fused_tail=1), dict(mg_fused_tail=0)):
            h = _cabi.Handle(**{**REF, **opts})
            try:
                compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)], handle=h)
                t0 = time.time()
                out, info = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)], return_info=True, handle=h)
                print(f"## {name} {opts}: |dn| {np.abs(out[0][1] - gg['n_tight']).max():.1e} inner {info[0]['inner_iters']} dev_ms {h.last_stats()['device_ms']:.1f} wall {time.time()-t0:.3f}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"## {name} {opts}: FAILED {e}", flush=True)
            h.close()
if "d" in which:
    run(64, "default")
    run(64, "transfer_tiled0", transfer_tiled=0)
    run(16, "default")
    run(16, "transfer_tiled0", transfer_tiled=0)
    # warm start across windows: 128 problems = two windows of 64
    for ws in (0, 1):
        wl = W.headline(nf=256)
        h = _cabi.Handle(**{**REF, "max_batch": 64, "warm_start": ws})
        probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in wl.freqs[:128]]
        t0 = time.time()
        out, info = compute_modes_batch(probs, return_info=True, handle=h, want_fields=False)
        st = h.last_stats()
        gs = np.load("/root/repo/tests/golden/headline_512_sweep.npz")
        dn = max(np.abs(out[int(i)][1] - gs["n_ref"][j]).max() for j, i in enumerate(gs["idx"]) if i < 128)
        print(f"## warm_start={ws}: 128 problems dev_ms {st['device_ms']:.0f} op/solve {st['op_applies']/128:.1f} inner/solve {st['inner_iters']/128:.0f} max|dn| vs golden {dn:.1e} wall {time.time()-t0:.1f}", flush=True)
        h.close()
if "e" in which:
    run(64, "default")
    run(16, "default")
    for kk in (8, 10, 14, 16):
        run(16, "ks_keep", ks_keep=kk)
    run(16, "inner_tol 3e-8", inner_tol=3e-8)
    run(16, "ncv 16", ncv=16)
if "f" in which:
    run(64, "default")
    run(64, "mg_fuse_first0", mg_fuse_first=0)
    run(16, "default")
    run(16, "mg_fuse_first0", mg_fuse_first=0)
    from tests.golden.cases import CASES
    for name in ("c3_128", "c4_128", "nonuniform_56", "c3_512"):
        fac, kw, _ = CASES[name]
        wl = fac()
        gg = np.load(f"/root/repo/tests/golden/{name}.npz")
        for opts in (dict(mg_fuse_first=1), dict(mg_fuse_first=0)):
            h = _cabi.Handle(**{**REF, **opts})
            out, info = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)], return_info=True, handle=h)
            print(f"## {name} {opts}: |dn| {np.abs(out[0][1] - gg['n_tight']).max():.1e} inner {info[0]['inner_iters']} dev_ms {h.last_stats()['device_ms']:.1f}", flush=True)
            h.close()
if "g" in which:
    for nb in (32, 16, 64):
        run(nb, "wave rule")
        run(nb, "no wave rule", stencil_variant=3)
