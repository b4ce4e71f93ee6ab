"""DEV TOOL: accuracy / cost of solver options on golden cases."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from tests.golden.cases import CASES
from tidy3d_b200 import _cabi, compute_modes_batch
from tidy3d_b200 import workloads as W
H = _cabi.Handle()
def run(name, nrep=1, **opts):
    fac, kw, _ = CASES[name]; wl = fac()
    H.set_options(**opts)
    g = np.load(f"/root/repo/tests/golden/{name}.npz")
    probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)] * nrep
    t0 = time.time()
    try:
        out, info = compute_modes_batch(probs, return_info=True, handle=H)
        i = info[0]
        print(f"{name:16s} {str(opts):70s} |dn| {np.abs(out[0][1]-g['n_tight']).max():.1e} op {i['op_applies']} inner {i['inner_iters']} launches {i['stencil_applies']} res {i['max_residual']:.1e} ms {i['solve_ms']:.0f}", flush=True)
    except Exception as e:
        print(name, opts, "FAILED", e, flush=True)
for name, nrep in [("headline_512_f0", 32), ("c3_512", 2), ("strip_128_m4", 1)]:
    run(name, nrep=nrep, mg_nu_growth=0)
    run(name, nrep=nrep, mg_nu_growth=1)
    run(name, nrep=nrep, mg_nu_growth=2)
    run(name, nrep=nrep, mg_nu_growth=0, mg_omega=0.9)
    run(name, nrep=nrep, mg_nu_growth=0, mg_omega=0.7)
    H.set_options(mg_omega=0.8, mg_nu_growth=0)
