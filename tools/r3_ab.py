"""DEV TOOL: A/B of library options on one 64-problem batch of the headline sweep (run under gpurun).
usage: python tools/r3_ab.py "k=v,k=v" "k=v" ...   (an empty string = defaults)"""
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo")
from tidy3d_b200 import _cabi, compute_modes_batch  # noqa: E402
from tidy3d_b200 import workloads as W  # noqa: E402

REF = dict(eig_tol=1.1920928955078125e-07, inner_tol=1e-8)
g = np.load("/root/repo/tests/golden/headline_512_f0.npz")
wl = W.headline(nf=256, n=512)
nb = 64
probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in wl.freqs[:nb]]
for spec in sys.argv[1:] or [""]:
    opts = {}
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        opts[k] = float(v) if "." in v or "e" in v else int(v)
    h = _cabi.Handle(**{**REF, "max_batch": 64, **opts})
    try:
        for rep in range(2):
            t0 = time.time()
            out, info = compute_modes_batch(probs, return_info=True, handle=h)
            dt = time.time() - t0
        st = h.last_stats()
        dn = np.abs(out[0][1] - g["n_tight"]).max()
        print(f"## B={nb} {opts}: |dn| {dn:.1e} op {info[0]['op_applies']} inner {info[0]['inner_iters']} dev_ms {st['device_ms']:.0f} "
              f"({st['device_ms'] / max(1, info[0]['inner_iters']):.3f}/it) wall {dt:.2f} launches {st['launches']}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"## B={nb} {opts}: FAILED {e}", flush=True)
    h.close()
