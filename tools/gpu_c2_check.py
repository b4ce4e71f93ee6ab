"""DEV TOOL: BASELINE config 2 at scale (256x256, 64 of the 256 sweep frequencies in one device batch) + parity spot checks."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from oracle import restatement as R
from tidy3d_b200 import compute_modes_batch
from tidy3d_b200 import workloads as W
wl = W.c2(nf=256, n=256)
fr = wl.freqs[::4]
probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in fr]
compute_modes_batch(probs[:8])
t0 = time.time()
out, info = compute_modes_batch(probs, return_info=True)
dt = time.time() - t0
print(f"c2 256x256 x{len(probs)}: wall {dt:.2f}s -> {len(probs)/dt:.1f} solves/s e2e; device {info[0]['solve_ms']:.0f} ms; inner {info[0]['inner_iters']} op {info[0]['op_applies']}")
for i in (0, 63):
    _, n0, _ = R.compute_modes(wl.eps_cross, wl.coords, fr[i], wl.mode_spec, tol=1e-12)
    print(i, "max|dn| vs tight oracle", np.abs(out[i][1] - n0).max())
ns = np.array([o[1] for o in out])
print("monotone n_eff(lambda):", bool((np.diff(ns[:, 0].real) < 0).all()))
