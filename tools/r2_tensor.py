"""DEV TOOL: throughput of the tensorial (angled) path at a given size.  `python tools/r2_tensor.py gpu|ref [n] [nf]`.
gpu: one batch of nf frequencies through the library; ref: the unmodified reference on the first frequency (build container)."""
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo")
from tidy3d_b200 import workloads as W  # noqa: E402

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 8
wl = W.angled(n, theta=0.2, phi=0.0, num_modes=4)
freqs = wl.freqs[0] * np.linspace(0.97, 1.03, nf)
if which == "ref":
    from oracle import ref_shim

    t0 = time.time()
    f, nc, spec = ref_shim.compute_modes(wl.eps_cross, wl.coords, freqs[0], wl.mode_spec)
    print(f"## reference angled {n}: {time.time() - t0:.1f} s  n = {nc}  {spec}")
else:
    from tidy3d_b200 import compute_modes_batch
    from tidy3d_b200.solver import get_handle

    probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in freqs]
    h = get_handle()
    for rep in range(2):
        t0 = time.time()
        out, info = compute_modes_batch(probs, return_info=True, handle=h)
        wall = time.time() - t0
    st = h.last_stats()
    print(f"## gpu angled {n} x {nf}: wall {wall:.2f} s = {nf / wall:.2f} solves/s; n[0] = {out[0][1]}; op {info[0]['op_applies']}"
          f" inner {info[0]['inner_iters']} stats {st}")
