"""DEV TOOL: the pair-marching / TMA stencil kernels, the packed-list restriction and the 128-bit prolongation on small grids, for
compute-sanitizer (memcheck): one strip with TMA staging (ny % 4 == 0), the register fallback (ny % 4 == 2), several strips."""
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
from tidy3d_b200 import compute_modes_batch  # noqa: E402
from tidy3d_b200 import workloads as W  # noqa: E402

for nx, ny in ((64, 64), (70, 150), (48, 600)):
    eps, coords = W.strip_eps(nx, ny)
    wl = W.Workload(name=f"strip_{nx}x{ny}", eps_cross=W._iso(eps), coords=coords, freqs=np.array([W.C_0 / 1.55]),
                    mode_spec=W.ModeSpecLike(num_modes=2, precision="double"))
    out = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec)])
    print(wl.name, out[0][1], flush=True)
