"""DEV TOOL: small solves of every code path for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
from tests.golden.cases import CASES, relative_case  # noqa: E402
from tidy3d_b200 import compute_modes_batch  # noqa: E402

names = sys.argv[1:] or ["c1_64", "c3_96", "c4_96", "lossy_48", "slab1d_x1", "angled_48_minus", "offdiag_48", "pec_block_40"]
for name in names:
    fac, kw, _ = CASES[name]
    wl = fac()
    g = np.load(f"/root/repo/tests/golden/{name}.npz")
    out = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)] * 2)
    print(name, "max|dn|", np.abs(out[0][1] - g["n_tight"]).max(), flush=True)
wl = relative_case()
g = np.load("/root/repo/tests/golden/relative_48.npz")
out = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, solver_basis_fields=g["basis"])])
print("relative", np.abs(out[0][1] - g["n_ref"]).max())
