"""DEV TOOL: a short pass through every kernel family on small grids, for compute-sanitizer (memcheck / racecheck)."""
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
from tests.golden.cases import CASES, resolve_kwargs  # noqa: E402
from tidy3d_b200 import compute_modes_batch  # noqa: E402
from tidy3d_b200 import workloads as W  # noqa: E402
from tidy3d_b200.sections import Medium, Rect, Section  # noqa: E402

for name in ("c1_64", "lossy_48", "nonuniform_56", "pec_split_40", "offdiag_48", "slab1d_x1"):
    fac, kw, _ = CASES[name]
    wl = fac()
    kw = resolve_kwargs(wl, kw)
    out = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)])
    print(name, out[0][1], flush=True)
wl = W.c2(nf=3, n=80)
out, info = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in wl.freqs],
                                post=("gauge", "normalize", "flux", "overlaps"), return_info=True)
print("post", info[1]["flux"], flush=True)
c = np.linspace(-1.5, 1.5, 49)
sec = Section(Medium(2.0736), [(Rect((0, 0), (0.45, 0.22)), Medium(12.1))])
print("section", compute_modes_batch([dict(section=sec, coords=[c, c], freq=W.C_0 / 1.55, mode_spec=W.ModeSpecLike(num_modes=2, precision="single"))])[0][1])
