"""DEV TOOL: one headline batch with a given inner mode (for ncu launch lists)."""
import sys

sys.path.insert(0, "/root/repo")
from tidy3d_b200 import _cabi, compute_modes_batch  # noqa: E402
from tidy3d_b200 import workloads as W  # noqa: E402

mode = int(sys.argv[1])
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wl = W.headline(nf=256)
h = _cabi.Handle(eig_tol=1.1920928955078125e-07, inner_tol=1e-8, max_batch=64, inner_mode=mode)
probs = [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in wl.freqs[:nb]]
out, info = compute_modes_batch(probs, return_info=True, handle=h, want_fields=False)
print(info[0], h.last_stats())
