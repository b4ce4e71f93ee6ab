"""DEV TOOL: inner-solver diagnostics on the large indefinite configs."""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from tests.golden.cases import CASES
from tidy3d_b200 import _cabi, compute_modes_batch
L = _cabi.lib(); H = _cabi.Handle()
rng = np.random.default_rng(3)
def solve(name, **opts):
    fac, kw, _ = CASES[name]; wl = fac()
    H.set_options(**opts)
    pk = _cabi.PackedProblem(wl.eps_cross, wl.coords, wl.freqs[0], wl.mode_spec, kw.get("symmetry", (0, 0)))
    shapes = (C.c_int * 40)()
    nl = L.b200ms_debug_hierarchy(C.byref(pk.struct), C.byref(H.options), 20, shapes)
    r = rng.standard_normal((2, pk.nx, pk.ny)) + 0j
    r[0][:, 0] = 0; r[1][0, :] = 0
    ra = np.ascontiguousarray(r.ravel()); x = np.zeros(r.size, complex); it = C.c_int(); rr = C.c_double()
    t0 = time.time()
    rc = L.b200ms_debug_solve(H._h, C.byref(pk.struct), _cabi._ptr(ra.view(float)), _cabi._ptr(x.view(float)), C.byref(it), C.byref(rr))
    print(f"{name} {opts} levels {[(shapes[2*i], shapes[2*i+1]) for i in range(nl)]} rc {rc} iters {it.value} res {rr.value:.2e} wall {time.time()-t0:.2f}s", flush=True)
for name in sys.argv[1:] or ["c3_512", "c4_512"]:
    solve(name, mg_pml_phase=0.785)
    solve(name, mg_pml_phase=1.0)
    solve(name, mg_pml_phase=0.785, mg_coarse_iters=32)
    solve(name, mg_pml_phase=0.785, mg_ppw=0.0)
    solve(name, mg_pml_phase=0.785, mg_nu=3)
    solve(name, mg_pml_phase=0.785, mg_ppw=8.0, mg_coarse_iters=32)
H.set_options(mg_pml_phase=0.785, mg_ppw=4.0, mg_coarse_iters=16, mg_nu=2)
for name in ["c3_512", "c4_512"]:
    fac, kw, _ = CASES[name]; wl = fac()
    g = np.load(f"/root/repo/tests/golden/{name}.npz")
    t0 = time.time()
    try:
        out, info = compute_modes_batch([dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=wl.freqs[0], mode_spec=wl.mode_spec, **kw)], return_info=True, handle=H)
        print(name, "max|dn|", np.abs(out[0][1] - g["n_tight"]).max(), info[0], f"wall {time.time()-t0:.1f}")
    except Exception as e:
        print(name, "FAILED", e)
