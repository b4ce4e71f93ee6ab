"""Golden n_complex along the BASELINE sweeps, from the UNMODIFIED reference (oracle/ref_shim.py; build container only).

    python tests/golden/make_sweep_golden.py headline   # 512x512, 4 modes: 33 of the 256 sweep points (every 8th + last)
    python tests/golden/make_sweep_golden.py c5         # config 5: 4 of the 32 planes x 3 of the 128 frequencies (256x256)

bench.py asserts, inside the timed run, |n - n_ref| <= 1e-6 at the golden points and |n - spline(n_ref)| <= 1e-6 in
between (n_eff(lambda) is smooth; the spline is validated leave-one-out by tests/test_oracle_pinning.py).
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _one(a):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import ref_shim
    from tidy3d_b200 import workloads as W

    kind, plane, fi = a
    if kind == "headline":
        wl = W.headline(nf=256)
    else:
        wl = W.c5_planes()[plane]
    t0 = time.time()
    _, n, _ = ref_shim.compute_modes(wl.eps_cross, wl.coords, wl.freqs[fi], wl.mode_spec)
    print(kind, plane, fi, n, f"{time.time() - t0:.0f}s", flush=True)
    return n


if __name__ == "__main__":
    which = sys.argv[1]
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    if which == "headline":
        idx = sorted(set(list(range(0, 256, 8)) + [255]))
        jobs = [("headline", 0, i) for i in idx]
        with ProcessPoolExecutor(workers) as ex:
            ns = list(ex.map(_one, jobs))
        from tidy3d_b200 import workloads as W

        np.savez_compressed(os.path.join(HERE, "headline_512_sweep.npz"), idx=np.array(idx), freqs=W.sweep_freqs(256)[idx], n_ref=np.array(ns))
    else:
        planes, fis = [0, 10, 21, 31], [0, 64, 127]
        jobs = [("c5", p, f) for p in planes for f in fis]
        with ProcessPoolExecutor(workers) as ex:
            ns = list(ex.map(_one, jobs))
        np.savez_compressed(os.path.join(HERE, "c5_subset.npz"), planes=np.array(planes), fidx=np.array(fis),
                            n_ref=np.array(ns).reshape(len(planes), len(fis), -1))
