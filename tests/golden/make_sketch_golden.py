"""Field goldens for the full-size cases without storing 50-100 MB of fields: a seeded random SKETCH of every mode.

    python tests/golden/make_sketch_golden.py [case ...]      (build container only: runs the UNMODIFIED reference)

For each mode the six field components (E block, and the H block scaled by ETA_0 so that both weigh equally) are projected
on K = 16 fixed complex Gaussian vectors (numpy default_rng(20260923), regenerated at test time); `<case>_sketch.npz` holds
the K x M complex projections of the reference's fields at TOL_EIGS = 1e-12.  Two unit vectors whose K-dimensional sketches
agree up to a global phase to 1e-3 agree themselves to ~1e-3 with overwhelming probability (Johnson-Lindenstrauss), so the
GPU test compares |<s_gpu, s_ref>| / (|s_gpu| |s_ref|) per mode (tests/test_gpu_parity.py::test_full_size_field_sketch).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.helpers import SKETCH_K as K, SKETCH_SEED as SEED, sketch  # noqa: E402


if __name__ == "__main__":
    from oracle import ref_shim
    from tests.golden.cases import CASES, resolve_kwargs

    for name in sys.argv[1:] or ["c2_256_f0", "headline_512_f0", "c4_512", "c3_512"]:
        fac, kw, _ = CASES[name]
        wl = fac()
        kw = resolve_kwargs(wl, kw)
        ref = ref_shim.load()
        old = ref.TOL_EIGS
        ref.TOL_EIGS = 1e-12
        t0 = time.time()
        try:
            fields, n, spec = ref_shim.compute_modes(wl.eps_cross, wl.coords, wl.freqs[0], wl.mode_spec, **kw)
        finally:
            ref.TOL_EIGS = old
        np.savez_compressed(os.path.join(HERE, name + "_sketch.npz"), sketch=sketch(fields), n_tight=n, K=K, seed=SEED)
        print(name, n, f"{time.time() - t0:.0f}s", flush=True)
