"""Generates tests/golden/post_<case>.npz: for every case of tests/post_cases.py the seeded inputs and what
the UNMODIFIED reference's post-processing code makes of them -- ``ModeSolver.data_raw`` with its steps (gauge, colocation,
flux normalisation with finite-grid correction, polarisation filter, mode tracking), ``flux``, ``pol_fraction``, ``dot`` and
``outer_dot`` -- executed by oracle/ref_post.py (the reference's own method bodies over a stand-in for xarray; see there).

Run in the build container (needs /root/reference):   python tests/golden/make_post_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tests import post_cases as PC  # noqa: E402


def main():
    import warnings

    for name in PC.CASES:
        c = PC.inputs(name)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)  # scipy's 0/0 in the one-point interpolation (see skipped_keys)
            ref = PC.reference_results(c)
        out = {f"in_{k}": v for k, v in PC.to_arrays(c).items()}
        out.update({f"ref_{k}": np.asarray(ref[k]) for k in PC.KEYS})
        path = os.path.join(ROOT, "tests", "golden", f"post_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB")


if __name__ == "__main__":
    main()
