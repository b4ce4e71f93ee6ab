"""Parity metrics shared by the tests (phase / degeneracy aware)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated parity contract: fp64, reference ARPACK tolerance is 1.19e-7 relative
N_TOL = 1e-6  # |n_eff - n_eff_ref|  and  |k_eff - k_eff_ref|
OVERLAP_MIN = 0.999  # normalised |<a,b>| over all six field components


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def overlap(a, b):
    a, b = a.ravel(), b.ravel()
    return abs(np.vdot(a, b)) / (np.linalg.norm(a) * np.linalg.norm(b))


def mode_overlaps(f, g):
    """Per-mode overlap of two (2,3,Nx,Ny,1,M) field arrays; near-degenerate modes are not merged here."""
    return np.array([overlap(f[..., m], g[..., m]) for m in range(f.shape[-1])])


def signature(fields):
    f = fields.reshape(6, -1, fields.shape[-1])
    scale = np.sqrt((np.abs(f[:2]) ** 2).sum(axis=(0, 1)))
    return (np.sqrt((np.abs(f) ** 2).sum(axis=1)) / scale).T
