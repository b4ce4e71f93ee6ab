"""Parity metrics shared by the tests (phase / degeneracy aware)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated parity contract: fp64, reference ARPACK tolerance is 1.19e-7 relative
N_TOL = 1e-6  # |n_eff - n_eff_ref|  and  |k_eff - k_eff_ref|   (library default = the reference's own tolerance)
N_TOL_TIGHT = 1e-8  # "tight" preset against the reference re-run with TOL_EIGS = 1e-12
OVERLAP_MIN = 0.999  # normalised |<a,b>|, separately for the E block and the H block
ETA_0 = 376.73031366686166


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def overlap(a, b):
    a, b = a.ravel(), b.ravel()
    return abs(np.vdot(a, b)) / (np.linalg.norm(a) * np.linalg.norm(b))


def mode_overlaps(f, g):
    """Per-mode overlap of two (2,3,Nx,Ny,1,M) field arrays: min of the E-block and the H-block overlaps (H is ~1/ETA_0
    of E, so a joint inner product would not see it), and the relative phase of the H block must equal that of the E
    block (catches sign / -i/eta0 / direction errors in H)."""
    out = []
    for m in range(f.shape[-1]):
        e_ov = np.vdot(g[0, ..., m].ravel(), f[0, ..., m].ravel())
        h_ov = np.vdot(g[1, ..., m].ravel(), f[1, ..., m].ravel())
        ne = np.linalg.norm(f[0, ..., m]) * np.linalg.norm(g[0, ..., m])
        nh = np.linalg.norm(f[1, ..., m]) * np.linalg.norm(g[1, ..., m])
        oe, oh = abs(e_ov) / ne, abs(h_ov) / nh
        # phase consistency: <g_E, f_E> and <g_H, f_H> carry the same phase factor when f = c * g
        phase = abs(e_ov / abs(e_ov) - h_ov / abs(h_ov)) if abs(e_ov) > 0 and abs(h_ov) > 0 else 2.0
        out.append(min(oe, oh) if phase < 2e-2 else -phase)
    return np.array(out)


def signature(fields):
    """(M, 6) component 2-norms per mode after unit-normalising the in-plane E part (phase independent amplitudes)."""
    f = fields.reshape(6, -1, fields.shape[-1])
    scale = np.sqrt((np.abs(f[:2]) ** 2).sum(axis=(0, 1)))
    return (np.sqrt((np.abs(f) ** 2).sum(axis=1)) / scale).T


def well_separated(n, gap=1e-4):
    n = np.asarray(n)
    gaps = np.abs(n[:, None] - n[None, :]) + np.eye(n.size)
    return gaps.min(axis=1) > gap


def cluster_overlaps(f, g, n, gap=1e-4):
    """Subspace overlap for clusters of near-degenerate modes: for each cluster (modes closer than `gap` in n) the
    singular values of the cross-Gram matrix of the E blocks (1 = same subspace)."""
    n = np.asarray(n)
    M = n.size
    order = list(range(M))
    clusters, cur = [], [order[0]]
    for i in order[1:]:
        if min(abs(n[i] - n[j]) for j in cur) < gap:
            cur.append(i)
        else:
            clusters.append(cur)
            cur = [i]
    clusters.append(cur)
    out = []
    for c in clusters:
        A = np.stack([f[0, ..., m].ravel() for m in c], axis=1)
        Bm = np.stack([g[0, ..., m].ravel() for m in c], axis=1)
        qa, _ = np.linalg.qr(A)
        qb, _ = np.linalg.qr(Bm)
        out.append((c, np.linalg.svd(qa.conj().T @ qb, compute_uv=False).min()))
    return out


SKETCH_K, SKETCH_SEED, ETA_0 = 16, 20260923, 376.73031366686166


def sketch(fields):
    """(K, M) complex random sketch of a (2,3,Nx,Ny,1,M) field array (H block scaled by ETA_0 so both blocks weigh equally);
    see tests/golden/make_sketch_golden.py.  Row by row: the K x 6N Gaussian matrix is never materialised."""
    f = np.asarray(fields, dtype=np.complex128).copy()
    f[1] *= ETA_0
    m = f.shape[-1]
    v = f.reshape(-1, m)
    rng = np.random.default_rng(SKETCH_SEED)
    out = np.zeros((SKETCH_K, m), complex)
    for k in range(SKETCH_K):
        p = rng.standard_normal(v.shape[0]) + 1j * rng.standard_normal(v.shape[0])
        out[k] = p @ v
    return out / np.linalg.norm(v, axis=0)


def sketch_similarity(a, b):
    """Per mode |<a, b>| / (|a| |b|) of two (K, M) sketches: 1 - O(err^2) when the underlying unit vectors agree up to a
    global phase, ~ 1/sqrt(K) for unrelated vectors."""
    return np.abs(np.sum(np.conj(a) * b, axis=0)) / (np.linalg.norm(a, axis=0) * np.linalg.norm(b, axis=0))
