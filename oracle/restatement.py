"""TEST INFRASTRUCTURE ONLY -- CPU (numpy/scipy) restatement of the reference mode-solver numerics.

This is the parity oracle for ``tidy3d_b200``.  It restates, in our own words, what
``/root/reference/tidy3d/plugins/mode/{solver,derivatives,transforms}.py`` compute; every function
cites the reference lines it follows.  It is pinned two ways (tests/test_oracle_pinning.py):
  * against the *unmodified* reference loaded by ``oracle/ref_shim.py`` (build container only), and
  * against the golden fixtures in ``tests/golden`` (generated from the unmodified reference by
    ``tests/golden/make_golden.py``; these travel to the GPU box),
plus the reference's own exact known-answer test ``test_pml_params``
(tests/test_plugins/test_mode_solver.py:783-806).

Third-party arithmetic: the eigen-decomposition is SciPy's ARPACK+SuperLU
(``scipy.sparse.linalg.eigs``; reference call site solver.py:744-746; reference locks scipy 1.13.1,
this image has 1.18.1), which is not vendored in the reference tree; we call the same routine.

Only ``tests/``, ``__graft_entry__.smoke`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file.  The product (``tidy3d_b200``) never does.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

# tidy3d/constants.py:16-64
C_0 = 2.99792458e14
MU_0 = 1.25663706212e-12
EPSILON_0 = 1.0 / (MU_0 * C_0**2)
ETA_0 = np.sqrt(MU_0 / EPSILON_0)
FP_EPS = float(np.finfo(np.float32).eps)
PEC_VAL = -1e8

TOL_TENSORIAL = 1e-6  # solver.py:22
TARGET_SHIFT = 10 * FP_EPS  # solver.py:24


# --------------------------------------------------------------------------------------------
# PML stretch (derivatives.py:79-232)
# --------------------------------------------------------------------------------------------
def s_value(dl, step, omega, avg_speed):
    """derivatives.py:200-232 with the defaults sigma_max=2, kappa 1..3, cubic profile."""
    p = step**3
    return (1.0 + 2.0 * p) + 1j * (2.0 * avg_speed / (ETA_0 * dl) * p) / (omega * EPSILON_0)


def sfactor(kind, omega, dls, n, n_pml, pml_at_min, speeds):
    """derivatives.py:158-196.  kind 'f': half-integer steps (H sites); 'b': integer (E sites)."""
    s = np.ones(n, dtype=complex)
    if n_pml == 0:
        return s
    for i in range(n):
        if kind == "f":
            if i <= n_pml - 1 and pml_at_min:
                s[i] = s_value(dls[0], (n_pml - i - 0.5) / n_pml, omega, speeds[0])
            elif i >= n - n_pml:
                s[i] = s_value(dls[-1], (i - (n - n_pml) + 0.5) / n_pml, omega, speeds[1])
        else:
            if i < n_pml and pml_at_min:
                s[i] = s_value(dls[0], (n_pml - i) / n_pml, omega, speeds[0])
            elif i > n - n_pml:
                s[i] = s_value(dls[-1], (i - (n - n_pml)) / n_pml, omega, speeds[1])
    return s


def pml_speeds(nx, ny, npml, eps_t, mu_t):
    """derivatives.py:129-155: 1/sqrt(<eps_diag><mu_diag>) over the four PML strips."""

    def strip_means(t):
        d = np.stack([t[0, 0], t[1, 1], t[2, 2]]).reshape(3, nx, ny)
        regions = (d[:, : npml[0], :], d[:, nx - npml[0] + 1 :, :], d[:, :, : npml[1]], d[:, :, ny - npml[1] + 1 :])
        return np.array([r.mean() if r.size else 1.0 for r in regions])

    return 1.0 / np.sqrt(strip_means(eps_t) * strip_means(mu_t))


# --------------------------------------------------------------------------------------------
# 1-D difference operators (derivatives.py:9-76) as (lo, hi) coefficient pairs
# --------------------------------------------------------------------------------------------
def diff_coeffs(n, dl_f, dl_b, pmc, s_f, s_b, k0):
    """Rows of the four bidiagonal operators along one axis, already scaled by 1/(s*k0).

    forward:  (D_f v)[i] = f0[i]*v[i] + f1[i]*v[i+1]     (f1[n-1] = 0: PEC at max)
    backward: (D_b v)[i] = b0[i]*v[i] + bm[i]*v[i-1]     (bm[0] = 0)
    n == 1 gives all-zero operators (derivatives.py:12-13).
    """
    f0 = np.zeros(n, complex)
    f1 = np.zeros(n, complex)
    b0 = np.zeros(n, complex)
    bm = np.zeros(n, complex)
    if n > 1:
        f0[:] = -1.0 / dl_f
        f1[:-1] = 1.0 / dl_f[:-1]
        if not pmc:
            f0[0] = 0.0
        b0[:] = 1.0 / dl_b
        bm[1:] = -1.0 / dl_b[1:]
        b0[0] = 2.0 / dl_b[0] if pmc else 0.0
        f0 /= s_f * k0
        f1 /= s_f * k0
        b0 /= s_b * k0
        bm /= s_b * k0
    return f0, f1, b0, bm


def _bidiag(n, d0, d1, off):
    return sp.diags([d0, d1[:-1] if off > 0 else d1[1:]], [0, off], shape=(n, n), format="csr")


# --------------------------------------------------------------------------------------------
# set-up shared by the assembled oracle and the matrix-free model
# --------------------------------------------------------------------------------------------
def setup(eps_cross, coords, freq, mode_spec, symmetry=(0, 0), mu_cross=None):
    """Everything ``EigSolver.compute_modes`` does before ``solver_em`` (solver.py:86-217)."""
    eps9 = [np.array(eps_cross[i], dtype=complex) for i in range(9)]
    if len(eps9) != 9:
        raise ValueError("Wrong input to mode solver pemittivity/permeability!")
    nx, ny = eps9[0].shape
    n = nx * ny
    if len(coords[0]) != nx + 1 or len(coords[1]) != ny + 1:
        raise ValueError("Mismatch between 'coords' and 'esp_cross' shapes.")
    omega = 2 * np.pi * freq
    k0 = omega / C_0
    eps_t = np.array(eps9).reshape(3, 3, n)
    mu_t = np.zeros((3, 3, n), complex)
    if mu_cross is None:
        for a in range(3):
            mu_t[a, a] = 1.0
    else:
        mu_t = np.array([np.array(m, complex) for m in mu_cross]).reshape(3, 3, n)

    new_coords = [np.array(c, float) for c in coords]
    jac_e = np.zeros((3, 3, n))
    jac_h = np.zeros((3, 3, n))
    for a in range(3):
        jac_e[a, a] = jac_h[a, a] = 1.0
    theta, phi = mode_spec.angle_theta, mode_spec.angle_phi
    if abs(theta) > 0:  # transforms.py:74-111
        jac_e[0, 2] = jac_h[0, 2] = -np.tan(theta) * np.cos(phi)
        jac_e[1, 2] = jac_h[1, 2] = -np.tan(theta) * np.sin(phi)
    if mode_spec.bend_radius is not None:  # transforms.py:14-71
        radius, bend_axis = mode_spec.bend_radius, mode_spec.bend_axis
        norm_axis = 0 if bend_axis == 1 else 1
        c = new_coords[norm_axis]
        c = c + (radius - c[(c.size - 1) // 2])
        new_coords[norm_axis] = c
        dw_e = radius / c[:-1]
        dw_h = 2 * radius / (c[:-1] + c[1:])
        shape = (nx, 1) if norm_axis == 0 else (1, ny)
        je = np.broadcast_to(dw_e.reshape(shape), (nx, ny)).ravel()
        jh = np.broadcast_to(dw_h.reshape(shape), (nx, ny)).ravel()
        # J_bend = diag(1,1,dwdz) applied on the left of the angle Jacobian (solver.py:147-148)
        jac_e = jac_e.copy()
        jac_h = jac_h.copy()
        jac_e[2] = jac_e[2] * je
        jac_h[2] = jac_h[2] * jh

    kxy = np.cos(theta) ** 2
    kp_to_k = np.array([kxy * np.sin(phi), kxy * np.cos(phi), np.cos(theta) * np.sin(theta)])
    knorm = float(np.linalg.norm(kp_to_k))

    def congruence(jac, t):  # solver.py:165-172
        det = np.linalg.det(np.moveaxis(jac, [0, 1], [-2, -1]))
        out = np.einsum("ijn,jpn->ipn", jac, t)
        out = np.einsum("ijn,pjn->ipn", out, jac)
        return out / det

    eps_t = congruence(jac_e, eps_t)
    mu_t = congruence(jac_h, mu_t)

    pmc = [s == 1 for s in symmetry]  # solver.py:184
    dl_f = [c[1:] - c[:-1] for c in new_coords]
    dl_b = [np.hstack((d[0], (d[:-1] + d[1:]) / 2)) for d in dl_f]  # solver.py:187-190
    pml_min = [s == 0 for s in symmetry]  # solver.py:197
    speeds = pml_speeds(nx, ny, mode_spec.num_pml, eps_t, mu_t)
    dims = (nx, ny)
    coef, slen = [], []
    for ax in range(2):
        s_f = sfactor("f", omega, dl_f[ax], dims[ax], mode_spec.num_pml[ax], pml_min[ax], speeds[2 * ax : 2 * ax + 2])
        s_b = sfactor("b", omega, dl_b[ax], dims[ax], mode_spec.num_pml[ax], pml_min[ax], speeds[2 * ax : 2 * ax + 2])
        coef.append(diff_coeffs(dims[ax], dl_f[ax], dl_b[ax], pmc[ax], s_f, s_b, k0))
        slen.append((s_f * dl_f[ax], s_b * dl_b[ax]))  # complex-stretched primal / dual lengths

    if mode_spec.target_neff is None:  # solver.py:204-217
        phys = np.array(eps_cross)
        phys = phys[np.abs(phys) < abs(PEC_VAL)]
        target = np.sqrt(np.max(np.abs(phys)))
    else:
        target = mode_spec.target_neff
    target = target / knorm
    if abs(TARGET_SHIFT) > abs(target * TARGET_SHIFT):
        target += TARGET_SHIFT
    else:
        target *= 1 + TARGET_SHIFT

    # PEC -> lossy metal model (solver.py:327-333)
    eps_t = eps_t.astype(complex)
    eps_t[eps_t <= 0.9 * PEC_VAL] = 1 + 1j * abs(PEC_VAL)
    off = ~np.eye(3, dtype=bool)
    tensorial = bool(np.any(np.abs(eps_t[off]) > TOL_TENSORIAL) or np.any(np.abs(mu_t[off]) > TOL_TENSORIAL))
    return dict(
        nx=nx, ny=ny, n=n, k0=k0, eps=eps_t, mu=mu_t, coef=coef, jac_e=jac_e, jac_h=jac_h,
        target=target, knorm=knorm, tensorial=tensorial, slen=slen, new_coords=new_coords, pmc=pmc,
    )  # fmt: skip


def d_matrices(st):
    """Kronecker lift of the 1-D operators (derivatives.py:18,32,46,60), C-order idx = ix*ny+iy."""
    nx, ny = st["nx"], st["ny"]
    (xf0, xf1, xb0, xbm), (yf0, yf1, yb0, ybm) = st["coef"]
    n = nx * ny
    if nx > 1:
        dxf = sp.kron(_bidiag(nx, xf0, xf1, 1), sp.eye(ny), format="csr")
        dxb = sp.kron(_bidiag(nx, xb0, xbm, -1), sp.eye(ny), format="csr")
    else:
        dxf = dxb = sp.csr_matrix((n, n), dtype=complex)
    if ny > 1:
        dyf = sp.kron(sp.eye(nx), _bidiag(ny, yf0, yf1, 1), format="csr")
        dyb = sp.kron(sp.eye(nx), _bidiag(ny, yb0, ybm, -1), format="csr")
    else:
        dyf = dyb = sp.csr_matrix((n, n), dtype=complex)
    return dxf, dxb, dyf, dyb


def _is_complex(a):
    """solver.py:779-795."""
    if sp.issparse(a):
        return spl.norm(a.imag) / (spl.norm(a) + FP_EPS) > FP_EPS
    return np.linalg.norm(a.imag) / (np.linalg.norm(a) + FP_EPS) > FP_EPS


def solver_dtype(st, precision):
    """solver.py:389-411: real unless eps, mu or any derivative has an imaginary part."""
    single = precision == "single"
    if st["tensorial"]:
        return np.complex64 if single else np.complex128
    # the Frobenius-norm ratio of D (x) I equals that of the 1-D operator, so test those
    ders = []
    for f0, f1, b0, bm in st["coef"]:
        ders += [np.concatenate((f0, f1)), np.concatenate((b0, bm))]
    cplx = _is_complex(st["eps"]) or _is_complex(st["mu"]) or any(_is_complex(d) for d in ders)
    if cplx:
        return np.complex64 if single else np.complex128
    return np.float32 if single else np.float64


def initial_vector(nx, ny, ncomp):
    """solver.py:822-857: rng(0) complex start vector, zero on the ix=0 / iy=0 rows, comp-major."""
    rng = np.random.default_rng(0)
    v = rng.random((nx, ny, ncomp)) + 1j * rng.random((nx, ny, ncomp))
    if nx > 1:
        v[0, :, :] = 0
    if ny > 1:
        v[:, 0, :] = 0
    return np.vstack(v).flatten("F")


def _dg(v):
    return sp.diags(v, 0, format="csr")


def assemble_diagonal(st):
    """P, Q and A = P.Q of solver.py:471-490 (2N x 2N)."""
    e, m = st["eps"], st["mu"]
    dxf, dxb, dyf, dyb = d_matrices(st)
    iez, imz = _dg(1 / e[2, 2]), _dg(1 / m[2, 2])
    pmat = sp.bmat([[-dxf @ iez @ dyb, dxf @ iez @ dxb + _dg(m[1, 1])], [-dyf @ iez @ dyb - _dg(m[0, 0]), dyf @ iez @ dxb]], format="csr")
    qmat = sp.bmat([[-dxb @ imz @ dyf, dxb @ imz @ dxf + _dg(e[1, 1])], [-dyb @ imz @ dyf - _dg(e[0, 0]), dyb @ imz @ dxf]], format="csr")
    return pmat, qmat, (pmat @ qmat).tocsr()


def assemble_tensorial(st):
    """The 4N x 4N first-order operator of solver.py:604-666 (before the -1j factor)."""
    e, m = st["eps"], st["mu"]
    dxf, dxb, dyf, dyb = d_matrices(st)
    iez, imz = _dg(1 / e[2, 2]), _dg(1 / m[2, 2])

    def r(t, a, b):  # t_ab / t_zz
        return _dg(t[a, b] / t[2, 2])

    def sc(t, a, b):  # Schur complement entry t_ab - t_az t_zb / t_zz
        return _dg(t[a, b] - t[a, 2] * t[2, b] / t[2, 2])

    blocks = [
        [-dxf @ r(e, 2, 0) - r(m, 1, 2) @ dyf, -dxf @ r(e, 2, 1) + r(m, 1, 2) @ dxf, -dxf @ iez @ dyb + sc(m, 1, 0), dxf @ iez @ dxb + sc(m, 1, 1)],
        [-dyf @ r(e, 2, 0) + r(m, 0, 2) @ dyf, -dyf @ r(e, 2, 1) - r(m, 0, 2) @ dxf, -dyf @ iez @ dyb - sc(m, 0, 0), dyf @ iez @ dxb - sc(m, 0, 1)],
        [-dxb @ imz @ dyf + sc(e, 1, 0), dxb @ imz @ dxf + sc(e, 1, 1), -dxb @ r(m, 2, 0) - r(e, 1, 2) @ dyb, -dxb @ r(m, 2, 1) + r(e, 1, 2) @ dxb],
        [-dyb @ imz @ dyf - sc(e, 0, 0), dyb @ imz @ dxf - sc(e, 0, 1), -dyb @ r(m, 2, 0) + r(e, 0, 2) @ dyb, -dyb @ r(m, 2, 1) - r(e, 0, 2) @ dxb],
    ]
    return sp.bmat(blocks, format="csr")


def _cast(a, dtype):
    """solver.py:797-819."""
    if np.issubdtype(dtype, np.complexfloating):
        return a.astype(dtype)
    return a.real.astype(dtype)


def _trim(mat):
    """solver.py:414-420 (single precision only)."""
    mx = np.amax(np.abs(mat))
    mat.data *= np.logical_or(np.abs(mat.data) / mx > FP_EPS, np.abs(mat.data) > FP_EPS)
    mat.eliminate_zeros()


def compute_modes(eps_cross, coords, freq, mode_spec, mu_cross=None, split_curl_scaling=None,
                  symmetry=(0, 0), direction="+", solver_basis_fields=None, tol=FP_EPS, info=None):  # fmt: skip
    """Restatement of ``compute_modes`` (solver.py:33-269, 941).  ``tol`` defaults to the
    reference's ARPACK tolerance (solver.py:20); tests pass 1e-12 for a tight oracle.
    the incidence-matrix variant used with ``mu_cross`` / ``split_curl_scaling`` (solver.py:441-449) is restated too
    (the product does not build it yet)."""
    if split_curl_scaling is not None:  # solver.py:122-124
        eps_cross = [np.array(eps_cross[i], dtype=complex, copy=True) for i in range(9)]
        for comp, idx in enumerate((0, 4, 8)):
            sc = np.asarray(split_curl_scaling[comp])
            outside = ~np.isclose(sc, 0)
            eps_cross[idx][outside] /= sc[outside]
    incidence = split_curl_scaling is not None or mu_cross is not None  # solver.py:93
    st = setup(eps_cross, coords, freq, mode_spec, symmetry, mu_cross)
    n, m_modes = st["n"], mode_spec.num_modes
    basis_vecs = None
    if solver_basis_fields is not None:  # solver.py:219-236, 523-528
        if st["tensorial"]:
            raise RuntimeError("Tensorial eps not yet supported in relative mode solver (with basis fields provided).")
        try:
            basis_e = np.asarray(solver_basis_fields)[:3, ...].reshape((3, n, m_modes))
        except ValueError:
            raise ValueError("Shape mismatch between 'basis_fields' and requested mode data.")
        jinv = np.moveaxis(np.linalg.inv(np.moveaxis(st["jac_e"], [0, 1], [-2, -1])), [-2, -1], [0, 1])
        basis_e = np.einsum("ijn,inm->jnm", jinv, basis_e)
        basis_vecs = np.concatenate((basis_e[0], basis_e[1]), axis=0)
    e, m = st["eps"], st["mu"]
    dtype = solver_dtype(st, mode_spec.precision)
    dxf, dxb, dyf, dyb = d_matrices(st)
    counter = {"applies": 0}

    def eigs(mat, sigma, v0):
        # count OP^-1 applies like SURVEY Appendix B by wrapping the factorised solve
        lu = spl.splu((mat - sigma * sp.identity(mat.shape[0], dtype=mat.dtype, format="csc")).tocsc())

        def opinv(x):
            counter["applies"] += 1
            return lu.solve(x)

        op = spl.LinearOperator(mat.shape, matvec=opinv, dtype=mat.dtype)
        return spl.eigs(mat, k=m_modes, sigma=sigma, tol=tol, v0=v0, OPinv=op)

    if not st["tensorial"]:
        spec = "diagonal"
        keep = None
        if incidence:
            # solver.py:441-449, 474-477: unknowns on PEC-valued cells are removed (incidence matrices) and 1/eps_zz is
            # zeroed there; identity when no cell is PEC
            thr = 0.9 * abs(PEC_VAL)
            st = dict(st)
            e = st["eps"] = st["eps"].copy()
            zz_pec = np.abs(e[2, 2]) >= thr
            keep = np.concatenate((np.abs(e[0, 0]) < thr, np.abs(e[1, 1]) < thr))
            ezz_inv = np.where(zz_pec, 0.0, 1.0 / e[2, 2])
            dxf_, dxb_, dyf_, dyb_ = d_matrices(st)
            iez_, imz_ = _dg(ezz_inv), _dg(1 / st["mu"][2, 2])
            m_ = st["mu"]
            pmat = sp.bmat([[-dxf_ @ iez_ @ dyb_, dxf_ @ iez_ @ dxb_ + _dg(m_[1, 1])], [-dyf_ @ iez_ @ dyb_ - _dg(m_[0, 0]), dyf_ @ iez_ @ dxb_]], format="csr")
            qmat = sp.bmat([[-dxb_ @ imz_ @ dyf_, dxb_ @ imz_ @ dxf_ + _dg(e[1, 1])], [-dyb_ @ imz_ @ dyf_ - _dg(e[0, 0]), dyb_ @ imz_ @ dxf_]], format="csr")
            mat = (pmat @ qmat).tocsr()
        else:
            _, qmat, mat = assemble_diagonal(st)
        mat = _cast(mat, dtype)
        if mode_spec.precision == "single":
            _trim(mat)
        v0 = _cast(initial_vector(st["nx"], st["ny"], 2), dtype)
        sigma = _cast(np.array([-(st["target"] ** 2)]), dtype)[0]
        if keep is not None and not keep.all():  # solver.py:506-508
            mat = mat[keep][:, keep].tocsr()
            v0 = v0[keep]
        has_pec = bool(np.any(np.abs(np.stack([e[0, 0], e[1, 1], e[2, 2]])) >= 0.9 * abs(PEC_VAL)))
        if basis_vecs is None and has_pec:
            # solver.py:467-468, 510-514, 565-566: right-Jacobi preconditioned generalized problem
            # (mat D^-1) y = lambda D^-1 y, x = D^-1 y -- the same eigenpairs of mat
            precon = sp.diags(1 / mat.diagonal()).tocsr()
            vals, vecs = spl.eigs(mat @ precon, k=m_modes, sigma=sigma, tol=tol, v0=v0, M=precon)
            vecs = precon @ vecs
        elif basis_vecs is None:
            vals, vecs = eigs(mat, sigma, v0)
        else:  # solver_eigs_relative, solver.py:750-776: dense Rayleigh-Ritz in the span of the basis
            import scipy.linalg as sl

            qb, _ = np.linalg.qr(basis_vecs)
            vals, coeffs = sl.eig(np.conj(qb.T) @ (mat @ qb))
            vecs = qb @ coeffs
        if vals.size == 0:
            raise RuntimeError("Could not find any eigenmodes for this waveguide.")
        if keep is not None and not keep.all():  # solver.py:568-569: back to the full set of unknowns (zeros on PEC cells)
            full = np.zeros((keep.size, vecs.shape[1]), dtype=vecs.dtype)
            full[keep] = vecs
            vecs = full
        root = np.emath.sqrt(-vals + 0j)  # solver.py:884
        neff, keff = root.real, root.imag
        order = np.argsort(neff)[::-1]
        neff, keff, vecs = neff[order], keff[order], vecs[:, order]
        ex, ey = vecs[:n], vecs[n:]
        hq = qmat @ vecs
        hx = hq[:n] / (1j * neff - keff)
        hy = hq[n:] / (1j * neff - keff)
        hz = (dxf @ ey - dyf @ ex) / m[2, 2][:, None]
        ez = (dxb @ hy - dyb @ hx) * (ezz_inv[:, None] if incidence else 1 / e[2, 2][:, None])
        efield = np.stack((ex, ey, ez))
        hfield = np.stack((hx, hy, hz)) * (-1j / ETA_0)
        if direction == "-":  # solver.py:370-373
            hfield[0] *= -1
            hfield[1] *= -1
            efield[2] *= -1
    else:
        eps_complex = _is_complex(e)
        spec = "tensorial_complex" if eps_complex else "tensorial_real"
        mat = assemble_tensorial(st) * (-1j)
        if direction == "-" and eps_complex:  # solver.py:669-670
            mat = mat * -1
        mat = _cast(mat.tocsr(), dtype)
        if mode_spec.precision == "single":
            _trim(mat)
        v0 = _cast(initial_vector(st["nx"], st["ny"], 4), dtype)
        sigma = _cast(np.array([st["target"]]), dtype)[0]
        vals, vecs = eigs(mat, sigma, v0)
        if vals.size == 0:
            raise RuntimeError("Could not find any eigenmodes for this waveguide.")
        neff, keff = vals.real, vals.imag
        order = np.argsort(neff)[::-1]
        neff, keff, vecs = neff[order], keff[order], vecs[:, order]
        ex, ey, hx, hy = vecs[:n], vecs[n : 2 * n], vecs[2 * n : 3 * n], vecs[3 * n :]
        hz = (dxf @ ey - dyf @ ex - m[2, 0][:, None] * hx - m[2, 1][:, None] * hy) / m[2, 2][:, None]
        ez = (dxb @ hy - dyb @ hx - e[2, 0][:, None] * ex - e[2, 1][:, None] * ey) / e[2, 2][:, None]
        efield = np.stack((ex, ey, ez))
        hfield = np.stack((hx, hy, hz)) * (-1j / ETA_0)
        if direction == "-" and not eps_complex:  # solver.py:378-380
            efield = np.conj(efield)
            hfield = -np.conj(hfield)

    # back to the original axes, E = J^T E' (solver.py:254-259)
    efield = np.einsum("ijn,inm->jnm", st["jac_e"], efield)
    if split_curl_scaling is not None:  # solver.py:904-919
        sc = np.asarray(split_curl_scaling).reshape(3, -1)
        outside = ~np.isclose(sc, 0)
        efield = efield / np.where(outside, sc, 1.0)[:, :, None] * outside[:, :, None]
    hfield = np.einsum("ijn,inm->jnm", st["jac_h"], hfield)
    shape = (3, st["nx"], st["ny"], 1, m_modes)
    fields = np.stack((efield.reshape(shape), hfield.reshape(shape)))
    if mode_spec.precision == "single":
        fields = fields.astype(np.complex64)
    if info is not None:
        info.update(applies=counter["applies"], dtype=np.dtype(dtype).name, sigma=sigma)
    return fields, (neff + 1j * keff) * st["knorm"], spec


# --------------------------------------------------------------------------------------------
# matrix-free model of the diagonal operator (pins the CUDA stencil)
# --------------------------------------------------------------------------------------------
def apply_diagonal_matrix_free(st, v):
    """A.v for v = [Ex; Ey] (shape (2N,) or (2N, k)) using only 1-D coefficient vectors."""
    nx, ny, n = st["nx"], st["ny"], st["n"]
    (xf0, xf1, xb0, xbm), (yf0, yf1, yb0, ybm) = st["coef"]
    e, m = st["eps"], st["mu"]
    v = np.asarray(v)
    k = 1 if v.ndim == 1 else v.shape[1]
    f = v.reshape(2, nx, ny, k)

    def fx(a):
        out = xf0[:, None, None] * a
        out[:-1] += xf1[:-1, None, None] * a[1:]
        return out

    def bx(a):
        out = xb0[:, None, None] * a
        out[1:] += xbm[1:, None, None] * a[:-1]
        return out

    def fy(a):
        out = yf0[None, :, None] * a
        out[:, :-1] += yf1[None, :-1, None] * a[:, 1:]
        return out

    def by(a):
        out = yb0[None, :, None] * a
        out[:, 1:] += ybm[None, 1:, None] * a[:, :-1]
        return out

    g = lambda t: t.reshape(nx, ny, 1)
    v1, v2 = f[0], f[1]
    t = (fx(v2) - fy(v1)) / g(m[2, 2])
    q1 = bx(t) + g(e[1, 1]) * v2
    q2 = by(t) - g(e[0, 0]) * v1
    u = (bx(q2) - by(q1)) / g(e[2, 2])
    p1 = fx(u) + g(m[1, 1]) * q2
    p2 = fy(u) - g(m[0, 0]) * q1
    return np.stack((p1, p2)).reshape(v.shape)
