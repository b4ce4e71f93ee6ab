// Host-side problem set-up: everything EigSolver.compute_modes does before the eigen-solve
// (tidy3d/plugins/mode/solver.py:86-217, 327-339, 389-411) reduced to what the matrix-free device
// operator needs, plus the construction of the multigrid hierarchy (ours; no reference analogue).
#pragma once
#include <cmath>
#include <complex>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200ms.h"
#include "medium.cuh"

namespace b200ms {

using cd = std::complex<double>;

// tidy3d/constants.py:16-64
constexpr double kC0 = 2.99792458e14;
constexpr double kMu0 = 1.25663706212e-12;
constexpr double kEps0 = 1.0 / (kMu0 * kC0 * kC0);
constexpr double kFpEps = 1.1920928955078125e-07;  // np.finfo(np.float32).eps
constexpr double kPecVal = -1e8;
constexpr double kTolTensorial = 1e-6;  // solver.py:22
inline double eta0() { return std::sqrt(kMu0 / kEps0); }

// One axis of one grid level.  Lengths are complex-stretched (PML) and multiplied by k0.
struct Axis {
  int n = 0;
  bool pmc = false;             // PMC instead of PEC at the min wall (solver.py:184)
  std::vector<cd> lf, lb;       // primal (forward / H-site) and dual (backward / E-site) lengths
  std::vector<double> pos;      // n+1 real node coordinates (interpolation weights only)
  // rows of the bidiagonal difference operators, derivatives.py:9-62 with S.D/k0 folded in
  // (solver.py:201):  (Df v)[i] = f0[i] v[i] + f1[i] v[i+1],  (Db v)[i] = b0[i] v[i] + bm[i] v[i-1]
  void coefficients(std::vector<cd> &c) const {
    c.assign((size_t)4 * n, cd(0, 0));
    if (n <= 1) return;  // derivatives.py:12-13: a 1-cell axis has zero derivative matrices
    for (int i = 0; i < n; ++i) {
      c[i] = -1.0 / lf[i];
      c[n + i] = (i + 1 < n) ? 1.0 / lf[i] : cd(0, 0);
      c[2 * n + i] = 1.0 / lb[i];
      c[3 * n + i] = (i > 0) ? -1.0 / lb[i] : cd(0, 0);
    }
    if (!pmc) c[0] = 0.0;                      // derivatives.py:15-16
    c[2 * n] = pmc ? 2.0 / lb[0] : cd(0, 0);   // derivatives.py:28-31
  }
};

struct ProblemSetup {
  int nx = 0, ny = 0, num_modes = 0;
  int status = B200MS_OK;
  std::string error;
  bool is_complex = false;   // eigenproblem arithmetic (solver.py:389-411)
  bool coef_complex = false; // eps/mu fields have an imaginary part
  bool tensorial = false, has_mu = false;
  bool relative = false;     // solve in the span of a supplied basis (solver_eigs_relative, solver.py:750-776)
  int eps_spec = B200MS_SPEC_DIAGONAL;
  double k0 = 0, target = 0, knorm = 1;
  cd sigma;                  // eigenvalue shift -(target^2) (solver.py:504)
  int direction = 1;
  Axis ax[2];
  std::shared_ptr<std::vector<cd>> fp[6];  // exx, eyy, ezz, mxx, myy, mzz after Jacobian + PEC model, N each
  const std::vector<cd> &f(int k) const { return *fp[k]; }
  // frequency-independent by-products kept so that a sweep over one cross-section sets the medium up once
  cd speed[4];
  std::vector<double> dlf[2], dlb[2];
  double target_raw = 0;     // target before the knorm division and nudge (solver.py:204-210)
  bool eps_complex = false, mu_complex = false;
  // tensorial path (solver.py:594-721): 18 derived coefficient fields, see tensor_fields() below
  std::shared_ptr<std::vector<cd>> ft[18];
  cd sigma_t;                // shift of the first-order operator: target n_eff (solver.py:684)
  double jac_a = 0, jac_b = 0;  // J[0][2], J[1][2] of the angled transform (transforms.py:107-108)
  std::vector<double> jz_e, jz_h;  // bend back-transform E_z *= jz_e[ix or iy], H_z *= jz_h (solver.py:254-259)
  int jz_axis = -1;          // axis along which jz varies (-1: none)
  double max_k2 = 0;         // max over cells of Re(eps) - target^2 (positive => indefinite region)
  bool has_pec = false;      // some diagonal eps entry is PEC-valued (solver.py:327-333)
  bool masked = false;       // incidence-matrix formulation with PEC cells present: PEC unknowns are removed (solver.py:441-449)
  int medium = -1;           // slot of the raw eps/mu on the device (product path)
};

inline cd s_value(double dl, double step, double omega, cd avg_speed) {
  // derivatives.py:200-232 (sigma_max = 2, kappa 1..3, order 3)
  double p = step * step * step;
  cd sig = 2.0 * avg_speed / (eta0() * dl) * p;
  return cd(1.0 + 2.0 * p, 0.0) + cd(0, 1) * sig / (omega * kEps0);
}

// derivatives.py:174-196
inline void sfactors(double omega, const std::vector<double> &dlf, const std::vector<double> &dlb, int n,
                     int npml, bool pml_at_min, cd sp_min, cd sp_max, std::vector<cd> &sf, std::vector<cd> &sb) {
  sf.assign(n, cd(1, 0));
  sb.assign(n, cd(1, 0));
  if (npml == 0) return;
  for (int i = 0; i < n; ++i) {
    if (i <= npml - 1 && pml_at_min)
      sf[i] = s_value(dlf[0], (npml - i - 0.5) / npml, omega, sp_min);
    else if (i >= n - npml)
      sf[i] = s_value(dlf[n - 1], (i - (n - npml) + 0.5) / npml, omega, sp_max);
    if (i < npml && pml_at_min)
      sb[i] = s_value(dlb[0], double(npml - i) / npml, omega, sp_min);
    else if (i > n - npml)
      sb[i] = s_value(dlb[n - 1], double(i - (n - npml)) / npml, omega, sp_max);
  }
}

// Frequency-dependent part of the set-up: PML stretch and the k0-scaled lengths of both axes.  Needs s.k0,
// s.dlf/dlb, s.speed and s.ax[a].pos.  Returns whether any derivative operator is complex (solver.py:402).
inline bool setup_axes(const b200ms_problem &p, ProblemSetup &s) {
  const int nx = s.nx, ny = s.ny;
  const double omega = 2.0 * M_PI * p.freq;
  const std::vector<double> (&dlf)[2] = s.dlf, (&dlb)[2] = s.dlb;
  const cd (&speed)[4] = s.speed;
  bool der_complex = false;
  for (int a = 0; a < 2; ++a) {
    const int nn = a == 0 ? nx : ny;
    Axis &A = s.ax[a];
    A.n = nn;
    A.pmc = p.symmetry[a] == 1;
    std::vector<cd> sf, sb;
    sfactors(omega, dlf[a], dlb[a], nn, p.num_pml[a], p.symmetry[a] == 0, speed[2 * a], speed[2 * a + 1], sf, sb);
    A.lf.resize(nn);
    A.lb.resize(nn);
    for (int i = 0; i < nn; ++i) {
      A.lf[i] = sf[i] * dlf[a][i] * s.k0;
      A.lb[i] = sb[i] * dlb[a][i] * s.k0;
    }
    // complex test on the derivative matrices (solver.py:402, 779-793) via their 1-D factors
    if (nn > 1) {
      std::vector<cd> c;
      A.coefficients(c);
      for (int half = 0; half < 2; ++half) {
        double i2 = 0, a2 = 0;
        for (int i = 0; i < 2 * nn; ++i) {
          cd v = c[(size_t)half * 2 * nn + i];
          i2 += v.imag() * v.imag();
          a2 += std::norm(v);
        }
        if (std::sqrt(i2) / (std::sqrt(a2) + kFpEps) > kFpEps) der_complex = true;
      }
    }
  }

  return der_complex;
}

struct MediumScan {
  double v[kScanSlots];
};

inline MediumParams medium_params(const ProblemSetup &s, const b200ms_problem &p, const double *de, const double *dh) {
  MediumParams m;
  m.nx = s.nx; m.ny = s.ny; m.npml_x = p.num_pml[0]; m.npml_y = p.num_pml[1];
  m.a = s.jac_a; m.b = s.jac_b; m.norm_axis = s.jz_axis; m.de = de; m.dh = dh;
  m.incidence = p.incidence ? 1 : 0;
  return m;
}

// Stage A of the set-up: everything that does not touch the per-cell medium (solver.py:86-92, 142-162, 187-190;
// transforms.py:14-111): shapes, k0, k-vector norm, coordinates, grid steps, Jacobian factors.
inline void setup_geometry(const b200ms_problem &p, ProblemSetup &s) {
  const int nx = p.nx, ny = p.ny;
  s.nx = nx;
  s.ny = ny;
  s.num_modes = p.num_modes;
  s.direction = p.direction < 0 ? -1 : 1;
  if (nx < 1 || ny < 1 || (!p.eps && !p.section) || !p.coords_x || !p.coords_y || p.num_modes < 1) {
    s.status = B200MS_ERR_ARG;
    s.error = "bad problem description";
    return;
  }
  if (!p.eps) {  // geometric cross-section (include/b200ms.h b200ms_section): every index the rasteriser follows must be in range
    const b200ms_section &q = *p.section;
    bool ok = q.nrect >= 0 && q.nmedia >= 1 && q.eps_table && (q.nrect == 0 || (q.rects && q.medium));
    for (int r = 0; ok && r < q.nrect; ++r) {
      const int kind = q.shape ? q.shape[r] : B200MS_SHAPE_RECT;
      ok = q.medium[r] >= 0 && q.medium[r] < q.nmedia && kind >= B200MS_SHAPE_RECT && kind <= B200MS_SHAPE_POLYGON;
      if (ok && kind == B200MS_SHAPE_POLYGON)
        ok = q.poly_start && q.poly_xy && q.poly_start[r] >= 0 && q.poly_start[r + 1] - q.poly_start[r] >= 3;
      if (ok && q.poly_start) ok = q.poly_start[r + 1] >= q.poly_start[r];
    }
    if (!ok) {
      s.status = B200MS_ERR_ARG;
      s.error = "bad cross-section description (medium index, shape kind or polygon range)";
      return;
    }
  }
  if (p.mu) s.has_mu = true;
  s.k0 = 2.0 * M_PI * p.freq / kC0;
  const bool bend = !std::isnan(p.bend_radius);
  const bool angled = std::abs(p.angle_theta) > 0.0;
  {  // k-vector transformation, solver.py:160-162
    double c = std::cos(p.angle_theta), sn = std::sin(p.angle_theta);
    double kxy = c * c, kz = c * sn;
    double a = kxy * std::sin(p.angle_phi), b = kxy * std::cos(p.angle_phi);
    s.knorm = std::sqrt(a * a + b * b + kz * kz);
  }
  // coordinates and Jacobian (transforms.py:14-71); only dwdz != 1 for a bend
  std::vector<double> coords[2] = {std::vector<double>(p.coords_x, p.coords_x + nx + 1),
                                   std::vector<double>(p.coords_y, p.coords_y + ny + 1)};
  if (bend) {
    const int norm_axis = (p.bend_axis == 1) ? 0 : 1;
    std::vector<double> &c = coords[norm_axis];
    const int nn = (int)c.size() - 1;
    const double off = p.bend_radius - c[nn / 2];
    for (double &v : c) v += off;
    s.jz_e.resize(nn);
    s.jz_h.resize(nn);
    for (int i = 0; i < nn; ++i) {
      s.jz_e[i] = p.bend_radius / c[i];
      s.jz_h[i] = 2.0 * p.bend_radius / (c[i] + c[i + 1]);
    }
    s.jz_axis = norm_axis;
    s.has_mu = true;
  }
  if (angled) {
    s.tensorial = true;  // J has off-diagonals (transforms.py:74-111) -> solver.py:594
    s.jac_a = -std::tan(p.angle_theta) * std::cos(p.angle_phi);
    s.jac_b = -std::tan(p.angle_theta) * std::sin(p.angle_phi);
    s.has_mu = true;
  }
  // grid steps, solver.py:187-190
  for (int a = 0; a < 2; ++a) {
    const int nn = a == 0 ? nx : ny;
    s.dlf[a].resize(nn);
    s.dlb[a].resize(nn);
    for (int i = 0; i < nn; ++i) s.dlf[a][i] = coords[a][i + 1] - coords[a][i];
    s.dlb[a][0] = s.dlf[a][0];
    for (int i = 1; i < nn; ++i) s.dlb[a][i] = 0.5 * (s.dlf[a][i - 1] + s.dlf[a][i]);
    s.ax[a].pos = coords[a];
  }
  s.relative = p.basis_e != nullptr;
}

// Stage C: everything that follows from the reductions over the medium (solver.py:204-217 target; derivatives.py:129-155
// PML strip speeds; solver.py:336-339 tensorial test; :389-411 complex test).
inline void finish_setup(const b200ms_problem &p, ProblemSetup &s, const MediumScan &sc) {
  if (s.status != B200MS_OK) return;
  const int nx = s.nx, ny = s.ny;
  if (std::isnan(p.target_neff)) s.target = std::sqrt(sc.v[SC_MAX_ABS_EPS]);
  else s.target = p.target_neff;
  s.target_raw = s.target;
  s.target /= s.knorm;
  const double shift = 10 * kFpEps;
  if (std::abs(shift) > std::abs(s.target * shift)) s.target += shift;
  else s.target *= 1 + shift;
  s.sigma = cd(-(s.target * s.target), 0.0);
  {  // average relative speed in the four PML strips: [:npml], [N-npml+1:] (derivatives.py:147-150)
    const int npx = p.num_pml[0], npy = p.num_pml[1];
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    const double cells[4] = {(double)clampi(npx, 0, nx) * ny, (double)clampi(npx - 1, 0, nx) * ny, (double)clampi(npy, 0, ny) * nx,
                             (double)clampi(npy - 1, 0, ny) * nx};
    for (int r = 0; r < 4; ++r) {
      const double cnt = 3.0 * cells[r];
      cd ea = cnt > 0 ? cd(sc.v[SC_ESUM + 2 * r], sc.v[SC_ESUM + 2 * r + 1]) / cnt : cd(1, 0);
      cd ma = cnt > 0 ? cd(sc.v[SC_MSUM + 2 * r], sc.v[SC_MSUM + 2 * r + 1]) / cnt : cd(1, 0);
      s.speed[r] = 1.0 / std::sqrt(ea * ma);
    }
  }
  const bool der_complex = setup_axes(p, s);
  if (sc.v[SC_OFF_MAX] > kTolTensorial) s.tensorial = true;
  const bool eps_complex = std::sqrt(sc.v[SC_IM2]) / (std::sqrt(sc.v[SC_ALL2]) + kFpEps) > kFpEps;
  const bool mu_complex = std::sqrt(sc.v[SC_MU_IM2]) / (std::sqrt(sc.v[SC_MU_ALL2]) + kFpEps) > kFpEps;
  s.coef_complex = eps_complex || mu_complex;
  s.is_complex = s.coef_complex || der_complex;
  if (s.relative) s.is_complex = true;  // the supplied basis is complex (solver.py:771-775)
  s.eps_complex = eps_complex;
  s.mu_complex = mu_complex;
  s.has_pec = sc.v[SC_HAS_PEC] > 0.5;
  s.masked = s.has_pec && p.incidence != 0;
  s.max_k2 = std::max(0.0, sc.v[SC_MAX_RE] - s.target * s.target);
  s.sigma_t = cd(s.target, 0.0);
  if (s.tensorial) {
    s.has_mu = true;      // the tensorial kernels always carry the six diagonal-part fields
    s.is_complex = true;  // solver.py:395-396: the tensorial matrix is always complex
    s.eps_spec = eps_complex ? B200MS_SPEC_TENSORIAL_COMPLEX : B200MS_SPEC_TENSORIAL_REAL;
    if (s.relative) {
      s.status = B200MS_ERR_UNSUPPORTED;  // solver.py:357-361
      s.error = "Tensorial eps not yet supported in relative mode solver (with basis fields provided).";
    }
  }
  if (s.masked && s.tensorial) s.masked = false;  // solver_tensorial has no incidence matrices (solver.py:594-721)
  if (s.masked && s.relative) {
    s.status = B200MS_ERR_UNSUPPORTED;  // the reference leaves the basis un-reduced there (solver.py:520-528): shape error
    s.error = "basis fields together with mu_cross / split_curl_scaling and PEC-valued cells";
  }
  if (s.relative && p.num_modes > 20) {
    s.status = B200MS_ERR_UNSUPPORTED;
    s.error = "relative mode solver: at most 20 basis modes";
  }
}

// Same medium as `ref` (same eps / coords / bend / PML / symmetry), different frequency: only the axes change.
inline void setup_problem_like(const b200ms_problem &p, const ProblemSetup &ref, ProblemSetup &s) {
  s = ref;
  s.k0 = 2.0 * M_PI * p.freq / kC0;
  const bool der_complex = setup_axes(p, s);
  s.relative = p.basis_e != nullptr;
  s.is_complex = s.coef_complex || der_complex || s.tensorial || s.relative;
}

// ---- host mirror of the device medium kernels (debug hook b200ms_debug_setup and the CPU tests) ------------------
inline void scan_medium_host(const b200ms_problem &p, const ProblemSetup &s, MediumScan &sc) {
  const MediumParams mp = medium_params(s, p, s.jz_e.data(), s.jz_h.data());
  scan_init(sc.v);
  const cplx *eps = reinterpret_cast<const cplx *>(p.eps), *mu = reinterpret_cast<const cplx *>(p.mu);
  for (int ix = 0; ix < s.nx; ++ix)
    for (int iy = 0; iy < s.ny; ++iy) scan_cell(eps, mu, mp, ix, iy, sc.v);
}
inline void fill_fields_host(const b200ms_problem &p, ProblemSetup &s) {
  const MediumParams mp = medium_params(s, p, s.jz_e.data(), s.jz_h.data());
  const size_t n = (size_t)s.nx * s.ny;
  const cplx *eps = reinterpret_cast<const cplx *>(p.eps), *mu = reinterpret_cast<const cplx *>(p.mu);
  for (int k = 0; k < 6; ++k) s.fp[k] = std::make_shared<std::vector<cd>>(n);
  if (s.tensorial)
    for (int k = 0; k < 18; ++k) s.ft[k] = std::make_shared<std::vector<cd>>(n);
  for (int ix = 0; ix < s.nx; ++ix)
    for (int iy = 0; iy < s.ny; ++iy) {
      const size_t c = (size_t)ix * s.ny + iy;
      cplx f[18];
      cell_fields(eps, mu, mp, ix, iy, f, false);
      for (int k = 0; k < 6; ++k) {
        const cplx v = (k == 2 || k == 5) ? recip(f[k]) : f[k];  // the host mirror stores ezz / mzz themselves
        (*s.fp[k])[c] = cd(v.re, v.im);
      }
      if (s.tensorial) {
        cell_tensor_fields(eps, mu, mp, ix, iy, f);
        for (int k = 0; k < 18; ++k) (*s.ft[k])[c] = cd(f[k].re, f[k].im);
      }
    }
}

inline void setup_problem(const b200ms_problem &pin, ProblemSetup &s) {
  setup_geometry(pin, s);
  if (s.status != B200MS_OK) return;
  b200ms_problem p = pin;
  std::vector<cplx> raster;
  if (!p.eps && p.section) {  // host mirror of section_raster_kernel
    const b200ms_section &q = *p.section;
    raster.resize((size_t)9 * p.nx * p.ny);
    SectionDev sd{q.nrect, q.rects, q.medium, reinterpret_cast<const cplx *>(q.eps_table), p.coords_x, p.coords_y, q.nmedia, q.shape, q.poly_start, q.poly_xy, q.site_medium};
    for (int ix = 0; ix < p.nx; ++ix)
      for (int iy = 0; iy < p.ny; ++iy) section_cell(sd, p.nx, p.ny, ix, iy, raster.data());
    p.eps = reinterpret_cast<const double *>(raster.data());
  }
  MediumScan sc;
  scan_medium_host(p, s, sc);
  finish_setup(p, s, sc);
  fill_fields_host(p, s);
}

// ------------------------------------------------------------------------------------------------
// multigrid hierarchy (geometry part, shared by every problem of a batch)
// ------------------------------------------------------------------------------------------------
struct Transfer1D {
  // prolongation: fine i <- w0*coarse[i0] + w1*coarse[i1]
  std::vector<int> p_i0, p_i1;
  std::vector<double> p_w0, p_w1;
  // restriction (normalised transpose): coarse I <- sum_k r_w[k] * fine[r_idx[k]], k in [r_ptr[I], r_ptr[I+1])
  std::vector<int> r_ptr, r_idx;
  std::vector<double> r_w;
  void build_restriction(int nc) {
    const int n = (int)p_i0.size();
    std::vector<std::vector<std::pair<int, double>>> rows(nc);
    for (int i = 0; i < n; ++i) {
      if (p_w0[i] != 0.0) rows[p_i0[i]].push_back({i, p_w0[i]});
      if (p_w1[i] != 0.0 && p_i1[i] != p_i0[i]) rows[p_i1[i]].push_back({i, p_w1[i]});
      else if (p_w1[i] != 0.0) rows[p_i0[i]].back().second += p_w1[i];
    }
    r_ptr.assign(nc + 1, 0);
    r_idx.clear();
    r_w.clear();
    for (int I = 0; I < nc; ++I) {
      double tot = 0;
      for (auto &e : rows[I]) tot += e.second;
      for (auto &e : rows[I]) {
        r_idx.push_back(e.first);
        r_w.push_back(tot > 0 ? e.second / tot : 0.0);
      }
      r_ptr[I + 1] = (int)r_idx.size();
    }
  }
};

struct AxisTransfer {
  std::vector<int> start;  // aggregate starts, size nc+1
  Transfer1D node, edge;
};

// Greedy aggregation of neighbouring cells by |stretched length| (semi-coarsening that leaves cells
// already longer than the target alone -- PML layers and coarse regions of graded meshes).
inline std::vector<int> aggregate_axis(const std::vector<cd> &lf, double H, double slack = 1.25, int max_cells = 3) {
  const int n = (int)lf.size();
  std::vector<int> st;
  int i = 0;
  while (i < n) {
    st.push_back(i);
    double tot = std::abs(lf[i]);
    int cnt = 1;
    while (i + cnt < n && cnt < max_cells && tot + std::abs(lf[i + cnt]) <= slack * H) {
      tot += std::abs(lf[i + cnt]);
      ++cnt;
      if (cnt >= 2 && tot >= 0.75 * H) break;
    }
    i += cnt;
  }
  st.push_back(n);
  return st;
}

inline void build_axis_transfer(const std::vector<int> &a, const std::vector<double> &pos, AxisTransfer &t) {
  t.start = a;
  const int nc = (int)a.size() - 1, n = a.back();
  Transfer1D &nd = t.node, &ed = t.edge;
  nd.p_i0.assign(n, 0); nd.p_i1.assign(n, 0); nd.p_w0.assign(n, 0); nd.p_w1.assign(n, 0);
  ed.p_i0.assign(n, 0); ed.p_i1.assign(n, 0); ed.p_w0.assign(n, 0); ed.p_w1.assign(n, 0);
  std::vector<double> Xc(nc);
  for (int I = 0; I < nc; ++I) Xc[I] = 0.5 * (pos[a[I]] + pos[a[I + 1]]);
  for (int I = 0; I < nc; ++I) {
    const double x0 = pos[a[I]], x1 = pos[a[I + 1]];
    for (int i = a[I]; i < a[I + 1]; ++i) {
      // node type: coarse node I sits on fine node a[I]; past the last coarse node -> zero wall
      double tt = (pos[i] - x0) / (x1 - x0);
      nd.p_i0[i] = I;
      nd.p_i1[i] = std::min(I + 1, nc - 1);
      nd.p_w0[i] = 1.0 - tt;
      nd.p_w1[i] = (I + 1 < nc) ? tt : 0.0;
      // edge type: linear between coarse cell centres, constant at the ends
      double xc = 0.5 * (pos[i] + pos[i + 1]);
      if (xc < Xc[I] && I > 0) {
        double u = (xc - Xc[I - 1]) / (Xc[I] - Xc[I - 1]);
        ed.p_i0[i] = I - 1; ed.p_i1[i] = I; ed.p_w0[i] = 1.0 - u; ed.p_w1[i] = u;
      } else if (xc > Xc[I] && I + 1 < nc) {
        double u = (xc - Xc[I]) / (Xc[I + 1] - Xc[I]);
        ed.p_i0[i] = I; ed.p_i1[i] = I + 1; ed.p_w0[i] = 1.0 - u; ed.p_w1[i] = u;
      } else {
        ed.p_i0[i] = I; ed.p_i1[i] = I; ed.p_w0[i] = 1.0; ed.p_w1[i] = 0.0;
      }
    }
  }
  nd.build_restriction(nc);
  ed.build_restriction(nc);
}

inline void coarsen_axis(const Axis &fine, const std::vector<int> &a, Axis &c) {
  const int nc = (int)a.size() - 1;
  c.n = nc;
  c.pmc = fine.pmc;
  c.lf.assign(nc, cd(0, 0));
  c.lb.assign(nc, cd(0, 0));
  c.pos.resize(nc + 1);
  for (int I = 0; I < nc; ++I) {
    for (int i = a[I]; i < a[I + 1]; ++i) c.lf[I] += fine.lf[i];
    c.pos[I] = fine.pos[a[I]];
  }
  c.pos[nc] = fine.pos[a[nc]];
  c.lb[0] = c.lf[0];
  for (int I = 1; I < nc; ++I) c.lb[I] = 0.5 * (c.lf[I - 1] + c.lf[I]);
}

struct HierarchyPlan {
  // level l has shape (nx[l], ny[l]); tr[l] maps level l+1 -> l
  std::vector<int> nx, ny;
  std::vector<AxisTransfer> trx, try_;
};

// Geometry of the hierarchy from the axes of one representative problem.  `kh_limit` > 0 stops
// coarsening once k_max * H would exceed it (indefinite shifts: coarse grids must still resolve the
// local wavelength); k_max is in the k0-scaled units of the lengths.
inline void plan_hierarchy(const Axis fine_ax[2], int min_size, int max_levels, double kh_limit, double kmax,
                           HierarchyPlan &plan, std::vector<std::vector<Axis>> *axes_out = nullptr) {
  plan = HierarchyPlan();
  plan.nx.push_back(fine_ax[0].n);
  plan.ny.push_back(fine_ax[1].n);
  Axis cur[2] = {fine_ax[0], fine_ax[1]};
  if (axes_out) axes_out->push_back({cur[0], cur[1]});
  double h0 = 1e300;
  for (int a = 0; a < 2; ++a)
    if (cur[a].n > 1)
      for (auto &v : cur[a].lf) h0 = std::min(h0, std::abs(v));
  double H = h0;
  int guard = 0;
  while ((int)plan.nx.size() < max_levels && std::max(cur[0].n, cur[1].n) > min_size && guard++ < 40) {
    H *= 2;
    if (kh_limit > 0 && kmax * H > kh_limit) break;
    std::vector<int> agg[2];
    for (int a = 0; a < 2; ++a) {
      if (cur[a].n > 1)
        agg[a] = aggregate_axis(cur[a].lf, H);
      else
        agg[a] = {0, 1};
    }
    const int ncx = (int)agg[0].size() - 1, ncy = (int)agg[1].size() - 1;
    if (ncx == cur[0].n && ncy == cur[1].n) continue;  // nothing merged at this H, try a larger one
    AxisTransfer tx, ty;
    build_axis_transfer(agg[0], cur[0].pos, tx);
    build_axis_transfer(agg[1], cur[1].pos, ty);
    plan.trx.push_back(tx);
    plan.try_.push_back(ty);
    Axis nxt[2];
    coarsen_axis(cur[0], agg[0], nxt[0]);
    coarsen_axis(cur[1], agg[1], nxt[1]);
    cur[0] = nxt[0];
    cur[1] = nxt[1];
    plan.nx.push_back(ncx);
    plan.ny.push_back(ncy);
    if (axes_out) axes_out->push_back({cur[0], cur[1]});
  }
}

}  // namespace b200ms
