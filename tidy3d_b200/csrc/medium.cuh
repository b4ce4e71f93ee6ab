// Per-cell medium set-up, shared by the host mirror (debug hook / CPU tests) and the device kernels of the product path.
// Restates tidy3d/plugins/mode/solver.py:114-172 (tensor assembly, coordinate Jacobians), :204-207 (default target),
// :327-339 (PEC model, tensorial test), :389-411 (complex test), :604-653 (derived tensorial coefficient fields) and
// derivatives.py:129-155 (PML strip averages) cell by cell.  The product path uploads the raw eps_cross once per problem
// and runs medium_scan_kernel (reductions) and fields_kernel / tensor_fields_kernel (coefficient fields) on the device.
#pragma once
#include "kernels.cuh"

namespace b200ms {

constexpr double kPecValD = -1e8;  // tidy3d/constants.py pec_val

HD bool is_pec_val(cplx v) { return v.re < 0.9 * kPecValD || (v.re == 0.9 * kPecValD && v.im <= 0); }  // np.less on complex: lexicographic
HD cplx pec_model() { return mk(1.0, 1e8); }                                                        // solver.py:327-333
HD cplx cdivd(cplx a, double d) { return mk(a.re / d, a.im / d); }

// eps' = J eps J^T / det J and mu' = J mu J^T / det J (mu = identity when null) at cell c, for
// J = [[1,0,a],[0,1,b],[0,0,d]] (angled transform followed by the bend, solver.py:142-172; d differs between the E and H
// sites for a bend).  Raw tensors are component-major: eps[k*n + c], k = 3*row + col.
HD void cell_tensors(const cplx *eps, const cplx *mu, size_t n, size_t c, double a, double b, double d_e, double d_h, cplx e[9], cplx m[9]) {
  cplx raw[9], tmp[9];
  for (int k = 0; k < 9; ++k) raw[k] = eps[(size_t)k * n + c];
  const double Je[3][3] = {{1, 0, a}, {0, 1, b}, {0, 0, d_e}}, Jh[3][3] = {{1, 0, a}, {0, 1, b}, {0, 0, d_h}};
  for (int i = 0; i < 3; ++i)
    for (int q = 0; q < 3; ++q) {
      cplx acc = mk(0.0, 0.0);
      for (int j = 0; j < 3; ++j) acc += Je[i][j] * raw[3 * j + q];
      tmp[3 * i + q] = acc;
    }
  for (int i = 0; i < 3; ++i)
    for (int q = 0; q < 3; ++q) {
      cplx acc = mk(0.0, 0.0);
      for (int j = 0; j < 3; ++j) acc += tmp[3 * i + j] * Je[q][j];
      e[3 * i + q] = cdivd(acc, d_e);
    }
  if (!mu) {
    for (int i = 0; i < 3; ++i)
      for (int q = 0; q < 3; ++q) {
        double acc = 0.0;
        for (int j = 0; j < 3; ++j) acc += Jh[i][j] * Jh[q][j];
        m[3 * i + q] = mk(acc / d_h, 0.0);
      }
    return;
  }
  for (int k = 0; k < 9; ++k) raw[k] = mu[(size_t)k * n + c];
  for (int i = 0; i < 3; ++i)
    for (int q = 0; q < 3; ++q) {
      cplx acc = mk(0.0, 0.0);
      for (int j = 0; j < 3; ++j) acc += Jh[i][j] * raw[3 * j + q];
      tmp[3 * i + q] = acc;
    }
  for (int i = 0; i < 3; ++i)
    for (int q = 0; q < 3; ++q) {
      cplx acc = mk(0.0, 0.0);
      for (int j = 0; j < 3; ++j) acc += tmp[3 * i + j] * Jh[q][j];
      m[3 * i + q] = cdivd(acc, d_h);
    }
}

// Reductions over the cells of one cross-section.  Slots [0, kScanMax) are maxima, the rest are sums.
enum {
  SC_MAX_ABS_EPS = 0,  // max |eps_cross| over entries below |pec_val| (solver.py:204-207; raw input, all nine components)
  SC_OFF_MAX,          // max |off-diagonal| of eps' (PEC-modelled) and mu'  (tensorial test, solver.py:336-339)
  SC_MAX_RE,           // max Re(exx', eyy') over non-metal cells  (indefiniteness of the shifted operator; ours)
  SC_HAS_PEC,          // 1 if any diagonal entry is PEC-valued
  kScanMax,
  SC_IM2 = kScanMax, SC_ALL2, SC_MU_IM2, SC_MU_ALL2,  // ||Im||^2, ||.||^2 of eps' (PEC-modelled) and mu' (complex test, :399-401)
  SC_ESUM,                                            // 4 strips x (re, im): sum of exx'+eyy'+ezz' before the PEC model
  SC_MSUM = SC_ESUM + 8,                              // same for mu'
  kScanSlots = SC_MSUM + 8
};

struct MediumParams {
  int nx, ny, npml_x, npml_y;
  double a, b;        // angled transform J[0][2], J[1][2]
  int norm_axis;      // bend: axis along which d varies (0: x index, 1: y index), -1: no bend
  int incidence;      // solver.py:93 enable_incidence_matrices: PEC-valued unknowns are removed instead of modelled
  const double *de, *dh;  // dwdz at E / H sites along norm_axis (device or host pointers matching the caller), may be null
};

HD void scan_cell(const cplx *eps, const cplx *mu, const MediumParams &p, int ix, int iy, double acc[kScanSlots]) {
  const size_t n = (size_t)p.nx * p.ny, c = (size_t)ix * p.ny + iy;
  for (int k = 0; k < 9; ++k) {
    const double av = sqrt(abs2(eps[(size_t)k * n + c]));
    if (av < 1e8 && av > acc[SC_MAX_ABS_EPS]) acc[SC_MAX_ABS_EPS] = av;
  }
  double d_e = 1.0, d_h = 1.0;
  if (p.norm_axis >= 0) {
    const int t = p.norm_axis == 0 ? ix : iy;
    d_e = p.de[t];
    d_h = p.dh[t];
  }
  cplx e[9], m[9];
  cell_tensors(eps, mu, n, c, p.a, p.b, d_e, d_h, e, m);
  // PML strip averages use the tensors BEFORE the PEC model (derivatives.py:129-155 is called before solver_em)
  const bool in[4] = {ix < p.npml_x, ix >= p.nx - p.npml_x + 1, iy < p.npml_y, iy >= p.ny - p.npml_y + 1};
  const cplx es = e[0] + e[4] + e[8], ms = m[0] + m[4] + m[8];
  for (int r = 0; r < 4; ++r)
    if (in[r]) {
      acc[SC_ESUM + 2 * r] += es.re;
      acc[SC_ESUM + 2 * r + 1] += es.im;
      acc[SC_MSUM + 2 * r] += ms.re;
      acc[SC_MSUM + 2 * r + 1] += ms.im;
    }
  for (int k = 0; k < 9; ++k) {
    cplx v = e[k];
    const bool diag = (k == 0 || k == 4 || k == 8);
    if (is_pec_val(v)) {
      v = pec_model();
      if (diag) acc[SC_HAS_PEC] = 1.0;
    }
    acc[SC_IM2] += v.im * v.im;
    acc[SC_ALL2] += abs2(v);
    acc[SC_MU_IM2] += m[k].im * m[k].im;
    acc[SC_MU_ALL2] += abs2(m[k]);
    if (!diag) {
      const double av = fmax(sqrt(abs2(v)), sqrt(abs2(m[k])));
      if (av > acc[SC_OFF_MAX]) acc[SC_OFF_MAX] = av;
    }
    e[k] = v;
  }
  if (sqrt(abs2(e[0])) < 1e7 && sqrt(abs2(e[4])) < 1e7) {
    const double mr = fmax(e[0].re, e[4].re);
    if (mr > acc[SC_MAX_RE]) acc[SC_MAX_RE] = mr;
  }
}

HD void scan_init(double acc[kScanSlots]) {
  for (int k = 0; k < kScanSlots; ++k) acc[k] = 0.0;
  acc[SC_MAX_RE] = -1e300;
}

// one block-partial per CTA: partial[block][slot]; then medium_scan_final_kernel folds the blocks.
constexpr int kScanBlocks = 296;
__global__ void __launch_bounds__(256) medium_scan_kernel(const cplx *eps, const cplx *mu, MediumParams p, double *partial) {
  double acc[kScanSlots];
  scan_init(acc);
  const size_t n = (size_t)p.nx * p.ny;
  for (size_t c = (size_t)blockIdx.x * 256 + threadIdx.x; c < n; c += (size_t)gridDim.x * 256)
    scan_cell(eps, mu, p, (int)(c / p.ny), (int)(c % p.ny), acc);
  __shared__ double red[8][kScanSlots];
  for (int k = 0; k < kScanSlots; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) {
      const double w = __shfl_down_sync(0xffffffffu, v, o);
      v = k < kScanMax ? fmax(v, w) : v + w;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kScanSlots) {
    const int k = threadIdx.x;
    double v = red[0][k];
    for (int w = 1; w < 8; ++w) v = k < kScanMax ? fmax(v, red[w][k]) : v + red[w][k];
    partial[(size_t)blockIdx.x * kScanSlots + k] = v;
  }
}
__global__ void medium_scan_final_kernel(const double *partial, int nblocks, double *out) {
  const int k = threadIdx.x;
  if (k >= kScanSlots) return;
  double v = partial[k];
  for (int b = 1; b < nblocks; ++b) {
    const double w = partial[(size_t)b * kScanSlots + k];
    v = k < kScanMax ? fmax(v, w) : v + w;
  }
  out[k] = v;
}

// ---- coefficient fields ----------------------------------------------------------------------------------------
// exx, eyy, 1/ezz, (mxx, myy, 1/mzz) of the diagonal-path operator after Jacobian + PEC model, per cell.
// `marked` (incidence-matrix formulation, solver.py:441-449, 474-477): a PEC-valued exx / eyy is stored as 0 -- the marker the
// operator uses to drop that unknown -- and 1/ezz is zeroed on PEC-valued ezz.
HD void cell_fields(const cplx *eps, const cplx *mu, const MediumParams &p, int ix, int iy, cplx out[6], bool marked = false) {
  const size_t n = (size_t)p.nx * p.ny, c = (size_t)ix * p.ny + iy;
  double d_e = 1.0, d_h = 1.0;
  if (p.norm_axis >= 0) {
    const int t = p.norm_axis == 0 ? ix : iy;
    d_e = p.de[t];
    d_h = p.dh[t];
  }
  cplx e[9], m[9];
  cell_tensors(eps, mu, n, c, p.a, p.b, d_e, d_h, e, m);
  cplx ex = e[0], ey = e[4], ez = e[8];
  const bool px = is_pec_val(ex), py = is_pec_val(ey), pz = is_pec_val(ez);
  if (px) ex = marked ? mk(0.0, 0.0) : pec_model();
  if (py) ey = marked ? mk(0.0, 0.0) : pec_model();
  if (pz) ez = pec_model();
  out[0] = ex; out[1] = ey; out[2] = (pz && marked) ? mk(0.0, 0.0) : recip(ez);
  out[3] = m[0]; out[4] = m[4]; out[5] = recip(m[8]);
}

// 18 derived coefficient fields of the 4N first-order operator (solver.py:604-653), PEC model applied entry-wise:
// for eps (0..8) and mu (9..17): t_zx/t_zz, t_zy/t_zz, 1/t_zz, t_yz/t_zz, t_xz/t_zz, S_xx, S_xy, S_yx, S_yy.
HD void cell_tensor_fields(const cplx *eps, const cplx *mu, const MediumParams &p, int ix, int iy, cplx out[18]) {
  const size_t n = (size_t)p.nx * p.ny, c = (size_t)ix * p.ny + iy;
  double d_e = 1.0, d_h = 1.0;
  if (p.norm_axis >= 0) {
    const int t = p.norm_axis == 0 ? ix : iy;
    d_e = p.de[t];
    d_h = p.dh[t];
  }
  cplx et[9], mt[9];
  cell_tensors(eps, mu, n, c, p.a, p.b, d_e, d_h, et, mt);
  for (int q = 0; q < 9; ++q)
    if (is_pec_val(et[q])) et[q] = pec_model();
  for (int w = 0; w < 2; ++w) {
    const cplx *t = w == 0 ? et : mt;
    cplx *d = out + 9 * w;
    const cplx izz = recip(t[8]);
    d[0] = t[6] * izz;
    d[1] = t[7] * izz;
    d[2] = izz;
    d[3] = t[5] * izz;
    d[4] = t[2] * izz;
    d[5] = t[0] - t[2] * t[6] * izz;
    d[6] = t[1] - t[2] * t[7] * izz;
    d[7] = t[3] - t[5] * t[6] * izz;
    d[8] = t[4] - t[5] * t[7] * izz;
  }
}

struct MediumRef {  // one problem's raw medium on the device
  const cplx *eps, *mu;
  MediumParams p;
};

// fields[b][q][c] (C) and fields_p[b][q][c] (PC, the multigrid-precision twin; may alias) for q < nf
template <typename C, typename PC>
__global__ void __launch_bounds__(256) fields_kernel(const MediumRef *med, int nf, C *fields, PC *fields_p, size_t bstride) {
  const MediumRef r = med[blockIdx.y];
  const size_t n = (size_t)r.p.nx * r.p.ny;
  for (size_t c = (size_t)blockIdx.x * 256 + threadIdx.x; c < n; c += (size_t)gridDim.x * 256) {
    cplx f[6], fm[6];
    const bool marked = r.p.incidence != 0;
    cell_fields(r.eps, r.mu, r.p, (int)(c / r.p.ny), (int)(c % r.p.ny), f, marked);
    // the multigrid twin keeps the PEC model (a preconditioner for the full space; its output is masked afterwards)
    if (marked) cell_fields(r.eps, r.mu, r.p, (int)(c / r.p.ny), (int)(c % r.p.ny), fm, false);
    for (int q = 0; q < nf; ++q) {
      const C v = cast_to<C>(f[q]);
      fields[bstride * blockIdx.y + (size_t)q * n + c] = v;
      if ((const void *)fields_p != (const void *)fields) {
        PC w;
        convert(marked ? cast_to<C>(fm[q]) : v, w);
        fields_p[bstride * blockIdx.y + (size_t)q * n + c] = w;
      }
    }
  }
}
template <typename C>
__global__ void __launch_bounds__(256) tensor_fields_kernel(const MediumRef *med, C *ft, size_t bstride) {
  const MediumRef r = med[blockIdx.y];
  const size_t n = (size_t)r.p.nx * r.p.ny;
  for (size_t c = (size_t)blockIdx.x * 256 + threadIdx.x; c < n; c += (size_t)gridDim.x * 256) {
    cplx f[18];
    cell_tensor_fields(r.eps, r.mu, r.p, (int)(c / r.p.ny), (int)(c % r.p.ny), f);
    for (int q = 0; q < 18; ++q) ft[bstride * blockIdx.y + (size_t)q * n + c] = cast_to<C>(f[q]);
  }
}

// ---- rasterisation of a geometric cross-section (b200ms_section) into the raw eps layout [9][nx][ny] ------------------------
struct SectionDev {
  int nrect;
  const double *rects;  // nrect x 4, meaning per shape kind (include/b200ms.h)
  const int *medium;    // nrect
  const cplx *table;    // nmedia x 9
  const double *xs, *ys;  // cell boundaries, nx+1 / ny+1
  int nmedia;
  const int *shape;                  // nullptr: all rectangles
  const int *poly_start;             // nrect + 1 vertex offsets (polygons only)
  const double *poly_xy;             // (x, y) pairs
  const unsigned short *site_medium; // nullptr, or [3][nx][ny] medium found at the Ex / Ey / Ez sites before the shapes are drawn
};
// products / sums that must round like numpy's (no fused multiply-add): a site exactly on a circle must fall on the same side
// on the device, in the host mirror and in the reference
HD double section_mul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
HD double section_add(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
// even-odd crossing rule: a horizontal ray from (x, y) towards +x toggles at every edge it crosses
HD bool section_in_polygon(const double *v, int nv, double x, double y) {
  bool in = false;
  for (int i = 0, j = nv - 1; i < nv; j = i++) {
    const double xi = v[2 * i], yi = v[2 * i + 1], xj = v[2 * j], yj = v[2 * j + 1];
    if (((yi > y) != (yj > y)) && (x < section_add(section_mul(xj - xi, y - yi) / (yj - yi), xi))) in = !in;
  }
  return in;
}
HD int section_medium_at(const SectionDev &s, double x, double y, int base) {
  int med = base;
  for (int r = 0; r < s.nrect; ++r) {
    const double *q = s.rects + 4 * r;
    const int kind = s.shape ? s.shape[r] : 0;
    bool in;
    // Geometry.inside_meshgrid (geometry/base.py:195-204) evaluates `inside` only at the sites within the bounding box
    // (_inds_inside_bounds, :164-170: bounds[0] <= site <= bounds[1], bounds = centre -/+ radius or half size as computed in
    // floating point): a site on the rim can fall out by one rounding, here as there
    if (kind == 1) {  // disc: Cylinder.inside / Sphere.inside (geometry/primitives.py:600-632, 44-70)
      const double dx = fabs(x - q[0]), dy = fabs(y - q[1]);
      const bool box = q[0] - q[2] <= x && x <= q[0] + q[2] && q[1] - q[2] <= y && y <= q[1] + q[2];
      in = box && section_add(section_add(section_mul(dx, dx), section_mul(dy, dy)), section_mul(q[3], q[3])) <= section_mul(q[2], q[2]);
    } else if (kind == 2) {
      const int v0 = s.poly_start[r];
      in = section_in_polygon(s.poly_xy + 2 * (size_t)v0, s.poly_start[r + 1] - v0, x, y);
    } else {  // Box.inside (geometry/base.py:2042-2068): inclusive bounds
      const double hx = q[2] / 2, hy = q[3] / 2;
      in = q[0] - hx <= x && x <= q[0] + hx && q[1] - hy <= y && y <= q[1] + hy && fabs(x - q[0]) <= hx && fabs(y - q[1]) <= hy;
    }
    if (in) med = s.medium[r];  // later structures override (simulation.py:1199-1226)
  }
  return med;
}
HD void section_cell(const SectionDev &s, int nx, int ny, int ix, int iy, cplx *eps) {
  const size_t n = (size_t)nx * ny, c = (size_t)ix * ny + iy;
  const double xb = s.xs[ix], yb = s.ys[iy], xc = (s.xs[ix] + s.xs[ix + 1]) / 2, yc = (s.ys[iy] + s.ys[iy + 1]) / 2;
  int base[3] = {0, 0, 0};
  if (s.site_medium)
    for (int k = 0; k < 3; ++k) {
      const int m = s.site_medium[(size_t)k * n + c];
      base[k] = m < s.nmedia ? m : s.nmedia - 1;
    }
  const int med[3] = {section_medium_at(s, xc, yb, base[0]), section_medium_at(s, xb, yc, base[1]), section_medium_at(s, xb, yb, base[2])};  // Ex, Ey, Ez sites
  for (int row = 0; row < 3; ++row)
    for (int col = 0; col < 3; ++col) eps[(size_t)(3 * row + col) * n + c] = s.table[(size_t)med[row] * 9 + 3 * row + col];
}
__global__ void __launch_bounds__(256) section_raster_kernel(SectionDev s, int nx, int ny, cplx *eps) {
  const size_t n = (size_t)nx * ny;
  for (size_t c = (size_t)blockIdx.x * 256 + threadIdx.x; c < n; c += (size_t)gridDim.x * 256) section_cell(s, nx, ny, (int)(c / ny), (int)(c % ny), eps);
}

// y1 = 0 where exx == 0, y2 = 0 where eyy == 0 (the markers of removed PEC unknowns); strided two-component fields
template <typename T, typename C>
__global__ void __launch_bounds__(256) mask_kernel(T *y, size_t y_bstride, const C *fields, size_t f_bstride, size_t n) {
  const int b = blockIdx.y;
  const C *fb = fields + f_bstride * b;
  T *yb = y + y_bstride * b;
  for (size_t c = (size_t)blockIdx.x * 256 + threadIdx.x; c < n; c += (size_t)gridDim.x * 256) {
    if (abs2(fb[c]) == 0.0) yb[c] = zero_of<T>();
    if (abs2(fb[n + c]) == 0.0) yb[n + c] = zero_of<T>();
  }
}

}  // namespace b200ms
