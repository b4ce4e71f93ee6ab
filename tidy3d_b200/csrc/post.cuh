// On-device post-processing of the packed mode fields: what ModeSolver.data_raw does to the output of
// compute_modes before it becomes ModeSolverData, in solver-plane coordinates, without the fields leaving HBM:
//   gauge          mode_solver.py:802-810   phase such that the largest in-plane E entry is real positive
//   flux           monitor_data.py:523-539, 425-467, 582-618   colocate the tangential fields to the interior cell
//                  boundaries (linear interpolation from the Yee sites), 0.5 Re(E1 H2* - E2 H1*), trapezoid weights
//   normalise      mode_solver.py:517-521   all six components / sqrt(|flux|)
//   overlaps       monitor_data.py:640-697 (dot / outer_dot)   M x M modal overlap matrix between the modes of one problem
//                  and the next (adjacent frequencies), the input of overlap_sort (monitor_data.py:1295-1375)
// Fields layout: [2][3][nx][ny][M] (mode fastest), complex128 or complex64.
// Yee sites (c = cell centre, b = lower boundary):  Ex (c,b)  Ey (b,c)  Ez (b,b)  Hx (b,c)  Hy (c,b).
#pragma once
#include "kernels.cuh"

namespace b200ms {

constexpr int kPostChunks = 64;
constexpr int kPostMaxModes = 64;

// per-problem 1-D tables, all of length P (points along the axis): linear interpolation from centre sites and from
// boundary sites to the colocation points, and the integration weight of each point
struct PostAxis {
  const int *c_i0, *c_i1, *b_i0, *b_i1;
  const double *c_w0, *c_w1, *b_w0, *b_w1, *area;
  int P;
};
struct PostProblem {
  void *fields;  // [2][3][nx][ny][M]
  PostAxis ax, ay;
  double mult;   // 2^(number of symmetry planes)
  int flags;     // b200ms_problem.post: bit 0 gauge, bit 1 flux normalisation
  double ct, st, cp, sp;  // cos / sin of mode_spec.angle_theta, angle_phi: the TE fraction is taken in the propagation axes
  int symx, symy;         // the problem's symmetry values: which products of E components survive in the TE fraction (te_point_flags)
};

// Which products of E components the rotation to propagation axes creates survive in the reference's TE fraction: it integrates
// |E1|^2, |E2|^2 over the symmetry-EXPANDED plane (monitor_data.py:527-542, 1625-1652), where the product of two components of
// opposite parity under an active mirror cancels between a point and its mirror image.  Under the x mirror Ex is opposite to Ey and
// Ez, under the y mirror Ey is opposite to Ex and Ez (components/data/dataset.py:210-220).  A colocation point ON a symmetry
// plane (the first point along that axis, mode_solver.py:499-502) is its own image: nothing cancels there.
HD void te_point_flags(int symx, int symy, int p, int q, double &kxy, double &kxz, double &kyz) {
  kxz = (symx == 0 || p == 0) ? 1.0 : 0.0;
  kyz = (symy == 0 || q == 0) ? 1.0 : 0.0;
  kxy = kxz * kyz;
}
// |E1|^2 and |E2|^2 at one colocation point, E1 = ct (cp Ex + sp Ey) - st Ez, E2 = cp Ey - sp Ex (monitor_data.py:1603-1607 with
// the rotation matrices of components/transformation.py:112-131), written out so that the cross products can be switched off.
// Shared by post_scan_kernel and the host hook b200ms_debug_te_terms (tests/test_postprocess_pinning.py).
HD void te_tm_terms(cplx ex, cplx ey, cplx ez, double ct, double st, double cp, double sp, double kxy, double kxz, double kyz, double &te,
                    double &tm) {
  const double axx = ex.re * ex.re + ex.im * ex.im, ayy = ey.re * ey.re + ey.im * ey.im, azz = ez.re * ez.re + ez.im * ez.im;
  const double xy = ex.re * ey.re + ex.im * ey.im, xz = ex.re * ez.re + ex.im * ez.im, yz = ey.re * ez.re + ey.im * ez.im;  // Re(a b*)
  const double inplane = cp * cp * axx + sp * sp * ayy + kxy * 2.0 * cp * sp * xy;  // |cp Ex + sp Ey|^2
  te = ct * ct * inplane + st * st * azz - 2.0 * ct * st * (kxz * cp * xz + kyz * sp * yz);
  tm = cp * cp * ayy + sp * sp * axx - kxy * 2.0 * cp * sp * xy;
}

template <typename F> __device__ __forceinline__ cplx ldf(const F *p);
template <> __device__ __forceinline__ cplx ldf<cplx>(const cplx *p) { return *p; }
template <> __device__ __forceinline__ cplx ldf<cplxf>(const cplxf *p) { const cplxf v = *p; return mk((double)v.re, (double)v.im); }
__device__ __forceinline__ void stf(cplx *p, cplx v) { *p = v; }
__device__ __forceinline__ void stf(cplxf *p, cplx v) { *p = mkf((float)v.re, (float)v.im); }

// colocated value of component `comp` (0..5 = Ex,Ey,Ez,Hx,Hy,Hz) of mode m at colocation point (p, q)
template <typename F>
__device__ __forceinline__ cplx colocated(const F *f, int comp, bool xc, bool yc, const PostAxis &ax, const PostAxis &ay, int nx, int ny, int M,
                                          int p, int q, int m) {
  const int i0 = xc ? ax.c_i0[p] : ax.b_i0[p], i1 = xc ? ax.c_i1[p] : ax.b_i1[p];
  const int j0 = yc ? ay.c_i0[q] : ay.b_i0[q], j1 = yc ? ay.c_i1[q] : ay.b_i1[q];
  const double wx0 = xc ? ax.c_w0[p] : ax.b_w0[p], wx1 = xc ? ax.c_w1[p] : ax.b_w1[p];
  const double wy0 = yc ? ay.c_w0[q] : ay.b_w0[q], wy1 = yc ? ay.c_w1[q] : ay.b_w1[q];
  const F *g = f + (size_t)comp * nx * ny * M + m;
  const cplx v00 = ldf(g + ((size_t)i0 * ny + j0) * M), v01 = ldf(g + ((size_t)i0 * ny + j1) * M);
  const cplx v10 = ldf(g + ((size_t)i1 * ny + j0) * M), v11 = ldf(g + ((size_t)i1 * ny + j1) * M);
  return wx0 * (wy0 * v00 + wy1 * v01) + wx1 * (wy0 * v10 + wy1 * v11);
}

// partial[b][chunk][m] = {flux partial, max |E_inplane|^2, its raveled index (as double), re, im of that entry,
//                        int |E1|^2 dS, int |E2|^2 dS of the colocated field (pol_fraction, monitor_data.py:1625-1652)}
// GC (grid-correction factors present, mode_solver.py:847-904): slot 7 = the imaginary part of the integrated complex Poynting
// product, which the factor primal * conj(dual) mixes into the flux (post_final_kernel)
constexpr int kPostSlots = 8;
template <typename F, bool GC>
__global__ void __launch_bounds__(256) post_scan_kernel(const PostProblem *pp, int nx, int ny, int M, double *partial) {
  const PostProblem P = pp[blockIdx.y];
  const F *f = reinterpret_cast<const F *>(P.fields);
  const int chunk = blockIdx.x, nchunk = gridDim.x;
  const size_t N = (size_t)nx * ny;
  __shared__ double red[8][kPostSlots];
  for (int m = 0; m < M; ++m) {
    double fl = 0.0, best = -1.0, bidx = 0.0, bre = 0.0, bim = 0.0, te = 0.0, tm = 0.0, fli = 0.0;
    const size_t npts = (size_t)P.ax.P * P.ay.P;
    for (size_t t = (size_t)chunk * 256 + threadIdx.x; t < npts; t += (size_t)nchunk * 256) {
      const int p = (int)(t / P.ay.P), q = (int)(t % P.ay.P);
      const cplx ex = colocated(f, 0, true, false, P.ax, P.ay, nx, ny, M, p, q, m);
      const cplx ey = colocated(f, 1, false, true, P.ax, P.ay, nx, ny, M, p, q, m);
      const cplx hx = colocated(f, 3, false, true, P.ax, P.ay, nx, ny, M, p, q, m);
      const cplx hy = colocated(f, 4, true, false, P.ax, P.ay, nx, ny, M, p, q, m);
      const cplx s = ex * cj(hy) - ey * cj(hx);
      const double da = P.ax.area[p] * P.ay.area[q];
      fl += 0.5 * s.re * da;
      if constexpr (GC) fli += 0.5 * s.im * da;
      if (P.st != 0.0 || P.sp != 0.0) {
        // angled plane: the colocated E field is rotated by -phi around the normal, then by -theta around the second tangential
        // axis (monitor_data.py:1603-1607, rotation matrices components/transformation.py:112-131) before |E1|^2, |E2|^2
        const cplx ez = colocated(f, 2, false, false, P.ax, P.ay, nx, ny, M, p, q, m);
        double t1, t2;
        double kxy, kxz, kyz;
        te_point_flags(P.symx, P.symy, p, q, kxy, kxz, kyz);
        te_tm_terms(ex, ey, ez, P.ct, P.st, P.cp, P.sp, kxy, kxz, kyz, t1, t2);
        te += t1 * da;
        tm += t2 * da;
      } else {
        te += abs2(ex) * da;
        tm += abs2(ey) * da;
      }
    }
    for (size_t c = (size_t)chunk * 256 + threadIdx.x; c < 2 * N; c += (size_t)nchunk * 256) {  // E[:2] raveled: comp, ix, iy
      const cplx v = ldf(f + c * M + m);
      const double a = abs2(v);
      if (a > best) { best = a; bidx = (double)c; bre = v.re; bim = v.im; }  // c ascends per thread: first maximum kept
    }
    // block reduction: flux by sum; gauge entry by (larger |.|^2, then smaller index)
    for (int o = 16; o > 0; o >>= 1) {
      fl += __shfl_down_sync(0xffffffffu, fl, o);
      te += __shfl_down_sync(0xffffffffu, te, o);
      tm += __shfl_down_sync(0xffffffffu, tm, o);
      if constexpr (GC) fli += __shfl_down_sync(0xffffffffu, fli, o);
      const double ob = __shfl_down_sync(0xffffffffu, best, o), oi = __shfl_down_sync(0xffffffffu, bidx, o);
      const double orr = __shfl_down_sync(0xffffffffu, bre, o), oim = __shfl_down_sync(0xffffffffu, bim, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; bre = orr; bim = oim; }
    }
    if ((threadIdx.x & 31) == 0) {
      double *r = red[threadIdx.x >> 5];
      r[0] = fl; r[1] = best; r[2] = bidx; r[3] = bre; r[4] = bim; r[5] = te; r[6] = tm; r[7] = fli;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) {
        fl += red[w][0];
        te += red[w][5];
        tm += red[w][6];
        if constexpr (GC) fli += red[w][7];
        if (red[w][1] > best || (red[w][1] == best && red[w][2] < bidx)) { best = red[w][1]; bidx = red[w][2]; bre = red[w][3]; bim = red[w][4]; }
      }
      double *o = partial + (((size_t)blockIdx.y * nchunk + chunk) * M + m) * kPostSlots;
      o[0] = fl; o[1] = best; o[2] = bidx; o[3] = bre; o[4] = bim; o[5] = te; o[6] = tm; o[7] = fli;
    }
    __syncthreads();
  }
}

// flux[b][m], scal[b][m] = exp(-i phi) / sqrt(|flux|)  (either factor optional).  gc: nullptr, or [b][m] = primal * conj(dual) of the
// grid-correction factors: flux = 0.5 Re(gc * int (E1 H2* - E2 H1*) dS)  (monitor_data.py:488-503, 582-618)
__global__ void post_final_kernel(const PostProblem *pp, const double *partial, int nchunk, int M, double *flux, cplx *scal, double *te_frac,
                                  const cplx *gc) {
  const int b = blockIdx.x, m = threadIdx.x;
  if (m >= M) return;
  const int do_gauge = pp[b].flags & 1, do_norm = pp[b].flags & 2;
  double fl = 0.0, best = -1.0, bidx = 0.0, bre = 1.0, bim = 0.0, te = 0.0, tm = 0.0, fli = 0.0;
  for (int c = 0; c < nchunk; ++c) {
    const double *o = partial + (((size_t)b * nchunk + c) * M + m) * kPostSlots;
    fl += o[0];
    if (gc) fli += o[7];
    te += o[5];
    tm += o[6];
    if (o[1] > best || (o[1] == best && o[2] < bidx)) { best = o[1]; bidx = o[2]; bre = o[3]; bim = o[4]; }
  }
  if (gc) {
    const cplx g = gc[(size_t)b * M + m];
    fl = g.re * fl - g.im * fli;
  }
  fl *= pp[b].mult;
  flux[(size_t)b * M + m] = fl;
  te_frac[(size_t)b * M + m] = te / (te + tm);  // NaN for an all-zero field, like the reference
  cplx s = mk(1.0, 0.0);
  if (do_gauge && best > 0.0) {
    const double inv = rsqrt(best);
    s = mk(bre * inv, -bim * inv);  // exp(-i phi)
  }
  if (do_norm && fabs(fl) > 0.0) s = rsqrt(fabs(fl)) * s;
  scal[(size_t)b * M + m] = s;
}

template <typename F>
__global__ void __launch_bounds__(256) post_apply_kernel(const PostProblem *pp, size_t cells6, int M, const cplx *scal) {
  const int b = blockIdx.y;
  F *f = reinterpret_cast<F *>(pp[b].fields);
  __shared__ cplx s[kPostMaxModes];
  for (int m = threadIdx.x; m < M; m += 256) s[m] = scal[(size_t)b * M + m];
  __syncthreads();
  const size_t tot = cells6 * M;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < tot; e += (size_t)gridDim.x * 256) stf(f + e, s[e % M] * ldf(f + e));
}

// partial[pair][chunk][m][m'] of dot(modes of A, modes of B) = 1/4 sum (Ea* x Hb - Ha* x Eb) dS  (monitor_data.py:640-697)
struct PostPair {
  const void *a, *b;  // fields of the two problems
  int prob;           // index of problem B in the PostProblem array (its tables are used; both share the grid)
};
// GC: the two cross products are kept apart, partial[pair][chunk][2][m][m'] = {1/4 sum Ea* x Hb dS, 1/4 sum Ha* x Eb dS}: the host
// combines them with the grid-correction factors of the two data sets, conj(primal_a) dual_b and conj(dual_a) primal_b
template <typename F, bool GC>
__global__ void __launch_bounds__(256) post_dot_kernel(const PostProblem *pp, const PostPair *pairs, int nx, int ny, int M, cplx *partial) {
  const PostPair pr = pairs[blockIdx.y];
  const PostProblem P = pp[pr.prob];
  const F *fa = reinterpret_cast<const F *>(pr.a), *fb = reinterpret_cast<const F *>(pr.b);
  const int ma = blockIdx.z / M, mb = blockIdx.z % M;
  const int chunk = blockIdx.x, nchunk = gridDim.x;
  cplx acc = mk(0.0, 0.0), acc2 = mk(0.0, 0.0);
  const size_t npts = (size_t)P.ax.P * P.ay.P;
  for (size_t t = (size_t)chunk * 256 + threadIdx.x; t < npts; t += (size_t)nchunk * 256) {
    const int p = (int)(t / P.ay.P), q = (int)(t % P.ay.P);
    const cplx ea1 = cj(colocated(fa, 0, true, false, P.ax, P.ay, nx, ny, M, p, q, ma)), ea2 = cj(colocated(fa, 1, false, true, P.ax, P.ay, nx, ny, M, p, q, ma));
    const cplx ha1 = cj(colocated(fa, 3, false, true, P.ax, P.ay, nx, ny, M, p, q, ma)), ha2 = cj(colocated(fa, 4, true, false, P.ax, P.ay, nx, ny, M, p, q, ma));
    const cplx eb1 = colocated(fb, 0, true, false, P.ax, P.ay, nx, ny, M, p, q, mb), eb2 = colocated(fb, 1, false, true, P.ax, P.ay, nx, ny, M, p, q, mb);
    const cplx hb1 = colocated(fb, 3, false, true, P.ax, P.ay, nx, ny, M, p, q, mb), hb2 = colocated(fb, 4, true, false, P.ax, P.ay, nx, ny, M, p, q, mb);
    if constexpr (GC) {
      acc += (P.ax.area[p] * P.ay.area[q]) * (ea1 * hb2 - ea2 * hb1);
      acc2 += (P.ax.area[p] * P.ay.area[q]) * (ha1 * eb2 - ha2 * eb1);
    } else {
      const cplx v = (ea1 * hb2 - ea2 * hb1) - (ha1 * eb2 - ha2 * eb1);
      acc += (P.ax.area[p] * P.ay.area[q]) * v;
    }
  }
  acc = warp_sum(acc);
  if constexpr (GC) acc2 = warp_sum(acc2);
  __shared__ cplx red[8], red2[8];
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = acc;
    if constexpr (GC) red2[threadIdx.x >> 5] = acc2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cplx s = red[0];
    for (int w = 1; w < 8; ++w) s += red[w];
    if constexpr (GC) {
      cplx s2 = red2[0];
      for (int w = 1; w < 8; ++w) s2 += red2[w];
      cplx *o = partial + ((size_t)blockIdx.y * nchunk + chunk) * 2 * M * M;
      o[(size_t)ma * M + mb] = (0.25 * P.mult) * s;
      o[(size_t)M * M + (size_t)ma * M + mb] = (0.25 * P.mult) * s2;
    } else {
      partial[(((size_t)blockIdx.y * nchunk + chunk) * M + ma) * M + mb] = (0.25 * P.mult) * s;
    }
  }
}
// out[pair][e] = sum over chunks of partial[pair][chunk][e], e < E (E = M^2, or 2 M^2 with grid-correction factors)
__global__ void post_dot_final_kernel(const cplx *partial, int nchunk, int E, cplx *out) {
  const int pair = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    cplx s = mk(0.0, 0.0);
    for (int c = 0; c < nchunk; ++c) s += partial[((size_t)pair * nchunk + c) * E + e];
    out[(size_t)pair * E + e] = s;
  }
}

}  // namespace b200ms
