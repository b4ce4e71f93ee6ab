// C ABI of libb200ms.so (see include/b200ms.h for the reference interfaces each entry point replaces).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/b200ms.h"
#include "solver.cuh"

using namespace b200ms;

struct b200ms_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  Arena arena;
  b200ms_options opt;
  std::string err;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned char *flush_buf = nullptr;
  size_t flush_bytes = 0;
  b200ms_stats stats;
};

extern "C" int b200ms_version(void) { return B200MS_VERSION; }

extern "C" void b200ms_default_options(b200ms_options *o) {
  o->eig_tol = kFpEps;   // the reference's ARPACK tolerance (TOL_EIGS = fp_eps, solver.py:20)
  o->inner_tol = 1e-8;
  o->ncv = 0;
  o->max_restarts = 100;
  o->gmres_restart = 40;
  o->gmres_maxit = 400;
  o->mg_nu = 2;
  o->mg_min_size = 12;
  o->mg_coarse_iters = 16;
  o->max_batch = 64;
  o->mg_omega = 0.8;
  o->mg_ppw = 4.0;
  o->verbose = 0;
  o->mg_pml_phase = 0.7853981633974483;
  o->stencil_variant = 0;
  o->mg_precision = 1;
  o->mg_cycles = 1;
  o->use_graph = 1;
  o->mg_nu_growth = 0;
  o->gmres_cgs2 = 2;
  o->inner_relax = 1.0;
  o->inner_relax_cap = 1e-4;
  o->inner_mode = 1;
  o->inner_ir = 1;
  o->ir_floor = 2e-5;
  o->ir_trust = 3e-5;
}

extern "C" int b200ms_create(int device, b200ms_handle **out) {
  if (!out) return B200MS_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return B200MS_ERR_CUDA;  // no CPU fallback
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return B200MS_ERR_CUDA;
  }
  if (device >= count || cudaSetDevice(device) != cudaSuccess) return B200MS_ERR_CUDA;
  auto *h = new b200ms_handle();
  h->device = device;
  b200ms_default_options(&h->opt);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
    delete h;
    return B200MS_ERR_CUDA;
  }
  *out = h;
  return B200MS_OK;
}

extern "C" int b200ms_destroy(b200ms_handle *h) {
  if (!h) return B200MS_OK;
  cudaSetDevice(h->device);
  h->arena.release();
  if (h->flush_buf) cudaFree(h->flush_buf);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return B200MS_OK;
}

extern "C" void *b200ms_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
extern "C" void b200ms_host_free(void *ptr) {
  if (ptr) cudaFreeHost(ptr);
}

extern "C" int b200ms_set_options(b200ms_handle *h, const b200ms_options *opt) {
  if (!h || !opt) return B200MS_ERR_ARG;
  h->opt = *opt;
  return B200MS_OK;
}

extern "C" const char *b200ms_last_error(b200ms_handle *h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int b200ms_get_stats(b200ms_handle *h, b200ms_stats *out) {
  if (!h || !out) return B200MS_ERR_ARG;
  *out = h->stats;
  return B200MS_OK;
}

namespace {

struct GroupKey {
  int nx, ny, k, kind, has_mu, sx, sy, jz_axis, dir, rel, tens, prec;
  double theta, phi;
  bool operator<(const GroupKey &o) const {
    return std::tie(nx, ny, k, kind, has_mu, sx, sy, jz_axis, dir, rel, tens, prec, theta, phi) <
           std::tie(o.nx, o.ny, o.k, o.kind, o.has_mu, o.sx, o.sy, o.jz_axis, o.dir, o.rel, o.tens, o.prec, o.theta, o.phi);
  }
};
struct MediumKey {
  const double *eps, *mu, *cx, *cy;
  int nx, ny, k, p0, p1, s0, s1, bend_axis, dir;
  double bend_radius, target, theta, phi;
  bool operator==(const MediumKey &o) const {
    auto same = [](double a, double b) { return (std::isnan(a) && std::isnan(b)) || a == b; };
    return eps == o.eps && mu == o.mu && cx == o.cx && cy == o.cy && nx == o.nx && ny == o.ny && k == o.k && p0 == o.p0 && p1 == o.p1 &&
           s0 == o.s0 && s1 == o.s1 && bend_axis == o.bend_axis && dir == o.dir && same(bend_radius, o.bend_radius) &&
           same(target, o.target) && theta == o.theta && phi == o.phi;
  }
};
// kind: 0 real, 1 complex vectors + real fields, 2 all complex
int kind_of(const ProblemSetup &s) { return !s.is_complex ? 0 : (s.coef_complex ? 2 : 1); }

template <typename T, typename C, typename P, typename PC>
void solve_group(b200ms_handle *h, const std::vector<int> &ids, const std::vector<ProblemSetup> &setups,
                 const b200ms_problem *prob, b200ms_result *res) {
  const int B = (int)ids.size();
  std::vector<const ProblemSetup *> ps(B);
  bool share = true;
  for (int b = 0; b < B; ++b) {
    ps[b] = &setups[ids[b]];
    const b200ms_problem &p = prob[ids[b]], &p0 = prob[ids[0]];
    if (p.eps != p0.eps || p.mu != p0.mu || p.coords_x != p0.coords_x || p.coords_y != p0.coords_y ||
        !((std::isnan(p.bend_radius) && std::isnan(p0.bend_radius)) || p.bend_radius == p0.bend_radius) ||
        p.bend_axis != p0.bend_axis)
      share = false;
  }
  BatchSolver<T, C, P, PC> S(h->arena, h->stream, h->opt);
  auto wall0 = std::chrono::steady_clock::now();
  S.single_out_ = prob[ids[0]].precision == 1;
  S.build(ps, share);
  h->stats.setup_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));  // inputs are resident in HBM from here on
  const bool real_arith = std::is_same<T, double>::value;
  const bool relative = ps[0]->relative;
  const int k = S.k;
  typename BatchSolver<T, C, P, PC>::EigResult eig;
  std::vector<cd> rel_vals;
  if (relative) {
    std::vector<const cd *> basis(B);
    for (int b = 0; b < B; ++b) basis[b] = reinterpret_cast<const cd *>(prob[ids[b]].basis_e);
    rel_vals = S.solve_relative(basis);
    eig.nconv.assign(B, k);
    eig.resid.assign(B, 0.0);
  } else {
    S.init_start_vector(0);
    eig = S.krylov_schur(real_arith);
  }
  // eigenvalues of A: lambda = sigma + 1/theta; n = sqrt(-lambda) (principal root, solver.py:884)
  std::vector<cd> nsorted((size_t)B * k), lam((size_t)B * k);
  std::vector<int> perm((size_t)B * k);
  for (int b = 0; b < B; ++b) {
    std::vector<cd> nn(k), ll(k);
    for (int q = 0; q < k; ++q) {
      if (relative) {
        ll[q] = rel_vals[(size_t)b * k + q];
      } else {
        cd th = eig.theta[(size_t)b * k + q];
        ll[q] = (S.tensor_ ? ps[b]->sigma_t : ps[b]->sigma) + (std::abs(th) > 0 ? 1.0 / th : cd(0, 0));
      }
      if (S.tensor_) {
        nn[q] = ll[q];  // the first-order operator's eigenvalue is n_eff + i k_eff itself (solver.py:879-880)
      } else {
        nn[q] = std::sqrt(-ll[q]);
        if (nn[q].real() < 0) nn[q] = -nn[q];
      }
    }
    std::vector<int> o(k);
    for (int q = 0; q < k; ++q) o[q] = q;
    std::stable_sort(o.begin(), o.end(), [&](int a, int c) { return nn[a].real() > nn[c].real(); });  // solver.py:559
    for (int q = 0; q < k; ++q) {
      perm[(size_t)b * k + q] = o[q];
      nsorted[(size_t)b * k + q] = nn[o[q]];
      lam[(size_t)b * k + q] = ll[o[q]];
    }
  }
  std::vector<cplx *> dst(B);
  bool want_fields = false;
  for (int b = 0; b < B; ++b) {
    dst[b] = reinterpret_cast<cplx *>(res[ids[b]].fields);
    if (dst[b]) want_fields = true;
  }
  S.epilogue(nsorted, perm, ps, dst.data(), want_fields);
  // true residuals on the sorted Ritz vectors (they sit in the FGMRES Z scratch after the permutation)
  std::vector<double> maxres(B, 0.0);
  {
    // reuse eigen_residuals on the permuted vectors: copy them back into the Ritz slots
    CUDA_CHECK(cudaMemcpyAsync(S.ritz_ptr(), S.gmres_z(), (size_t)k * S.vstride * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    for (int q = 0; q < k; ++q) {
      std::vector<cd> lq(B);
      for (int b = 0; b < B; ++b) lq[b] = lam[(size_t)b * k + q];
      auto r = S.eigen_residuals(q, lq);
      for (int b = 0; b < B; ++b) maxres[b] = std::max(maxres[b], r[b]);
    }
  }
  CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  CUDA_CHECK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  auto dl0 = std::chrono::steady_clock::now();
  if (want_fields) S.copy_fields_out(dst.data());
  h->stats.download_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - dl0).count();
  h->stats.device_ms += ms;
  h->stats.launches += S.stats.launches;
  h->stats.inner_iters += (long long)S.stats.inner_iters * B;
  h->stats.op_applies += (long long)S.stats.op_applies * B;
  h->stats.host_syncs += S.stats.host_syncs;
  h->stats.device_batches += 1;
  const double total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  for (int b = 0; b < B; ++b) {
    b200ms_result &r = res[ids[b]];
    for (int q = 0; q < k; ++q) {
      cd n = nsorted[(size_t)b * k + q] * ps[b]->knorm;  // solver.py:262-263
      r.n_complex[2 * q] = n.real();
      r.n_complex[2 * q + 1] = n.imag();
    }
    r.eps_spec = ps[b]->eps_spec;
    r.converged = eig.nconv[b];
    r.outer_iters = S.stats.restarts;
    r.op_applies = S.stats.op_applies;
    r.inner_iters = S.stats.inner_iters;
    r.stencil_applies = (int)std::min<long>(S.stats.launches, 2147483647L);
    r.is_complex = real_arith ? 0 : 1;
    r.solve_ms = ms;
    r.total_ms = total_ms;
    r.max_residual = maxres[b];
    // converged == Ritz residuals of OP below eig_tol (ARPACK's criterion).  The residual with respect to A itself is
    // reported for information: it is amplified by ||A - sigma|| ~ 1/(k0 dl)^2 (and by 1e8 next to PEC cells) and only
    // screened for garbage (O(1)) here.
    r.status = (relative || (eig.nconv[b] == k && eig.ok && maxres[b] < 1e-1 && S.stats.inner_failures == 0)) ? B200MS_OK : B200MS_ERR_NOCONV;
    if (h->opt.verbose)
      fprintf(stderr, "[b200ms] prob %d: conv %d/%d restarts %d op %d inner %d stencil %ld res %.2e ms %.1f\n", ids[b],
              eig.nconv[b], k, S.stats.restarts, S.stats.op_applies, S.stats.inner_iters, S.stats.stencil_applies,
              maxres[b], ms);
  }
}

size_t bytes_per_problem(const ProblemSetup &s, const b200ms_options &opt) {
  const size_t len = (size_t)2 * s.nx * s.ny;
  const size_t sv = s.is_complex ? 16 : 8;
  const int m = opt.ncv > 0 ? opt.ncv : std::max(2 * s.num_modes + 1, 20);
  size_t vecs = (m + 1) + (2 * opt.gmres_restart + 1) + 2 + s.num_modes + 6 /* level work x 4/3 */ + 3;
  return vecs * len * sv + (size_t)6 * s.nx * s.ny * s.num_modes * 16 + (size_t)8 * s.nx * s.ny * 16;
}

}  // namespace

extern "C" int b200ms_solve_batch(b200ms_handle *h, int nprob, const b200ms_problem *prob, b200ms_result *res) {
  if (!h || nprob < 0 || (nprob > 0 && (!prob || !res))) return B200MS_ERR_ARG;
  if (cudaSetDevice(h->device) != cudaSuccess) return B200MS_ERR_CUDA;
  h->err.clear();
  std::memset(&h->stats, 0, sizeof(h->stats));
  h->stats.nprob = nprob;
  const auto call0 = std::chrono::steady_clock::now();
  int first_err = B200MS_OK;
  try {
    const auto su0 = std::chrono::steady_clock::now();
    std::vector<ProblemSetup> setups(nprob);
    std::map<GroupKey, std::vector<int>> groups;
    std::vector<std::pair<MediumKey, int>> seen;
    for (int i = 0; i < nprob; ++i) {
      res[i].status = B200MS_OK;
      res[i].converged = 0;
      if (!res[i].n_complex) {
        res[i].status = B200MS_ERR_ARG;
      } else {
        // a sweep over one cross-section sets the medium up once (pointer equality == identical content)
        MediumKey mk{prob[i].eps, prob[i].mu, prob[i].coords_x, prob[i].coords_y, prob[i].nx, prob[i].ny, prob[i].num_modes,
                     prob[i].num_pml[0], prob[i].num_pml[1], prob[i].symmetry[0], prob[i].symmetry[1], prob[i].bend_axis,
                     prob[i].direction, prob[i].bend_radius, prob[i].target_neff, prob[i].angle_theta, prob[i].angle_phi};
        auto it = std::find_if(seen.begin(), seen.end(), [&](const std::pair<MediumKey, int> &e) { return e.first == mk; });
        if (it != seen.end() && setups[it->second].status == B200MS_OK) {
          setup_problem_like(prob[i], setups[it->second], setups[i]);
        } else {
          setup_problem(prob[i], setups[i]);
          seen.push_back({mk, i});
        }
        res[i].status = setups[i].status;
        res[i].eps_spec = setups[i].eps_spec;
        res[i].is_complex = setups[i].is_complex;
      }
      if (res[i].status != B200MS_OK) {
        if (first_err == B200MS_OK) {
          first_err = res[i].status;
          h->err = setups[i].error;
        }
        continue;
      }
      const ProblemSetup &s = setups[i];
      GroupKey key{s.nx, s.ny, s.num_modes, kind_of(s), s.has_mu ? 1 : 0, prob[i].symmetry[0], prob[i].symmetry[1],
                   s.jz_axis, s.direction, s.relative ? 1 : 0, s.tensorial ? (s.eps_complex ? 2 : 1) : 0, prob[i].precision == 1 ? 1 : 0,
                   prob[i].angle_theta, prob[i].angle_phi};
      groups[key].push_back(i);
    }
    h->stats.setup_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - su0).count();
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    free_b += h->arena.cap;
    for (auto &kv : groups) {
      const std::vector<int> &all = kv.second;
      const size_t per = bytes_per_problem(setups[all[0]], h->opt);
      int bmax = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, h->opt.max_batch), (size_t)(0.6 * free_b) / per));
      for (size_t s0 = 0; s0 < all.size(); s0 += bmax) {
        std::vector<int> ids(all.begin() + s0, all.begin() + std::min(all.size(), s0 + bmax));
        const bool f32 = h->opt.mg_precision == 1;
        switch (kv.first.kind) {
          case 0:
            if (f32) solve_group<double, double, float, float>(h, ids, setups, prob, res);
            else solve_group<double, double, double, double>(h, ids, setups, prob, res);
            break;
          case 1:
            if (f32) solve_group<cplx, double, cplxf, float>(h, ids, setups, prob, res);
            else solve_group<cplx, double, cplx, double>(h, ids, setups, prob, res);
            break;
          default:
            if (f32) solve_group<cplx, cplx, cplxf, cplxf>(h, ids, setups, prob, res);
            else solve_group<cplx, cplx, cplx, cplx>(h, ids, setups, prob, res);
            break;
        }
        for (int id : ids)
          if (res[id].status != B200MS_OK && first_err == B200MS_OK) {
            first_err = res[id].status;
            h->err = "eigen-iteration did not converge";
          }
      }
    }
  } catch (const std::exception &e) {
    h->err = e.what();
    return B200MS_ERR_CUDA;
  }
  h->stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - call0).count();
  return first_err;
}

// ---- benchmark hook -----------------------------------------------------------------------------------
// mode 0: the fp64 operator apply of the Krylov iteration (true PML);  mode 1: the production smoother sweep
// (stored-diagonal Jacobi in the multigrid precision, fp32 by default)
template <typename T, typename C, typename P, typename PC>
static void bench_group(b200ms_handle *h, const ProblemSetup &s, int nbatch, int mode, int nrep, int flush_l2,
                        const double *x, double *y, double *ms_out, double *bytes_out) {
  std::vector<const ProblemSetup *> ps(nbatch, &s);
  BatchSolver<T, C, P, PC> S(h->arena, h->stream, h->opt);
  S.build(ps, false);
  const size_t len = S.len;
  std::vector<T> hx(len);
  std::vector<P> hp(len);
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  for (size_t e = 0; e < len; ++e) {
    cd v = x ? cd(x[2 * e], x[2 * e + 1]) : cd(U(rng), U(rng));
    hx[e] = from_cd<T>(v);
    hp[e] = from_cd<P>(v);
  }
  T *dx = S.basis0(), *dy = S.basis0() + S.vstride, *drhs = S.basis0() + 2 * S.vstride;
  P *px = reinterpret_cast<P *>(S.basis0() + 3 * S.vstride), *py = px + S.vstride, *prhs = px + 2 * S.vstride;
  for (int b = 0; b < nbatch; ++b) {
    CUDA_CHECK(cudaMemcpyAsync(dx + (size_t)b * len, hx.data(), len * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(drhs + (size_t)b * len, hx.data(), len * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(px + (size_t)b * len, hp.data(), len * sizeof(P), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(prhs + (size_t)b * len, hp.data(), len * sizeof(P), cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (flush_l2 && !h->flush_buf) {
    h->flush_bytes = (size_t)256 << 20;
    CUDA_CHECK(cudaMalloc(&h->flush_buf, h->flush_bytes));
  }
  // mode 2: one fused CGS2 step in the multigrid precision against 4 basis vectors (slots 0..3 of the FGMRES basis)
  P *gsV = reinterpret_cast<P *>(S.gmres_v()), *gsW = gsV + 4 * S.vstride;
  if (mode == 2)
    for (int q = 0; q < 5; ++q)
      for (int b = 0; b < nbatch; ++b)
        CUDA_CHECK(cudaMemcpyAsync(gsV + (size_t)q * S.vstride + (size_t)b * len, hp.data(), len * sizeof(P), cudaMemcpyHostToDevice, h->stream));
  auto run = [&]() {
    if (mode == 2) S.template gs_cgs2<P>(gsV, S.vstride, len, 4, gsW, gsW, nullptr, reinterpret_cast<P *>(S.hessenberg()), (size_t)S.restart * (S.restart + 4));
    else if (mode == 1) S.apply(0, MODE_JACOBI, px, prhs, py);
    else S.apply_true(MODE_APPLY, dx, drhs, dy);
  };
  for (int w = 0; w < 3; ++w) run();
  double total = 0.0;
  if (flush_l2) {
    for (int r = 0; r < nrep; ++r) {
      CUDA_CHECK(cudaMemsetAsync(h->flush_buf, r & 0xff, h->flush_bytes, h->stream));
      CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
      run();
      CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
      CUDA_CHECK(cudaEventSynchronize(h->ev1));
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
      total += ms;
    }
  } else {
    CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
    for (int r = 0; r < nrep; ++r) run();
    CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
    CUDA_CHECK(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    total = ms;
  }
  CUDA_CHECK(cudaGetLastError());
  if (ms_out) *ms_out = total / nrep;
  // algorithmic bytes: apply = read v (2) + write Av (2) + nf coefficient fields per cell;
  // smoother sweep = read x, rhs, omega/diag (6) + write x' (2) + nf coefficient fields per cell
  const double ncell = (double)S.N * nbatch;
  if (bytes_out && mode == 2)  // dots: 4+1 vectors; update+dots: 4 + r/w w; update+norm: 4 + r/w w; scale: r + w
    *bytes_out = (double)len * nbatch * sizeof(P) * 19.0;
  else if (bytes_out)
    *bytes_out = mode == 1 ? ncell * (8.0 * sizeof(P) + (double)S.nf * sizeof(PC)) : ncell * (4.0 * sizeof(T) + (double)S.nf * sizeof(C));
  if (y) {
    for (size_t e = 0; e < len; ++e) {
      cd v;
      if (mode == 1) {
        P t;
        CUDA_CHECK(cudaMemcpy(&t, py + e, sizeof(P), cudaMemcpyDeviceToHost));
        v = to_cd(t);
      } else {
        T t;
        CUDA_CHECK(cudaMemcpy(&t, dy + e, sizeof(T), cudaMemcpyDeviceToHost));
        v = to_cd(t);
      }
      y[2 * e] = v.real();
      y[2 * e + 1] = v.imag();
    }
  }
}

extern "C" int b200ms_bench_stencil(b200ms_handle *h, const b200ms_problem *prob, int nbatch, int mode, int nrep,
                                    int flush_l2, const double *x, double *y, double *ms_per_launch,
                                    double *bytes_per_apply) {
  if (!h || !prob || nbatch < 1 || nrep < 1) return B200MS_ERR_ARG;
  if (cudaSetDevice(h->device) != cudaSuccess) return B200MS_ERR_CUDA;
  try {
    ProblemSetup s;
    setup_problem(*prob, s);
    if (s.status != B200MS_OK) {
      h->err = s.error;
      return s.status;
    }
    const bool f32 = h->opt.mg_precision == 1;
    switch (kind_of(s)) {
      case 0:
        if (f32) bench_group<double, double, float, float>(h, s, nbatch, mode, nrep, flush_l2, x, y, ms_per_launch, bytes_per_apply);
        else bench_group<double, double, double, double>(h, s, nbatch, mode, nrep, flush_l2, x, y, ms_per_launch, bytes_per_apply);
        break;
      case 1:
        if (f32) bench_group<cplx, double, cplxf, float>(h, s, nbatch, mode, nrep, flush_l2, x, y, ms_per_launch, bytes_per_apply);
        else bench_group<cplx, double, cplx, double>(h, s, nbatch, mode, nrep, flush_l2, x, y, ms_per_launch, bytes_per_apply);
        break;
      default:
        if (f32) bench_group<cplx, cplx, cplxf, cplxf>(h, s, nbatch, mode, nrep, flush_l2, x, y, ms_per_launch, bytes_per_apply);
        else bench_group<cplx, cplx, cplx, cplx>(h, s, nbatch, mode, nrep, flush_l2, x, y, ms_per_launch, bytes_per_apply);
        break;
    }
  } catch (const std::exception &e) {
    h->err = e.what();
    return B200MS_ERR_CUDA;
  }
  return B200MS_OK;
}

// ---- host-only debug hooks ----------------------------------------------------------------------------
extern "C" int b200ms_debug_schur(int n, const double *a, double *t, double *q) {
  if (n < 1 || !a || !t || !q) return B200MS_ERR_ARG;
  CMat A(n, n), Q;
  for (int i = 0; i < n * n; ++i) A.a[i] = cd(a[2 * i], a[2 * i + 1]);
  bool ok = schur(A, Q);
  for (int i = 0; i < n * n; ++i) {
    t[2 * i] = A.a[i].real();
    t[2 * i + 1] = A.a[i].imag();
    q[2 * i] = Q.a[i].real();
    q[2 * i + 1] = Q.a[i].imag();
  }
  return ok ? B200MS_OK : B200MS_ERR_NOCONV;
}

extern "C" int b200ms_debug_setup(const b200ms_problem *prob, double *sigma, int *flags, double *target, double *knorm,
                                  double *coef_x, double *coef_y, double *fields) {
  if (!prob) return B200MS_ERR_ARG;
  ProblemSetup s;
  setup_problem(*prob, s);
  if (s.status != B200MS_OK && s.status != B200MS_ERR_UNSUPPORTED) return s.status;
  if (sigma) {
    sigma[0] = s.sigma.real();
    sigma[1] = s.sigma.imag();
  }
  if (flags) {
    flags[0] = s.is_complex;
    flags[1] = s.tensorial;
    flags[2] = s.has_mu;
    flags[3] = s.coef_complex;
  }
  if (target) *target = s.target;
  if (knorm) *knorm = s.knorm;
  for (int a = 0; a < 2; ++a) {
    double *dst = a == 0 ? coef_x : coef_y;
    if (!dst) continue;
    std::vector<cd> c;
    s.ax[a].coefficients(c);
    for (size_t i = 0; i < c.size(); ++i) {
      dst[2 * i] = c[i].real();
      dst[2 * i + 1] = c[i].imag();
    }
  }
  if (fields) {
    const size_t n = (size_t)s.nx * s.ny;
    for (int q = 0; q < 6; ++q)
      for (size_t i = 0; i < n; ++i) {
        fields[2 * (q * n + i)] = s.f(q)[i].real();
        fields[2 * (q * n + i) + 1] = s.f(q)[i].imag();
      }
  }
  return s.status;
}

extern "C" int b200ms_debug_hierarchy(const b200ms_problem *prob, const b200ms_options *opt, int max_levels, int *shapes) {
  if (!prob || !shapes) return -1;
  b200ms_options o;
  b200ms_default_options(&o);
  if (opt) o = *opt;
  ProblemSetup s;
  setup_problem(*prob, s);
  if (s.status != B200MS_OK && s.status != B200MS_ERR_UNSUPPORTED) return -1;
  double kh_limit = 0.0, kmax = 0.0;
  if (s.max_k2 > 1e-6 && o.mg_ppw > 0) {
    kmax = std::sqrt(s.max_k2);
    kh_limit = 2.0 * M_PI / o.mg_ppw;
  }
  HierarchyPlan plan;
  plan_hierarchy(s.ax, o.mg_min_size, 12, kh_limit, kmax, plan);
  int n = std::min<int>(max_levels, (int)plan.nx.size());
  for (int l = 0; l < n; ++l) {
    shapes[2 * l] = plan.nx[l];
    shapes[2 * l + 1] = plan.ny[l];
  }
  return (int)plan.nx.size();
}

// ---- device debug hooks -------------------------------------------------------------------------------
namespace {
template <typename T, typename C, typename P, typename PC>
int debug_run(b200ms_handle *h, const ProblemSetup &s, int what, int level, int mode, const double *in0, const double *in1,
              double *out, int *iters, double *relres) {
  std::vector<const ProblemSetup *> ps(1, &s);
  BatchSolver<T, C, P, PC> S(h->arena, h->stream, h->opt);
  S.build(ps, false);
  if (what == 0 && (level > 0 || mode == 3)) {  // multigrid operator of level `level` in preconditioner precision
    if (level < 0 || level >= (int)S.lv.size()) return B200MS_ERR_ARG;
    const size_t n2 = 2 * S.lv[level].N;
    std::vector<P> a(n2), b(n2), o(n2);
    for (size_t e = 0; e < n2; ++e) {
      a[e] = from_cd<P>(cd(in0[2 * e], in0[2 * e + 1]));
      b[e] = in1 ? from_cd<P>(cd(in1[2 * e], in1[2 * e + 1])) : zero_of<P>();
    }
    P *d0 = reinterpret_cast<P *>(S.basis0()), *d1 = d0 + S.vstride, *d2 = d0 + 2 * S.vstride;
    CUDA_CHECK(cudaMemcpy(d0, a.data(), n2 * sizeof(P), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d1, b.data(), n2 * sizeof(P), cudaMemcpyHostToDevice));
    if (mode == 3) S.jacobi0(level, d1, d2);
    else S.apply(level, mode, d0, d1, d2);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(o.data(), d2, n2 * sizeof(P), cudaMemcpyDeviceToHost));
    for (size_t e = 0; e < n2; ++e) {
      cd v = to_cd(o[e]);
      out[2 * e] = v.real();
      out[2 * e + 1] = v.imag();
    }
    return B200MS_OK;
  }
  if (level < 0 || level >= (int)S.lv.size()) return B200MS_ERR_ARG;
  const size_t n2 = what == 0 ? 2 * S.lv[level].N : S.len;
  std::vector<T> a(n2), b(n2);
  for (size_t e = 0; e < n2; ++e) {
    a[e] = from_cd<T>(cd(in0[2 * e], in0[2 * e + 1]));
    b[e] = in1 ? from_cd<T>(cd(in1[2 * e], in1[2 * e + 1])) : zero_of<T>();
  }
  T *d0 = S.basis0(), *d1 = S.basis0() + S.vstride, *d2 = S.basis0() + 2 * S.vstride;
  CUDA_CHECK(cudaMemcpy(d0, a.data(), n2 * sizeof(T), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(d1, b.data(), n2 * sizeof(T), cudaMemcpyHostToDevice));
  if (what == 0) {
    S.apply_true(mode, d0, d1, d2);
  } else if (what == 1) {
    S.precondition(d0, d2);
  } else {
    int it = 0;
    double rr = S.fgmres(d0, d2, it);
    if (iters) *iters = it;
    if (relres) *relres = rr;
  }
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  CUDA_CHECK(cudaGetLastError());
  std::vector<T> o(n2);
  CUDA_CHECK(cudaMemcpy(o.data(), d2, n2 * sizeof(T), cudaMemcpyDeviceToHost));
  for (size_t e = 0; e < n2; ++e) {
    cd v = to_cd(o[e]);
    out[2 * e] = v.real();
    out[2 * e + 1] = v.imag();
  }
  return B200MS_OK;
}

int debug_dispatch(b200ms_handle *h, const b200ms_problem *prob, int what, int level, int mode, const double *in0,
                   const double *in1, double *out, int *iters, double *relres) {
  if (!h || !prob || !in0 || !out) return B200MS_ERR_ARG;
  if (cudaSetDevice(h->device) != cudaSuccess) return B200MS_ERR_CUDA;
  try {
    ProblemSetup s;
    setup_problem(*prob, s);
    if (s.status != B200MS_OK) {
      h->err = s.error;
      return s.status;
    }
    const bool f32 = h->opt.mg_precision == 1;
    switch (kind_of(s)) {
      case 0:
        return f32 ? debug_run<double, double, float, float>(h, s, what, level, mode, in0, in1, out, iters, relres)
                   : debug_run<double, double, double, double>(h, s, what, level, mode, in0, in1, out, iters, relres);
      case 1:
        return f32 ? debug_run<cplx, double, cplxf, float>(h, s, what, level, mode, in0, in1, out, iters, relres)
                   : debug_run<cplx, double, cplx, double>(h, s, what, level, mode, in0, in1, out, iters, relres);
      default:
        return f32 ? debug_run<cplx, cplx, cplxf, cplxf>(h, s, what, level, mode, in0, in1, out, iters, relres)
                   : debug_run<cplx, cplx, cplx, cplx>(h, s, what, level, mode, in0, in1, out, iters, relres);
    }
  } catch (const std::exception &e) {
    h->err = e.what();
    return B200MS_ERR_CUDA;
  }
}
}  // namespace

extern "C" int b200ms_debug_apply(b200ms_handle *h, const b200ms_problem *prob, int level, int mode, const double *x,
                                  const double *rhs, double *y) {
  return debug_dispatch(h, prob, 0, level, mode, x, rhs, y, nullptr, nullptr);
}
extern "C" int b200ms_debug_vcycle(b200ms_handle *h, const b200ms_problem *prob, const double *r, double *z) {
  return debug_dispatch(h, prob, 1, 0, 0, r, nullptr, z, nullptr, nullptr);
}
extern "C" int b200ms_debug_solve(b200ms_handle *h, const b200ms_problem *prob, const double *b, double *x, int *iters,
                                  double *relres) {
  return debug_dispatch(h, prob, 2, 0, 0, b, nullptr, x, iters, relres);
}
