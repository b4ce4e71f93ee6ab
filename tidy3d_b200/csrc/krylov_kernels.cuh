// Batched Krylov-side vector kernels, round 2: 128-bit loads, several independent loads in flight per thread,
// and the reductions of one kernel folded into the prologue of the next so that a full CGS2 Gram-Schmidt step is
//   dots (read V, w)  ->  update + dots (read V, r/w w)  ->  update + norm (read V, r/w w)  ->  scale
// i.e. three reads of the basis instead of four and no stand-alone reduction launches.  Replaces the scalar
// multidot / multiaxpy kernels of round 1 on the hot path (those remain for the coarsest-level solver).
//
// Layout as everywhere: basis vector i of problem b at V + i*vstride + b*len; partial sums at
// partial[(b*nchunk + chunk)*pstride + slot].
#pragma once
#include "kernels.cuh"

namespace b200ms {

// ---- 16-byte packs -------------------------------------------------------------------------------
template <typename U> struct PackTraits;
template <> struct PackTraits<double> { static constexpr int EPV = 2; };
template <> struct PackTraits<cplx> { static constexpr int EPV = 1; };
template <> struct PackTraits<float> { static constexpr int EPV = 4; };
template <> struct PackTraits<cplxf> { static constexpr int EPV = 2; };

template <typename U, int E>
struct Pack {
  U v[E];
};

template <typename U, int E>
__device__ __forceinline__ Pack<U, E> ld_pack(const U *p) {
  Pack<U, E> r;
  if constexpr (E * sizeof(U) == 16) {
    union { float4 raw; Pack<U, E> pk; } u;
    u.raw = __ldg(reinterpret_cast<const float4 *>(p));
    r = u.pk;
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) r.v[e] = ldg(p + e);
  }
  return r;
}
template <typename U, int E>
__device__ __forceinline__ Pack<U, E> ld_pack_rw(const U *p) {  // data this kernel also writes: plain load
  Pack<U, E> r;
  if constexpr (E * sizeof(U) == 16) {
    union { float4 raw; Pack<U, E> pk; } u;
    u.raw = *reinterpret_cast<const float4 *>(p);
    r = u.pk;
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) r.v[e] = p[e];
  }
  return r;
}
template <typename U, int E>
__device__ __forceinline__ void st_pack(U *p, const Pack<U, E> &v) {
  if constexpr (E * sizeof(U) == 16) {
    union { float4 raw; Pack<U, E> pk; } u;
    u.pk = v;
    *reinterpret_cast<float4 *>(p) = u.raw;
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) p[e] = v.v[e];
  }
}

constexpr int kGsGroup = 8;     // dot products / fused-update vectors per launch
constexpr int kGsMaxCoef = 64;  // coefficients a prologue can reduce (>= m + 2, restart + 2)

// block-wide sums of NG accumulators -> dst[0..ng)
template <typename U, int NG>
__device__ __forceinline__ void block_reduce_store(U (&acc)[NG], int ng, U *dst) {
  __shared__ U red[8][NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    U v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < ng) {
    U s = zero_of<U>();
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    dst[threadIdx.x] = s;
  }
  __syncthreads();
}

// sum over chunks of partial[b][chunk][off + k], k < n  -> sh[k]  (double accumulation).  One warp per coefficient: the
// chunk loop is a strided warp reduction instead of a serial chain of up to 64 dependent loads in every CTA's prologue.
template <typename U>
__device__ __forceinline__ void reduce_partials(const U *partial, int b, int nchunk, int pstride, int off, int n, U *sh) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < n; k += nwarps) {
    cplx s = mk(0.0, 0.0);
    for (int c = lane; c < nchunk; c += 32) s += to_cplx_any(partial[((size_t)b * nchunk + c) * pstride + off + k]);
    s = warp_sum(s);
    if (lane == 0) sh[k] = from_cplx_any<U>(s);
  }
}

// K1: partial[b][chunk][poff + k] = sum_{e in chunk} conj(V_{g0+k}[b][e]) w[b][e],  k < ng <= NG
template <typename U, int EPV, int NG>
__global__ void __launch_bounds__(256) gs_dots_kernel(const U *__restrict__ V, size_t vstride, size_t len, const U *__restrict__ w,
                                                      int g0, int ng, U *partial, int pstride, int poff) {
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const size_t packs = len / EPV;
  const size_t per = (packs + nchunk - 1) / nchunk;
  const size_t p0 = (size_t)chunk * per, p1 = (p0 + per < packs) ? p0 + per : packs;
  const U *wb = w + (size_t)b * len;
  const U *vb = V + (size_t)g0 * vstride + (size_t)b * len;
  U acc[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) acc[k] = zero_of<U>();
  if (ng == NG) {  // full group: no predicates, all NG + 1 loads of a step issued back to back
#pragma unroll(NG >= 8 ? 1 : (NG >= 4 ? 2 : 4))
    for (size_t p = p0 + threadIdx.x; p < p1; p += 256) {
      const Pack<U, EPV> wv = ld_pack<U, EPV>(wb + p * EPV);
      Pack<U, EPV> vv[NG];
#pragma unroll
      for (int k = 0; k < NG; ++k) vv[k] = ld_pack<U, EPV>(vb + (size_t)k * vstride + p * EPV);
#pragma unroll
      for (int k = 0; k < NG; ++k) {
#pragma unroll
        for (int e = 0; e < EPV; ++e) acc[k] += cj(vv[k].v[e]) * wv.v[e];
      }
    }
  } else {
    for (size_t p = p0 + threadIdx.x; p < p1; p += 256) {
      const Pack<U, EPV> wv = ld_pack<U, EPV>(wb + p * EPV);
      Pack<U, EPV> vv[NG];
#pragma unroll
      for (int k = 0; k < NG; ++k)
        if (k < ng) vv[k] = ld_pack<U, EPV>(vb + (size_t)k * vstride + p * EPV);
#pragma unroll
      for (int k = 0; k < NG; ++k)
        if (k < ng) {
#pragma unroll
          for (int e = 0; e < EPV; ++e) acc[k] += cj(vv[k].v[e]) * wv.v[e];
        }
    }
  }
  block_reduce_store<U, NG>(acc, ng, partial + ((size_t)b * nchunk + chunk) * pstride + poff);
}

// K2 (nv <= kGsGroup): h = sum_chunks pin;  w -= V h;  pout[k] = partial of conj(V_k) . w_new;  chunk 0 exports h.
template <typename U, int EPV>
__global__ void __launch_bounds__(256) gs_update_dots_kernel(const U *__restrict__ V, size_t vstride, size_t len, U *w, int nv,
                                                             const U *pin, int nchunk_in, int pstride, int pin_off, U *pout,
                                                             int pout_off, U *hexp, size_t hstride, int hoff, int accumulate) {
  constexpr int NG = kGsGroup;
  __shared__ U sh[NG];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  reduce_partials<U>(pin, b, nchunk_in, pstride, pin_off, nv, sh);
  __syncthreads();
  if (chunk == 0 && hexp && threadIdx.x < nv) {
    U *d = hexp + (size_t)b * hstride + hoff + threadIdx.x;
    *d = accumulate ? (*d + sh[threadIdx.x]) : sh[threadIdx.x];
  }
  const size_t packs = len / EPV;
  const size_t per = (packs + nchunk - 1) / nchunk;
  const size_t p0 = (size_t)chunk * per, p1 = (p0 + per < packs) ? p0 + per : packs;
  U *wb = w + (size_t)b * len;
  const U *vb = V + (size_t)b * len;
  U hc[NG], acc[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    hc[k] = k < nv ? sh[k] : zero_of<U>();
    acc[k] = zero_of<U>();
  }
  for (size_t p = p0 + threadIdx.x; p < p1; p += 256) {
    Pack<U, EPV> wv = ld_pack_rw<U, EPV>(wb + p * EPV);
    Pack<U, EPV> vv[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k)
      if (k < nv) vv[k] = ld_pack<U, EPV>(vb + (size_t)k * vstride + p * EPV);
#pragma unroll
    for (int k = 0; k < NG; ++k)
      if (k < nv) {
#pragma unroll
        for (int e = 0; e < EPV; ++e) wv.v[e] -= hc[k] * vv[k].v[e];
      }
    st_pack<U, EPV>(wb + p * EPV, wv);
#pragma unroll
    for (int k = 0; k < NG; ++k)
      if (k < nv) {
#pragma unroll
        for (int e = 0; e < EPV; ++e) acc[k] += cj(vv[k].v[e]) * wv.v[e];
      }
  }
  block_reduce_store<U, NG>(acc, nv, pout + ((size_t)b * nchunk + chunk) * pstride + pout_off);
}

// K3 (any nv <= kGsMaxCoef): h = sum_chunks pin;  w -= V h;  optionally pout[pout_off] = partial of ||w_new||^2;
// chunk 0 exports h (= or +=).
template <typename U, int EPV>
__global__ void __launch_bounds__(256) gs_update_norm_kernel(const U *__restrict__ V, size_t vstride, size_t len, U *w, int nv,
                                                             const U *pin, int nchunk_in, int pstride, int pin_off, U *pout,
                                                             int pout_off, int do_norm, U *hexp, size_t hstride, int hoff,
                                                             int accumulate) {
  __shared__ U sh[kGsMaxCoef];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  reduce_partials<U>(pin, b, nchunk_in, pstride, pin_off, nv, sh);
  __syncthreads();
  if (chunk == 0 && hexp)
    for (int k = threadIdx.x; k < nv; k += 256) {
      U *d = hexp + (size_t)b * hstride + hoff + k;
      *d = accumulate ? (*d + sh[k]) : sh[k];
    }
  const size_t packs = len / EPV;
  const size_t per = (packs + nchunk - 1) / nchunk;
  const size_t p0 = (size_t)chunk * per, p1 = (p0 + per < packs) ? p0 + per : packs;
  U *wb = w + (size_t)b * len;
  const U *vb = V + (size_t)b * len;
  U nacc[1] = {zero_of<U>()};
  for (size_t p = p0 + threadIdx.x; p < p1; p += 256) {
    Pack<U, EPV> wv = ld_pack_rw<U, EPV>(wb + p * EPV);
    int k = 0;
    for (; k + 4 <= nv; k += 4) {
      Pack<U, EPV> v0 = ld_pack<U, EPV>(vb + (size_t)k * vstride + p * EPV), v1 = ld_pack<U, EPV>(vb + (size_t)(k + 1) * vstride + p * EPV),
                   v2 = ld_pack<U, EPV>(vb + (size_t)(k + 2) * vstride + p * EPV), v3 = ld_pack<U, EPV>(vb + (size_t)(k + 3) * vstride + p * EPV);
#pragma unroll
      for (int e = 0; e < EPV; ++e) wv.v[e] -= sh[k] * v0.v[e] + sh[k + 1] * v1.v[e] + sh[k + 2] * v2.v[e] + sh[k + 3] * v3.v[e];
    }
    for (; k < nv; ++k) {
      Pack<U, EPV> v0 = ld_pack<U, EPV>(vb + (size_t)k * vstride + p * EPV);
#pragma unroll
      for (int e = 0; e < EPV; ++e) wv.v[e] -= sh[k] * v0.v[e];
    }
    st_pack<U, EPV>(wb + p * EPV, wv);
    if (do_norm) {
#pragma unroll
      for (int e = 0; e < EPV; ++e) nacc[0] += from_real<U>(abs2(wv.v[e]));
    }
  }
  if (do_norm) block_reduce_store<U, 1>(nacc, 1, pout + ((size_t)b * nchunk + chunk) * pstride + pout_off);
}

// y = w / ||w|| with ||w||^2 = sum_chunks pin[pin_off] (0 -> y = 0); optional second destination y2 (the multigrid input
// buffer); block 0 exports ||w||^2.
template <typename U, int EPV>
__global__ void __launch_bounds__(256) gs_scale_kernel(const U *w, U *y, U *y2, size_t len, const U *pin, int nchunk_in,
                                                       int pstride, int pin_off, U *hexp, size_t hstride, int hoff) {
  __shared__ U sh[1];
  const int b = blockIdx.y;
  reduce_partials<U>(pin, b, nchunk_in, pstride, pin_off, 1, sh);
  __syncthreads();
  const double n2 = real_part(sh[0]);
  if (blockIdx.x == 0 && threadIdx.x == 0 && hexp) hexp[(size_t)b * hstride + hoff] = from_real<U>(n2 > 0.0 ? n2 : 0.0);
  const U al = from_real<U>(n2 > 1e-60 ? rsqrt(n2) : 0.0);
  const size_t packs = len / EPV, stride = (size_t)gridDim.x * 256;
  const U *wb = w + (size_t)b * len;
  U *yb = y + (size_t)b * len, *y2b = y2 ? y2 + (size_t)b * len : nullptr;
  for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < packs; p += stride) {
    Pack<U, EPV> v = ld_pack_rw<U, EPV>(wb + p * EPV);
#pragma unroll
    for (int e = 0; e < EPV; ++e) v.v[e] = al * v.v[e];
    st_pack<U, EPV>(yb + p * EPV, v);
    if (y2b) st_pack<U, EPV>(y2b + p * EPV, v);
  }
}

// ---- mixed-precision iterative refinement helpers --------------------------------------------------
// rp[b] = (P) (r[b] / ||r[b]||),  ||r||^2 = sum_chunks pin[pin_off];  block 0 exports ||r||^2 (double) to n2out[b].
// Problems flagged in `skip` get rp = 0.
template <typename T, typename P, int EP>
__global__ void __launch_bounds__(256) ir_begin_kernel(const T *__restrict__ r, P *rp, P *rp2, size_t len, const T *pin, int nchunk_in,
                                                       int pstride, int pin_off, double *n2out, const unsigned char *skip) {
  __shared__ T sh[1];
  const int b = blockIdx.y;
  reduce_partials<T>(pin, b, nchunk_in, pstride, pin_off, 1, sh);
  __syncthreads();
  const double n2 = real_part(sh[0]);
  if (blockIdx.x == 0 && threadIdx.x == 0 && n2out) n2out[b] = n2 > 0.0 ? n2 : 0.0;
  const double al = (n2 > 1e-300 && !(skip && skip[b])) ? rsqrt(n2) : 0.0;
  const size_t packs = len / EP, stride = (size_t)gridDim.x * 256;
  const T *rb = r + (size_t)b * len;
  P *ob = rp + (size_t)b * len, *ob2 = rp2 ? rp2 + (size_t)b * len : nullptr;
  for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < packs; p += stride) {
    Pack<P, EP> o;
#pragma unroll
    for (int e = 0; e < EP; ++e) {
      P d;
      convert(al * ldg(rb + p * EP + e), d);
      o.v[e] = d;
    }
    st_pack<P, EP>(ob + p * EP, o);
    if (ob2) st_pack<P, EP>(ob2 + p * EP, o);
  }
}

// x[b] += s_b * sum_{k<nv} y[b][k] Z_k[b],   s_b = sqrt(n2[b]) (or 1 when n2 is null).  Z, y in P; x in T.
template <typename T, typename P, int EP>
__global__ void __launch_bounds__(256) ir_update_kernel(const P *__restrict__ Z, size_t vstride, size_t len, const P *y, int ystride,
                                                        int nv, const double *n2, T *x) {
  __shared__ P sc[kGsMaxCoef];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < nv; i += 256) sc[i] = y[(size_t)b * ystride + i];
  __syncthreads();
  const double s = n2 ? sqrt(n2[b]) : 1.0;
  const size_t packs = len / EP, stride = (size_t)gridDim.x * 256;
  const P *zb = Z + (size_t)b * len;
  T *xb = x + (size_t)b * len;
  for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < packs; p += stride) {
    Pack<P, EP> acc;
#pragma unroll
    for (int e = 0; e < EP; ++e) acc.v[e] = zero_of<P>();
    int k = 0;
    for (; k + 2 <= nv; k += 2) {
      const Pack<P, EP> z0 = ld_pack<P, EP>(zb + (size_t)k * vstride + p * EP), z1 = ld_pack<P, EP>(zb + (size_t)(k + 1) * vstride + p * EP);
#pragma unroll
      for (int e = 0; e < EP; ++e) acc.v[e] += sc[k] * z0.v[e] + sc[k + 1] * z1.v[e];
    }
    for (; k < nv; ++k) {
      const Pack<P, EP> z0 = ld_pack<P, EP>(zb + (size_t)k * vstride + p * EP);
#pragma unroll
      for (int e = 0; e < EP; ++e) acc.v[e] += sc[k] * z0.v[e];
    }
#pragma unroll
    for (int e = 0; e < EP; ++e) {
      T d;
      convert(acc.v[e], d);
      T *q = xb + p * EP + e;
      *q = *q + s * d;
    }
  }
}

// out[b] = Re sum_chunks partial[b][chunk][off]
template <typename U>
__global__ void sum_partials_kernel(const U *partial, int nchunk, int pstride, int off, double *out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += real_part(partial[((size_t)b * nchunk + c) * pstride + off]);
  out[b] = s > 0.0 ? s : 0.0;
}

// ---- least-squares problem of one FGMRES cycle, one thread per problem ---------------------------------
// H: [B][restart][ld] as exported by the Gram-Schmidt kernels (column j: h_0j..h_jj, then ||w_j||^2 at row j+1).
// The right-hand side is e_1 (the cycle starts from a unit-norm residual).  Columns are processed until the
// residual estimate |g_{j+1}| falls below tol[b] (tol < 0: problem not active, y = 0) or kc columns are used.
// Output: y[b][0..kc) (zero past the columns used), res[b] = estimate, jused[b].
// work: [B][(restart+1)*restart + 3*restart + 2] cplx.
template <typename U>
__global__ void fgmres_lsq_kernel(const U *H, int restart, int ld, int kc, const double *tol, cplx *work, U *y, int ystride,
                                  double *res_out, int *jused_out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  U *yb = y + (size_t)b * ystride;
  const double tolb = tol[b];
  if (tolb < 0.0) {
    for (int i = 0; i < kc; ++i) yb[i] = zero_of<U>();
    res_out[b] = 0.0;
    jused_out[b] = 0;
    return;
  }
  const int rs1 = restart + 1;
  cplx *R = work + (size_t)b * ((size_t)rs1 * restart + 3 * restart + 2);
  cplx *cs = R + (size_t)rs1 * restart, *sn = cs + restart, *g = sn + restart;
  g[0] = mk(1.0, 0.0);
  int ju = 0;
  double res = 1.0;
  for (int j = 0; j < kc; ++j) {
    const U *col = H + ((size_t)b * restart + j) * ld;
    cplx *Rj = R + (size_t)j * rs1;
    for (int i = 0; i <= j; ++i) Rj[i] = to_cplx_any(col[i]);
    const double sub = sqrt(fmax(0.0, real_part(col[j + 1])));
    for (int i = 0; i < j; ++i) {
      const cplx a0 = Rj[i], a1 = Rj[i + 1];
      Rj[i] = cj(cs[i]) * a0 + cj(sn[i]) * a1;
      Rj[i + 1] = cs[i] * a1 - sn[i] * a0;
    }
    const cplx a0 = Rj[j];
    const double d = sqrt(abs2(a0) + sub * sub);
    cplx c = mk(1.0, 0.0), s = mk(0.0, 0.0);
    if (d > 0.0) {
      c = (1.0 / d) * a0;
      s = mk(sub / d, 0.0);
    }
    cs[j] = c;
    sn[j] = s;
    Rj[j] = mk(d, 0.0);
    g[j + 1] = -(s * g[j]);
    g[j] = cj(c) * g[j];
    ju = j + 1;
    res = sqrt(abs2(g[j + 1]));
    if (res <= tolb || !(sub > 0.0)) break;
  }
  for (int i = kc - 1; i >= ju; --i) yb[i] = zero_of<U>();
  for (int i = ju - 1; i >= 0; --i) {
    cplx acc = g[i];
    for (int j = i + 1; j < ju; ++j) acc -= R[(size_t)j * rs1 + i] * to_cplx_any(yb[j]);
    const cplx rii = R[(size_t)i * rs1 + i];
    yb[i] = abs2(rii) > 1e-300 ? from_cplx_any<U>(recip(rii) * acc) : zero_of<U>();
  }
  res_out[b] = res;
  jused_out[b] = ju;
}

}  // namespace b200ms
