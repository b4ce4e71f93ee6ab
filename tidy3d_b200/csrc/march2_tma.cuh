// Pair-marching stencil with TMA row staging (cp.async.bulk + mbarrier), sm_100a.
//
// Same march and the same arithmetic as stencil_march2_kernel (csrc/march2.cuh); what changes is how a row reaches the thread
// that owns it.  A strip's slice of a grid row is contiguous in HBM (W column pairs = 8 W bytes per array), so one elected
// thread moves it with a bulk asynchronous copy -- one `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` per
// array and row: v1, v2, exx, eyy, 1/ezz of row k+2 and rhs / omega-over-diagonal of row k -- into a ring of D stages in shared
// memory, and every stage completes on its own mbarrier (expect_tx = the bytes of that step).  Rows in flight cost no registers
// and no per-thread load instructions: with D = 3 two to three rows (18 KB each at W = 256) are in flight per CTA, against one
// row of prefetch registers in the register version, which is what the register version's bandwidth is limited by
// (Little's law: ~40 KB in flight per SM at 7.7 TB/s).  A stage is refilled by the elected thread right after the
// __syncthreads that ends the step which consumed it (the march needs that barrier anyway), so no "empty" barriers exist.
// Requirements beyond the register version: ny % 4 == 0 (16-byte aligned row slices).  MODE_JACOBI_D0 stays with the register
// version.  A wait that does not complete within ~2^24 polls traps instead of hanging the device.
#pragma once

namespace b200ms {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int MODE>
struct M2TmaArrays {
  static constexpr int n = 5 + (MODE != MODE_APPLY ? 2 : 0) + (MODE == MODE_JACOBI_D ? 2 : 0);
};

template <int MODE, int D>
__global__ void __launch_bounds__(kM2MaxW, 2) stencil_march2_tma_kernel(StencilArgs<float, float> a, int rows) {
  static_assert(MODE == MODE_APPLY || MODE == MODE_RESID || MODE == MODE_JACOBI_D, "TMA pair kernel: apply / residual / stored-diagonal sweep");
  static_assert(kM2Unroll % D == 0, "ring depth must divide the unroll factor");
  using T = float;
  using V = float2;
  constexpr bool HAS_RHS = (MODE != MODE_APPLY);
  constexpr bool DSTORED = (MODE == MODE_JACOBI_D);
  constexpr int NARR = M2TmaArrays<MODE>::n;
  __shared__ T sB[2][kM2MaxW + 2], sV[2][kM2MaxW + 2], sU[2][kM2MaxW + 2], sTt[2][kM2MaxW + 2];
  __shared__ RowCoef<T> sX[kM2MaxSteps + 2];
  __shared__ __align__(8) unsigned long long sBar[D];
  extern __shared__ __align__(128) unsigned char m2_ring_raw[];  // [D][NARR][W] float2
  V *ring = reinterpret_cast<V *>(m2_ring_raw);

  const int nx = a.nx, ny = a.ny;
  const int W = blockDim.x;
  const unsigned N = (unsigned)nx * (unsigned)ny;
  const int b = blockIdx.z;
  const int c = threadIdx.x;
  const int npairs = ny >> 1;
  const int q0 = (int)blockIdx.x * (W - 2);  // first column pair of the strip
  const int q = q0 + c;
  const int gj = 2 * q;
  const int i0 = blockIdx.y * rows;
  const int iend = (i0 + rows < nx) ? i0 + rows : nx;
  const int k0 = i0 - 3;
  const int nsteps = rows + 3;
  const bool colv = q < npairs;
  const bool outc = colv && (c >= 1 || blockIdx.x == 0) && (c <= W - 2 || q == npairs - 1);
  const T *x1 = a.x + (size_t)b * 2 * N, *x2 = x1 + N;
  const T *fb = a.fields + a.field_bstride * b;
  const T *exx = fb, *eyy = fb + N, *iez = fb + 2 * (size_t)N;
  const T *cx = a.cx + (size_t)b * 4 * nx, *cy = a.cy + (size_t)b * 4 * ny;
  const T *r1 = HAS_RHS ? a.rhs + (size_t)b * 2 * N : nullptr, *r2 = HAS_RHS ? r1 + N : nullptr;
  const T *d1 = DSTORED ? a.dinv + (size_t)b * 2 * N : nullptr, *d2 = DSTORED ? d1 + N : nullptr;
  T *y1 = a.y + (size_t)b * 2 * N, *y2 = y1 + N;
  const T zT = 0.0f;
  const V zV = {zT, zT};
  const int wcopy = (npairs - q0 < W) ? npairs - q0 : W;  // column pairs of this strip that exist
  const unsigned nbytes = (unsigned)wcopy * (unsigned)sizeof(V);
  const int raw_lo = (i0 - 1 > 0) ? i0 - 1 : 0, raw_hi = (i0 + rows < nx - 1) ? i0 + rows : nx - 1;

  // lanes 0 .. NARR-1 of warp 0 each own one array: lane 0 arms the stage of step jj (expect_tx = the bytes of that step), then
  // every owner lane issues the bulk copy of its array's row slice
  const T *my_arr = nullptr;
  if (c == 0) my_arr = x1; else if (c == 1) my_arr = x2; else if (c == 2) my_arr = exx; else if (c == 3) my_arr = eyy; else if (c == 4) my_arr = iez;
  else if (c == 5) my_arr = r1; else if (c == 6) my_arr = r2; else if (c == 7) my_arr = d1; else if (c == 8) my_arr = d2;
  my_arr += 2 * (size_t)q0;
  const bool my_raw = c < 5;
  auto issue = [&](int jj, int stage) {
    const int gr = k0 + jj + 2, go = k0 + jj;
    const bool rawv = gr >= raw_lo && gr <= raw_hi;
    const bool outv = HAS_RHS && go >= i0 && go < iend;
    unsigned long long *bar = &sBar[stage];
    if (c == 0) mbar_expect_tx(bar, (rawv ? 5u : 0u) * nbytes + (outv ? (unsigned)(NARR - 5) : 0u) * nbytes);
    const bool mine = my_raw ? rawv : outv;
    if (mine) bulk_g2s(ring + ((size_t)stage * NARR + c) * W, my_arr + (size_t)(my_raw ? gr : go) * ny, nbytes, bar);
  };

  if (c == 0) {
#pragma unroll
    for (int s = 0; s < D; ++s) mbar_init(&sBar[s], 1);
    mbar_fence_init();
  }
  for (int r = c; r < nsteps + 1; r += W) {  // coefficients of rows k0 .. k0 + nsteps
    const int gi = k0 + r;
    RowCoef<T> rc;
    rc.f0 = rc.f1 = rc.b0 = rc.bm = zT;
    if (gi >= 0 && gi < nx) { rc.f0 = __ldg(cx + gi); rc.f1 = __ldg(cx + nx + gi); rc.b0 = __ldg(cx + 2 * nx + gi); rc.bm = __ldg(cx + 3 * nx + gi); }
    sX[r] = rc;
  }
  for (int s = 0; s < 2; ++s) { sB[s][c + 1] = zT; sV[s][c + 1] = zT; sU[s][c + 1] = zT; sTt[s][c + 1] = zT; }
  if (c < 2)
    for (int s = 0; s < 2; ++s) {
      const int e = c == 0 ? 0 : W + 1;
      sB[s][e] = zT; sV[s][e] = zT; sU[s][e] = zT; sTt[s][e] = zT;
    }
  V yf0 = zV, yf1 = zV, yb0 = zV, ybm = zV;
  if (colv) { yf0 = ldg2(cy + gj); yf1 = ldg2(cy + ny + gj); yb0 = ldg2(cy + 2 * ny + gj); ybm = ldg2(cy + 3 * ny + gj); }
  const T sg = __ldg(a.sigma + b);
  __syncthreads();  // barriers initialised, coefficient rows and exchange rows in place
  if (c < NARR) {
#pragma unroll
    for (int s = 0; s < D; ++s)
      if (s < nsteps) issue(s, s);
  }

  V a1 = zV, b1 = zV, v11 = zV, v21 = zV, a0 = zV, b0 = zV, v10 = zV, v20 = zV, ie1 = zV;
  V u0 = zV, t0 = zV, tm = zV;
  RowCoef<T> xc = sX[0];  // coefficients of row k

  for (int j0 = 0; j0 < nsteps; j0 += kM2Unroll) {
    const unsigned parity0 = (unsigned)(j0 / D);  // ring pass of the first step of this group (kM2Unroll / D passes per group)
#pragma unroll
    for (int u = 0; u < kM2Unroll; ++u) {
      const int j = j0 + u, k = k0 + j;
      const int s = u & 1, sp = s ^ 1;
      const int stage = u % D;
      const unsigned parity = (parity0 + (unsigned)(u / D)) & 1u;
      const int gr = k + 2;
      const bool rawv = gr >= raw_lo && gr <= raw_hi;
      const bool outv = HAS_RHS && k >= i0 && k < iend;
      // step 0: row k+2 (and rhs / diagonal of row k) out of the ring
      mbar_wait(&sBar[stage], parity);
      const V *st = ring + (size_t)stage * NARR * W;
      V v12 = zV, v22 = zV, ex2 = zV, ey2 = zV, ie2 = zV;
      if (rawv && colv) { v12 = st[0 * W + c]; v22 = st[1 * W + c]; ex2 = st[2 * W + c]; ey2 = st[3 * W + c]; ie2 = st[4 * W + c]; }
      V cr1 = zV, cr2 = zV, cd1 = zV, cd2 = zV;
      if (HAS_RHS && outv && outc) {
        cr1 = st[5 * W + c]; cr2 = st[6 * W + c];
        if (DSTORED) { cd1 = st[7 * W + c]; cd2 = st[8 * W + c]; }
      }
      V a2, b2;
      a2.x = ex2.x * v12.x; a2.y = ex2.y * v12.y;
      b2.x = ey2.x * v22.x; b2.y = ey2.y * v22.y;
      // step 1: publish what the neighbouring pairs need of row k+2
      sB[s][c + 1] = b2.y;
      sV[s][c + 1] = v12.x;
      // step 2: u[k+1], t[k+1]
      const T bl = sB[sp][c], v1r = sV[sp][c + 2];
      const RowCoef<T> xn = sX[j + 1];
      V u1, t1;
      u1.x = -(ie1.x * (xn.b0 * a1.x + xn.bm * a0.x + yb0.x * b1.x + ybm.x * bl));
      u1.y = -(ie1.y * (xn.b0 * a1.y + xn.bm * a0.y + yb0.y * b1.y + ybm.y * b1.x));
      t1.x = xn.f0 * v21.x + xn.f1 * v22.x - yf0.x * v11.x - yf1.x * v11.y;
      t1.y = xn.f0 * v21.y + xn.f1 * v22.y - yf0.y * v11.y - yf1.y * v1r;
      sU[s][c + 1] = u1.x;
      sTt[s][c + 1] = t1.y;
      // step 3: outputs of row k
      if (outc && k >= i0 && k < iend) {
        const T ur = sU[sp][c + 2], tl = sTt[sp][c];
        T p1x = xc.f0 * u0.x + xc.f1 * u1.x, p1y = xc.f0 * u0.y + xc.f1 * u1.y;
        T p2x = yf0.x * u0.x + yf1.x * u0.y, p2y = yf0.y * u0.y + yf1.y * ur;
        const T c1x = yb0.x * t0.x + ybm.x * tl - a0.x, c1y = yb0.y * t0.y + ybm.y * t0.x - a0.y;
        const T c2x = xc.b0 * t0.x + xc.bm * tm.x + b0.x, c2y = xc.b0 * t0.y + xc.bm * tm.y + b0.y;
        p1x += c1x; p1y += c1y; p2x -= c2x; p2y -= c2y;
        const T o1x = p1x - sg * v10.x, o1y = p1y - sg * v10.y, o2x = p2x - sg * v20.x, o2y = p2y - sg * v20.y;
        const unsigned g = (unsigned)k * (unsigned)ny + (unsigned)gj;
        if (MODE == MODE_APPLY) {
          stg2(y1 + g, o1x, o1y); stg2(y2 + g, o2x, o2y);
        } else if (MODE == MODE_RESID) {
          stg2(y1 + g, cr1.x - o1x, cr1.y - o1y); stg2(y2 + g, cr2.x - o2x, cr2.y - o2y);
        } else {
          stg2(y1 + g, v10.x + cd1.x * (cr1.x - o1x), v10.y + cd1.y * (cr1.y - o1y));
          stg2(y2 + g, v20.x + cd2.x * (cr2.x - o2x), v20.y + cd2.y * (cr2.y - o2y));
        }
      }
      tm = t0; t0 = t1; u0 = u1;
      a0 = a1; b0 = b1; v10 = v11; v20 = v21; a1 = a2; b1 = b2; v11 = v12; v21 = v22;
      ie1 = ie2;
      xc = xn;
      __syncthreads();
      // every thread has read stage `stage`: refill it with the rows of step j + D
      if (c < NARR && j + D < nsteps) issue(j + D, stage);
    }
  }
}

}  // namespace b200ms
