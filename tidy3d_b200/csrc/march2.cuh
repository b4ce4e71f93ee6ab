// Pair-marching variant of the fused curl-curl stencil (real arithmetic, no mu fields): two grid columns per thread.
//
// Why: ncu on the one-column marching kernel (profiles/r02_sweep_full_ncu.txt) shows the fp32 multigrid sweep is bound by
// instruction issue, not by HBM (issue slots 76 % busy at 0.70 of the measured copy bandwidth): 197 SASS instructions per
// cell and row step of which ~35 are floating point -- the rest is 64-bit address arithmetic for nine scalar loads, shared-memory
// index arithmetic, eight LDS of the row coefficients, register shifts of the three-row pipeline, and a fifth CTA per 512-column
// row that is 94 % idle.  This kernel does the same march with the same arithmetic (expression for expression) but
//   * every thread owns an aligned pair of columns: 64-bit (fp32) / 128-bit (fp64) loads and stores, one 32-bit element offset
//     shared by all eleven arrays, half the shared-memory exchanges per cell (only the pair's outer neighbours come from
//     other threads);
//   * the CTA width W (threads) is a launch parameter chosen so that the strips tile the row without idle warps
//     (512 columns = one 256-thread CTA, no halo columns at all);
//   * the row loop is unrolled six-fold (the lcm of the double-buffered exchange rows and the three-row register pipeline), so
//     the pipeline shifts are register renames and the buffer indices compile-time constants;
//   * the four row coefficients arrive as one 128-bit shared-memory load;
//   * rows per CTA are a launch parameter (6 m - 3), chosen on the host against the number of resident CTAs.
// Reference for what is computed: DESIGN.md section 3 (the radius-1 form of P.Q, tidy3d/plugins/mode/solver.py:479-490).
#pragma once

namespace b200ms {

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T> struct __align__(4 * sizeof(T)) RowCoef { T f0, f1, b0, bm; };

__device__ __forceinline__ float2 ldg2(const float *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void stg2(float *p, float x, float y) { *reinterpret_cast<float2 *>(p) = make_float2(x, y); }
__device__ __forceinline__ void stg2(double *p, double x, double y) { *reinterpret_cast<double2 *>(p) = make_double2(x, y); }

constexpr int kM2MaxW = 256;      // threads per CTA (column pairs per strip)
constexpr int kM2MaxSteps = 72;   // row steps per CTA (rows per CTA + 3), multiple of 6
constexpr int kM2Unroll = 6;

// Strip geometry (host and device agree on it; tests/test_host_logic.py restates the rule): strip i covers the column pairs
// [i (W-2), i (W-2) + W); a pair is an output of its strip unless it is the strip's first pair (except in strip 0) or its last pair
// (except when it is the last pair of the row).  One strip of W >= ny/2 threads has no halo pairs at all.

template <typename T, int MODE, int PF>
__global__ void __launch_bounds__(kM2MaxW, 2) stencil_march2_kernel(StencilArgs<T, T> a, int rows) {
  static_assert(MODE == MODE_APPLY || MODE == MODE_RESID || MODE == MODE_JACOBI_D || MODE == MODE_JACOBI_D0, "march2 modes");
  static_assert(kM2Unroll % PF == 0, "prefetch depth must divide the unroll factor");
  using V = typename Vec2<T>::type;
  constexpr bool HAS_RHS = (MODE != MODE_APPLY);
  constexpr bool DSTORED = (MODE == MODE_JACOBI_D || MODE == MODE_JACOBI_D0);
  __shared__ T sB[2][kM2MaxW + 2], sV[2][kM2MaxW + 2], sU[2][kM2MaxW + 2], sTt[2][kM2MaxW + 2];
  __shared__ RowCoef<T> sX[kM2MaxSteps + 2];

  const int nx = a.nx, ny = a.ny;
  const int W = blockDim.x;
  const unsigned N = (unsigned)nx * (unsigned)ny;
  const int b = blockIdx.z;
  const int c = threadIdx.x;
  const int npairs = ny >> 1;
  const int q = (int)blockIdx.x * (W - 2) + c;  // column pair of this thread
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
  const T zT = T(0);

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
  V yf0 = {zT, zT}, yf1 = {zT, zT}, yb0 = {zT, zT}, ybm = {zT, zT};
  if (colv) { yf0 = ldg2(cy + gj); yf1 = ldg2(cy + ny + gj); yb0 = ldg2(cy + 2 * ny + gj); ybm = ldg2(cy + 3 * ny + gj); }
  const T sg = __ldg(a.sigma + b);

  // raw rows in flight (PF of them), rhs / stored diagonal of the output rows in flight
  V pv1[PF], pv2[PF], pex[PF], pey[PF], pie[PF];
  V qr1[PF], qr2[PF], qd1[PF], qd2[PF];
  const V zV = {zT, zT};
  auto load_row = [&](int gi, int slot) {
    pv1[slot] = zV; pv2[slot] = zV; pex[slot] = zV; pey[slot] = zV; pie[slot] = zV;
    if (colv && gi >= 0 && gi < nx && gi >= i0 - 1 && gi <= i0 + rows) {
      const unsigned g = (unsigned)gi * (unsigned)ny + (unsigned)gj;
      if (MODE == MODE_JACOBI_D0) {  // x0 = dinv * rhs formed on the fly
        const V da = ldg2(d1 + g), db = ldg2(d2 + g), ra = ldg2(r1 + g), rb = ldg2(r2 + g);
        pv1[slot].x = da.x * ra.x; pv1[slot].y = da.y * ra.y;
        pv2[slot].x = db.x * rb.x; pv2[slot].y = db.y * rb.y;
      } else {
        pv1[slot] = ldg2(x1 + g); pv2[slot] = ldg2(x2 + g);
      }
      pex[slot] = ldg2(exx + g); pey[slot] = ldg2(eyy + g); pie[slot] = ldg2(iez + g);
    }
  };
  auto load_out = [&](int gi, int slot) {
    if (HAS_RHS) {
      qr1[slot] = zV; qr2[slot] = zV;
      if (DSTORED) { qd1[slot] = zV; qd2[slot] = zV; }
      if (outc && gi >= i0 && gi < iend) {
        const unsigned g = (unsigned)gi * (unsigned)ny + (unsigned)gj;
        qr1[slot] = ldg2(r1 + g); qr2[slot] = ldg2(r2 + g);
        if (DSTORED) { qd1[slot] = ldg2(d1 + g); qd2[slot] = ldg2(d2 + g); }
      }
    }
  };
#pragma unroll
  for (int p = 0; p < PF; ++p) {
    load_row(k0 + 2 + p, p);
    load_out(k0 + p, p);
  }

  V a1 = zV, b1 = zV, v11 = zV, v21 = zV, a0 = zV, b0 = zV, v10 = zV, v20 = zV, ie1 = zV;
  V u0 = zV, t0 = zV, tm = zV;
  RowCoef<T> xc;  // coefficients of row k
  xc.f0 = xc.f1 = xc.b0 = xc.bm = zT;
  __syncthreads();
  xc = sX[0];

  for (int j0 = 0; j0 < nsteps; j0 += kM2Unroll) {
#pragma unroll
    for (int u = 0; u < kM2Unroll; ++u) {
      const int j = j0 + u, k = k0 + j;
      const int s = u & 1, sp = s ^ 1;
      const int slot = u % PF;
      // step 0: row k+2 out of the prefetch registers; issue the loads of row k+2+PF and of the rhs of row k+PF
      const V v12 = pv1[slot], v22 = pv2[slot];
      V a2, b2;
      a2.x = pex[slot].x * v12.x; a2.y = pex[slot].y * v12.y;
      b2.x = pey[slot].x * v22.x; b2.y = pey[slot].y * v22.y;
      const V ie2 = pie[slot];
      V cr1 = zV, cr2 = zV, cd1 = zV, cd2 = zV;
      if (HAS_RHS) { cr1 = qr1[slot]; cr2 = qr2[slot]; }
      if (DSTORED) { cd1 = qd1[slot]; cd2 = qd2[slot]; }
      load_row(k + 2 + PF, slot);
      load_out(k + PF, slot);
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
      // shift the register pipeline (renames once the loop is unrolled)
      tm = t0; t0 = t1; u0 = u1;
      a0 = a1; b0 = b1; v10 = v11; v20 = v21; a1 = a2; b1 = b2; v11 = v12; v21 = v22;
      ie1 = ie2;
      xc = xn;
      __syncthreads();
    }
  }
}

// Host side: strip width and rows per CTA.
// W: the multiple of 32 (<= 256) that covers the ny/2 column pairs with the fewest thread slots (ties: the narrower CTA).
inline void march2_strips(int ny, int &W, int &nstrips) {
  const int npairs = ny / 2;
  long best = -1;
  W = 256; nstrips = 1;
  for (int w = 32; w <= kM2MaxW; w += 32) {
    const int s = (npairs <= w) ? 1 : (npairs - 2 + (w - 3)) / (w - 2);
    const long cost = (long)s * w;
    if (best < 0 || cost < best) { best = cost; W = w; nstrips = s; }
  }
}
// rows per CTA = 6 m - 3: minimise (waves of resident CTAs) x (row steps per CTA + prologue)
inline int march2_rows(int nx, int nstrips, int B, int resident_ctas) {
  int best_rows = 9;
  double best = -1.0;
  for (int m = 2; 6 * m <= kM2MaxSteps; ++m) {
    const int rows = 6 * m - 3;
    const double ctas = (double)nstrips * ((nx + rows - 1) / rows) * B;
    const double waves = std::ceil(ctas / std::max(1, resident_ctas));
    const double cost = waves * (6 * m + 5);
    if (best < 0 || cost < best * 0.999 || (cost <= best * 1.001 && rows > best_rows && rows <= nx)) { best = cost; best_rows = rows; }
    if (rows >= nx) break;
  }
  return best_rows;
}

}  // namespace b200ms
