// Small dense complex linear algebra for the Rayleigh-Ritz step of the Krylov-Schur iteration
// (matrices are at most ~64 x 64).  Plays the role LAPACK's zgehrd/zhseqr/ztrexc/ztrevc play inside
// ARPACK's dneupd/zneupd, which the reference reaches through scipy.sparse.linalg.eigs
// (tidy3d/plugins/mode/solver.py:744-746).  Host code, no dependencies.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace b200ms {

using cd = std::complex<double>;

struct CMat {
  int rows = 0, cols = 0;
  std::vector<cd> a;
  CMat() {}
  CMat(int r, int c) : rows(r), cols(c), a((size_t)r * c, cd(0, 0)) {}
  cd &operator()(int i, int j) { return a[(size_t)i * cols + j]; }
  const cd &operator()(int i, int j) const { return a[(size_t)i * cols + j]; }
  static CMat identity(int n) {
    CMat m(n, n);
    for (int i = 0; i < n; ++i) m(i, i) = 1.0;
    return m;
  }
};

// Givens rotation G = [c s; -conj(s) c] (c real) with G [x; y] = [r; 0].
struct Givens {
  double c;
  cd s;
  Givens(cd x, cd y) {
    double ax = std::abs(x), ay = std::abs(y);
    if (ay == 0.0) {
      c = 1.0;
      s = 0.0;
    } else if (ax == 0.0) {
      c = 0.0;
      s = std::conj(y) / ay;
    } else {
      double nrm = std::hypot(ax, ay);
      c = ax / nrm;
      s = (x / ax) * std::conj(y) / nrm;
    }
  }
  // rows (p, q) of M <- G * rows, columns j0..j1-1
  void rows(CMat &m, int p, int q, int j0, int j1) const {
    for (int j = j0; j < j1; ++j) {
      cd a = m(p, j), b = m(q, j);
      m(p, j) = c * a + s * b;
      m(q, j) = -std::conj(s) * a + c * b;
    }
  }
  // columns (p, q) of M <- cols * G^H, rows i0..i1-1
  void cols(CMat &m, int p, int q, int i0, int i1) const {
    for (int i = i0; i < i1; ++i) {
      cd a = m(i, p), b = m(i, q);
      m(i, p) = c * a + std::conj(s) * b;
      m(i, q) = -s * a + c * b;
    }
  }
};

// A <- Q^H A Q upper Hessenberg, Q accumulated (must enter as identity or a prior transform).
inline void hessenberg(CMat &a, CMat &q) {
  const int n = a.rows;
  for (int j = 0; j + 2 < n; ++j)
    for (int i = j + 2; i < n; ++i) {
      if (a(i, j) == cd(0, 0)) continue;
      Givens g(a(j + 1, j), a(i, j));
      g.rows(a, j + 1, i, 0, n);
      g.cols(a, j + 1, i, 0, n);
      g.cols(q, j + 1, i, 0, q.rows);
      a(i, j) = 0.0;
    }
}

// Shifted QR iteration on an upper Hessenberg matrix: h <- T (upper triangular), q <- q * Z.
// Returns false if an eigenvalue failed to converge.
inline bool schur_hessenberg(CMat &h, CMat &q) {
  const int n = h.rows;
  const double eps = 2.220446049250313e-16;
  int hi = n - 1;
  int iter = 0;
  while (hi >= 1) {
    // look for a negligible sub-diagonal
    int lo = hi;
    while (lo > 0) {
      double s = std::abs(h(lo - 1, lo - 1)) + std::abs(h(lo, lo));
      if (s == 0.0) s = 1.0;
      if (std::abs(h(lo, lo - 1)) <= eps * s) {
        h(lo, lo - 1) = 0.0;
        break;
      }
      --lo;
    }
    if (lo == hi) {
      --hi;
      iter = 0;
      continue;
    }
    if (++iter > 60 * n) return false;
    // Wilkinson shift: eigenvalue of the trailing 2x2 closest to h(hi,hi)
    cd mu;
    if (iter % 11 == 10) {
      mu = h(hi, hi) + cd(0.75 * std::abs(h(hi, hi - 1)), 0.0);  // exceptional shift
    } else {
      cd a = h(hi - 1, hi - 1), b = h(hi - 1, hi), c = h(hi, hi - 1), d = h(hi, hi);
      cd tr = a + d, det = a * d - b * c;
      cd disc = std::sqrt(tr * tr * 0.25 - det);
      cd l1 = tr * 0.5 + disc, l2 = tr * 0.5 - disc;
      mu = (std::abs(l1 - d) < std::abs(l2 - d)) ? l1 : l2;
    }
    // implicit single-shift QR sweep on the active block [lo, hi]
    for (int k = lo; k < hi; ++k) {
      cd x = (k == lo) ? h(lo, lo) - mu : h(k, k - 1);
      cd y = (k == lo) ? h(lo + 1, lo) : h(k + 1, k - 1);
      Givens g(x, y);
      g.rows(h, k, k + 1, (k == lo) ? lo : k - 1, n);
      if (k > lo) h(k + 1, k - 1) = 0.0;
      g.cols(h, k, k + 1, 0, std::min(k + 3, hi + 1));
      g.cols(q, k, k + 1, 0, q.rows);
    }
  }
  for (int i = 1; i < n; ++i)
    for (int j = 0; j < i; ++j) h(i, j) = 0.0;
  return true;
}

// General complex Schur decomposition a = q t q^H.  `a` is overwritten by t.
inline bool schur(CMat &a, CMat &q) {
  q = CMat::identity(a.rows);
  hessenberg(a, q);
  return schur_hessenberg(a, q);
}

// Swap the diagonal entries k and k+1 of the upper-triangular t, updating q (ztrexc step).
inline void schur_swap(CMat &t, CMat &q, int k) {
  const int n = t.rows;
  cd t11 = t(k, k), t22 = t(k + 1, k + 1);
  Givens g(t(k, k + 1), t22 - t11);
  g.rows(t, k, k + 1, k, n);
  g.cols(t, k, k + 1, 0, k + 2);
  g.cols(q, k, k + 1, 0, q.rows);
  t(k, k) = t22;
  t(k + 1, k + 1) = t11;
  t(k + 1, k) = 0.0;
}

// Reorder so that the diagonal entries listed in `order` (indices into the current diagonal) come
// first, in that sequence.
inline void schur_reorder(CMat &t, CMat &q, const std::vector<int> &order) {
  const int n = t.rows;
  std::vector<int> pos(n);
  for (int i = 0; i < n; ++i) pos[i] = i;  // pos[j] = original index now at diagonal slot j
  for (size_t dst = 0; dst < order.size(); ++dst) {
    int cur = -1;
    for (int j = (int)dst; j < n; ++j)
      if (pos[j] == order[dst]) {
        cur = j;
        break;
      }
    for (int j = cur; j > (int)dst; --j) {
      schur_swap(t, q, j - 1);
      std::swap(pos[j], pos[j - 1]);
    }
  }
}

// Eigenvectors of the leading k x k block of upper-triangular t: column i of s (k x k) solves
// t s_i = t_ii s_i, normalised to unit 2-norm.
inline CMat tri_eigvecs(const CMat &t, int k) {
  CMat s(k, k);
  double tnorm = 0.0;
  for (int i = 0; i < k; ++i)
    for (int j = i; j < k; ++j) tnorm = std::max(tnorm, std::abs(t(i, j)));
  const double small = std::max(tnorm, 1e-300) * 2.220446049250313e-16;
  for (int i = 0; i < k; ++i) {
    s(i, i) = 1.0;
    for (int j = i - 1; j >= 0; --j) {
      cd acc = 0.0;
      for (int l = j + 1; l <= i; ++l) acc += t(j, l) * s(l, i);
      cd den = t(j, j) - t(i, i);
      if (std::abs(den) < small) den = small;
      s(j, i) = -acc / den;
    }
    double nrm = 0.0;
    for (int j = 0; j <= i; ++j) nrm += std::norm(s(j, i));
    nrm = std::sqrt(nrm);
    for (int j = 0; j <= i; ++j) s(j, i) /= nrm;
  }
  return s;
}

// Condition numbers of the eigenvalues of an upper-triangular T from its (unit-norm, upper-triangular) eigenvector matrix S:
// kappa_i = ||row i of S^-1||_2 (* ||column i of S||_2 = 1), i.e. 1 / |y_i^H x_i| for unit left/right eigenvectors.
inline std::vector<double> eig_condition(const CMat &s, int k) {
  // inverse of an upper-triangular matrix by back-substitution, row by row from the bottom
  CMat inv(k, k);
  for (int i = k - 1; i >= 0; --i) {
    inv(i, i) = 1.0 / s(i, i);
    for (int j = i + 1; j < k; ++j) {
      cd acc = 0.0;
      for (int l = i + 1; l <= j; ++l) acc += s(i, l) * inv(l, j);
      inv(i, j) = -acc / s(i, i);
    }
  }
  std::vector<double> kappa(k);
  for (int i = 0; i < k; ++i) {
    double r = 0.0;
    for (int j = i; j < k; ++j) r += std::norm(inv(i, j));
    kappa[i] = std::sqrt(r);
  }
  return kappa;
}

inline CMat matmul(const CMat &a, const CMat &b) {
  CMat c(a.rows, b.cols);
  for (int i = 0; i < a.rows; ++i)
    for (int l = 0; l < a.cols; ++l) {
      cd v = a(i, l);
      if (v == cd(0, 0)) continue;
      for (int j = 0; j < b.cols; ++j) c(i, j) += v * b(l, j);
    }
  return c;
}

// Real orthonormal basis (rows x k, returned as complex with zero imaginary part) of the span of the
// columns of the complex matrix `qc` (rows x k), assuming that span is closed under conjugation.
// Pivoted modified Gram-Schmidt on [Re qc, Im qc].  Returns the achieved rank.
inline int real_basis(const CMat &qc, CMat &out) {
  const int m = qc.rows, k = qc.cols;
  std::vector<std::vector<double>> cand(2 * k, std::vector<double>(m));
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < m; ++i) {
      cand[j][i] = qc(i, j).real();
      cand[k + j][i] = qc(i, j).imag();
    }
  out = CMat(m, k);
  std::vector<std::vector<double>> basis;
  int rank = 0;
  for (int r = 0; r < k; ++r) {
    int best = -1;
    double bestn = 0.0;
    for (int c = 0; c < 2 * k; ++c) {
      double nn = 0.0;
      for (int i = 0; i < m; ++i) nn += cand[c][i] * cand[c][i];
      if (nn > bestn) {
        bestn = nn;
        best = c;
      }
    }
    if (best < 0 || bestn < 1e-20) break;
    std::vector<double> v = cand[best];
    for (int pass = 0; pass < 2; ++pass)
      for (auto &b : basis) {
        double d = 0.0;
        for (int i = 0; i < m; ++i) d += b[i] * v[i];
        for (int i = 0; i < m; ++i) v[i] -= d * b[i];
      }
    double nn = 0.0;
    for (int i = 0; i < m; ++i) nn += v[i] * v[i];
    nn = std::sqrt(nn);
    for (int i = 0; i < m; ++i) v[i] /= nn;
    basis.push_back(v);
    for (int i = 0; i < m; ++i) out(i, rank) = v[i];
    ++rank;
    for (int c = 0; c < 2 * k; ++c) {
      double d = 0.0;
      for (int i = 0; i < m; ++i) d += v[i] * cand[c][i];
      for (int i = 0; i < m; ++i) cand[c][i] -= d * v[i];
    }
  }
  return rank;
}

}  // namespace b200ms
