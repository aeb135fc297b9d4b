// cg_ritz.hpp — host side of the CG deflation basis (k_cg<.., KD = kDeflK>, ba_kernels.cuh).
//
// The first linear solve of a run is undeflated and records, per iteration j, the residual r_j (device) and the scalars
// alpha_j, beta_j, gamma_j = |r_j|^2 (read back with the iteration's scalar block).  CG is the Lanczos process in disguise:
// in the basis r_j / |r_j| the operator is the tridiagonal
//     T[j][j] = 1/alpha_j + beta_j/alpha_{j-1},   T[j][j+1] = -sqrt(beta_{j+1}) / alpha_j        (beta_0 = 0)
// whose lowest eigenpairs (Ritz pairs) approximate the smallest eigenvalues of S~ -- the modes the plain iteration spends its
// plateau on.  This file finds the lowest K of them (Sturm bisection + inverse iteration; O(K m) work, m <= 320) and returns
// the coefficient matrix Y with 1/|r_j| folded in, so that the device forms  W~ = sum_j Y[j][.] r_j  in one pass.
// Lanczos "ghosts" (copies of an already converged Ritz value that reappear once orthogonality is lost) are skipped.
// Plain host C++ (no CUDA): unit-tested on the CPU by tests/test_host_ritz.py against numpy.
#pragma once
#include <math.h>

#include <algorithm>
#include <vector>

namespace ptz {

struct TriDiag {
  std::vector<double> a, b;  // diagonal [m], off-diagonal [m-1]
  int m() const { return (int)a.size(); }
};

// number of eigenvalues of T below x
inline int sturm_count(const TriDiag& T, double x) {
  const int m = T.m();
  int cnt = 0;
  double q = T.a[0] - x;
  if (q < 0) ++cnt;
  for (int j = 1; j < m; ++j) {
    if (fabs(q) < 1e-300) q = q < 0 ? -1e-300 : 1e-300;
    q = T.a[j] - x - T.b[j - 1] * T.b[j - 1] / q;
    if (q < 0) ++cnt;
  }
  return cnt;
}

// i-th eigenvalue (ascending, 0-based) by bisection inside the Gershgorin interval
inline double tri_eigenvalue(const TriDiag& T, int i) {
  const int m = T.m();
  double lo = T.a[0], hi = T.a[0];
  for (int j = 0; j < m; ++j) {
    const double r = (j > 0 ? fabs(T.b[j - 1]) : 0.0) + (j + 1 < m ? fabs(T.b[j]) : 0.0);
    lo = std::min(lo, T.a[j] - r);
    hi = std::max(hi, T.a[j] + r);
  }
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (sturm_count(T, mid) > i) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

// solves (T - shift I) z = y in place: Gaussian elimination with partial pivoting on the tridiagonal
inline void tri_shifted_solve(const TriDiag& T, double shift, std::vector<double>& y) {
  const int m = T.m();
  std::vector<double> d(m), u1(m, 0.0), u2(m, 0.0);
  d[0] = T.a[0] - shift;
  if (m > 1) u1[0] = T.b[0];
  for (int j = 0; j + 1 < m; ++j) {
    double n0 = T.b[j], n1 = T.a[j + 1] - shift, n2 = (j + 2 < m) ? T.b[j + 1] : 0.0;  // row j+1 at columns j, j+1, j+2
    if (fabs(n0) > fabs(d[j])) {
      std::swap(d[j], n0); std::swap(u1[j], n1); std::swap(u2[j], n2);
      std::swap(y[j], y[j + 1]);
    }
    if (d[j] == 0.0) d[j] = 1e-300;
    const double f = n0 / d[j];
    d[j + 1] = n1 - f * u1[j];
    u1[j + 1] = n2 - f * u2[j];
    y[j + 1] -= f * y[j];
  }
  if (d[m - 1] == 0.0) d[m - 1] = 1e-300;
  for (int j = m - 1; j >= 0; --j) {
    double s = y[j];
    if (j + 1 < m) s -= u1[j] * y[j + 1];
    if (j + 2 < m) s -= u2[j] * y[j + 2];
    y[j] = s / d[j];
  }
}

// abg = [alpha_j, beta_j, gamma_j] of iterations 0..m-1 of an undeflated k_cg run.  Writes Y [m][ld] (column c = coefficients of
// the c-th kept Ritz vector over the residuals, 1/|r_j| included; columns >= the return value are zero) and the kept Ritz values.
// Returns how many vectors were kept (<= kmax).
inline int lowest_ritz_vectors(const double* abg, int m, int kmax, int ld, std::vector<double>& Y, std::vector<double>* values = nullptr) {
  Y.assign((size_t)std::max(m, 0) * ld, 0.0);
  if (values) values->clear();
  if (m < 2 || kmax <= 0) return 0;
  TriDiag T;
  T.a.resize(m);
  T.b.resize(m - 1);
  for (int j = 0; j < m; ++j) {
    const double al = abg[3 * j], be = abg[3 * j + 1];
    if (!(al > 0) || !(abg[3 * j + 2] > 0) || !isfinite(al)) return 0;
    T.a[j] = 1.0 / al + (j > 0 ? be / abg[3 * (j - 1)] : 0.0);
    if (j + 1 < m) {
      const double bn = abg[3 * (j + 1) + 1];
      if (!(bn > 0)) return 0;
      T.b[j] = -sqrt(bn) / al;
    }
  }
  double tnorm = 0;
  for (int j = 0; j < m; ++j) tnorm = std::max(tnorm, fabs(T.a[j]) + (j > 0 ? fabs(T.b[j - 1]) : 0.0) + (j + 1 < m ? fabs(T.b[j]) : 0.0));
  std::vector<std::vector<double>> kept;
  std::vector<double> kept_val;
  const int scan = std::min(m, 3 * kmax);
  for (int i = 0; i < scan && (int)kept.size() < kmax; ++i) {
    const double lam = tri_eigenvalue(T, i);
    if (!kept_val.empty() && fabs(lam - kept_val.back()) <= 1e-6 * std::max(fabs(lam), 1e-300)) continue;  // ghost of the previous one
    std::vector<double> z(m);
    for (int j = 0; j < m; ++j) z[j] = 1.0 + 0.37 * ((j * 7919) % 13);  // fixed, generic start vector
    const double shift = lam + 1e-14 * tnorm + 1e-10 * fabs(lam);
    bool ok = true;
    for (int rep = 0; rep < 3 && ok; ++rep) {
      tri_shifted_solve(T, shift, z);
      for (size_t q = 0; q < kept.size(); ++q) {  // keep clusters apart
        double dot = 0;
        for (int j = 0; j < m; ++j) dot += z[j] * kept[q][j];
        for (int j = 0; j < m; ++j) z[j] -= dot * kept[q][j];
      }
      double nn = 0;
      for (int j = 0; j < m; ++j) nn += z[j] * z[j];
      if (!(nn > 0) || !isfinite(nn)) { ok = false; break; }
      nn = 1.0 / sqrt(nn);
      for (int j = 0; j < m; ++j) z[j] *= nn;
    }
    if (!ok) continue;
    kept.push_back(z);
    kept_val.push_back(lam);
  }
  for (size_t c = 0; c < kept.size(); ++c)
    for (int j = 0; j < m; ++j) Y[(size_t)j * ld + c] = kept[c][j] / sqrt(abg[3 * j + 2]);
  if (values) *values = kept_val;
  return (int)kept.size();
}

}  // namespace ptz
