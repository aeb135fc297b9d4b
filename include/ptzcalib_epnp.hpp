// ptzcalib_epnp.hpp — EPnP pose from 2d-3d correspondences, host-side C++17, header only.
//
// PTZRayOptimizer::SetInitTransLocalToWorld (src/core/ptzray_optimizer.cc:562-633) seeds the local->world transform of the
// georeferencing BA with cv::solvePnP(pts3d, pixels, K, dist, rvec, tvec, false, cv::SOLVEPNP_EPNP) (:572) on the first annotated
// candidate view.  OpenCV is an un-vendored dependency (4.5.3, install_deps.sh:44-73) and not part of this build, so the published
// algorithm is restated here: Lepetit, Moreno-Noguer, Fua, "EPnP: An Accurate O(n) Solution to the PnP Problem", IJCV 2009, in the
// arrangement of OpenCV's calib3d/src/epnp.cpp (control points from the PCA of the object points, 2n x 12 system, null space of
// M^T M, the three beta approximations each refined by 5 Gauss-Newton steps, Horn's absolute orientation, smallest reprojection
// error wins), preceded by what solvePnP does for this flag: cv::undistortPoints to normalised coordinates, which come back as
// float32 because the reference's pixels are cv::Point2f.  SURVEY.md §8f row 3.  A few tens of points, once per georeferencing
// run: host code, not a kernel.  Pinned against cv2.solvePnP(SOLVEPNP_EPNP) 4.13 by tests/golden/epnp_kat.npz.
#ifndef PTZCALIB_EPNP_HPP
#define PTZCALIB_EPNP_HPP

#include <algorithm>
#include <cmath>
#include <vector>

#include "../ptz-calib_b200/csrc/ptz_math.cuh"  // rodrigues_jac / rodrigues_inv (plain C++ when not compiled by nvcc)

namespace ptzcalib {
namespace epnp {

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (row-major, destroyed).  evals descending, evecs[k*n + i] =
// component i of eigenvector k (rows), like the U^T that cvSVD(..., CV_SVD_U_T) returns for a symmetric positive matrix.
inline void sym_eigen(int n, std::vector<double>& A, std::vector<double>& evals, std::vector<double>& evecs) {
  std::vector<double> V((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns p, q
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p, q
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] > A[(size_t)b * n + b]; });
  evals.resize(n);
  evecs.assign((size_t)n * n, 0.0);
  for (int k = 0; k < n; ++k) {
    evals[k] = A[(size_t)order[k] * n + order[k]];
    for (int i = 0; i < n; ++i) evecs[(size_t)k * n + i] = V[(size_t)i * n + order[k]];
  }
}

// minimum-norm least squares x = pinv(A) b for a small m x n system, m >= n (cvSolve(..., CV_SVD)): one-sided Jacobi SVD
// (Hestenes) on A itself, so the condition number is not squared as it would be through A^T A
inline void lstsq(int m, int n, const double* A, const double* b, double* x) {
  std::vector<double> W(A, A + (size_t)m * n), V((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double app = 0, aqq = 0, apq = 0;
        for (int k = 0; k < m; ++k) { const double wp = W[(size_t)k * n + p], wq = W[(size_t)k * n + q]; app += wp * wp; aqq += wq * wq; apq += wp * wq; }
        if (std::fabs(apq) <= 1e-16 * std::sqrt(app * aqq) || apq == 0.0) continue;
        rotated = true;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < m; ++k) {
          const double wp = W[(size_t)k * n + p], wq = W[(size_t)k * n + q];
          W[(size_t)k * n + p] = c * wp - sn * wq;
          W[(size_t)k * n + q] = sn * wp + c * wq;
        }
        for (int k = 0; k < n; ++k) {
          const double vp = V[(size_t)k * n + p], vq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vp - sn * vq;
          V[(size_t)k * n + q] = sn * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  // columns of W are now u_j * s_j
  double smax = 0;
  std::vector<double> s2(n, 0.0);
  for (int j = 0; j < n; ++j) {
    for (int k = 0; k < m; ++k) s2[j] += W[(size_t)k * n + j] * W[(size_t)k * n + j];
    smax = std::max(smax, s2[j]);
  }
  for (int i = 0; i < n; ++i) x[i] = 0.0;
  for (int j = 0; j < n; ++j) {
    if (s2[j] <= 1e-28 * smax || s2[j] <= 0) continue;  // singular direction left out: minimum norm
    double c = 0;
    for (int k = 0; k < m; ++k) c += W[(size_t)k * n + j] * b[k];
    c /= s2[j];
    for (int i = 0; i < n; ++i) x[i] += c * V[(size_t)i * n + j];
  }
}

// least squares by Householder QR without pivoting (the Gauss-Newton step of EPnP); false when a column is exactly zero
inline bool qr_solve(int m, int n, const double* A0, const double* b0, double* x) {
  std::vector<double> A(A0, A0 + (size_t)m * n), b(b0, b0 + m), beta(n), diag(n);
  for (int k = 0; k < n; ++k) {
    double eta = 0;
    for (int i = k; i < m; ++i) eta = std::max(eta, std::fabs(A[(size_t)i * n + k]));
    if (eta == 0.0) return false;
    double sum = 0;
    for (int i = k; i < m; ++i) { A[(size_t)i * n + k] /= eta; sum += A[(size_t)i * n + k] * A[(size_t)i * n + k]; }
    double sigma = std::sqrt(sum);
    if (A[(size_t)k * n + k] < 0) sigma = -sigma;
    A[(size_t)k * n + k] += sigma;
    beta[k] = sigma * A[(size_t)k * n + k];
    diag[k] = -eta * sigma;
    for (int j = k + 1; j < n; ++j) {
      double dotp = 0;
      for (int i = k; i < m; ++i) dotp += A[(size_t)i * n + k] * A[(size_t)i * n + j];
      const double tau = dotp / beta[k];
      for (int i = k; i < m; ++i) A[(size_t)i * n + j] -= tau * A[(size_t)i * n + k];
    }
  }
  for (int j = 0; j < n; ++j) {  // b <- Q^T b
    double dotp = 0;
    for (int i = j; i < m; ++i) dotp += A[(size_t)i * n + j] * b[i];
    const double tau = dotp / beta[j];
    for (int i = j; i < m; ++i) b[i] -= tau * A[(size_t)i * n + j];
  }
  for (int i = n - 1; i >= 0; --i) {  // R x = b
    double sum = 0;
    for (int j = i + 1; j < n; ++j) sum += A[(size_t)i * n + j] * x[j];
    x[i] = (b[i] - sum) / diag[i];
  }
  return true;
}

// A = U diag(s) V^T for a 3 x 3 matrix (row-major), s descending; rank-deficient A completed with cross products
inline void svd3(const double A[9], double U[9], double s[3], double V[9]) {
  std::vector<double> AtA(9, 0.0), ev, evec;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) AtA[3 * i + j] += A[3 * k + i] * A[3 * k + j];
  sym_eigen(3, AtA, ev, evec);
  double v[3][3], u[3][3];
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 3; ++i) v[k][i] = evec[3 * k + i];
  // right-handed V
  const double cx = v[0][1] * v[1][2] - v[0][2] * v[1][1], cy = v[0][2] * v[1][0] - v[0][0] * v[1][2], cz = v[0][0] * v[1][1] - v[0][1] * v[1][0];
  if (cx * v[2][0] + cy * v[2][1] + cz * v[2][2] < 0) for (int i = 0; i < 3; ++i) v[2][i] = -v[2][i];
  int rank = 0;
  for (int k = 0; k < 3; ++k) {
    s[k] = ev[k] > 0 ? std::sqrt(ev[k]) : 0.0;
    if (s[k] > 1e-12 * (s[0] > 0 ? s[0] : 1.0)) {
      for (int i = 0; i < 3; ++i) u[k][i] = (A[3 * i] * v[k][0] + A[3 * i + 1] * v[k][1] + A[3 * i + 2] * v[k][2]) / s[k];
      ++rank;
    }
  }
  if (rank == 2) {
    u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1]; u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2]; u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
  } else if (rank < 2) {
    for (int k = rank; k < 3; ++k) for (int i = 0; i < 3; ++i) u[k][i] = (i == k) ? 1.0 : 0.0;  // degenerate input: any completion
  }
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 3; ++i) { U[3 * i + k] = u[k][i]; V[3 * i + k] = v[k][i]; }
}

struct Solver {
  int n = 0;
  std::vector<double> pws, us, alphas, pcs;  // [3n] object points, [2n] normalised image points, [4n] barycentric, [3n] camera points
  double cws[4][3], ccs[4][3];
  // The principal axes of the object points are defined up to sign; which sign an SVD returns is an implementation detail of the
  // library (OpenCV's JacobiSVD here), and with noisy pixels each of the 8 choices gives a slightly different, equally valid pose
  // (they coincide on exact data).  Default +,+,+; tests flip them to show that cv2's answer is one of the eight.
  double axis_sign[3] = {1.0, 1.0, 1.0};

  static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
  static double dist2(const double* a, const double* b) { return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]); }

  void choose_control_points() {  // centroid + principal axes scaled by sqrt(eigenvalue / n)
    for (int j = 0; j < 3; ++j) {
      cws[0][j] = 0;
      for (int i = 0; i < n; ++i) cws[0][j] += pws[3 * i + j];
      cws[0][j] /= n;
    }
    std::vector<double> C(9, 0.0), ev, evec;
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) C[3 * a + b] += (pws[3 * i + a] - cws[0][a]) * (pws[3 * i + b] - cws[0][b]);
    sym_eigen(3, C, ev, evec);
    for (int i = 1; i < 4; ++i) {
      const double k = axis_sign[i - 1] * std::sqrt(std::max(ev[i - 1], 0.0) / n);
      for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + k * evec[3 * (i - 1) + j];
    }
  }
  void compute_barycentric_coordinates() {
    double cc[9], cci[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
    // pseudo-inverse through the SVD, as cvInvert(..., CV_SVD): coplanar object points give a rank-2 basis
    double U[9], sv[3], V[9];
    svd3(cc, U, sv, V);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double acc = 0;
        for (int k = 0; k < 3; ++k)
          if (sv[k] > 1e-12 * sv[0]) acc += V[3 * i + k] * U[3 * j + k] / sv[k];
        cci[3 * i + j] = acc;
      }
    alphas.resize(4 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      double* a = &alphas[4 * (size_t)i];
      const double d[3] = {pws[3 * i] - cws[0][0], pws[3 * i + 1] - cws[0][1], pws[3 * i + 2] - cws[0][2]};
      for (int j = 0; j < 3; ++j) a[1 + j] = cci[3 * j] * d[0] + cci[3 * j + 1] * d[1] + cci[3 * j + 2] * d[2];
      a[0] = 1.0 - a[1] - a[2] - a[3];
    }
  }
  void compute_L_6x10(const double* ut, double* L) const {
    const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
    double dv[4][6][3];
    for (int i = 0; i < 4; ++i) {
      int a = 0, b = 1;
      for (int j = 0; j < 6; ++j) {
        for (int k = 0; k < 3; ++k) dv[i][j][k] = v[i][3 * a + k] - v[i][3 * b + k];
        if (++b > 3) { ++a; b = a + 1; }
      }
    }
    for (int i = 0; i < 6; ++i) {
      double* row = L + 10 * i;
      row[0] = dot3(dv[0][i], dv[0][i]); row[1] = 2.0 * dot3(dv[0][i], dv[1][i]); row[2] = dot3(dv[1][i], dv[1][i]);
      row[3] = 2.0 * dot3(dv[0][i], dv[2][i]); row[4] = 2.0 * dot3(dv[1][i], dv[2][i]); row[5] = dot3(dv[2][i], dv[2][i]);
      row[6] = 2.0 * dot3(dv[0][i], dv[3][i]); row[7] = 2.0 * dot3(dv[1][i], dv[3][i]); row[8] = 2.0 * dot3(dv[2][i], dv[3][i]);
      row[9] = dot3(dv[3][i], dv[3][i]);
    }
  }
  void compute_rho(double* rho) const {
    rho[0] = dist2(cws[0], cws[1]); rho[1] = dist2(cws[0], cws[2]); rho[2] = dist2(cws[0], cws[3]);
    rho[3] = dist2(cws[1], cws[2]); rho[4] = dist2(cws[1], cws[3]); rho[5] = dist2(cws[2], cws[3]);
  }
  // L columns are the products [B11 B12 B22 B13 B23 B33 B14 B24 B34 B44]
  static void betas_approx_1(const double* L, const double* rho, double* betas) {  // unknowns B11 B12 B13 B14
    double A[24], b4[4];
    for (int i = 0; i < 6; ++i) { A[4 * i] = L[10 * i]; A[4 * i + 1] = L[10 * i + 1]; A[4 * i + 2] = L[10 * i + 3]; A[4 * i + 3] = L[10 * i + 6]; }
    lstsq(6, 4, A, rho, b4);
    if (b4[0] < 0) { betas[0] = std::sqrt(-b4[0]); betas[1] = -b4[1] / betas[0]; betas[2] = -b4[2] / betas[0]; betas[3] = -b4[3] / betas[0]; }
    else { betas[0] = std::sqrt(b4[0]); betas[1] = b4[1] / betas[0]; betas[2] = b4[2] / betas[0]; betas[3] = b4[3] / betas[0]; }
  }
  static void betas_approx_2(const double* L, const double* rho, double* betas) {  // unknowns B11 B12 B22
    double A[18], b3[3];
    for (int i = 0; i < 6; ++i) { A[3 * i] = L[10 * i]; A[3 * i + 1] = L[10 * i + 1]; A[3 * i + 2] = L[10 * i + 2]; }
    lstsq(6, 3, A, rho, b3);
    if (b3[0] < 0) { betas[0] = std::sqrt(-b3[0]); betas[1] = b3[2] < 0 ? std::sqrt(-b3[2]) : 0.0; }
    else { betas[0] = std::sqrt(b3[0]); betas[1] = b3[2] > 0 ? std::sqrt(b3[2]) : 0.0; }
    if (b3[1] < 0) betas[0] = -betas[0];
    betas[2] = 0.0; betas[3] = 0.0;
  }
  static void betas_approx_3(const double* L, const double* rho, double* betas) {  // unknowns B11 B12 B22 B13 B23
    double A[30], b5[5];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 5; ++j) A[5 * i + j] = L[10 * i + j];
    lstsq(6, 5, A, rho, b5);
    if (b5[0] < 0) { betas[0] = std::sqrt(-b5[0]); betas[1] = b5[2] < 0 ? std::sqrt(-b5[2]) : 0.0; }
    else { betas[0] = std::sqrt(b5[0]); betas[1] = b5[2] > 0 ? std::sqrt(b5[2]) : 0.0; }
    if (b5[1] < 0) betas[0] = -betas[0];
    betas[2] = b5[3] / betas[0]; betas[3] = 0.0;
  }
  static void gauss_newton(const double* L, const double* rho, double* be) {
    for (int it = 0; it < 5; ++it) {
      double A[24], b[6], x[4];
      for (int i = 0; i < 6; ++i) {
        const double* l = L + 10 * i;
        A[4 * i] = 2 * l[0] * be[0] + l[1] * be[1] + l[3] * be[2] + l[6] * be[3];
        A[4 * i + 1] = l[1] * be[0] + 2 * l[2] * be[1] + l[4] * be[2] + l[7] * be[3];
        A[4 * i + 2] = l[3] * be[0] + l[4] * be[1] + 2 * l[5] * be[2] + l[8] * be[3];
        A[4 * i + 3] = l[6] * be[0] + l[7] * be[1] + l[8] * be[2] + 2 * l[9] * be[3];
        b[i] = rho[i] - (l[0] * be[0] * be[0] + l[1] * be[0] * be[1] + l[2] * be[1] * be[1] + l[3] * be[0] * be[2] + l[4] * be[1] * be[2] +
                         l[5] * be[2] * be[2] + l[6] * be[0] * be[3] + l[7] * be[1] * be[3] + l[8] * be[2] * be[3] + l[9] * be[3] * be[3]);
      }
      if (!qr_solve(6, 4, A, b, x)) return;
      for (int i = 0; i < 4; ++i) be[i] += x[i];
    }
  }
  void estimate_R_and_t(double R[9], double t[3]) const {  // Horn: R = U V^T of sum (pc - pc0)(pw - pw0)^T
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) { pc0[j] += pcs[3 * i + j]; pw0[j] += pws[3 * i + j]; }
    for (int j = 0; j < 3; ++j) { pc0[j] /= n; pw0[j] /= n; }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, U[9], s[3], V[9];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) abt[3 * j + k] += (pcs[3 * i + j] - pc0[j]) * (pws[3 * i + k] - pw0[k]);
    svd3(abt, U, s, V);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[3 * i + j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1] + U[3 * i + 2] * V[3 * j + 2];
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
    for (int j = 0; j < 3; ++j) t[j] = pc0[j] - (R[3 * j] * pw0[0] + R[3 * j + 1] * pw0[1] + R[3 * j + 2] * pw0[2]);
  }
  double reprojection_error(const double R[9], const double t[3]) const {
    double sum = 0;
    for (int i = 0; i < n; ++i) {
      const double* p = &pws[3 * (size_t)i];
      const double X = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0], Y = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1];
      const double iz = 1.0 / (R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2]);
      const double du = us[2 * i] - X * iz, dv = us[2 * i + 1] - Y * iz;
      sum += std::sqrt(du * du + dv * dv);
    }
    return sum / n;
  }
  double compute_R_and_t(const double* ut, const double* betas, double R[9], double t[3]) {
    for (int i = 0; i < 4; ++i) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double* v = ut + 12 * (11 - i);
      for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 3; ++k) ccs[j][k] += betas[i] * v[3 * j + k];
    }
    pcs.resize(3 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      const double* a = &alphas[4 * (size_t)i];
      for (int j = 0; j < 3; ++j) pcs[3 * i + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
    }
    if (pcs[2] < 0.0) {  // the first point must lie in front of the camera
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) ccs[i][j] = -ccs[i][j];
      for (size_t i = 0; i < pcs.size(); ++i) pcs[i] = -pcs[i];
    }
    estimate_R_and_t(R, t);
    return reprojection_error(R, t);
  }
  // object points [3n], NORMALISED image points [2n] (K = I).  R row-major (camera <- world), t.
  void compute_pose(int num, const double* object_pts, const double* image_pts, double R[9], double t[3]) {
    n = num;
    pws.assign(object_pts, object_pts + 3 * (size_t)n);
    us.assign(image_pts, image_pts + 2 * (size_t)n);
    choose_control_points();
    compute_barycentric_coordinates();
    std::vector<double> MtM(144, 0.0), ev, ut;
    for (int i = 0; i < n; ++i) {
      const double* a = &alphas[4 * (size_t)i];
      double m1[12], m2[12];
      for (int j = 0; j < 4; ++j) {
        m1[3 * j] = a[j]; m1[3 * j + 1] = 0.0; m1[3 * j + 2] = a[j] * (0.0 - us[2 * i]);
        m2[3 * j] = 0.0; m2[3 * j + 1] = a[j]; m2[3 * j + 2] = a[j] * (0.0 - us[2 * i + 1]);
      }
      for (int p = 0; p < 12; ++p)
        for (int q = 0; q < 12; ++q) MtM[12 * p + q] += m1[p] * m1[q] + m2[p] * m2[q];
    }
    sym_eigen(12, MtM, ev, ut);
    double L[60], rho[6], betas[4][4], Rs[4][9], ts[4][3], err[4];
    compute_L_6x10(ut.data(), L);
    compute_rho(rho);
    betas_approx_1(L, rho, betas[1]); gauss_newton(L, rho, betas[1]); err[1] = compute_R_and_t(ut.data(), betas[1], Rs[1], ts[1]);
    betas_approx_2(L, rho, betas[2]); gauss_newton(L, rho, betas[2]); err[2] = compute_R_and_t(ut.data(), betas[2], Rs[2], ts[2]);
    betas_approx_3(L, rho, betas[3]); gauss_newton(L, rho, betas[3]); err[3] = compute_R_and_t(ut.data(), betas[3], Rs[3], ts[3]);
    int N = 1;
    if (err[2] < err[1]) N = 2;
    if (err[3] < err[N]) N = 3;
    for (int i = 0; i < 9; ++i) R[i] = Rs[N][i];
    for (int i = 0; i < 3; ++i) t[i] = ts[N][i];
  }
};

// cv::undistortPoints(pixels, K, dist) without P: normalised coordinates after 5 fixed-point iterations (OpenCV coefficient order
// k1,k2,p1,p2,k3), returned as float32 because the input is cv::Point2f
inline void undistort_normalized_f32(float u, float v, double fx, double fy, double cx, double cy, const double d[5], float out[2]) {
  double x = ((double)u - cx) / fx, y = ((double)v - cy) / fy;
  const double x0 = x, y0 = y;
  const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
  if (k1 != 0 || k2 != 0 || p1 != 0 || p2 != 0 || k3 != 0) {
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
      if (icdist < 0) { x = x0; y = y0; break; }
      const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x), dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
  }
  out[0] = (float)x;
  out[1] = (float)y;
}

// cv::solvePnP(object (double), pixels (float), K, dist, rvec, tvec, false, SOLVEPNP_EPNP): R (row-major) and t.  false for n < 4.
inline bool solve_pnp_epnp(int n, const double* object_pts, const float* pixels, const double K[9], const double dist[5], double R[9], double t[3],
                           const double* axis_sign = nullptr) {
  if (n < 4) return false;
  std::vector<double> us(2 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    float o[2];
    undistort_normalized_f32(pixels[2 * i], pixels[2 * i + 1], K[0], K[4], K[2], K[5], dist, o);
    us[2 * i] = o[0]; us[2 * i + 1] = o[1];
  }
  Solver s;
  if (axis_sign) for (int i = 0; i < 3; ++i) s.axis_sign[i] = axis_sign[i];
  s.compute_pose(n, object_pts, us.data(), R, t);
  for (int i = 0; i < 9; ++i) if (!std::isfinite(R[i])) return false;
  return std::isfinite(t[0]) && std::isfinite(t[1]) && std::isfinite(t[2]);
}

// One view's attempt of PTZRayOptimizer::SetInitTransLocalToWorld (ptzray_optimizer.cc:566-616): EPnP pose of the annotated view, the
// gates of :582-604 (first point in front, det R >= 0, reprojection RMS of the float32 points <= 300 px without distortion) and
// T_l_w = T_i_l^-1 T_i_w (:606-616) as [rvec | t].  Ri, ti: the view's rotation (row-major) and translation in the local frame.
inline bool init_tlw_from_view(int n, const double* obj, const float* pix, const double K[9], const double dist[5], const double Ri[9], const double ti[3],
                               double tlw[6]) {
  double R[9], tv[3], rv[3];
  if (n <= 0 || !solve_pnp_epnp(n, obj, pix, K, dist, R, tv)) return false;
  // cv::Rodrigues(rvec, R) of the rvec solvePnP returns: the round trip re-orthonormalises R
  ptz::rodrigues_inv(R, rv);
  ptz::rodrigues_jac(rv, R, nullptr);
  const double z0 = R[6] * obj[0] + R[7] * obj[1] + R[8] * obj[2] + tv[2];
  const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
  if (z0 < 0 || det < 0.0) return false;  // .cc:582-587
  double sq = 0;  // .cc:589-604
  for (int j = 0; j < n; ++j) {
    const double X = (float)obj[3 * j], Y = (float)obj[3 * j + 1], Z = (float)obj[3 * j + 2];
    const double xc = R[0] * X + R[1] * Y + R[2] * Z + tv[0], yc = R[3] * X + R[4] * Y + R[5] * Z + tv[1], zc = R[6] * X + R[7] * Y + R[8] * Z + tv[2];
    const float pu = (float)(K[0] * (xc / zc) + K[2]), pv = (float)(K[4] * (yc / zc) + K[5]);
    sq += (double)(pu - pix[2 * j]) * (pu - pix[2 * j]) + (double)(pv - pix[2 * j + 1]) * (pv - pix[2 * j + 1]);
  }
  if (std::sqrt(sq / n) > 300) return false;
  double Rlw[9];  // R_lw = R_i^T R, t_lw = R_i^T (t - t_i)
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rlw[3 * r + c] = Ri[r] * R[c] + Ri[3 + r] * R[3 + c] + Ri[6 + r] * R[6 + c];
    tlw[3 + r] = Ri[r] * (tv[0] - ti[0]) + Ri[3 + r] * (tv[1] - ti[1]) + Ri[6 + r] * (tv[2] - ti[2]);
  }
  ptz::rodrigues_inv(Rlw, tlw);
  return true;
}

}  // namespace epnp
}  // namespace ptzcalib
#endif  // PTZCALIB_EPNP_HPP
