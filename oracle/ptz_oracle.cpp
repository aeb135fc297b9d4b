/*
 * ptz_oracle.cpp — CPU restatement of PTZ-Calib's Ceres-backed NLS path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the timed CPU baseline.  The product (ptz-calib_b200/csrc) never links,
 * loads or calls it and has no CPU fallback.
 *
 * What is restated, and from where (paths under /root/reference; [ext] = un-vendored third-party code,
 * restated from its published algorithm: Ceres Solver 1.14.0 and OpenCV 4.5.3, install_deps.sh:44-73,114-129):
 *   - Camera::FromVector / ToVector, 15-vector layout                      src/core/types.cc:32-73
 *   - cv::Rodrigues (vector->matrix and matrix->vector)                    [ext] calib3d cvRodrigues2
 *   - cv::Mat::inv() for 3x3 (cofactor formula)                            [ext] core/lapack.cpp cv::invert
 *   - cv::undistortPoints (5 fixed-point iterations, float32 output)       [ext] calib3d/undistort.dispatch.cpp
 *   - cv::projectPoints (k1,k2,p1,p2,k3 ordering)                          [ext] calib3d cvProjectPoints2
 *   - the six PTZ-ray / 2d-3d functors                                     src/core/ptzray_optimizer.cc:20-401
 *   - the six KRT functors                                                 src/core/krt_optimizer.cc:22-248
 *   - problem wiring: weights, blocks, fixed coordinates                   ptzray_optimizer.cc:799-958, krt_optimizer.cc:265-348
 *   - initial rays (Pix2Ray)                                               ptzray_optimizer.cc:768-797
 *   - NumericDiffCostFunction<CENTRAL>                                     [ext] ceres/internal/numeric_diff.h
 *   - ScaledLoss corrector, SubsetParameterization                         [ext] ceres/loss_function.cc, corrector.cc
 *   - TrustRegionMinimizer + LevenbergMarquardtStrategy                    [ext] ceres/trust_region_minimizer.cc,
 *                                                                                levenberg_marquardt_strategy.cc
 *   - SPARSE_SCHUR (exact Schur elimination of rays + Cholesky)            [ext] ceres/schur_eliminator_impl.h
 *   - DENSE_QR                                                             [ext] ceres/dense_qr_solver.cc
 *   - acceptance logic and reported errors                                 ptzray_optimizer.cc:482,960-1072; krt_optimizer.cc:504-567
 *
 * PINNING STATUS: the reference ships no tests, fixtures or golden vectors, and neither Ceres
 * nor the reference can be built offline.  The functor arithmetic (Rodrigues, projection, distortion,
 * undistort) IS pinned against OpenCV-python 4.13 and 50-digit mpmath derivatives by tests/golden/ (generated
 * by tests/golden/make_golden.py); the minimiser is cross-checked against scipy.optimize.least_squares on
 * gauge-fixed problems.  The LM trajectory (iteration table) is "parity unpinned" against Ceres itself.
 *
 * An exact Jacobian is obtained by forward-mode dual numbers over the same templated functors; it is
 * deliberately a different derivation from the hand-written closed forms in the CUDA kernels.
 */
#include "../include/ptzcalib_b200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#include <cstdlib>
#endif

namespace {

// =====================================================================================================
// forward-mode dual numbers
// =====================================================================================================
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double x, int k) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1; }
};
#define JOP template <int N> inline
JOP Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
JOP Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
JOP Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
JOP Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
JOP Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; double iy = 1.0 / y.a; r.a = x.a * iy; for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * iy; return r; }
JOP Jet<N> operator+(const Jet<N>& x, double y) { Jet<N> r = x; r.a += y; return r; }
JOP Jet<N> operator+(double y, const Jet<N>& x) { Jet<N> r = x; r.a += y; return r; }
JOP Jet<N> operator-(const Jet<N>& x, double y) { Jet<N> r = x; r.a -= y; return r; }
JOP Jet<N> operator-(double y, const Jet<N>& x) { Jet<N> r = -x; r.a += y; return r; }
JOP Jet<N> operator*(const Jet<N>& x, double y) { Jet<N> r; r.a = x.a * y; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * y; return r; }
JOP Jet<N> operator*(double y, const Jet<N>& x) { return x * y; }
JOP Jet<N> operator/(const Jet<N>& x, double y) { return x * (1.0 / y); }
JOP Jet<N> operator/(double y, const Jet<N>& x) { return Jet<N>(y) / x; }
JOP Jet<N>& operator+=(Jet<N>& x, const Jet<N>& y) { x = x + y; return x; }
JOP Jet<N>& operator/=(Jet<N>& x, const Jet<N>& y) { x = x / y; return x; }
JOP Jet<N> Sqrt(const Jet<N>& x) { Jet<N> r; r.a = std::sqrt(x.a); double d = 0.5 / r.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
JOP Jet<N> Sin(const Jet<N>& x) { Jet<N> r; r.a = std::sin(x.a); double d = std::cos(x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
JOP Jet<N> Cos(const Jet<N>& x) { Jet<N> r; r.a = std::cos(x.a); double d = -std::sin(x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
JOP double Val(const Jet<N>& x) { return x.a; }
inline double Sqrt(double x) { return std::sqrt(x); }
inline double Sin(double x) { return std::sin(x); }
inline double Cos(double x) { return std::cos(x); }
inline double Val(double x) { return x; }

// =====================================================================================================
// OpenCV pieces
// =====================================================================================================
// cv::Rodrigues vector -> matrix [ext]: theta < DBL_EPSILON -> I, else c*I + (1-c)*k k^T + s*[k]x
void rodrigues(const double r[3], double R[9]) {
  double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < DBL_EPSILON) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
  double x = r[0] * it, y = r[1] * it, z = r[2] * it;
  R[0] = c + c1 * (x * x); R[1] = c1 * (x * y) + s * (-z); R[2] = c1 * (x * z) + s * y;
  R[3] = c1 * (x * y) + s * z; R[4] = c + c1 * (y * y); R[5] = c1 * (y * z) + s * (-x);
  R[6] = c1 * (x * z) + s * (-y); R[7] = c1 * (y * z) + s * x; R[8] = c + c1 * (z * z);
}
// dual-number version: same function, written as I + a[r]x + b[r]x^2 with the theta^2 series near 0 so that
// the derivative at rvec = 0 (start point of every reloc query) is the true one, not that of the I-branch.
template <int N>
void rodrigues(const Jet<N> r[3], Jet<N> R[9]) {
  typedef Jet<N> T;
  T t2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  T a, b;
  if (t2.a < 1e-6) {
    a = 1.0 - t2 * (1.0 / 6.0) * (1.0 - t2 * (1.0 / 20.0) * (1.0 - t2 * (1.0 / 42.0)));
    b = 0.5 - t2 * (1.0 / 24.0) * (1.0 - t2 * (1.0 / 30.0) * (1.0 - t2 * (1.0 / 56.0)));
  } else {
    T th = Sqrt(t2);
    a = Sin(th) / th;
    b = (1.0 - Cos(th)) / t2;
  }
  T x = r[0], y = r[1], z = r[2];
  R[0] = 1.0 - b * (y * y + z * z); R[1] = b * (x * y) - a * z; R[2] = b * (x * z) + a * y;
  R[3] = b * (x * y) + a * z; R[4] = 1.0 - b * (x * x + z * z); R[5] = b * (y * z) - a * x;
  R[6] = b * (x * z) - a * y; R[7] = b * (y * z) + a * x; R[8] = 1.0 - b * (x * x + y * y);
}

// cv::invert for 3x3 [ext]: cofactors times 1/det (returns zeros when det == 0)
void inv3(const double S[9], double D[9]) {
  double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (d == 0) { for (int i = 0; i < 9; ++i) D[i] = 0; return; }
  d = 1.0 / d;
  double t[9];
  t[0] = (S[4] * S[8] - S[5] * S[7]) * d; t[1] = (S[2] * S[7] - S[1] * S[8]) * d; t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  t[3] = (S[5] * S[6] - S[3] * S[8]) * d; t[4] = (S[0] * S[8] - S[2] * S[6]) * d; t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  t[6] = (S[3] * S[7] - S[4] * S[6]) * d; t[7] = (S[1] * S[6] - S[0] * S[7]) * d; t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  for (int i = 0; i < 9; ++i) D[i] = t[i];
}
void mul33(const double A[9], const double B[9], double C[9]) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  for (int i = 0; i < 9; ++i) C[i] = t[i];
}
void mul31(const double A[9], const double b[3], double c[3]) {
  double t[3];
  for (int i = 0; i < 3; ++i) t[i] = A[3 * i] * b[0] + A[3 * i + 1] * b[1] + A[3 * i + 2] * b[2];
  c[0] = t[0]; c[1] = t[1]; c[2] = t[2];
}

// cv::Rodrigues matrix -> vector [ext].  OpenCV first replaces R by U*Vt of its SVD (the nearest rotation);
// here that polar factor is obtained by Newton iteration R <- (R + R^-T)/2, which converges to the same matrix.
void rodrigues_inv(const double Rin[9], double r[3]) {
  double R[9];
  for (int i = 0; i < 9; ++i) R[i] = Rin[i];
  for (int it = 0; it < 20; ++it) {
    double Ri[9];
    inv3(R, Ri);
    double diff = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double nv = 0.5 * (R[3 * i + j] + Ri[3 * j + i]);
        diff = std::max(diff, std::fabs(nv - R[3 * i + j]));
        R[3 * i + j] = nv;
      }
    if (diff < 1e-16) break;
  }
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = std::acos(c);
  if (s < 1e-5) {
    if (c > 0) { rx = ry = rz = 0; }
    else {
      double t;
      t = (R[0] + 1) * 0.5; rx = std::sqrt(std::max(t, 0.));
      t = (R[4] + 1) * 0.5; ry = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
      t = (R[8] + 1) * 0.5; rz = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
      if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
      theta /= std::sqrt(rx * rx + ry * ry + rz * rz);
      rx *= theta; ry *= theta; rz *= theta;
    }
  } else {
    double vth = 1 / (2 * s);
    vth *= theta;
    rx *= vth; ry *= vth; rz *= vth;
  }
  r[0] = rx; r[1] = ry; r[2] = rz;
}

// cv::undistortPoints(src, dst, K, dist, noArray(), P=K) for one float point [ext]: 5 fixed-point iterations
// (TermCriteria(MAX_ITER,5,0.01)), distortion read in OpenCV order (k1,k2,p1,p2,k3), result cast to float32.
void undistort_point_f32(const float uv[2], const double K4[4], const double d[5], float out[2]) {
  double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
  double ifx = 1. / fx, ify = 1. / fy;
  double x = ((double)uv[0] - cx) * ifx, y = ((double)uv[1] - cy) * ify;
  double x0 = x, y0 = y;
  double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
  bool any = (k1 != 0 || k2 != 0 || p1 != 0 || p2 != 0 || k3 != 0);
  if (any) {
    for (int j = 0; j < 5; ++j) {
      double r2 = x * x + y * y;
      double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
      if (icdist < 0) { x = x0; y = y0; break; }
      double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
      double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
  }
  // P = K: xx = fx*x + cx
  out[0] = (float)(x * fx + cx);
  out[1] = (float)(y * fy + cy);
}

// =====================================================================================================
// functors (templated on the scalar so the same text gives values and exact derivatives)
// =====================================================================================================
template <class T> struct RodSel { static void run(const T r[3], T R[9]) { rodrigues(r, R); } };

// Brown model as written in the hand-coded factors: (k1,k2,k3,p1,p2) = v[10..14] (ptzray_optimizer.cc:107-120)
template <class T>
inline void brown_ref(const T& x, const T& y, const T& k1, const T& k2, const T& k3, const T& p1, const T& p2, T& xd, T& yd) {
  T r2 = x * x + y * y;
  T r4 = r2 * r2;
  T r6 = r2 * r2 * r2;
  T xy = x * y, x2 = x * x, y2 = y * y;
  T radial = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
  xd = x * radial + 2.0 * p1 * xy + p2 * (r2 + 2.0 * x2);
  yd = y * radial + 2.0 * p2 * xy + p1 * (r2 + 2.0 * y2);
}

// PTZRayFactor / PTZRayDistFactor / PTZRayFxfyDistFactor / PTZRayDistDispFactor (ptzray_optimizer.cc:20-264)
template <class T>
void ba_ray_factor(int type, const T* intr, const T* ext, const T* ray, const T* disp, double u, double v, T* res) {
  T R[9];
  RodSel<T>::run(ext, R);
  T n[3] = {ray[0], ray[1], ray[2]};
  if (type != PTZ_BA_PTZRAY_DIST) {  // :46, :164, :227 normalise; :91 does not
    T nn = Sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] = n[0] / nn; n[1] = n[1] / nn; n[2] = n[2] / nn;
  }
  T fx = intr[0], fy = (type == PTZ_BA_PTZRAY_FXFY_DIST) ? intr[1] : intr[0], cx = intr[2], cy = intr[3];
  if (type == PTZ_BA_PTZRAY) {
    // :49-50  p = K*R*n (MatExpr evaluates (K*R)*n), then p /= p.z
    T KR[9];
    for (int j = 0; j < 3; ++j) {
      KR[j] = fx * R[j] + cx * R[6 + j];
      KR[3 + j] = fy * R[3 + j] + cy * R[6 + j];
      KR[6 + j] = R[6 + j];
    }
    T p0 = KR[0] * n[0] + KR[1] * n[1] + KR[2] * n[2];
    T p1 = KR[3] * n[0] + KR[4] * n[1] + KR[5] * n[2];
    T p2 = KR[6] * n[0] + KR[7] * n[1] + KR[8] * n[2];
    res[0] = u - p0 / p2;
    res[1] = v - p1 / p2;
    return;
  }
  T X = R[0] * n[0] + R[1] * n[1] + R[2] * n[2];
  T Y = R[3] * n[0] + R[4] * n[1] + R[5] * n[2];
  T Z = R[6] * n[0] + R[7] * n[1] + R[8] * n[2];
  if (type == PTZ_BA_PTZRAY_DIST && Val(Z) < 0) {  // :97-102 constant penalty
    res[0] = T(1000000.0);
    res[1] = T(1000000.0);
    return;
  }
  if (type == PTZ_BA_PTZRAY_DIST_DISP) Z = Z + (disp[0] + disp[1] * fx + disp[2] * fx * fx);  // :231-232
  T x = X / Z, y = Y / Z;
  T xd, yd;
  brown_ref(x, y, intr[4], intr[5], intr[6], intr[7], intr[8], xd, yd);
  res[0] = u - (fx * xd + cx);
  res[1] = v - (fy * yd + cy);
}

// Reproj2d3dFactor / Reproj2d3dDispFactor (ptzray_optimizer.cc:268-401): fx,fy both live; no camera t.
template <class T>
void ba_pt_factor(int type, const T* intr, const T* ext, const T* tlw, const T* disp, double u, double v, const double Xw[3], T* res) {
  T R[9], Rl[9];
  RodSel<T>::run(ext, R);
  RodSel<T>::run(tlw, Rl);
  T Xl[3];
  for (int i = 0; i < 3; ++i) Xl[i] = Rl[3 * i] * Xw[0] + Rl[3 * i + 1] * Xw[1] + Rl[3 * i + 2] * Xw[2] + tlw[3 + i];
  T X = R[0] * Xl[0] + R[1] * Xl[1] + R[2] * Xl[2];
  T Y = R[3] * Xl[0] + R[4] * Xl[1] + R[5] * Xl[2];
  T Z = R[6] * Xl[0] + R[7] * Xl[1] + R[8] * Xl[2];
  T fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  if (type == PTZ_BA_PTZRAY_DIST_DISP) Z = Z + (disp[0] + disp[1] * fx + disp[2] * fx * fx);  // :368-369
  T x = X / Z, y = Y / Z;
  T xd, yd;
  brown_ref(x, y, intr[4], intr[5], intr[6], intr[7], intr[8], xd, yd);
  res[0] = u - (fx * xd + cx);
  res[1] = v - (fy * yd + cy);
}

// per-match constants of the KRT 2d-2d factors: everything that does not depend on the optimised camera
struct KrtPre {
  double ray1[3];  // normalised or not, see below
  int masked;      // Factor2d2dDist border mask (krt_optimizer.cc:95-101)
};
// cam1 = reference camera in its own frame: R = I, t = 0 (krt_optimizer.cc:272-276)
void krt_precompute(int type, const double refK4[4], const double refd[5], const float uv1[2], KrtPre* pre) {
  double K[9] = {refK4[0], 0, refK4[2], 0, refK4[1], refK4[3], 0, 0, 1}, Ki[9];
  inv3(K, Ki);
  double pt[3];
  pre->masked = 0;
  if (type == PTZ_KRT_FDIST || type == PTZ_KRT_FXFYDIST) {
    float und[2];
    undistort_point_f32(uv1, refK4, refd, und);  // :89-92 float32 result
    pt[0] = und[0]; pt[1] = und[1]; pt[2] = 1;
    double w1 = refK4[2] * 2, h1 = refK4[3] * 2;
    if (und[0] < 0 || und[0] >= w1 || und[1] < 0 || und[1] >= h1) pre->masked = 1;
  } else {
    pt[0] = uv1[0]; pt[1] = uv1[1]; pt[2] = 1;
  }
  // ray1 = R1^-1 * K1^-1 * pt1 with R1 = I
  mul31(Ki, pt, pre->ray1);
  if (type != PTZ_KRT_FXFY) {  // :34, :113, :162 normalise; Factor2d2dFxfy (:61) does not
    double nn = std::sqrt(pre->ray1[0] * pre->ray1[0] + pre->ray1[1] * pre->ray1[1] + pre->ray1[2] * pre->ray1[2]);
    pre->ray1[0] /= nn; pre->ray1[1] /= nn; pre->ray1[2] /= nn;
  }
}

// Factor2d2d / Factor2d2dFxfy / Factor2d2dDist / Factor2d2dFxfyDist (krt_optimizer.cc:22-197)
template <class T>
void krt_2d2d_factor(int type, const T* cam, const KrtPre& pre, double u2, double v2, T* res) {
  if (pre.masked) { res[0] = T(0.0); res[1] = T(0.0); return; }
  T R[9];
  RodSel<T>::run(cam + 4, R);
  T fx = cam[0], fy = (type == PTZ_KRT_FXFY || type == PTZ_KRT_FXFYDIST) ? cam[1] : cam[0], cx = cam[2], cy = cam[3];
  const double* n = pre.ray1;
  if (type == PTZ_KRT_F || type == PTZ_KRT_FXFY) {
    T KR[9];
    for (int j = 0; j < 3; ++j) {
      KR[j] = fx * R[j] + cx * R[6 + j];
      KR[3 + j] = fy * R[3 + j] + cy * R[6 + j];
      KR[6 + j] = R[6 + j];
    }
    T p0 = KR[0] * n[0] + KR[1] * n[1] + KR[2] * n[2];
    T p1 = KR[3] * n[0] + KR[4] * n[1] + KR[5] * n[2];
    T p2 = KR[6] * n[0] + KR[7] * n[1] + KR[8] * n[2];
    res[0] = u2 - p0 / p2;
    res[1] = v2 - p1 / p2;
    return;
  }
  T X = R[0] * n[0] + R[1] * n[1] + R[2] * n[2];
  T Y = R[3] * n[0] + R[4] * n[1] + R[5] * n[2];
  T Z = R[6] * n[0] + R[7] * n[1] + R[8] * n[2];
  T x = X / Z, y = Y / Z, xd, yd;
  brown_ref(x, y, cam[10], cam[11], cam[12], cam[13], cam[14], xd, yd);
  res[0] = u2 - (fx * xd + cx);
  res[1] = v2 - (fy * yd + cy);
}

// Factor2d3dDist / Factor2d3dFxfyDist (krt_optimizer.cc:201-248) via cv::projectPoints [ext]:
// X = R*P + t, x = X/Z, OpenCV distortion order (k1,k2,p1,p2,k3) on v[10..14].
// (cam.rvec() re-derives rvec from R by cv::Rodrigues; for a valid rotation that is the identity map.)
template <class T>
void krt_2d3d_factor(int type, const T* cam, const double P[3], double u, double v, T* res) {
  T R[9];
  RodSel<T>::run(cam + 4, R);
  T fx = cam[0], fy = (type == PTZ_KRT_FXFY || type == PTZ_KRT_FXFYDIST) ? cam[1] : cam[0], cx = cam[2], cy = cam[3];
  T X = R[0] * P[0] + R[1] * P[1] + R[2] * P[2] + cam[7];
  T Y = R[3] * P[0] + R[4] * P[1] + R[5] * P[2] + cam[8];
  T Z = R[6] * P[0] + R[7] * P[1] + R[8] * P[2] + cam[9];
  T x = X / Z, y = Y / Z;
  T k1 = cam[10], k2 = cam[11], p1 = cam[12], p2 = cam[13], k3 = cam[14];
  T r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
  T a1 = 2.0 * x * y, a2 = r2 + 2.0 * x * x, a3 = r2 + 2.0 * y * y;
  T cdist = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
  T xd = x * cdist + p1 * a1 + p2 * a2;
  T yd = y * cdist + p1 * a3 + p2 * a1;
  res[0] = u - (xd * fx + cx);
  res[1] = v - (yd * fy + cy);
}

// =====================================================================================================
// generic least-squares problem: ambient parameter vector, free-coordinate list (SubsetParameterization),
// residual blocks of size 2 touching a few ambient coordinates
// =====================================================================================================
const int kMaxCols = 24;  // 9+6+3+6 (Reproj2d3dDispFactor)
typedef Jet<kMaxCols> JetT;

struct BlockJ {          // one residual block evaluated: weighted residual and tangent Jacobian
  double r[2];
  int nc;                // tangent columns touched
  int col[kMaxCols];     // tangent index of each
  double J[2][kMaxCols];
};

struct Problem {
  int num_ambient = 0, num_tangent = 0, num_blocks = 0;
  std::vector<int> amb2tan;  // ambient index -> tangent index or -1 (fixed)
  std::vector<int> tan2amb;
  std::vector<char> amb_active;  // coordinate belongs to a parameter block that is in the Ceres problem
  virtual ~Problem() {}
  // ambient indices of the global parameter blocks of residual block k, in functor argument order
  virtual int block_params(int k, int* amb_idx) const = 0;
  virtual double block_weight(int k) const = 0;
  virtual void eval_d(int k, const double* args, double* res) const = 0;  // args = gathered ambient values
  virtual void eval_j(int k, const JetT* args, JetT* res) const = 0;
  // optional structure for the Schur path
  int ray_tan_begin = 0, ray_tan_end = 0;  // tangent range eliminated first (empty for KRT)
  int cam_block = 0, num_cam_blocks = 0;   // reduced columns [0, cam_block*num_cam_blocks) come in per-view blocks; the rest is one border block
};

// evaluate one block: raw residual, raw Jacobian over ALL its global coordinates (mode 0 exact, 1 Ceres CENTRAL)
void eval_block_raw(const Problem& P, int k, const double* x, int mode, double res[2], int* na, int* amb, double Jg[2][kMaxCols], bool want_j) {
  *na = P.block_params(k, amb);
  double a[kMaxCols];
  for (int i = 0; i < *na; ++i) a[i] = x[amb[i]];
  P.eval_d(k, a, res);
  if (!want_j) return;
  if (mode == 0) {
    JetT ja[kMaxCols], jr[2];
    for (int i = 0; i < *na; ++i) ja[i] = JetT(a[i], i);
    P.eval_j(k, ja, jr);
    for (int i = 0; i < *na; ++i) { Jg[0][i] = jr[0].v[i]; Jg[1][i] = jr[1].v[i]; }
  } else {
    // ceres::internal::NumericDiff<CENTRAL>: delta = max(|x|*1e-6, sqrt(eps)); (f(x+d)-f(x-d)) * (1/d/2)
    const double min_step = std::sqrt(std::numeric_limits<double>::epsilon());
    for (int i = 0; i < *na; ++i) {
      double delta = std::max(min_step, std::fabs(a[i]) * 1e-6);
      double save = a[i], rp[2], rm[2];
      a[i] = save + delta;
      P.eval_d(k, a, rp);
      a[i] = save - delta;
      P.eval_d(k, a, rm);
      a[i] = save;
      double one_over_delta = 1.0 / delta;
      one_over_delta /= 2;
      Jg[0][i] = (rp[0] - rm[0]) * one_over_delta;
      Jg[1][i] = (rp[1] - rm[1]) * one_over_delta;
    }
  }
}

// ResidualBlock::Evaluate with ScaledLoss(NULL, w): residual and Jacobian times sqrt(w), cost 1/2 w |r|^2
double eval_all(const Problem& P, const double* x, int mode, std::vector<BlockJ>* out, int nthreads) {
  double cost = 0;
  const bool want_j = out != nullptr;
  if (want_j) out->resize(P.num_blocks);
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : cost) schedule(static) num_threads(nthreads)
#endif
  for (int k = 0; k < P.num_blocks; ++k) {
    double res[2], Jg[2][kMaxCols];
    int na, amb[kMaxCols];
    eval_block_raw(P, k, x, mode, res, &na, amb, Jg, want_j);
    double w = P.block_weight(k), sw = std::sqrt(w);
    cost += 0.5 * w * (res[0] * res[0] + res[1] * res[1]);
    if (want_j) {
      BlockJ& b = (*out)[k];
      b.r[0] = res[0] * sw; b.r[1] = res[1] * sw;
      b.nc = 0;
      for (int i = 0; i < na; ++i) {
        int t = P.amb2tan[amb[i]];
        if (t < 0) continue;
        b.col[b.nc] = t; b.J[0][b.nc] = Jg[0][i] * sw; b.J[1][b.nc] = Jg[1][i] * sw;
        ++b.nc;
      }
    }
  }
  (void)nthreads;
  return cost;
}

// out[col] += f(block, c) over every (block, tangent column) pair, in parallel: thread-private accumulators over contiguous
// ranges of blocks, merged in thread order (the result depends on the thread count only in the last bits)
template <class F>
void scatter_cols(const std::vector<BlockJ>& B, int nt, int nthreads, std::vector<double>& out, F f) {
  std::fill(out.begin(), out.end(), 0.0);
  const int nb = (int)B.size();
  int T = std::max(1, std::min(nthreads, nb / 4096));
  if (T == 1) {
    for (const BlockJ& b : B)
      for (int c = 0; c < b.nc; ++c) out[b.col[c]] += f(b, c);
    return;
  }
  std::vector<double> priv((size_t)T * nt, 0.0);
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(T)
#endif
  for (int t = 0; t < T; ++t) {
    double* o = &priv[(size_t)t * nt];
    const int k0 = (int)((long long)nb * t / T), k1 = (int)((long long)nb * (t + 1) / T);
    for (int k = k0; k < k1; ++k)
      for (int c = 0; c < B[k].nc; ++c) o[B[k].col[c]] += f(B[k], c);
  }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(T)
#endif
  for (int j = 0; j < nt; ++j) {
    double sacc = 0;
    for (int t = 0; t < T; ++t) sacc += priv[(size_t)t * nt + j];
    out[j] = sacc;
  }
}

// ----------------------------------------------------------------------------------------------------
// dense helpers
// ----------------------------------------------------------------------------------------------------
bool cholesky_inplace(std::vector<double>& A, int n) {  // lower triangle, row-major
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    double id = 1.0 / d;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (n - j > 256)
#endif
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      const double* ai = &A[(size_t)i * n];
      const double* aj = &A[(size_t)j * n];
      for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
      A[(size_t)i * n + j] = s * id;
    }
  }
  return true;
}
void cholesky_solve(const std::vector<double>& L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}
bool chol3(const double A[6], double L[6]) {  // A = [a00 a10 a11 a20 a21 a22] lower
  double l00 = A[0]; if (!(l00 > 0)) return false; l00 = std::sqrt(l00);
  double l10 = A[1] / l00;
  double l11 = A[2] - l10 * l10; if (!(l11 > 0)) return false; l11 = std::sqrt(l11);
  double l20 = A[3] / l00;
  double l21 = (A[4] - l20 * l10) / l11;
  double l22 = A[5] - l20 * l20 - l21 * l21; if (!(l22 > 0)) return false; l22 = std::sqrt(l22);
  L[0] = l00; L[1] = l10; L[2] = l11; L[3] = l20; L[4] = l21; L[5] = l22;
  return true;
}
inline void chol3_solve(const double L[6], double b[3]) {
  b[0] = b[0] / L[0];
  b[1] = (b[1] - L[1] * b[0]) / L[2];
  b[2] = (b[2] - L[3] * b[0] - L[4] * b[1]) / L[5];
  b[2] = b[2] / L[5];
  b[1] = (b[1] - L[4] * b[2]) / L[2];
  b[0] = (b[0] - L[1] * b[1] - L[3] * b[2]) / L[0];
}

// ----------------------------------------------------------------------------------------------------
// linear solvers: minimise |J y - r|^2 + |D y|^2  (LevenbergMarquardtStrategy::ComputeStep solves J y = r)
// ----------------------------------------------------------------------------------------------------
// DENSE_QR [ext]: Householder QR of the (2N+n) x n matrix [J; D]
bool solve_dense_qr(const std::vector<BlockJ>& B, int n, const double* D, double* y) {
  const int m = (int)B.size() * 2 + n;
  std::vector<double> A((size_t)m * n, 0.0), b(m, 0.0);
  for (size_t k = 0; k < B.size(); ++k)
    for (int r = 0; r < 2; ++r) {
      for (int c = 0; c < B[k].nc; ++c) A[(2 * k + r) * n + B[k].col[c]] += B[k].J[r][c];
      b[2 * k + r] = B[k].r[r];
    }
  for (int j = 0; j < n; ++j) A[(size_t)(2 * B.size() + j) * n + j] = D[j];
  for (int j = 0; j < n; ++j) {
    double nrm = 0;
    for (int i = j; i < m; ++i) nrm += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    nrm = std::sqrt(nrm);
    if (nrm == 0) return false;
    double alpha = A[(size_t)j * n + j] > 0 ? -nrm : nrm;
    std::vector<double> v(m - j);
    for (int i = j; i < m; ++i) v[i - j] = A[(size_t)i * n + j];
    v[0] -= alpha;
    double vn = 0;
    for (double t : v) vn += t * t;
    if (vn == 0) continue;
    for (int c = j; c < n; ++c) {
      double d = 0;
      for (int i = j; i < m; ++i) d += v[i - j] * A[(size_t)i * n + c];
      d = 2 * d / vn;
      for (int i = j; i < m; ++i) A[(size_t)i * n + c] -= d * v[i - j];
    }
    double d = 0;
    for (int i = j; i < m; ++i) d += v[i - j] * b[i];
    d = 2 * d / vn;
    for (int i = j; i < m; ++i) b[i] -= d * v[i - j];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= A[(size_t)i * n + k] * y[k];
    y[i] = s / A[(size_t)i * n + i];
  }
  return true;
}

// SPARSE_SCHUR [ext]: eliminate the ray blocks (tangent range [rb,re), 3 columns each), solve the reduced system
// exactly (dense Cholesky) or by block-Jacobi PCG to tight tolerance, back-substitute.
//
// Threading (Ceres runs its SchurEliminator on options.num_threads = 32 threads, ptzray_optimizer.cc:473): the rays are
// eliminated in parallel (one track per task), and the reduced system is accumulated in parallel BY BLOCK ROW -- a task owns
// the rows of one camera block and walks the tracks / residual blocks that touch it, so no two threads write the same entry
// and every sum has a fixed order (results do not depend on the thread count).  The index structure is built once per solve.
struct SchurWork {
  std::vector<std::vector<int>> track_blocks;  // residual blocks of each ray
  std::vector<int> cam_blocks;                 // residual blocks with no ray column
  // per track: reduced camera columns it touches (grouped by camera block, ascending), and the block of each
  std::vector<std::vector<int>> tcols;
  // per block row: (track, first, count) -> tcols[track][first .. first+count) are the columns of this block
  struct RowRef { int track, first, count; };
  std::vector<std::vector<RowRef>> row_tracks;
  std::vector<std::vector<int>> row_res;       // per block row: residual blocks with a column in this block
  // block-sparse pattern (sparse mode): per block row the sorted block columns and the slot of each
  std::vector<std::vector<int>> pat_col, pat_slot;
  int nslots = 0;
  bool built = false;
};
bool solve_schur(const Problem& P, const std::vector<BlockJ>& B, const double* D, double* y, SchurWork& W, int linear_solver,
                 const ptz_solver_options& opt, int* lin_iters) {
  const int rb = P.ray_tan_begin, re = P.ray_tan_end, nt = P.num_tangent;
  const int nP = (re - rb) / 3;
  const int nC = nt - (re - rb);
  const int nthreads = opt.num_threads > 0 ? opt.num_threads : 1;
  (void)nthreads;
  auto cidx = [&](int t) { return t < rb ? t : t - (re - rb); };  // tangent -> reduced index
  auto is_ray = [&](int t) { return t >= rb && t < re; };
  // camera blocks own their rows; everything behind them is one border block.  (The dense solver keeps the same ownership
  // for the accumulation; only the storage differs.)
  const bool sparse = (linear_solver == 1);
  const int bs = P.cam_block > 0 ? P.cam_block : 1, nvb = P.cam_block > 0 ? P.num_cam_blocks : 0, nblk = nvb + 1, bbs = nC - nvb * bs;
  auto blk_of = [&](int c) { return c < nvb * bs ? c / bs : nvb; };
  auto blk_off = [&](int c) { return c < nvb * bs ? c % bs : c - nvb * bs; };
  auto blk_size = [&](int b) { return b < nvb ? bs : bbs; };
  const int mbs = std::max(bs, bbs);
  if (!W.built) {
    W.track_blocks.assign(nP, std::vector<int>());
    W.cam_blocks.clear();
    for (int k = 0; k < (int)B.size(); ++k) {
      int p = -1;
      for (int c = 0; c < B[k].nc; ++c)
        if (is_ray(B[k].col[c])) { p = (B[k].col[c] - rb) / 3; break; }
      if (p >= 0) W.track_blocks[p].push_back(k); else W.cam_blocks.push_back(k);
    }
    W.tcols.assign(nP, std::vector<int>());
    W.row_tracks.assign(nblk, std::vector<SchurWork::RowRef>());
    W.row_res.assign(nblk, std::vector<int>());
    std::vector<std::vector<int>> pat(nblk);
    auto add_pairs = [&](const std::vector<int>& blocks) {
      for (int a : blocks) for (int c : blocks) pat[a].push_back(c);
    };
    std::vector<int> blocks;
    for (int p = 0; p < nP; ++p) {
      std::vector<int>& tc = W.tcols[p];
      for (int k : W.track_blocks[p])
        for (int a = 0; a < B[k].nc; ++a)
          if (!is_ray(B[k].col[a])) tc.push_back(cidx(B[k].col[a]));
      std::sort(tc.begin(), tc.end());
      tc.erase(std::unique(tc.begin(), tc.end()), tc.end());
      blocks.clear();
      for (size_t i = 0; i < tc.size();) {
        const int b = blk_of(tc[i]);
        size_t j = i;
        while (j < tc.size() && blk_of(tc[j]) == b) ++j;
        W.row_tracks[b].push_back({p, (int)i, (int)(j - i)});
        blocks.push_back(b);
        i = j;
      }
      if (sparse) add_pairs(blocks);
    }
    for (int k = 0; k < (int)B.size(); ++k) {
      blocks.clear();
      for (int a = 0; a < B[k].nc; ++a)
        if (!is_ray(B[k].col[a])) blocks.push_back(blk_of(cidx(B[k].col[a])));
      std::sort(blocks.begin(), blocks.end());
      blocks.erase(std::unique(blocks.begin(), blocks.end()), blocks.end());
      for (int b : blocks) W.row_res[b].push_back(k);
      if (sparse) add_pairs(blocks);
    }
    W.pat_col.assign(nblk, std::vector<int>());
    W.pat_slot.assign(nblk, std::vector<int>());
    W.nslots = 0;
    if (sparse)
      for (int b = 0; b < nblk; ++b) {
        pat[b].push_back(b);  // the diagonal block always exists (LM diagonal)
        std::sort(pat[b].begin(), pat[b].end());
        pat[b].erase(std::unique(pat[b].begin(), pat[b].end()), pat[b].end());
        W.pat_col[b] = pat[b];
        W.pat_slot[b].resize(pat[b].size());
        for (size_t i = 0; i < pat[b].size(); ++i) W.pat_slot[b][i] = W.nslots++;
      }
    W.built = true;
  }
  std::vector<double> S(sparse ? 0 : (size_t)nC * nC, 0.0), rhs(nC, 0.0);
  std::vector<double> blocks(sparse ? (size_t)W.nslots * mbs * mbs : 0, 0.0);  // slot -> mbs*mbs values
  auto slot_of = [&](int bi, int bj) {
    const std::vector<int>& pc = W.pat_col[bi];
    return W.pat_slot[bi][std::lower_bound(pc.begin(), pc.end(), bj) - pc.begin()];
  };
  // ---- phase A: per track, V = E^T E + D^2 = L L^T, t = L^-1 h, What = (F^T E) L^-T for every camera column it touches
  std::vector<double> Lp((size_t)nP * 6), tp((size_t)nP * 3);
  std::vector<char> pok(nP, 0);
  std::vector<size_t> woff(nP + 1, 0);
  for (int p = 0; p < nP; ++p) woff[p + 1] = woff[p] + W.tcols[p].size() * 3;
  std::vector<double> What(woff[nP], 0.0);
  int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) reduction(| : bad)
#endif
  for (int p = 0; p < nP; ++p) {
    const std::vector<int>& tb = W.track_blocks[p];
    if (tb.empty()) continue;
    const std::vector<int>& tc = W.tcols[p];
    double* Wm = &What[woff[p]];
    double Vm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, h[3] = {0, 0, 0};
    for (int k : tb) {
      const BlockJ& b = B[k];
      double E[2][3] = {{0, 0, 0}, {0, 0, 0}};
      for (int a = 0; a < b.nc; ++a)
        if (is_ray(b.col[a])) { int e = (b.col[a] - rb) % 3; E[0][e] = b.J[0][a]; E[1][e] = b.J[1][a]; }
      for (int i = 0; i < 3; ++i) {
        h[i] += E[0][i] * b.r[0] + E[1][i] * b.r[1];
        for (int j = 0; j < 3; ++j) Vm[i][j] += E[0][i] * E[0][j] + E[1][i] * E[1][j];
      }
      for (int a = 0; a < b.nc; ++a)
        if (!is_ray(b.col[a])) {
          const int li = (int)(std::lower_bound(tc.begin(), tc.end(), cidx(b.col[a])) - tc.begin());
          for (int j = 0; j < 3; ++j) Wm[li * 3 + j] += b.J[0][a] * E[0][j] + b.J[1][a] * E[1][j];
        }
    }
    for (int i = 0; i < 3; ++i) Vm[i][i] += D[rb + 3 * p + i] * D[rb + 3 * p + i];
    double A6[6] = {Vm[0][0], Vm[1][0], Vm[1][1], Vm[2][0], Vm[2][1], Vm[2][2]};
    if (!chol3(A6, &Lp[(size_t)p * 6])) { bad |= 1; continue; }
    pok[p] = 1;
    const double* L = &Lp[(size_t)p * 6];
    double t0 = h[0] / L[0], t1 = (h[1] - L[1] * t0) / L[2], t2 = (h[2] - L[3] * t0 - L[4] * t1) / L[5];
    tp[3 * p] = t0; tp[3 * p + 1] = t1; tp[3 * p + 2] = t2;
    for (size_t a = 0; a < tc.size(); ++a) {
      double w0 = Wm[a * 3] / L[0], w1 = (Wm[a * 3 + 1] - L[1] * w0) / L[2], w2 = (Wm[a * 3 + 2] - L[3] * w0 - L[4] * w1) / L[5];
      Wm[a * 3] = w0; Wm[a * 3 + 1] = w1; Wm[a * 3 + 2] = w2;
    }
  }
  if (bad) return false;
  // ---- phase B: one task per block row: D^2, F^T F and F^T r of its residual blocks, minus the Schur terms of its tracks
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
  for (int bi = 0; bi < nblk; ++bi) {
    const int rows_n = blk_size(bi);
    if (rows_n == 0) continue;
    const int row0 = bi < nvb ? bi * bs : nvb * bs;
    auto Sadd = [&](int i, int j, double v) {  // i is a reduced row of block bi
      if (!sparse) { S[(size_t)i * nC + j] += v; return; }
      blocks[(size_t)slot_of(bi, blk_of(j)) * mbs * mbs + blk_off(i) * mbs + blk_off(j)] += v;
    };
    for (int i = row0; i < row0 + rows_n; ++i) {
      const int t = i < rb ? i : i + (re - rb);  // reduced -> tangent
      Sadd(i, i, D[t] * D[t]);
    }
    for (int k : W.row_res[bi]) {
      const BlockJ& b = B[k];
      for (int a = 0; a < b.nc; ++a) {
        if (is_ray(b.col[a])) continue;
        const int ia = cidx(b.col[a]);
        if (blk_of(ia) != bi) continue;
        rhs[ia] += b.J[0][a] * b.r[0] + b.J[1][a] * b.r[1];
        for (int c = 0; c < b.nc; ++c) {
          if (is_ray(b.col[c])) continue;
          Sadd(ia, cidx(b.col[c]), b.J[0][a] * b.J[0][c] + b.J[1][a] * b.J[1][c]);
        }
      }
    }
    for (const SchurWork::RowRef& rr : W.row_tracks[bi]) {
      const int p = rr.track;
      if (!pok[p]) continue;
      const std::vector<int>& tc = W.tcols[p];
      const double* Wm = &What[woff[p]];
      const double t0 = tp[3 * p], t1 = tp[3 * p + 1], t2 = tp[3 * p + 2];
      for (int a = rr.first; a < rr.first + rr.count; ++a) {
        rhs[tc[a]] -= Wm[a * 3] * t0 + Wm[a * 3 + 1] * t1 + Wm[a * 3 + 2] * t2;
        for (size_t c = 0; c < tc.size(); ++c)
          Sadd(tc[a], tc[c], -(Wm[a * 3] * Wm[c * 3] + Wm[a * 3 + 1] * Wm[c * 3 + 1] + Wm[a * 3 + 2] * Wm[c * 3 + 2]));
      }
    }
  }
  std::vector<double> yc(rhs);
  *lin_iters = 0;
  if (!sparse) {
    std::vector<double> L(S);
    if (!cholesky_inplace(L, nC)) return false;
    cholesky_solve(L, nC, yc.data());
  } else {
    // block-Jacobi preconditioned CG on the block-sparse reduced system, run to opt.pcg_rel_tolerance ("exact" stand-in
    // for the sparse Cholesky of SPARSE_SCHUR on problems where a dense factorisation would dominate the CPU baseline)
    std::vector<std::vector<double>> Minv(nblk);
    int bad_block = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(| : bad_block)
#endif
    for (int b = 0; b < nblk; ++b) {
      const int n = blk_size(b);
      if (n == 0) continue;
      std::vector<double> A((size_t)n * n);
      const double* Bd = &blocks[(size_t)slot_of(b, b) * mbs * mbs];
      for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = Bd[i * mbs + j];
      if (!cholesky_inplace(A, n)) { bad_block |= 1; continue; }
      Minv[b].assign((size_t)n * n, 0.0);
      for (int c = 0; c < n; ++c) {
        std::vector<double> e(n, 0.0);
        e[c] = 1.0;
        cholesky_solve(A, n, e.data());
        for (int i = 0; i < n; ++i) Minv[b][(size_t)i * n + c] = e[i];
      }
    }
    if (bad_block) return false;
    auto boff = [&](int b) { return b < nvb ? b * bs : nvb * bs; };
    auto apply_M = [&](const std::vector<double>& r, std::vector<double>& z) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
      for (int b = 0; b < nblk; ++b) {
        const int n = blk_size(b), o = boff(b);
        for (int i = 0; i < n; ++i) { double sacc = 0; for (int j = 0; j < n; ++j) sacc += Minv[b][(size_t)i * n + j] * r[o + j]; z[o + i] = sacc; }
      }
    };
    auto apply_S = [&](const std::vector<double>& x, std::vector<double>& out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
#endif
      for (int bi = 0; bi < nblk; ++bi) {
        const int ni = blk_size(bi), oi = boff(bi);
        for (int i = 0; i < ni; ++i) out[oi + i] = 0;
        for (size_t e = 0; e < W.pat_col[bi].size(); ++e) {
          const int bj = W.pat_col[bi][e], nj = blk_size(bj), oj = boff(bj);
          const double* Bd = &blocks[(size_t)W.pat_slot[bi][e] * mbs * mbs];
          for (int i = 0; i < ni; ++i) { double sacc = 0; for (int j = 0; j < nj; ++j) sacc += Bd[i * mbs + j] * x[oj + j]; out[oi + i] += sacc; }
        }
      }
    };
    std::vector<double> x(nC, 0.0), r(rhs), z(nC), p(nC), Ap(nC);
    double bn = 0;
    for (int i = 0; i < nC; ++i) bn += rhs[i] * rhs[i];
    bn = std::sqrt(bn);
    apply_M(r, z);
    p = z;
    double rz = 0;
    for (int i = 0; i < nC; ++i) rz += r[i] * z[i];
    int it = 0;
    for (; it < opt.pcg_max_iterations && bn > 0;) {
      apply_S(p, Ap);
      double pAp = 0;
      for (int i = 0; i < nC; ++i) pAp += p[i] * Ap[i];
      if (!(pAp > 0)) return false;
      double alpha = rz / pAp, rn = 0;
      for (int i = 0; i < nC; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; rn += r[i] * r[i]; }
      ++it;
      if (std::sqrt(rn) <= opt.pcg_rel_tolerance * bn) break;
      apply_M(r, z);
      double rz2 = 0;
      for (int i = 0; i < nC; ++i) rz2 += r[i] * z[i];
      double beta = rz2 / rz;
      rz = rz2;
      for (int i = 0; i < nC; ++i) p[i] = z[i] + beta * p[i];
    }
    *lin_iters = it;
    yc = x;
  }
  for (int t = 0; t < nt; ++t)
    if (t < rb || t >= re) y[t] = yc[cidx(t)];
  // back-substitution: y_p = L^-T (t_p - What^T y_c)
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
#endif
  for (int p = 0; p < nP; ++p) {
    if (!pok[p]) { y[rb + 3 * p] = y[rb + 3 * p + 1] = y[rb + 3 * p + 2] = 0; continue; }
    const double* L = &Lp[(size_t)p * 6];
    double acc[3] = {0, 0, 0};
    for (int k : W.track_blocks[p]) {
      const BlockJ& b = B[k];
      double E[2][3] = {{0, 0, 0}, {0, 0, 0}}, fy[2] = {0, 0};
      for (int a = 0; a < b.nc; ++a) {
        if (is_ray(b.col[a])) { int e = (b.col[a] - rb) % 3; E[0][e] = b.J[0][a]; E[1][e] = b.J[1][a]; }
        else { fy[0] += b.J[0][a] * y[b.col[a]]; fy[1] += b.J[1][a] * y[b.col[a]]; }
      }
      for (int j = 0; j < 3; ++j) acc[j] += E[0][j] * fy[0] + E[1][j] * fy[1];  // W^T y_c = E^T F y_c
    }
    // solve (V+D^2) y_p = h - W^T y_c, with h = L t
    double hh[3] = {L[0] * tp[3 * p], L[1] * tp[3 * p] + L[2] * tp[3 * p + 1], L[3] * tp[3 * p] + L[4] * tp[3 * p + 1] + L[5] * tp[3 * p + 2]};
    double bb[3] = {hh[0] - acc[0], hh[1] - acc[1], hh[2] - acc[2]};
    chol3_solve(L, bb);
    y[rb + 3 * p] = bb[0]; y[rb + 3 * p + 1] = bb[1]; y[rb + 3 * p + 2] = bb[2];
  }
  return true;
}

// ----------------------------------------------------------------------------------------------------
// TrustRegionMinimizer + LevenbergMarquardtStrategy, Ceres 1.14.0 [ext]
// ----------------------------------------------------------------------------------------------------
struct LmSummary {
  int termination = PTZ_NO_CONVERGENCE;
  int num_iterations = 0, num_successful = 0, num_unsuccessful = 0, lin_iters = 0;
  double initial_cost = 0, final_cost = 0;
  std::vector<ptz_iter_log> log;
};

void minimize(const Problem& P, std::vector<double>& x, const ptz_solver_options& o, bool use_schur, LmSummary& sum) {
  const int nt = P.num_tangent, na = P.num_ambient;
  const bool timing = std::getenv("ORC_TIMING") != nullptr;  // where an iteration's time goes (stderr)
  double t_eval = 0, t_grad = 0, t_lin = 0, t_model = 0, t_cost = 0;
  auto now = []() { return omp_get_wtime(); };
  const int nthreads = o.num_threads > 0 ? o.num_threads : 1;
  std::vector<BlockJ> B;
  std::vector<double> scale(nt, 1.0), grad(nt), diag(nt), D(nt), ystep(nt), delta(nt), cand(x);
  SchurWork SW;
  auto xnorm = [&](const std::vector<double>& v) { double s = 0; for (int i = 0; i < na; ++i) if (P.amb_active[i]) s += v[i] * v[i]; return std::sqrt(s); };
  double x_cost = 0, x_norm = xnorm(x), radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int iteration = 0, num_consecutive_invalid = 0;
  double grad_max = 0;
  // EvaluateGradientAndJacobian
  auto eval_gj = [&]() {
    double t0 = now();
    x_cost = eval_all(P, x.data(), o.jacobian_mode, &B, nthreads);
    t_eval += now() - t0; t0 = now();
    scatter_cols(B, nt, nthreads, grad, [](const BlockJ& b, int c) { return b.J[0][c] * b.r[0] + b.J[1][c] * b.r[1]; });  // unscaled J
    if (o.jacobi_scaling) {
      if (iteration == 0) {
        std::vector<double> cn(nt, 0.0);
        scatter_cols(B, nt, nthreads, cn, [](const BlockJ& b, int c) { return b.J[0][c] * b.J[0][c] + b.J[1][c] * b.J[1][c]; });
        for (int j = 0; j < nt; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(cn[j]));
      }
      const int nbk = (int)B.size();
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
      for (int k = 0; k < nbk; ++k) {
        BlockJ& b = B[k];
        for (int c = 0; c < b.nc; ++c) { b.J[0][c] *= scale[b.col[c]]; b.J[1][c] *= scale[b.col[c]]; }
      }
    }
    // |x - Plus(x, -g)|_inf
    grad_max = 0;
    for (int j = 0; j < nt; ++j) {
      double xv = x[P.tan2amb[j]];
      double d = std::fabs(xv - (xv + (-grad[j])));
      if (d > grad_max) grad_max = d;
    }
    t_grad += now() - t0;
  };
  auto push_log = [&](double cost, double cost_change, double step_norm, double rho, int lin, int ok) {
    ptz_iter_log l;
    l.cost = cost; l.cost_change = cost_change; l.gradient_max_norm = grad_max; l.step_norm = step_norm;
    l.relative_decrease = rho; l.trust_region_radius = radius; l.linear_solver_iterations = lin; l.step_is_successful = ok;
    sum.log.push_back(l);
  };
  // IterationZero
  eval_gj();
  sum.initial_cost = x_cost;
  double min_cost = x_cost;
  bool last_successful = true;
  ++sum.num_successful;
  push_log(x_cost, 0, 0, 0, 0, 1);
  if (o.verbose) std::printf("[orc] it %3d cost %.10e |g| %.3e radius %.3e\n", 0, x_cost, grad_max, radius);
  sum.termination = PTZ_NO_CONVERGENCE;
  while (true) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue (checks only)
    if (iteration >= o.max_num_iterations) { sum.termination = PTZ_NO_CONVERGENCE; break; }
    if (last_successful && grad_max <= o.gradient_tolerance) { sum.termination = PTZ_CONVERGENCE; break; }
    if (radius <= o.min_trust_region_radius) { sum.termination = PTZ_CONVERGENCE; break; }
    ++iteration;
    // ComputeTrustRegionStep -> LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      scatter_cols(B, nt, nthreads, diag, [](const BlockJ& b, int c) { return b.J[0][c] * b.J[0][c] + b.J[1][c] * b.J[1][c]; });
      for (int j = 0; j < nt; ++j) diag[j] = std::min(std::max(diag[j], o.min_lm_diagonal), o.max_lm_diagonal);
    }
    for (int j = 0; j < nt; ++j) D[j] = std::sqrt(diag[j] / radius);
    int lin = 0;
    double t0 = now();
    bool ok = use_schur ? solve_schur(P, B, D.data(), ystep.data(), SW, o.linear_solver, o, &lin) : solve_dense_qr(B, nt, D.data(), ystep.data());
    t_lin += now() - t0; t0 = now();
    reuse_diagonal = true;
    sum.lin_iters += lin;
    if (ok)
      for (int j = 0; j < nt; ++j) if (!std::isfinite(ystep[j])) { ok = false; break; }
    double model_cost_change = 0;
    if (ok) {
      for (int j = 0; j < nt; ++j) ystep[j] = -ystep[j];
      // model_cost_change = -(J step)^T (r + J step / 2)
      {
        const int nbk = (int)B.size(), Tm = std::max(1, std::min(nthreads, nbk / 4096));
        std::vector<double> part(Tm, 0.0);
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(Tm)
#endif
        for (int t = 0; t < Tm; ++t) {
          double sacc = 0;
          for (int k = (int)((long long)nbk * t / Tm); k < (int)((long long)nbk * (t + 1) / Tm); ++k) {
            const BlockJ& b = B[k];
            double m0 = 0, m1 = 0;
            for (int c = 0; c < b.nc; ++c) { m0 += b.J[0][c] * ystep[b.col[c]]; m1 += b.J[1][c] * ystep[b.col[c]]; }
            sacc += m0 * (b.r[0] + m0 / 2.0) + m1 * (b.r[1] + m1 / 2.0);
          }
          part[t] = sacc;
        }
        for (int t = 0; t < Tm; ++t) model_cost_change -= part[t];
      }
      ok = model_cost_change > 0.0;
    }
    t_model += now() - t0;
    if (!ok) {
      // HandleInvalidStep
      ++num_consecutive_invalid;
      if (num_consecutive_invalid >= o.max_num_consecutive_invalid_steps) { sum.termination = PTZ_FAILURE; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepIsInvalid -> StepRejected(0)
      last_successful = false;
      ++sum.num_unsuccessful;
      push_log(x_cost, 0, 0, 0, lin, -1);
      continue;
    }
    num_consecutive_invalid = 0;
    for (int j = 0; j < nt; ++j) delta[j] = ystep[j] * scale[j];
    // ComputeCandidatePointAndEvaluateCost
    cand = x;
    for (int j = 0; j < nt; ++j) cand[P.tan2amb[j]] = x[P.tan2amb[j]] + delta[j];
    t0 = now();
    double cand_cost = eval_all(P, cand.data(), 0, nullptr, nthreads);
    t_cost += now() - t0;
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
    // ParameterToleranceReached
    double step_norm = 0;
    for (int i = 0; i < na; ++i) if (P.amb_active[i]) step_norm += (x[i] - cand[i]) * (x[i] - cand[i]);
    step_norm = std::sqrt(step_norm);
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { sum.termination = PTZ_CONVERGENCE; break; }
    // FunctionToleranceReached
    double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= o.function_tolerance * x_cost) { sum.termination = PTZ_CONVERGENCE; break; }
    // IsStepSuccessful (monotonic steps: StepQuality = (cost - candidate)/model_cost_change)
    double rho = cost_change / model_cost_change;
    if (rho > o.min_relative_decrease) {
      // HandleSuccessfulStep
      x = cand;
      x_norm = xnorm(x);
      eval_gj();
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      last_successful = true;
      ++sum.num_successful;
      if (x_cost < min_cost) min_cost = x_cost;
      push_log(x_cost, cost_change, step_norm, rho, lin, 1);
    } else {
      // HandleUnsuccessfulStep
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      last_successful = false;
      ++sum.num_unsuccessful;
      push_log(cand_cost, cost_change, step_norm, rho, lin, 0);
    }
    if (o.verbose)
      std::printf("[orc] it %3d cost %.10e change %.3e |g| %.3e |step| %.3e rho %.3e radius %.3e lin %d %s\n", iteration, sum.log.back().cost,
                  cost_change, grad_max, step_norm, rho, radius, lin, last_successful ? "ok" : "rej");
  }
  if (timing)
    std::fprintf(stderr, "[orc timing, %d threads] jacobian %.3f s  gradient/scaling %.3f s  linear solve %.3f s  model change %.3f s  cost %.3f s\n", nthreads,
                 t_eval, t_grad, t_lin, t_model, t_cost);
  sum.num_iterations = (int)sum.log.size() - 1;
  sum.final_cost = min_cost;  // SetSummaryFinalCost: min over accepted iterates == cost at returned x
}

// =====================================================================================================
// the BA problem (PTZRayOptimizer)
// =====================================================================================================
struct BaProblem : Problem {
  const ptzba_problem* p;
  int type, V, Pn, M, A;
  int off_ext, off_ray, off_disp, off_tlw;  // ambient offsets: intr V*9 | ext V*6 | ray P*3 | disp 3 | tlw 6
  int nci, ncv;
  bool use_disp, use_tlw;
  // SetSharedIntrinsics (ptzray_optimizer.cc:497-505): views with the same shared_ic_id use ONE intrinsics block, the one of the
  // first such view (SetUpInitialCameraParams :640-650 inserts the block at the first candidate of the id).  rep[v] = that view.
  std::vector<int> rep;
  BaProblem(const ptzba_problem* pp) : p(pp) {
    type = p->factor_type; V = p->num_views; Pn = p->num_tracks; M = p->num_obs; A = p->num_pts3d;
    off_ext = V * 9; off_ray = off_ext + V * 6; off_disp = off_ray + Pn * 3; off_tlw = off_disp + 3;
    num_ambient = off_tlw + 6 + 1;  // (+1: a scratch coordinate that tangent slots without a parameter point to)
    rep.resize(V);
    {
      std::map<int, int> first;
      for (int i = 0; i < V; ++i) {
        const int id = p->shared_ic_id ? p->shared_ic_id[i] : i;
        auto it = first.find(id);
        if (it == first.end()) { first[id] = i; rep[i] = i; } else rep[i] = it->second;
      }
    }
    std::vector<int> gsize(V, 0);
    for (int i = 0; i < V; ++i) ++gsize[rep[i]];
    num_blocks = M + A;
    use_disp = (type == PTZ_BA_PTZRAY_DIST_DISP);
    use_tlw = A > 0;
    nci = (type == PTZ_BA_PTZRAY) ? 2 : 3;
    ncv = nci + 3;
    amb2tan.assign(num_ambient, -1);
    amb_active.assign(num_ambient, 0);
    std::vector<char> view_used(V, 0), track_used(Pn, 0);
    for (int k = 0; k < M; ++k) { view_used[p->obs_view[k]] = 1; track_used[p->obs_track[k]] = 1; }
    for (int k = 0; k < A; ++k) view_used[p->pt_view[k]] = 1;
    // tangent layout: header ptzba_eval_out.  Views / tracks that appear in no residual block are not in the
    // Ceres problem; they keep tangent slots here (zero columns, zero step) but are excluded from the norms.
    int t = 0;
    for (int i = 0; i < V; ++i) {
      amb2tan[i * 9 + 0] = t++; amb2tan[i * 9 + 1] = t++;              // fx, fy  (:861-872)
      if (nci == 3) amb2tan[i * 9 + 4] = t++;                          // k1
      for (int j = 0; j < 3; ++j) amb2tan[off_ext + i * 6 + j] = t++;  // rvec; t fixed (:881)
      if (view_used[i]) { for (int j = 0; j < 9; ++j) amb_active[i * 9 + j] = 1; for (int j = 0; j < 6; ++j) amb_active[off_ext + i * 6 + j] = 1; }
    }
    ray_tan_begin = t;
    for (int q = 0; q < Pn; ++q)
      for (int j = 0; j < 3; ++j) { amb2tan[off_ray + q * 3 + j] = t++; if (track_used[q]) amb_active[off_ray + q * 3 + j] = 1; }
    ray_tan_end = t;
    if (use_disp) for (int j = 0; j < 3; ++j) { amb2tan[off_disp + j] = t++; amb_active[off_disp + j] = 1; }
    if (use_tlw) for (int j = 0; j < 6; ++j) { amb2tan[off_tlw + j] = t++; amb_active[off_tlw + j] = 1; }
    // shared intrinsics blocks: their free coordinates (fx, fy, (k1)) sit behind everything else, in the border of the reduced
    // system; the per-view slots of the representative keep their place in the camera block but point to no parameter (zero
    // columns, zero step), those of the other members point to their own, unused, intrinsics
    std::vector<int> orphan;
    for (int i = 0; i < V; ++i) {
      if (rep[i] != i || gsize[i] < 2) continue;
      bool used = false;
      for (int v = 0; v < V; ++v) if (rep[v] == i && view_used[v]) used = true;
      const int idx[3] = {0, 1, 4};
      for (int c = 0; c < nci; ++c) { orphan.push_back(amb2tan[i * 9 + idx[c]]); amb2tan[i * 9 + idx[c]] = t++; }
      for (int j = 0; j < 9; ++j) amb_active[i * 9 + j] = used ? 1 : 0;
    }
    for (int i = 0; i < V; ++i)
      if (rep[i] != i) for (int j = 0; j < 9; ++j) amb_active[i * 9 + j] = 0;  // not parameter blocks of the problem
    num_tangent = t;
    tan2amb.assign(t, num_ambient - 1);
    for (int i = 0; i < num_ambient; ++i) if (amb2tan[i] >= 0) tan2amb[amb2tan[i]] = i;
    for (int o : orphan) tan2amb[o] = num_ambient - 1;
    cam_block = ncv; num_cam_blocks = V;
  }
  // functor argument order: 2d-2d: intr(9) [disp(3)] ext(6) ray(3); 2d-3d: intr(9) [disp(3)] ext(6) tlw(6)
  int block_params(int k, int* amb) const override {
    int n = 0;
    int view = k < M ? p->obs_view[k] : p->pt_view[k - M];
    for (int j = 0; j < 9; ++j) amb[n++] = rep[view] * 9 + j;  // intrinsics_param_.at(ic_id), :821-848
    if (use_disp) for (int j = 0; j < 3; ++j) amb[n++] = off_disp + j;
    for (int j = 0; j < 6; ++j) amb[n++] = off_ext + view * 6 + j;
    if (k < M) for (int j = 0; j < 3; ++j) amb[n++] = off_ray + p->obs_track[k] * 3 + j;
    else for (int j = 0; j < 6; ++j) amb[n++] = off_tlw + j;
    return n;
  }
  double block_weight(int k) const override { return k < M ? p->track_weight[p->obs_track[k]] : 1.0; }
  template <class T> void eval_t(int k, const T* a, T* res) const {
    const T* intr = a;
    const T* disp = use_disp ? a + 9 : nullptr;
    const T* ext = a + 9 + (use_disp ? 3 : 0);
    const T* last = ext + 6;
    if (k < M) ba_ray_factor<T>(type, intr, ext, last, disp, p->obs_uv[2 * k], p->obs_uv[2 * k + 1], res);
    else ba_pt_factor<T>(type, intr, ext, last, disp, p->pt_uv[2 * (k - M)], p->pt_uv[2 * (k - M) + 1], p->pt_xyz + 3 * (k - M), res);
  }
  void eval_d(int k, const double* a, double* res) const override { eval_t<double>(k, a, res); }
  void eval_j(int k, const JetT* a, JetT* res) const override { eval_t<JetT>(k, a, res); }
};

// Pix2Ray (ptzray_optimizer.cc:768-797): mean over the candidate views of normalise((R^-1 K^-1) [u,v,1]), normalised
void init_rays(const ptzba_problem* p, const double* intr, const double* ext, double* ray) {
  const int V = p->num_views, Pn = p->num_tracks;
  std::vector<double> RiKi((size_t)V * 9);
  for (int i = 0; i < V; ++i) {
    double R[9], Ri[9], Ki[9];
    rodrigues(ext + 6 * i, R);
    double K[9] = {intr[9 * i], 0, intr[9 * i + 2], 0, intr[9 * i + 1], intr[9 * i + 3], 0, 0, 1};
    inv3(R, Ri);
    inv3(K, Ki);
    mul33(Ri, Ki, &RiKi[(size_t)i * 9]);
  }
  std::vector<double> acc((size_t)Pn * 3, 0.0);
  std::vector<int> cnt(Pn, 0);
  for (int k = 0; k < p->num_obs; ++k) {
    double uv[3] = {(double)p->obs_uv[2 * k], (double)p->obs_uv[2 * k + 1], 1.0}, r[3];
    mul31(&RiKi[(size_t)p->obs_view[k] * 9], uv, r);
    double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    int q = p->obs_track[k];
    acc[3 * q] += r[0] / n; acc[3 * q + 1] += r[1] / n; acc[3 * q + 2] += r[2] / n;
    ++cnt[q];
  }
  for (int q = 0; q < Pn; ++q) {
    double r[3] = {acc[3 * q] / cnt[q], acc[3 * q + 1] / cnt[q], acc[3 * q + 2] / cnt[q]};
    double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    ray[3 * q] = r[0] / n; ray[3 * q + 1] = r[1] / n; ray[3 * q + 2] = r[2] / n;
  }
}

void ba_initial_x(const ptzba_problem* p, const BaProblem& bp, std::vector<double>& x, const double* disp) {
  x.assign(bp.num_ambient, 0.0);
  std::memcpy(&x[0], p->intr, sizeof(double) * 9 * bp.V);
  std::memcpy(&x[bp.off_ext], p->ext, sizeof(double) * 6 * bp.V);
  if (p->ray0) std::memcpy(&x[bp.off_ray], p->ray0, sizeof(double) * 3 * bp.Pn);
  else if (bp.Pn > 0) init_rays(p, p->intr, p->ext, &x[bp.off_ray]);
  if (disp) std::memcpy(&x[bp.off_disp], disp, sizeof(double) * 3);
  if (p->tlw0) std::memcpy(&x[bp.off_tlw], p->tlw0, sizeof(double) * 6);
}

int ba_check(const ptzba_problem* p) {
  if (!p) return PTZ_ERR_INVALID;
  if (p->num_views <= 0 || p->num_tracks < 0 || p->num_obs < 0 || p->num_pts3d < 0) return PTZ_ERR_INVALID;  // CheckValid :515-535
  if (!p->intr || !p->ext) return PTZ_ERR_INVALID;
  if (p->num_obs > 0 && (!p->obs_uv || !p->obs_view || !p->obs_track || !p->track_weight)) return PTZ_ERR_INVALID;
  if (p->num_pts3d > 0 && (!p->pt_uv || !p->pt_xyz || !p->pt_view)) return PTZ_ERR_INVALID;
  if (p->factor_type < 0 || p->factor_type > 3) return PTZ_ERR_INVALID;
  for (int k = 0; k < p->num_obs; ++k)
    if (p->obs_view[k] < 0 || p->obs_view[k] >= p->num_views || p->obs_track[k] < 0 || p->obs_track[k] >= p->num_tracks) return PTZ_ERR_INVALID;
  for (int k = 0; k < p->num_pts3d; ++k)
    if (p->pt_view[k] < 0 || p->pt_view[k] >= p->num_views) return PTZ_ERR_INVALID;
  if (p->shared_ic_id && p->num_pts3d > 0)  // (not restated: shared blocks together with 2d-3d terms)
    for (int i = 0; i < p->num_views; ++i) if (p->shared_ic_id[i] != i) return PTZ_ERR_UNSUPPORTED;
  return PTZ_OK;
}

// world-frame outputs (ObtainRefinedCameraParams, ptzray_optimizer.cc:672-766)
void ba_world_outputs(const BaProblem& bp, const std::vector<double>& x, ptzba_result* out) {
  const double* tlw = &x[bp.off_tlw];
  double Rlw[9];
  rodrigues(tlw, Rlw);
  const double* d = &x[bp.off_disp];
  if (out->cams_world) {
    for (int i = 0; i < bp.V; ++i) {
      const double* in = &x[bp.rep[i] * 9];  // intrinsics_param_.at(ic_id), :679-727
      const double* ex = &x[bp.off_ext + i * 6];
      double* c = out->cams_world + 21 * i;
      double fx = in[0];
      c[0] = fx; c[1] = (bp.type == PTZ_BA_PTZRAY_FXFY_DIST) ? in[1] : in[0]; c[2] = in[2]; c[3] = in[3];
      double R[9];
      rodrigues(ex, R);
      double t[3] = {ex[3], ex[4], ex[5] + (d[0] + d[1] * fx + d[2] * fx * fx)};  // :693-694
      double Rt[3];
      mul31(R, tlw + 3, Rt);                      // t <- R*t_lw + t   (:738)
      for (int j = 0; j < 3; ++j) c[13 + j] = Rt[j] + t[j];
      mul33(R, Rlw, c + 4);                       // R <- R*R_lw       (:739)
      for (int j = 0; j < 5; ++j) c[16 + j] = in[4 + j];
    }
  }
  if (out->rays_world) {
    // ray_w = R_lw^T ray_l - R_lw^T t_lw (:748-754)
    for (int q = 0; q < bp.Pn; ++q) {
      const double* r = &x[bp.off_ray + 3 * q];
      for (int j = 0; j < 3; ++j)
        out->rays_world[3 * q + j] = Rlw[j] * (r[0] - tlw[3]) + Rlw[3 + j] * (r[1] - tlw[4]) + Rlw[6 + j] * (r[2] - tlw[5]);
    }
  }
}

// =====================================================================================================
// the KRT problem (KRTOptimizer), one query
// =====================================================================================================
struct KrtProblem : Problem {
  int type, N, Np = 0;
  const float* uv2;
  const float* puv = nullptr;
  std::vector<KrtPre> pre;
  std::vector<double> plocal;  // 2d-3d points in the reference-local frame (Add2d3dConstraints, krt_optimizer.cc:357-362)
  void add_points(int np, const float* pt_uv, const double* pt_xyz_world, const double* ref21) {
    Np = np; puv = pt_uv;
    plocal.resize(3 * (size_t)np);
    for (int i = 0; i < np; ++i) {
      double t[3];
      mul31(ref21 + 4, pt_xyz_world + 3 * i, t);
      for (int a = 0; a < 3; ++a) plocal[3 * (size_t)i + a] = t[a] + ref21[13 + a];
    }
    num_blocks = N + Np;
  }
  KrtProblem(int type_, int n, const float* uv_ref, const float* uv_cur, const double refK4[4], const double refd[5]) : type(type_), N(n), uv2(uv_cur) {
    num_ambient = 15;
    num_blocks = N;
    pre.resize(N);
    for (int k = 0; k < N; ++k) krt_precompute(type, refK4, refd, uv_ref + 2 * k, &pre[k]);
    amb2tan.assign(15, -1);
    amb_active.assign(15, 1);
    // SubsetParameterization constants, krt_optimizer.cc:318-347
    std::vector<int> freec;
    switch (type) {
      case PTZ_KRT_F: freec = {0, 4, 5, 6}; break;
      case PTZ_KRT_FXFY: freec = {0, 1, 4, 5, 6}; break;
      case PTZ_KRT_FDIST: freec = {0, 4, 5, 6, 10}; break;
      default: freec = {0, 1, 4, 5, 6, 10}; break;
    }
    num_tangent = (int)freec.size();
    tan2amb = freec;
    for (int j = 0; j < num_tangent; ++j) amb2tan[freec[j]] = j;
  }
  int block_params(int, int* amb) const override { for (int j = 0; j < 15; ++j) amb[j] = j; return 15; }
  double block_weight(int) const override { return 1.0; }  // nullptr loss (:294)
  void eval_d(int k, const double* a, double* res) const override {
    if (k < N) krt_2d2d_factor<double>(type, a, pre[k], uv2[2 * k], uv2[2 * k + 1], res);
    else krt_2d3d_factor<double>(type, a, &plocal[3 * (size_t)(k - N)], puv[2 * (k - N)], puv[2 * (k - N) + 1], res);
  }
  void eval_j(int k, const JetT* a, JetT* res) const override {
    if (k < N) krt_2d2d_factor<JetT>(type, a, pre[k], uv2[2 * k], uv2[2 * k + 1], res);
    else krt_2d3d_factor<JetT>(type, a, &plocal[3 * (size_t)(k - N)], puv[2 * (k - N)], puv[2 * (k - N) + 1], res);
  }
};

// Add2d2dConstraints frame change (krt_optimizer.cc:269-286): local = reference frame
void krt_to_local(const double* ref21, const double* init21, double local15[15]) {
  const double* Rr = ref21 + 4; const double* tr = ref21 + 13;
  const double* Rc = init21 + 4; const double* tc = init21 + 13;
  double Rri[9], Rl[9], tl[3];
  inv3(Rr, Rri);
  mul33(Rc, Rri, Rl);                                   // R_curr_world * R_local_world^-1
  double nRc[9];
  for (int i = 0; i < 9; ++i) nRc[i] = -Rc[i];
  double tmp[9];
  mul33(nRc, Rri, tmp);
  mul31(tmp, tr, tl);
  for (int j = 0; j < 3; ++j) tl[j] += tc[j];           // -R_c R_r^-1 t_r + t_c
  local15[0] = init21[0]; local15[1] = init21[1]; local15[2] = init21[2]; local15[3] = init21[3];
  rodrigues_inv(Rl, local15 + 4);                       // ToVector (types.cc:39-44)
  local15[7] = tl[0]; local15[8] = tl[1]; local15[9] = tl[2];
  for (int j = 0; j < 5; ++j) local15[10 + j] = init21[16 + j];
}
// ObtainRefinedCameraParams (krt_optimizer.cc:535-567)
void krt_to_world(int type, const double* ref21, const double local15[15], double out21[21]) {
  double fx = local15[0], fy = (type == PTZ_KRT_F || type == PTZ_KRT_FDIST) ? local15[0] : local15[1];
  out21[0] = fx; out21[1] = fy; out21[2] = local15[2]; out21[3] = local15[3];
  double Rl[9];
  rodrigues(local15 + 4, Rl);
  mul33(Rl, ref21 + 4, out21 + 4);                      // R_local * R_local_world
  double t[3];
  mul31(Rl, ref21 + 13, t);
  for (int j = 0; j < 3; ++j) out21[13 + j] = t[j] + local15[7 + j];
  for (int j = 0; j < 5; ++j) out21[16 + j] = local15[10 + j];
}

thread_local char g_err[256] = "";

}  // namespace

// =====================================================================================================
// exported C functions (same structs as include/ptzcalib_b200.h, prefix orc_)
// =====================================================================================================
extern "C" {

void orc_solver_options_default(ptz_solver_options* o) {
  o->max_num_iterations = 50;  // Ceres default; the reference always overrides it
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->pcg_max_iterations = 2000;
  o->pcg_rel_tolerance = 1e-13;
  o->jacobian_mode = 0;
  o->linear_solver = 0;
  o->num_threads = 1;
  o->verbose = 0;
}
const char* orc_last_error(void) { return g_err; }

void orc_rodrigues(const double* r, double* R) { rodrigues(r, R); }
void orc_rodrigues_inv(const double* R, double* r) { rodrigues_inv(R, r); }
void orc_undistort_point(const float* uv, const double* K4, const double* dist5, float* out) { undistort_point_f32(uv, K4, dist5, out); }
void orc_inv3(const double* S, double* D) { inv3(S, D); }
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// raw functor calls (for golden-vector tests)
void orc_ba_ray_factor(int type, const double* intr, const double* ext, const double* ray, const double* disp, const float* uv, double* res) {
  double z[3] = {0, 0, 0};
  ba_ray_factor<double>(type, intr, ext, ray, disp ? disp : z, uv[0], uv[1], res);
}
void orc_ba_pt_factor(int type, const double* intr, const double* ext, const double* tlw, const double* disp, const float* uv, const double* xyz, double* res) {
  double z[3] = {0, 0, 0};
  ba_pt_factor<double>(type, intr, ext, tlw, disp ? disp : z, uv[0], uv[1], xyz, res);
}
void orc_krt_2d2d_factor(int type, const double* cam15, const double* refK4, const double* refd5, const float* uv1, const float* uv2, double* res) {
  KrtPre pre;
  krt_precompute(type, refK4, refd5, uv1, &pre);
  krt_2d2d_factor<double>(type, cam15, pre, uv2[0], uv2[1], res);
}
void orc_krt_2d3d_factor(int type, const double* cam15, const float* uv, const double* xyz, double* res) {
  krt_2d3d_factor<double>(type, cam15, xyz, uv[0], uv[1], res);
}
// exact / Ceres-numeric Jacobian of a 2d-3d KRT factor over all 15 coordinates, row-major [2][15]
void orc_krt_2d3d_jac(int type, const double* cam15, const float* uv, const double* xyz, double* J) {
  JetT a[15], r[2];
  for (int i = 0; i < 15; ++i) a[i] = JetT(cam15[i], i);
  krt_2d3d_factor<JetT>(type, a, xyz, uv[0], uv[1], r);
  for (int i = 0; i < 15; ++i) { J[i] = r[0].v[i]; J[15 + i] = r[1].v[i]; }
}

int orc_ptzba_eval_mode(const ptzba_problem* prob, const double* disp, ptzba_eval_out* out, int jacobian_mode) {
  int rc = ba_check(prob);
  if (rc != PTZ_OK) return rc;
  BaProblem bp(prob);
  std::vector<double> x;
  ba_initial_x(prob, bp, x, disp);
  const int M = bp.M, A = bp.A, ncv = bp.ncv;
  const int wo = ncv + 3 + (bp.use_disp ? 3 : 0), wp = ncv + 6 + (bp.use_disp ? 3 : 0);
  out->num_tangent = bp.num_tangent;
  if (out->gradient) std::fill(out->gradient, out->gradient + bp.num_tangent, 0.0);
  if (out->jac_obs) std::fill(out->jac_obs, out->jac_obs + (size_t)M * 2 * wo, 0.0);
  if (out->jac_pts) std::fill(out->jac_pts, out->jac_pts + (size_t)A * 2 * wp, 0.0);
  double cost = 0;
  for (int k = 0; k < M + A; ++k) {
    double res[2], Jg[2][kMaxCols];
    int na, amb[kMaxCols];
    eval_block_raw(bp, k, x.data(), jacobian_mode, res, &na, amb, Jg, true);
    double w = bp.block_weight(k);
    cost += 0.5 * w * (res[0] * res[0] + res[1] * res[1]);
    if (out->residuals) { out->residuals[2 * k] = res[0]; out->residuals[2 * k + 1] = res[1]; }
    // columns in the documented order: view cols, then ray|tlw, then disp
    int view = k < M ? prob->obs_view[k] : prob->pt_view[k - M];
    for (int i = 0; i < na; ++i) {
      int t = bp.amb2tan[amb[i]];
      if (t < 0) continue;
      if (out->gradient) out->gradient[t] += w * (Jg[0][i] * res[0] + Jg[1][i] * res[1]);
      int col;
      if (amb[i] < bp.off_ext) col = t - view * ncv;                       // intr free
      else if (amb[i] < bp.off_ray) col = t - view * ncv;                  // rvec
      else if (amb[i] < bp.off_disp) col = ncv + (amb[i] - bp.off_ray) % 3;   // ray
      else if (amb[i] < bp.off_tlw) col = ncv + (k < M ? 3 : 6) + (amb[i] - bp.off_disp);
      else col = ncv + (amb[i] - bp.off_tlw);
      if (k < M && out->jac_obs) { out->jac_obs[((size_t)k * 2) * wo + col] = Jg[0][i]; out->jac_obs[((size_t)k * 2 + 1) * wo + col] = Jg[1][i]; }
      if (k >= M && out->jac_pts) { out->jac_pts[((size_t)(k - M) * 2) * wp + col] = Jg[0][i]; out->jac_pts[((size_t)(k - M) * 2 + 1) * wp + col] = Jg[1][i]; }
    }
  }
  out->cost = cost;
  return PTZ_OK;
}
int orc_ptzba_eval(const ptzba_problem* prob, const double* disp, ptzba_eval_out* out) { return orc_ptzba_eval_mode(prob, disp, out, 0); }

// initial rays only (Pix2Ray)
int orc_ptzba_init_rays(const ptzba_problem* prob, double* ray) {
  int rc = ba_check(prob);
  if (rc != PTZ_OK) return rc;
  init_rays(prob, prob->intr, prob->ext, ray);
  return PTZ_OK;
}

int orc_ptzba_solve(const ptzba_problem* prob, const ptz_solver_options* opt, ptzba_result* out) {
  int rc = ba_check(prob);
  if (rc != PTZ_OK) return rc;
  if (!opt || !out || opt->max_num_iterations <= 0) return PTZ_ERR_INVALID;  // CheckValid: max_iter_ <= 0
  BaProblem bp(prob);
  std::vector<double> x;
  ba_initial_x(prob, bp, x, nullptr);
  LmSummary sum;
  minimize(bp, x, *opt, true, sum);
  out->termination = sum.termination;
  out->num_iterations = sum.num_iterations;
  out->num_successful_steps = sum.num_successful;
  out->num_unsuccessful_steps = sum.num_unsuccessful;
  out->num_residuals = 2 * (bp.M + bp.A);
  out->linear_solver_iterations = sum.lin_iters;
  out->initial_cost = sum.initial_cost;
  out->final_cost = sum.final_cost;
  out->init_reproj_error_all = std::sqrt(2.0) * std::sqrt((2 * sum.initial_cost) / out->num_residuals);
  out->final_reproj_error_all = std::sqrt(2.0) * std::sqrt((2 * sum.final_cost) / out->num_residuals);
  // CalReprojError2d2d / 2d3d: unweighted RMS at the final parameters
  double s2 = 0, s3 = 0;
  for (int k = 0; k < bp.M + bp.A; ++k) {
    double res[2], Jg[2][kMaxCols];
    int na, amb[kMaxCols];
    eval_block_raw(bp, k, x.data(), 0, res, &na, amb, Jg, false);
    (k < bp.M ? s2 : s3) += res[0] * res[0] + res[1] * res[1];
  }
  out->final_reproj_error_2d2d = std::sqrt(s2 / bp.M);
  out->final_reproj_error_2d3d = std::sqrt(s3 / bp.A);
  if (out->intr) for (int i = 0; i < bp.V; ++i) std::memcpy(out->intr + 9 * i, &x[bp.rep[i] * 9], sizeof(double) * 9);
  if (out->ext) std::memcpy(out->ext, &x[bp.off_ext], sizeof(double) * 6 * bp.V);
  if (out->ray) std::memcpy(out->ray, &x[bp.off_ray], sizeof(double) * 3 * bp.Pn);
  if (out->disp) std::memcpy(out->disp, &x[bp.off_disp], sizeof(double) * 3);
  if (out->tlw) std::memcpy(out->tlw, &x[bp.off_tlw], sizeof(double) * 6);
  ba_world_outputs(bp, x, out);
  out->log_count = 0;
  if (out->log)
    for (size_t i = 0; i < sum.log.size() && (int)i < out->log_capacity; ++i) out->log[out->log_count++] = sum.log[i];
  return PTZ_OK;
}

int orc_ptzreloc_eval_mode(int type, int N, const float* uv_ref, const float* uv_cur, const double* ref21, const double* local15, double* residuals,
                           double* jac, double* cost, double* gradient, int jacobian_mode) {
  if (N < 0 || type < 0 || type > 3) return PTZ_ERR_INVALID;
  KrtProblem kp(type, N, uv_ref, uv_cur, ref21, ref21 + 16);
  const int nf = kp.num_tangent;
  if (gradient) std::fill(gradient, gradient + nf, 0.0);
  double c = 0;
  for (int k = 0; k < N; ++k) {
    double res[2], Jg[2][kMaxCols];
    int na, amb[kMaxCols];
    eval_block_raw(kp, k, local15, jacobian_mode, res, &na, amb, Jg, true);
    c += 0.5 * (res[0] * res[0] + res[1] * res[1]);
    if (residuals) { residuals[2 * k] = res[0]; residuals[2 * k + 1] = res[1]; }
    for (int j = 0; j < nf; ++j) {
      int a = kp.tan2amb[j];
      if (jac) { jac[((size_t)k * 2) * nf + j] = Jg[0][a]; jac[((size_t)k * 2 + 1) * nf + j] = Jg[1][a]; }
      if (gradient) gradient[j] += Jg[0][a] * res[0] + Jg[1][a] * res[1];
    }
  }
  if (cost) *cost = c;
  return PTZ_OK;
}
int orc_ptzreloc_eval(int type, int N, const float* uv_ref, const float* uv_cur, const double* ref21, const double* local15, double* residuals, double* jac,
                      double* cost, double* gradient) {
  return orc_ptzreloc_eval_mode(type, N, uv_ref, uv_cur, ref21, local15, residuals, jac, cost, gradient, 0);
}
void orc_krt_to_local(const double* ref21, const double* init21, double* local15) { krt_to_local(ref21, init21, local15); }
// KRTOptimizer::Cal2d2dReprojError (krt_optimizer.cc:406-455) / Cal2d3dReprojError (:457-500) at local15
int orc_ptzreloc_reproj_error(int type, const double* ref21, const double* local15, int N, const float* uv_ref, const float* uv_cur, int npts,
                              const float* pt_uv, const double* pt_xyz, double* err_2d2d, double* err_2d3d) {
  double s22 = 0, s23 = 0;
  const double refK4[4] = {ref21[0], ref21[1], ref21[2], ref21[3]};
  for (int i = 0; i < N; ++i) {
    KrtPre pre;
    krt_precompute(type, refK4, ref21 + 16, uv_ref + 2 * i, &pre);
    double r[2];
    krt_2d2d_factor<double>(type, local15, pre, uv_cur[2 * i], uv_cur[2 * i + 1], r);
    s22 += r[0] * r[0] + r[1] * r[1];
  }
  for (int i = 0; i < npts; ++i) {
    double P[3], r[2];
    mul31(ref21 + 4, pt_xyz + 3 * i, P);
    for (int a = 0; a < 3; ++a) P[a] += ref21[13 + a];
    krt_2d3d_factor<double>(type, local15, P, pt_uv[2 * i], pt_uv[2 * i + 1], r);
    s23 += r[0] * r[0] + r[1] * r[1];
  }
  if (err_2d2d) *err_2d2d = std::sqrt(s22 / (double)N);
  if (err_2d3d) *err_2d3d = npts > 0 ? std::sqrt(s23 / (double)npts) : -1.0;
  return PTZ_OK;
}
void orc_krt_to_world(int type, const double* ref21, const double* local15, double* out21) { krt_to_world(type, ref21, local15, out21); }

int orc_ptzreloc_solve_batch(const ptzreloc_batch* b, const ptz_solver_options* opt, ptzreloc_result* out) {
  if (!b || !opt || !out || b->num_queries < 0 || b->factor_type < 0 || b->factor_type > 3) return PTZ_ERR_INVALID;
  const int B = b->num_queries;
  const int nthreads = opt->num_threads > 0 ? opt->num_threads : 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
#endif
  for (int q = 0; q < B; ++q) {
    const int64_t o0 = b->match_offset[q];
    const int N = (int)(b->match_offset[q + 1] - o0);
    const double* ref = b->ref_cam + 21 * (size_t)q;
    std::vector<double> x(15);
    krt_to_local(ref, b->init_cam + 21 * (size_t)q, x.data());
    KrtProblem kp(b->factor_type, N, b->uv_ref + 2 * o0, b->uv_cur + 2 * o0, ref, ref + 16);
    if (b->pt_offset && b->pt_uv && b->pt_xyz) {
      const int64_t p0 = b->pt_offset[q];
      kp.add_points((int)(b->pt_offset[q + 1] - p0), b->pt_uv + 2 * p0, b->pt_xyz + 3 * p0, ref);
    }
    ptz_solver_options o = *opt;
    o.max_num_iterations = b->max_iter;
    o.num_threads = 1;
    o.verbose = 0;
    LmSummary sum;
    minimize(kp, x, o, false, sum);
    // CheckResults (krt_optimizer.cc:504-533)
    const int nres = 2 * kp.num_blocks;
    double final_rms = std::sqrt(2.0) * std::sqrt((2 * sum.final_cost) / nres);
    int ok = 1;
    if (sum.termination != PTZ_CONVERGENCE) ok = 0;
    else if (final_rms >= b->max_reproj_error) ok = 0;
    if (ok) {
      double fx = x[0], fy = x[1], cx = x[2], cy = x[3];  // FromVector(cam_curr_local_param_): fy NOT tied here (:522-524)
      double fov_x = std::atan(cx / fx) * 2 * 180 / M_PI, fov_y = std::atan(cy / fy) * 2 * 180 / M_PI;
      if (fov_x < 0 || fov_x > 170 || fov_y < 0 || fov_y > 170) ok = 0;
    }
    out->success[q] = ok;
    out->termination[q] = sum.termination;
    out->num_iter[q] = sum.num_successful;
    if (out->iterations) out->iterations[q] = sum.num_iterations;
    if (out->initial_cost) out->initial_cost[q] = sum.initial_cost;
    if (out->final_cost) out->final_cost[q] = sum.final_cost;
    if (out->final_rms) out->final_rms[q] = final_rms;
    if (out->local_cam15) std::memcpy(out->local_cam15 + 15 * (size_t)q, x.data(), sizeof(double) * 15);
    krt_to_world(b->factor_type, ref, x.data(), out->cam + 21 * (size_t)q);
  }
  (void)nthreads;
  return PTZ_OK;
}

// single-query solve with the iteration table (trajectory tests)
int orc_ptzreloc_solve_one(int type, int N, const float* uv_ref, const float* uv_cur, const double* ref21, const double* init21, const ptz_solver_options* opt,
                           double* local15_out, ptz_iter_log* log, int log_capacity, int* log_count, int* termination) {
  std::vector<double> x(15);
  krt_to_local(ref21, init21, x.data());
  KrtProblem kp(type, N, uv_ref, uv_cur, ref21, ref21 + 16);
  LmSummary sum;
  minimize(kp, x, *opt, false, sum);
  std::memcpy(local15_out, x.data(), sizeof(double) * 15);
  int n = 0;
  if (log) for (size_t i = 0; i < sum.log.size() && (int)i < log_capacity; ++i) log[n++] = sum.log[i];
  if (log_count) *log_count = n;
  if (termination) *termination = sum.termination;
  return PTZ_OK;
}

// timing hook for bench.py: one Jacobian evaluation pass over all 2d-2d observations (mode 0 exact, 1 Ceres CENTRAL)
double orc_ptzba_time_jacobian(const ptzba_problem* prob, int jacobian_mode, int num_threads, int repeats) {
  if (ba_check(prob) != PTZ_OK) return -1;
  BaProblem bp(prob);
  std::vector<double> x;
  ba_initial_x(prob, bp, x, nullptr);
  std::vector<BlockJ> B;
  double acc = 0;
#ifdef _OPENMP
  double t0 = omp_get_wtime();
#endif
  for (int r = 0; r < repeats; ++r) acc += eval_all(bp, x.data(), jacobian_mode, &B, num_threads);
#ifdef _OPENMP
  double t1 = omp_get_wtime();
  return (t1 - t0) / repeats + 0 * acc;
#else
  return 0 * acc;
#endif
}

}  // extern "C"
