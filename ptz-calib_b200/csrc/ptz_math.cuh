// ptz_math.cuh — per-observation PTZ reprojection residuals and hand-derived analytic Jacobians (fp64).
//
// These replace ceres::NumericDiffCostFunction<..., CENTRAL, ...> around the reference's functors
// (src/core/ptzray_optimizer.cc:20-401, src/core/krt_optimizer.cc:22-197).  Functions are __host__ __device__
// so the closed forms can be unit-checked on the build host (tests/host_math_check.cpp); the product only ever
// calls them from kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PTZ_HD __host__ __device__ __forceinline__
#else
#define PTZ_HD inline
#endif

namespace ptz {

enum { BA_PTZRAY = 0, BA_PTZRAY_DIST = 1, BA_PTZRAY_FXFY_DIST = 2, BA_PTZRAY_DIST_DISP = 3 };
enum { KRT_F = 0, KRT_FDIST = 1, KRT_FXFY = 2, KRT_FXFYDIST = 3 };

// live (non-identically-zero) camera columns of a 2d-2d observation:
//   PTZRay: fx, w(3) | PTZRayDist, PTZRayDistDisp: fx, k1, w(3) | PTZRayFxfyDist: fx, fy, k1, w(3)
PTZ_HD constexpr int ba_ncl(int type) { return type == BA_PTZRAY ? 4 : (type == BA_PTZRAY_FXFY_DIST ? 6 : 5); }

// per-view table: everything a thread needs about its view, computed once per evaluation
struct ViewTab {
  double R[9];      // cv::Rodrigues(rvec)
  double Jl[9];     // left Jacobian of SO(3) at rvec, row-major: d(R n)/dw_k = (Jl e_k) x (R n)
  double fx, fy, cx, cy;
  double k1, k2, k3, p1, p2;  // hand-written factors read dist as (k1,k2,k3,p1,p2): ptzray_optimizer.cc:108-109
  double sc[6];     // Jacobi scales of the view's live camera columns (k_resjac takes the whole table in ONE bulk copy)
  double pad[7];
};
constexpr int kViewTabDoubles = 40;
static_assert(sizeof(ViewTab) == kViewTabDoubles * 8 && sizeof(ViewTab) % 16 == 0, "ViewTab is 40 doubles");

// R(w) = I + a[w]x + b[w]x^2 and its derivatives.  a = sin t/t, b = (1-cos t)/t^2; da = a'(t)/t, db = b'(t)/t.
// Series below t^2 = 1e-2 (truncation < 3e-18), closed forms above.  Same function as cv::Rodrigues; OpenCV
// returns exactly I below t = DBL_EPSILON, which this also does to rounding.
PTZ_HD void rodrigues_jac(const double w[3], double R[9], double dR[27]) {
  const double x = w[0], y = w[1], z = w[2];
  const double t2 = x * x + y * y + z * z;
  double a, b, da, db;
  if (t2 < 1e-2) {
    a = 1.0 + t2 * (-1.0 / 6 + t2 * (1.0 / 120 + t2 * (-1.0 / 5040 + t2 * (1.0 / 362880))));
    b = 0.5 + t2 * (-1.0 / 24 + t2 * (1.0 / 720 + t2 * (-1.0 / 40320 + t2 * (1.0 / 3628800))));
    da = -1.0 / 3 + t2 * (1.0 / 30 + t2 * (-1.0 / 840 + t2 * (1.0 / 45360 + t2 * (-1.0 / 3991680))));
    db = -1.0 / 12 + t2 * (1.0 / 180 + t2 * (-1.0 / 6720 + t2 * (1.0 / 453600 + t2 * (-1.0 / 47900160))));
  } else {
    const double t = sqrt(t2);
    const double s = sin(t), c = cos(t);
    a = s / t;
    b = (1.0 - c) / t2;
    da = (c - a) / t2;
    db = (a - 2.0 * b) / t2;
  }
  // K = [w]x ; K2 = w w^T - t2 I
  const double K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  const double K2[9] = {x * x - t2, x * y, x * z, x * y, y * y - t2, y * z, x * z, y * z, z * z - t2};
  for (int i = 0; i < 9; ++i) R[i] = a * K[i] + b * K2[i];
  R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
  if (dR) {
    for (int k = 0; k < 3; ++k) {
      const double wk = w[k];
      double* D = dR + 9 * k;
      for (int i = 0; i < 9; ++i) D[i] = (da * wk) * K[i] + (db * wk) * K2[i];
      // a * G_k, G_k = [e_k]x
      if (k == 0) { D[5] -= a; D[7] += a; }
      if (k == 1) { D[2] += a; D[6] -= a; }
      if (k == 2) { D[1] -= a; D[3] += a; }
      // b * (w e_k^T + e_k w^T - 2 w_k I)
      for (int i = 0; i < 3; ++i) { D[3 * i + k] += b * w[i]; D[3 * k + i] += b * w[i]; }
      D[0] -= 2.0 * b * wk; D[4] -= 2.0 * b * wk; D[8] -= 2.0 * b * wk;
    }
  }
}

// R(w) and the left Jacobian J_l(w) = I + b[w]x + c[w]x^2, b = (1-cos t)/t^2, c = (t - sin t)/t^3.  For any vector n,
// d(R(w) n)/dw_k = (J_l e_k) x (R(w) n): the rotation-vector derivative without the 27-entry dR/dw table.
PTZ_HD void rodrigues_jl(const double w[3], double R[9], double Jl[9]) {
  const double x = w[0], y = w[1], z = w[2];
  const double t2 = x * x + y * y + z * z;
  double a, b, c;
  if (t2 < 1e-2) {
    a = 1.0 + t2 * (-1.0 / 6 + t2 * (1.0 / 120 + t2 * (-1.0 / 5040 + t2 * (1.0 / 362880))));
    b = 0.5 + t2 * (-1.0 / 24 + t2 * (1.0 / 720 + t2 * (-1.0 / 40320 + t2 * (1.0 / 3628800))));
    c = 1.0 / 6 + t2 * (-1.0 / 120 + t2 * (1.0 / 5040 + t2 * (-1.0 / 362880 + t2 * (1.0 / 39916800))));
  } else {
    const double t = sqrt(t2);
    const double s = sin(t), co = cos(t);
    a = s / t;
    b = (1.0 - co) / t2;
    c = (1.0 - a) / t2;
  }
  const double K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  const double K2[9] = {x * x - t2, x * y, x * z, x * y, y * y - t2, y * z, x * z, y * z, z * z - t2};
  for (int i = 0; i < 9; ++i) { R[i] = a * K[i] + b * K2[i]; if (Jl) Jl[i] = b * K[i] + c * K2[i]; }
  R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
  if (Jl) { Jl[0] += 1.0; Jl[4] += 1.0; Jl[8] += 1.0; }
}
// derivative of X = R n w.r.t. w_k from the k-th column of Jl:  (Jl e_k) x X
PTZ_HD void dX_dw(const double Jl[9], int k, double X, double Y, double Z, double& dX, double& dY, double& dZ) {
  const double j0 = Jl[k], j1 = Jl[3 + k], j2 = Jl[6 + k];
  dX = j1 * Z - j2 * Y;
  dY = j2 * X - j0 * Z;
  dZ = j0 * Y - j1 * X;
}

PTZ_HD void make_view_tab(const double intr[9], const double ext[6], ViewTab* vt, bool with_jac) {
  rodrigues_jl(ext, vt->R, vt->Jl);
  (void)with_jac;
  vt->fx = intr[0]; vt->fy = intr[1]; vt->cx = intr[2]; vt->cy = intr[3];
  vt->k1 = intr[4]; vt->k2 = intr[5]; vt->k3 = intr[6]; vt->p1 = intr[7]; vt->p2 = intr[8];
  for (int i = 0; i < 6; ++i) vt->sc[i] = 1.0;
  for (int i = 0; i < 7; ++i) vt->pad[i] = 0.0;
}

// Brown model of the hand-written factors and its 2x2 Jacobian
struct Brown {
  double xd, yd, xd_x, xd_y, yd_x, yd_y, r2;
};
PTZ_HD Brown brown(double x, double y, double k1, double k2, double k3, double p1, double p2, bool jac) {
  Brown o;
  const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r2 * r2 * r2;
  const double rad = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
  o.r2 = r2;
  o.xd = x * rad + 2.0 * p1 * (x * y) + p2 * (r2 + 2.0 * (x * x));
  o.yd = y * rad + 2.0 * p2 * (x * y) + p1 * (r2 + 2.0 * (y * y));
  if (jac) {
    const double g = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;  // d rad / d r2
    o.xd_x = rad + 2.0 * x * x * g + 2.0 * p1 * y + 6.0 * p2 * x;
    o.xd_y = 2.0 * x * y * g + 2.0 * p1 * x + 2.0 * p2 * y;
    o.yd_x = 2.0 * x * y * g + 2.0 * p2 * y + 2.0 * p1 * x;
    o.yd_y = rad + 2.0 * y * y * g + 2.0 * p2 * x + 6.0 * p1 * y;
  }
  return o;
}

// -----------------------------------------------------------------------------------------------------------
// 2d-2d observation of a PTZ-ray factor.  Outputs the raw residual r[2] and (when JAC) the raw Jacobian blocks
//   F[2][NCL]  columns as listed at ba_ncl()
//   E[2][3]    d r / d ray
//   Fd[2][3]   d r / d disp  (PTZRayDistDisp only)
// Reference: PTZRayFactor (:20-56), PTZRayDistFactor (:65-129), PTZRayFxfyDistFactor (:138-193),
// PTZRayDistDispFactor (:202-259).
// -----------------------------------------------------------------------------------------------------------
template <int TYPE, bool JAC>
PTZ_HD void ba_obs(const ViewTab& vt, const double ray[3], const double disp[3], double u, double v, double r[2], double* F /*[2*NCL]*/,
                   double* E /*[6]*/, double* Fd /*[6]*/) {
  constexpr int NCL = ba_ncl(TYPE);
  const double* R = vt.R;
  double n[3] = {ray[0], ray[1], ray[2]};
  double inv_norm = 1.0;
  if (TYPE != BA_PTZRAY_DIST) {
    inv_norm = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] *= inv_norm; n[1] *= inv_norm; n[2] *= inv_norm;
  }
  const double X = R[0] * n[0] + R[1] * n[1] + R[2] * n[2];
  const double Y = R[3] * n[0] + R[4] * n[1] + R[5] * n[2];
  double Z = R[6] * n[0] + R[7] * n[1] + R[8] * n[2];
  if (TYPE == BA_PTZRAY_DIST && Z < 0) {  // constant penalty, zero derivative (:97-102)
    r[0] = 1000000.0; r[1] = 1000000.0;
    if (JAC) {
      for (int i = 0; i < 2 * NCL; ++i) F[i] = 0.0;
      for (int i = 0; i < 6; ++i) E[i] = 0.0;
    }
    return;
  }
  const double fx = vt.fx, fy = (TYPE == BA_PTZRAY_FXFY_DIST) ? vt.fy : vt.fx;
  const double Zr = Z;  // before displacement
  double dZ_df = 0.0;
  if (TYPE == BA_PTZRAY_DIST_DISP) {
    Z = Z + (disp[0] + disp[1] * fx + disp[2] * fx * fx);
    dZ_df = disp[1] + 2.0 * disp[2] * fx;
  }
  const double iz = 1.0 / Z;
  const double x = X * iz, y = Y * iz;
  Brown bw;
  if (TYPE == BA_PTZRAY) {
    bw.xd = x; bw.yd = y; bw.xd_x = 1.0; bw.xd_y = 0.0; bw.yd_x = 0.0; bw.yd_y = 1.0; bw.r2 = 0.0;
  } else {
    bw = brown(x, y, vt.k1, vt.k2, vt.k3, vt.p1, vt.p2, JAC);
  }
  r[0] = u - (fx * bw.xd + vt.cx);
  r[1] = v - (fy * bw.yd + vt.cy);
  if (!JAC) return;
  // d(u,v)/d(X,Y,Z)
  const double ux = fx * bw.xd_x * iz, uy = fx * bw.xd_y * iz, uz = -(ux * x + uy * y);
  const double vx = fy * bw.yd_x * iz, vy = fy * bw.yd_y * iz, vz = -(vx * x + vy * y);
  int c = 0;
  // fx (and the tied fy): d u/d fx = xd (+ disp path through Z)
  F[c] = -(bw.xd + uz * dZ_df);
  F[NCL + c] = (TYPE == BA_PTZRAY_FXFY_DIST) ? -(vz * dZ_df) : -(bw.yd + vz * dZ_df);
  ++c;
  if (TYPE == BA_PTZRAY_FXFY_DIST) { F[c] = 0.0; F[NCL + c] = -bw.yd; ++c; }
  if (TYPE != BA_PTZRAY) { F[c] = -(fx * x * bw.r2); F[NCL + c] = -(fy * y * bw.r2); ++c; }
  // rotation: dX/dw_k = (Jl e_k) x (R n)   (X, Y, Zr: before any displacement)
  for (int k = 0; k < 3; ++k) {
    double dX, dY, dZ;
    dX_dw(vt.Jl, k, X, Y, Zr, dX, dY, dZ);
    F[c + k] = -(ux * dX + uy * dY + uz * dZ);
    F[NCL + c + k] = -(vx * dX + vy * dY + vz * dZ);
  }
  // ray: dX/d rho = (R - X_r n^T) / |rho| when normalised (X_r = R n before displacement), R otherwise
  for (int j = 0; j < 3; ++j) {
    double dX, dY, dZ;
    if (TYPE != BA_PTZRAY_DIST) {
      dX = (R[j] - X * n[j]) * inv_norm;
      dY = (R[3 + j] - Y * n[j]) * inv_norm;
      dZ = (R[6 + j] - Zr * n[j]) * inv_norm;
    } else {
      dX = R[j]; dY = R[3 + j]; dZ = R[6 + j];
    }
    E[j] = -(ux * dX + uy * dY + uz * dZ);
    E[3 + j] = -(vx * dX + vy * dY + vz * dZ);
  }
  if (TYPE == BA_PTZRAY_DIST_DISP && Fd) {
    Fd[0] = -uz; Fd[1] = -uz * fx; Fd[2] = -uz * fx * fx;
    Fd[3] = -vz; Fd[4] = -vz * fx; Fd[5] = -vz * fx * fx;
  }
}

// -----------------------------------------------------------------------------------------------------------
// 2d-3d annotated point (Reproj2d3dFactor :268-326, Reproj2d3dDispFactor :335-396): fx and fy both live.
//   Jc[2][6]  columns fx, fy, k1, w(3)     (k1 column is meaningful for the distortion types only)
//   Jt[2][6]  d r / d tlw (rvec 3, t 3)
//   Jd[2][3]  d r / d disp (DISP only)
// -----------------------------------------------------------------------------------------------------------
template <bool DISP, bool JAC>
PTZ_HD void ba_pt(const ViewTab& vt, const double Rl[9], const double Jll[9] /* left Jacobian at tlw rvec */, const double tl[3], const double disp[3],
                  const double Xw[3], double u, double v, double r[2], double* Jc, double* Jt, double* Jd) {
  const double* R = vt.R;
  double Xl[3];
  for (int i = 0; i < 3; ++i) Xl[i] = Rl[3 * i] * Xw[0] + Rl[3 * i + 1] * Xw[1] + Rl[3 * i + 2] * Xw[2] + tl[i];
  const double X = R[0] * Xl[0] + R[1] * Xl[1] + R[2] * Xl[2];
  const double Y = R[3] * Xl[0] + R[4] * Xl[1] + R[5] * Xl[2];
  double Z = R[6] * Xl[0] + R[7] * Xl[1] + R[8] * Xl[2];
  const double Zr = Z;
  const double fx = vt.fx, fy = vt.fy;
  double dZ_df = 0.0;
  if (DISP) { Z = Z + (disp[0] + disp[1] * fx + disp[2] * fx * fx); dZ_df = disp[1] + 2.0 * disp[2] * fx; }
  const double iz = 1.0 / Z, x = X * iz, y = Y * iz;
  const Brown bw = brown(x, y, vt.k1, vt.k2, vt.k3, vt.p1, vt.p2, JAC);
  r[0] = u - (fx * bw.xd + vt.cx);
  r[1] = v - (fy * bw.yd + vt.cy);
  if (!JAC) return;
  const double ux = fx * bw.xd_x * iz, uy = fx * bw.xd_y * iz, uz = -(ux * x + uy * y);
  const double vx = fy * bw.yd_x * iz, vy = fy * bw.yd_y * iz, vz = -(vx * x + vy * y);
  Jc[0] = -(bw.xd + uz * dZ_df); Jc[6 + 0] = -(vz * dZ_df);
  Jc[1] = 0.0;                   Jc[6 + 1] = -bw.yd;
  Jc[2] = -(fx * x * bw.r2);     Jc[6 + 2] = -(fy * y * bw.r2);
  for (int k = 0; k < 3; ++k) {
    double dX, dY, dZ;
    dX_dw(vt.Jl, k, X, Y, Zr, dX, dY, dZ);
    Jc[3 + k] = -(ux * dX + uy * dY + uz * dZ);
    Jc[6 + 3 + k] = -(vx * dX + vy * dY + vz * dZ);
  }
  const double Rw[3] = {Xl[0] - tl[0], Xl[1] - tl[1], Xl[2] - tl[2]};  // R_lw X_w
  for (int k = 0; k < 3; ++k) {  // tlw rotation: dXl = (Jll e_k) x (R_lw X_w), dXcam = R dXl
    double d[3];
    dX_dw(Jll, k, Rw[0], Rw[1], Rw[2], d[0], d[1], d[2]);
    const double dX = R[0] * d[0] + R[1] * d[1] + R[2] * d[2];
    const double dY = R[3] * d[0] + R[4] * d[1] + R[5] * d[2];
    const double dZ = R[6] * d[0] + R[7] * d[1] + R[8] * d[2];
    Jt[k] = -(ux * dX + uy * dY + uz * dZ);
    Jt[6 + k] = -(vx * dX + vy * dY + vz * dZ);
  }
  for (int j = 0; j < 3; ++j) {  // tlw translation: dXcam = R[:, j]
    Jt[3 + j] = -(ux * R[j] + uy * R[3 + j] + uz * R[6 + j]);
    Jt[6 + 3 + j] = -(vx * R[j] + vy * R[3 + j] + vz * R[6 + j]);
  }
  if (DISP && Jd) {
    Jd[0] = -uz; Jd[1] = -uz * fx; Jd[2] = -uz * fx * fx;
    Jd[3] = -vz; Jd[4] = -vz * fx; Jd[5] = -vz * fx * fx;
  }
}

// -----------------------------------------------------------------------------------------------------------
// KRT 2d-2d factors (krt_optimizer.cc:22-197) with the parameter-independent part (ray1, border mask) hoisted.
// Free columns in ascending parameter index: F: fx,w | Fxfy: fx,fy,w | FDist: fx,w,k1 | FxfyDist: fx,fy,w,k1
// -----------------------------------------------------------------------------------------------------------
PTZ_HD constexpr int krt_nfree(int type) { return type == KRT_F ? 4 : (type == KRT_FXFYDIST ? 6 : 5); }

// cv::invert 3x3 of K = [fx 0 cx; 0 fy cy; 0 0 1] applied to (u, v, 1): cofactors times 1/det, as OpenCV does
PTZ_HD void kinv_apply(double fx, double fy, double cx, double cy, double u, double v, double out[3]) {
  const double d = 1.0 / (fx * fy);
  const double t00 = fy * d, t02 = (0.0 - cx * fy) * d, t11 = fx * d, t12 = (0.0 - fx * cy) * d, t22 = (fx * fy) * d;
  out[0] = t00 * u + 0.0 * v + t02;
  out[1] = 0.0 * u + t11 * v + t12;
  out[2] = t22;
}

// cv::undistortPoints(src, dst, K, dist, noArray(), K): 5 fixed-point iterations, OpenCV coefficient order
// (k1,k2,p1,p2,k3) on the stored vector, float32 result (krt_optimizer.cc:89-92)
PTZ_HD void undistort_f32(float u, float v, double fx, double fy, double cx, double cy, const double d[5], float out[2]) {
  const double ifx = 1.0 / fx, ify = 1.0 / fy;
  double x = ((double)u - cx) * ifx, y = ((double)v - cy) * ify;
  const double x0 = x, y0 = y;
  const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
  if (k1 != 0 || k2 != 0 || p1 != 0 || p2 != 0 || k3 != 0) {
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
      if (icdist < 0) { x = x0; y = y0; break; }
      const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
      const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
  }
  out[0] = (float)(x * fx + cx);
  out[1] = (float)(y * fy + cy);
}

// returns false when the match is masked out (residual identically 0)
PTZ_HD bool krt_precompute(int type, const double refK4[4], const double refd[5], float u1, float v1, double ray1[3]) {
  double pu = u1, pv = v1;
  if (type == KRT_FDIST || type == KRT_FXFYDIST) {
    float und[2];
    undistort_f32(u1, v1, refK4[0], refK4[1], refK4[2], refK4[3], refd, und);
    const double w1 = refK4[2] * 2, h1 = refK4[3] * 2;
    if (und[0] < 0 || und[0] >= w1 || und[1] < 0 || und[1] >= h1) { ray1[0] = ray1[1] = 0; ray1[2] = 1; return false; }
    pu = und[0]; pv = und[1];
  }
  kinv_apply(refK4[0], refK4[1], refK4[2], refK4[3], pu, pv, ray1);
  if (type != KRT_FXFY) {
    const double inn = 1.0 / sqrt(ray1[0] * ray1[0] + ray1[1] * ray1[1] + ray1[2] * ray1[2]);
    ray1[0] *= inn; ray1[1] *= inn; ray1[2] *= inn;
  }
  return true;
}

// camera-side constants of one KRT evaluation
struct KrtCam {
  double R[9], Jl[9];
  double fx, fy, cx, cy, k1, k2, k3, p1, p2;
};
template <int TYPE>
PTZ_HD void krt_make_cam(const double cam[15], KrtCam* kc, bool with_jac) {
  rodrigues_jl(cam + 4, kc->R, with_jac ? kc->Jl : nullptr);
  kc->fx = cam[0];
  kc->fy = (TYPE == KRT_FXFY || TYPE == KRT_FXFYDIST) ? cam[1] : cam[0];
  kc->cx = cam[2]; kc->cy = cam[3];
  kc->k1 = cam[10]; kc->k2 = cam[11]; kc->k3 = cam[12]; kc->p1 = cam[13]; kc->p2 = cam[14];
}
template <int TYPE, bool JAC>
PTZ_HD void krt_obs(const KrtCam& kc, const double n[3], double u2, double v2, double r[2], double* J /*[2*NFREE]*/) {
  constexpr int NF = krt_nfree(TYPE);
  constexpr bool DIST = (TYPE == KRT_FDIST || TYPE == KRT_FXFYDIST);
  constexpr bool FXFY = (TYPE == KRT_FXFY || TYPE == KRT_FXFYDIST);
  const double* R = kc.R;
  const double X = R[0] * n[0] + R[1] * n[1] + R[2] * n[2];
  const double Y = R[3] * n[0] + R[4] * n[1] + R[5] * n[2];
  const double Z = R[6] * n[0] + R[7] * n[1] + R[8] * n[2];
  const double iz = 1.0 / Z, x = X * iz, y = Y * iz;
  Brown bw;
  if (DIST) bw = brown(x, y, kc.k1, kc.k2, kc.k3, kc.p1, kc.p2, JAC);
  else { bw.xd = x; bw.yd = y; bw.xd_x = 1.0; bw.xd_y = 0.0; bw.yd_x = 0.0; bw.yd_y = 1.0; bw.r2 = 0.0; }
  r[0] = u2 - (kc.fx * bw.xd + kc.cx);
  r[1] = v2 - (kc.fy * bw.yd + kc.cy);
  if (!JAC) return;
  const double ux = kc.fx * bw.xd_x * iz, uy = kc.fx * bw.xd_y * iz, uz = -(ux * x + uy * y);
  const double vx = kc.fy * bw.yd_x * iz, vy = kc.fy * bw.yd_y * iz, vz = -(vx * x + vy * y);
  int c = 0;
  J[c] = -bw.xd; J[NF + c] = FXFY ? 0.0 : -bw.yd; ++c;
  if (FXFY) { J[c] = 0.0; J[NF + c] = -bw.yd; ++c; }
  for (int k = 0; k < 3; ++k) {
    double dX, dY, dZ;
    dX_dw(kc.Jl, k, X, Y, Z, dX, dY, dZ);
    J[c + k] = -(ux * dX + uy * dY + uz * dZ);
    J[NF + c + k] = -(vx * dX + vy * dY + vz * dZ);
  }
  c += 3;
  if (DIST) { J[c] = -(kc.fx * x * bw.r2); J[NF + c] = -(kc.fy * y * bw.r2); }
}

// Factor2d3dDist / Factor2d3dFxfyDist (krt_optimizer.cc:201-248): cv::projectPoints(P; rvec, t, K, dist) with the stored
// distortion vector read in OpenCV order (k1,k2,p1,p2,k3) = v[10..14] and the camera translation t = v[7..9] (not free).
// Free columns as krt_obs.  P is the point in the reference-local frame.
template <int TYPE, bool JAC>
PTZ_HD void krt_obs3d(const KrtCam& kc, const double t[3], const double P[3], double u, double v, double r[2], double* J /*[2*NFREE]*/) {
  constexpr int NF = krt_nfree(TYPE);
  constexpr bool DIST = (TYPE == KRT_FDIST || TYPE == KRT_FXFYDIST);
  constexpr bool FXFY = (TYPE == KRT_FXFY || TYPE == KRT_FXFYDIST);
  const double* R = kc.R;
  const double RX = R[0] * P[0] + R[1] * P[1] + R[2] * P[2];
  const double RY = R[3] * P[0] + R[4] * P[1] + R[5] * P[2];
  const double RZ = R[6] * P[0] + R[7] * P[1] + R[8] * P[2];
  const double X = RX + t[0], Y = RY + t[1], Z = RZ + t[2];
  const double iz = 1.0 / Z, x = X * iz, y = Y * iz;
  // hand order (k1,k2,k3,p1,p2) <- OpenCV order (v10,v11,v14,v12,v13); kc holds v10..v14 as k1,k2,k3,p1,p2
  const Brown bw = brown(x, y, kc.k1, kc.k2, kc.p2, kc.k3, kc.p1, JAC);
  r[0] = u - (kc.fx * bw.xd + kc.cx);
  r[1] = v - (kc.fy * bw.yd + kc.cy);
  if (!JAC) return;
  const double ux = kc.fx * bw.xd_x * iz, uy = kc.fx * bw.xd_y * iz, uz = -(ux * x + uy * y);
  const double vx = kc.fy * bw.yd_x * iz, vy = kc.fy * bw.yd_y * iz, vz = -(vx * x + vy * y);
  int c = 0;
  J[c] = -bw.xd; J[NF + c] = FXFY ? 0.0 : -bw.yd; ++c;
  if (FXFY) { J[c] = 0.0; J[NF + c] = -bw.yd; ++c; }
  for (int k = 0; k < 3; ++k) {
    double dX, dY, dZ;
    dX_dw(kc.Jl, k, RX, RY, RZ, dX, dY, dZ);  // d(R P)/dw_k; t does not depend on w
    J[c + k] = -(ux * dX + uy * dY + uz * dZ);
    J[NF + c + k] = -(vx * dX + vy * dY + vz * dZ);
  }
  c += 3;
  if (DIST) { J[c] = -(kc.fx * x * bw.r2); J[NF + c] = -(kc.fy * y * bw.r2); }
}

// -----------------------------------------------------------------------------------------------------------
// frame changes of KRTOptimizer (krt_optimizer.cc:269-286, 535-567) with OpenCV's 3x3 inverse and cv::Rodrigues(R->r)
// -----------------------------------------------------------------------------------------------------------
PTZ_HD void inv3(const double S[9], double D[9]) {  // cv::invert, cofactor formula
  double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (d == 0) { for (int i = 0; i < 9; ++i) D[i] = 0; return; }
  d = 1.0 / d;
  double t[9];
  t[0] = (S[4] * S[8] - S[5] * S[7]) * d; t[1] = (S[2] * S[7] - S[1] * S[8]) * d; t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  t[3] = (S[5] * S[6] - S[3] * S[8]) * d; t[4] = (S[0] * S[8] - S[2] * S[6]) * d; t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  t[6] = (S[3] * S[7] - S[4] * S[6]) * d; t[7] = (S[1] * S[6] - S[0] * S[7]) * d; t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  for (int i = 0; i < 9; ++i) D[i] = t[i];
}
PTZ_HD void mul33(const double A[9], const double B[9], double C[9]) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  for (int i = 0; i < 9; ++i) C[i] = t[i];
}
// cv::Rodrigues matrix -> vector: nearest rotation first (OpenCV: U*Vt of the SVD; here the equivalent polar Newton
// iteration), then the axis-angle extraction with OpenCV's small-angle and near-pi branches.
PTZ_HD void rodrigues_inv(const double Rin[9], double r[3]) {
  double R[9];
  for (int i = 0; i < 9; ++i) R[i] = Rin[i];
  for (int it = 0; it < 20; ++it) {
    double Ri[9];
    inv3(R, Ri);
    double diff = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double nv = 0.5 * (R[3 * i + j] + Ri[3 * j + i]);
        diff = fmax(diff, fabs(nv - R[3 * i + j]));
        R[3 * i + j] = nv;
      }
    if (diff < 1e-16) break;
  }
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = acos(c);
  if (s < 1e-5) {
    if (c > 0) { rx = ry = rz = 0; }
    else {
      double t;
      t = (R[0] + 1) * 0.5; rx = sqrt(fmax(t, 0.));
      t = (R[4] + 1) * 0.5; ry = sqrt(fmax(t, 0.)) * (R[1] < 0 ? -1. : 1.);
      t = (R[8] + 1) * 0.5; rz = sqrt(fmax(t, 0.)) * (R[2] < 0 ? -1. : 1.);
      if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
      theta /= sqrt(rx * rx + ry * ry + rz * rz);
      rx *= theta; ry *= theta; rz *= theta;
    }
  } else {
    double vth = 1 / (2 * s);
    vth *= theta;
    rx *= vth; ry *= vth; rz *= vth;
  }
  r[0] = rx; r[1] = ry; r[2] = rz;
}
// world -> reference-local: R_loc = R_cur R_ref^-1, t_loc = -R_cur R_ref^-1 t_ref + t_cur, then Camera::ToVector
PTZ_HD void krt_to_local(const double* ref21, const double* init21, double local15[15]) {
  const double* Rr = ref21 + 4; const double* tr = ref21 + 13;
  const double* Rc = init21 + 4; const double* tc = init21 + 13;
  double Rri[9], Rl[9], nRc[9], tmp[9];
  inv3(Rr, Rri);
  mul33(Rc, Rri, Rl);
  for (int i = 0; i < 9; ++i) nRc[i] = -Rc[i];
  mul33(nRc, Rri, tmp);
  local15[0] = init21[0]; local15[1] = init21[1]; local15[2] = init21[2]; local15[3] = init21[3];
  rodrigues_inv(Rl, local15 + 4);
  for (int i = 0; i < 3; ++i) local15[7 + i] = (tmp[3 * i] * tr[0] + tmp[3 * i + 1] * tr[1] + tmp[3 * i + 2] * tr[2]) + tc[i];
  for (int j = 0; j < 5; ++j) local15[10 + j] = init21[16 + j];
}
// reference-local -> world: R = R_loc R_ref, t = R_loc t_ref + t_loc; fy := fx for the tied types
PTZ_HD void krt_to_world(int type, const double* ref21, const double local15[15], double out21[21]) {
  out21[0] = local15[0];
  out21[1] = (type == KRT_F || type == KRT_FDIST) ? local15[0] : local15[1];
  out21[2] = local15[2]; out21[3] = local15[3];
  double Rl[9];
  rodrigues_jac(local15 + 4, Rl, nullptr);
  mul33(Rl, ref21 + 4, out21 + 4);
  for (int i = 0; i < 3; ++i) out21[13 + i] = (Rl[3 * i] * ref21[13] + Rl[3 * i + 1] * ref21[14] + Rl[3 * i + 2] * ref21[15]) + local15[7 + i];
  for (int j = 0; j < 5; ++j) out21[16 + j] = local15[10 + j];
}

// -----------------------------------------------------------------------------------------------------------
// small dense helpers
// -----------------------------------------------------------------------------------------------------------
// Cholesky of a 3x3 SPD matrix given as lower triangle [a00 a10 a11 a20 a21 a22]; false when not positive
PTZ_HD bool chol3(const double A[6], double L[6]) {
  double l00 = A[0];
  if (!(l00 > 0)) return false;
  l00 = sqrt(l00);
  const double l10 = A[1] / l00;
  double l11 = A[2] - l10 * l10;
  if (!(l11 > 0)) return false;
  l11 = sqrt(l11);
  const double l20 = A[3] / l00;
  const double l21 = (A[4] - l20 * l10) / l11;
  double l22 = A[5] - l20 * l20 - l21 * l21;
  if (!(l22 > 0)) return false;
  l22 = sqrt(l22);
  L[0] = l00; L[1] = l10; L[2] = l11; L[3] = l20; L[4] = l21; L[5] = l22;
  return true;
}
// in-place Cholesky of an n x n SPD matrix (row-major, lower used), n <= a few tens; false when not positive
PTZ_HD bool chol_n(double* A, int n, int ld) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * ld + j];
    for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
    if (!(d > 0)) return false;
    d = sqrt(d);
    A[j * ld + j] = d;
    const double id = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * ld + j];
      for (int k = 0; k < j; ++k) s -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = s * id;
    }
  }
  return true;
}
PTZ_HD void chol_solve_n(const double* L, int n, int ld, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * ld + k] * b[k];
    b[i] = s / L[i * ld + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * ld + i] * b[k];
    b[i] = s / L[i * ld + i];
  }
}

}  // namespace ptz
