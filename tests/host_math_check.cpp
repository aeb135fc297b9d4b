// Host-side compilation of the kernels' per-observation math (ptz-calib_b200/csrc/ptz_math.cuh) so the hand-derived
// Jacobians can be checked against the oracle without a GPU.  Test harness only: never part of the product library.
#include "../ptz-calib_b200/csrc/ptz_math.cuh"
using namespace ptz;
template <int T> static void ba_obs_t(const double* intr, const double* ext, const double* ray, const double* disp, const float* uv, double* r, double* J) {
  constexpr int NCL = ba_ncl(T);
  ViewTab vt; make_view_tab(intr, ext, &vt, true);
  double F[2 * NCL], E[6], Fd[6] = {0, 0, 0, 0, 0, 0};
  ba_obs<T, true>(vt, ray, disp, uv[0], uv[1], r, F, E, Fd);
  // eval layout: [fx, fy, (k1), w(3), ray(3), (disp 3)]
  const int ncv = (T == BA_PTZRAY) ? 5 : 6, wo = ncv + 3 + (T == BA_PTZRAY_DIST_DISP ? 3 : 0);
  for (int row = 0; row < 2; ++row) {
    double* o = J + row * wo; const double* f = F + row * NCL;
    int c = 0;
    o[0] = f[c++];
    o[1] = (T == BA_PTZRAY_FXFY_DIST) ? f[c++] : 0.0;
    if (T != BA_PTZRAY) o[2] = f[c++];
    for (int k = 0; k < 3; ++k) o[ncv - 3 + k] = f[c + k];
    for (int k = 0; k < 3; ++k) o[ncv + k] = E[row * 3 + k];
    if (T == BA_PTZRAY_DIST_DISP) for (int k = 0; k < 3; ++k) o[ncv + 3 + k] = Fd[row * 3 + k];
  }
}
extern "C" {
void hm_ba_obs(int type, const double* intr, const double* ext, const double* ray, const double* disp, const float* uv, double* r, double* J) {
  switch (type) {
    case 0: ba_obs_t<0>(intr, ext, ray, disp, uv, r, J); break;
    case 1: ba_obs_t<1>(intr, ext, ray, disp, uv, r, J); break;
    case 2: ba_obs_t<2>(intr, ext, ray, disp, uv, r, J); break;
    default: ba_obs_t<3>(intr, ext, ray, disp, uv, r, J); break;
  }
}
void hm_ba_pt(int disp_on, const double* intr, const double* ext, const double* tlw, const double* disp, const float* uv, const double* xyz, double* r,
              double* Jc, double* Jt, double* Jd) {
  ViewTab vt; make_view_tab(intr, ext, &vt, true);
  double Rl[9], dRl[9];
  rodrigues_jl(tlw, Rl, dRl);
  if (disp_on) ba_pt<true, true>(vt, Rl, dRl, tlw + 3, disp, xyz, uv[0], uv[1], r, Jc, Jt, Jd);
  else ba_pt<false, true>(vt, Rl, dRl, tlw + 3, disp, xyz, uv[0], uv[1], r, Jc, Jt, Jd);
}
int hm_krt_obs(int type, const double* cam15, const double* refK4, const double* refd, const float* uv1, const float* uv2, double* r, double* J) {
  double ray1[3];
  bool ok = krt_precompute(type, refK4, refd, uv1[0], uv1[1], ray1);
  const int nf = krt_nfree(type);
  if (!ok) { r[0] = r[1] = 0; for (int i = 0; i < 2 * nf; ++i) J[i] = 0; return 0; }
  KrtCam kc;
  switch (type) {
    case 0: krt_make_cam<0>(cam15, &kc, true); krt_obs<0, true>(kc, ray1, uv2[0], uv2[1], r, J); break;
    case 1: krt_make_cam<1>(cam15, &kc, true); krt_obs<1, true>(kc, ray1, uv2[0], uv2[1], r, J); break;
    case 2: krt_make_cam<2>(cam15, &kc, true); krt_obs<2, true>(kc, ray1, uv2[0], uv2[1], r, J); break;
    default: krt_make_cam<3>(cam15, &kc, true); krt_obs<3, true>(kc, ray1, uv2[0], uv2[1], r, J); break;
  }
  return 1;
}
void hm_rodrigues_jac(const double* w, double* R, double* dR) { rodrigues_jac(w, R, dR); }
void hm_krt_obs3d(int type, const double* cam15, const float* uv, const double* P, double* r, double* J) {
  KrtCam kc;
  switch (type) {
    case 0: krt_make_cam<0>(cam15, &kc, true); krt_obs3d<0, true>(kc, cam15 + 7, P, uv[0], uv[1], r, J); break;
    case 1: krt_make_cam<1>(cam15, &kc, true); krt_obs3d<1, true>(kc, cam15 + 7, P, uv[0], uv[1], r, J); break;
    case 2: krt_make_cam<2>(cam15, &kc, true); krt_obs3d<2, true>(kc, cam15 + 7, P, uv[0], uv[1], r, J); break;
    default: krt_make_cam<3>(cam15, &kc, true); krt_obs3d<3, true>(kc, cam15 + 7, P, uv[0], uv[1], r, J); break;
  }
}
}
