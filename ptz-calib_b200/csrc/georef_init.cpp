// ptzgeo_init_tlw: the EPnP initialisation of T_l_w (PTZRayOptimizer::SetInitTransLocalToWorld, ptzray_optimizer.cc:562-633) behind the C ABI,
// for hosts that do not go through the C++ adaptor (the Python mirror's from_matches).  Host code only; the work is in
// include/ptzcalib_epnp.hpp (epnp::init_tlw_from_view), which the adaptor's member function calls too.
#include "../../include/ptzcalib_b200.h"
#include "../../include/ptzcalib_epnp.hpp"

extern "C" int ptzgeo_init_tlw(int32_t num_views, const double* cams21, const int64_t* pt_offset, const float* pt_uv, const double* pt_xyz, double tlw[6],
                               int32_t* view_used) {
  if (num_views < 0 || !tlw || (num_views > 0 && (!cams21 || !pt_offset))) return PTZ_ERR_INVALID;
  for (int k = 0; k < 6; ++k) tlw[k] = 0.0;
  if (view_used) *view_used = -1;
  for (int32_t i = 0; i < num_views; ++i) {
    const int64_t lo = pt_offset[i], n = pt_offset[i + 1] - lo;
    if (n < 0 || lo < 0) return PTZ_ERR_INVALID;
    if (n == 0) continue;
    if (!pt_uv || !pt_xyz) return PTZ_ERR_INVALID;
    const double* c = cams21 + 21 * (size_t)i;
    const double K[9] = {c[0], 0, c[2], 0, c[1], c[3], 0, 0, 1};
    try {
      if (!ptzcalib::epnp::init_tlw_from_view((int)n, pt_xyz + 3 * lo, pt_uv + 2 * lo, K, c + 16, c + 4, c + 13, tlw)) {
        for (int k = 0; k < 6; ++k) tlw[k] = 0.0;
        continue;
      }
    } catch (...) { return PTZ_ERR_INVALID; }
    if (view_used) *view_used = i;
    return PTZ_OK;
  }
  return PTZ_OK;
}
