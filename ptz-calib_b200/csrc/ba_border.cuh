// ba_border.cuh — the georeferencing terms of PTZ-BA: annotated 2d-3d points (AddConstraints2d3d,
// ptzray_optimizer.cc:887-958; Reproj2d3dFactor :268-326).  There are only tens of them, in a handful of views, so they
// are handled by ONE CTA with fixed-order loops (deterministic, no atomics).  They couple the annotated views to a
// small dense BORDER of the reduced system:
//     border = [ tlw(6) | fy of each annotated view (factor types that tie fy:=fx in the ray terms) ]
// (fy is identically-zero-column in the ray factors of those types and is driven only by these terms: SURVEY.md A6.)
#pragma once
#include "ba_kernels.cuh"

namespace ptz {

struct PtsArgs {
  int A, nav, nb, fy_in_border;     // fy_in_border = 1 for PTZRay / PTZRayDist
  const float2* uv; const double* xyz; const int* view;   // sorted by view
  const int* ann_view; const int* ann_off;                // [nav], [nav+1]
  const ViewTab* vt; const double* tlw;                   // current parameters
  const double* scale_cam; const double* scale_b;
  double* scratch;                  // [A][2 + 2*NCL + 2*nb] scaled rows
  double* raw;                      // optional [A][2 + 12 + 12]: r, Jc(2x6), Jt(2x6) unscaled (ptzba_eval), or nullptr
  // outputs
  double* U; double* g; double* gabs;   // += on the annotated views
  double* C;                        // [nav][NCL][nb]
  double* Hbb; double* gb;          // [nb*nb], [nb]
  double* cost_pts;                 // [2]: 1/2 sum r^2, sum r^2
  double* gabs_b;                   // [nb]
};

template <int TYPE>
__global__ void __launch_bounds__(128) k_pts(PtsArgs a) {
  constexpr int NCL = ba_ncl(TYPE);
  const int nb = a.nb, RW = 2 + 2 * NCL + 2 * nb;
  __shared__ double sR[9], sdR[9], st[3];
  if (threadIdx.x == 0) {
    double w[3] = {a.tlw[0], a.tlw[1], a.tlw[2]};
    double R[9], Jl[9];
    rodrigues_jl(w, R, Jl);
    for (int i = 0; i < 9; ++i) sR[i] = R[i];
    for (int i = 0; i < 9; ++i) sdR[i] = Jl[i];
    st[0] = a.tlw[3]; st[1] = a.tlw[4]; st[2] = a.tlw[5];
  }
  __syncthreads();
  // phase 1: one thread per annotated point
  for (int i = threadIdx.x; i < a.A; i += blockDim.x) {
    const int v = a.view[i];
    int k = 0;
    while (a.ann_view[k] != v) ++k;
    const ViewTab vt = a.vt[v];
    const double Xw[3] = {a.xyz[3 * i], a.xyz[3 * i + 1], a.xyz[3 * i + 2]};
    double r[2], Jc[12], Jt[12];
    const double dz[3] = {0, 0, 0};
    ba_pt<false, true>(vt, sR, sdR, st, dz, Xw, (double)a.uv[i].x, (double)a.uv[i].y, r, Jc, Jt, nullptr);
    if (a.raw) {
      double* o = a.raw + (size_t)i * 26;
      o[0] = r[0]; o[1] = r[1];
      for (int j = 0; j < 12; ++j) { o[2 + j] = Jc[j]; o[14 + j] = Jt[j]; }
    }
    double* row = a.scratch + (size_t)i * RW;
    row[0] = r[0]; row[1] = r[1];
    double* Fc = row + 2;            // [2][NCL]
    double* Bj = row + 2 + 2 * NCL;  // [2][nb]
    for (int j = 0; j < 2 * nb; ++j) Bj[j] = 0.0;
    for (int rr = 0; rr < 2; ++rr) {
      const double* jc = Jc + 6 * rr;
      double live[6];
      int n = 0;
      if (TYPE == BA_PTZRAY) { live[n++] = jc[0]; live[n++] = jc[3]; live[n++] = jc[4]; live[n++] = jc[5]; }
      else if (TYPE == BA_PTZRAY_FXFY_DIST) { for (int j = 0; j < 6; ++j) live[n++] = jc[j]; }
      else { live[n++] = jc[0]; live[n++] = jc[2]; live[n++] = jc[3]; live[n++] = jc[4]; live[n++] = jc[5]; }
      for (int c = 0; c < NCL; ++c) Fc[rr * NCL + c] = live[c] * a.scale_cam[v * NCL + c];
      for (int j = 0; j < 6; ++j) Bj[rr * nb + j] = Jt[6 * rr + j] * a.scale_b[j];
      if (a.fy_in_border) Bj[rr * nb + 6 + k] = jc[1] * a.scale_b[6 + k];
    }
  }
  __syncthreads();
  // phase 2a: per annotated view, camera block / gradient / coupling strip (fixed order over its points)
  for (int k = threadIdx.x; k < a.nav; k += blockDim.x) {
    const int v = a.ann_view[k];
    double Uadd[NCL * NCL], gadd[NCL];
    for (int i = 0; i < NCL * NCL; ++i) Uadd[i] = 0;
    for (int i = 0; i < NCL; ++i) gadd[i] = 0;
    double* Ck = a.C + (size_t)k * NCL * nb;
    for (int i = 0; i < NCL * nb; ++i) Ck[i] = 0;
    for (int i = a.ann_off[k]; i < a.ann_off[k + 1]; ++i) {
      const double* row = a.scratch + (size_t)i * RW;
      const double* Fc = row + 2;
      const double* Bj = row + 2 + 2 * NCL;
      for (int c = 0; c < NCL; ++c) {
        gadd[c] += Fc[c] * row[0] + Fc[NCL + c] * row[1];
        for (int d = 0; d < NCL; ++d) Uadd[c * NCL + d] += Fc[c] * Fc[d] + Fc[NCL + c] * Fc[NCL + d];
        for (int j = 0; j < nb; ++j) Ck[c * nb + j] += Fc[c] * Bj[j] + Fc[NCL + c] * Bj[nb + j];
      }
    }
    for (int i = 0; i < NCL * NCL; ++i) a.U[(size_t)v * NCL * NCL + i] += Uadd[i];
    for (int c = 0; c < NCL; ++c) {
      const double gv = a.g[v * NCL + c] + gadd[c];
      a.g[v * NCL + c] = gv;
      a.gabs[v * NCL + c] = fabs(gv / a.scale_cam[v * NCL + c]);
    }
  }
  // phase 2b: border block and gradient
  for (int e = threadIdx.x; e < nb * nb + nb; e += blockDim.x) {
    double s = 0;
    if (e < nb * nb) {
      const int i = e / nb, j = e % nb;
      for (int p = 0; p < a.A; ++p) {
        const double* Bj = a.scratch + (size_t)p * RW + 2 + 2 * NCL;
        s += Bj[i] * Bj[j] + Bj[nb + i] * Bj[nb + j];
      }
      a.Hbb[e] = s;
    } else {
      const int i = e - nb * nb;
      for (int p = 0; p < a.A; ++p) {
        const double* row = a.scratch + (size_t)p * RW;
        const double* Bj = row + 2 + 2 * NCL;
        s += Bj[i] * row[0] + Bj[nb + i] * row[1];
      }
      a.gb[i] = s;
      a.gabs_b[i] = fabs(s / a.scale_b[i]);
    }
  }
  if (threadIdx.x == 0) {
    double s = 0;
    for (int p = 0; p < a.A; ++p) { const double* row = a.scratch + (size_t)p * RW; s += row[0] * row[0] + row[1] * row[1]; }
    a.cost_pts[0] = 0.5 * s;
    a.cost_pts[1] = s;
  }
}

// cost of the annotated points at the candidate parameters (one warp)
__global__ void k_pts_cost(int A, const float2* __restrict__ uv, const double* __restrict__ xyz, const int* __restrict__ view, const ViewTab* __restrict__ vt,
                           const double* __restrict__ tlw, double* __restrict__ out2) {
  __shared__ double sR[9], st[3];
  if (threadIdx.x == 0) {
    double w[3] = {tlw[0], tlw[1], tlw[2]}, R[9];
    rodrigues_jac(w, R, nullptr);
    for (int i = 0; i < 9; ++i) sR[i] = R[i];
    st[0] = tlw[3]; st[1] = tlw[4]; st[2] = tlw[5];
  }
  __syncthreads();
  double s = 0;
  for (int i = threadIdx.x; i < A; i += 32) {
    const ViewTab t = vt[view[i]];
    const double Xw[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    const double dz[3] = {0, 0, 0};
    double r[2];
    ba_pt<false, false>(t, sR, nullptr, st, dz, Xw, (double)uv[i].x, (double)uv[i].y, r, nullptr, nullptr, nullptr);
    s += r[0] * r[0] + r[1] * r[1];
  }
  s = warp_sum(s);
  if (threadIdx.x == 0) { out2[0] = 0.5 * s; out2[1] = s; }
}

// Jacobi scaling of the border columns at iteration 0
__global__ void k_border_scales(int nb, const double* __restrict__ Hbb, double* __restrict__ scale_b) {
  const int i = threadIdx.x;
  if (i < nb) scale_b[i] = 1.0 / (1.0 + sqrt(Hbb[i * nb + i]));
}

// S_bb = H_bb + D_b^2, rhs_b = g_b  (D_b^2 = clamp(diag H_bb)/mu, refreshed after accepted steps only)
__global__ void k_border_system(int nb, const double* __restrict__ Hbb, const double* __restrict__ gb, double mu, int refresh_diag, double min_diag,
                                double max_diag, double* __restrict__ diag_b, double* __restrict__ Sbb, double* __restrict__ rhs_b) {
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    double s = Hbb[e];
    if (i == j) {
      double d;
      if (refresh_diag) { d = fmin(fmax(Hbb[e], min_diag), max_diag); diag_b[i] = d; }
      else d = diag_b[i];
      s += d / mu;
      rhs_b[i] = gb[i];
    }
    Sbb[e] = s;
  }
}

// candidate border parameters; part3 = {model cost change, |step|^2, |x_cand|^2} contributions (fy corrections included)
__global__ void k_border_update(int nb, int nav, int fy_in_border, const int* __restrict__ ann_view, const double* __restrict__ yb,
                                const double* __restrict__ scale_b, const double* __restrict__ gb, const double* __restrict__ diag_b, double mu,
                                const double* __restrict__ tlw, double* __restrict__ tlw_c, const double* __restrict__ intr, double* __restrict__ intr_c,
                                double* __restrict__ part3) {
  if (threadIdx.x != 0) return;
  double dm = 0, st = 0, xn = 0;
  for (int j = 0; j < nb; ++j) dm += 0.5 * yb[j] * (gb[j] + diag_b[j] / mu * yb[j]);
  for (int j = 0; j < 6; ++j) {
    const double c = tlw[j] + (-scale_b[j] * yb[j]);
    tlw_c[j] = c;
    st += (tlw[j] - c) * (tlw[j] - c);
    xn += c * c;
  }
  if (fy_in_border)
    for (int k = 0; k < nav; ++k) {
      const int v = ann_view[k];
      const double x = intr[9 * v + 1], c = x + (-scale_b[6 + k] * yb[6 + k]);
      intr_c[9 * v + 1] = c;
      st += (x - c) * (x - c);
      xn += c * c - x * x;  // k_cam_update counted the unchanged fy
    }
  part3[0] = dm; part3[1] = st; part3[2] = xn;
}

}  // namespace ptz
