// ba_border.cuh — the georeferencing terms of PTZ-BA: annotated 2d-3d points (AddConstraints2d3d,
// ptzray_optimizer.cc:887-958; Reproj2d3dFactor :268-326).  They are few next to the ray observations, so ONE CTA handles
// them with fixed-order loops (deterministic, no atomics).  They bring two kinds of unknowns besides the cameras:
//     border  = [ tlw(6) | disp(3) ]                 <= 9 unknowns, a dense row/column of the reduced system (k_cg's border row)
//     fy_k    = fy of annotated view k               (factor types that tie fy := fx in the ray terms: the column is identically
//                                                     zero there and is driven only by these terms)
// fy_k touches nothing but its own view's points, i.e. the camera block of view k and the border: it is ELIMINATED like a ray
// (1x1 Schur complement, k_border_system / k_fy_eliminate / back-substitution in k_border_update), so any number of views may be
// annotated.  Vectors over [border | fy] (scales, LM diagonal, gradient) are stored with the fy part behind the nb border entries.
#pragma once
#include "ba_kernels.cuh"

namespace ptz {

struct PtsArgs {
  int A, nav, nb, nf;               // nf = nav when fy lives outside the camera block (fy := fx factor types), else 0
  int bo_tlw, bo_disp;              // first border column of tlw(6) / disp(3), or -1
  const int* ann_strip;             // strip row of annotated view k inside C (k itself, or the view id when every view has a strip)
  const double* disp;               // current disp[3] (PTZRayDistDisp) or nullptr
  const float2* uv; const double* xyz; const int* view;   // sorted by view
  const int* ann_view; const int* ann_off;                // [nav], [nav+1]
  const ViewTab* vt; const double* tlw;                   // current parameters
  const double* scale_cam; const double* scale_b;         // scale_b: [nb + nf]
  double* scratch;                  // [A][2 + 2*NCL + 2*nb + 2] scaled rows: r, Fc, B (border columns), bf (fy column)
  double* raw;                      // optional [A][2 + 12 + 12 + 6]: r, Jc(2x6), Jt(2x6), Jd(2x3) unscaled (ptzba_eval), or nullptr
  // outputs
  double* U; double* g; double* gabs;   // += on the annotated views
  double* C;                        // [strips][NCL][nb]
  double* Hbb; double* gb;          // [nb*nb], [nb + nf]
  double* Cf; double* Hrf; double* Hff;  // fy_k: coupling to its camera block [nf][NCL], to the border [nf][nb], own diagonal [nf]
  double* cost_pts;                 // [2]: 1/2 sum r^2, sum r^2
  double* gabs_b;                   // [nb + nf]
};

template <int TYPE>
__global__ void __launch_bounds__(128) k_pts(PtsArgs a) {
  constexpr int NCL = ba_ncl(TYPE);
  const int nb = a.nb, RW = 2 + 2 * NCL + 2 * nb + 2;
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
    {  // annotated-view index of v (ann_view is ascending)
      int lo = 0, hi = a.nav - 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.ann_view[mid] < v) lo = mid + 1; else hi = mid; }
      k = lo;
    }
    const ViewTab vt = a.vt[v];
    const double Xw[3] = {a.xyz[3 * i], a.xyz[3 * i + 1], a.xyz[3 * i + 2]};
    double r[2], Jc[12], Jt[12], Jd[6] = {0, 0, 0, 0, 0, 0};
    double dz[3] = {0, 0, 0};
    if (TYPE == BA_PTZRAY_DIST_DISP) {
      dz[0] = a.disp[0]; dz[1] = a.disp[1]; dz[2] = a.disp[2];
      ba_pt<true, true>(vt, sR, sdR, st, dz, Xw, (double)a.uv[i].x, (double)a.uv[i].y, r, Jc, Jt, Jd);
    } else {
      ba_pt<false, true>(vt, sR, sdR, st, dz, Xw, (double)a.uv[i].x, (double)a.uv[i].y, r, Jc, Jt, nullptr);
    }
    if (a.raw) {
      double* o = a.raw + (size_t)i * 32;
      o[0] = r[0]; o[1] = r[1];
      for (int j = 0; j < 12; ++j) { o[2 + j] = Jc[j]; o[14 + j] = Jt[j]; }
      for (int j = 0; j < 6; ++j) o[26 + j] = Jd[j];
    }
    double* row = a.scratch + (size_t)i * RW;
    row[0] = r[0]; row[1] = r[1];
    double* Fc = row + 2;            // [2][NCL]
    double* Bj = row + 2 + 2 * NCL;  // [2][nb]
    double* bf = Bj + 2 * nb;        // [2]
    for (int j = 0; j < 2 * nb; ++j) Bj[j] = 0.0;
    for (int rr = 0; rr < 2; ++rr) {
      const double* jc = Jc + 6 * rr;
      double live[6];
      int n = 0;
      if (TYPE == BA_PTZRAY) { live[n++] = jc[0]; live[n++] = jc[3]; live[n++] = jc[4]; live[n++] = jc[5]; }
      else if (TYPE == BA_PTZRAY_FXFY_DIST) { for (int j = 0; j < 6; ++j) live[n++] = jc[j]; }
      else { live[n++] = jc[0]; live[n++] = jc[2]; live[n++] = jc[3]; live[n++] = jc[4]; live[n++] = jc[5]; }
      for (int c = 0; c < NCL; ++c) Fc[rr * NCL + c] = live[c] * a.scale_cam[v * NCL + c];
      for (int j = 0; j < 6; ++j) Bj[rr * nb + a.bo_tlw + j] = Jt[6 * rr + j] * a.scale_b[a.bo_tlw + j];
      bf[rr] = a.nf > 0 ? jc[1] * a.scale_b[nb + k] : 0.0;
      if (TYPE == BA_PTZRAY_DIST_DISP) for (int j = 0; j < 3; ++j) Bj[rr * nb + a.bo_disp + j] = Jd[3 * rr + j] * a.scale_b[a.bo_disp + j];
    }
  }
  __syncthreads();
  // phase 2a: per annotated view, camera block / gradient / coupling strip and the fy terms (fixed order over its points)
  for (int k = threadIdx.x; k < a.nav; k += blockDim.x) {
    const int v = a.ann_view[k];
    double Uadd[NCL * NCL], gadd[NCL], cf[NCL], hff = 0, gf = 0;
    for (int i = 0; i < NCL * NCL; ++i) Uadd[i] = 0;
    for (int i = 0; i < NCL; ++i) { gadd[i] = 0; cf[i] = 0; }
    double* Ck = a.C + (size_t)a.ann_strip[k] * NCL * nb;  // zeroed by the caller; the disp kernels add to it as well
    double* Hk = a.nf > 0 ? a.Hrf + (size_t)k * nb : nullptr;  // zeroed by the caller
    for (int i = a.ann_off[k]; i < a.ann_off[k + 1]; ++i) {
      const double* row = a.scratch + (size_t)i * RW;
      const double* Fc = row + 2;
      const double* Bj = row + 2 + 2 * NCL;
      const double* bf = Bj + 2 * nb;
      for (int c = 0; c < NCL; ++c) {
        gadd[c] += Fc[c] * row[0] + Fc[NCL + c] * row[1];
        for (int d = 0; d < NCL; ++d) Uadd[c * NCL + d] += Fc[c] * Fc[d] + Fc[NCL + c] * Fc[NCL + d];
        for (int j = 0; j < nb; ++j) Ck[c * nb + j] += Fc[c] * Bj[j] + Fc[NCL + c] * Bj[nb + j];
        cf[c] += Fc[c] * bf[0] + Fc[NCL + c] * bf[1];
      }
      if (a.nf > 0) {
        for (int j = 0; j < nb; ++j) Hk[j] += bf[0] * Bj[j] + bf[1] * Bj[nb + j];
        hff += bf[0] * bf[0] + bf[1] * bf[1];
        gf += bf[0] * row[0] + bf[1] * row[1];
      }
    }
    for (int i = 0; i < NCL * NCL; ++i) a.U[(size_t)v * NCL * NCL + i] += Uadd[i];
    for (int c = 0; c < NCL; ++c) {
      const double gv = a.g[v * NCL + c] + gadd[c];
      a.g[v * NCL + c] = gv;
      a.gabs[v * NCL + c] = fabs(gv / a.scale_cam[v * NCL + c]);
    }
    if (a.nf > 0) {
      for (int c = 0; c < NCL; ++c) a.Cf[(size_t)k * NCL + c] = cf[c];
      a.Hff[k] = hff;
      a.gb[nb + k] = gf;
      a.gabs_b[nb + k] = fabs(gf / a.scale_b[nb + k]);
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
      a.Hbb[e] += s;
    } else {
      const int i = e - nb * nb;
      for (int p = 0; p < a.A; ++p) {
        const double* row = a.scratch + (size_t)p * RW;
        const double* Bj = row + 2 + 2 * NCL;
        s += Bj[i] * row[0] + Bj[nb + i] * row[1];
      }
      a.gb[i] += s;
      a.gabs_b[i] = fabs(a.gb[i] / a.scale_b[i]);
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
                           const double* __restrict__ tlw, const double* __restrict__ disp /* or nullptr */, double* __restrict__ out2) {
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
    double r[2];
    if (disp) {
      const double dz[3] = {disp[0], disp[1], disp[2]};
      ba_pt<true, false>(t, sR, nullptr, st, dz, Xw, (double)uv[i].x, (double)uv[i].y, r, nullptr, nullptr, nullptr);
    } else {
      const double dz[3] = {0, 0, 0};
      ba_pt<false, false>(t, sR, nullptr, st, dz, Xw, (double)uv[i].x, (double)uv[i].y, r, nullptr, nullptr, nullptr);
    }
    s += r[0] * r[0] + r[1] * r[1];
  }
  s = warp_sum(s);
  if (threadIdx.x == 0) { out2[0] = 0.5 * s; out2[1] = s; }
}

// Jacobi scaling of the border and fy columns at iteration 0
__global__ void k_border_scales(int nb, int nf, const double* __restrict__ Hbb, const double* __restrict__ Hff, double* __restrict__ scale_b,
                                int nb_plain, const double* __restrict__ sh_h /* column norms^2 of the shared intrinsics unknowns */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb_plain) scale_b[i] = 1.0 / (1.0 + sqrt(Hbb[i * nb + i]));
  else if (i < nb) scale_b[i] = 1.0 / (1.0 + sqrt(sh_h[i - nb_plain]));
  else if (i < nb + nf) scale_b[i] = 1.0 / (1.0 + sqrt(Hff[i - nb]));
}
// |g_j / s_j| of every camera, border and fy column from the (all-reduced) gradient: the sharded problem's replacement for the
// values k_view_finalize / k_pts wrote from rank-local sums
__global__ void k_grad_abs(int ncam, const double* __restrict__ g, const double* __restrict__ scale_cam, double* __restrict__ gabs, int nbt,
                           const double* __restrict__ gb, const double* __restrict__ scale_b, double* __restrict__ gabs_b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncam) gabs[i] = fabs(g[i] / scale_cam[i]);
  else if (i < ncam + nbt) gabs_b[i - ncam] = fabs(gb[i - ncam] / scale_b[i - ncam]);
}

// Border system of one linear solve, with the fy unknowns eliminated:
//   h_k = Hff_k + D_k^2          S_bb = H_bb + D_b^2 - sum_k Hrf_k Hrf_k^T / h_k          rhs_b = g_b - sum_k Hrf_k gf_k / h_k
// (D^2 = clamp(diag)/mu, refreshed after accepted steps only; sums over k in ascending order).  One CTA.
__global__ void __launch_bounds__(128) k_border_system(int nb, int nf, const double* __restrict__ Hbb, const double* __restrict__ Hrf,
                                                       const double* __restrict__ Hff, const double* __restrict__ gb, double mu, int refresh_diag,
                                                       double min_diag, double max_diag, int own /* sharded problem: rank 0 contributes, the others write zeros */,
                                                       double* __restrict__ diag_b, double* __restrict__ hinv, double* __restrict__ Sbb,
                                                       double* __restrict__ rhs_b, int nb_plain /* entries behind it: shared intrinsics, k_shared_border */) {
  for (int k = threadIdx.x; k < nf; k += blockDim.x) {
    double d;
    if (refresh_diag) { d = fmin(fmax(Hff[k], min_diag), max_diag); diag_b[nb + k] = d; }
    else d = diag_b[nb + k];
    hinv[k] = 1.0 / (Hff[k] + d / mu);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    double s = Hbb[e];
    if (i >= nb_plain || j >= nb_plain) { Sbb[e] = 0.0; if (i == j) rhs_b[i] = 0.0; continue; }
    if (i == j) {
      double d;
      if (refresh_diag) { d = fmin(fmax(Hbb[e], min_diag), max_diag); diag_b[i] = d; }
      else d = diag_b[i];
      s += d / mu;
      double r = gb[i];
      for (int k = 0; k < nf; ++k) r -= Hrf[(size_t)k * nb + i] * gb[nb + k] * hinv[k];
      rhs_b[i] = own ? r : 0.0;
    }
    for (int k = 0; k < nf; ++k) s -= Hrf[(size_t)k * nb + i] * Hrf[(size_t)k * nb + j] * hinv[k];
    Sbb[e] = own ? s : 0.0;
  }
}
// camera side of the fy elimination (one thread per annotated view k, view v):
//   S_vv -= Cf_k Cf_k^T / h_k      rhs_v -= Cf_k gf_k / h_k      strip_k = C_k - Cf_k Hrf_k^T / h_k   (into the working strips Cw)
template <int NCL>
__global__ void k_fy_eliminate(int nf, int nb, const int* __restrict__ ann_view, const int* __restrict__ ann_strip, const double* __restrict__ Cf,
                               const double* __restrict__ Hrf, const double* __restrict__ gf, const double* __restrict__ hinv,
                               const double* __restrict__ Csrc, double* __restrict__ Cw, const int* __restrict__ diag_pos, double* __restrict__ Sval,
                               double* __restrict__ rhs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nf) return;
  const int v = ann_view[k];
  const double hi = hinv[k];
  double cf[NCL];
#pragma unroll
  for (int a = 0; a < NCL; ++a) cf[a] = Cf[(size_t)k * NCL + a];
  double* S = Sval + (size_t)diag_pos[v] * NCL * NCL;
#pragma unroll
  for (int a = 0; a < NCL; ++a) {
#pragma unroll
    for (int b = 0; b < NCL; ++b) S[a * NCL + b] -= cf[a] * cf[b] * hi;
    rhs[v * NCL + a] -= cf[a] * gf[k] * hi;
  }
  const size_t so = (size_t)ann_strip[k] * NCL * nb;
  for (int a = 0; a < NCL; ++a)
    for (int j = 0; j < nb; ++j) Cw[so + a * nb + j] = Csrc[so + a * nb + j] - cf[a] * Hrf[(size_t)k * nb + j] * hi;
}

// candidate border parameters and the back-substituted fy steps  y_fk = (gf_k - Cf_k^T y_v - Hrf_k^T y_b) / h_k;
// part3 = {model cost change, |step|^2, |x_cand|^2} contributions.  One warp, lanes stride the annotated views.
template <int NCL>
__global__ void __launch_bounds__(32) k_border_update(int nb, int nf, int bo_tlw, int bo_disp, const int* __restrict__ ann_view, const double* __restrict__ y,
                                                      int ncam, const double* __restrict__ scale_b, const double* __restrict__ gb,
                                                      const double* __restrict__ diag_b, double mu, const double* __restrict__ Cf,
                                                      const double* __restrict__ Hrf, const double* __restrict__ hinv, const double* __restrict__ tlw,
                                                      double* __restrict__ tlw_c, const double* __restrict__ intr, double* __restrict__ intr_c,
                                                      const double* __restrict__ disp, double* __restrict__ disp_c, double* __restrict__ part3, int nb_plain) {
  const int lane = threadIdx.x;
  const double* yb = y + ncam;
  double dm = 0, st = 0, xn = 0;
  if (lane == 0) {
    for (int j = 0; j < nb_plain; ++j) dm += 0.5 * yb[j] * (gb[j] + diag_b[j] / mu * yb[j]);
    // shared intrinsics unknowns: k_cam_update already summed 1/2 y g over the views of the group; the damping term is theirs once
    for (int j = nb_plain; j < nb; ++j) dm += 0.5 * yb[j] * (diag_b[j] / mu * yb[j]);
    if (bo_tlw >= 0)
      for (int j = 0; j < 6; ++j) {
        const double c = tlw[j] + (-scale_b[bo_tlw + j] * yb[bo_tlw + j]);
        tlw_c[j] = c;
        st += (tlw[j] - c) * (tlw[j] - c);
        xn += c * c;
      }
    if (bo_disp >= 0)
      for (int j = 0; j < 3; ++j) {
        const double c = disp[j] + (-scale_b[bo_disp + j] * yb[bo_disp + j]);
        disp_c[j] = c;
        st += (disp[j] - c) * (disp[j] - c);
        xn += c * c;
      }
  }
  for (int k = lane; k < nf; k += 32) {
    const int v = ann_view[k];
    double t = gb[nb + k];
#pragma unroll
    for (int a = 0; a < NCL; ++a) t -= Cf[(size_t)k * NCL + a] * y[v * NCL + a];
    for (int j = 0; j < nb; ++j) t -= Hrf[(size_t)k * nb + j] * yb[j];
    const double yf = t * hinv[k];
    dm += 0.5 * yf * (gb[nb + k] + diag_b[nb + k] / mu * yf);
    const double x = intr[9 * v + 1], c = x + (-scale_b[nb + k] * yf);
    intr_c[9 * v + 1] = c;
    st += (x - c) * (x - c);
    xn += c * c - x * x;  // k_cam_update counted the unchanged fy
  }
  dm = warp_sum(dm); st = warp_sum(st); xn = warp_sum(xn);
  if (lane == 0) { part3[0] = dm; part3[1] = st; part3[2] = xn; }
}

// ---------------------------------------------------------------------------------------------------------------------
// SetSharedIntrinsics (ptzray_optimizer.cc:497-505; wiring :640-650, :821-848): the views of a group use ONE intrinsics block.
// Everything up to the reduced camera system is formed per view as if the blocks were independent (records, U, g, What, S); the
// constraint "the intrinsics columns of all views of group G are the same unknown" is then applied to the reduced system itself,
// S' = T^T S T: the nI = NCL - 3 intrinsics unknowns of every group of two or more views move into the dense border
//     border = [ tlw | disp | group 0: fx (k1) (fy) | group 1: ... ]
// and the rows / columns they had in the camera blocks are deactivated (unit diagonal, zero right-hand side).  Exact: rays are
// eliminated per observation, and T only adds camera columns.  Jacobi scale, LM diagonal and gradient of a shared column come from
// the sums over the group (k_shared_grad); the step is copied back into the per-view slots (k_shared_expand), so the per-view
// update kernels run unchanged and the views of a group stay bit-identical.
// ---------------------------------------------------------------------------------------------------------------------
// after every Jacobian evaluation (and all-reduce): gradient and squared column norm of each shared unknown j = (group, column)
template <int NCL>
__global__ void k_shared_grad(int ns, const int* __restrict__ grp_off, const int* __restrict__ grp_view, const double* __restrict__ U,
                              const double* __restrict__ g, const double* __restrict__ scale_b, int bo_sh, double* __restrict__ gb,
                              double* __restrict__ gabs_b, double* __restrict__ sh_h, double* __restrict__ gabs) {
  constexpr int NI = NCL - 3;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const int grp = j / NI, c = j % NI;
  double gs = 0, hs = 0;
  for (int k = grp_off[grp]; k < grp_off[grp + 1]; ++k) {
    const int v = grp_view[k];
    gs += g[v * NCL + c];
    hs += U[(size_t)v * NCL * NCL + c * NCL + c];
    gabs[v * NCL + c] = 0.0;  // not a column of its own
  }
  gb[bo_sh + j] = gs;
  gabs_b[bo_sh + j] = fabs(gs / scale_b[bo_sh + j]);
  sh_h[j] = hs;
}
// iteration 0: every view of a group takes the group's Jacobi scale
template <int NCL>
__global__ void k_shared_scales(int V, const int* __restrict__ grp_of, const double* __restrict__ scale_b, int bo_sh, double* __restrict__ scale_cam) {
  constexpr int NI = NCL - 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = i / NI, c = i % NI;
  if (v >= V || grp_of[v] < 0) return;
  scale_cam[v * NCL + c] = scale_b[bo_sh + grp_of[v] * NI + c];
}
// per linear solve, one thread per block row w of S: the shared columns of every block (w, v) are summed into the strip of row w
// and zeroed; if w is itself in a group its intrinsics rows move to `rowpart` (summed over the group by k_shared_border) and are
// deactivated.  Every entry of S is touched by the thread of its row only.
template <int NCL>
__global__ void k_shared_fold(int V, int nb, int nb_plain, const int* __restrict__ grp_of, const int* __restrict__ rowptr, const int* __restrict__ col,
                              double* __restrict__ Sval, double* __restrict__ rhs, const double* __restrict__ Csrc /* strips so far, or nullptr */,
                              double* __restrict__ Cw, double* __restrict__ rowpart /* [V][NI][nb + 1] */) {
  constexpr int NI = NCL - 3, NB = NCL * NCL;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= V) return;
  double strip[NCL][kMaxBorder];
#pragma unroll
  for (int a = 0; a < NCL; ++a)
#pragma unroll
    for (int j = 0; j < kMaxBorder; ++j) strip[a][j] = (Csrc != nullptr && j < nb_plain) ? Csrc[((size_t)w * NCL + a) * nb + j] : 0.0;
  const int gw = grp_of[w];
  for (int k = rowptr[w]; k < rowptr[w + 1]; ++k) {
    const int v = col[k], gv = grp_of[v];
    double* B = Sval + (size_t)k * NB;
    if (gv >= 0) {
#pragma unroll
      for (int c = 0; c < NI; ++c) {
        const int j = nb_plain + gv * NI + c;
#pragma unroll
        for (int a = 0; a < NCL; ++a) {
#pragma unroll
          for (int jj = 0; jj < kMaxBorder; ++jj)
            if (jj == j) strip[a][jj] += B[a * NCL + c];
          B[a * NCL + c] = (v == w && a == c) ? 1.0 : 0.0;
        }
      }
    }
    if (gw >= 0) {
#pragma unroll
      for (int a = 0; a < NI; ++a)
#pragma unroll
        for (int c = 0; c < NCL; ++c) B[a * NCL + c] = (v == w && a == c) ? 1.0 : 0.0;  // (its content went into strip[a][.] / the other rows' strips)
    }
  }
  // NOTE: the intrinsics rows of a grouped w were zeroed above AFTER their shared columns were taken into strip[a][.]; their
  // non-shared columns (the coupling to the rotation unknowns of the neighbours) are the transposes of what the neighbours' threads
  // put into THEIR strips, so nothing is lost.
  if (gw >= 0) {
#pragma unroll
    for (int a = 0; a < NI; ++a) {
      double* rp = rowpart + ((size_t)w * NI + a) * (nb + 1);
#pragma unroll
      for (int j = 0; j < kMaxBorder; ++j)
        if (j < nb) { rp[j] = strip[a][j]; strip[a][j] = 0.0; }
      rp[nb] = rhs[w * NCL + a];
      rhs[w * NCL + a] = 0.0;
    }
  }
#pragma unroll
  for (int a = 0; a < NCL; ++a)
#pragma unroll
    for (int j = 0; j < kMaxBorder; ++j)
      if (j < nb) Cw[((size_t)w * NCL + a) * nb + j] = strip[a][j];
}
// one thread per shared unknown (row of the border): sum its group's row parts in member order, add the LM diagonal
template <int NCL>
__global__ void k_shared_border(int ns, int nb, int nb_plain, const int* __restrict__ grp_off, const int* __restrict__ grp_view,
                                const double* __restrict__ rowpart, const double* __restrict__ sh_h, double mu, int refresh_diag, double min_diag,
                                double max_diag, double* __restrict__ diag_b, double* __restrict__ Sbb, double* __restrict__ rhs_b) {
  constexpr int NI = NCL - 3;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const int grp = j / NI, a = j % NI, row = nb_plain + j;
  double acc[kMaxBorder + 1];
#pragma unroll
  for (int q = 0; q <= kMaxBorder; ++q) acc[q] = 0.0;
  for (int k = grp_off[grp]; k < grp_off[grp + 1]; ++k) {
    const double* rp = rowpart + ((size_t)grp_view[k] * NI + a) * (nb + 1);
#pragma unroll
    for (int q = 0; q < kMaxBorder; ++q)
      if (q < nb) acc[q] += rp[q];
    acc[kMaxBorder] += rp[nb];
  }
  double d;
  if (refresh_diag) { d = fmin(fmax(sh_h[j], min_diag), max_diag); diag_b[row] = d; }
  else d = diag_b[row];
#pragma unroll
  for (int q = 0; q < kMaxBorder; ++q) {
    if (q >= nb) continue;
    const double val = acc[q] + (q == row ? d / mu : 0.0);
    Sbb[row * nb + q] = val;
    if (q < nb_plain) Sbb[q * nb + row] = val;
  }
  rhs_b[row] = acc[kMaxBorder];
}
// the solved step of every shared unknown back into the per-view slots of its group
template <int NCL>
__global__ void k_shared_expand(int V, const int* __restrict__ grp_of, int ncam, int bo_sh, double* __restrict__ y) {
  constexpr int NI = NCL - 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = i / NI, c = i % NI;
  if (v >= V || grp_of[v] < 0) return;
  y[v * NCL + c] = y[ncam + bo_sh + grp_of[v] * NI + c];
}

// ---------------------------------------------------------------------------------------------------------------------
// PTZRayDistDisp (ptzray_optimizer.cc:202-264): the global disp[3] block is touched by EVERY ray observation.  It lives in
// the border; its per-observation Jacobian Fd (2x3, sqrt(w)- and Jacobi-scaled) is kept in recd[M][6].  This factor type
// is not selected by the reference's apps and only meets small scenes, so these kernels are plain (one thread per view or
// per track, fixed-order loops), not tuned.
// ---------------------------------------------------------------------------------------------------------------------
// per view: coupling strip C[v][a][disp] += sum F^T Fd, per-view partials of Hdd = sum Fd^T Fd (6) and gd = sum Fd^T r (3)
template <int NCL>
__global__ void k_disp_view(int V, const int* __restrict__ view_off, const double* __restrict__ recA, const double* __restrict__ recF,
                            const double* __restrict__ recd, int nb, int bo_disp,
                            double* __restrict__ C, double* __restrict__ dpart) {
  typedef Dims<NCL> D;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double c[NCL * 3], h[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
  for (int i = 0; i < NCL * 3; ++i) c[i] = 0;
  for (int o = view_off[v]; o < view_off[v + 1]; ++o) {
    const double* rp = recA + (size_t)o * D::RA;   // r at rp[0..1]
    const double* fp = recF + (size_t)o * D::RF;   // F0[NCL], F1[NCL]
    const double* fd = recd + (size_t)o * 6;
    for (int a = 0; a < NCL; ++a)
      for (int j = 0; j < 3; ++j) c[a * 3 + j] += fp[a] * fd[j] + fp[NCL + a] * fd[3 + j];
    int k = 0;
    for (int i = 0; i < 3; ++i) {
      g[i] += fd[i] * rp[0] + fd[3 + i] * rp[1];
      for (int j = i; j < 3; ++j) h[k++] += fd[i] * fd[j] + fd[3 + i] * fd[3 + j];
    }
  }
  for (int a = 0; a < NCL; ++a)
    for (int j = 0; j < 3; ++j) C[((size_t)v * NCL + a) * nb + bo_disp + j] += c[a * 3 + j];
  for (int i = 0; i < 6; ++i) dpart[(size_t)v * 9 + i] = h[i];
  for (int i = 0; i < 3; ++i) dpart[(size_t)v * 9 + 6 + i] = g[i];
}
// one thread: sums the per-view partials in view order into the disp block of Hbb and gb
__global__ void k_disp_total(int V, const double* __restrict__ dpart, int nb, int bo_disp, const double* __restrict__ scale_b, double* __restrict__ Hbb,
                             double* __restrict__ gb, double* __restrict__ gabs_b) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int v = 0; v < V; ++v)
    for (int i = 0; i < 9; ++i) s[i] += dpart[(size_t)v * 9 + i];
  int k = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      Hbb[(bo_disp + i) * nb + bo_disp + j] += s[k];
      if (j != i) Hbb[(bo_disp + j) * nb + bo_disp + i] += s[k];
      ++k;
    }
  for (int i = 0; i < 3; ++i) {
    gb[bo_disp + i] += s[6 + i];
    gabs_b[bo_disp + i] = fabs(gb[bo_disp + i] / scale_b[bo_disp + i]);
  }
}
// per track (after k_track_factor): Wd = sum Fd^T E, Wdh = Wd L^-T (3x3), qd = Wdh t; Wdh[P][12]
template <int NCL>
__global__ void k_disp_track(int P, const int* __restrict__ t_off, const int* __restrict__ t_obs, const double* __restrict__ recA,
                             const double* __restrict__ recd, const double* __restrict__ Lt, double* __restrict__ Wdh) {
  typedef Dims<NCL> D;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double w[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = t_off[p]; i < t_off[p + 1]; ++i) {
    const int o = t_obs[i];
    const double* rp = recA + (size_t)o * D::RA;   // E at rp[2..7]: E0[3], E1[3]
    const double* fd = recd + (size_t)o * 6;
    for (int a = 0; a < 3; ++a)
      for (int j = 0; j < 3; ++j) w[a * 3 + j] += fd[a] * rp[2 + j] + fd[3 + a] * rp[5 + j];
  }
  const double* lt = Lt + (size_t)p * 10;
  double* out = Wdh + (size_t)p * 12;
  for (int a = 0; a < 3; ++a) {
    const double x0 = w[a * 3] / lt[0], x1 = (w[a * 3 + 1] - lt[1] * x0) / lt[2], x2 = (w[a * 3 + 2] - lt[3] * x0 - lt[4] * x1) / lt[5];
    out[a * 3] = x0; out[a * 3 + 1] = x1; out[a * 3 + 2] = x2;
    out[9 + a] = x0 * lt[6] + x1 * lt[7] + x2 * lt[8];
  }
}
// per view: Schur term of the camera-disp coupling  sum_o What_o Wdh_{p(o)}^T  (NCL x 3), subtracted from the strip's disp columns
// of the working copy Cw (= C at the last Jacobian evaluation)
template <int NCL>
__global__ void k_disp_schur_view(int V, const int* __restrict__ view_off, const int* __restrict__ o_track, const double* __restrict__ What,
                                  const double* __restrict__ Wdh, int nb, int bo_disp, int own, const double* __restrict__ C, double* __restrict__ Cw) {
  typedef Dims<NCL> D;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double c[NCL * 3];
  for (int i = 0; i < NCL * 3; ++i) c[i] = 0;
  for (int o = view_off[v]; o < view_off[v + 1]; ++o) {
    const double* w = What + (size_t)o * D::WS;
    const double* wd = Wdh + (size_t)o_track[o] * 12;
    for (int a = 0; a < NCL; ++a)
      for (int j = 0; j < 3; ++j) c[a * 3 + j] += w[3 * a] * wd[3 * j] + w[3 * a + 1] * wd[3 * j + 1] + w[3 * a + 2] * wd[3 * j + 2];
  }
  for (int a = 0; a < NCL; ++a)
    for (int j = 0; j < nb; ++j) {
      double val = own ? C[((size_t)v * NCL + a) * nb + j] : 0.0;  // (sharded problem: the ranks' pieces are summed by the all-reduce)
      if (j >= bo_disp && j < bo_disp + 3) val -= c[a * 3 + (j - bo_disp)];
      Cw[((size_t)v * NCL + a) * nb + j] = val;
    }
}
// one thread: disp block of S_bb -= sum_p Wdh Wdh^T, rhs_disp -= sum_p qd, in track order
__global__ void k_disp_schur_border(int P, const int* __restrict__ t_off, const double* __restrict__ Wdh, int nb, int bo_disp, double* __restrict__ Sbb,
                                    double* __restrict__ rhs_b) {
  // 9 threads: entry (i, j) of the 3x3 block; 3 more for the right-hand side; every sum in track order
  const int e = threadIdx.x;
  if (blockIdx.x != 0 || e >= 12) return;
  double sacc = 0;
  if (e < 9) {
    const int i = e / 3, j = e % 3;
    for (int p = 0; p < P; ++p) {
      if (t_off[p] == t_off[p + 1]) continue;
      const double* w = Wdh + (size_t)p * 12;
      sacc += w[3 * i] * w[3 * j] + w[3 * i + 1] * w[3 * j + 1] + w[3 * i + 2] * w[3 * j + 2];
    }
    Sbb[(bo_disp + i) * nb + bo_disp + j] -= sacc;
  } else {
    const int i = e - 9;
    for (int p = 0; p < P; ++p) {
      if (t_off[p] == t_off[p + 1]) continue;
      sacc += Wdh[(size_t)p * 12 + 9 + i];
    }
    rhs_b[bo_disp + i] -= sacc;
  }
}

}  // namespace ptz
