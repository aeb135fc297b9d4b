// reloc.cu — batched PTZ relocalisation: B independent KRTOptimizer problems (krt_optimizer.cc:257-404), ONE CTA EACH.
//
// A query is a 4..6-dof problem over N = 64..512 matches (run_ptz_reloc.cc:68-118).  The CTA is one warp: the matches
// are staged once into shared memory as (unit ray of the reference pixel, current pixel), every LM iteration is one
// pass over them with the normal equations (<= 21+6+1 doubles) butterfly-reduced across the warp, and the
// Levenberg–Marquardt logic of Ceres' TrustRegionMinimizer runs redundantly and identically in every lane, so there
// is no block barrier and no host round trip inside a solve.  Bound: the fp64 pipe, not HBM.
//
// DENSE_QR (krt_optimizer.cc:389) is replaced by a Cholesky solve of the Jacobi-scaled damped normal equations:
// same minimiser of |J y - r|^2 + |D y|^2.
#include <math.h>

#include "common.cuh"
#include "ptz_math.cuh"

#ifndef PTZ_RELOC_MINB
#define PTZ_RELOC_MINB 12 // resident warps (= CTAs) per SM the register budget of k_reloc is sized for
#endif
namespace ptz {

struct RelocArgs {
  int B;
  const int64_t* off;
  const float2* uv_ref; const float2* uv_cur;
  const double* ref_cam; const double* init_cam;
  const int64_t* pt_off; const float2* pt_uv; const double* pt_xyz;  // optional 2d-3d terms (Add2d3dConstraints)
  int max_iter; double max_reproj_error;
  ptz_solver_options opt;
  int smem_matches;  // matches that fit the dynamic shared memory of this launch; larger queries recompute from global
  double* cam; int* success; int* termination; int* num_iter; int* iterations; double* initial_cost; double* final_cost; double* final_rms;
  double* local15;
};

// all lanes end with the same NV warp totals: reduce-scatter (NP-1 shuffles) + one broadcast per value, instead of NV
// five-step butterflies; every lane receives the very same bits, which the lane-redundant LM logic below relies on
template <int NV>
__device__ __forceinline__ void warp_allreduce(double (&v)[NV], int lane) {
  constexpr int NP = NV <= 16 ? 16 : 32;
  static_assert(NV <= 32, "at most 32 sums");
  double t[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) t[i] = i < NV ? v[i] : 0.0;
  const double tot = warp_reduce_scatter<NP>(t, lane);
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __shfl_sync(0xffffffffu, tot, i * (32 / NP));
}

template <int TYPE>
__device__ __forceinline__ int krt_free_index(int j) {
  // F: fx,w | Fxfy: fx,fy,w | FDist: fx,w,k1 | FxfyDist: fx,fy,w,k1
  constexpr bool FXFY = (TYPE == KRT_FXFY || TYPE == KRT_FXFYDIST);
  if (j == 0) return 0;
  if (FXFY) { if (j == 1) return 1; if (j < 5) return 4 + (j - 2); return 10; }
  if (j < 4) return 4 + (j - 1);
  return 10;
}

// one pass over the query's matches at camera x: cost, and (when JAC) A = J^T J (upper, row-major), g = J^T r
template <int TYPE, bool JAC>
__device__ __forceinline__ void reloc_pass(const double* x /*[15], shared*/, int N, int lane, const double4* sm, int smem_matches, const float2* uv_ref,
                                           const float2* uv_cur, int npts, const float2* puv, const double* pxyz, const double* ref21,
                                           double* out /*[NA+NF+1]*/) {
  constexpr int NF = krt_nfree(TYPE), NA = NF * (NF + 1) / 2, NV = NA + NF + 1;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  KrtCam kc;
  krt_make_cam<TYPE>(x, &kc, JAC);
  for (int i = lane; i < N; i += 32) {
    double n[3];
    float2 uv2;
    bool ok;
    if (i < smem_matches) {
      const double4 m = sm[i];
      n[0] = m.x; n[1] = m.y; n[2] = m.z;
      uv2 = *reinterpret_cast<const float2*>(&m.w);
      ok = !isnan(uv2.x);
    } else {
      double refK4[4], refd[5];  // beyond the staging cap (rare): recompute from global
      for (int j = 0; j < 4; ++j) refK4[j] = ref21[j];
      for (int j = 0; j < 5; ++j) refd[j] = ref21[16 + j];
      const float2 u1 = uv_ref[i];
      ok = krt_precompute(TYPE, refK4, refd, u1.x, u1.y, n);
      uv2 = uv_cur[i];
    }
    if (!ok) continue;  // masked: residual identically 0 (krt_optimizer.cc:95-101)
    double r[2], J[2 * NF];
    krt_obs<TYPE, JAC>(kc, n, (double)uv2.x, (double)uv2.y, r, J);
    acc[NA + NF] += 0.5 * (r[0] * r[0] + r[1] * r[1]);
    if (JAC) {
      int k = 0;
#pragma unroll
      for (int a = 0; a < NF; ++a)
#pragma unroll
        for (int b = a; b < NF; ++b) acc[k++] += J[a] * J[b] + J[NF + a] * J[NF + b];
#pragma unroll
      for (int a = 0; a < NF; ++a) acc[NA + a] += J[a] * r[0] + J[NF + a] * r[1];
    }
  }
  // 2d-3d terms: the world point goes to the reference-local frame (R_ref X + t_ref, krt_optimizer.cc:357-362), then
  // Factor2d3dDist / Factor2d3dFxfyDist with the camera's own t
  for (int i = lane; i < npts; i += 32) {
    const double* Rr = ref21 + 4;
    const double Xw[3] = {pxyz[3 * i], pxyz[3 * i + 1], pxyz[3 * i + 2]};
    double P[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) P[a] = Rr[3 * a] * Xw[0] + Rr[3 * a + 1] * Xw[1] + Rr[3 * a + 2] * Xw[2] + ref21[13 + a];
    double r[2], J[2 * NF];
    krt_obs3d<TYPE, JAC>(kc, x + 7, P, (double)puv[i].x, (double)puv[i].y, r, J);
    acc[NA + NF] += 0.5 * (r[0] * r[0] + r[1] * r[1]);
    if (JAC) {
      int k = 0;
#pragma unroll
      for (int a = 0; a < NF; ++a)
#pragma unroll
        for (int b = a; b < NF; ++b) acc[k++] += J[a] * J[b] + J[NF + a] * J[NF + b];
#pragma unroll
      for (int a = 0; a < NF; ++a) acc[NA + a] += J[a] * r[0] + J[NF + a] * r[1];
    }
  }
  warp_allreduce<NV>(acc, lane);
#pragma unroll
  for (int i = 0; i < NV; ++i) out[i] = acc[i];
}

template <int TYPE>
__global__ void __launch_bounds__(32, PTZ_RELOC_MINB) k_reloc(RelocArgs a) {
  constexpr int NF = krt_nfree(TYPE), NA = NF * (NF + 1) / 2, NV = NA + NF + 1;
  extern __shared__ double4 sm[];
  const int q = blockIdx.x, lane = threadIdx.x;
  if (q >= a.B) return;
  const int64_t o0 = a.off[q];
  const int N = (int)(a.off[q + 1] - o0);
  const double* ref = a.ref_cam + 21 * (size_t)q;
  const float2* uv_ref = a.uv_ref + o0;
  const float2* uv_cur = a.uv_cur + o0;
  const int64_t p0 = a.pt_off ? a.pt_off[q] : 0;
  const int npts = a.pt_off ? (int)(a.pt_off[q + 1] - p0) : 0;
  const float2* puv = a.pt_uv + p0;
  const double* pxyz = a.pt_xyz + 3 * p0;
  // stage the parameter-independent part of every match
  {
    double refK4[4], refd[5];
    for (int j = 0; j < 4; ++j) refK4[j] = ref[j];
    for (int j = 0; j < 5; ++j) refd[j] = ref[16 + j];
    for (int i = lane; i < N && i < a.smem_matches; i += 32) {
      double n[3];
      const float2 u1 = uv_ref[i];
      const bool ok = krt_precompute(TYPE, refK4, refd, u1.x, u1.y, n);
      float2 u2 = uv_cur[i];
      if (!ok) u2.x = nanf("");
      double4 m;
      m.x = n[0]; m.y = n[1]; m.z = n[2];
      *reinterpret_cast<float2*>(&m.w) = u2;
      sm[i] = m;
    }
  }
  // The LM state is the same in every lane; it lives in shared memory (one copy per warp, broadcast reads) so that the
  // registers of the match loop are not shared with it.  Every lane stores the same bits; __syncwarp orders them.
  __shared__ double x[15], cand[15], A[NA], g[NF], scale[NF], diag[NF];
  {
    double x0[15];
    krt_to_local(ref, a.init_cam + 21 * (size_t)q, x0);  // Add2d2dConstraints: reference-local frame
    if (lane == 0) for (int j = 0; j < 15; ++j) x[j] = x0[j];
  }
  __syncwarp();
  const ptz_solver_options& o = a.opt;
  // ---- TrustRegionMinimizer (Ceres 1.14), dense normal equations
  double ev[NV];
  reloc_pass<TYPE, true>(x, N, lane, sm, a.smem_matches, uv_ref, uv_cur, npts, puv, pxyz, ref, ev);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) A[i] = ev[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) g[i] = ev[NA + i];
    int k = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) { scale[i] = o.jacobi_scaling ? 1.0 / (1.0 + sqrt(ev[k])) : 1.0; k += NF - i; }
  }
  __syncwarp();
  double x_cost = ev[NA + NF];
  const double initial_cost = x_cost;
  double min_cost = x_cost;
  double x_norm = 0;
  for (int j = 0; j < 15; ++j) x_norm += x[j] * x[j];
  x_norm = sqrt(x_norm);
  double grad_max = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) grad_max = fmax(grad_max, fabs(g[i]));
  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false, last_successful = true;
  int iteration = 0, logged = 0, num_invalid = 0, num_successful = 1, termination = PTZ_NO_CONVERGENCE;
  const int max_iter = a.max_iter;
  while (true) {
    if (iteration >= max_iter) { termination = PTZ_NO_CONVERGENCE; break; }
    if (last_successful && grad_max <= o.gradient_tolerance) { termination = PTZ_CONVERGENCE; break; }
    if (radius <= o.min_trust_region_radius) { termination = PTZ_CONVERGENCE; break; }
    ++iteration;
    // scaled system  As = s A s,  gs = s g ; D^2 = clamp(diag As) / radius
    double L[NF * NF], y[NF];
    {
      int k = 0;
#pragma unroll
      for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int j = i; j < NF; ++j) { const double v = A[k++] * scale[i] * scale[j]; L[i * NF + j] = v; L[j * NF + i] = v; }
    }
    if (!reuse_diagonal) {
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NF; ++i) diag[i] = fmin(fmax(L[i * NF + i], o.min_lm_diagonal), o.max_lm_diagonal);
      }
      __syncwarp();
    }
    double ytAy = 0, ygs = 0;
    double As[NF * NF];
#pragma unroll
    for (int i = 0; i < NF * NF; ++i) As[i] = L[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) { L[i * NF + i] += diag[i] / radius; y[i] = g[i] * scale[i]; }
    reuse_diagonal = true;
    bool ok = chol_n(L, NF, NF);
    if (ok) {
      chol_solve_n(L, NF, NF, y);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        if (!isfinite(y[i])) ok = false;
        double s = 0;
#pragma unroll
        for (int j = 0; j < NF; ++j) s += As[i * NF + j] * y[j];
        ytAy += y[i] * s;
        ygs += y[i] * g[i] * scale[i];
      }
    }
    // model_cost_change = -(J d)^T (r + J d / 2) with d = -y
    const double model_cost_change = ygs - 0.5 * ytAy;
    if (ok) ok = model_cost_change > 0.0;
    if (!ok) {
      ++num_invalid;
      if (num_invalid >= o.max_num_consecutive_invalid_steps) { termination = PTZ_FAILURE; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      last_successful = false;
      ++logged;
      continue;
    }
    num_invalid = 0;
    double step2 = 0;
    {
      double c[15];
#pragma unroll
      for (int j = 0; j < 15; ++j) c[j] = x[j];
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int idx = krt_free_index<TYPE>(i);
        c[idx] = x[idx] + (-y[i] * scale[i]);
        step2 += (x[idx] - c[idx]) * (x[idx] - c[idx]);
      }
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 15; ++j) cand[j] = c[j];
      }
      __syncwarp();
    }
    // candidate cost together with its normal equations (they are needed as soon as the step is accepted)
    reloc_pass<TYPE, true>(cand, N, lane, sm, a.smem_matches, uv_ref, uv_cur, npts, puv, pxyz, ref, ev);
    double cand_cost = ev[NA + NF];
    if (!isfinite(cand_cost)) cand_cost = 1.7976931348623157e308;
    const double step_norm = sqrt(step2);
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { termination = PTZ_CONVERGENCE; break; }
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= o.function_tolerance * x_cost) { termination = PTZ_CONVERGENCE; break; }
    const double rho = cost_change / model_cost_change;
    if (rho > o.min_relative_decrease) {
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 15; ++j) x[j] = cand[j];
#pragma unroll
        for (int i = 0; i < NA; ++i) A[i] = ev[i];
#pragma unroll
        for (int i = 0; i < NF; ++i) g[i] = ev[NA + i];
      }
      __syncwarp();
      x_norm = 0;
      for (int j = 0; j < 15; ++j) x_norm += x[j] * x[j];
      x_norm = sqrt(x_norm);
      grad_max = 0;
#pragma unroll
      for (int i = 0; i < NF; ++i) grad_max = fmax(grad_max, fabs(g[i]));
      x_cost = cand_cost;
      radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3.0));
      radius = fmin(o.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      last_successful = true;
      ++num_successful;
      if (x_cost < min_cost) min_cost = x_cost;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      last_successful = false;
    }
    ++logged;  // rows of Ceres' iteration table: the iteration that trips a tolerance is not recorded
  }
  // CheckResults (krt_optimizer.cc:504-533) and ObtainRefinedCameraParams (:535-567)
  if (lane == 0) {
    const double nres = 2.0 * (N + npts);
    const double final_rms = sqrt(2.0) * sqrt((2 * min_cost) / nres);
    int ok = 1;
    if (termination != PTZ_CONVERGENCE) ok = 0;
    else if (final_rms >= a.max_reproj_error) ok = 0;
    if (ok) {
      const double fov_x = atan(x[2] / x[0]) * 2 * 180 / M_PI, fov_y = atan(x[3] / x[1]) * 2 * 180 / M_PI;
      if (fov_x < 0 || fov_x > 170 || fov_y < 0 || fov_y > 170) ok = 0;
    }
    a.success[q] = ok;
    a.termination[q] = termination;
    a.num_iter[q] = num_successful;
    if (a.iterations) a.iterations[q] = logged;
    if (a.initial_cost) a.initial_cost[q] = initial_cost;
    if (a.final_cost) a.final_cost[q] = min_cost;
    if (a.final_rms) a.final_rms[q] = final_rms;
    if (a.local15) for (int j = 0; j < 15; ++j) a.local15[15 * (size_t)q + j] = x[j];
    double w[21];
    krt_to_world(TYPE, ref, x, w);
    for (int j = 0; j < 21; ++j) a.cam[21 * (size_t)q + j] = w[j];
  }
}

// stage-level hook: per-match residuals and Jacobian rows of one query (one thread per match)
template <int TYPE>
__global__ void k_reloc_eval(int N, const float2* uv_ref, const float2* uv_cur, const double* ref21, const double* local15, double* res, double* jac) {
  constexpr int NF = krt_nfree(TYPE);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double refK4[4], refd[5], x[15];
  for (int j = 0; j < 4; ++j) refK4[j] = ref21[j];
  for (int j = 0; j < 5; ++j) refd[j] = ref21[16 + j];
  for (int j = 0; j < 15; ++j) x[j] = local15[j];
  KrtCam kc;
  krt_make_cam<TYPE>(x, &kc, true);
  double n[3], r[2] = {0, 0}, J[2 * NF];
  for (int j = 0; j < 2 * NF; ++j) J[j] = 0;
  if (krt_precompute(TYPE, refK4, refd, uv_ref[i].x, uv_ref[i].y, n)) krt_obs<TYPE, true>(kc, n, (double)uv_cur[i].x, (double)uv_cur[i].y, r, J);
  res[2 * i] = r[0]; res[2 * i + 1] = r[1];
  for (int j = 0; j < 2 * NF; ++j) jac[(size_t)i * 2 * NF + j] = J[j];
}

// KRTOptimizer::Cal2d2dReprojError / Cal2d3dReprojError (krt_optimizer.cc:406-500): sums of squared functor residuals at the
// given local parameters, one CTA; out[0] = sum over the matches, out[1] = sum over the 2d-3d points
template <int TYPE>
__global__ void __launch_bounds__(256) k_reloc_reproj(int N, const float2* __restrict__ uv_ref, const float2* __restrict__ uv_cur, int npts,
                                                      const float2* __restrict__ puv, const double* __restrict__ pxyz, const double* __restrict__ ref21,
                                                      const double* __restrict__ local15, double* __restrict__ out) {
  __shared__ double sred[2 * 8];
  double refK4[4], refd[5], x[15];
  for (int j = 0; j < 4; ++j) refK4[j] = ref21[j];
  for (int j = 0; j < 5; ++j) refd[j] = ref21[16 + j];
  for (int j = 0; j < 15; ++j) x[j] = local15[j];
  KrtCam kc;
  krt_make_cam<TYPE>(x, &kc, false);
  double acc[2] = {0, 0};
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double n[3], r[2] = {0, 0};
    if (krt_precompute(TYPE, refK4, refd, uv_ref[i].x, uv_ref[i].y, n)) krt_obs<TYPE, false>(kc, n, (double)uv_cur[i].x, (double)uv_cur[i].y, r, nullptr);
    acc[0] += r[0] * r[0] + r[1] * r[1];
  }
  for (int i = threadIdx.x; i < npts; i += blockDim.x) {
    const double* Rr = ref21 + 4;
    double P[3], r[2];
    for (int a = 0; a < 3; ++a) P[a] = Rr[3 * a] * pxyz[3 * i] + Rr[3 * a + 1] * pxyz[3 * i + 1] + Rr[3 * a + 2] * pxyz[3 * i + 2] + ref21[13 + a];
    krt_obs3d<TYPE, false>(kc, x + 7, P, (double)puv[i].x, (double)puv[i].y, r, nullptr);
    acc[1] += r[0] * r[0] + r[1] * r[1];
  }
  block_sum<2>(acc, sred);
  if (threadIdx.x == 0) { out[0] = acc[0]; out[1] = acc[1]; }
}

static void launch_reloc(int type, const RelocArgs& a, int max_matches, cudaStream_t s) {
  // dynamic shared memory: 32 B per staged match, capped so that several CTAs still share an SM
  const int cap_matches = 2048;
  RelocArgs b = a;
  b.smem_matches = std::min(max_matches, cap_matches);
  const size_t smem = (size_t)std::max(b.smem_matches, 1) * sizeof(double4);
  auto go = [&](auto kern) {
    PTZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<a.B, 32, smem, s>>>(b);
    PTZ_CUDA(cudaGetLastError());
  };
  switch (type) {
    case KRT_F: go(k_reloc<KRT_F>); break;
    case KRT_FDIST: go(k_reloc<KRT_FDIST>); break;
    case KRT_FXFY: go(k_reloc<KRT_FXFY>); break;
    default: go(k_reloc<KRT_FXFYDIST>); break;
  }
}

// fp64 FMA throughput of the device, measured: the denominator of the reloc kernel's roofline (it is bound by the fp64 pipe,
// not by HBM).  8 independent chains per thread, every SM full.
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double* __restrict__ out) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

template <class F>
static int guarded_r(F&& f) {
  try {
    return f();
  } catch (const CudaError& e) {
    set_last_error("%s", e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error("%s", e.what());
    return PTZ_ERR_CUDA;
  }
}

}  // namespace ptz

using namespace ptz;

extern "C" {

int ptzreloc_solve_batch_dev(const ptzreloc_batch* b, const ptz_solver_options* opt, ptzreloc_result* out, void* cuda_stream) {
  if (!b || !opt || !out || b->num_queries < 0 || b->factor_type < 0 || b->factor_type > 3) return PTZ_ERR_INVALID;
  if (b->num_queries == 0) return PTZ_OK;
  return guarded_r([&]() {
    RelocArgs a;
    a.B = b->num_queries; a.off = b->match_offset;
    a.uv_ref = reinterpret_cast<const float2*>(b->uv_ref); a.uv_cur = reinterpret_cast<const float2*>(b->uv_cur);
    a.ref_cam = b->ref_cam; a.init_cam = b->init_cam; a.max_iter = b->max_iter; a.max_reproj_error = b->max_reproj_error;
    a.pt_off = b->pt_offset; a.pt_uv = reinterpret_cast<const float2*>(b->pt_uv); a.pt_xyz = b->pt_xyz;
    a.opt = *opt;
    a.cam = out->cam; a.success = out->success; a.termination = out->termination; a.num_iter = out->num_iter; a.iterations = out->iterations;
    a.initial_cost = out->initial_cost; a.final_cost = out->final_cost; a.final_rms = out->final_rms; a.local15 = out->local_cam15;
    // the device-resident variant cannot look at the offsets: size shared memory for the common case, larger queries recompute
    launch_reloc(b->factor_type, a, 512, (cudaStream_t)cuda_stream);
    return (int)PTZ_OK;
  });
}

int ptzreloc_solve_batch(const ptzreloc_batch* b, const ptz_solver_options* opt, ptzreloc_result* out) {
  if (!b || !opt || !out || b->num_queries < 0 || b->factor_type < 0 || b->factor_type > 3) return PTZ_ERR_INVALID;
  if (b->num_queries > 0 && (!b->match_offset || !b->ref_cam || !b->init_cam || !out->cam || !out->success || !out->termination || !out->num_iter))
    return PTZ_ERR_INVALID;
  if (b->num_queries == 0) return PTZ_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_last_error("no CUDA device; this library has no CPU fallback"); return PTZ_ERR_NO_DEVICE; }
  return guarded_r([&]() {
    const int B = b->num_queries;
    const int64_t N = b->match_offset[B];
    int max_matches = 0;
    for (int q = 0; q < B; ++q) {
      int64_t n = b->match_offset[q + 1] - b->match_offset[q];
      if (n < 0) throw CudaError(PTZ_ERR_INVALID, "match_offset is not monotone");
      max_matches = std::max<int64_t>(max_matches, n);
    }
    enable_memory_pool();
    StreamHolder sh;  // declared before the buffers: destroyed after them
    sh.create();
    cudaStream_t s = sh.s;
    DevBuf<int64_t> d_off, d_poff;
    DevBuf<float2> d_ur, d_uc, d_puv;
    DevBuf<double> d_pxyz;
    const bool has_pts = b->pt_offset && b->pt_uv && b->pt_xyz;
    if (has_pts) {
      const int64_t Np = b->pt_offset[B];
      d_poff.upload(b->pt_offset, B + 1, s);
      d_puv.upload(reinterpret_cast<const float2*>(b->pt_uv), Np, s);
      d_pxyz.upload(b->pt_xyz, 3 * (size_t)Np, s);
      if (Np == 0) { d_puv.alloc(1, s); d_pxyz.alloc(1, s); }
    }
    DevBuf<double> d_ref, d_init, d_cam, d_ic, d_fc, d_rms, d_loc;
    DevBuf<int> d_succ, d_term, d_ni, d_it;
    d_off.upload(b->match_offset, B + 1, s);
    d_ur.alloc(N, s); d_uc.alloc(N, s); d_ref.alloc(21 * (size_t)B, s); d_init.alloc(21 * (size_t)B, s);
    d_cam.alloc(21 * (size_t)B, s); d_ic.alloc(B, s); d_fc.alloc(B, s); d_rms.alloc(B, s); d_loc.alloc(15 * (size_t)B, s);
    d_succ.alloc(B, s); d_term.alloc(B, s); d_ni.alloc(B, s); d_it.alloc(B, s);
    // The batch is PCIe-bound end to end (16 B per match in, the kernel reads them once): cut it into slices of ~1 M matches and
    // rotate them over three streams, so that the copy-in of slice k+1 and the copy-out of slice k-1 run under the kernel of slice k
    // (the copies only overlap when the caller's buffers are pinned; with pageable memory this degrades to the serial order).
    const int64_t kSliceMatches = 1 << 20;
    const int nslices = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / kSliceMatches));
    StreamHolder extra[2];
    cudaStream_t st[3] = {s, s, s};
    cudaEvent_t ready = nullptr, done[2] = {nullptr, nullptr};
    if (nslices > 1) {
      extra[0].create(); extra[1].create();
      st[1] = extra[0].s; st[2] = extra[1].s;
      PTZ_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
      PTZ_CUDA(cudaEventRecord(ready, s));  // allocations and the offsets are stream-ordered on s
      for (int k = 0; k < 2; ++k) { PTZ_CUDA(cudaStreamWaitEvent(st[k + 1], ready, 0)); PTZ_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming)); }
    }
    int q0 = 0;
    for (int k = 0; k < nslices; ++k) {
      int q1 = B;
      if (k + 1 < nslices) {  // first query whose matches start at or behind the slice's share
        const int64_t target = N * (int64_t)(k + 1) / nslices;
        q1 = (int)(std::lower_bound(b->match_offset + q0, b->match_offset + B, target) - b->match_offset);
        q1 = std::max(q1, q0);
      }
      const int nq = q1 - q0;
      if (nq == 0) continue;
      cudaStream_t cs = st[k % 3];
      const int64_t m0 = b->match_offset[q0], m1 = b->match_offset[q1];
      if (m1 > m0) {
        PTZ_CUDA(cudaMemcpyAsync(d_ur.p + m0, reinterpret_cast<const float2*>(b->uv_ref) + m0, (size_t)(m1 - m0) * sizeof(float2), cudaMemcpyHostToDevice, cs));
        PTZ_CUDA(cudaMemcpyAsync(d_uc.p + m0, reinterpret_cast<const float2*>(b->uv_cur) + m0, (size_t)(m1 - m0) * sizeof(float2), cudaMemcpyHostToDevice, cs));
      }
      PTZ_CUDA(cudaMemcpyAsync(d_ref.p + 21 * (size_t)q0, b->ref_cam + 21 * (size_t)q0, 21 * (size_t)nq * 8, cudaMemcpyHostToDevice, cs));
      PTZ_CUDA(cudaMemcpyAsync(d_init.p + 21 * (size_t)q0, b->init_cam + 21 * (size_t)q0, 21 * (size_t)nq * 8, cudaMemcpyHostToDevice, cs));
      RelocArgs a;
      a.B = nq; a.off = d_off.p + q0; a.uv_ref = d_ur.p; a.uv_cur = d_uc.p; a.ref_cam = d_ref.p + 21 * (size_t)q0; a.init_cam = d_init.p + 21 * (size_t)q0;
      a.max_iter = b->max_iter; a.max_reproj_error = b->max_reproj_error; a.opt = *opt;
      a.pt_off = has_pts ? d_poff.p + q0 : nullptr; a.pt_uv = d_puv.p; a.pt_xyz = d_pxyz.p;
      a.cam = d_cam.p + 21 * (size_t)q0; a.success = d_succ.p + q0; a.termination = d_term.p + q0; a.num_iter = d_ni.p + q0; a.iterations = d_it.p + q0;
      a.initial_cost = d_ic.p + q0; a.final_cost = d_fc.p + q0; a.final_rms = d_rms.p + q0; a.local15 = d_loc.p + 15 * (size_t)q0;
      launch_reloc(b->factor_type, a, max_matches, cs);
      q0 = q1;
    }
    if (nslices > 1) {  // the main stream (on which the buffers are released) waits for the other two
      for (int k = 0; k < 2; ++k) { PTZ_CUDA(cudaEventRecord(done[k], st[k + 1])); PTZ_CUDA(cudaStreamWaitEvent(s, done[k], 0)); }
    }
    // results (small next to the matches) leave at the end: a copy into PAGEABLE caller memory blocks the host until it is done, and
    // issued per slice it would stall the loop above behind every kernel
    d_cam.download(out->cam, 21 * (size_t)B, s);
    d_succ.download(out->success, B, s);
    d_term.download(out->termination, B, s);
    d_ni.download(out->num_iter, B, s);
    if (out->iterations) d_it.download(out->iterations, B, s);
    if (out->initial_cost) d_ic.download(out->initial_cost, B, s);
    if (out->final_cost) d_fc.download(out->final_cost, B, s);
    if (out->final_rms) d_rms.download(out->final_rms, B, s);
    if (out->local_cam15) d_loc.download(out->local_cam15, 15 * (size_t)B, s);
    PTZ_CUDA(cudaStreamSynchronize(s));
    if (ready) cudaEventDestroy(ready);
    for (int k = 0; k < 2; ++k) if (done[k]) cudaEventDestroy(done[k]);
    return (int)PTZ_OK;
  });
}


int ptzreloc_eval(int type, int N, const float* uv_ref, const float* uv_cur, const double* ref21, const double* local15, double* residuals, double* jac,
                  double* cost, double* gradient) {
  if (type < 0 || type > 3 || N < 0 || !ref21 || !local15) return PTZ_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_last_error("no CUDA device; this library has no CPU fallback"); return PTZ_ERR_NO_DEVICE; }
  return guarded_r([&]() {
    const int nf = krt_nfree(type);
    cudaStream_t s = 0;
    DevBuf<float2> d_ur, d_uc;
    DevBuf<double> d_ref, d_x, d_res, d_jac;
    d_ur.upload(reinterpret_cast<const float2*>(uv_ref), N, s);
    d_uc.upload(reinterpret_cast<const float2*>(uv_cur), N, s);
    d_ref.upload(ref21, 21, s);
    d_x.upload(local15, 15, s);
    d_res.alloc(2 * (size_t)std::max(N, 1));
    d_jac.alloc(2 * (size_t)std::max(N, 1) * nf);
    if (N > 0) {
      const int g = (N + 127) / 128;
      switch (type) {
        case KRT_F: k_reloc_eval<KRT_F><<<g, 128, 0, s>>>(N, d_ur.p, d_uc.p, d_ref.p, d_x.p, d_res.p, d_jac.p); break;
        case KRT_FDIST: k_reloc_eval<KRT_FDIST><<<g, 128, 0, s>>>(N, d_ur.p, d_uc.p, d_ref.p, d_x.p, d_res.p, d_jac.p); break;
        case KRT_FXFY: k_reloc_eval<KRT_FXFY><<<g, 128, 0, s>>>(N, d_ur.p, d_uc.p, d_ref.p, d_x.p, d_res.p, d_jac.p); break;
        default: k_reloc_eval<KRT_FXFYDIST><<<g, 128, 0, s>>>(N, d_ur.p, d_uc.p, d_ref.p, d_x.p, d_res.p, d_jac.p); break;
      }
      PTZ_CUDA(cudaGetLastError());
    }
    std::vector<double> r(2 * (size_t)N), J(2 * (size_t)N * nf);
    d_res.download(r.data(), r.size(), s);
    d_jac.download(J.data(), J.size(), s);
    PTZ_CUDA(cudaStreamSynchronize(s));
    double c = 0;
    std::vector<double> g(nf, 0.0);
    for (int i = 0; i < N; ++i) {
      c += 0.5 * (r[2 * i] * r[2 * i] + r[2 * i + 1] * r[2 * i + 1]);
      for (int j = 0; j < nf; ++j) g[j] += J[(size_t)i * 2 * nf + j] * r[2 * i] + J[(size_t)i * 2 * nf + nf + j] * r[2 * i + 1];
    }
    if (residuals) memcpy(residuals, r.data(), r.size() * 8);
    if (jac) memcpy(jac, J.data(), J.size() * 8);
    if (cost) *cost = c;
    if (gradient) memcpy(gradient, g.data(), nf * 8);
    return (int)PTZ_OK;
  });
}

int ptz_measure_fp64_gflops(double* gflops) {
  if (!gflops) return PTZ_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_last_error("no CUDA device"); return PTZ_ERR_NO_DEVICE; }
  return guarded_r([&]() {
    int dev = 0, sms = 0;
    PTZ_CUDA(cudaGetDevice(&dev));
    PTZ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DevBuf<double> d;
    d.alloc(1);
    cudaEvent_t e0, e1;
    PTZ_CUDA(cudaEventCreate(&e0));
    PTZ_CUDA(cudaEventCreate(&e1));
    const int iters = 1 << 14, grid = sms * 8;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
      PTZ_CUDA(cudaEventRecord(e0, 0));
      k_fp64_peak<<<grid, 256>>>(iters, 1.0, d.p);
      PTZ_CUDA(cudaEventRecord(e1, 0));
      PTZ_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      PTZ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double gf = 2.0 * 8.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e9;
      if (rep > 0 && gf > best) best = gf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gflops = best;
    return (int)PTZ_OK;
  });
}

int ptzreloc_local_params(const double* ref21, const double* init21, double* local15) {
  if (!ref21 || !init21 || !local15) return PTZ_ERR_INVALID;
  krt_to_local(ref21, init21, local15);  // host instantiation of the routine the kernel runs per query
  return PTZ_OK;
}

int ptzreloc_reproj_error(int type, const double* ref21, const double* local15, int N, const float* uv_ref, const float* uv_cur, int npts,
                          const float* pt_uv, const double* pt_xyz, double* err_2d2d, double* err_2d3d) {
  if (type < 0 || type > 3 || N < 0 || npts < 0 || !ref21 || !local15) return PTZ_ERR_INVALID;
  if ((N > 0 && (!uv_ref || !uv_cur)) || (npts > 0 && (!pt_uv || !pt_xyz))) return PTZ_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_last_error("no CUDA device; this library has no CPU fallback"); return PTZ_ERR_NO_DEVICE; }
  return guarded_r([&]() {
    cudaStream_t s = 0;
    DevBuf<float2> d_ur, d_uc, d_pu;
    DevBuf<double> d_ref, d_x, d_px, d_out;
    d_ur.upload(reinterpret_cast<const float2*>(uv_ref), N, s);
    d_uc.upload(reinterpret_cast<const float2*>(uv_cur), N, s);
    d_pu.upload(reinterpret_cast<const float2*>(pt_uv), npts, s);
    d_px.upload(pt_xyz, 3 * (size_t)npts, s);
    d_ref.upload(ref21, 21, s);
    d_x.upload(local15, 15, s);
    d_out.alloc(2);
    switch (type) {
      case KRT_F: k_reloc_reproj<KRT_F><<<1, 256, 0, s>>>(N, d_ur.p, d_uc.p, npts, d_pu.p, d_px.p, d_ref.p, d_x.p, d_out.p); break;
      case KRT_FDIST: k_reloc_reproj<KRT_FDIST><<<1, 256, 0, s>>>(N, d_ur.p, d_uc.p, npts, d_pu.p, d_px.p, d_ref.p, d_x.p, d_out.p); break;
      case KRT_FXFY: k_reloc_reproj<KRT_FXFY><<<1, 256, 0, s>>>(N, d_ur.p, d_uc.p, npts, d_pu.p, d_px.p, d_ref.p, d_x.p, d_out.p); break;
      default: k_reloc_reproj<KRT_FXFYDIST><<<1, 256, 0, s>>>(N, d_ur.p, d_uc.p, npts, d_pu.p, d_px.p, d_ref.p, d_x.p, d_out.p); break;
    }
    PTZ_CUDA(cudaGetLastError());
    double h[2] = {0, 0};
    d_out.download(h, 2, s);
    PTZ_CUDA(cudaStreamSynchronize(s));
    // sqrt(sum / count): NaN for an empty match list, as the reference's 0/0; -1 without points (krt_optimizer.cc:459-460)
    if (err_2d2d) *err_2d2d = sqrt(h[0] / (double)N);
    if (err_2d3d) *err_2d3d = npts > 0 ? sqrt(h[1] / (double)npts) : -1.0;
    return (int)PTZ_OK;
  });
}

}  // extern "C"
