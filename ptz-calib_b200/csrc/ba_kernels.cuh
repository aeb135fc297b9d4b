// ba_kernels.cuh — the kernels of one Levenberg–Marquardt iteration of PTZ bundle adjustment (fp64, sm_100a).
//
// Stage 1  k_view_prep, k_resjac, k_view_finalize, k_track_accum      residual + analytic Jacobian, block reductions
// Stage 2  k_track_solve, k_schur_diag, k_schur_offdiag, k_precond    ray elimination, Schur complement onto cameras
// Stage 3  k_pcg                                                      block-Jacobi PCG on the reduced camera system
// Stage 4  k_track_backsub, k_cam_update, k_cost, k_scalars           back-substitution, damping bookkeeping, cost
//
// They replace what ceres::Solve does per iteration behind ptzray_optimizer.cc:475.
// Memory-bound small-block work: no tensor cores; layouts are chosen so that every pass streams contiguous,
// 16-byte-aligned records and the per-view tables are CTA-uniform.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "ptz_math.cuh"

namespace ptz {
namespace cg = cooperative_groups;

constexpr int kChunk = 128;      // observations per CTA in the streaming passes; chunks never straddle a view
constexpr int kMaxBorder = 16;   // dense border of the reduced system: tlw(6) + disp(3) + shared intrinsics unknowns; one warp owns its
                                 // row in the CG (the fy of annotated views are eliminated before the CG, ba_border.cuh)

template <int NCL>
struct Dims {
  static constexpr int RS = 8 + 2 * NCL;            // record: r(2) E(6) | F(2*NCL), kept in TWO arrays:     [doubles]
  static constexpr int RA = 8, RF = 2 * NCL;        //   recA[M][8] = r, E (what the by-track passes gather: 64 bytes, never half a line
                                                    //   of something else) and recF[M][2 NCL] = F
  static constexpr int WS = (3 * NCL + 3) & ~3;     // What record: NCL x 3, padded to whole 32-byte pieces (256-bit gathers)
  static constexpr int NU = NCL * (NCL + 1) / 2;    // upper triangle of a camera block
  static constexpr int NPART = NU + NCL + 1;        // chunk partial: U upper, g, cost
};

// per-track parameter record: ray(3), sqrt(weight), Jacobi scale(3), pad
constexpr int kTrk = 8;

// position of 16-byte chunk p of thread t's record inside its shared-memory slot.  Records of 8 chunks (NCL=4) would put
// chunk p of every thread in the same bank group: rotate by the thread index.  9- and 10-chunk records spread by themselves.
template <int NCL>
__device__ __forceinline__ int rec_swz(int t, int p) { return NCL == 4 ? (p ^ (t & 7)) : p; }
// same rule keyed on the number C of 16-byte chunks per record
template <int C>
__device__ __forceinline__ int chunk_swz(int t, int p) { return C == 8 ? (p ^ (t & 7)) : p; }

// -------------------------------------------------------------------------------------------------------------
// stage 1
// -------------------------------------------------------------------------------------------------------------
__global__ void k_view_prep(int V, const double* __restrict__ intr, const double* __restrict__ ext, ViewTab* __restrict__ vt, int with_jac,
                            const double* __restrict__ scale_cam, int ncl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  double in[9], ex[6];
  for (int j = 0; j < 9; ++j) in[j] = intr[9 * i + j];
  for (int j = 0; j < 6; ++j) ex[j] = ext[6 * i + j];
  ViewTab t;
  make_view_tab(in, ex, &t, with_jac != 0);
  for (int a = 0; a < ncl; ++a) t.sc[a] = scale_cam[i * ncl + a];
  vt[i] = t;
}

// Each thread: one observation -> weighted, Jacobi-scaled residual/Jacobian record, plus the chunk's partial camera block
// (U upper, g) and cost by a fixed-order block reduction.
//
// Persistent CTAs: CTA b owns the contiguous run of chunks [b*per, (b+1)*per) and walks it with a software pipeline, so that the
// dependent gather chain  chunk -> o_track[o] -> trk[track]  (three DRAM/L2 latencies, which one-chunk-per-CTA launches exposed in
// full at only 24 resident warps per SM) overlaps the record write-out and the block reduction of the previous chunk:
//   iteration k:  compute chunk k from registers / shared memory, stage its records in shared memory (the image of the two
//                 contiguous global blocks recA[begin..begin+cnt), recF[begin..begin+cnt))
//                 ld.global.256  trk records of chunk k+1 (64 B per thread, index loaded one iteration earlier)  -> registers
//                 ld.global      uv, track index of chunk k+2                                                    -> registers
//                 ONE thread: cp.async.bulk (TMA engine) the two staged blocks to global memory, and the 320-byte view table
//                 of chunk k+1 (with its Jacobi scales) from global memory onto an mbarrier -- no LSU traffic, no registers
//                 block reduction of the chunk partials (own shared-memory scratch: the bulk store drains behind it)
// (A first pipelined version gathered the track records with per-thread cp.async: ncu showed each such copy costing ~30 shared-memory
// wavefronts per warp -- half of the kernel's shared-memory traffic -- so the gather stays in registers, issued late.  Round 1 staged
// the view table with 18 cp.async and wrote the records out with 8 LDS + 8 STG per thread.)
constexpr int kResjacMaxPer = 64;  // chunks per CTA at most (their metadata sits in shared memory)
#ifndef PTZ_RJ_MINB
#define PTZ_RJ_MINB 5  // resident CTAs per SM the register budget of k_resjac is sized for
#endif
template <int NCL>
struct ResjacSmem {
  typedef Dims<NCL> D;
  static constexpr int kABytes = kChunk * D::RA * 8, kFBytes = kChunk * D::RF * 8, kRedBytes = (D::NPART * (kChunk + 4) + D::NPART) * 8;
  static constexpr int kBytes = kABytes + kFBytes + kRedBytes;
};
template <int TYPE>
__global__ void __launch_bounds__(kChunk, PTZ_RJ_MINB) k_resjac(int nchunks, int per, const int* __restrict__ chunk_view, const int* __restrict__ chunk_begin,
                                                                const int* __restrict__ chunk_cnt, const float2* __restrict__ o_uv,
                                                                const int* __restrict__ o_track, const ViewTab* __restrict__ vt,
                                                                const double* __restrict__ trk, const double* __restrict__ disp, int weighted,
                                                                double* __restrict__ recA, double* __restrict__ recF, double* __restrict__ part,
                                                                double* __restrict__ recd /* PTZRayDistDisp: [M][6] d r/d disp */,
                                                                const double* __restrict__ scale_d) {
  constexpr int NCL = ba_ncl(TYPE);
  typedef Dims<NCL> D;
  typedef ResjacSmem<NCL> SM;
  extern __shared__ __align__(128) unsigned char rj_smem[];
  __shared__ __align__(16) ViewTab svt[2];
  __shared__ __align__(8) unsigned long long vbar[2];
  __shared__ int s_view[kResjacMaxPer], s_begin[kResjacMaxPer], s_cnt[kResjacMaxPer];
  double2* sA = reinterpret_cast<double2*>(rj_smem);                  // [cnt][4] 16-byte pieces: the image of recA's block
  double2* sF = reinterpret_cast<double2*>(rj_smem + SM::kABytes);    // [cnt][NCL]
  double* sred = reinterpret_cast<double*>(rj_smem + SM::kABytes + SM::kFBytes);
  const int t = threadIdx.x;
  const int c0 = blockIdx.x * per, n = min(per, nchunks - c0);
  if (n <= 0) return;
  if (t < n) { s_view[t] = chunk_view[c0 + t]; s_begin[t] = chunk_begin[c0 + t]; s_cnt[t] = chunk_cnt[c0 + t]; }
  if (t == 0) { mbar_init(&vbar[0], 1); mbar_init(&vbar[1], 1); mbar_init_fence(); }
#ifdef PTZ_L2_HINTS
  const unsigned long long pol_stream = l2_evict_first_policy();
#endif
  __syncthreads();
  auto stage_view = [&](int k) {  // bulk copy of chunk k's view table into buffer k & 1 (thread 0 only)
    const int b = k & 1;
    mbar_expect_tx(&vbar[b], (unsigned)sizeof(ViewTab));
    bulk_load(&svt[b], vt + s_view[k], (unsigned)sizeof(ViewTab), &vbar[b]);
  };
  float2 uv_a = make_float2(0.f, 0.f), uv_b = uv_a, uv_c = uv_a;
  int p_b = -1, p_c = -1;
  double t0x = 0, t0y = 0, t0z = 1, t0w = 0, t1x = 1, t1y = 1, t1z = 1, t1w = 0;
  if (t < s_cnt[0]) {
    uv_a = o_uv[s_begin[0] + t];
    const int p_a = o_track[s_begin[0] + t];
    ld256(trk + (size_t)p_a * kTrk, t0x, t0y, t0z, t0w);
    ld256(trk + (size_t)p_a * kTrk + 4, t1x, t1y, t1z, t1w);
  }
  if (n > 1 && t < s_cnt[1]) { uv_b = o_uv[s_begin[1] + t]; p_b = o_track[s_begin[1] + t]; }
  if (t == 0) stage_view(0);
  double dz[3] = {0, 0, 0};
  if (TYPE == BA_PTZRAY_DIST_DISP) { dz[0] = disp[0]; dz[1] = disp[1]; dz[2] = disp[2]; }
  for (int k = 0; k < n; ++k) {
    if (t == 0 && k > 0) bulk_wait_read();  // the TMA engine is done reading the previous chunk's staged records
    __syncthreads();                        // ... and everybody is done with the previous chunk's buffers
    if (t == 0 && k + 1 < n) stage_view(k + 1);
    mbar_wait(&vbar[k & 1], (unsigned)((k >> 1) & 1));  // chunk k's view table has landed in svt[k & 1]
    const int b = k & 1, chunk = c0 + k, begin = s_begin[k], cnt = s_cnt[k];
    double acc[D::NPART];
#pragma unroll
    for (int i = 0; i < D::NPART; ++i) acc[i] = 0.0;
    if (t < cnt) {
      const double ray[3] = {t0x, t0y, t0z};
      double r[2], F[2 * NCL], E[6], Fd[6];
      ba_obs<TYPE, true>(svt[b], ray, dz, (double)uv_a.x, (double)uv_a.y, r, F, E, TYPE == BA_PTZRAY_DIST_DISP ? Fd : nullptr);
      const double sw = weighted ? t0w : 1.0;
      if (TYPE == BA_PTZRAY_DIST_DISP) {
        double* fo = recd + (size_t)(begin + t) * 6;
#pragma unroll
        for (int j = 0; j < 3; ++j) { fo[j] = Fd[j] * sw * scale_d[j]; fo[3 + j] = Fd[3 + j] * sw * scale_d[j]; }
      }
      r[0] *= sw; r[1] *= sw;
      const double sr[3] = {t1x * sw, t1y * sw, t1z * sw};
#pragma unroll
      for (int j = 0; j < 3; ++j) { E[j] *= sr[j]; E[3 + j] *= sr[j]; }
#pragma unroll
      for (int a = 0; a < NCL; ++a) { const double s = svt[b].sc[a] * sw; F[a] *= s; F[NCL + a] *= s; }
      // the record in 16-byte pieces, linear in shared memory (= its global image).  Rows of 4 pieces would put piece p of every
      // thread into the same two bank groups: thread t stores its pieces in the rotated order (s + t/2) & 3 instead
      const double2 ra[4] = {make_double2(r[0], r[1]), make_double2(E[0], E[1]), make_double2(E[2], E[3]), make_double2(E[4], E[5])};
      const int rot = (t >> 1) & 3;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const int pch = (s4 + rot) & 3;
        const double2 lo = (pch & 1) ? ra[1] : ra[0], hi = (pch & 1) ? ra[3] : ra[2];
        sA[t * 4 + pch] = (pch & 2) ? hi : lo;
      }
      if (NCL == 4) {
        const double2 rf[4] = {make_double2(F[0], F[1]), make_double2(F[2], F[3]), make_double2(F[4], F[5]), make_double2(F[6], F[7])};
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int pch = (s4 + rot) & 3;
          const double2 lo = (pch & 1) ? rf[1] : rf[0], hi = (pch & 1) ? rf[3] : rf[2];
          sF[t * 4 + pch] = (pch & 2) ? hi : lo;
        }
      } else {
#pragma unroll
        for (int a = 0; a < NCL; ++a) sF[t * NCL + a] = make_double2(F[2 * a], F[2 * a + 1]);  // F stored flat [2*NCL]; 5- and 6-piece rows spread
      }
      int kk = 0;
#pragma unroll
      for (int a = 0; a < NCL; ++a)
#pragma unroll
        for (int bb = a; bb < NCL; ++bb) acc[kk++] = F[a] * F[bb] + F[NCL + a] * F[NCL + bb];
#pragma unroll
      for (int a = 0; a < NCL; ++a) acc[D::NU + a] = F[a] * r[0] + F[NCL + a] * r[1];
      acc[D::NU + NCL] = 0.5 * (r[0] * r[0] + r[1] * r[1]);
    }
    // gathers of the next two chunks: in flight during the write-out and the reduction below
    if (p_b >= 0) {
      ld256(trk + (size_t)p_b * kTrk, t0x, t0y, t0z, t0w);
      ld256(trk + (size_t)p_b * kTrk + 4, t1x, t1y, t1z, t1w);
    }
    p_c = -1;
    if (k + 2 < n && t < s_cnt[k + 2]) { uv_c = o_uv[s_begin[k + 2] + t]; p_c = o_track[s_begin[k + 2] + t]; }
    bulk_store_fence();  // this thread's shared-memory writes become visible to the TMA engine ...
    __syncthreads();     // ... and all of them are done
    if (t == 0) {        // the chunk's records are contiguous in both arrays: two bulk stores
#ifdef PTZ_L2_HINTS
      bulk_store_hint(recA + (size_t)begin * D::RA, sA, (unsigned)(cnt * D::RA * 8), pol_stream);
      bulk_store_hint(recF + (size_t)begin * D::RF, sF, (unsigned)(cnt * D::RF * 8), pol_stream);
#else
      bulk_store(recA + (size_t)begin * D::RA, sA, (unsigned)(cnt * D::RA * 8));
      bulk_store(recF + (size_t)begin * D::RF, sF, (unsigned)(cnt * D::RF * 8));
#endif
      bulk_commit();
    }
    // (one reduction per CHUNK here: keeping the 15 partial sums in registers across the chunks of a view, as k_obs_what does, costs
    // k_resjac its fifth resident CTA or 176 bytes of spills -- measured 88 -> 98 / 117 us)
    block_sum_sm<D::NPART, kChunk>(acc, sred);
    if (t == 0) {
#pragma unroll
      for (int i = 0; i < D::NPART; ++i) part[(size_t)chunk * D::NPART + i] = acc[i];
    }
    uv_a = uv_b; uv_b = uv_c; p_b = p_c;
  }
  if (t == 0) bulk_wait_read();  // (shared memory must outlive the last store's reads)
}

// per view: sum the chunk partials in chunk order -> full symmetric U[NCL*NCL], g[NCL], cost; |g/s| for the gradient norm
template <int NCL>
__global__ void k_view_finalize(int V, const int* __restrict__ view_chunk_off, const double* __restrict__ part, const double* __restrict__ scale_cam,
                                double* __restrict__ U, double* __restrict__ g, double* __restrict__ cost_view, double* __restrict__ gabs) {
  typedef Dims<NCL> D;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = idx / D::NPART, e = idx % D::NPART;
  if (v >= V) return;
  double s = 0;
  for (int c = view_chunk_off[v]; c < view_chunk_off[v + 1]; ++c) s += part[(size_t)c * D::NPART + e];
  if (e < D::NU) {
    int a = 0, rem = e;
    while (rem >= NCL - a) { rem -= NCL - a; ++a; }
    const int b = a + rem;
    U[(size_t)v * NCL * NCL + a * NCL + b] = s;
    U[(size_t)v * NCL * NCL + b * NCL + a] = s;
  } else if (e < D::NU + NCL) {
    const int a = e - D::NU;
    g[v * NCL + a] = s;
    gabs[v * NCL + a] = fabs(s / scale_cam[v * NCL + a]);
  } else {
    cost_view[v] = s;
  }
}

// damp and factor one ray block: (V + D^2) = L L^T, t = L^-1 h (D^2 = clamp(diag V)/mu, refreshed after accepted steps only)
__device__ __forceinline__ void track_factor_one(int p, bool empty, const double (&vh)[9], double mu, int refresh_diag, double min_diag, double max_diag,
                                                 double* __restrict__ diag_ray, double* __restrict__ Lt, int* __restrict__ fail) {
  double d0, d1, d2;
  if (refresh_diag) {
    d0 = fmin(fmax(vh[0], min_diag), max_diag); d1 = fmin(fmax(vh[2], min_diag), max_diag); d2 = fmin(fmax(vh[5], min_diag), max_diag);
    diag_ray[3 * (size_t)p] = d0; diag_ray[3 * (size_t)p + 1] = d1; diag_ray[3 * (size_t)p + 2] = d2;
  } else {
    d0 = diag_ray[3 * (size_t)p]; d1 = diag_ray[3 * (size_t)p + 1]; d2 = diag_ray[3 * (size_t)p + 2];
  }
  const double A6[6] = {vh[0] + d0 / mu, vh[1], vh[2] + d1 / mu, vh[3], vh[4], vh[5] + d2 / mu};
  double L[6] = {1, 0, 1, 0, 0, 1};
  double* lt = Lt + (size_t)p * 10;
  if (!empty && !chol3(A6, L)) {
    atomicExch(fail, 1);
    L[0] = L[2] = L[5] = 1.0; L[1] = L[3] = L[4] = 0.0;
  }
  const double t0 = vh[6] / L[0], t1 = (vh[7] - L[1] * t0) / L[2], t2 = (vh[8] - L[3] * t0 - L[4] * t1) / L[5];
  lt[0] = L[0]; lt[1] = L[1]; lt[2] = L[2]; lt[3] = L[3]; lt[4] = L[4]; lt[5] = L[5];
  lt[6] = empty ? 0.0 : t0; lt[7] = empty ? 0.0 : t1; lt[8] = empty ? 0.0 : t2; lt[9] = 0;
}
// per track: V = sum E^T E (lower 6), h = sum E^T r.  Once per Jacobian evaluation (gradient, column norms, LM diagonal).
// factor_mu > 0: the damped Cholesky factor of the NEXT linear solve is formed right here from the sums still in registers (after an
// accepted step the new radius is known before the Jacobian is evaluated), which saves that solve its k_track_factor launch.
__global__ void k_track_accum(int P, const int* __restrict__ t_off, const int* __restrict__ t_obs, const double* __restrict__ recA,
                              const double* __restrict__ trk, double* __restrict__ Vh, double* __restrict__ gmax_part, double factor_mu, double min_diag,
                              double max_diag, double* __restrict__ diag_ray, double* __restrict__ Lt, int* __restrict__ fail) {
  __shared__ double sm[8];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double gm = 0;
  if (p < P) {
    double v00 = 0, v10 = 0, v11 = 0, v20 = 0, v21 = 0, v22 = 0, h0 = 0, h1 = 0, h2 = 0;
    const int tb = t_off[p], te = t_off[p + 1];
#ifndef PTZ_TA_U
#define PTZ_TA_U 6
#endif
    constexpr int U = PTZ_TA_U;
    for (int i = tb; i < te; i += U) {
      // U records in flight (indices clamped, extra ones weighted 0)
      double2 rr[U], ea[U], eb[U], ec[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // a 64-byte record = two 256-bit loads
        const double* q = recA + (size_t)t_obs[min(i + u, te - 1)] * 8;
        ld256(q, rr[u].x, rr[u].y, ea[u].x, ea[u].y);
        ld256(q + 4, eb[u].x, eb[u].y, ec[u].x, ec[u].y);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i + u >= te) break;
        const double a0 = ea[u].x, a1 = ea[u].y, a2 = eb[u].x, b0 = eb[u].y, b1 = ec[u].x, b2 = ec[u].y;
        v00 += a0 * a0 + b0 * b0; v10 += a1 * a0 + b1 * b0; v11 += a1 * a1 + b1 * b1;
        v20 += a2 * a0 + b2 * b0; v21 += a2 * a1 + b2 * b1; v22 += a2 * a2 + b2 * b2;
        h0 += a0 * rr[u].x + b0 * rr[u].y; h1 += a1 * rr[u].x + b1 * rr[u].y; h2 += a2 * rr[u].x + b2 * rr[u].y;
      }
    }
    double* o = Vh + (size_t)p * 10;
    o[0] = v00; o[1] = v10; o[2] = v11; o[3] = v20; o[4] = v21; o[5] = v22; o[6] = h0; o[7] = h1; o[8] = h2; o[9] = 0;
    if (factor_mu > 0.0) {
      const double vh[9] = {v00, v10, v11, v20, v21, v22, h0, h1, h2};
      track_factor_one(p, tb == te, vh, factor_mu, 1, min_diag, max_diag, diag_ray, Lt, fail);
    }
    const double* t = trk + (size_t)p * kTrk;
    gm = fmax(fabs(h0 / t[4]), fmax(fabs(h1 / t[5]), fabs(h2 / t[6])));
  }
  gm = warp_max(gm);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = gm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, sm[w]);
    gmax_part[blockIdx.x] = m;
  }
}

// ---- by-track passes with ONE LANE PER (track, observation) ENTRY (opt-in experiment, PTZ_BYTRACK_WARP=1: see ba_solver.cu) ----
// The by-track list (t_off / t_obs) is cut into groups: group g = the tracks whose first entry lies in [32 g, 32 g + 32).  A warp takes a
// group: every lane gathers ONE record (all loads of a round independent: 32 gathers in flight per warp instead of 4 per thread), the
// per-track sums are formed by a segmented shuffle reduction (entries of a track are neighbours; fixed tree order: bit-reproducible),
// and the lane at the head of each segment finishes the track.  A group has at most 32 + (longest track - 1) entries: the tail
// of its last track is a second round whose partial sums are carried.  The default kernels run one THREAD per track, four records in
// flight: k_track_accum 62 us, k_track_backsub 87 us against 36 / 55 us for a plain random gather of the same records.  This
// version measured 110 / 131 us: too little work per warp for the length of its dependent chain.
constexpr int kNoGroup = 0x7f7f7f7f;  // (what cudaMemset(0x7f) leaves: larger than any entry index)
__global__ void k_entry_tracks(int P, const int* __restrict__ t_off, int* __restrict__ t_trk, int* __restrict__ grp_e0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int b = t_off[p], e = t_off[p + 1];
  for (int i = b; i < e; ++i) t_trk[i] = p;
  // first entry of the first non-empty track that starts in group b / 32 (entries ascend with p: the minimum wins)
  if (e > b) atomicMin(grp_e0 + (b >> 5), b);
}
// reduce NV values over runs of equal `key` among neighbouring lanes; afterwards the first lane of a run holds the run's sum
template <int NV>
__device__ __forceinline__ void seg_reduce(double (&v)[NV], int key, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int k2 = __shfl_down_sync(0xffffffffu, key, off);
    const bool take = lane + off < 32 && k2 == key;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double t = __shfl_down_sync(0xffffffffu, v[j], off);
      if (take) v[j] += t;
    }
  }
}
// V = sum E^T E (lower 6), h = sum E^T r per track; gradient max-norm partial per warp
__global__ void __launch_bounds__(256) k_track_accum_w(int G, int M, const int* __restrict__ grp_e0, const int* __restrict__ t_trk, const int* __restrict__ t_obs,
                                                       const double* __restrict__ recA, const double* __restrict__ trk, double* __restrict__ Vh,
                                                       double* __restrict__ gmax_part) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= G) return;
  double gm = 0.0;
  const int e0 = grp_e0[g];
  if (e0 != kNoGroup) {
    int gg = g + 1;
    while (gg < G && grp_e0[gg] == kNoGroup) ++gg;
    const int e1 = gg < G ? grp_e0[gg] : M;
    double carry[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) carry[j] = 0.0;
    int carry_trk = -1;
    for (int base = e0; base < e1; base += 32) {
      const int e = base + lane;
      const bool valid = e < e1;
      const int key = valid ? t_trk[e] : -2 - lane;
      double v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = 0.0;
      if (valid) {
        const double* q = recA + (size_t)t_obs[e] * 8;
        double r0, r1, a0, a1, a2, b0, b1, b2;
        ld256(q, r0, r1, a0, a1);
        ld256(q + 4, a2, b0, b1, b2);
        v[0] = a0 * a0 + b0 * b0; v[1] = a1 * a0 + b1 * b0; v[2] = a1 * a1 + b1 * b1;
        v[3] = a2 * a0 + b2 * b0; v[4] = a2 * a1 + b2 * b1; v[5] = a2 * a2 + b2 * b2;
        v[6] = a0 * r0 + b0 * r1; v[7] = a1 * r0 + b1 * r1; v[8] = a2 * r0 + b2 * r1;
      }
      seg_reduce<9>(v, key, lane);
      const int prev = __shfl_up_sync(0xffffffffu, key, 1);
      const bool head = valid && (lane == 0 || prev != key);
      if (head && lane == 0 && key == carry_trk) {
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] += carry[j];
      }
      // does the last track of this round go on in the next one?
      const int last_key = __shfl_sync(0xffffffffu, key, 31);
      const bool goes_on = base + 32 < e1 && last_key >= 0 && t_trk[base + 32] == last_key;
      // the head of that track hands its sum to the next round instead of writing it
      const bool is_carry = head && goes_on && key == last_key;
      const unsigned cm = __ballot_sync(0xffffffffu, is_carry);
      if (cm) {
        const int src = __ffs(cm) - 1;
#pragma unroll
        for (int j = 0; j < 9; ++j) carry[j] = __shfl_sync(0xffffffffu, v[j], src);
        carry_trk = last_key;
      }
      if (head && !is_carry) {
        double* o = Vh + (size_t)key * 10;
#pragma unroll
        for (int j = 0; j < 9; ++j) o[j] = v[j];
        o[9] = 0.0;
        const double* t = trk + (size_t)key * kTrk;
        gm = fmax(gm, fmax(fabs(v[6] / t[4]), fmax(fabs(v[7] / t[5]), fabs(v[8] / t[6]))));
      }
    }
  }
  gm = warp_max(gm);
  if (lane == 0) gmax_part[g] = gm;
}
// back-substitution of the rays, same scheme: per entry  c = What_o^T y_view(o), summed per track; the head lane solves
// L^T y_p = t - c, writes the candidate ray; partial sums (model cost change, |step|^2, |x_cand|^2) per warp
template <int NCL>
__global__ void __launch_bounds__(256) k_track_backsub_w(int G, int M, const int* __restrict__ grp_e0, const int* __restrict__ t_trk, const int* __restrict__ t_obs,
                                                         const int* __restrict__ t_view, const double* __restrict__ What, const double* __restrict__ y,
                                                         const double* __restrict__ Lt, const double* __restrict__ Vh, const double* __restrict__ diag_ray,
                                                         double mu, const double* __restrict__ trk, double* __restrict__ trk_cand,
                                                         double* __restrict__ part3, const double* __restrict__ Wdh /* or nullptr */,
                                                         const double* __restrict__ ydisp) {
  typedef Dims<NCL> D;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= G) return;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  const int e0 = grp_e0[g];
  if (e0 != kNoGroup) {
    int gg = g + 1;
    while (gg < G && grp_e0[gg] == kNoGroup) ++gg;
    const int e1 = gg < G ? grp_e0[gg] : M;
    double carry[3] = {0.0, 0.0, 0.0};
    int carry_trk = -1;
    for (int base = e0; base < e1; base += 32) {
      const int e = base + lane;
      const bool valid = e < e1;
      const int key = valid ? t_trk[e] : -2 - lane;
      double v[3] = {0.0, 0.0, 0.0};
      if (valid) {
        const double* w4 = What + (size_t)t_obs[e] * D::WS;
        const double* yv = y + (size_t)t_view[e] * NCL;
        double w[D::WS];
#pragma unroll
        for (int k = 0; k < D::WS / 4; ++k) ld256(w4 + 4 * k, w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
#pragma unroll
        for (int a = 0; a < NCL; ++a) { const double ya = yv[a]; v[0] += w[3 * a] * ya; v[1] += w[3 * a + 1] * ya; v[2] += w[3 * a + 2] * ya; }
      }
      seg_reduce<3>(v, key, lane);
      const int prev = __shfl_up_sync(0xffffffffu, key, 1);
      const bool head = valid && (lane == 0 || prev != key);
      if (head && lane == 0 && key == carry_trk) { v[0] += carry[0]; v[1] += carry[1]; v[2] += carry[2]; }
      const int last_key = __shfl_sync(0xffffffffu, key, 31);
      const bool goes_on = base + 32 < e1 && last_key >= 0 && t_trk[base + 32] == last_key;
      const bool is_carry = head && goes_on && key == last_key;
      const unsigned cm = __ballot_sync(0xffffffffu, is_carry);
      if (cm) {
        const int src = __ffs(cm) - 1;
#pragma unroll
        for (int j = 0; j < 3; ++j) carry[j] = __shfl_sync(0xffffffffu, v[j], src);
        carry_trk = last_key;
      }
      if (head && !is_carry) {
        const int p = key;
        const double* lt = Lt + (size_t)p * 10;
        const double* t = trk + (size_t)p * kTrk;
        double b0 = lt[6] - v[0], b1 = lt[7] - v[1], b2 = lt[8] - v[2];
        if (Wdh) {  // PTZRayDistDisp: - Wdh^T y_disp
          const double* wd = Wdh + (size_t)p * 12;
#pragma unroll
          for (int a = 0; a < 3; ++a) { const double ya = ydisp[a]; b0 -= wd[3 * a] * ya; b1 -= wd[3 * a + 1] * ya; b2 -= wd[3 * a + 2] * ya; }
        }
        const double y2 = b2 / lt[5], y1 = (b1 - lt[4] * y2) / lt[2], y0 = (b0 - lt[1] * y1 - lt[3] * y2) / lt[0];
        const double* vh = Vh + (size_t)p * 10;
        const double* dg = diag_ray + 3 * (size_t)p;
        acc0 += 0.5 * (y0 * (vh[6] + dg[0] / mu * y0) + y1 * (vh[7] + dg[1] / mu * y1) + y2 * (vh[8] + dg[2] / mu * y2));
        const double c0 = t[0] + (-t[4] * y0), c1 = t[1] + (-t[5] * y1), c2 = t[2] + (-t[6] * y2);
        double* tc = trk_cand + (size_t)p * kTrk;
        tc[0] = c0; tc[1] = c1; tc[2] = c2;
        acc1 += (t[0] - c0) * (t[0] - c0) + (t[1] - c1) * (t[1] - c1) + (t[2] - c2) * (t[2] - c2);
        acc2 += c0 * c0 + c1 * c1 + c2 * c2;
      }
    }
  }
  acc0 = warp_sum(acc0); acc1 = warp_sum(acc1); acc2 = warp_sum(acc2);
  if (lane == 0) { part3[3 * (size_t)g] = acc0; part3[3 * (size_t)g + 1] = acc1; part3[3 * (size_t)g + 2] = acc2; }
}

// Jacobi scaling, once at iteration 0: s = 1 / (1 + sqrt(column norm^2))      (TrustRegionMinimizer::EvaluateGradientAndJacobian)
template <int NCL>
__global__ void k_make_scales(int V, int P, const double* __restrict__ U, const double* __restrict__ Vh, double* __restrict__ scale_cam,
                              double* __restrict__ trk0, double* __restrict__ trk1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V * NCL) {
    const int v = i / NCL, a = i % NCL;
    scale_cam[i] = 1.0 / (1.0 + sqrt(U[(size_t)v * NCL * NCL + a * NCL + a]));
  }
  if (i < P) {
    const double* vh = Vh + (size_t)i * 10;
    const double s0 = 1.0 / (1.0 + sqrt(vh[0])), s1 = 1.0 / (1.0 + sqrt(vh[2])), s2 = 1.0 / (1.0 + sqrt(vh[5]));
    trk0[(size_t)i * kTrk + 4] = s0; trk0[(size_t)i * kTrk + 5] = s1; trk0[(size_t)i * kTrk + 6] = s2;
    trk1[(size_t)i * kTrk + 4] = s0; trk1[(size_t)i * kTrk + 5] = s1; trk1[(size_t)i * kTrk + 6] = s2;
  }
}

// -------------------------------------------------------------------------------------------------------------
// stage 2
// -------------------------------------------------------------------------------------------------------------
// per track: damp and factor (V + D^2) = L L^T, t = L^-1 h
__global__ void k_track_factor(int P, const int* __restrict__ t_off, const double* __restrict__ Vh, double mu, int refresh_diag, double min_diag,
                               double max_diag, double* __restrict__ diag_ray, double* __restrict__ Lt, int* __restrict__ fail) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double* q = Vh + (size_t)p * 10;
  const double vh[9] = {q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8]};
  track_factor_one(p, t_off[p] == t_off[p + 1], vh, mu, refresh_diag, min_diag, max_diag, diag_ray, Lt, fail);
}
// per observation (one thread each; a chunk of one view per step, so records stream and the view's Schur terms reduce in the
// CTA): What = (F^T E) L^-T, q = What t; chunk partials of sum What What^T (upper) and sum q for k_schur_diag.
//
// Persistent CTAs over contiguous runs of chunks, as k_resjac.  The records of chunk k+1 stream into a second shared-memory
// buffer with coalesced 16-byte cp.async (no register staging) while chunk k computes; the per-track Cholesky factors of
// chunk k+1 are gathered into registers after the arithmetic of chunk k, so they fly during its write-out and reduction.
#ifndef PTZ_OW_MINB
#define PTZ_OW_MINB 4
#endif
template <int NCL>
struct ObsWhatSmem {
  typedef Dims<NCL> D;
  static constexpr int NV = D::NU + NCL, RC = D::RS / 2, WC = D::WS / 2;
  static constexpr int kRecBytes = kChunk * RC * 16, kWBytes = kChunk * WC * 16, kRedBytes = (NV * (kChunk + 4) + NV) * 8;
  static constexpr int kBytes = 2 * kRecBytes + (kWBytes > kRedBytes ? kWBytes : kRedBytes);
};
template <int NCL>
__global__ void __launch_bounds__(kChunk, PTZ_OW_MINB) k_obs_what(int nchunks, int per, const int* __restrict__ chunk_view, const int* __restrict__ chunk_begin,
                                                                  const int* __restrict__ chunk_cnt,
                                                                  const int* __restrict__ o_track, const double* __restrict__ recA,
                                                                  const double* __restrict__ recF, const double* __restrict__ Lt,
                                                                  double* __restrict__ What, double* __restrict__ wpart) {
  typedef Dims<NCL> D;
  typedef ObsWhatSmem<NCL> SM;
  constexpr int NV = SM::NV, RC = SM::RC, WC = SM::WC;
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ int s_begin[kResjacMaxPer], s_cnt[kResjacMaxPer], s_view[kResjacMaxPer];
  double2* sw = reinterpret_cast<double2*>(dsm + 2 * SM::kRecBytes);
  double* sred = reinterpret_cast<double*>(dsm + 2 * SM::kRecBytes);  // reused after the staged What has been written out
  const int t = threadIdx.x;
  const int c0 = blockIdx.x * per, n = min(per, nchunks - c0);
  if (n <= 0) return;
  if (t < n) { s_begin[t] = chunk_begin[c0 + t]; s_cnt[t] = chunk_cnt[c0 + t]; s_view[t] = chunk_view[c0 + t]; }
  __syncthreads();
  auto stage_rec = [&](int k) {  // the chunk's records are two contiguous blocks: coalesced 16-byte copies into (swizzled) shared memory
    double2* dst = reinterpret_cast<double2*>(dsm + (k & 1) * SM::kRecBytes);
    const double2* ga = reinterpret_cast<const double2*>(recA + (size_t)s_begin[k] * D::RA);
    const double2* gf = reinterpret_cast<const double2*>(recF + (size_t)s_begin[k] * D::RF);
    const int cnt = s_cnt[k];
#ifdef PTZ_L2_HINTS
    const unsigned long long pol_stream = l2_evict_first_policy();
#define PTZ_OW_CP(d, g) cp_async16_hint(d, g, pol_stream)
#else
#define PTZ_OW_CP(d, g) cp_async16(d, g)
#endif
    for (int gch = t; gch < cnt * 4; gch += kChunk) {
      const int tt = gch >> 2, pch = gch & 3;
      PTZ_OW_CP(&dst[tt * RC + chunk_swz<RC>(tt, pch)], ga + gch);
    }
    for (int gch = t; gch < cnt * NCL; gch += kChunk) {
      const int tt = gch / NCL, pch = 4 + gch % NCL;
      PTZ_OW_CP(&dst[tt * RC + chunk_swz<RC>(tt, pch)], gf + gch);
    }
#undef PTZ_OW_CP
    cp_async_commit();
  };
  double2 l01 = make_double2(1, 0), l23 = make_double2(1, 0), l45 = make_double2(0, 1), l67 = make_double2(0, 0), l89 = make_double2(0, 0);
  int p_b = -1, p_c = -1;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  stage_rec(0);
  if (t < s_cnt[0]) {
    const double2* lp = reinterpret_cast<const double2*>(Lt + (size_t)o_track[s_begin[0] + t] * 10);
    l01 = lp[0]; l23 = lp[1]; l45 = lp[2]; l67 = lp[3]; l89 = lp[4];
  }
  if (n > 1 && t < s_cnt[1]) p_b = o_track[s_begin[1] + t];
  for (int k = 0; k < n; ++k) {
    if (k + 1 < n) stage_rec(k + 1); else cp_async_commit();  // (an empty group keeps the wait below uniform)
    cp_async_wait<1>();
    __syncthreads();
    const double2* srec = reinterpret_cast<const double2*>(dsm + (k & 1) * SM::kRecBytes);
    const int chunk = c0 + k, begin = s_begin[k], cnt = s_cnt[k];
    if (t < cnt) {
      const double L1 = l01.y, L3 = l23.y, L4 = l45.x, t0 = l67.x, t1 = l67.y, t2 = l89.x;
      const double i00 = 1.0 / l01.x, i11 = 1.0 / l23.x, i22 = 1.0 / l45.y;
      const double2 e01 = srec[t * RC + chunk_swz<RC>(t, 1)], e23 = srec[t * RC + chunk_swz<RC>(t, 2)], e45 = srec[t * RC + chunk_swz<RC>(t, 3)];
      const double a0 = e01.x, a1 = e01.y, a2 = e23.x, b0 = e23.y, b1 = e45.x, b2 = e45.y;
      double Fv[2 * NCL];
#pragma unroll
      for (int a = 0; a < NCL; ++a) { const double2 f = srec[t * RC + chunk_swz<RC>(t, 4 + a)]; Fv[2 * a] = f.x; Fv[2 * a + 1] = f.y; }
      double w[D::WS];
#pragma unroll
      for (int a = 0; a < NCL; ++a) {
        const double f0 = Fv[a], f1 = Fv[NCL + a];
        const double w0 = f0 * a0 + f1 * b0, w1 = f0 * a1 + f1 * b1, w2 = f0 * a2 + f1 * b2;  // row a of F^T E
        const double x0 = w0 * i00, x1 = (w1 - L1 * x0) * i11, x2 = (w2 - L3 * x0 - L4 * x1) * i22;
        w[3 * a] = x0; w[3 * a + 1] = x1; w[3 * a + 2] = x2;
        acc[D::NU + a] += x0 * t0 + x1 * t1 + x2 * t2;  // q_a
      }
#pragma unroll
      for (int i = 3 * NCL; i < D::WS; ++i) w[i] = 0.0;
#pragma unroll
      for (int kk = 0; kk < WC; ++kk) sw[t * WC + chunk_swz<WC>(t, kk)] = make_double2(w[2 * kk], w[2 * kk + 1]);
      int kk = 0;
#pragma unroll
      for (int a = 0; a < NCL; ++a)
#pragma unroll
        for (int bb = a; bb < NCL; ++bb) acc[kk++] += w[3 * a] * w[3 * bb] + w[3 * a + 1] * w[3 * bb + 1] + w[3 * a + 2] * w[3 * bb + 2];
    }
    // gathers for the next two chunks: in flight during the write-out and the reduction below
    if (p_b >= 0) {
      const double2* lp = reinterpret_cast<const double2*>(Lt + (size_t)p_b * 10);
      l01 = lp[0]; l23 = lp[1]; l45 = lp[2]; l67 = lp[3]; l89 = lp[4];
    }
    p_c = -1;
    if (k + 2 < n && t < s_cnt[k + 2]) p_c = o_track[s_begin[k + 2] + t];
    if constexpr (WC != 8) {
      // the staged What records are linear in shared memory (= the chunk's contiguous global block): ONE bulk store by the TMA
      // engine instead of WC LDS + WC STG per thread; the scratch of the reduction below aliases the staging area, so wait until
      // the engine has read it
      bulk_store_fence();
      __syncthreads();
      if (t == 0) {
        bulk_store(What + (size_t)begin * D::WS, sw, (unsigned)(cnt * D::WS * 8));
        bulk_commit();
        bulk_wait_read();
      }
    } else {
      __syncthreads();
      double2* gout = reinterpret_cast<double2*>(What + (size_t)begin * D::WS);
      for (int gch = t; gch < cnt * WC; gch += kChunk) {
        const int tt = gch / WC, pch = gch % WC;
        gout[gch] = sw[tt * WC + chunk_swz<WC>(tt, pch)];
      }
    }
    __syncthreads();
    if (k + 1 == n || s_view[k + 1] != s_view[k]) {  // one block reduction per view of the CTA's run (as k_resjac): the other slots stay zero
      block_sum_sm<NV, kChunk>(acc, sred);
      if (t == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) wpart[(size_t)chunk * NV + i] = acc[i];
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    }
    p_b = p_c;
  }
}

// per (view, value) one thread: S_cc = [U + D^2] - sum_chunks(partial What What^T), rhs_c = [g] - sum_chunks(partial q), chunk order
// fixed.  The bracketed terms are added by the rank that owns the camera blocks (add_own); D^2 = clamp(diag U)/mu, refreshed after
// accepted steps only.  (Round 1 ran one thread per VIEW, 14-27 dependent sums each: 31 us for 1000 views.)
template <int NCL>
__global__ void k_schur_diag(int V, const int* __restrict__ view_chunk_off, const double* __restrict__ wpart, const double* __restrict__ U,
                             const double* __restrict__ g, double mu, int refresh_diag, double min_diag, double max_diag, int add_own,
                             double* __restrict__ diag_cam, const int* __restrict__ diag_pos, double* __restrict__ Sval, double* __restrict__ rhs,
                             const int* __restrict__ grp_of /* shared intrinsics: group of the view or -1; nullptr = none */) {
  typedef Dims<NCL> D;
  constexpr int NV = D::NU + NCL;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = idx / NV, e = idx % NV;
  if (v >= V) return;
  double acc = 0.0;
  const int c1 = view_chunk_off[v + 1];
  int c = view_chunk_off[v];
  for (; c + 4 <= c1; c += 4) {  // four loads in flight, summed in chunk order
    const double a0 = wpart[(size_t)c * NV + e], a1 = wpart[(size_t)(c + 1) * NV + e], a2 = wpart[(size_t)(c + 2) * NV + e], a3 = wpart[(size_t)(c + 3) * NV + e];
    acc += a0; acc += a1; acc += a2; acc += a3;
  }
  for (; c < c1; ++c) acc += wpart[(size_t)c * NV + e];
  if (e >= D::NU) {
    const int a = e - D::NU;
    rhs[v * NCL + a] = (add_own ? g[v * NCL + a] : 0.0) - acc;
    return;
  }
  int a = 0, rem = e;
  while (rem >= NCL - a) { rem -= NCL - a; ++a; }
  const int b = a + rem;
  const double* Uv = U + (size_t)v * NCL * NCL;
  double* S = Sval + (size_t)(diag_pos != nullptr ? diag_pos[v] : v) * NCL * NCL;  // (nullptr: packed layout of the sharded problem)
  double sacc = -acc;
  if (add_own) sacc += Uv[a * NCL + b];
  if (a == b) {
    double d;
    if (grp_of != nullptr && a < NCL - 3 && grp_of[v] >= 0) { d = 0.0; diag_cam[v * NCL + a] = 0.0; }  // damped once, in the border (k_shared_border)
    else if (refresh_diag) { d = fmin(fmax(Uv[a * NCL + a], min_diag), max_diag); diag_cam[v * NCL + a] = d; }
    else d = diag_cam[v * NCL + a];
    if (add_own) sacc += d / mu;
  }
  S[a * NCL + b] = sacc;
  S[b * NCL + a] = sacc;
}

// per upper off-diagonal block (one warp): S_rc = - sum over observation pairs What_o What_o'^T ; also writes S_cr = S_rc^T
#ifndef PTZ_OD_MINB
#define PTZ_OD_MINB 1
#endif
template <int NCL>
__global__ void __launch_bounds__(256, PTZ_OD_MINB) k_schur_offdiag(int nlist, const int* __restrict__ ub_list, const int64_t* __restrict__ pair_off,
                                                       const int* __restrict__ pair_a, const int* __restrict__ pair_b, const double* __restrict__ What,
                                                       const int* __restrict__ ub_pos, const int* __restrict__ ub_pos_t, double* __restrict__ Sval) {
  typedef Dims<NCL> D;
  const int li = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (li >= nlist) return;
  // sharded problem: the block pattern is the union over the ranks, this rank has pairs in ~1/W of the blocks; ub_list names those (the
  // others were zeroed by a memset), nullptr = every block
  const int b = ub_list != nullptr ? ub_list[li] : li;
  double acc[NCL * NCL];
#pragma unroll
  for (int i = 0; i < NCL * NCL; ++i) acc[i] = 0.0;
  // the pair indices of the next round are fetched before the gathers of this one: one dependent latency less per round
  const int64_t ke = pair_off[b + 1];
  int64_t k = pair_off[b] + lane;
  int ia = 0, ib = 0;
  if (k < ke) { ia = pair_a[k]; ib = pair_b[k]; }
  for (; k < ke; k += 32) {
    const double* wa = What + (size_t)ia * D::WS;
    const double* wb = What + (size_t)ib * D::WS;
    if (k + 32 < ke) { ia = pair_a[k + 32]; ib = pair_b[k + 32]; }
    double x[D::WS], y[D::WS];  // 256-bit loads: half the load instructions and L1 wavefronts of the 16-byte version
#pragma unroll
    for (int i = 0; i < D::WS / 4; ++i) ld256(wa + 4 * i, x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
#pragma unroll
    for (int i = 0; i < D::WS / 4; ++i) ld256(wb + 4 * i, y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
#pragma unroll
    for (int a = 0; a < NCL; ++a)
#pragma unroll
      for (int c = 0; c < NCL; ++c) acc[a * NCL + c] += x[3 * a] * y[3 * c] + x[3 * a + 1] * y[3 * c + 1] + x[3 * a + 2] * y[3 * c + 2];
  }
  // reduce-scatter over the warp: element e of the block ends up in lane e * (32 / NP); those lanes store it (and its transpose)
  constexpr int NN = NCL * NCL, NP = NN <= 16 ? 16 : 32, REST = NN > 32 ? NN - 32 : 0;
  // packed layout (sharded problem, ub_pos == nullptr): block b itself, no transpose -- k_unpack_S mirrors after the all-reduce
  const bool packed = ub_pos == nullptr;
  double* S = Sval + (size_t)(packed ? b : ub_pos[b]) * NN;
  double* St = packed ? S : Sval + (size_t)ub_pos_t[b] * NN;
  {
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) v[i] = i < NN ? acc[i] : 0.0;
    const double tot = warp_reduce_scatter<NP>(v, lane);
    const int e = lane / (32 / NP);
    if (lane % (32 / NP) == 0 && e < NN) { S[e] = -tot; if (!packed) St[(e % NCL) * NCL + e / NCL] = -tot; }
  }
  if (REST > 0) {  // NCL = 6: elements 32..35
    constexpr int RP = 4;
    double v[RP];
#pragma unroll
    for (int i = 0; i < RP; ++i) v[i] = (REST > 0 && 32 + i < NN) ? acc[(32 + i) < NN ? 32 + i : 0] : 0.0;
    const double tot = warp_reduce_scatter<RP>(v, lane);
    const int e = 32 + lane / (32 / RP);
    if (lane % (32 / RP) == 0 && e < NN) { S[e] = -tot; if (!packed) St[(e % NCL) * NCL + e / NCL] = -tot; }
  }
}

// A SUB-warp per block (GROUP = 16 or 8 lanes: two or four blocks per warp).  More blocks in flight per SM hide the gather latency
// better than a whole warp per block even at ~100 pairs per block (cfg 4 on one GPU: 155 -> 145 us with half-warps), and on sharded
// problems the pairs of a block are spread over the ranks (~25 per block and rank at 4 ranks, ~12 at 8), where a whole warp leaves most
// lanes idle.  NCL = 4 only (16 entries: one or two per lane).
template <int NCL, int GROUP>
__global__ void __launch_bounds__(256, PTZ_OD_MINB) k_schur_offdiag_sub(int nlist, const int* __restrict__ ub_list, const int64_t* __restrict__ pair_off,
                                                           const int* __restrict__ pair_a, const int* __restrict__ pair_b,
                                                           const double* __restrict__ What, const int* __restrict__ ub_pos,
                                                           const int* __restrict__ ub_pos_t, double* __restrict__ Sval) {
  static_assert(NCL == 4 && (GROUP == 16 || GROUP == 8), "16 block entries over a half or a quarter warp");
  typedef Dims<NCL> D;
  const int li = blockIdx.x * (blockDim.x / GROUP) + (threadIdx.x / GROUP);
  const int lane = threadIdx.x & 31, gl = lane & (GROUP - 1);
  const bool active = li < nlist;  // (no early return: all groups of a warp take part in the shuffles below)
  const int b = active ? (ub_list != nullptr ? ub_list[li] : li) : 0;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0;
  const int64_t ke = active ? pair_off[b + 1] : 0;
  int64_t k = active ? pair_off[b] + gl : 0;
  int ia = 0, ib = 0;
  if (k < ke) { ia = pair_a[k]; ib = pair_b[k]; }
  for (; k < ke; k += GROUP) {
    const double* wa = What + (size_t)ia * D::WS;
    const double* wb = What + (size_t)ib * D::WS;
    if (k + GROUP < ke) { ia = pair_a[k + GROUP]; ib = pair_b[k + GROUP]; }
    double x[D::WS], y[D::WS];
#pragma unroll
    for (int i = 0; i < D::WS / 4; ++i) ld256(wa + 4 * i, x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
#pragma unroll
    for (int i = 0; i < D::WS / 4; ++i) ld256(wb + 4 * i, y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
#pragma unroll
    for (int a = 0; a < NCL; ++a)
#pragma unroll
      for (int c = 0; c < NCL; ++c) acc[a * NCL + c] += x[3 * a] * y[3 * c] + x[3 * a + 1] * y[3 * c + 1] + x[3 * a + 2] * y[3 * c + 2];
  }
  // reduce-scatter inside each group: lane gl ends up with the totals of entries [gl * 16/GROUP, (gl + 1) * 16/GROUP)
  WarpRS<16, GROUP / 2>::run(acc, lane);
  if (active) {
    constexpr int PER = 16 / GROUP;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int e = gl * PER + u;
      if (ub_pos == nullptr) { Sval[(size_t)b * 16 + e] = -acc[u]; continue; }  // packed layout: mirrored after the all-reduce
      Sval[(size_t)ub_pos[b] * 16 + e] = -acc[u];
      Sval[(size_t)ub_pos_t[b] * 16 + (e % NCL) * NCL + e / NCL] = -acc[u];
    }
  }
}

// Sharded problem: the ranks' pieces of S travel through the all-reduce PACKED -- diagonal blocks, then the upper blocks once --
// instead of in the block-CSR layout that stores both triangles: half the bytes on the wire (92 -> 46 MB at 8 ranks).  One thread
// per entry puts them (and the transposes) into place afterwards.
template <int NCL>
__global__ void k_unpack_S(int V, int nub, const double* __restrict__ pack, const int* __restrict__ diag_pos, const int* __restrict__ ub_pos,
                           const int* __restrict__ ub_pos_t, double* __restrict__ Sval) {
  constexpr int NB = NCL * NCL;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t blk = idx / NB;
  const int e = (int)(idx % NB);
  if (blk >= (size_t)(V + nub)) return;
  const double v = pack[idx];
  if (blk < (size_t)V) { Sval[(size_t)diag_pos[blk] * NB + e] = v; return; }
  const int b = (int)(blk - V);
  Sval[(size_t)ub_pos[b] * NB + e] = v;
  Sval[(size_t)ub_pos_t[b] * NB + (e % NCL) * NCL + e / NCL] = v;
}

// ---- block-Jacobi preconditioning applied as a SYMMETRIC SCALING of the reduced system --------------------------------
// With S_cc = L_c L_c^T:  S~ = Linv S Linv^T (unit diagonal blocks), b~ = Linv b, y = Linv^T y~.  Plain CG on S~ has exactly
// the iterates of block-Jacobi PCG on S, but no preconditioner application inside the iteration.
//
// k_precond: Linv_c (explicit inverse of the Cholesky factor of every diagonal block)
template <int NCL>
__global__ void k_precond(int V, const int* __restrict__ diag_pos, const double* __restrict__ Sval, double* __restrict__ Linv, int* __restrict__ fail,
                          double* __restrict__ Lfac /* the factor itself (deflation basis scaling), or nullptr */) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double A[NCL * NCL];
  const double* S = Sval + (size_t)diag_pos[v] * NCL * NCL;
#pragma unroll
  for (int i = 0; i < NCL * NCL; ++i) A[i] = S[i];
  double* M = Linv + (size_t)v * NCL * NCL;
  if (!chol_n(A, NCL, NCL)) {
    atomicExch(fail, 2);
    for (int i = 0; i < NCL * NCL; ++i) M[i] = (i / NCL == i % NCL) ? 1.0 : 0.0;
    if (Lfac) for (int i = 0; i < NCL * NCL; ++i) Lfac[(size_t)v * NCL * NCL + i] = (i / NCL == i % NCL) ? 1.0 : 0.0;
    return;
  }
  if (Lfac) {
#pragma unroll
    for (int i = 0; i < NCL * NCL; ++i) Lfac[(size_t)v * NCL * NCL + i] = (i % NCL <= i / NCL) ? A[i] : 0.0;
  }
  // X = L^-1 by forward substitution, column by column (lower triangular)
  for (int c = 0; c < NCL; ++c) {
    double x[NCL];
    for (int i = 0; i < NCL; ++i) {
      double sacc = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) sacc -= A[i * NCL + k] * x[k];
      x[i] = (i < c) ? 0.0 : sacc / A[i * NCL + i];
    }
    for (int i = 0; i < NCL; ++i) M[i * NCL + c] = x[i];
  }
}
// border block (nb <= kMaxBorder = 16): Cholesky in shared memory, one lane per column of L^-1
__global__ void __launch_bounds__(32) k_precond_border(int nb, const double* __restrict__ Sbb, double* __restrict__ Linv_b, int* __restrict__ fail) {
  __shared__ double A[kMaxBorder * kMaxBorder];
  __shared__ int ok;
  for (int i = threadIdx.x; i < nb * nb; i += 32) A[i] = Sbb[i];
  __syncwarp();
  if (threadIdx.x == 0) { ok = chol_n(A, nb, nb) ? 1 : 0; if (!ok) atomicExch(fail, 3); }
  __syncwarp();
  const int c = threadIdx.x;
  if (c < nb) {
    if (!ok) {
      for (int i = 0; i < nb; ++i) Linv_b[i * nb + c] = (i == c) ? 1.0 : 0.0;
    } else {
      double x[kMaxBorder];
      for (int i = 0; i < nb; ++i) {
        double sacc = (i == c) ? 1.0 : 0.0;
        for (int k = c; k < i; ++k) sacc -= A[i * nb + k] * x[k];
        x[i] = (i < c) ? 0.0 : sacc / A[i * nb + i];
      }
      for (int i = 0; i < nb; ++i) Linv_b[i * nb + c] = x[i];
    }
  }
}

// CG state: per row three vectors (r, w = S~ r, s = S~ p), interleaved so that one neighbour is one contiguous gather:
//   cams  : st[(c*3 + comp)*NCL + j]        border: st[3*V*NCL + comp*nb + j]
// k_scale_system: one thread per block of S: B <- Linv_r B Linv_c^T (diagonal blocks become exactly I); the first V threads
// also scale the right-hand side and write the initial CG state (r = b~, w = s = 0, x = p = 0).
template <int NCL>
__global__ void k_scale_system(int V, int nnzb, const int* __restrict__ blk_row, const int* __restrict__ col, const double* __restrict__ Linv,
                               double* __restrict__ Sval, const double* __restrict__ rhs, double* __restrict__ st0, double* __restrict__ x,
                               double* __restrict__ p, const unsigned char* __restrict__ row_owner, int my_rank) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  // sharded CG: a rank only ever reads the rows of S~ it owns (my_rank < 0: all rows)
  if (k < nnzb && (my_rank < 0 || row_owner[blk_row[k]] == my_rank)) {
    const int r = blk_row[k], c = col[k];
    double* B = Sval + (size_t)k * NCL * NCL;
    if (r == c) {
#pragma unroll
      for (int i = 0; i < NCL * NCL; ++i) B[i] = (i / NCL == i % NCL) ? 1.0 : 0.0;
    } else {
      double Lr[NCL * NCL], Lc[NCL * NCL], T[NCL * NCL], Bv[NCL * NCL];
#pragma unroll
      for (int i = 0; i < NCL * NCL; ++i) { Lr[i] = Linv[(size_t)r * NCL * NCL + i]; Lc[i] = Linv[(size_t)c * NCL * NCL + i]; Bv[i] = B[i]; }
#pragma unroll
      for (int i = 0; i < NCL; ++i)
#pragma unroll
        for (int j = 0; j < NCL; ++j) {
          double sacc = 0;
#pragma unroll
          for (int q = 0; q <= i; ++q) sacc += Lr[i * NCL + q] * Bv[q * NCL + j];
          T[i * NCL + j] = sacc;
        }
#pragma unroll
      for (int i = 0; i < NCL; ++i)
#pragma unroll
        for (int j = 0; j < NCL; ++j) {
          double sacc = 0;
#pragma unroll
          for (int q = 0; q <= j; ++q) sacc += T[i * NCL + q] * Lc[j * NCL + q];
          B[i * NCL + j] = sacc;
        }
    }
  }
  if (k < V) {
    double b[NCL];
#pragma unroll
    for (int i = 0; i < NCL; ++i) b[i] = rhs[k * NCL + i];
#pragma unroll
    for (int i = 0; i < NCL; ++i) {
      double sacc = 0;
#pragma unroll
      for (int q = 0; q <= i; ++q) sacc += Linv[(size_t)k * NCL * NCL + i * NCL + q] * b[q];
      st0[(k * 3 + 0) * NCL + i] = sacc; st0[(k * 3 + 1) * NCL + i] = 0.0; st0[(k * 3 + 2) * NCL + i] = 0.0;
      x[k * NCL + i] = 0.0; p[k * NCL + i] = 0.0;
    }
  }
}
// border part of the scaling: strips C_k <- Linv_{c_k} C_k Linv_b^T, b_b <- Linv_b b_b, initial state of the border row
template <int NCL>
__global__ void __launch_bounds__(128) k_scale_border(int V, int nb, int nav, const int* __restrict__ ann_view, const double* __restrict__ Linv,
                                                      const double* __restrict__ Linv_b, const double* __restrict__ C, double* __restrict__ Cs,
                                                      const double* __restrict__ rhs, double* __restrict__ st0, double* __restrict__ x, double* __restrict__ p) {
  for (int e = threadIdx.x; e < nav * NCL * nb; e += blockDim.x) {
    const int k = e / (NCL * nb), i = (e / nb) % NCL, j = e % nb;
    const double* Lr = Linv + (size_t)ann_view[k] * NCL * NCL;
    const double* Ck = C + (size_t)k * NCL * nb;
    double sacc = 0;
    for (int q = 0; q <= i; ++q) {
      double t = 0;
      for (int m = 0; m <= j; ++m) t += Ck[q * nb + m] * Linv_b[j * nb + m];
      sacc += Lr[i * NCL + q] * t;
    }
    Cs[e] = sacc;
  }
  const int boff = V * NCL;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    double sacc = 0;
    for (int q = 0; q <= i; ++q) sacc += Linv_b[i * nb + q] * rhs[boff + q];
    st0[3 * boff + 0 * nb + i] = sacc; st0[3 * boff + 1 * nb + i] = 0.0; st0[3 * boff + 2 * nb + i] = 0.0;
    x[boff + i] = 0.0; p[boff + i] = 0.0;
  }
}
// y = Linv^T y~ : back to the unscaled unknowns
template <int NCL>
__global__ void k_unscale(int V, int nb, const double* __restrict__ Linv, const double* __restrict__ Linv_b, const double* __restrict__ xs,
                          double* __restrict__ y) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V) {
#pragma unroll
    for (int i = 0; i < NCL; ++i) {
      double sacc = 0;
#pragma unroll
      for (int q = i; q < NCL; ++q) sacc += Linv[(size_t)v * NCL * NCL + q * NCL + i] * xs[v * NCL + q];
      y[v * NCL + i] = sacc;
    }
  } else if (v < V + nb) {
    const int i = v - V, boff = V * NCL;
    double sacc = 0;
    for (int q = i; q < nb; ++q) sacc += Linv_b[q * nb + i] * xs[boff + q];
    y[boff + i] = sacc;
  }
}

// -------------------------------------------------------------------------------------------------------------
// stage 3 for SMALL reduced systems (n = V*NCL + nb <= kDenseMaxN: BASELINE cfg 1 / cfg 2, every BA of an IBA run): a dense
// Cholesky factorisation in ONE CTA instead of the CG.  At this size the CG is pure latency -- ~100 iterations of a grid-wide
// barrier each, 400 us per solve at V = 36 -- while the whole matrix fits the shared memory of one SM (n <= 160).  It is also
// what SPARSE_SCHUR does in the reference: an exact factorisation of the reduced camera system (ptzray_optimizer.cc:471).
//   k_dense_assemble: lower triangle of S (blocks, border strips, border block) and the right-hand side as row n of A[(n+1) x n]
//   k_dense_chol    : the matrix is copied into shared memory once (odd row pitch: no bank conflicts) and factored there by a
//                     blocked right-looking Cholesky, NB = 16: diagonal block factored by one warp in registers, panel rows (one
//                     thread each, registers) X = A L^-T written in place, trailing update straight from the panel columns (a warp
//                     = 4 rows, a lane = 4 columns: 8 LDS per 16 FMA).  A first version kept the matrix in global memory: the
//                     read-modify-write latency of the trailing update made it 4x slower (161 us at n = 144, no gain at n = 300).  Carrying b as an extra ROW makes z = L^-1 b fall out of the panel
//                     steps; L^T y = z is then solved block by block from the bottom.  Fixed order: bit-reproducible.
// -------------------------------------------------------------------------------------------------------------
constexpr int kDenseMaxN = 160;  // (n + 1) x (n | 1) doubles of shared memory: 207 KB at n = 160
template <int NCL>
__global__ void k_dense_assemble(int V, int nb, int nnzb, const int* __restrict__ blk_row, const int* __restrict__ col, const double* __restrict__ Sval,
                                 const double* __restrict__ rhs, int ncpl, const int* __restrict__ cpl_view, const double* __restrict__ C,
                                 const double* __restrict__ Sbb, double* __restrict__ A) {
  const int n = V * NCL + nb, ld = n;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnzb) {
    const int r = blk_row[k], c = col[k];
    if (c <= r) {
      const double* B = Sval + (size_t)k * NCL * NCL;
#pragma unroll
      for (int i = 0; i < NCL; ++i)
#pragma unroll
        for (int j = 0; j < NCL; ++j) A[(size_t)(r * NCL + i) * ld + c * NCL + j] = B[i * NCL + j];
    }
    return;
  }
  int e = k - nnzb;
  if (e < ncpl * NCL * nb) {  // border rows x camera columns: transposed strips
    const int q = e / (NCL * nb), a = (e / nb) % NCL, j = e % nb;
    A[(size_t)(V * NCL + j) * ld + cpl_view[q] * NCL + a] = C[e];
    return;
  }
  e -= ncpl * NCL * nb;
  if (e < nb * nb) { A[(size_t)(V * NCL + e / nb) * ld + V * NCL + e % nb] = Sbb[e]; return; }
  e -= nb * nb;
  if (e < n) A[(size_t)n * ld + e] = rhs[e];
}
#ifndef PTZ_DENSE_THREADS
#define PTZ_DENSE_THREADS 256
#endif
constexpr int kDenseThreads = PTZ_DENSE_THREADS;
constexpr int kDenseNB = 16, kDenseLd = kDenseNB + 1;  // block size: the serial chains (diagonal factor, row substitution) grow with NB^2
__global__ void __launch_bounds__(kDenseThreads) k_dense_chol(int n, const double* __restrict__ Ag, double* __restrict__ y, int* __restrict__ info,
                                                              int* __restrict__ fail) {
  constexpr int NB = kDenseNB, LD = kDenseLd;
  extern __shared__ double dsm_chol[];
  const int ldp = n | 1;                        // odd pitch: the rows a warp reads side by side fall into different banks
  double* Ld = dsm_chol;                        // [NB][LD] diagonal block (identity-padded)
  double* red = Ld + NB * LD;                   // [<= 32 warps][NB]
  double* yv = red + 32 * NB;                   // [n]
  double* rdiag = yv + ((n + 3) & ~3);          // [n] 1 / L[i][i]: the substitutions multiply instead of paying a dependent divide per step
  double* A = rdiag + ((n + 3) & ~3);           // [(n + 1)][ldp]: lower triangle of S, the right-hand side as row n
  __shared__ int s_ok;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nw = blockDim.x >> 5;
  if (t == 0) s_ok = 1;
  for (int r = wid; r <= n; r += nw) {  // a row per warp, coalesced, four loads in flight
    const double* src = Ag + (size_t)r * n;
    double* dst = A + r * ldp;
    int c = lane;
    for (; c + 96 < n; c += 128) {
      const double v0 = src[c], v1 = src[c + 32], v2 = src[c + 64], v3 = src[c + 96];
      dst[c] = v0; dst[c + 32] = v1; dst[c + 64] = v2; dst[c + 96] = v3;
    }
    for (; c < n; c += 32) dst[c] = src[c];
  }
  __syncthreads();
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int bs = min(NB, n - j0);
    if (wid == 0) {
      // the NB x NB block in registers, one row per lane (lanes >= NB idle); column c of L leaves by shuffles: no shared-memory
      // round trips in the chain.  Constant loop bounds + predicates so that everything unrolls and a[] stays in registers.
      const int row = lane < bs ? lane : -1;
      double a[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = (row >= 0 && c < bs) ? (c <= row ? A[(j0 + row) * ldp + j0 + c] : 0.0) : (lane == c ? 1.0 : 0.0);
      bool ok = true;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const double dcc = __shfl_sync(0xffffffffu, a[c], c);
        const bool good = dcc > 0.0 && isfinite(dcc);  // (uniform)
        ok = ok && good;
        const double dinv = good ? rsqrt(dcc) : 1.0, d = good ? dcc * dinv : 1.0;
        const double l = lane > c ? a[c] * dinv : (lane == c ? d : 0.0);
        a[c] = l;
        if (lane == c && c < bs) rdiag[j0 + c] = dinv;
#pragma unroll
        for (int cc = 0; cc < NB; ++cc) {
          if (cc > c) {
            const double lcc = __shfl_sync(0xffffffffu, l, cc);  // L[cc][c]
            if (lane >= cc) a[cc] -= l * lcc;
          }
        }
      }
      if (!ok && lane == 0) s_ok = 0;
      if (lane < NB) {
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          Ld[lane * LD + c] = a[c];
          if (lane < bs && c <= lane) A[(j0 + lane) * ldp + j0 + c] = a[c];
        }
      }
    }
    __syncthreads();
    if (!s_ok) break;
    // panel: rows below the block, the right-hand side row (row n) last: X = A L^-T in place, one thread per row
    const int m = n - j0 - bs;
    for (int pr = t; pr <= m; pr += blockDim.x) {
      double* arow = A + (j0 + bs + pr) * ldp + j0;
      double x[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) x[c] = c < bs ? arow[c] : 0.0;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        double s0 = x[c], s1 = 0.0;  // two chains
#pragma unroll
        for (int q = 0; q < NB; q += 2) {
          if (q < c) s0 -= x[q] * Ld[c * LD + q];
          if (q + 1 < c) s1 -= x[q + 1] * Ld[c * LD + q + 1];
        }
        x[c] = (s0 + s1) * (c < bs ? rdiag[j0 + c] : 1.0);
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) if (c < bs) arow[c] = x[c];
    }
    __syncthreads();
    // trailing update: A[r][c] -= X_r . X_c for c <= r < m, and the right-hand side row (r == m) for all c < m.
    // A warp = 4 rows, a lane = 4 columns (c0 + lane + 32 b): 8 LDS per 16 FMA
    const double* X = A + (j0 + bs) * ldp + j0;  // panel: X[row * ldp + c], c < bs
    double* T = A + (j0 + bs) * ldp + j0 + bs;   // trailing matrix
    for (int r0 = wid * 4; r0 <= m; r0 += nw * 4) {
      const int rmax = min(r0 + 3, m);
      const int cend = rmax == m ? m : rmax + 1;  // columns [0, cend)
      for (int c0 = 0; c0 < cend; c0 += 128) {
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        int pc[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) pc[b] = min(c0 + lane + 32 * b, m);  // (clamped lanes read a valid row and discard)
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          if (c < bs) {
            double pv[4], qv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) pv[a] = X[min(r0 + a, m) * ldp + c];
#pragma unroll
            for (int b = 0; b < 4; ++b) qv[b] = X[pc[b] * ldp + c];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) acc[a][b] += pv[a] * qv[b];
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int r = r0 + a;
          if (r > m) continue;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int c = c0 + lane + 32 * b;
            if (c < m && (c <= r || r == m)) T[r * ldp + c] -= acc[a][b];
          }
        }
      }
    }
    __syncthreads();
  }
  if (!s_ok) {
    for (int i = t; i < n; i += blockDim.x) y[i] = 0.0;
    if (t == 0) { atomicExch(fail, 2); info[0] = 0; info[1] = 0; }
    return;
  }
  // L^T y = z, z = row n of A; blocks from the bottom
  const int last = ((n - 1) / NB) * NB;
  for (int j0 = last; j0 >= 0; j0 -= NB) {
    const int bs = min(NB, n - j0);
    double part = 0.0;
    if (lane < bs)
      for (int i = j0 + bs + wid; i < n; i += nw) part += A[i * ldp + j0 + lane] * yv[i];
    if (lane < NB) red[wid * NB + lane] = part;
    __syncthreads();
    if (wid == 0) {
      double tc = 0.0;
      if (lane < bs) {
        tc = A[n * ldp + j0 + lane];
        for (int w = 0; w < nw; ++w) tc -= red[w * NB + lane];
      }
#pragma unroll
      for (int r = NB - 1; r >= 0; --r) {
        if (r < bs) {
          const double yr = __shfl_sync(0xffffffffu, tc, r) * rdiag[j0 + r];
          if (lane == r) tc = yr;
          else if (lane < r) tc -= A[(j0 + r) * ldp + j0 + lane] * yr;
        }
      }
      if (lane < bs) { yv[j0 + lane] = tc; y[j0 + lane] = tc; }
    }
    __syncthreads();
  }
  if (t == 0) { info[0] = 1; info[1] = 0; }
}

// -------------------------------------------------------------------------------------------------------------
// stage 3: conjugate gradients on the scaled reduced system, ONE cooperative launch per linear solve and ONE grid-wide
// barrier per iteration.  Chronopoulos-Gear recurrences (single fused reduction of (r,r) and (S~r, r)):
//     p = r + beta p ; s = w + beta s ; x += alpha p ; r' = r - alpha s ; w' = S~ r' ; gamma' = (r',r') ; delta = (w',r')
//     beta' = gamma'/gamma ; alpha' = gamma' / (delta - beta' gamma'/alpha)
// A neighbour's r' is recomputed on the fly from its PREVIOUS (r, w, s) (one contiguous gather), so no barrier is needed
// between the vector update and the sparse product; the state is ping-ponged so readers never race the owner.
//
// Multi-GPU (MULTI): the ROWS are sharded across the ranks of one NVLink/NVSwitch box; W kernels (one per GPU) run the same
// iteration in lock step and exchange through peer-mapped arenas (CUDA IPC) -- sparse product, vector update, exchange and
// reduction are ONE kernel, there is no NCCL call inside a linear solve.  The exchange is latency-bound (a few hundred
// bytes per row), so it uses self-validating "LL" words instead of fence + flag: a double travels as two 8-byte words, each
// holding 32 data bits and the 32-bit tag of the iteration it belongs to.  The receiver polls the DATA until both tags
// match: one NVLink hop, no release fence waiting for write acknowledgements, no atomics.
//   * state: the owner of a row keeps it in its local plain replica (read by the CTAs of the same GPU through L1) and
//     pushes it as LL words into the inbox of every rank that owns a neighbouring row (peer_mask);
//   * reduction + barrier: every CTA pushes its partial (r,r), (S~r,r) as LL words into its slot on every rank; warp 0 of
//     every CTA polls all W x grid slots and sums them in slot order -- bit-identical totals on every rank, which therefore
//     take the same decisions and leave together.  State and slots are double-buffered by iteration parity.
// PTZ_CG_VRANKS=k (debug) runs the same protocol on ONE GPU with the grid split into k virtual ranks.
// -------------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
// Deflation (KD > 0): a small basis W (KD vectors: Ritz vectors harvested from the residual history of the first solve of a run)
// is projected out of the Krylov space -- the handful of tiny eigenvalues of S~ (global gauge rotations and the smooth bending
// modes of the view graph) are what the plain iteration spends most of its steps on.  Saad/Yeung/Erhel/Guyomarc'h's deflated
// CG in the single-reduction form:
//     mu = E^-1 (AW)^T r            E = W^T S~ W   (KD x KD, factored once per solve)
//     p  = r + beta p - W mu ;  s = S~ p = w + beta s - (AW) mu ;  x += alpha p ;  r' = r - alpha s
//     (p, S~ p) = delta - beta gamma / alpha_prev - mu^T nu         nu = (AW)^T r rides in the ONE fused reduction (KD + 2 values)
// The single grid barrier per iteration survives: a neighbour's r' is still recomputed from its OLD (r, w, s) alone, and the
// part of it that depends on mu is added by the owner of the row through Z = S~ (S~ W), precomputed once per solve:
//     w'_i = sum_j B_ij [ r_j - alpha (w_j + beta s_j) ]  +  alpha (Z mu)_i
// (validated statement for statement in tests/scripts/deflated_cg_kernel_model.py).
constexpr int kDeflK = 16;
constexpr int kCgMaxVals = 2 + kDeflK;
constexpr int kCgMaxCtas = 192;  // CTAs per rank at most (one per SM): bounds the shared-memory copy of the reduction partials
struct CgArgs {
  int V, nb, n;            // n = V*NCL + nb
  const int* rowptr; const int* col; const double* Sval;   // col: bit 31 set = the column is owned by another rank than the row
  int nav; const int* ann_view; const int* ann_idx; const double* C;   // scaled coupling strips [nav][NCL][nb]
  const int* order;        // slot -> row: consecutive slots are neighbouring views (Cuthill-McKee), one contiguous run per CTA
  const unsigned char* peer_mask;  // per row: ranks (bits) owning a neighbour of the row, i.e. that need its state
  // arena replicas (index = rank; [rank] is local).  Same layout everywhere: ctrl | slots | st0 | st1 | x | ll0 | ll1
  char* arena[kMaxPeers];
  size_t off_partial, off_st0, off_st1, off_x, off_ll0, off_ll1;
  int W, rank;             // ranks sharing the rows; rank r owns slots [r*slots_per_rank, (r+1)*slots_per_rank)
  int vranks;              // > 1: virtual ranks inside one launch (debug), rank = blockIdx / (gridDim / vranks)
  int slots_per_rank;
  double* p;               // search direction (every row has one owner)
  int smem_blocks;         // blocks of S (and their column indices) each warp keeps in shared memory for the whole solve
  int smem_defl_rows;      // rows per warp whose deflation data (W, AW, Z entries) live in shared memory as well
  int debug;               // timing experiments only (PTZ_CG_DEBUG): 1 = skip the sparse product, 2 = skip the grid barrier
  int max_iter; double tol;
  int* out_info;           // [0] iterations, [1] status (0 converged, 1 hit cap, 2 breakdown, 3 peer timeout, 4 deflated solve stagnated in the endgame)
  const double* gamma0_ptr;  // plain kernel warm-started from a deflated solve's x: |b~|^2 to measure the residual against (else nullptr)
  double* out_res;         // [0] |r~| / |b~|
  // deflation (KD > 0): W, AW = S~ W, Z = S~ AW as [n][KD]; Einv [KD*KD]; dscal = { |b~|^2, basis usable (1) or not (0) }
  const double* dW; const double* dAW; const double* dZ; const double* dEinv; const double* dscal;
  // residual history of an undeflated solve (harvested into the deflation basis by the host afterwards): hist [hist_cap][n],
  // abg [.. ][3] = alpha, beta, gamma of every iteration.  nullptr = off
  double* hist; int hist_cap; double* abg;
  double* prof;            // PTZ_CG_DEBUG & 8: cycles of CTA 0 / thread 0 per phase of an iteration, summed over the solve
};
// arena control block (u64 words): [0] barrier arrival counter of the single-GPU path (never reset), [1] arrivals consumed by
// the barriers passed so far, [2] next unused LL tag (the same on every rank; carried from solve to solve)
constexpr size_t kArenaCtrlBytes = 256;
// reduction slots: per-CTA partials [2][NV][G] doubles, then (multi-rank) LL rank totals [2][W][NV] x 16 bytes
__host__ __device__ constexpr size_t cg_slot_region_bytes(int ctas) {
  return 2 * (size_t)kCgMaxVals * ctas * sizeof(double) + 2 * (size_t)kMaxPeers * kCgMaxVals * 16;
}
constexpr long long kPeerTimeoutCycles = 40000000000ll;  // ~20 s

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// CTA-level part of both reductions: v[0], v[1] may live in any lane, v[2..] only in the first 8 lanes of a warp (the lanes
// that own a row entry).  Leaves the CTA totals of the NV values in sred[0..NV) (valid for warp 0 after its __syncwarp).
template <int NV>
__device__ __forceinline__ void cta_reduceN(double (&v)[NV], double* sred /* [nwarp][NV] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const double a = warp_sum(v[0]), b = warp_sum(v[1]);
  if (lane == 0) { sred[wid * NV] = a; sred[wid * NV + 1] = b; }
  if constexpr (NV > 2) {
    // reduce-scatter over the 8-lane group: 8 + 4 + 2 shuffles instead of 3 x 16; lane l ends up with elements e0, e0 + 1
    constexpr int KD = NV - 2;
    static_assert(KD == 16, "deflation width is 16");
    double u[KD];
#pragma unroll
    for (int i = 0; i < KD; ++i) u[i] = v[2 + i];
    WarpRS<KD, 4>::run(u, lane);
    const int e0 = ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
    if (lane < 8) { sred[wid * NV + 2 + e0] = u[0]; sred[wid * NV + 2 + e0 + 1] = u[1]; }
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int w = 0; w < nwarp; ++w) s += sred[w * NV + threadIdx.x];
    v[0] = s;  // thread c holds the CTA total of value c
  }
}
// Sums of the value-major partial array buf[NV][G] (G <= kCgMaxCtas): the NW warps of the CTA share the NV values between them
// and every lane issues ALL its loads (<= ceil(NV/NW) * kCgMaxCtas/32) before the first one is consumed -- one L2 round trip
// for the whole fetch.  Fixed order; the caller does something with (value index, total) in every lane of the owning warp.
template <int NV, int NW, class F>
__device__ __forceinline__ void sum_partials(const double* buf, int G, F&& emit) {
  constexpr int VPW = (NV + NW - 1) / NW, LPV = kCgMaxCtas / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double t[VPW][LPV];
#pragma unroll
  for (int q = 0; q < VPW; ++q) {
    const int c = wid + q * NW;
#pragma unroll
    for (int u = 0; u < LPV; ++u) {
      const int i = lane + 32 * u;
      t[q][u] = (c < NV && i < G) ? __ldcg(buf + (size_t)c * G + i) : 0.0;
    }
  }
#pragma unroll
  for (int q = 0; q < VPW; ++q) {
    const int c = wid + q * NW;
    double s = 0;
#pragma unroll
    for (int u = 0; u < LPV; ++u) s += t[q][u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (c < NV) emit(c, s);
  }
}
// single GPU: CTA partials -> value-major slots, counter barrier; then the warps share the NV sums between them (sum_partials)
template <int NV, int NW>
__device__ __forceinline__ void grid_reduceN(const CgArgs& A, double (&v)[NV], int parity, unsigned long long& arrived, double* sred, double* s_out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, G = gridDim.x;
  const bool profiling = A.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long tprof = profiling ? clock64() : 0;
#define PTZ_RED_PROF(slot)                                                       \
  if (profiling) { const long long tn = clock64(); A.prof[slot] += (double)(tn - tprof); tprof = tn; }
  cta_reduceN<NV>(v, sred);
  PTZ_RED_PROF(3)
  double* buf = reinterpret_cast<double*>(A.arena[0] + A.off_partial) + (size_t)(parity & 1) * NV * G;
  if (wid == 0) {
    if (lane < NV) __stcg(buf + (size_t)lane * G + blockIdx.x, v[0]);
    __syncwarp();  // the NV stores happen-before lane 0's release below
    if (lane == 0) {
      unsigned long long* bar = reinterpret_cast<unsigned long long*>(A.arena[0]);
      asm volatile("red.release.gpu.global.add.u64 [%0], 1;" :: "l"(bar) : "memory");
      const unsigned long long target = arrived + (unsigned long long)G;
      while (ld_acquire_u64(bar) < target) { }
    }
  }
  __syncthreads();
  PTZ_RED_PROF(4)
  arrived += (unsigned long long)G;
  // every thread acquires: its later plain (L1-cached) loads must not be served from lines older than this barrier
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  sum_partials<NV, NW>(buf, G, [&](int c, double tot) { if (lane == 0) s_out[c] = tot; });
  PTZ_RED_PROF(5)
  __syncthreads();
  PTZ_RED_PROF(6)
#undef PTZ_RED_PROF
}

// ---- LL words
__device__ __forceinline__ void ll_store(ulonglong2* dst, double v, unsigned int tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
  const unsigned long long lo = t | (bits & 0xffffffffull), hi = t | (bits >> 32);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" :: "l"(dst), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ bool ll_load(const ulonglong2* src, unsigned int tag, double& v) {
  unsigned long long lo, hi;
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
  v = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
  return (unsigned int)(lo >> 32) == tag && (unsigned int)(hi >> 32) == tag;
}
// multi-rank reduction + barrier, two levels.  Inside a rank: plain partials + acq_rel arrival counter; the LAST CTA to arrive
// fetches the rank's partials (all its warps, one L2 round trip), sums them in CTA order and pushes the NV rank totals as LL
// words into the rank's slots on every rank.  Across ranks: the first W x NV threads of every CTA poll one local slot each; the
// totals are added in rank order -- bit-identical on every CTA of every rank.  sys_release: make this CTA's earlier PLAIN stores
// to peer memory visible first (end-of-solve push of x).  Totals in s_out[0..NV); returns false on a peer timeout.
template <int NV, int NW>
__device__ __forceinline__ bool ll_reduceN(const CgArgs& A, double (&v)[NV], int parity, unsigned int tag, int my_rank, int cta, int G, int W,
                                           bool sys_release, unsigned long long& arrived, double* sred, double* s_out, double* s_part /* [kMaxPeers*NV] */,
                                           int* s_flag) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  cta_reduceN<NV>(v, sred);
  char* const mine = A.arena[my_rank];
  double* lbuf = reinterpret_cast<double*>(mine + A.off_partial) + (size_t)(parity & 1) * NV * G;
  const size_t roff = A.off_partial + 2 * (size_t)NV * G * sizeof(double) + (size_t)(parity & 1) * W * NV * 16;
  if (wid == 0) {
    if (lane < NV) __stcg(lbuf + (size_t)lane * G + cta, v[0]);
    __syncwarp();
    if (lane == 0) {
      unsigned long long old = 0;
      if (sys_release) asm volatile("fence.acq_rel.sys;" ::: "memory");
      // arrival: releases this CTA's writes, acquires those of earlier arrivers
      asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], 1;" : "=l"(old) : "l"(reinterpret_cast<unsigned long long*>(mine)) : "memory");
      s_flag[0] = (old + 1ull == arrived + (unsigned long long)G) ? 1 : 0;  // last CTA of this rank
      s_flag[1] = 1;                                                        // no timeout so far
    }
  }
  __syncthreads();
  if (s_flag[0]) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    sum_partials<NV, NW>(lbuf, G, [&](int c, double tot) {
      if (lane < W) ll_store(reinterpret_cast<ulonglong2*>(A.arena[lane] + roff) + (size_t)my_rank * NV + c, tot, tag);
    });
  }
  // all CTAs: wait for the W x NV rank totals, one slot per thread
  if ((int)threadIdx.x < W * NV) {
    const ulonglong2* slot = reinterpret_cast<const ulonglong2*>(mine + roff) + threadIdx.x;
    const long long t_begin = clock64();
    double val = 0;
    while (!ll_load(slot, tag, val)) {
      if (clock64() - t_begin > kPeerTimeoutCycles) { s_flag[1] = 0; break; }
    }
    s_part[threadIdx.x] = val;
  }
  __syncthreads();
  if ((int)threadIdx.x < NV) {
    double s = 0;
    for (int k = 0; k < W; ++k) s += s_part[k * NV + threadIdx.x];
    s_out[threadIdx.x] = s;
  }
  __syncthreads();
  arrived += (unsigned long long)G;
  // every thread acquires: its later plain (L1-cached) loads must not be served from lines older than this barrier
  if (sys_release) asm volatile("fence.acq_rel.sys;" ::: "memory");
  else asm volatile("fence.acq_rel.gpu;" ::: "memory");
  return s_flag[1] != 0;
}

// MAXT = CTA width (256 / 512 threads): wide CTAs keep one row per warp on larger systems (one CTA per SM either way; beyond
// 16 rows per SM a warp walks several rows).  KD = 0: plain CG; KD = kDeflK: deflated.
template <int NCL, int MAXT, bool MULTI, int KD>
__global__ void __launch_bounds__(MAXT) k_cg(CgArgs A) {
  constexpr int SLOTS = 32 / NCL, NB = NCL * NCL, NV = 2 + KD;
  extern __shared__ double cg_smem[];
  __shared__ double sred[(MAXT / 32) * NV];
  __shared__ double s_out[NV + 1];
  __shared__ double s_part[MULTI ? kMaxPeers * NV : 1];
  __shared__ double s_einv[KD > 0 ? KD * KD : 1];
  __shared__ int s_flag[2];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;
  const int la = lane % NCL, ls = lane / NCL;
  const bool lact = lane < SLOTS * NCL;
  const int V = A.V, nb = A.nb, nrows = V + (nb > 0 ? 1 : 0), boff = V * NCL;
  const int W = MULTI ? A.W : 1;
  // rank of this CTA, its index inside the rank and the CTAs per rank (virtual ranks split one launch)
  const int G = (MULTI && A.vranks > 1) ? (int)gridDim.x / A.vranks : (int)gridDim.x;
  const int my_rank = MULTI ? (A.vranks > 1 ? (int)blockIdx.x / G : A.rank) : 0;
  const int cta = (MULTI && A.vranks > 1) ? (int)blockIdx.x % G : (int)blockIdx.x;
  char* const mine = A.arena[my_rank];
  // ---- the rows of a warp are the same in every iteration: keep (a prefix of) their S blocks and column indices on chip
  const int cap = A.smem_blocks;
  double* Bs = cg_smem + (size_t)wid * cap * NB;
  int* cs = reinterpret_cast<int*>(cg_smem + (size_t)wpb * cap * NB) + (size_t)wid * cap;
  // this rank's slots [slot0, slot1); inside, slots [cta*per, (cta+1)*per) belong to this CTA; warp w takes every wpb-th
  const int slot0 = my_rank * A.slots_per_rank, slot1 = min(nrows, slot0 + A.slots_per_rank);
  const int per = (A.slots_per_rank + G - 1) / G;
  const int cta0 = slot0 + cta * per;
  int ncached = 0;
  for (int sl = wid; sl < per && ncached < cap; sl += wpb) {
    const int slot = cta0 + sl;
    if (slot >= min(V, slot1)) break;
    const int row = A.order[slot];
    const int b0 = A.rowptr[row], take = min(A.rowptr[row + 1] - b0, cap - ncached);
    for (int e = lane; e < take * NB; e += 32) Bs[(size_t)ncached * NB + e] = A.Sval[(size_t)b0 * NB + e];
    for (int e = lane; e < take; e += 32) cs[ncached + e] = A.col[b0 + e];
    ncached += take;
  }
  bool defl = false;
  // deflation data of the first `dcap` rows of every warp: [row][W | AW | Z][KD][NCL] (an entry's KD values at stride NCL:
  // the NCL owner lanes read consecutive words)
  const int dcap = KD > 0 ? A.smem_defl_rows : 0;
  double* Ds = cg_smem + (((size_t)wpb * cap * (NB * sizeof(double) + sizeof(int)) + 15) / 16) * 2 + (size_t)wid * dcap * 3 * KD * NCL;
  if (KD > 0) {
    defl = A.dscal[1] != 0.0;  // the basis passed the Cholesky of E (else: plain CG from x = 0)
    for (int e = threadIdx.x; e < KD * KD; e += blockDim.x) s_einv[e] = A.dEinv[e];
    int ri = 0;
    for (int sl = wid; sl < per && ri < dcap; sl += wpb, ++ri) {
      const int slot = cta0 + sl;
      if (slot >= min(V, slot1)) break;
      const size_t g0 = (size_t)A.order[slot] * NCL * KD;
      for (int e = lane; e < NCL * KD; e += 32) {
        const int a = e / KD, dd = e % KD;
        double* q = Ds + (size_t)ri * 3 * KD * NCL + dd * NCL + a;
        q[0] = A.dW[g0 + e]; q[KD * NCL] = A.dAW[g0 + e]; q[2 * KD * NCL] = A.dZ[g0 + e];
      }
    }
    __syncthreads();
  }
  __syncwarp();
  unsigned long long* ctrl = reinterpret_cast<unsigned long long*>(mine);
  unsigned long long arrived = ctrl[1];             // single GPU: arrivals consumed so far on this arena
  const unsigned int tag0 = (unsigned int)ctrl[2];  // multi: first LL tag of this solve (same on every rank)
  double alpha = 0.0, beta = 0.0, gamma_old = 0.0, gamma0 = 0.0, gamma_last = 0.0, g_best = 1.7976931348623157e308;
  int it_best = 0;
  double mus[KD > 0 ? KD : 1];  // mu, replicated in every lane
#pragma unroll
  for (int dd = 0; dd < (KD > 0 ? KD : 1); ++dd) mus[dd] = 0.0;
  size_t off_o = A.off_st0, off_n = A.off_st1;  // previous state (read by everyone) / next state (written by the owner only)
  double* const xl = reinterpret_cast<double*>(mine + A.off_x);
  int it = 0, status = 1;
  bool peers_ok = true;
  const bool profiling = A.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long tprof = profiling ? clock64() : 0;
#define PTZ_CG_PROF(slot)                                                        \
  if (profiling) { const long long tn = clock64(); A.prof[slot] += (double)(tn - tprof); tprof = tn; }
  for (;; ++it) {
    PTZ_CG_PROF(7)
    const double* so = reinterpret_cast<const double*>(mine + off_o);
    double* const sn = reinterpret_cast<double*>(mine + off_n);
    // LL inboxes: state(it-1) was pushed with tag0 + it into buffer (it-1)&1; state(it) goes out with tag0 + it + 1 into it&1
    const ulonglong2* llo = reinterpret_cast<const ulonglong2*>(mine + ((it & 1) ? A.off_ll0 : A.off_ll1));
    const size_t off_lln = (it & 1) ? A.off_ll1 : A.off_ll0;
    const unsigned int tag_rd = tag0 + (unsigned int)it, tag_wr = tag_rd + 1u;
    const bool use_ll = MULTI && it > 0;  // iteration 0 reads the initial state, which every rank computed for all rows
    double acc[NV];  // acc[0] = (r',r'), acc[1] = (w',r'), acc[2..] = (AW)^T r'
#pragma unroll
    for (int c = 0; c < NV; ++c) acc[c] = 0.0;
    int bi = 0;  // running index of this warp's blocks
    int ri = -1; // running index of this warp's rows
    for (int sl = wid; sl < per; sl += wpb) {
      const int slot = cta0 + sl;
      if (slot >= slot1) break;
      const int row = slot < V ? A.order[slot] : V;
      ++ri;
      if (row < V) {
        double rn = 0, s_n = 0, zm = 0;
        unsigned int pmask = 0;
        if (MULTI) pmask = A.peer_mask[row];
        if (lane < NCL) {
          const int i = row * NCL + lane;
          const double ro = __ldcg(so + (row * 3 + 0) * NCL + lane), wo = __ldcg(so + (row * 3 + 1) * NCL + lane), s_o = __ldcg(so + (row * 3 + 2) * NCL + lane);
          double pn = ro + beta * A.p[i];
          s_n = wo + beta * s_o;
          if (KD > 0) {
            // deflation terms of this row entry: (W mu)_i, (AW mu)_i, (Z mu)_i
            double wm = 0, awm = 0;
            if (ri < dcap) {
              const double* q = Ds + (size_t)ri * 3 * KD * NCL + lane;
#pragma unroll
              for (int dd = 0; dd < KD; ++dd) {
                wm += q[dd * NCL] * mus[dd];
                awm += q[(KD + dd) * NCL] * mus[dd];
                zm += q[(2 * KD + dd) * NCL] * mus[dd];
              }
            } else {
              const double2* Wi = reinterpret_cast<const double2*>(A.dW + (size_t)i * KD);
              const double2* AWi = reinterpret_cast<const double2*>(A.dAW + (size_t)i * KD);
              const double2* Zi = reinterpret_cast<const double2*>(A.dZ + (size_t)i * KD);
#pragma unroll
              for (int dd = 0; dd < KD / 2; ++dd) {
                const double2 a = __ldg(Wi + dd), b = __ldg(AWi + dd), c = __ldg(Zi + dd);
                wm += a.x * mus[2 * dd] + a.y * mus[2 * dd + 1];
                awm += b.x * mus[2 * dd] + b.y * mus[2 * dd + 1];
                zm += c.x * mus[2 * dd] + c.y * mus[2 * dd + 1];
              }
            }
            pn -= wm;
            s_n -= awm;
          }
          A.p[i] = pn;
          xl[i] += alpha * pn;
          rn = ro - alpha * s_n;
          sn[(row * 3 + 0) * NCL + lane] = rn;
          sn[(row * 3 + 2) * NCL + lane] = s_n;
          if (A.hist != nullptr && it < A.hist_cap) A.hist[(size_t)it * A.n + i] = rn;
          if (MULTI) {  // r and s leave now, so that the remote stores drain behind the sparse product
            for (int k = 0; k < W; ++k)
              if ((pmask >> k) & 1u) {
                ulonglong2* q = reinterpret_cast<ulonglong2*>(A.arena[k] + off_lln) + (size_t)row * (3 * NCL) + lane;
                ll_store(q, rn, tag_wr);
                ll_store(q + 2 * NCL, s_n, tag_wr);
              }
          }
        }
        const int b0 = A.rowptr[row], nblk = A.rowptr[row + 1] - b0;
        double sum = 0;
        {
          // Every lane of a slot fetches ONE entry of the neighbour's (r, w, s) -- 3 loads, each 32-byte sector requested once --
          // forms its entry of r' and the NCL lanes of the slot exchange by shuffle.  Branch-free: lanes without a block read
          // their own row (valid memory) and discard; the raw loads of CH steps are issued before any is consumed.
#ifndef PTZ_CG_CH
#define PTZ_CG_CH 6
#endif
          constexpr int CH = PTZ_CG_CH;
          const int nsteps = (A.debug & 1) ? 0 : (nblk + SLOTS - 1) / SLOTS;
          const int lim = lact ? nblk : 0;
          const int base = ls * NCL;
          const bool cached = bi + nblk <= ncached;  // uniform over the warp
          const int* ci = cached ? (cs + bi) : nullptr;
          for (int t0 = 0; t0 < nsteps; t0 += CH) {
            int off[CH];
            bool rem[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) {
              const int k = (t0 + u) * SLOTS + ls;
              int c = row;
              if (k < lim) c = cached ? ci[k] : __ldg(A.col + b0 + k);
              rem[u] = use_ll && c < 0;  // neighbour owned by another rank: its state arrives in the LL inbox
              off[u] = (c & 0x7fffffff) * (3 * NCL) + la;
            }
            double rr[CH], ww[CH], ss[CH];
#pragma unroll
            // plain (L1-cached) loads: the rows of a CTA are neighbours and share most of their gathers; the acquire in the grid
            // barrier invalidates L1 every iteration, so a cached line is never older than the last barrier
            for (int u = 0; u < CH; ++u) {
              if (!MULTI || !rem[u]) { rr[u] = so[off[u]]; ww[u] = so[off[u] + NCL]; ss[u] = so[off[u] + 2 * NCL]; }
            }
            if (MULTI) {
              // remote neighbours: one batched attempt (the words normally arrived long ago), then poll the stragglers
              unsigned int bad = 0;
#pragma unroll
              for (int u = 0; u < CH; ++u) {
                if (rem[u]) {
                  const ulonglong2* q = llo + off[u];
                  if (!(ll_load(q, tag_rd, rr[u]) & ll_load(q + NCL, tag_rd, ww[u]) & ll_load(q + 2 * NCL, tag_rd, ss[u]))) bad |= 1u << u;
                }
              }
              if (bad) {
                const long long t_begin = clock64();
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                  if ((bad >> u) & 1u) {
                    const ulonglong2* q = llo + off[u];
                    while (!(ll_load(q, tag_rd, rr[u]) & ll_load(q + NCL, tag_rd, ww[u]) & ll_load(q + 2 * NCL, tag_rd, ss[u]))) {
                      if (clock64() - t_begin > kPeerTimeoutCycles) { peers_ok = false; break; }
                    }
                  }
                }
              }
            }
#pragma unroll
            for (int u = 0; u < CH; ++u) {
              const int k = (t0 + u) * SLOTS + ls;
              const double mine_r = rr[u] - alpha * (ww[u] + beta * ss[u]);
              double rj[NCL];
#pragma unroll
              for (int j = 0; j < NCL; ++j) rj[j] = __shfl_sync(0xffffffffu, mine_r, (base + j) & 31);
              if (k < lim) {
                if (cached) {
                  const double* B = Bs + (size_t)(bi + k) * NB + la * NCL;
#pragma unroll
                  for (int j = 0; j < NCL; ++j) sum += B[j] * rj[j];
                } else {
                  const double* B = A.Sval + (size_t)(b0 + k) * NB + la * NCL;
#pragma unroll
                  for (int j = 0; j < NCL; ++j) sum += __ldg(B + j) * rj[j];
                }
              }
            }
          }
        }
        bi += nblk;
        double tot = sum;
#pragma unroll
        for (int sft = 1; sft < SLOTS; ++sft) tot += __shfl_down_sync(0xffffffffu, sum, sft * NCL);
        if (lane < NCL) {
          if (nb > 0) {
            const int k = A.ann_idx[row];
            if (k >= 0) {
              const double* strip = A.C + ((size_t)k * NCL + lane) * nb;
              const double* q = so + 3 * (size_t)boff;
              for (int j = 0; j < nb; ++j) tot += strip[j] * (__ldcg(q + j) - alpha * (__ldcg(q + nb + j) + beta * __ldcg(q + 2 * nb + j)));
            }
          }
          if (KD > 0) tot += alpha * zm;  // the neighbours' deflation terms, through Z = S~ (S~ W)
          sn[(row * 3 + 1) * NCL + lane] = tot;
          if (MULTI) {
            for (int k = 0; k < W; ++k)
              if ((pmask >> k) & 1u)
                ll_store(reinterpret_cast<ulonglong2*>(A.arena[k] + off_lln) + (size_t)row * (3 * NCL) + NCL + lane, tot, tag_wr);
          }
          acc[0] += rn * rn; acc[1] += tot * rn;
          if (KD > 0) {
            if (ri < dcap) {
              const double* q = Ds + (size_t)ri * 3 * KD * NCL + KD * NCL + lane;
#pragma unroll
              for (int dd = 0; dd < KD; ++dd) acc[2 + dd] += q[dd * NCL] * rn;
            } else {
              const double2* AWi = reinterpret_cast<const double2*>(A.dAW + (size_t)(row * NCL + lane) * KD);
#pragma unroll
              for (int dd = 0; dd < KD / 2; ++dd) {
                const double2 b = __ldg(AWi + dd);
                acc[2 + 2 * dd] += b.x * rn;
                acc[2 + 2 * dd + 1] += b.y * rn;
              }
            }
          }
        }
      } else if (KD == 0) {  // dense border row (never deflated, never sharded): the lanes split the coupled views
        const double* q = so + 3 * (size_t)boff;
        double part[kMaxBorder];
#pragma unroll
        for (int j = 0; j < kMaxBorder; ++j) part[j] = 0.0;
        for (int k = lane; k < A.nav; k += 32) {
          const double* qc = so + (size_t)A.ann_view[k] * 3 * NCL;
#pragma unroll
          for (int a = 0; a < NCL; ++a) {
            const double rk = __ldcg(qc + a) - alpha * (__ldcg(qc + NCL + a) + beta * __ldcg(qc + 2 * NCL + a));
            const double* crow = A.C + ((size_t)k * NCL + a) * nb;
#pragma unroll
            for (int j = 0; j < kMaxBorder; ++j)
              if (j < nb) part[j] += crow[j] * rk;
          }
        }
        double mine_part = 0.0;
#pragma unroll
        for (int j = 0; j < kMaxBorder; ++j) {
          double t = part[j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == j) mine_part = t;
        }
        if (lane < nb) {
          const int i = boff + lane;
          const double ro = __ldcg(q + lane), wo = __ldcg(q + nb + lane), s_o = __ldcg(q + 2 * nb + lane);
          const double pn = ro + beta * A.p[i];
          const double s_n = wo + beta * s_o;
          A.p[i] = pn;
          xl[i] += alpha * pn;
          const double rn = ro - alpha * s_n;
          double* qn = sn + 3 * (size_t)boff;
          qn[lane] = rn; qn[2 * nb + lane] = s_n;
          const double tot = rn + mine_part;  // unit diagonal block
          qn[nb + lane] = tot;
          acc[0] += rn * rn; acc[1] += tot * rn;
        }
      }
    }
    PTZ_CG_PROF(0)  // rows: owner update + sparse product
    if (A.debug & 2) { s_out[0] = 1.0 / (it + 1.0); s_out[1] = 1.0; __syncthreads(); }
    else if (MULTI) {
      if (!ll_reduceN<NV, MAXT / 32>(A, acc, it, tag_wr, my_rank, cta, G, W, false, arrived, sred, s_out, s_part, s_flag)) peers_ok = false;
      if (__syncthreads_or(peers_ok ? 0 : 1)) { status = 3; break; }
    } else grid_reduceN<NV, MAXT / 32>(A, acc, it, arrived, sred, s_out);
    PTZ_CG_PROF(1)  // reduction + barrier
    const double g = s_out[0], d = s_out[1];
    // mu = E^-1 nu (lane c < KD computes mu_c, then every lane collects the vector) and mu^T nu, identically in every warp
    double mu_nu = 0.0;
    if (KD > 0) {
      double mu_l = 0.0, prod = 0.0;
      if (defl && lane < KD) {
#pragma unroll
        for (int dd = 0; dd < KD; ++dd) mu_l += s_einv[lane * KD + dd] * s_out[2 + dd];
        prod = mu_l * s_out[2 + lane];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) prod += __shfl_xor_sync(0xffffffffu, prod, o);
      mu_nu = prod;
#pragma unroll
      for (int dd = 0; dd < KD; ++dd) mus[dd] = __shfl_sync(0xffffffffu, mu_l, dd);
    }
    // (s_out is next written behind the barriers of the next reduction: no extra barrier needed here)
    PTZ_CG_PROF(2)  // mu
    gamma_last = g;
    if (it == 0) {
      // deflated: x0 = W E^-1 W^T b~ already took part of b~ away; measure against |b~| (warm start of the plain kernel likewise)
      gamma0 = (KD > 0) ? A.dscal[0] : (A.gamma0_ptr != nullptr ? A.gamma0_ptr[0] : g);
      if (!(g > 0)) { status = 0; break; }          // zero right-hand side: x = 0
      const double den = d - mu_nu;
      if (!(den > 0) || !isfinite(den)) { status = 2; break; }
      beta = 0.0; alpha = g / den;
    } else {
      if (sqrt(g) <= A.tol * sqrt(gamma0)) { status = 0; break; }
      if (!isfinite(g) || !isfinite(d)) { status = 2; break; }
      if (it >= A.max_iter) { status = (KD > 0 && g < 1e-18 * gamma0) ? 4 : 1; break; }  // (a deflated solve at its cap but nearly done: polish)
      if (KD > 0) {
        // The recurrence residual of the deflated single-reduction iteration can level off a digit or two above the tolerance on
        // ill-conditioned systems (V >= 4000: 1e-12 relative for a tolerance of 1e-13).  Below 1e-10 relative, forty iterations
        // without a new minimum (by 10 %) are that plateau: leave, the host polishes with the plain iteration from this x.  (Twenty
        // iterations below 1e-9 fired on solves that were still converging, V = 8000, and the polish is not cheap: the restarted
        // plain iteration has lost the Krylov space.)
        if (g < 0.9 * g_best) { g_best = g; it_best = it; }
        else if (it - it_best >= 40 && g < 1e-20 * gamma0) { status = 4; break; }
      }
      beta = g / gamma_old;
      const double den = d - beta * g / alpha - mu_nu;
      if (!(den > 0)) { status = 2; break; }
      alpha = g / den;
    }
    if (A.abg != nullptr && cta == 0 && threadIdx.x == 0 && it < A.hist_cap) { A.abg[3 * it] = alpha; A.abg[3 * it + 1] = beta; A.abg[3 * it + 2] = g; }
    gamma_old = g;
    const size_t t = off_o; off_o = off_n; off_n = t;
  }
#undef PTZ_CG_PROF
  if (MULTI && status != 3) {
    // every rank needs the whole solution: push the owned rows of x (plain stores) into the other replicas, then meet once
    // more -- this time behind a system-scope release -- so that nobody leaves before all pushes have landed
    for (int sl = wid; sl < per; sl += wpb) {
      const int slot = cta0 + sl;
      if (slot >= min(V, slot1)) break;
      const int row = A.order[slot];
      if (lane < NCL) {
        const double v = xl[row * NCL + lane];
        for (int k = 0; k < W; ++k)
          if (k != my_rank) reinterpret_cast<double*>(A.arena[k] + A.off_x)[row * NCL + lane] = v;
      }
    }
    double z[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) z[c] = 0.0;
    if (!ll_reduceN<NV, MAXT / 32>(A, z, it + 1, tag0 + (unsigned int)it + 2u, my_rank, cta, G, W, true, arrived, sred, s_out, s_part, s_flag)) status = 3;
  }
  if (cta == 0 && threadIdx.x == 0) {
    ctrl[1] = arrived;
    ctrl[2] = (unsigned long long)(tag0 + (unsigned int)it + 3u);
    if (my_rank == (MULTI ? A.rank : 0)) {
      A.out_info[0] = it; A.out_info[1] = status;
      A.out_res[0] = gamma0 > 0 ? sqrt(gamma_last / gamma0) : 0.0;
    }
  }
}

// ---- deflation: the small kernels around k_cg<.., KD = kDeflK> (once per linear solve; the basis itself once per run) ---------
// W~ = L^T W_y per view block: the basis is kept in the unscaled unknowns y and follows every solve's own block-Jacobi scaling
// (y~ = L^T y).  trans_inv: W_y = Linv^T W~ instead (harvest time).  One thread per (view, column).
template <int NCL>
__global__ void k_defl_scale_basis(int V, const double* __restrict__ T /* L, or Linv */, const double* __restrict__ in, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = idx / kDeflK, c = idx % kDeflK;
  if (v >= V) return;
  const double* L = T + (size_t)v * NCL * NCL;
  double x[NCL];
#pragma unroll
  for (int a = 0; a < NCL; ++a) x[a] = in[((size_t)v * NCL + a) * kDeflK + c];
#pragma unroll
  for (int a = 0; a < NCL; ++a) {
    double s = 0;
#pragma unroll
    for (int q = a; q < NCL; ++q) s += L[q * NCL + a] * x[q];  // (T^T x)_a, T lower triangular
    out[((size_t)v * NCL + a) * kDeflK + c] = s;
  }
}
// Out = S~ In for the kDeflK columns at once.  One CTA (4 warps) per row: lane = column + 16 * half; the 8 half-warps take
// the blocks of the row round-robin, four blocks in flight each (the walk is latency-bound: index -> gather), and the partial
// sums meet in shared memory in a fixed order.  Rows of other ranks (owner != my_rank >= 0) are written as zeros: the caller
// sums the ranks' pieces.
template <int NCL>
__global__ void __launch_bounds__(128) k_defl_spmm(int V, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ Sval,
                                                   const double* __restrict__ in, double* __restrict__ out, const unsigned char* __restrict__ row_owner,
                                                   int my_rank) {
  static_assert(kDeflK == 16, "lane layout");
  __shared__ double sacc[8][NCL][kDeflK];
  const int r = blockIdx.x, t = threadIdx.x, c = t & (kDeflK - 1), hw = t >> 4;  // hw: half-warp 0..7
  double acc[NCL];
#pragma unroll
  for (int a = 0; a < NCL; ++a) acc[a] = 0.0;
  if (my_rank < 0 || row_owner[r] == my_rank) {
    const int k1 = rowptr[r + 1];
    for (int k = rowptr[r] + hw; k < k1; k += 32) {
      int j[4];
      double x[4][NCL];
#pragma unroll
      for (int u = 0; u < 4; ++u) j[u] = (k + 8 * u < k1) ? col[k + 8 * u] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int b = 0; b < NCL; ++b) x[u][b] = j[u] >= 0 ? in[((size_t)j[u] * NCL + b) * kDeflK + c] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j[u] < 0) continue;
        const double* B = Sval + (size_t)(k + 8 * u) * NCL * NCL;
#pragma unroll
        for (int a = 0; a < NCL; ++a)
#pragma unroll
          for (int b = 0; b < NCL; ++b) acc[a] += B[a * NCL + b] * x[u][b];
      }
    }
  }
#pragma unroll
  for (int a = 0; a < NCL; ++a) sacc[hw][a][c] = acc[a];
  __syncthreads();
  for (int e = t; e < NCL * kDeflK; e += 128) {
    const int a = e / kDeflK, cc = e % kDeflK;
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += sacc[q][a][cc];
    out[((size_t)r * NCL + a) * kDeflK + cc] = s;
  }
}
// E = W^T (AW), nu0 = W^T b~, |b~|^2 as per-CTA partial sums over chunks of 256 unknowns (b~ = the r component of the initial CG
// state): partial[chunk][0..K*K) = E, [K*K..K*K+K) = nu0, [K*K+K] = |b~|^2.  Also keeps a copy of b~ for an undeflated retry.
constexpr int kDeflGramVals = kDeflK * kDeflK + kDeflK + 1;
constexpr int kDeflGramChunk = 128;  // unknowns per CTA (two 16 KB panels in shared memory)
template <int NCL>
__global__ void __launch_bounds__(256) k_defl_gram(int V, const double* __restrict__ Wm, const double* __restrict__ AW, const double* __restrict__ st0,
                                                   double* __restrict__ partial, double* __restrict__ b_copy) {
  __shared__ double sw[kDeflGramChunk * kDeflK], sa[kDeflGramChunk * kDeflK], sb[kDeflGramChunk];
  const int n = V * NCL, i0 = blockIdx.x * kDeflGramChunk, cnt = min(kDeflGramChunk, n - i0), t = threadIdx.x;
  for (int e = t; e < kDeflGramChunk * kDeflK; e += 256) {
    const bool in = e < cnt * kDeflK;
    sw[e] = in ? Wm[(size_t)i0 * kDeflK + e] : 0.0;
    sa[e] = in ? AW[(size_t)i0 * kDeflK + e] : 0.0;
  }
  if (t < kDeflGramChunk) {
    const int i = i0 + t;
    const double bt = t < cnt ? st0[((size_t)(i / NCL) * 3) * NCL + (i % NCL)] : 0.0;
    sb[t] = bt;
    if (t < cnt) b_copy[i] = bt;
  }
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * kDeflGramVals;
  {
    const int c = t / kDeflK, d = t % kDeflK;
    double s = 0;
    for (int i = 0; i < kDeflGramChunk; ++i) s += sw[i * kDeflK + c] * sa[i * kDeflK + d];
    out[t] = s;
  }
  if (t < kDeflK) {
    double s = 0;
    for (int i = 0; i < kDeflGramChunk; ++i) s += sw[i * kDeflK + t] * sb[i];
    out[kDeflK * kDeflK + t] = s;
  } else if (t == 32) {
    double s = 0;
    for (int i = 0; i < kDeflGramChunk; ++i) s += sb[i] * sb[i];
    out[kDeflK * kDeflK + kDeflK] = s;
  }
}
// one warp: sums the chunk partials in chunk order, Cholesky of E (symmetrised), Einv, c0 = Einv nu0; dscal[0] = |b~|^2,
// dscal[1] = 1 when every pivot is safely positive, else 0 (the solve then runs undeflated).  kd <= kDeflK columns are in use;
// the rest of Einv is zero.
constexpr int kDeflSmallThreads = 288;  // >= kDeflGramVals: one thread per reduced value, then warp 0 factors E
__global__ void __launch_bounds__(kDeflSmallThreads) k_defl_small(int kd, int nchunk, const double* __restrict__ partial, double* __restrict__ Einv,
                                                                  double* __restrict__ c0, double* __restrict__ dscal) {
  __shared__ double A[kDeflK * kDeflK], Li[kDeflK * kDeflK], nu[kDeflK];
  static_assert(kDeflSmallThreads >= kDeflGramVals, "one thread per value");
  {
    const int e = threadIdx.x;
    if (e < kDeflGramVals) {
      double s = 0;
      int c = 0;
      for (; c + 8 <= nchunk; c += 8) {  // eight loads in flight; summed in chunk order
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = partial[(size_t)(c + u) * kDeflGramVals + e];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
      for (; c < nchunk; ++c) s += partial[(size_t)c * kDeflGramVals + e];
      if (e < kDeflK * kDeflK) A[e] = s;
      else if (e < kDeflK * kDeflK + kDeflK) nu[e - kDeflK * kDeflK] = s;
      else dscal[0] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  for (int e = lane; e < kDeflK * kDeflK; e += 32) {  // symmetrise (lower triangle is what the factorisation reads)
    const int i = e / kDeflK, j = e % kDeflK;
    if (i > j) A[e] = 0.5 * (A[i * kDeflK + j] + A[j * kDeflK + i]);
    Li[e] = 0.0;
  }
  __syncwarp();
  double dmax = 0;
  for (int j = 0; j < kd; ++j) dmax = fmax(dmax, A[j * kDeflK + j]);
  int ok = kd > 0 ? 1 : 0;
  for (int j = 0; j < kd && ok; ++j) {  // column j of L, row `lane`
    double s = 0;
    if (lane >= j && lane < kd) {
      s = A[lane * kDeflK + j];
      for (int k = 0; k < j; ++k) s -= A[lane * kDeflK + k] * A[j * kDeflK + k];
    }
    double dj = __shfl_sync(0xffffffffu, s, j);
    if (!(dj > 1e-12 * dmax) || !isfinite(dj)) { ok = 0; break; }
    dj = sqrt(dj);
    __syncwarp();
    if (lane == j) A[j * kDeflK + j] = dj;
    else if (lane > j && lane < kd) A[lane * kDeflK + j] = s / dj;
    __syncwarp();
  }
  // column c of L^-1 by forward substitution (lane c touches only its own column)
  if (ok && lane < kd) {
    const int c = lane;
    for (int i = c; i < kd; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= A[i * kDeflK + k] * Li[k * kDeflK + c];
      Li[i * kDeflK + c] = s / A[i * kDeflK + i];
    }
  }
  __syncwarp();
  // Einv = L^-T L^-1 (into A, and out)
  double ev[(kDeflK * kDeflK) / 32];
#pragma unroll
  for (int q = 0; q < (kDeflK * kDeflK) / 32; ++q) {
    const int e = lane + 32 * q, i = e / kDeflK, j = e % kDeflK;
    double s = 0;
    if (ok) for (int k = max(i, j); k < kd; ++k) s += Li[k * kDeflK + i] * Li[k * kDeflK + j];
    ev[q] = s;
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < (kDeflK * kDeflK) / 32; ++q) { A[lane + 32 * q] = ev[q]; Einv[lane + 32 * q] = ev[q]; }
  __syncwarp();
  if (lane < kDeflK) {
    double s = 0;
    for (int j = 0; j < kDeflK; ++j) s += A[lane * kDeflK + j] * nu[j];
    c0[lane] = ok ? s : 0.0;
  }
  if (lane == 0) dscal[1] = ok ? 1.0 : 0.0;
}
// x0 = W c0, r0 = b~ - (AW) c0 into the initial CG state (every rank, all rows); nothing to do when the basis was rejected
template <int NCL>
__global__ void k_defl_start(int V, const double* __restrict__ Wm, const double* __restrict__ AW, const double* __restrict__ c0, const double* __restrict__ dscal,
                             double* __restrict__ st0, double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * NCL || dscal[1] == 0.0) return;
  double xs = 0, rs = 0;
#pragma unroll
  for (int c = 0; c < kDeflK; ++c) { xs += Wm[(size_t)i * kDeflK + c] * c0[c]; rs += AW[(size_t)i * kDeflK + c] * c0[c]; }
  x[i] = xs;
  st0[((size_t)(i / NCL) * 3) * NCL + (i % NCL)] -= rs;
}
// back to the plain initial state (r = b~, w = s = 0, x = p = 0) for an undeflated retry after a breakdown
template <int NCL>
__global__ void k_cg_restart(int V, const double* __restrict__ b_copy, double* __restrict__ st0, double* __restrict__ x, double* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * NCL) return;
  const size_t o = ((size_t)(i / NCL) * 3) * NCL + (i % NCL);
  st0[o] = b_copy[i]; st0[o + NCL] = 0.0; st0[o + 2 * NCL] = 0.0;
  x[i] = 0.0; p[i] = 0.0;
}
// warm start of the plain iteration from a given x~ (a deflated solve that stagnated just above the tolerance): the TRUE residual
// r = b~ - S~ x~ into the CG state (w = s = 0, p = 0); one warp per row, lanes over the row's blocks, fixed order.  Rows of other
// ranks are written as zeros (the caller sums the ranks' pieces).
template <int NCL>
__global__ void __launch_bounds__(256) k_cg_residual(int V, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ Sval,
                                                     const double* __restrict__ x, const double* __restrict__ b_copy, double* __restrict__ st0,
                                                     double* __restrict__ p, const unsigned char* __restrict__ row_owner, int my_rank) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= V) return;
  double acc[NCL];
#pragma unroll
  for (int a = 0; a < NCL; ++a) acc[a] = 0.0;
  const bool mine = my_rank < 0 || row_owner[r] == my_rank;
  if (mine) {
    for (int k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) {
      const int c = col[k] & 0x7fffffff;
      const double* B = Sval + (size_t)k * NCL * NCL;
      double xc[NCL];
#pragma unroll
      for (int b = 0; b < NCL; ++b) xc[b] = x[c * NCL + b];
#pragma unroll
      for (int a = 0; a < NCL; ++a)
#pragma unroll
        for (int b = 0; b < NCL; ++b) acc[a] += B[a * NCL + b] * xc[b];
    }
  }
#pragma unroll
  for (int a = 0; a < NCL; ++a) acc[a] = warp_sum(acc[a]);
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < NCL; ++a) {
      st0[(r * 3 + 0) * NCL + a] = mine ? b_copy[r * NCL + a] - acc[a] : 0.0;
      st0[(r * 3 + 1) * NCL + a] = 0.0;
      st0[(r * 3 + 2) * NCL + a] = 0.0;
      p[r * NCL + a] = 0.0;
    }
  }
}
// harvest: W~[i][c] = sum_j Y[j][c] hist[j][i]  (Y already carries 1 / |r_j|); rows of other ranks -> 0 (summed by the caller)
__global__ void k_defl_harvest(int n, int m, int kd, const double* __restrict__ hist, const double* __restrict__ Y /* [m][kDeflK] */,
                               double* __restrict__ Wt, const unsigned char* __restrict__ row_owner, int ncl, int my_rank) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = idx / kDeflK, c = idx % kDeflK;
  if (i >= n) return;
  double s = 0;
  if (c < kd && (my_rank < 0 || row_owner[i / ncl] == my_rank))
    for (int j = 0; j < m; ++j) s += Y[(size_t)j * kDeflK + c] * hist[(size_t)j * n + i];
  Wt[(size_t)i * kDeflK + c] = s;
}

// -------------------------------------------------------------------------------------------------------------
// stage 4
// -------------------------------------------------------------------------------------------------------------
__global__ void k_gather_int(int n, const int* __restrict__ idx, const int* __restrict__ src, int* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
// per track: y_p = L^-T (t - sum_o What_o^T y_c(o)); candidate ray = ray - s*y.  Partials (per CTA): model cost change,
// |step|^2, |x_cand|^2.
template <int NCL>
__global__ void k_track_backsub(int P, const int* __restrict__ t_off, const int* __restrict__ t_obs, const int* __restrict__ t_view /* view of t_obs[i]: saves the dependent o_view lookup */,
                                const double* __restrict__ What, const double* __restrict__ y, const double* __restrict__ Lt, const double* __restrict__ Vh,
                                const double* __restrict__ diag_ray, double mu, const double* __restrict__ trk, double* __restrict__ trk_cand,
                                double* __restrict__ part3, const double* __restrict__ Wdh /* or nullptr */, const double* __restrict__ ydisp) {
  typedef Dims<NCL> D;
  __shared__ double sred[3 * 8];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[3] = {0, 0, 0};
  if (p < P) {
    const double* t = trk + (size_t)p * kTrk;
    double* tc = trk_cand + (size_t)p * kTrk;
    if (t_off[p] == t_off[p + 1]) {
      tc[0] = t[0]; tc[1] = t[1]; tc[2] = t[2];
    } else {
      const double* lt = Lt + (size_t)p * 10;
      double b0 = lt[6], b1 = lt[7], b2 = lt[8];
      const int tb = t_off[p], te = t_off[p + 1];
      for (int i = tb; i < te; i += 4) {
        int ob[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) ob[u] = t_obs[min(i + u, te - 1)];
        int vw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) vw[u] = t_view[min(i + u, te - 1)];
        double wv[4][D::WS], yv[4][NCL];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double* w4 = What + (size_t)ob[u] * D::WS;
#pragma unroll
          for (int k = 0; k < D::WS / 4; ++k) ld256(w4 + 4 * k, wv[u][4 * k], wv[u][4 * k + 1], wv[u][4 * k + 2], wv[u][4 * k + 3]);
#pragma unroll
          for (int a = 0; a < NCL; ++a) yv[u][a] = y[(size_t)vw[u] * NCL + a];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i + u >= te) break;
#pragma unroll
          for (int a = 0; a < NCL; ++a) { const double ya = yv[u][a]; b0 -= wv[u][3 * a] * ya; b1 -= wv[u][3 * a + 1] * ya; b2 -= wv[u][3 * a + 2] * ya; }
        }
      }
      if (Wdh) {  // PTZRayDistDisp: - Wdh^T y_disp
        const double* wd = Wdh + (size_t)p * 12;
#pragma unroll
        for (int a = 0; a < 3; ++a) { const double ya = ydisp[a]; b0 -= wd[3 * a] * ya; b1 -= wd[3 * a + 1] * ya; b2 -= wd[3 * a + 2] * ya; }
      }
      // L^T y = b
      const double y2 = b2 / lt[5], y1 = (b1 - lt[4] * y2) / lt[2], y0 = (b0 - lt[1] * y1 - lt[3] * y2) / lt[0];
      const double* vh = Vh + (size_t)p * 10;
      const double* dg = diag_ray + 3 * (size_t)p;
      acc[0] = 0.5 * (y0 * (vh[6] + dg[0] / mu * y0) + y1 * (vh[7] + dg[1] / mu * y1) + y2 * (vh[8] + dg[2] / mu * y2));
      const double d0 = -t[4] * y0, d1 = -t[5] * y1, d2 = -t[6] * y2;
      const double c0 = t[0] + d0, c1 = t[1] + d1, c2 = t[2] + d2;
      tc[0] = c0; tc[1] = c1; tc[2] = c2;
      acc[1] = (t[0] - c0) * (t[0] - c0) + (t[1] - c1) * (t[1] - c1) + (t[2] - c2) * (t[2] - c2);
      acc[2] = c0 * c0 + c1 * c1 + c2 * c2;
    }
  }
  block_sum<3>(acc, sred);
  if (threadIdx.x == 0) { part3[3 * blockIdx.x] = acc[0]; part3[3 * blockIdx.x + 1] = acc[1]; part3[3 * blockIdx.x + 2] = acc[2]; }
}

// ambient index of live camera column a: into intr (0..8) or ext (9 + 0..5)
template <int TYPE>
__host__ __device__ constexpr int live_col_slot(int a) {
  // PTZRay: fx,w | Dist, DistDisp: fx,k1,w | FxfyDist: fx,fy,k1,w
  return TYPE == BA_PTZRAY ? (a == 0 ? 0 : 9 + (a - 1))
       : TYPE == BA_PTZRAY_FXFY_DIST ? (a == 0 ? 0 : a == 1 ? 1 : a == 2 ? 4 : 9 + (a - 3))
       : (a == 0 ? 0 : a == 1 ? 4 : 9 + (a - 2));
}

// per view: candidate camera = camera - s*y on the live columns; partials as above (one value triple per CTA)
template <int TYPE>
__global__ void k_cam_update(int V, const double* __restrict__ y, const double* __restrict__ scale_cam, const double* __restrict__ g,
                             const double* __restrict__ diag_cam, double mu, const int* __restrict__ view_active, const double* __restrict__ intr,
                             const double* __restrict__ ext, double* __restrict__ intr_c, double* __restrict__ ext_c, double* __restrict__ part3,
                             const int* __restrict__ intr_counted /* shared intrinsics: 0 = the block belongs to another view; nullptr = all own */) {
  constexpr int NCL = ba_ncl(TYPE);
  __shared__ double sred[3 * 8];
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[3] = {0, 0, 0};
  if (v < V) {
    double x[15], c[15];
    for (int j = 0; j < 9; ++j) x[j] = intr[9 * v + j];
    for (int j = 0; j < 6; ++j) x[9 + j] = ext[6 * v + j];
    for (int j = 0; j < 15; ++j) c[j] = x[j];
#pragma unroll
    for (int a = 0; a < NCL; ++a) {
      const double ya = y[v * NCL + a];
      acc[0] += 0.5 * ya * (g[v * NCL + a] + diag_cam[v * NCL + a] / mu * ya);
      c[live_col_slot<TYPE>(a)] = x[live_col_slot<TYPE>(a)] + (-scale_cam[v * NCL + a] * ya);
    }
    for (int j = 0; j < 9; ++j) intr_c[9 * v + j] = c[j];
    for (int j = 0; j < 6; ++j) ext_c[6 * v + j] = c[9 + j];
    if (view_active[v]) {
      const int j0 = (intr_counted != nullptr && !intr_counted[v]) ? 9 : 0;  // a shared intrinsics block counts once, at its first view
      for (int j = j0; j < 15; ++j) { acc[1] += (x[j] - c[j]) * (x[j] - c[j]); acc[2] += c[j] * c[j]; }
    }
  }
  block_sum<3>(acc, sred);
  if (threadIdx.x == 0) { part3[3 * blockIdx.x] = acc[0]; part3[3 * blockIdx.x + 1] = acc[1]; part3[3 * blockIdx.x + 2] = acc[2]; }
}

// cost-only pass at the candidate point; `raw` != 0 also accumulates the unweighted squared residual (CalReprojError2d2d)
template <int TYPE>
__global__ void __launch_bounds__(kChunk) k_cost(const int* __restrict__ chunk_view, const int* __restrict__ chunk_begin, const int* __restrict__ chunk_cnt,
                                                 const float2* __restrict__ o_uv, const int* __restrict__ o_track, const ViewTab* __restrict__ vt,
                                                 const double* __restrict__ trk, const double* __restrict__ disp, double* __restrict__ part2) {
  __shared__ ViewTab svt;
  __shared__ double sred[2 * (kChunk / 32)];
  const int chunk = blockIdx.x;
  const int view = chunk_view[chunk], begin = chunk_begin[chunk], cnt = chunk_cnt[chunk];
  if (threadIdx.x < kViewTabDoubles) reinterpret_cast<double*>(&svt)[threadIdx.x] = reinterpret_cast<const double*>(vt + view)[threadIdx.x];
  // the observation and its track record are fetched BEFORE the barrier: the gather chain (index -> record) runs beside the view
  // table's load instead of behind it (the kernel is a chain of dependent latencies: 35 -> 27 us)
  float2 uv = make_float2(0.f, 0.f);
  double ray[3] = {0, 0, 1}, tw = 0;
  if (threadIdx.x < cnt) {
    const int o = begin + threadIdx.x;
    uv = o_uv[o];
    const int p = o_track[o];
    ld256(trk + (size_t)p * kTrk, ray[0], ray[1], ray[2], tw);  // the first 32 bytes of the track record: ray, sqrt(weight)
  }
  __syncthreads();
  double acc[2] = {0, 0};
  if (threadIdx.x < cnt) {
    double dz[3] = {0, 0, 0};
    if (TYPE == BA_PTZRAY_DIST_DISP) { dz[0] = disp[0]; dz[1] = disp[1]; dz[2] = disp[2]; }
    double r[2];
    ba_obs<TYPE, false>(svt, ray, dz, (double)uv.x, (double)uv.y, r, nullptr, nullptr, nullptr);
    const double s = r[0] * r[0] + r[1] * r[1];
    acc[0] = 0.5 * tw * tw * s;
    acc[1] = s;
  }
  block_sum<2>(acc, sred);
  if (threadIdx.x == 0) { part2[2 * (size_t)chunk] = acc[0]; part2[2 * (size_t)chunk + 1] = acc[1]; }
}

// fixed-order single-CTA reductions of the partial arrays into the scalar block the host reads once per iteration
struct ScalarJobs {
  // sums: (ptr, count, stride, offset) -> out slot ; maxes likewise
  const double* sum_ptr[12]; int sum_n[12]; int sum_stride[12]; int sum_slot[12]; int nsum;
  const double* max_ptr[4]; int max_n[4]; int max_slot[4]; int nmax;
};
__global__ void __launch_bounds__(1024) k_scalars(ScalarJobs J, double* __restrict__ out) {
  // one CTA per job (grid = nsum + nmax), fixed-order tree inside the CTA
  __shared__ double sred[32];
  const int job = blockIdx.x;
  if (job < J.nsum) {
    const int j = job;
    double s[1] = {0};
    for (int i = threadIdx.x; i < J.sum_n[j]; i += blockDim.x) s[0] += J.sum_ptr[j][(size_t)i * J.sum_stride[j]];
    block_sum<1>(s, sred);
    if (threadIdx.x == 0) out[J.sum_slot[j]] = s[0];
  } else if (job < J.nsum + J.nmax) {
    const int j = job - J.nsum;
    double m = 0;
    for (int i = threadIdx.x; i < J.max_n[j]; i += blockDim.x) m = fmax(m, J.max_ptr[j][i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      double mm = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mm = fmax(mm, sred[w]);
      out[J.max_slot[j]] = mm;
    }
  }
}

// What the host reads once per step attempt, in ONE pinned, device-mapped block: the kernel writes it over PCIe and raises
// `seq` behind a system-scope fence; the host spins on `seq` (no copy engine, no stream synchronise on the critical path).
struct PublishBlock {
  double sc[32];          // the scalar slots
  double dscal[2];        // deflation: |b~|^2, basis usable
  int info[4];            // pcg iterations, pcg status, factorisation failure flag
  unsigned long long seq;
};
__global__ void __launch_bounds__(64) k_publish(const double* __restrict__ scalars, const int* __restrict__ pcg_info, const int* __restrict__ fail,
                                               const double* __restrict__ dscal /* or nullptr */, PublishBlock* host, unsigned long long seq) {
  const int t = threadIdx.x;
  if (t < 32) host->sc[t] = scalars[t];
  else if (t < 34) host->info[t - 32] = pcg_info[t - 32];
  else if (t == 34) host->info[2] = fail[0];
  else if (t < 37 && dscal) host->dscal[t - 35] = dscal[t - 35];
  __threadfence_system();
  __syncthreads();
  if (t == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&host->seq) = seq;
  }
}

// |x|^2 of the current point over the coordinates that are in the Ceres problem: partial sums per CTA
__global__ void k_xnorm2(int V, int P, const int* __restrict__ view_active, const double* __restrict__ intr, const double* __restrict__ ext,
                         const int* __restrict__ t_off, const double* __restrict__ trk, double* __restrict__ part2,
                         const int* __restrict__ intr_counted) {
  __shared__ double sred[2 * 8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[2] = {0, 0};
  if (i < V && view_active[i]) {
    if (intr_counted == nullptr || intr_counted[i])
      for (int j = 0; j < 9; ++j) acc[0] += intr[9 * i + j] * intr[9 * i + j];
    for (int j = 0; j < 6; ++j) acc[0] += ext[6 * i + j] * ext[6 * i + j];
  }
  if (i < P && t_off[i + 1] > t_off[i]) {
    const double* t = trk + (size_t)i * kTrk;
    acc[1] = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
  }
  block_sum<2>(acc, sred);
  if (threadIdx.x == 0) { part2[2 * blockIdx.x] = acc[0]; part2[2 * blockIdx.x + 1] = acc[1]; }
}
// rays in the world frame (ptzray_optimizer.cc:748-754): R_lw^T (ray - t_lw), compact [P][3]
__global__ void k_rays_out(int P, const double* __restrict__ trk, const double* __restrict__ tlw, double* __restrict__ ray, double* __restrict__ ray_w) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double w[3] = {tlw[0], tlw[1], tlw[2]}, R[9];
  rodrigues_jac(w, R, nullptr);
  const double* t = trk + (size_t)p * kTrk;
  const double r0 = t[0], r1 = t[1], r2 = t[2];
  ray[3 * (size_t)p] = r0; ray[3 * (size_t)p + 1] = r1; ray[3 * (size_t)p + 2] = r2;
  for (int j = 0; j < 3; ++j) ray_w[3 * (size_t)p + j] = R[j] * (r0 - tlw[3]) + R[3 + j] * (r1 - tlw[4]) + R[6 + j] * (r2 - tlw[5]);
}

// initial rays (Pix2Ray, ptzray_optimizer.cc:768-797): mean over the track's views of normalise((R^-1 K^-1)[u,v,1]), normalised
// initial track records from the caller's arrays: ray0 (or zero, then k_init_rays), sqrt(ScaledLoss weight), unit Jacobi scales
__global__ void k_trk_init(int P, const double* __restrict__ weight, const double* __restrict__ ray0, double* __restrict__ trk) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double4 a = make_double4(0, 0, 0, sqrt(weight[p]));
  if (ray0) { a.x = ray0[3 * (size_t)p]; a.y = ray0[3 * (size_t)p + 1]; a.z = ray0[3 * (size_t)p + 2]; }
  reinterpret_cast<double4*>(trk + (size_t)p * kTrk)[0] = a;
  reinterpret_cast<double4*>(trk + (size_t)p * kTrk)[1] = make_double4(1.0, 1.0, 1.0, 0.0);
}
__global__ void k_init_rays(int P, const int* __restrict__ t_off, const int* __restrict__ t_obs, const int* __restrict__ o_view,
                            const float2* __restrict__ o_uv, const double* __restrict__ RiKi, double* __restrict__ trk) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double a0 = 0, a1 = 0, a2 = 0;
  const int n = t_off[p + 1] - t_off[p];
  for (int i = t_off[p]; i < t_off[p + 1]; ++i) {
    const int o = t_obs[i];
    const double* Mv = RiKi + 9 * (size_t)o_view[o];
    const double u = o_uv[o].x, v = o_uv[o].y;
    const double x = Mv[0] * u + Mv[1] * v + Mv[2], yv = Mv[3] * u + Mv[4] * v + Mv[5], z = Mv[6] * u + Mv[7] * v + Mv[8];
    const double nn = sqrt(x * x + yv * yv + z * z);
    a0 += x / nn; a1 += yv / nn; a2 += z / nn;
  }
  a0 /= n; a1 /= n; a2 /= n;
  const double nn = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
  trk[(size_t)p * kTrk] = a0 / nn; trk[(size_t)p * kTrk + 1] = a1 / nn; trk[(size_t)p * kTrk + 2] = a2 / nn;
}
// (R^-1 K^-1) per view with OpenCV's 3x3 cofactor inverse (cv::Mat::inv, ptzray_optimizer.cc:786)
__global__ void k_rikI(int V, const double* __restrict__ intr, const double* __restrict__ ext, double* __restrict__ RiKi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  double w[3] = {ext[6 * i], ext[6 * i + 1], ext[6 * i + 2]}, R[9];
  rodrigues_jac(w, R, nullptr);
  double S[9], Ri[9], Ki[9];
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 0) for (int j = 0; j < 9; ++j) S[j] = R[j];
    else { S[0] = intr[9 * i]; S[1] = 0; S[2] = intr[9 * i + 2]; S[3] = 0; S[4] = intr[9 * i + 1]; S[5] = intr[9 * i + 3]; S[6] = 0; S[7] = 0; S[8] = 1; }
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    d = 1.0 / d;
    double* T = pass == 0 ? Ri : Ki;
    T[0] = (S[4] * S[8] - S[5] * S[7]) * d; T[1] = (S[2] * S[7] - S[1] * S[8]) * d; T[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    T[3] = (S[5] * S[6] - S[3] * S[8]) * d; T[4] = (S[0] * S[8] - S[2] * S[6]) * d; T[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    T[6] = (S[3] * S[7] - S[4] * S[6]) * d; T[7] = (S[1] * S[6] - S[0] * S[7]) * d; T[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) RiKi[9 * (size_t)i + 3 * r + c] = Ri[3 * r] * Ki[c] + Ri[3 * r + 1] * Ki[3 + c] + Ri[3 * r + 2] * Ki[6 + c];
}

}  // namespace ptz
