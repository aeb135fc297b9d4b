// ba_solver.cu — host side of the PTZ-BA hot path: problem upload, the Levenberg–Marquardt loop that sequences the
// stage kernels, and the extern "C" entry points of include/ptzcalib_b200.h.
//
// The loop restates Ceres 1.14's TrustRegionMinimizer + LevenbergMarquardtStrategy (what runs behind
// ptzray_optimizer.cc:475) with the decisions taken on the host from ONE small device->host read per iteration;
// every O(M), O(P) or O(V) piece of work is a kernel.  There is no CPU solver path in this library.
#include <math.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <chrono>
#include <memory>

#include "ba_border.cuh"
#include "ba_kernels.cuh"
#include "ba_setup.cuh"
#include "ba_structure.hpp"
#include "cg_ritz.hpp"

namespace ptz {

// --------------------------------------------------------------------------------------------------------------
// errors, NCCL state
// --------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct NcclState {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};
static NcclState g_nccl;

#define PTZ_NCCL(expr)                                                                                       \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) {                                                                                 \
      char _b[512];                                                                                          \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r));        \
      throw ::ptz::CudaError(PTZ_ERR_NCCL, _b);                                                              \
    }                                                                                                        \
  } while (0)

// ---- peer arena of the row-sharded CG (k_cg): one cudaMalloc'ed block per rank, mapped into every other rank of the box
// through CUDA IPC.  Persistent (grown on demand, collectively): its barrier counter and epoch carry over between solves.
struct PeerArena {
  char* base[kMaxPeers] = {nullptr};
  size_t bytes = 0;
  int world = 0;  // world size it was built for
};
static PeerArena g_arena;

static void arena_release() {
  if (!g_arena.bytes) return;
  cudaDeviceSynchronize();
  for (int k = 0; k < kMaxPeers; ++k) {
    if (!g_arena.base[k]) continue;
    if (g_arena.world <= 1 || k == g_nccl.rank) cudaFree(g_arena.base[k]);  // own block, or local virtual replicas (world < 0)
    else cudaIpcCloseMemHandle(g_arena.base[k]);
    g_arena.base[k] = nullptr;
  }
  g_arena.bytes = 0;
  g_arena.world = 0;
}

// Collective over the ranks of the communicator (every rank asks for the same size: the reduced system is replicated).
// vranks > 1 (debug, single process): that many replicas on this device instead of peer mappings.
static void arena_ensure(size_t bytes, cudaStream_t s, int vranks) {
  const int W = g_nccl.world, R = g_nccl.rank;
  if (vranks > 1) {
    if (g_arena.bytes >= bytes && g_arena.world == -vranks) return;
    arena_release();
    const size_t want = std::max(bytes + bytes / 2, (size_t)8 << 20);
    for (int k = 0; k < vranks; ++k) {
      PTZ_CUDA(cudaMalloc((void**)&g_arena.base[k], want));
      PTZ_CUDA(cudaMemset(g_arena.base[k], 0, want));
      const unsigned long long one = 1;
      PTZ_CUDA(cudaMemcpy(g_arena.base[k] + 16, &one, 8, cudaMemcpyHostToDevice));
    }
    PTZ_CUDA(cudaDeviceSynchronize());
    g_arena.world = -vranks;
    g_arena.bytes = want;
    return;
  }
  if (g_arena.bytes >= bytes && g_arena.world == W) return;
  if (W > kMaxPeers) throw CudaError(PTZ_ERR_UNSUPPORTED, "more than 8 ranks per box");
  if (W > 1) {  // nobody may still be spinning on, or writing into, the old arena
    PTZ_CUDA(cudaStreamSynchronize(s));
    DevBuf<double> d_tok;
    d_tok.alloc(1, nullptr);
    PTZ_CUDA(cudaMemsetAsync(d_tok.p, 0, 8, s));
    PTZ_NCCL(ncclAllReduce(d_tok.p, d_tok.p, 1, ncclDouble, ncclSum, g_nccl.comm, s));
    PTZ_CUDA(cudaStreamSynchronize(s));
  }
  arena_release();
  const size_t want = std::max(bytes + bytes / 2, (size_t)8 << 20);
  char* mine = nullptr;
  PTZ_CUDA(cudaMalloc((void**)&mine, want));
  PTZ_CUDA(cudaMemset(mine, 0, want));
  {
    const unsigned long long one = 1;  // first LL tag: zero-filled memory must never look valid
    PTZ_CUDA(cudaMemcpy(mine + 16, &one, 8, cudaMemcpyHostToDevice));
  }
  PTZ_CUDA(cudaDeviceSynchronize());
  g_arena.base[R] = mine;
  g_arena.world = W;
  g_arena.bytes = want;
  if (W > 1) {
    cudaIpcMemHandle_t h;
    PTZ_CUDA(cudaIpcGetMemHandle(&h, mine));
    DevBuf<char> d_h, d_all;
    d_h.alloc(sizeof(h), nullptr);
    d_all.alloc(sizeof(h) * W, nullptr);
    PTZ_CUDA(cudaMemcpyAsync(d_h.p, &h, sizeof(h), cudaMemcpyHostToDevice, s));
    PTZ_NCCL(ncclAllGather(d_h.p, d_all.p, sizeof(h), ncclChar, g_nccl.comm, s));  // also the barrier after everybody's memset
    std::vector<cudaIpcMemHandle_t> all(W);
    PTZ_CUDA(cudaMemcpyAsync(all.data(), d_all.p, sizeof(h) * W, cudaMemcpyDeviceToHost, s));
    PTZ_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < W; ++k) {
      if (k == R) continue;
      void* ptr = nullptr;
      PTZ_CUDA(cudaIpcOpenMemHandle(&ptr, all[k], cudaIpcMemLazyEnablePeerAccess));
      g_arena.base[k] = (char*)ptr;
    }
  }
}

static void allreduce_sum(double* buf, size_t n, cudaStream_t s) {
  if (g_nccl.world > 1 && n) PTZ_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, g_nccl.comm, s));
}
static void allreduce_max(double* buf, size_t n, cudaStream_t s) {
  if (g_nccl.world > 1 && n) PTZ_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, ncclMax, g_nccl.comm, s));
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// scalar slots read by the host once per LM iteration
enum Slot {
  // summed across ranks (local partial sums)
  S_COST_CAND = 0, S_RAW2_CAND, S_DM_RAY, S_STEP2_RAY, S_XN2_RAY, S_SUM_END = 8,
  // max across ranks
  S_GMAX_RAY = 8, S_MAX_END = 10,
  // replicated (computed from all-reduced data or from replicated camera parameters)
  S_COST_X = 10, S_GMAX_CAM, S_GMAX_B, S_DM_CAM, S_STEP2_CAM, S_XN2_CAM, S_DM_B, S_STEP2_B, S_XN2_B, S_COSTPTS_X, S_RAWPTS_X, S_COSTPTS_CAND, S_RAWPTS_CAND,
  S_COUNT = 32
};

// Process-wide free lists of CUDA events and small pinned host blocks.  The many short-lived solver objects of one IBA run (and
// bench.py's end-to-end leg) would otherwise call cudaEventCreate / cudaMallocHost / cudaFreeHost per solve: driver-lock calls that
// were seen to stall for tens of milliseconds now and then (e.g. while nvidia-smi polls the device).
struct HostCache {
  static std::mutex& lock() { static std::mutex m; return m; }
  static std::vector<cudaEvent_t>& events() { static std::vector<cudaEvent_t> v; return v; }
  static std::vector<void*>& pinned() { static std::vector<void*> v; return v; }
  static constexpr size_t kPinnedBytes = 512;
  static cudaEvent_t get_event() {
    {
      std::lock_guard<std::mutex> g(lock());
      if (!events().empty()) { cudaEvent_t e = events().back(); events().pop_back(); return e; }
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  static void put_events(const std::vector<cudaEvent_t>& ev) {
    std::lock_guard<std::mutex> g(lock());
    events().insert(events().end(), ev.begin(), ev.end());
  }
  static void* get_pinned() {
    {
      std::lock_guard<std::mutex> g(lock());
      if (!pinned().empty()) { void* p = pinned().back(); pinned().pop_back(); return p; }
    }
    void* p = nullptr;
    PTZ_CUDA(cudaMallocHost(&p, kPinnedBytes));
    return p;
  }
  static void put_pinned(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(lock());
    pinned().push_back(p);
  }
};

struct StageClock {
  std::vector<cudaEvent_t> ev;
  std::vector<int> tag;  // kernel id of interval [2i, 2i+1]; -1 = a whole ptzba_run
  size_t used = 0;
  float ms[PTZ_K_COUNT] = {0};
  int launches[PTZ_K_COUNT] = {0};
  float ms_run = 0;
  cudaStream_t stream = nullptr;
  // per-kernel events: two cudaEventRecord per launch, which open small gaps between dependent kernels on the device (measured
  // below).  On by default (ptzba_get_stage_times reports them); ptzba_set_stage_timing(h, 0) keeps only the span of the runs.
  bool per_kernel = true;
  void init(cudaStream_t s) { stream = s; const char* e = getenv("PTZ_STAGE_TIMES"); if (e) per_kernel = atoi(e) != 0; }
  void begin(int id) {
    if (used + 2 > ev.size()) {
      for (int i = 0; i < 256; ++i) ev.push_back(HostCache::get_event());
    }
    cudaEventRecord(ev[used], stream);
    tag.push_back(id);
    ++used;
  }
  void end() { cudaEventRecord(ev[used], stream); ++used; }
  void collect() {  // call after a stream synchronise
    for (size_t i = 0; i + 1 < used; i += 2) {
      float t = 0;
      cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      if (tag[i / 2] < 0) ms_run += t; else ms[tag[i / 2]] += t;
    }
    used = 0;
    tag.clear();
  }
  void reset() { collect(); for (int i = 0; i < PTZ_K_COUNT; ++i) { ms[i] = 0; launches[i] = 0; } ms_run = 0; }
  ~StageClock() { HostCache::put_events(ev); }  // (the owning solver synchronised its stream before this runs)
};
// time one kernel launch (or one collective) under its id
#define PTZ_TIMED(id, ...) do { if (clk.per_kernel) { clk.begin(id); __VA_ARGS__; clk.end(); } else { __VA_ARGS__; } ++clk.launches[id]; } while (0)

// --------------------------------------------------------------------------------------------------------------
// the solver, specialised on the factor type
// --------------------------------------------------------------------------------------------------------------
struct BaSolverBase {
  virtual ~BaSolverBase() {}
  virtual void reset() = 0;
  virtual void run(int max_new_iterations, ptzba_result* out) = 0;
  virtual void eval(const double* disp, ptzba_eval_out* out) = 0;
  virtual void stage_times(ptzba_stage_times* t) = 0;
  virtual void set_stage_timing(bool on) = 0;
};

template <int TYPE>
struct BaSolver : BaSolverBase {
  static constexpr int NCL = ba_ncl(TYPE);
  typedef Dims<NCL> D;
  static constexpr bool kFyBorder = (TYPE != BA_PTZRAY_FXFY_DIST);
  static constexpr bool kDisp = (TYPE == BA_PTZRAY_DIST_DISP);

  StreamHolder sh;       // first member: destroyed last, after every buffer allocated on its stream
  int V, P, M, A, nb = 0, nav = 0, n = 0;
  int nf = 0, nbt = 0;                 // fy unknowns of annotated views (eliminated before the CG, ba_border.cuh); nbt = nb + nf
  int bo_tlw = -1, bo_disp = -1;       // first border column of tlw(6) / disp(3)
  bool rows_sharded = false;           // the CG's rows are split across the ranks (camera-only systems); else every rank solves it all
  // SetSharedIntrinsics: groups of >= 2 views with one intrinsics block (ba_border.cuh).  ns = G * (NCL - 3) unknowns in the border
  int ngrp = 0, ns = 0, bo_sh = 0, nb_plain = 0;
  std::vector<int> h_rep, h_grp_of, h_grp_off, h_grp_view, h_intr_counted;
  DevBuf<int> d_grp_of, d_grp_off, d_grp_view, d_intr_counted;
  DevBuf<double> d_sh_h, d_rowpart;
  int ncpl = 0;                               // views with a coupling strip to the border (annotated ones, or all of them with disp)
  std::vector<int> h_cpl_view, h_cpl_idx, h_ann_strip;
  ptz_solver_options opt;
  DevStructure ds;       // orderings + block pattern, device-resident
  std::vector<int> h_perm;  // view-major position -> caller's observation index (fetched lazily, ptzba_eval only)
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  StageClock clk;
  double seconds_setup = 0;

  // host copies of the inputs that outputs need
  std::vector<double> h_intr0, h_ext0, h_tlw0;
  bool have_ray0 = false;
  std::vector<int> h_view_active, h_ann_view, h_ann_off, h_ann_idx, h_pt_perm;

  // device: structure lives in `ds`; short aliases are set in upload()
  // device: parameters (two copies: current / candidate)
  DevBuf<double> d_intr[2], d_ext[2], d_trk[2], d_tlw[2], d_intr_init, d_ext_init, d_trk_init, d_tlw_init;
  int cur = 0;
  int od_group = 32;       // lanes per block in the off-diagonal Schur kernel: 32 (k_schur_offdiag), 16 or 8 (k_schur_offdiag_sub, NCL = 4)
  int nub_local = 0;       // upper blocks of S with observation pairs on THIS rank (sharded problem: ~1/W of the union pattern)
  DevBuf<int> d_ub_list;   // their indices; empty = all blocks
  bool use_dense = false;  // n <= kDenseMaxN: stage 3 is a dense Cholesky in one CTA (k_dense_chol) instead of the CG
  DevBuf<double> d_dense;  // [(n + 1) x n]
  size_t dense_smem_bytes() const { return (size_t)(kDenseNB * kDenseLd + 32 * kDenseNB + 2 * ((n + 3) & ~3) + (n + 1) * (n | 1)) * sizeof(double); }
  double fuse_factor_mu = 0.0;  // > 0: the next k_track_accum also factors the ray blocks for this radius
  double lt_mu = 0.0;           // radius d_Lt was factored for inside k_track_accum (0: k_track_factor has to run)
  int vt_for = -1;         // parameter copy (0/1) the view table d_vt was built from, -1 = stale
  bool vt_scaled = false;  // ... with the final Jacobi scales in it
  // device: work
  DevBuf<ViewTab> d_vt;
  DevBuf<double> d_scale_cam, d_scale_b, d_recA, d_recF, d_part, d_viewred, d_Vh, d_gmax_part, d_diag_ray, d_diag_cam, d_diag_b, d_Lt, d_What, d_q, d_sys, d_Linv,
      d_Linv_b, d_Cs, d_cgp, d_y, d_pcg_res, d_part3_ray, d_part3_cam, d_part3_b, d_cost_part, d_scalars, d_RiKi, d_pts_scratch, d_pts_raw,
      d_pts_xyz;
  DevBuf<float2> d_pts_uv;
  DevBuf<int> d_pts_view, d_ann_view, d_ann_off, d_ann_idx, d_fail, d_pcg_info, d_cpl_view, d_cpl_idx, d_ann_strip;
  DevBuf<double> d_recd, d_dpart, d_Wdh, d_Cw, d_hinv, d_dispp[2], d_disp_init;
  DevBuf<int> d_cg_order, d_t_view, d_t_trk, d_grp_e0;
  int track_groups = 0, nray_parts = 1;  // groups of 32 by-track entries (0: thread-per-track kernels); per-warp / per-CTA partial counts
  // views into d_viewred (all-reduced once per Jacobian evaluation): U | g | cost_view | C | Hbb | gb | cost_pts(2)
  double *p_U, *p_g, *p_cost_view, *p_C, *p_Hbb, *p_gb, *p_cost_pts, *p_gabs, *p_gabs_b, *p_Cf, *p_Hrf, *p_Hff;
  DevBuf<double> d_gabs;
  size_t viewred_n = 0;
  // views into d_sys (all-reduced once per linear solve): Sval | rhs(n)
  double *p_Sval, *p_rhs, *p_Sbb, *p_Cw, *p_Spack;
  bool s_packed = false;
  DevBuf<double> d_Sfull;
  size_t sys_n = 0;
  // what the host reads once per step attempt.  The block is pinned AND device-mapped: k_publish writes it straight over PCIe and
  // raises `seq` behind a system-scope fence; the host spins on `seq` instead of paying three cudaMemcpyAsync + a stream
  // synchronise (measured: the device sat idle ~60 us per round trip, two round trips per accepted LM step).
  PublishBlock* h_pub = nullptr;
  unsigned long long pub_seq = 0;
  bool pub_spin = true;         // PTZ_SYNC_COPY=1: the old copy + synchronise path (A/B measurements)
  double* h_scalars = nullptr;  // = h_pub->sc
  int* h_info = nullptr;        // = h_pub->info: pcg info(2), fail(1)
  int rj_per = 1, rj_grid = 1, ow_per = 1, ow_grid = 1;  // k_resjac / k_obs_what launch shapes (persistent CTAs)
  int nblk_ray = 0, nblk_cam = 0, cg_cap = 1, cg_wpb = 8, cg_grid = 1, cg_slots_per_rank = 1;
  size_t ar_partial = 0, ar_st0 = 0, ar_st1 = 0, ar_x = 0, ar_ll0 = 0, ar_ll1 = 0;  // arena offsets (bytes), identical on every rank
  int cgW = 1, cgR = 0, cg_vranks = 1;  // ranks sharing the rows of the CG, this rank; virtual ranks (debug: PTZ_CG_VRANKS)
  DevBuf<int> d_cg_col;                 // column indices, bit 31 = owned by another rank than the row
  DevBuf<unsigned char> d_cg_mask;      // per row: ranks that own one of its neighbours
  DevBuf<unsigned char> d_cg_owner;     // per row: owning rank
  double* arena_ptr(size_t off) const { return reinterpret_cast<double*>(g_arena.base[cgR] + off); }
  int cg_defl_rows = 0;  // rows per warp with their deflation data in shared memory
  // dynamic shared memory of k_cg: S blocks + column indices (16-byte aligned), then the deflation rows
  size_t cg_smem_bytes(bool deflated) const {
    const size_t blocks = (((size_t)cg_wpb * cg_cap * (NCL * NCL * sizeof(double) + sizeof(int)) + 15) / 16) * 16;
    return blocks + (deflated ? (size_t)cg_wpb * cg_defl_rows * 3 * NCL * kDeflK * sizeof(double) : 0) + 16;
  }
  template <bool MULTI, int KD>
  const void* cg_kernel_sel() const { return cg_wpb == 8 ? (const void*)k_cg<NCL, 256, MULTI, KD> : (const void*)k_cg<NCL, 512, MULTI, KD>; }
  const void* cg_kernel(bool deflated) const {
    if (cgW > 1) return deflated ? cg_kernel_sel<true, kDeflK>() : cg_kernel_sel<true, 0>();
    return deflated ? cg_kernel_sel<false, kDeflK>() : cg_kernel_sel<false, 0>();
  }
  // ---- deflation of the CG (k_cg<.., KD = kDeflK>): basis harvested from the residual history of the first solve of a run
  bool defl_enabled = false, defl_allowed = false, have_W = false, defl_active = true;
  int defl_kd = 0, defl_solves = 0, defl_rejects = 0, defl_polishes = 0, last_lin_iters = -1, defl_ref_iters = 0;
  static constexpr int kHistCap = 320, kMinHarvest = 40, kDeflOffBelow = 12, kDeflOnAbove = 30;
  DevBuf<double> d_hist, d_abg, d_Wy, d_Wt, d_AW, d_Z, d_gram, d_Einv, d_c0, d_dscal, d_bcopy, d_Lfac, d_Y;
  std::vector<double> h_abg;
  double h_dscal[2] = {0, 0};
  DevBuf<double> d_prof;

  // LM state (names follow ceres::internal::TrustRegionMinimizer / LevenbergMarquardtStrategy)
  double radius = 0, decrease_factor = 2.0, x_cost = 0, x_norm = 0, min_cost = 0, initial_cost = 0, grad_max = 0;
  bool reuse_diagonal = false, last_successful = true, started = false, finished = false;
  int iteration = 0, num_consecutive_invalid = 0, termination = PTZ_NO_CONVERGENCE;
  int num_successful = 0, num_unsuccessful = 0, lin_iters_total = 0, jac_evals = 0, cost_evals = 0;
  long long life_pcg = 0, life_lm = 0;  // since create, not cleared by reset (bench.py takes differences)
  std::vector<ptz_iter_log> log;

  BaSolver(const ptzba_problem* prob, const ptz_solver_options* o) {
    auto t0 = std::chrono::steady_clock::now();
    const bool dbg = getenv("PTZ_SETUP_DEBUG") != nullptr;  // prints where the set-up time goes (adds a stream sync per phase)
    auto phase = [&](const char* what) {
      if (!dbg) return;
      cudaStreamSynchronize(stream);
      fprintf(stderr, "[ptzba setup] %-28s %8.3f ms\n", what, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    opt = *o;
    V = prob->num_views; P = prob->num_tracks; M = prob->num_obs; A = prob->num_pts3d;
    int dev = 0;
    PTZ_CUDA(cudaGetDevice(&dev));
    PTZ_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    enable_memory_pool();
    sh.create();
    stream = sh.s;
    clk.init(stream);
    phase("stream + clock");
    if (opt.verbose & 4) {
      // reference path: the unit-tested host builder
      BaStructure st;
      build_structure(V, P, M, prob->obs_uv, prob->obs_view, prob->obs_track, kChunk, st, nullptr);
      if (g_nccl.world > 1) {
        std::vector<int64_t> keys(st.nub());
        for (int b = 0; b < st.nub(); ++b) keys[b] = ub_key(st.ub_row[b], st.ub_col[b]);
        std::vector<int64_t> all = union_of_keys(keys);
        build_structure(V, P, M, prob->obs_uv, prob->obs_view, prob->obs_track, kChunk, st, &all);
      }
      upload_structure(st, ds, stream);
    } else {
      build_structure_device_obs(V, P, M, prob->obs_uv, prob->obs_view, prob->obs_track, kChunk, ds, stream);
      phase("orderings (obs)");
      if (g_nccl.world > 1) {
        // every rank must hold the same block pattern of S: all-gather the local upper-block keys, union on the device
        const int W = g_nccl.world;
        DevBuf<int64_t> d_cnt, d_my;
        d_cnt.alloc(W, stream);
        int64_t my = ds.n_local;
        std::vector<int64_t> counts(W);
        d_my.upload(&my, 1, stream);
        PTZ_NCCL(ncclAllGather(d_my.p, d_cnt.p, 1, ncclInt64, g_nccl.comm, stream));
        d_cnt.download(counts.data(), W, stream);
        PTZ_CUDA(cudaStreamSynchronize(stream));
        const int64_t mx = std::max<int64_t>(1, *std::max_element(counts.begin(), counts.end()));
        DevBuf<unsigned long long> d_pad, d_all, d_union;
        d_pad.alloc((size_t)mx, stream);
        PTZ_CUDA(cudaMemsetAsync(d_pad.p, 0xff, (size_t)mx * 8, stream));
        if (my) PTZ_CUDA(cudaMemcpyAsync(d_pad.p, ds.uniq_local.p, (size_t)my * 8, cudaMemcpyDeviceToDevice, stream));
        d_all.alloc((size_t)mx * W, stream);
        PTZ_NCCL(ncclAllGather(d_pad.p, d_all.p, mx, ncclUint64, g_nccl.comm, stream));
        const int nun = union_keys_device(d_all.p, mx * W, d_union, stream);
        build_structure_device_blocks(ds, nullptr, stream, d_union.p, nun);
      } else {
        build_structure_device_blocks(ds, nullptr, stream);
      }
    }
    nub_local = ds.nub;
    if (g_nccl.world > 1 && ds.nub > 0) {
      std::vector<int64_t> poff((size_t)ds.nub + 1);
      ds.pair_off.download(poff.data(), poff.size(), stream);
      PTZ_CUDA(cudaStreamSynchronize(stream));
      std::vector<int> list;
      for (int b = 0; b < ds.nub; ++b) if (poff[b + 1] > poff[b]) list.push_back(b);
      if ((int)list.size() < ds.nub) { nub_local = (int)list.size(); if (list.empty()) list.push_back(0); d_ub_list.upload(list, stream); }
    }
    {
      const char* e = getenv("PTZ_OD_GROUP");  // tuning hook: 32 / 16 / 8 force the variant
      od_group = e ? atoi(e) : (ds.npairs < 64ll * std::max(nub_local, 1) ? 8 : 16);
      if (NCL != 4 || (od_group != 16 && od_group != 8)) od_group = 32;
    }
    phase("block pattern + pair lists");
    // annotated points: sort by view, list annotated views
    h_ann_idx.assign(V, -1);
    if (A > 0) {
      h_pt_perm.resize(A);
      std::iota(h_pt_perm.begin(), h_pt_perm.end(), 0);
      std::stable_sort(h_pt_perm.begin(), h_pt_perm.end(), [&](int a, int b) { return prob->pt_view[a] < prob->pt_view[b]; });
      for (int i = 0; i < A; ++i) {
        int v = prob->pt_view[h_pt_perm[i]];
        if (h_ann_view.empty() || h_ann_view.back() != v) { h_ann_idx[v] = (int)h_ann_view.size(); h_ann_view.push_back(v); h_ann_off.push_back(i); }
      }
      h_ann_off.push_back(A);
      nav = (int)h_ann_view.size();
      bo_tlw = 0; nb = 6;
      if (kFyBorder) nf = nav;
    }
    if (kDisp) { bo_disp = nb; nb += 3; }
    nb_plain = nb;
    // shared intrinsics blocks (SetSharedIntrinsics): the block of an id is the one of its first view (ptzray_optimizer.cc:640-650)
    h_rep.resize(V); h_grp_of.assign(V, -1); h_intr_counted.assign(V, 1);
    {
      std::map<int, int> first;
      std::vector<int> gsize(V, 0);
      for (int i = 0; i < V; ++i) {
        const int id = prob->shared_ic_id ? prob->shared_ic_id[i] : i;
        auto it = first.find(id);
        if (it == first.end()) { first[id] = i; h_rep[i] = i; } else h_rep[i] = it->second;
        ++gsize[h_rep[i]];
      }
      std::vector<int> gidx(V, -1);
      for (int i = 0; i < V; ++i) if (h_rep[i] == i && gsize[i] >= 2) gidx[i] = ngrp++;
      h_grp_off.assign(ngrp + 1, 0);
      for (int i = 0; i < V; ++i) {
        h_grp_of[i] = gidx[h_rep[i]];
        if (h_grp_of[i] >= 0) { ++h_grp_off[h_grp_of[i] + 1]; if (h_rep[i] != i) h_intr_counted[i] = 0; }
      }
      for (int q = 0; q < ngrp; ++q) h_grp_off[q + 1] += h_grp_off[q];
      h_grp_view.resize(h_grp_off[ngrp]);
      std::vector<int> fill(h_grp_off.begin(), h_grp_off.end() - 1);
      for (int i = 0; i < V; ++i) if (h_grp_of[i] >= 0) h_grp_view[fill[h_grp_of[i]]++] = i;
    }
    ns = ngrp * (NCL - 3);
    if (ns > 0 && A > 0) throw CudaError(PTZ_ERR_UNSUPPORTED, "shared intrinsics together with 2d-3d terms");
    bo_sh = nb; nb += ns;
    if (nb > kMaxBorder)
      throw CudaError(PTZ_ERR_UNSUPPORTED, "more shared intrinsics unknowns than the dense border holds (16 minus tlw / disp): too many groups");
    nbt = nb + nf;
    static_assert(kMaxBorder >= 9, "tlw(6) + disp(3)");
    h_ann_strip.resize(nav);
    if (kDisp || ns > 0) {
      ncpl = V;
      h_cpl_view.resize(V); std::iota(h_cpl_view.begin(), h_cpl_view.end(), 0);
      h_cpl_idx = h_cpl_view;
      for (int k = 0; k < nav; ++k) h_ann_strip[k] = h_ann_view[k];
    } else {
      ncpl = nav;
      h_cpl_view = h_ann_view;
      h_cpl_idx = h_ann_idx;
      for (int k = 0; k < nav; ++k) h_ann_strip[k] = k;
    }
    n = V * NCL + nb;
    {
      // CG launch shape: one CTA per SM, as many warps per CTA (8/16) as it takes to give every warp at most one row
      // where possible.  Rows are handed out in Cuthill-McKee order of the view graph, a contiguous run per CTA, so that the
      // rows of a CTA are neighbouring views whose gathers overlap (L1 hits).  Then: how many blocks of S fit 200 KB of smem.
      // Sharded problem: the rows are split across the ranks (a contiguous run of the ordering each), see k_cg.
      const int nrows = V + (nb > 0 ? 1 : 0);
      // a system with a border row is solved replicated (every rank runs the single-GPU kernel on the all-reduced system): the
      // border row would sit on one rank and serialise the others behind an NVLink hop per iteration
      cgW = (nb > 0) ? 1 : g_nccl.world; cgR = g_nccl.rank; cg_vranks = 1;
      rows_sharded = cgW > 1;
      if (g_nccl.world == 1 && nb == 0) {
        const char* e = getenv("PTZ_CG_VRANKS");
        const int vr = e ? atoi(e) : 1;
        if (vr > 1 && vr <= kMaxPeers && V >= 2 * vr) { cg_vranks = vr; cgW = vr; }
      }
      const int W = cgW, R = rows_sharded ? cgR : 0;
      const int sms_per_rank = cg_vranks > 1 ? num_sms / cg_vranks : num_sms;
      cg_slots_per_rank = cdiv(nrows, W);
      const int my0 = R * cg_slots_per_rank, my1 = std::min(nrows, my0 + cg_slots_per_rank);
      const int need = cdiv(cg_slots_per_rank, sms_per_rank);
      cg_wpb = need <= 8 ? 8 : 16;  // beyond 16 rows per SM a warp walks several rows
      if (const char* e = getenv("PTZ_CG_WPB")) { const int w = atoi(e); if (w == 8 || w == 16) cg_wpb = w; }  // tuning hook
      cg_grid = std::min(sms_per_rank, cdiv(cg_slots_per_rank, cg_wpb));  // CTAs per rank
      std::vector<int> h_col(ds.nnzb);
      ds.s_col.download(h_col.data(), ds.nnzb, stream);
      PTZ_CUDA(cudaStreamSynchronize(stream));
      std::vector<int> order;
      order.reserve(V);
      {
        std::vector<char> seen(V, 0);
        std::vector<int> deg(V), byd(V);
        for (int v = 0; v < V; ++v) { deg[v] = ds.h_rowptr[v + 1] - ds.h_rowptr[v]; byd[v] = v; }
        std::stable_sort(byd.begin(), byd.end(), [&](int a, int b) { return deg[a] < deg[b]; });
        std::vector<int> nbrs;
        for (int start : byd) {
          if (seen[start]) continue;
          size_t head = order.size();
          order.push_back(start); seen[start] = 1;
          while (head < order.size()) {
            const int v = order[head++];
            nbrs.clear();
            for (int k = ds.h_rowptr[v]; k < ds.h_rowptr[v + 1]; ++k) if (!seen[h_col[k]]) { nbrs.push_back(h_col[k]); seen[h_col[k]] = 1; }
            std::stable_sort(nbrs.begin(), nbrs.end(), [&](int a, int b) { return deg[a] < deg[b]; });
            order.insert(order.end(), nbrs.begin(), nbrs.end());
          }
        }
      }
      d_cg_order.upload(order, stream);
      const int per = cdiv(cg_slots_per_rank, cg_grid);
      int worst = 0;
      for (int r = (cg_vranks > 1 ? 0 : R); r < (cg_vranks > 1 ? W : R + 1); ++r)
        for (int c = 0; c < cg_grid; ++c)
          for (int w = 0; w < cg_wpb; ++w) {
            int cnt = 0;
            for (int sl = w; sl < per; sl += cg_wpb) {
              const int slot = r * cg_slots_per_rank + c * per + sl;
              if (slot >= std::min(V, std::min(nrows, (r + 1) * cg_slots_per_rank))) break;
              cnt += ds.h_rowptr[order[slot] + 1] - ds.h_rowptr[order[slot]];
            }
            worst = std::max(worst, cnt);
          }
      (void)my0; (void)my1;
      {
        // ownership marks for the sharded CG: which columns of a row live on another rank, which ranks need a row's state
        std::vector<int> owner(V, 0), colx(h_col);
        for (int sl = 0; sl < V; ++sl) owner[order[sl]] = sl / cg_slots_per_rank;
        std::vector<unsigned char> mask(V, 0);
        for (int r = 0; r < V; ++r)
          for (int k = ds.h_rowptr[r]; k < ds.h_rowptr[r + 1]; ++k) {
            const int c = h_col[k];
            if (owner[c] != owner[r]) { colx[k] = c | (int)0x80000000; mask[c] |= (unsigned char)(1u << owner[r]); }
          }
        d_cg_col.upload(colx, stream);
        d_cg_mask.upload(mask, stream);
        std::vector<unsigned char> own8(owner.begin(), owner.end());
        d_cg_owner.upload(own8, stream);
      }
      const size_t per_block = NCL * NCL * sizeof(double) + sizeof(int);
      const int fit = (int)((150 * 1024) / (cg_wpb * per_block));  // leave >= 60 KB of the SM's 228 KB to the L1 (the kernel has ~8 KB of static shared memory)
      cg_cap = std::max(1, std::min(worst, fit));
      if (rows_sharded) {  // every rank launches the same shape (the slot of a CTA's partial sums is rank * grid + cta)
        DevBuf<double> d_m;
        double m[2] = {(double)cg_cap, 0.0};
        d_m.upload(m, 2, stream);
        PTZ_NCCL(ncclAllReduce(d_m.p, d_m.p, 2, ncclDouble, ncclMax, g_nccl.comm, stream));
        d_m.download(m, 2, stream);
        PTZ_CUDA(cudaStreamSynchronize(stream));
        cg_cap = (int)m[0];
      }
      // deflation data of the first rows of every warp go to shared memory too (W, AW, Z entries: 3 * NCL * kDeflK doubles a row)
      {
        const size_t left = (size_t)190 * 1024 - std::min((size_t)190 * 1024, cg_smem_bytes(false));
        const int rows_per_warp = cdiv(cdiv(cg_slots_per_rank, cg_grid), cg_wpb);
        cg_defl_rows = std::min(rows_per_warp, (int)(left / ((size_t)cg_wpb * 3 * NCL * kDeflK * sizeof(double))));
      }
      if (cg_grid > kCgMaxCtas) throw CudaError(PTZ_ERR_UNSUPPORTED, "more SMs than the CG kernel's reduction buffer holds");
      PTZ_CUDA(cudaFuncSetAttribute(cg_kernel(false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cg_smem_bytes(false)));
      PTZ_CUDA(cudaFuncSetAttribute(cg_kernel(true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cg_smem_bytes(true)));
      // arena layout: ctrl | slots [2][W*grid] | st0 [3n] | st1 [3n] | x [n] | LL inboxes ll0, ll1 [V*3*NCL] 16-byte words
      auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
      ar_partial = kArenaCtrlBytes;
      ar_st0 = al(ar_partial + cg_slot_region_bytes(num_sms));
      ar_st1 = al(ar_st0 + 3 * (size_t)n * sizeof(double));
      ar_x = al(ar_st1 + 3 * (size_t)n * sizeof(double));
      ar_ll0 = al(ar_x + (size_t)n * sizeof(double));
      const size_t ll_bytes = W > 1 ? (size_t)V * 3 * NCL * 16 : 0;
      ar_ll1 = al(ar_ll0 + ll_bytes);
      arena_ensure(al(ar_ll1 + ll_bytes), stream, cg_vranks);
    }
    phase("CG shape + arena");
    upload(prob);
    PTZ_CUDA(cudaStreamSynchronize(stream));
    phase("parameters + work buffers");
    seconds_setup = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }

  ~BaSolver() {
    cudaStreamSynchronize(stream);
    if (d_prof.n) {
      double pr[16];
      cudaMemcpy(pr, d_prof.p, sizeof(pr), cudaMemcpyDeviceToHost);
      const char* names[8] = {"rows (update + product)", "reduction + barrier (total)", "mu", "  cta reduce", "  store + arrive + spin", "  fetch partials", "  sum partials", "loop head"};
      for (int k = 0; k < 2; ++k) {
        fprintf(stderr, "[k_cg profile, %s] cycles of CTA 0 / thread 0 over all solves:\n", k ? "deflated" : "plain");
        for (int i = 0; i < 8; ++i) fprintf(stderr, "   %-30s %14.0f\n", names[i], pr[8 * k + i]);
      }
    }
    HostCache::put_pinned(h_pub);
  }

  // every rank must hold the same block pattern of S: all-gather the local upper block keys, return their sorted union
  std::vector<int64_t> union_of_keys(const std::vector<int64_t>& keys) {
    const int W = g_nccl.world;
    DevBuf<int64_t> d_cnt, d_all, d_my, d_pad;
    d_cnt.alloc(W, stream);
    int64_t my = (int64_t)keys.size();
    std::vector<int64_t> counts(W);
    d_my.upload(&my, 1, stream);
    PTZ_NCCL(ncclAllGather(d_my.p, d_cnt.p, 1, ncclInt64, g_nccl.comm, stream));
    d_cnt.download(counts.data(), W, stream);
    PTZ_CUDA(cudaStreamSynchronize(stream));
    int64_t mx = *std::max_element(counts.begin(), counts.end());
    if (mx == 0) return std::vector<int64_t>();
    std::vector<int64_t> padded(mx, -1);
    std::copy(keys.begin(), keys.end(), padded.begin());
    d_pad.upload(padded, stream);
    d_all.alloc((size_t)mx * W, stream);
    PTZ_NCCL(ncclAllGather(d_pad.p, d_all.p, mx, ncclInt64, g_nccl.comm, stream));
    std::vector<int64_t> all((size_t)mx * W);
    d_all.download(all.data(), all.size(), stream);
    PTZ_CUDA(cudaStreamSynchronize(stream));
    all.erase(std::remove(all.begin(), all.end(), (int64_t)-1), all.end());
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    return all;
  }

  void upload(const ptzba_problem* prob) {
    cudaStream_t s = stream;
    if (A > 0) {  // views that carry only annotated points are active too
      std::vector<int> act(V);
      ds.view_active.download(act.data(), V, s);
      PTZ_CUDA(cudaStreamSynchronize(s));
      for (int v : h_ann_view) act[v] = 1;
      ds.view_active.upload(act, s);
    }
    // parameters
    h_intr0.assign(prob->intr, prob->intr + 9 * (size_t)V);
    h_ext0.assign(prob->ext, prob->ext + 6 * (size_t)V);
    h_tlw0.assign(6, 0.0);
    if (prob->tlw0) h_tlw0.assign(prob->tlw0, prob->tlw0 + 6);
    have_ray0 = prob->ray0 != nullptr;
    d_intr_init.upload(h_intr0, s); d_ext_init.upload(h_ext0, s); d_tlw_init.upload(h_tlw0, s);
    if (ns > 0) { PTZ_CUDA(cudaStreamSynchronize(s)); for (int i = 0; i < V; ++i) for (int j = 0; j < 9; ++j) h_intr0[9 * (size_t)i + j] = prob->intr[9 * (size_t)h_rep[i] + j]; }
    // track records [ray(3), sqrt(weight), Jacobi scale(3), pad] are laid out on the device from the caller's arrays
    const size_t trk_n = (size_t)std::max(P, 1) * kTrk;
    d_trk_init.alloc(trk_n, s);
    if (P > 0) {
      DevBuf<double> d_w, d_r0;
      d_w.upload(prob->track_weight, (size_t)P, s);
      if (have_ray0) d_r0.upload(prob->ray0, 3 * (size_t)P, s);
      k_trk_init<<<cdiv(P, 256), 256, 0, s>>>(P, d_w.p, have_ray0 ? d_r0.p : nullptr, d_trk_init.p);
      PTZ_CUDA(cudaGetLastError());
    } else {
      d_trk_init.zero(s);
    }
    for (int i = 0; i < 2; ++i) { d_intr[i].alloc(9 * (size_t)V, stream); d_ext[i].alloc(6 * (size_t)V, stream); d_trk[i].alloc(trk_n, stream); d_tlw[i].alloc(6, stream); }
    d_vt.alloc(V, stream);
    d_RiKi.alloc(9 * (size_t)V, stream);
    if (!have_ray0 && P > 0) {  // Pix2Ray on the device, into the initial track records
      k_rikI<<<cdiv(V, 128), 128, 0, s>>>(V, d_intr_init.p, d_ext_init.p, d_RiKi.p);
      k_init_rays<<<cdiv(P, 128), 128, 0, s>>>(P, ds.t_off.p, ds.t_obs.p, ds.o_view.p, ds.o_uv.p, d_RiKi.p, d_trk_init.p);
      PTZ_CUDA(cudaGetLastError());
    }
    if (ns > 0) {
      // (Pix2Ray above used every view's own K, as the reference does: :768-797 read cameras_.)  The solve starts from the block of
      // the group's first view
      d_intr_init.upload(h_intr0, s);
      d_grp_of.upload(h_grp_of, s); d_grp_off.upload(h_grp_off, s); d_grp_view.upload(h_grp_view, s); d_intr_counted.upload(h_intr_counted, s);
      d_sh_h.alloc(ns, s); d_sh_h.zero(s);
      d_rowpart.alloc((size_t)V * (NCL - 3) * (nb + 1), s); d_rowpart.zero(s);
    }
    // annotated points
    if (A > 0) {
      std::vector<float> puv(2 * (size_t)A);
      std::vector<double> pxyz(3 * (size_t)A);
      std::vector<int> pview(A);
      for (int i = 0; i < A; ++i) {
        int k = h_pt_perm[i];
        puv[2 * i] = prob->pt_uv[2 * k]; puv[2 * i + 1] = prob->pt_uv[2 * k + 1];
        for (int j = 0; j < 3; ++j) pxyz[3 * i + j] = prob->pt_xyz[3 * k + j];
        pview[i] = prob->pt_view[k];
      }
      d_pts_uv.upload(reinterpret_cast<const float2*>(puv.data()), A, s);
      d_pts_xyz.upload(pxyz, s); d_pts_view.upload(pview, s);
      d_ann_view.upload(h_ann_view, s); d_ann_off.upload(h_ann_off, s);
      d_pts_scratch.alloc((size_t)A * (2 + 2 * NCL + 2 * nb + 2), stream);
      d_pts_raw.alloc((size_t)A * 32, stream);
    }
    d_ann_idx.upload(h_ann_idx, s);
    if (h_cpl_view.empty()) d_cpl_view.alloc(1, s); else d_cpl_view.upload(h_cpl_view, s);
    d_cpl_idx.upload(h_cpl_idx, s);
    if (h_ann_strip.empty()) d_ann_strip.alloc(1, s); else d_ann_strip.upload(h_ann_strip, s);
    {
      std::vector<double> z3(3, 0.0);
      d_disp_init.upload(z3, s);
      d_dispp[0].alloc(3, s); d_dispp[1].alloc(3, s);
    }
    if (kDisp) {
      d_recd.alloc((size_t)std::max(M, 1) * 6, s); d_dpart.alloc((size_t)V * 9, s); d_Wdh.alloc((size_t)std::max(P, 1) * 12, s);
    }
    if (!kDisp && (nf > 0 || ns > 0)) d_Cw.alloc((size_t)std::max(ncpl, 1) * NCL * nb, s);  // working copy of the coupling strips of one linear solve
    d_hinv.alloc(std::max(nf, 1), s);
    // work buffers
    d_scale_cam.alloc((size_t)V * NCL, stream); d_scale_b.alloc(std::max(nbt, 1), stream);
    d_recA.alloc((size_t)std::max(M, 1) * D::RA, stream); d_recF.alloc((size_t)std::max(M, 1) * D::RF, stream);
    {
      // k_resjac runs persistent CTAs, one resident wave, each over a contiguous run of chunks (<= kResjacMaxPer of them)
      int occ = 1;
      PTZ_CUDA(cudaFuncSetAttribute(k_resjac<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ResjacSmem<NCL>::kBytes));
      PTZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_resjac<TYPE>, kChunk, ResjacSmem<NCL>::kBytes));
      const int wave = std::max(1, num_sms * std::max(occ, 1));
      rj_per = std::min(kResjacMaxPer, std::max(1, cdiv(ds.nchunks, wave)));
      if (const char* e = getenv("PTZ_RJ_PER")) rj_per = std::min(kResjacMaxPer, std::max(1, atoi(e)));  // tuning hook
      rj_grid = std::max(1, cdiv(ds.nchunks, rj_per));
      // k_obs_what likewise (dynamic shared memory: two record buffers + staging)
      PTZ_CUDA(cudaFuncSetAttribute(k_obs_what<NCL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ObsWhatSmem<NCL>::kBytes));
      PTZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_obs_what<NCL>, kChunk, ObsWhatSmem<NCL>::kBytes));
      const int wave2 = std::max(1, num_sms * std::max(occ, 1));
      ow_per = std::min(kResjacMaxPer, std::max(1, cdiv(ds.nchunks, wave2)));
      if (const char* e = getenv("PTZ_OW_PER")) ow_per = std::min(kResjacMaxPer, std::max(1, atoi(e)));
      ow_grid = std::max(1, cdiv(ds.nchunks, ow_per));
    }
    d_part.alloc((size_t)std::max(ds.nchunks, 1) * D::NPART, stream); d_part.zero(s);  // (only the last chunk of a view-run is ever written)
    viewred_n = (size_t)V * NCL * NCL + (size_t)V * NCL + V + (size_t)ncpl * NCL * nb + (size_t)nb * nb + nbt + 2 + (size_t)nf * (NCL + nb + 1);
    d_viewred.alloc(viewred_n, stream);
    d_viewred.zero(s);
    p_U = d_viewred.p; p_g = p_U + (size_t)V * NCL * NCL; p_cost_view = p_g + (size_t)V * NCL; p_C = p_cost_view + V;
    p_Hbb = p_C + (size_t)ncpl * NCL * nb; p_gb = p_Hbb + (size_t)nb * nb; p_cost_pts = p_gb + nbt;
    p_Cf = p_cost_pts + 2; p_Hrf = p_Cf + (size_t)nf * NCL; p_Hff = p_Hrf + (size_t)nf * nb;
    d_gabs.alloc((size_t)V * NCL + std::max(nbt, 1), stream);
    d_gabs.zero(s);
    p_gabs = d_gabs.p; p_gabs_b = d_gabs.p + (size_t)V * NCL;
    d_Vh.alloc((size_t)std::max(P, 1) * 10, stream); d_Vh.zero(s);  // (tracks without observations are never written)
    d_t_view.alloc(std::max(M, 1), stream);
    if (M > 0) k_gather_int<<<cdiv(M, 256), 256, 0, s>>>(M, ds.t_obs.p, ds.o_view.p, d_t_view.p);
    nblk_ray = std::max(cdiv(P, 128), 1); nblk_cam = std::max(cdiv(V, 128), 1);
    {
      // by-track passes: a lane per (track, observation) entry, a warp per group of 32 entries (k_track_accum_w / k_track_backsub_w)
      // OPT-IN (PTZ_BYTRACK_WARP=1): measured SLOWER than one thread per track at cfg 4 (k_track_accum 62 -> 110 us, k_track_backsub
      // 87 -> 131 us, the 62 k per-warp partials cost k_scalars 25 us more): a warp that handles only 32 records pays the dependent chain
      // group -> indices -> records -> head epilogue five times as often as a warp of 32 tracks.  Kept for the record and for a
      // persistent, pipelined version; parity-tested (pytest with the variable set).
      const char* e = getenv("PTZ_BYTRACK_WARP");
      track_groups = (M > 0 && P > 0 && e && atoi(e) != 0) ? cdiv(M, 32) : 0;
      if (track_groups > 0) {
        d_t_trk.alloc(M, s); d_grp_e0.alloc(track_groups, s);
        PTZ_CUDA(cudaMemsetAsync(d_grp_e0.p, 0x7f, (size_t)track_groups * sizeof(int), s));
        k_entry_tracks<<<cdiv(P, 256), 256, 0, s>>>(P, ds.t_off.p, d_t_trk.p, d_grp_e0.p);
        PTZ_CUDA(cudaGetLastError());
      }
      nray_parts = track_groups > 0 ? track_groups : nblk_ray;
    }
    d_gmax_part.alloc(nray_parts, stream); d_gmax_part.zero(s);
    d_diag_ray.alloc(3 * (size_t)std::max(P, 1), stream); d_diag_cam.alloc((size_t)V * NCL, stream); d_diag_b.alloc(std::max(nbt, 1), stream);
    d_Lt.alloc((size_t)std::max(P, 1) * 10, stream);
    d_What.alloc((size_t)std::max(M, 1) * D::WS, stream); d_What.zero(s);
    d_q.alloc((size_t)std::max(ds.nchunks, 1) * (D::NU + NCL), stream); d_q.zero(s);  // chunk partials of sum What What^T, sum q (likewise)
    // d_sys (all-reduced once per linear solve): Sval | rhs(n) | Sbb(nb^2) | PTZRayDistDisp: the working strips Cw (they carry
    // rank-local Schur terms of the rays)
    // Sharded problem: S goes through the all-reduce packed (diagonal blocks + upper blocks once, k_unpack_S), the block-CSR copy
    // the later kernels work on is a buffer of its own
    s_packed = g_nccl.world > 1;
    const size_t s_reduced = s_packed ? (size_t)(V + ds.nub) * NCL * NCL : (size_t)ds.nnzb * NCL * NCL;
    sys_n = s_reduced + n + (size_t)nb * nb + (kDisp ? (size_t)ncpl * NCL * nb : 0);
    d_sys.alloc(sys_n, stream);
    if (s_packed) { d_Sfull.alloc((size_t)ds.nnzb * NCL * NCL, stream); p_Sval = d_Sfull.p; } else p_Sval = d_sys.p;
    p_Spack = d_sys.p; p_rhs = d_sys.p + s_reduced; p_Sbb = p_rhs + n;
    p_Cw = kDisp ? p_Sbb + (size_t)nb * nb : d_Cw.p;
    d_Linv.alloc((size_t)V * NCL * NCL, stream); d_Linv_b.alloc(kMaxBorder * kMaxBorder, stream);
    d_Cs.alloc((size_t)std::max(ncpl, 1) * NCL * std::max(nb, 1), stream);
    {
      // deflated CG (camera-only reduced systems of some size; PTZ_CG_DEFLATE=0 switches it off for A/B measurements)
      const char* e = getenv("PTZ_CG_DEFLATE");
      defl_enabled = (!e || atoi(e) != 0) && nb == 0 && V >= 64;  // (tried V >= 48 for cfg 2, V = 60: the deflated solves get rejected and redone there, 965 -> 1480 us per LM iteration)
      have_W = false;
      if (defl_enabled) {
        const size_t nk = (size_t)V * NCL * kDeflK;
        d_hist.alloc((size_t)kHistCap * V * NCL, stream);
        d_abg.alloc(3 * (size_t)kHistCap, stream); d_abg.zero(s);
        d_Wy.alloc(nk, stream); d_Wt.alloc(nk, stream); d_AW.alloc(nk, stream); d_Z.alloc(nk, stream);
        d_gram.alloc((size_t)cdiv(V * NCL, kDeflGramChunk) * kDeflGramVals, stream);
        d_Einv.alloc(kDeflK * kDeflK, stream); d_c0.alloc(kDeflK, stream); d_dscal.alloc(2, stream); d_dscal.zero(s);
        d_bcopy.alloc((size_t)V * NCL, stream); d_Lfac.alloc((size_t)V * NCL * NCL, stream); d_Y.alloc((size_t)kHistCap * kDeflK, stream);
        h_abg.assign(3 * (size_t)kHistCap, 0.0);
      }
    }
    {
      const char* e = getenv("PTZ_DENSE_MAX_N");  // (0 switches the dense path off: A/B measurements, tests of the CG on small scenes)
      const int cap = e ? atoi(e) : kDenseMaxN;
      use_dense = n <= std::min(cap, (int)kDenseMaxN) && cg_vranks == 1;
      if (use_dense) {
        d_dense.alloc((size_t)(n + 1) * n, stream);
        PTZ_CUDA(cudaFuncSetAttribute(k_dense_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem_bytes()));
        defl_enabled = false;
      }
    }
    defl_allowed = defl_enabled;
    d_cgp.alloc((size_t)n, stream); d_cgp.zero(s);
    d_y.alloc(n, stream); d_y.zero(s);
    d_pcg_res.alloc(2, stream); d_pcg_info.alloc(2, stream); d_fail.alloc(1, stream); d_fail.zero(s);
    d_part3_ray.alloc(3 * (size_t)nray_parts, stream); d_part3_ray.zero(s);
    d_part3_cam.alloc(3 * (size_t)nblk_cam, stream); d_part3_b.alloc(3, stream); d_part3_b.zero(s);
    d_cost_part.alloc(2 * (size_t)std::max(ds.nchunks, 1), stream); d_cost_part.zero(s);
    d_scalars.alloc(S_COUNT, stream); d_scalars.zero(s);

    static_assert(sizeof(PublishBlock) <= HostCache::kPinnedBytes && S_COUNT == 32, "pinned scratch block too small");
    h_pub = reinterpret_cast<PublishBlock*>(HostCache::get_pinned());
    memset(h_pub, 0, sizeof(PublishBlock));
    pub_seq = 0;
    h_scalars = h_pub->sc;
    h_info = h_pub->info;
    { const char* e = getenv("PTZ_SYNC_COPY"); pub_spin = !(e && atoi(e) != 0); }
    reset();
  }

  void reset() override {
    cudaStream_t s = stream;
    PTZ_CUDA(cudaMemcpyAsync(d_intr[0].p, d_intr_init.p, 9 * (size_t)V * 8, cudaMemcpyDeviceToDevice, s));
    PTZ_CUDA(cudaMemcpyAsync(d_ext[0].p, d_ext_init.p, 6 * (size_t)V * 8, cudaMemcpyDeviceToDevice, s));
    for (int i = 0; i < 2; ++i) {  // both copies: without a border the candidate copy is never written, yet `cur` flips to it
      PTZ_CUDA(cudaMemcpyAsync(d_tlw[i].p, d_tlw_init.p, 6 * 8, cudaMemcpyDeviceToDevice, s));
      PTZ_CUDA(cudaMemcpyAsync(d_dispp[i].p, d_disp_init.p, 3 * 8, cudaMemcpyDeviceToDevice, s));
      PTZ_CUDA(cudaMemcpyAsync(d_intr[i].p, d_intr_init.p, 9 * (size_t)V * 8, cudaMemcpyDeviceToDevice, s));
      PTZ_CUDA(cudaMemcpyAsync(d_ext[i].p, d_ext_init.p, 6 * (size_t)V * 8, cudaMemcpyDeviceToDevice, s));
    }
    for (int i = 0; i < 2; ++i) PTZ_CUDA(cudaMemcpyAsync(d_trk[i].p, d_trk_init.p, d_trk_init.n * 8, cudaMemcpyDeviceToDevice, s));
    cur = 0;
    vt_for = -1; vt_scaled = false;
    fuse_factor_mu = 0.0; lt_mu = 0.0;
    started = finished = false;
    iteration = 0; num_consecutive_invalid = 0; num_successful = num_unsuccessful = lin_iters_total = jac_evals = cost_evals = 0;
    radius = opt.initial_trust_region_radius; decrease_factor = 2.0; reuse_diagonal = false; last_successful = true;
    termination = PTZ_NO_CONVERGENCE;
    log.clear();
    defl_enabled = defl_allowed; defl_rejects = 0;   // (a run that gave deflation up does not decide for the next one)
    if (have_W || (defl_enabled && d_hist.n == 0)) {  // every solve harvests its own deflation basis: runs are bit-reproducible
      have_W = false;
      defl_active = true; last_lin_iters = -1;
      if (defl_enabled && d_hist.n == 0) d_hist.alloc((size_t)kHistCap * V * NCL, stream);
    }
    // kernel timings keep accumulating across resets (bench.py times several solves); see ptzba_get_stage_times
  }

  // ---- stage 1 at the current point.  scale arrays must be valid (all ones on the very first pass).
  void launch_resjac(int weighted) {
    cudaStream_t s = stream;
    // (after an accepted step the table is already the one of the new point: the cost pass built it for the candidate)
    if (vt_for != cur || !vt_scaled)
      PTZ_TIMED(PTZ_K_VIEW_PREP, k_view_prep<<<cdiv(V, 128), 128, 0, s>>>(V, d_intr[cur].p, d_ext[cur].p, d_vt.p, 1, d_scale_cam.p, NCL));
    vt_for = cur; vt_scaled = true;
    if (nb > 0) PTZ_CUDA(cudaMemsetAsync(p_C, 0, (viewred_n - (size_t)(p_C - d_viewred.p)) * sizeof(double), s));  // C | Hbb | gb | cost_pts accumulate
    if (ds.nchunks > 0)
      PTZ_TIMED(PTZ_K_RESJAC, k_resjac<TYPE><<<rj_grid, kChunk, ResjacSmem<NCL>::kBytes, s>>>(ds.nchunks, rj_per, ds.chunk_view.p, ds.chunk_begin.p, ds.chunk_cnt.p,
                                                                                                   ds.o_uv.p, ds.o_track.p, d_vt.p, d_trk[cur].p, d_dispp[cur].p, weighted,
                                                                                                   d_recA.p, d_recF.p, d_part.p, d_recd.p,
                                                                                                   d_scale_b.p + (kDisp ? bo_disp : 0)));
    PTZ_TIMED(PTZ_K_VIEW_FINALIZE,
              k_view_finalize<NCL><<<cdiv(V * D::NPART, 256), 256, 0, s>>>(V, ds.view_chunk_off.p, d_part.p, d_scale_cam.p, p_U, p_g, p_cost_view, p_gabs));
    if (P > 0)
    {
      if (track_groups > 0)
        PTZ_TIMED(PTZ_K_TRACK_ACCUM, k_track_accum_w<<<cdiv(track_groups, 8), 256, 0, s>>>(track_groups, M, d_grp_e0.p, d_t_trk.p, ds.t_obs.p, d_recA.p, d_trk[cur].p,
                                                                                            d_Vh.p, d_gmax_part.p));
      else
        PTZ_TIMED(PTZ_K_TRACK_ACCUM, k_track_accum<<<nblk_ray, 128, 0, s>>>(P, ds.t_off.p, ds.t_obs.p, d_recA.p, d_trk[cur].p, d_Vh.p, d_gmax_part.p, fuse_factor_mu,
                                                                            opt.min_lm_diagonal, opt.max_lm_diagonal, d_diag_ray.p, d_Lt.p, d_fail.p));
      lt_mu = (track_groups > 0) ? 0.0 : fuse_factor_mu;  // d_Lt now holds the factor for this radius (fresh diagonal), or nothing
      fuse_factor_mu = 0.0;
    }
    if (A > 0) {
      PtsArgs a;
      a.A = A; a.nav = nav; a.nb = nb; a.nf = nf;
      a.bo_tlw = bo_tlw; a.bo_disp = bo_disp; a.ann_strip = d_ann_strip.p; a.disp = d_dispp[cur].p;
      a.uv = d_pts_uv.p; a.xyz = d_pts_xyz.p; a.view = d_pts_view.p; a.ann_view = d_ann_view.p; a.ann_off = d_ann_off.p;
      a.vt = d_vt.p; a.tlw = d_tlw[cur].p; a.scale_cam = d_scale_cam.p; a.scale_b = d_scale_b.p; a.scratch = d_pts_scratch.p;
      a.raw = weighted ? nullptr : d_pts_raw.p;
      a.U = p_U; a.g = p_g; a.gabs = p_gabs; a.C = p_C; a.Hbb = p_Hbb; a.gb = p_gb; a.cost_pts = p_cost_pts; a.gabs_b = p_gabs_b;
      a.Cf = p_Cf; a.Hrf = p_Hrf; a.Hff = p_Hff;
      // sharded problem: the annotated points belong to rank 0 and reach the others through the all-reduce below
      if (g_nccl.rank == 0) PTZ_TIMED(PTZ_K_PTS, k_pts<TYPE><<<1, 128, 0, s>>>(a));
    }
    if (kDisp) {
      k_disp_view<NCL><<<cdiv(V, 64), 64, 0, s>>>(V, ds.view_off.p, d_recA.p, d_recF.p, d_recd.p, nb, bo_disp, p_C, d_dpart.p);
      k_disp_total<<<1, 32, 0, s>>>(V, d_dpart.p, nb, bo_disp, d_scale_b.p, p_Hbb, p_gb, p_gabs_b);
    }
    PTZ_CUDA(cudaGetLastError());
    if (g_nccl.world > 1) {
      PTZ_TIMED(PTZ_K_ALLREDUCE, allreduce_sum(d_viewred.p, viewred_n, s));
      // |g/s| was formed from rank-local sums: redo it from the reduced gradient
      k_grad_abs<<<cdiv(V * NCL + nbt, 256), 256, 0, s>>>(V * NCL, p_g, d_scale_cam.p, p_gabs, nbt, p_gb, d_scale_b.p, p_gabs_b);
      PTZ_CUDA(cudaGetLastError());
    }
    if (ns > 0) {
      k_shared_grad<NCL><<<cdiv(ns, 32), 32, 0, s>>>(ns, d_grp_off.p, d_grp_view.p, p_U, p_g, d_scale_b.p, bo_sh, p_gb, p_gabs_b, d_sh_h.p, p_gabs);
      PTZ_CUDA(cudaGetLastError());
    }
  }

  void fill_ones(double* p, size_t count) {
    std::vector<double> ones(count, 1.0);
    PTZ_CUDA(cudaMemcpyAsync(p, ones.data(), count * 8, cudaMemcpyHostToDevice, stream));
    PTZ_CUDA(cudaStreamSynchronize(stream));
  }

  // EvaluateGradientAndJacobian: cost, scaled Jacobian records, gradient max-norm
  void evaluate_jacobian(bool first) {
    if (first) {
      fill_ones(d_scale_cam.p, (size_t)V * NCL);
      fill_ones(d_scale_b.p, std::max(nbt, 1));
      if (opt.jacobi_scaling) {
        launch_resjac(1);
        k_make_scales<NCL><<<cdiv(std::max(V * NCL, P), 256), 256, 0, stream>>>(V, P, p_U, d_Vh.p, d_scale_cam.p, d_trk[0].p, d_trk[1].p);
        if (nbt > 0) k_border_scales<<<cdiv(nbt, 128), 128, 0, stream>>>(nb, nf, p_Hbb, p_Hff, d_scale_b.p, nb_plain, d_sh_h.p);
        if (ns > 0) k_shared_scales<NCL><<<cdiv(V * (NCL - 3), 128), 128, 0, stream>>>(V, d_grp_of.p, d_scale_b.p, bo_sh, d_scale_cam.p);
      }
      vt_scaled = false;  // the table carries the scales: rebuild it with the final ones
    }
    launch_resjac(1);
    ++jac_evals;
    ScalarJobs J;
    J.nsum = 0; J.nmax = 0;
    auto add_sum = [&](const double* p, int cnt, int stride, int slot) { J.sum_ptr[J.nsum] = p; J.sum_n[J.nsum] = cnt; J.sum_stride[J.nsum] = stride; J.sum_slot[J.nsum] = slot; ++J.nsum; };
    auto add_max = [&](const double* p, int cnt, int slot) { J.max_ptr[J.nmax] = p; J.max_n[J.nmax] = cnt; J.max_slot[J.nmax] = slot; ++J.nmax; };
    add_sum(p_cost_view, V, 1, S_COST_X);
    add_sum(p_cost_pts, A > 0 ? 1 : 0, 1, S_COSTPTS_X);
    add_sum(p_cost_pts + 1, A > 0 ? 1 : 0, 1, S_RAWPTS_X);
    add_max(p_gabs, V * NCL, S_GMAX_CAM);
    add_max(d_gmax_part.p, P > 0 ? nray_parts : 0, S_GMAX_RAY);
    add_max(p_gabs_b, nbt, S_GMAX_B);
    PTZ_TIMED(PTZ_K_SCALARS, k_scalars<<<J.nsum + J.nmax, 1024, 0, stream>>>(J, d_scalars.p));
    PTZ_CUDA(cudaGetLastError());
    allreduce_max(d_scalars.p + S_GMAX_RAY, S_MAX_END - S_GMAX_RAY, stream);
    read_scalars();
    x_cost = h_scalars[S_COST_X] + h_scalars[S_COSTPTS_X];
    grad_max = std::max(h_scalars[S_GMAX_CAM], std::max(h_scalars[S_GMAX_RAY], h_scalars[S_GMAX_B]));
  }

  void read_scalars() {
    const bool want_dscal = defl_enabled && have_W;
    if (!pub_spin) {
      d_scalars.download(h_scalars, S_COUNT, stream);
      d_pcg_info.download(h_info, 2, stream);
      d_fail.download(h_info + 2, 1, stream);
      if (want_dscal) d_dscal.download(h_dscal, 2, stream);
      PTZ_CUDA(cudaStreamSynchronize(stream));
      return;
    }
    ++pub_seq;
    k_publish<<<1, 64, 0, stream>>>(d_scalars.p, d_pcg_info.p, d_fail.p, want_dscal ? d_dscal.p : nullptr, h_pub, pub_seq);
    PTZ_CUDA(cudaGetLastError());
    volatile unsigned long long* flag = &h_pub->seq;
    for (unsigned spins = 1; *flag != pub_seq; ++spins) {
      if ((spins & 0xfffu) == 0) {  // every few tens of microseconds: has the stream died under us?
        const cudaError_t e = cudaStreamQuery(stream);
        if (e == cudaSuccess) {
          if (*flag != pub_seq) throw CudaError(PTZ_ERR_CUDA, "k_publish finished without raising its sequence number");
          break;
        }
        if (e != cudaErrorNotReady) PTZ_CUDA(e);
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (want_dscal) { h_dscal[0] = h_pub->dscal[0]; h_dscal[1] = h_pub->dscal[1]; }
  }

  // ---- stages 2-4: one trust-region step attempt at the current Jacobian.  Returns false on linear-solver failure.
  bool compute_step_and_candidate(int* lin_iters) {
    cudaStream_t s = stream;
    const double mu = radius;
    const int refresh = reuse_diagonal ? 0 : 1;
    const int own = (g_nccl.rank == 0) ? 1 : 0;
    // (after an accepted step the ray blocks were already damped and factored for this radius inside k_track_accum: their failure flag,
    // if any, is in d_fail and must survive)
    const bool lt_ready = P > 0 && refresh && lt_mu == mu;
    lt_mu = 0.0;
    if (!lt_ready) PTZ_CUDA(cudaMemsetAsync(d_fail.p, 0, sizeof(int), s));
    if (P > 0)
      PTZ_TIMED(PTZ_K_TRACK_SOLVE, {
        if (!lt_ready)
          k_track_factor<<<nblk_ray, 128, 0, s>>>(P, ds.t_off.p, d_Vh.p, mu, refresh, opt.min_lm_diagonal, opt.max_lm_diagonal, d_diag_ray.p, d_Lt.p, d_fail.p);
        if (ds.nchunks > 0)
          k_obs_what<NCL><<<ow_grid, kChunk, ObsWhatSmem<NCL>::kBytes, s>>>(ds.nchunks, ow_per, ds.chunk_view.p, ds.chunk_begin.p, ds.chunk_cnt.p, ds.o_track.p, d_recA.p, d_recF.p, d_Lt.p,
                                                                          d_What.p, d_q.p);
        if (kDisp) k_disp_track<NCL><<<nblk_ray, 128, 0, s>>>(P, ds.t_off.p, ds.t_obs.p, d_recA.p, d_recd.p, d_Lt.p, d_Wdh.p);
      });
    if (d_ub_list.n) PTZ_CUDA(cudaMemsetAsync(s_packed ? p_Spack : p_Sval, 0, (size_t)(s_packed ? V + ds.nub : ds.nnzb) * NCL * NCL * sizeof(double), s));  // blocks without local pairs stay zero
    // stage-2 targets: the block-CSR S, or (sharded) the packed buffer the all-reduce carries
    const int* t_diag = s_packed ? nullptr : ds.diag_pos.p;
    const int* t_ub = s_packed ? nullptr : ds.ub_pos.p;
    const int* t_ubt = s_packed ? nullptr : ds.ub_pos_t.p;
    double* t_Sd = s_packed ? p_Spack : p_Sval;
    double* t_So = s_packed ? p_Spack + (size_t)V * NCL * NCL : p_Sval;
    PTZ_TIMED(PTZ_K_SCHUR_DIAG, k_schur_diag<NCL><<<cdiv(V * (D::NU + NCL), 128), 128, 0, s>>>(V, ds.view_chunk_off.p, d_q.p, p_U, p_g, mu, refresh, opt.min_lm_diagonal,
                                                                               opt.max_lm_diagonal, own, d_diag_cam.p, t_diag, t_Sd, p_rhs, ns > 0 ? d_grp_of.p : nullptr));
    if (ds.nub > 0) {
      bool half = false;
      if constexpr (NCL == 4) {
        half = od_group != 32;
        const int* list = d_ub_list.n ? d_ub_list.p : nullptr;
        if (od_group == 16)
          PTZ_TIMED(PTZ_K_SCHUR_OFFDIAG, (k_schur_offdiag_sub<NCL, 16><<<cdiv(std::max(nub_local, 1), 16), 256, 0, s>>>(
                                             nub_local, list, ds.pair_off.p, ds.pair_a.p, ds.pair_b.p, d_What.p, t_ub, t_ubt, t_So)));
        else if (od_group == 8)
          PTZ_TIMED(PTZ_K_SCHUR_OFFDIAG, (k_schur_offdiag_sub<NCL, 8><<<cdiv(std::max(nub_local, 1), 32), 256, 0, s>>>(
                                             nub_local, list, ds.pair_off.p, ds.pair_a.p, ds.pair_b.p, d_What.p, t_ub, t_ubt, t_So)));
      }
      if (!half)
        PTZ_TIMED(PTZ_K_SCHUR_OFFDIAG, k_schur_offdiag<NCL><<<cdiv(std::max(nub_local, 1), 8), 256, 0, s>>>(nub_local, d_ub_list.n ? d_ub_list.p : nullptr, ds.pair_off.p,
                                                                                                            ds.pair_a.p, ds.pair_b.p, d_What.p, t_ub,
                                                                                                            t_ubt, t_So));
    }
    if (nb > 0)
      k_border_system<<<1, 128, 0, s>>>(nb, nf, p_Hbb, p_Hrf, p_Hff, p_gb, mu, refresh, opt.min_lm_diagonal, opt.max_lm_diagonal, own, d_diag_b.p, d_hinv.p,
                                        p_Sbb, p_rhs + (size_t)V * NCL, nb_plain);
    if (kDisp) {
      k_disp_schur_view<NCL><<<cdiv(V, 64), 64, 0, s>>>(V, ds.view_off.p, ds.o_track.p, d_What.p, d_Wdh.p, nb, bo_disp, own, p_C, p_Cw);
      k_disp_schur_border<<<1, 32, 0, s>>>(P, ds.t_off.p, d_Wdh.p, nb, bo_disp, p_Sbb, p_rhs + (size_t)V * NCL);
    }
    PTZ_CUDA(cudaGetLastError());
    if (g_nccl.world > 1) {
      PTZ_TIMED(PTZ_K_ALLREDUCE, allreduce_sum(d_sys.p, sys_n, s));
      k_unpack_S<NCL><<<cdiv((int)((size_t)(V + ds.nub) * NCL * NCL), 256), 256, 0, s>>>(V, ds.nub, p_Spack, ds.diag_pos.p, ds.ub_pos.p, ds.ub_pos_t.p, p_Sval);
      PTZ_CUDA(cudaGetLastError());
    }
    if (nf > 0) {  // fy elimination, camera side (replicated: after the reduce)
      k_fy_eliminate<NCL><<<cdiv(nf, 64), 64, 0, s>>>(nf, nb, d_ann_view.p, d_ann_strip.p, p_Cf, p_Hrf, p_gb + nb, d_hinv.p, kDisp ? p_Cw : p_C, p_Cw,
                                                     ds.diag_pos.p, p_Sval, p_rhs);
      PTZ_CUDA(cudaGetLastError());
    }
    if (ns > 0) {  // shared intrinsics: fold their columns of S into the border (replicated: after the reduce)
      k_shared_fold<NCL><<<cdiv(V, 64), 64, 0, s>>>(V, nb, nb_plain, d_grp_of.p, ds.s_rowptr.p, ds.s_col.p, p_Sval, p_rhs, kDisp ? p_Cw : nullptr, p_Cw,
                                                    d_rowpart.p);
      k_shared_border<NCL><<<cdiv(ns, 32), 32, 0, s>>>(ns, nb, nb_plain, d_grp_off.p, d_grp_view.p, d_rowpart.p, d_sh_h.p, mu, refresh, opt.min_lm_diagonal,
                                                       opt.max_lm_diagonal, d_diag_b.p, p_Sbb, p_rhs + (size_t)V * NCL);
      PTZ_CUDA(cudaGetLastError());
    }
    if (use_dense) {
      // small system: exact dense Cholesky in one CTA instead of the CG (k_dense_chol)
      PTZ_TIMED(PTZ_K_PCG, {
        PTZ_CUDA(cudaMemsetAsync(d_dense.p, 0, d_dense.n * sizeof(double), s));
        const int nthreads = ds.nnzb + ncpl * NCL * nb + nb * nb + n;
        k_dense_assemble<NCL><<<cdiv(nthreads, 128), 128, 0, s>>>(V, nb, ds.nnzb, ds.blk_row.p, ds.s_col.p, p_Sval, p_rhs, ncpl, d_cpl_view.p,
                                                                  (kDisp || nf > 0 || ns > 0) ? p_Cw : p_C, p_Sbb, d_dense.p);
        k_dense_chol<<<1, kDenseThreads, dense_smem_bytes(), s>>>(n, d_dense.p, d_y.p, d_pcg_info.p, d_fail.p);
        if (ns > 0) k_shared_expand<NCL><<<cdiv(V * (NCL - 3), 128), 128, 0, s>>>(V, d_grp_of.p, V * NCL, bo_sh, d_y.p);
      });
      PTZ_CUDA(cudaGetLastError());
      launch_stage4(mu);
      read_scalars();
      ++cost_evals;
      *lin_iters = h_info[0];
      last_lin_iters = h_info[0];
      return h_info[2] == 0;
    }
    PTZ_TIMED(PTZ_K_PRECOND, {
      k_precond<NCL><<<cdiv(V, 128), 128, 0, s>>>(V, ds.diag_pos.p, p_Sval, d_Linv.p, d_fail.p, defl_enabled ? d_Lfac.p : nullptr);
      if (nb > 0) k_precond_border<<<1, 32, 0, s>>>(nb, p_Sbb, d_Linv_b.p, d_fail.p);
      k_scale_system<NCL><<<cdiv(std::max(ds.nnzb, V), 128), 128, 0, s>>>(V, ds.nnzb, ds.blk_row.p, ds.s_col.p, d_Linv.p, p_Sval, p_rhs, arena_ptr(ar_st0),
                                                                            arena_ptr(ar_x), d_cgp.p, d_cg_owner.p,
                                                                            rows_sharded && g_nccl.world > 1 ? cgR : -1);
      if (nb > 0)
        k_scale_border<NCL><<<1, 128, 0, s>>>(V, nb, ncpl, d_cpl_view.p, d_Linv.p, d_Linv_b.p, (kDisp || nf > 0 || ns > 0) ? p_Cw : p_C, d_Cs.p, p_rhs, arena_ptr(ar_st0), arena_ptr(ar_x),
                                              d_cgp.p);
    });
    // ---- stages 3 and 4
    // Deflation pays when the plain iteration would sit on its plateau (hundreds of steps while the step is large); near
    // convergence a solve takes a handful of iterations and the set-up of the deflated one (~0.1 ms) costs more than it saves.
    // The previous solve's count decides: plain after a short deflated solve, deflated again after a long plain one.
    if (have_W) {
      if (defl_active && last_lin_iters >= 0 && last_lin_iters < kDeflOffBelow) defl_active = false;
      else if (!defl_active && last_lin_iters > kDeflOnAbove) defl_active = true;
    }
    const bool deflate = defl_enabled && have_W && defl_active;
    const bool record = defl_enabled && !have_W;  // undeflated solve: keep its residual history for the deflation basis
    linear_solve(deflate, record, false);
    launch_stage4(mu);
    read_scalars();
    if (deflate && h_info[1] == 4) {
      // the deflated solve levelled off just above the tolerance (k_cg, status 4): polish with the plain iteration from its x
      ++defl_polishes;
      const int it1 = h_info[0];
      linear_solve(false, false, false, true);
      launch_stage4(mu);
      read_scalars();
      if (opt.verbose || getenv("PTZ_DEFL_DEBUG"))
        fprintf(stderr, "[ptzba rank %d] LM iteration %d: deflated solve stagnated after %d iterations, polished with %d plain ones (status %d)\n", g_nccl.rank,
                iteration, it1, h_info[0], h_info[1]);
      h_info[0] += it1;
    }
    if (deflate) {
      ++defl_solves;
      const bool stalled = h_info[1] == 1 && h_info[0] < opt.pcg_max_iterations;  // hit the deflated solve's own cap
      if (stalled) h_info[1] = 2;
      if (h_dscal[1] == 0.0 || h_info[1] == 2) {
        // the (stale) basis lost rank, or the deflated single-reduction recurrences stagnated above the tolerance and broke down (seen
        // on sharded problems with V = 4000): drop the basis, redo this solve with the plain iteration and harvest a FRESH basis from
        // it; after three such rejections in one run the handle stays with the plain iteration
        ++defl_rejects;
        if (opt.verbose || getenv("PTZ_DEFL_DEBUG")) {
          double res = 0;
          d_pcg_res.download(&res, 1, stream);
          PTZ_CUDA(cudaStreamSynchronize(stream));
          fprintf(stderr, "[ptzba rank %d] deflated solve rejected at LM iteration %d: basis usable %g, pcg status %d%s after %d iterations, |r|/|b| = %.3e (%d rejections)\n",
                  g_nccl.rank, iteration, h_dscal[1], h_info[1], stalled ? " (stalled at its cap)" : "", h_info[0], res, defl_rejects);
        }
        have_W = false;
        if (defl_rejects >= 3) defl_enabled = false;
        if (defl_enabled && d_hist.n == 0) d_hist.alloc((size_t)kHistCap * V * NCL, stream);
        if (h_info[1] == 2) {
          const bool rec = defl_enabled;
          linear_solve(false, rec, true);
          launch_stage4(mu);
          read_scalars();
          if (rec && h_info[1] == 0 && h_info[0] >= kMinHarvest) harvest_basis(h_info[0]);
        }
      }
    } else if (record && h_info[1] == 0 && h_info[0] >= kMinHarvest) {
      harvest_basis(h_info[0]);
    }
    ++cost_evals;
    *lin_iters = h_info[0];
    last_lin_iters = h_info[0];
    if (h_info[2] != 0) return false;          // a 3x3 / camera / border block was not positive definite
    if (h_info[1] == 3) throw CudaError(PTZ_ERR_NCCL, "k_cg: a peer rank did not reach the cross-GPU barrier (timeout)");
    if (h_info[1] == 2) return false;          // PCG breakdown (non-finite or non-positive curvature)
    return true;
  }

  // stage 3: the deflation set-up (when a basis exists), ONE cooperative k_cg launch, back to the unscaled unknowns
  void linear_solve(bool deflate, bool record, bool restart, bool warm = false) {
    cudaStream_t s = stream;
    const int ncam = V * NCL;
    const size_t nk = (size_t)ncam * kDeflK;
    const int mr = (rows_sharded && g_nccl.world > 1) ? cgR : -1;
    CgArgs a;
    a.V = V; a.nb = nb; a.n = n;
    a.rowptr = ds.s_rowptr.p; a.col = d_cg_col.p; a.Sval = p_Sval; a.peer_mask = d_cg_mask.p;
    a.nav = ncpl; a.ann_view = d_cpl_view.p; a.ann_idx = d_cpl_idx.p; a.C = d_Cs.p; a.order = d_cg_order.p;
    for (int k = 0; k < kMaxPeers; ++k) a.arena[k] = g_arena.base[k];
    if (!rows_sharded && cg_vranks == 1) a.arena[0] = g_arena.base[cgR];  // the single-rank kernel works in arena[0]: this rank's own block
    a.off_partial = ar_partial; a.off_st0 = ar_st0; a.off_st1 = ar_st1; a.off_x = ar_x; a.off_ll0 = ar_ll0; a.off_ll1 = ar_ll1;
    a.W = cgW; a.rank = rows_sharded ? cgR : 0; a.vranks = cg_vranks; a.slots_per_rank = cg_slots_per_rank; a.p = d_cgp.p;
    // a deflated solve normally takes ~0.4x the iterations of the plain solve its basis came from (126 / 138 after 322 at V = 4000); one
    // that is not done after 0.6x has stagnated above the tolerance: stop it there (-> rejected, redone plain, fresh basis)
    a.max_iter = deflate ? std::min(opt.pcg_max_iterations, std::max(60, (defl_ref_iters * 3) / 5)) : opt.pcg_max_iterations;
    a.tol = opt.pcg_rel_tolerance;
    a.out_info = d_pcg_info.p; a.out_res = d_pcg_res.p;
    // shared-memory residency of S: every warp keeps up to cg_cap blocks (+ column indices) of its rows for the whole solve
    a.smem_blocks = cg_cap;
    a.smem_defl_rows = deflate ? cg_defl_rows : 0;
    { const char* e = getenv("PTZ_CG_DEBUG"); a.debug = e ? atoi(e) : 0; }
    a.dW = d_Wt.p; a.dAW = d_AW.p; a.dZ = d_Z.p; a.dEinv = d_Einv.p; a.dscal = d_dscal.p;
    a.hist = record ? d_hist.p : nullptr; a.hist_cap = kHistCap; a.abg = record ? d_abg.p : nullptr;
    a.prof = nullptr;
    a.gamma0_ptr = warm ? d_dscal.p : nullptr;
    if (a.debug & 8) {  // per-phase cycle counters of CTA 0 (printed by the destructor)
      if (d_prof.n == 0) { d_prof.alloc(16, s); d_prof.zero(s); }
      a.prof = d_prof.p + (deflate ? 8 : 0);
    }
    const size_t cg_smem = cg_smem_bytes(deflate);
    void* args[] = {&a};
    if (restart) k_cg_restart<NCL><<<cdiv(ncam, 256), 256, 0, s>>>(V, d_bcopy.p, arena_ptr(ar_st0), arena_ptr(ar_x), d_cgp.p);
    if (warm) {  // plain iteration from the x~ a stagnated deflated solve left behind: true residual into the state
      k_cg_residual<NCL><<<cdiv(V, 8), 256, 0, s>>>(V, ds.s_rowptr.p, ds.s_col.p, p_Sval, arena_ptr(ar_x), d_bcopy.p, arena_ptr(ar_st0), d_cgp.p, d_cg_owner.p, mr);
      allreduce_sum(arena_ptr(ar_st0), 3 * (size_t)ncam, s);  // sharded rows: every rank needs the whole initial state (no-op on one GPU)
    }
    if (deflate)
      PTZ_TIMED(PTZ_K_DEFLATE, {
        const int nchunk = cdiv(ncam, kDeflGramChunk);
        k_defl_scale_basis<NCL><<<cdiv(V * kDeflK, 256), 256, 0, s>>>(V, d_Lfac.p, d_Wy.p, d_Wt.p);  // W~ = L^T W_y
        k_defl_spmm<NCL><<<V, 128, 0, s>>>(V, ds.s_rowptr.p, ds.s_col.p, p_Sval, d_Wt.p, d_AW.p, d_cg_owner.p, mr);
        allreduce_sum(d_AW.p, nk, s);  // sharded rows: every rank needs all of AW (for E and for the start vector)
        k_defl_spmm<NCL><<<V, 128, 0, s>>>(V, ds.s_rowptr.p, ds.s_col.p, p_Sval, d_AW.p, d_Z.p, d_cg_owner.p, mr);
        k_defl_gram<NCL><<<nchunk, 256, 0, s>>>(V, d_Wt.p, d_AW.p, arena_ptr(ar_st0), d_gram.p, d_bcopy.p);
        k_defl_small<<<1, kDeflSmallThreads, 0, s>>>(defl_kd, nchunk, d_gram.p, d_Einv.p, d_c0.p, d_dscal.p);
        k_defl_start<NCL><<<cdiv(ncam, 256), 256, 0, s>>>(V, d_Wt.p, d_AW.p, d_c0.p, d_dscal.p, arena_ptr(ar_st0), arena_ptr(ar_x));
      });
    PTZ_TIMED(PTZ_K_PCG, {
      if (cg_vranks > 1) {  // debug: every virtual rank starts from the same initial state
        for (int k = 1; k < cg_vranks; ++k) {
          PTZ_CUDA(cudaMemcpyAsync(g_arena.base[k] + ar_st0, g_arena.base[0] + ar_st0, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
          PTZ_CUDA(cudaMemcpyAsync(g_arena.base[k] + ar_x, g_arena.base[0] + ar_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
        }
      }
      PTZ_CUDA(cudaLaunchCooperativeKernel(cg_kernel(deflate), dim3(cg_grid * cg_vranks), dim3(32 * cg_wpb), args, cg_smem, s));
      k_unscale<NCL><<<cdiv(V + nb, 128), 128, 0, s>>>(V, nb, d_Linv.p, d_Linv_b.p, arena_ptr(ar_x), d_y.p);
      if (ns > 0) k_shared_expand<NCL><<<cdiv(V * (NCL - 3), 128), 128, 0, s>>>(V, d_grp_of.p, V * NCL, bo_sh, d_y.p);
    });
    PTZ_CUDA(cudaGetLastError());
  }

  // after the first (undeflated, recorded) solve of a run: lowest Ritz vectors of its Lanczos tridiagonal -> deflation basis,
  // stored in the unscaled unknowns (W_y = Linv^T W~) so that it can follow the block-Jacobi scaling of the later solves
  void harvest_basis(int iterations) {
    cudaStream_t s = stream;
    const int m = std::min(iterations, (int)kHistCap), ncam = V * NCL;
    d_abg.download(h_abg.data(), 3 * (size_t)kHistCap, s);  // alpha, beta, gamma of the recorded solve
    PTZ_CUDA(cudaStreamSynchronize(s));
    std::vector<double> Y;
    const int kd = lowest_ritz_vectors(h_abg.data(), m, kDeflK, kDeflK, Y);
    if (kd < 4) return;
    d_Y.upload(Y.data(), (size_t)m * kDeflK, s);
    const int mr = (rows_sharded && g_nccl.world > 1) ? cgR : -1;
    k_defl_harvest<<<cdiv(ncam * kDeflK, 256), 256, 0, s>>>(ncam, m, kd, d_hist.p, d_Y.p, d_Wt.p, d_cg_owner.p, NCL, mr);
    allreduce_sum(d_Wt.p, (size_t)ncam * kDeflK, s);
    k_defl_scale_basis<NCL><<<cdiv(V * kDeflK, 256), 256, 0, s>>>(V, d_Linv.p, d_Wt.p, d_Wy.p);
    PTZ_CUDA(cudaGetLastError());
    have_W = true;
    defl_active = true;
    defl_kd = kd;
    defl_ref_iters = iterations;
    d_hist.release();  // (stream-ordered: goes back to the block cache behind the kernels above)
    if (opt.verbose) printf("[ptzba] deflation basis: %d Ritz vectors from %d CG iterations\n", kd, m);
  }

  void launch_stage4(double mu) {
    cudaStream_t s = stream;
    const int nxt = cur ^ 1;
    const double* y = d_y.p;
    if (P > 0 && track_groups > 0)
      PTZ_TIMED(PTZ_K_TRACK_BACKSUB, k_track_backsub_w<NCL><<<cdiv(track_groups, 8), 256, 0, s>>>(
                                         track_groups, M, d_grp_e0.p, d_t_trk.p, ds.t_obs.p, d_t_view.p, d_What.p, y, d_Lt.p, d_Vh.p, d_diag_ray.p, mu, d_trk[cur].p,
                                         d_trk[nxt].p, d_part3_ray.p, kDisp ? d_Wdh.p : nullptr, y + (size_t)V * NCL + (kDisp ? bo_disp : 0)));
    else if (P > 0)
      PTZ_TIMED(PTZ_K_TRACK_BACKSUB, k_track_backsub<NCL><<<nblk_ray, 128, 0, s>>>(P, ds.t_off.p, ds.t_obs.p, d_t_view.p, d_What.p, y, d_Lt.p, d_Vh.p, d_diag_ray.p,
                                                                                   mu, d_trk[cur].p, d_trk[nxt].p, d_part3_ray.p, kDisp ? d_Wdh.p : nullptr,
                                                                                   y + (size_t)V * NCL + (kDisp ? bo_disp : 0)));
    PTZ_TIMED(PTZ_K_CAM_UPDATE, k_cam_update<TYPE><<<nblk_cam, 128, 0, s>>>(V, y, d_scale_cam.p, p_g, d_diag_cam.p, mu, ds.view_active.p, d_intr[cur].p,
                                                                            d_ext[cur].p, d_intr[nxt].p, d_ext[nxt].p, d_part3_cam.p, ns > 0 ? d_intr_counted.p : nullptr));
    if (nb > 0)
      k_border_update<NCL><<<1, 32, 0, s>>>(nb, nf, bo_tlw, bo_disp, d_ann_view.p, y, V * NCL, d_scale_b.p, p_gb, d_diag_b.p, mu, p_Cf, p_Hrf, d_hinv.p, d_tlw[cur].p,
                                       d_tlw[nxt].p, d_intr[cur].p, d_intr[nxt].p, d_dispp[cur].p, d_dispp[nxt].p, d_part3_b.p, nb_plain);
    launch_cost(nxt);
    launch_step_scalars();
  }

  void launch_cost(int which) {
    cudaStream_t s = stream;
    PTZ_TIMED(PTZ_K_VIEW_PREP, k_view_prep<<<cdiv(V, 128), 128, 0, s>>>(V, d_intr[which].p, d_ext[which].p, d_vt.p, 0, d_scale_cam.p, NCL));
    vt_for = which;
    if (ds.nchunks > 0)
      PTZ_TIMED(PTZ_K_COST, k_cost<TYPE><<<ds.nchunks, kChunk, 0, s>>>(ds.chunk_view.p, ds.chunk_begin.p, ds.chunk_cnt.p, ds.o_uv.p, ds.o_track.p, d_vt.p,
                                                                         d_trk[which].p, d_dispp[which].p, d_cost_part.p));
    if (A > 0)
      k_pts_cost<<<1, 32, 0, s>>>(A, d_pts_uv.p, d_pts_xyz.p, d_pts_view.p, d_vt.p, d_tlw[which].p, kDisp ? d_dispp[which].p : nullptr,
                                  d_scalars.p + S_COSTPTS_CAND);
    PTZ_CUDA(cudaGetLastError());
  }

  void launch_step_scalars() {
    ScalarJobs J;
    J.nsum = 0; J.nmax = 0;
    auto add_sum = [&](const double* p, int cnt, int stride, int slot) { J.sum_ptr[J.nsum] = p; J.sum_n[J.nsum] = cnt; J.sum_stride[J.nsum] = stride; J.sum_slot[J.nsum] = slot; ++J.nsum; };
    add_sum(d_cost_part.p, ds.nchunks, 2, S_COST_CAND);
    add_sum(d_cost_part.p + 1, ds.nchunks, 2, S_RAW2_CAND);
    add_sum(d_part3_ray.p, P > 0 ? nray_parts : 0, 3, S_DM_RAY);
    add_sum(d_part3_ray.p + 1, P > 0 ? nray_parts : 0, 3, S_STEP2_RAY);
    add_sum(d_part3_ray.p + 2, P > 0 ? nray_parts : 0, 3, S_XN2_RAY);
    add_sum(d_part3_cam.p, nblk_cam, 3, S_DM_CAM);
    add_sum(d_part3_cam.p + 1, nblk_cam, 3, S_STEP2_CAM);
    add_sum(d_part3_cam.p + 2, nblk_cam, 3, S_XN2_CAM);
    add_sum(d_part3_b.p, 1, 3, S_DM_B);
    add_sum(d_part3_b.p + 1, 1, 3, S_STEP2_B);
    add_sum(d_part3_b.p + 2, 1, 3, S_XN2_B);
    PTZ_TIMED(PTZ_K_SCALARS, k_scalars<<<J.nsum + J.nmax, 1024, 0, stream>>>(J, d_scalars.p));
    PTZ_CUDA(cudaGetLastError());
    allreduce_sum(d_scalars.p, S_SUM_END, stream);
  }

  // |x| of the current point over the coordinates that are in the Ceres problem
  double current_x_norm() {
    const int nblk = std::max(cdiv(std::max(V, P), 256), 1);
    DevBuf<double> part;
    part.alloc(2 * (size_t)nblk, stream);
    k_xnorm2<<<nblk, 256, 0, stream>>>(V, P, ds.view_active.p, d_intr[cur].p, d_ext[cur].p, ds.t_off.p, d_trk[cur].p, part.p, ns > 0 ? d_intr_counted.p : nullptr);
    ScalarJobs J;
    J.nsum = 2; J.nmax = 0;
    J.sum_ptr[0] = part.p; J.sum_n[0] = nblk; J.sum_stride[0] = 2; J.sum_slot[0] = S_XN2_CAM;
    J.sum_ptr[1] = part.p + 1; J.sum_n[1] = nblk; J.sum_stride[1] = 2; J.sum_slot[1] = S_XN2_RAY;
    k_scalars<<<2, 1024, 0, stream>>>(J, d_scalars.p);
    PTZ_CUDA(cudaGetLastError());
    allreduce_sum(d_scalars.p + S_XN2_RAY, 1, stream);
    double tl[6] = {0, 0, 0, 0, 0, 0}, dd[3] = {0, 0, 0};
    if (A > 0) d_tlw[cur].download(tl, 6, stream);
    if (kDisp) d_dispp[cur].download(dd, 3, stream);
    read_scalars();
    double s = h_scalars[S_XN2_CAM] + h_scalars[S_XN2_RAY];
    if (A > 0) for (int j = 0; j < 6; ++j) s += tl[j] * tl[j];
    for (int j = 0; j < 3; ++j) s += dd[j] * dd[j];
    return sqrt(s);
  }

  void push_log(double cost, double cost_change, double step_norm, double rho, int lin, int ok) {
    ptz_iter_log l;
    l.cost = cost; l.cost_change = cost_change; l.gradient_max_norm = grad_max; l.step_norm = step_norm; l.relative_decrease = rho;
    l.trust_region_radius = radius; l.linear_solver_iterations = lin; l.step_is_successful = ok;
    log.push_back(l);
    if (opt.verbose)
      printf("[ptzba] it %3d cost %.10e change %.3e |g| %.3e |step| %.3e rho %.3e radius %.3e pcg %d %s\n", (int)log.size() - 1, cost, cost_change, grad_max,
             step_norm, rho, radius, lin, ok == 1 ? "ok" : ok == 0 ? "rej" : "invalid");
  }

  void run(int max_new_iterations, ptzba_result* out) override {
    auto t0 = std::chrono::steady_clock::now();
    clk.begin(-1);
    const size_t run_slot = clk.used;  // its end event is recorded below
    clk.used += 1;
    if (!started) {
      // IterationZero
      x_norm = current_x_norm();
      evaluate_jacobian(true);
      initial_cost = x_cost; min_cost = x_cost;
      last_successful = true;
      ++num_successful;
      push_log(x_cost, 0, 0, 0, 0, 1);
      started = true;
    }
    const int iter_cap = iteration + max_new_iterations;
    while (!finished) {
      // FinalizeIterationAndCheckIfMinimizerCanContinue
      if (iteration >= opt.max_num_iterations) { termination = PTZ_NO_CONVERGENCE; finished = true; break; }
      if (last_successful && grad_max <= opt.gradient_tolerance) { termination = PTZ_CONVERGENCE; finished = true; break; }
      if (radius <= opt.min_trust_region_radius) { termination = PTZ_CONVERGENCE; finished = true; break; }
      if (iteration >= iter_cap) break;  // caller's slice is used up; not a termination
      ++iteration;
      int lin = 0;
      bool ok = compute_step_and_candidate(&lin);
      reuse_diagonal = true;
      lin_iters_total += lin;
      life_pcg += lin; ++life_lm;
      const double* S = h_scalars;
      double model_cost_change = S[S_DM_RAY] + S[S_DM_CAM] + S[S_DM_B];
      if (ok && !(std::isfinite(model_cost_change) && std::isfinite(S[S_STEP2_RAY]) && std::isfinite(S[S_STEP2_CAM]))) ok = false;
      if (ok) ok = model_cost_change > 0.0;
      if (!ok) {
        // HandleInvalidStep
        ++num_consecutive_invalid;
        if (num_consecutive_invalid >= opt.max_num_consecutive_invalid_steps) { termination = PTZ_FAILURE; finished = true; break; }
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        last_successful = false;
        ++num_unsuccessful;
        push_log(x_cost, 0, 0, 0, lin, -1);
        continue;
      }
      num_consecutive_invalid = 0;
      double cand_cost = S[S_COST_CAND] + S[S_COSTPTS_CAND];
      if (!std::isfinite(cand_cost)) cand_cost = 1.7976931348623157e308;
      const double step_norm = sqrt(S[S_STEP2_RAY] + S[S_STEP2_CAM] + S[S_STEP2_B]);
      const double cand_norm = sqrt(S[S_XN2_RAY] + S[S_XN2_CAM] + S[S_XN2_B]);
      // ParameterToleranceReached
      if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { termination = PTZ_CONVERGENCE; finished = true; break; }
      // FunctionToleranceReached
      const double cost_change = x_cost - cand_cost;
      if (fabs(cost_change) <= opt.function_tolerance * x_cost) { termination = PTZ_CONVERGENCE; finished = true; break; }
      const double rho = cost_change / model_cost_change;
      if (rho > opt.min_relative_decrease) {
        // HandleSuccessfulStep
        cur ^= 1;
        x_norm = cand_norm;
        radius = radius / std::max(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3));  // (does not depend on the new Jacobian: known before it)
        radius = std::min(opt.max_trust_region_radius, radius);
        if (num_consecutive_invalid == 0) { PTZ_CUDA(cudaMemsetAsync(d_fail.p, 0, sizeof(int), stream)); fuse_factor_mu = radius; }
        evaluate_jacobian(false);
        decrease_factor = 2.0;
        reuse_diagonal = false;
        last_successful = true;
        ++num_successful;
        if (x_cost < min_cost) min_cost = x_cost;
        push_log(x_cost, cost_change, step_norm, rho, lin, 1);
      } else {
        // HandleUnsuccessfulStep
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        last_successful = false;
        ++num_unsuccessful;
        push_log(cand_cost, cost_change, step_norm, rho, lin, 0);
      }
    }
    cudaEventRecord(clk.ev[run_slot], stream);
    PTZ_CUDA(cudaStreamSynchronize(stream));
    clk.collect();
    double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out) fill_result(out, seconds);
  }

  void fill_result(ptzba_result* out, double seconds_solve) {
    out->termination = termination;
    out->num_iterations = (int)log.size() - 1;
    out->num_successful_steps = num_successful;
    out->num_unsuccessful_steps = num_unsuccessful;
    out->num_residuals = 2 * (M + A);
    if (g_nccl.world > 1) {
      // every rank reports the size of the whole problem
      double m = (double)M;
      DevBuf<double> d;
      d.upload(&m, 1, stream);
      allreduce_sum(d.p, 1, stream);
      d.download(&m, 1, stream);
      PTZ_CUDA(cudaStreamSynchronize(stream));
      out->num_residuals = 2 * ((int)llround(m) + A);
    }
    out->linear_solver_iterations = lin_iters_total;
    out->initial_cost = initial_cost;
    out->final_cost = min_cost;
    out->init_reproj_error_all = sqrt(2.0) * sqrt((2 * initial_cost) / out->num_residuals);
    out->final_reproj_error_all = sqrt(2.0) * sqrt((2 * min_cost) / out->num_residuals);
    // CalReprojError2d2d / 2d3d at the final parameters (unweighted)
    launch_cost(cur);
    ScalarJobs J;
    J.nsum = 1; J.nmax = 0;
    J.sum_ptr[0] = d_cost_part.p + 1; J.sum_n[0] = ds.nchunks; J.sum_stride[0] = 2; J.sum_slot[0] = S_RAW2_CAND;
    k_scalars<<<J.nsum + J.nmax, 1024, 0, stream>>>(J, d_scalars.p);
    allreduce_sum(d_scalars.p + S_RAW2_CAND, 1, stream);
    read_scalars();
    const double n2 = (double)(out->num_residuals / 2 - A);
    out->final_reproj_error_2d2d = sqrt(h_scalars[S_RAW2_CAND] / n2);
    out->final_reproj_error_2d3d = A > 0 ? sqrt(h_scalars[S_RAWPTS_CAND] / A) : sqrt(0.0 / 0.0);
    // parameters
    std::vector<double> intr(9 * (size_t)V), ext(6 * (size_t)V), tlw(6, 0.0), dsp(3, 0.0);
    d_dispp[cur].download(dsp.data(), 3, stream);
    d_intr[cur].download(intr.data(), intr.size(), stream);
    d_ext[cur].download(ext.data(), ext.size(), stream);
    d_tlw[cur].download(tlw.data(), 6, stream);
    if ((out->ray || out->rays_world) && P > 0) {
      // compact rays (local and world frame) on the device, straight into the caller's buffers
      DevBuf<double> d_ray, d_rayw;
      d_ray.alloc(3 * (size_t)P, stream); d_rayw.alloc(3 * (size_t)P, stream);
      k_rays_out<<<cdiv(P, 256), 256, 0, stream>>>(P, d_trk[cur].p, d_tlw[cur].p, d_ray.p, d_rayw.p);
      if (out->ray) d_ray.download(out->ray, 3 * (size_t)P, stream);
      if (out->rays_world) d_rayw.download(out->rays_world, 3 * (size_t)P, stream);
      PTZ_CUDA(cudaStreamSynchronize(stream));
    }
    PTZ_CUDA(cudaStreamSynchronize(stream));
    if (out->intr) memcpy(out->intr, intr.data(), intr.size() * 8);
    if (out->ext) memcpy(out->ext, ext.data(), ext.size() * 8);
    if (out->disp) memcpy(out->disp, dsp.data(), 24);
    if (out->tlw) memcpy(out->tlw, tlw.data(), 48);
    // ObtainRefinedCameraParams (ptzray_optimizer.cc:672-766)
    double Rlw[9];
    rodrigues_jac(tlw.data(), Rlw, nullptr);
    if (out->cams_world)
      for (int i = 0; i < V; ++i) {
        const double* in = &intr[9 * (size_t)i];
        const double* ex = &ext[6 * (size_t)i];
        double* c = out->cams_world + 21 * (size_t)i;
        c[0] = in[0]; c[1] = (TYPE == BA_PTZRAY_FXFY_DIST) ? in[1] : in[0]; c[2] = in[2]; c[3] = in[3];
        double R[9];
        rodrigues_jac(ex, R, nullptr);
        for (int r = 0; r < 3; ++r) {
          // t.z carries the displacement polynomial (ptzray_optimizer.cc:693-694,714-715)
          const double tz = (r == 2) ? (dsp[0] + dsp[1] * in[0] + dsp[2] * in[0] * in[0]) : 0.0;
          c[13 + r] = R[3 * r] * tlw[3] + R[3 * r + 1] * tlw[4] + R[3 * r + 2] * tlw[5] + ex[3 + r] + tz;
          for (int q = 0; q < 3; ++q) c[4 + 3 * r + q] = R[3 * r] * Rlw[q] + R[3 * r + 1] * Rlw[3 + q] + R[3 * r + 2] * Rlw[6 + q];
        }
        for (int j = 0; j < 5; ++j) c[16 + j] = in[4 + j];
      }
    out->log_count = 0;
    if (out->log)
      for (size_t i = 0; i < log.size() && (int)i < out->log_capacity; ++i) out->log[out->log_count++] = log[i];
    out->seconds_setup = seconds_setup;
    out->seconds_solve = seconds_solve;
  }

  // ptzba_eval: raw residuals + analytic Jacobian in the caller's observation order, weighted cost and gradient
  void eval(const double* disp, ptzba_eval_out* out) override {
    if (ns > 0) throw CudaError(PTZ_ERR_UNSUPPORTED, "ptzba_eval reports per-view columns: not defined with shared intrinsics");
    const int ncv = (TYPE == BA_PTZRAY) ? 5 : 6;
    const int dd = kDisp ? 3 : 0;
    const int wo = ncv + 3 + dd, wp = ncv + 6 + dd;
    fill_ones(d_scale_cam.p, (size_t)V * NCL);
    fill_ones(d_scale_b.p, std::max(nbt, 1));
    if (kDisp && disp) { PTZ_CUDA(cudaMemcpyAsync(d_dispp[cur].p, disp, 24, cudaMemcpyHostToDevice, stream)); PTZ_CUDA(cudaStreamSynchronize(stream)); }
    // pass 1: unweighted, unscaled records
    launch_resjac(0);
    std::vector<double> rec((size_t)std::max(M, 1) * D::RS), raw((size_t)std::max(A, 1) * 32), recd((size_t)std::max(M, 1) * 6, 0.0);
    {  // records back into one [r, E | F] row per observation
      std::vector<double> ra((size_t)std::max(M, 1) * D::RA), rf((size_t)std::max(M, 1) * D::RF);
      d_recA.download(ra.data(), (size_t)M * D::RA, stream);
      d_recF.download(rf.data(), (size_t)M * D::RF, stream);
      PTZ_CUDA(cudaStreamSynchronize(stream));
      for (int i = 0; i < M; ++i) {
        for (int j = 0; j < D::RA; ++j) rec[(size_t)i * D::RS + j] = ra[(size_t)i * D::RA + j];
        for (int j = 0; j < D::RF; ++j) rec[(size_t)i * D::RS + D::RA + j] = rf[(size_t)i * D::RF + j];
      }
    }
    if (kDisp && M > 0) d_recd.download(recd.data(), (size_t)M * 6, stream);
    if (A > 0) d_pts_raw.download(raw.data(), (size_t)A * 32, stream);
    PTZ_CUDA(cudaStreamSynchronize(stream));
    // live column a -> column of the documented layout [fx, fy, (k1), w(3)]
    int live2col[NCL];
    for (int a = 0; a < NCL; ++a) {
      if (TYPE == BA_PTZRAY) live2col[a] = a == 0 ? 0 : 1 + a;              // fx, w -> 0, 2,3,4
      else if (TYPE == BA_PTZRAY_FXFY_DIST) live2col[a] = a;                // fx, fy, k1, w
      else live2col[a] = a == 0 ? 0 : 1 + a;                                // fx, k1, w -> 0, 2, 3,4,5
    }
    h_perm.resize(std::max(M, 1));
    if (M > 0) { ds.perm.download(h_perm.data(), M, stream); PTZ_CUDA(cudaStreamSynchronize(stream)); }
    for (int i = 0; i < M; ++i) {
      const double* r = &rec[(size_t)i * D::RS];
      const int k = h_perm[i];
      if (out->residuals) { out->residuals[2 * (size_t)k] = r[0]; out->residuals[2 * (size_t)k + 1] = r[1]; }
      if (out->jac_obs)
        for (int row = 0; row < 2; ++row) {
          double* o = out->jac_obs + ((size_t)k * 2 + row) * wo;
          for (int c = 0; c < wo; ++c) o[c] = 0.0;
          for (int a = 0; a < NCL; ++a) o[live2col[a]] = r[8 + row * NCL + a];
          for (int j = 0; j < 3; ++j) o[ncv + j] = r[2 + row * 3 + j];
          for (int j = 0; j < dd; ++j) o[ncv + 3 + j] = recd[(size_t)i * 6 + row * 3 + j];
        }
    }
    for (int i = 0; i < A; ++i) {
      const double* r = &raw[(size_t)i * 32];
      const int k = h_pt_perm[i];
      if (out->residuals) { out->residuals[2 * (size_t)(M + k)] = r[0]; out->residuals[2 * (size_t)(M + k) + 1] = r[1]; }
      if (out->jac_pts)
        for (int row = 0; row < 2; ++row) {
          double* o = out->jac_pts + ((size_t)k * 2 + row) * wp;
          const double* jc = r + 2 + 6 * row;
          const double* jt = r + 14 + 6 * row;
          o[0] = jc[0]; o[1] = jc[1];
          if (ncv == 6) o[2] = jc[2];
          for (int j = 0; j < 3; ++j) o[ncv - 3 + j] = jc[3 + j];
          for (int j = 0; j < 6; ++j) o[ncv + j] = jt[j];
          for (int j = 0; j < dd; ++j) o[ncv + 6 + j] = r[26 + 3 * row + j];
        }
    }
    // pass 2: weighted (still unscaled) -> cost and gradient
    launch_resjac(1);
    std::vector<double> g((size_t)V * NCL), Vh((size_t)std::max(P, 1) * 10), gb(std::max(nbt, 1)), cv(V), cp(2, 0.0);
    PTZ_CUDA(cudaMemcpyAsync(g.data(), p_g, g.size() * 8, cudaMemcpyDeviceToHost, stream));
    PTZ_CUDA(cudaMemcpyAsync(cv.data(), p_cost_view, cv.size() * 8, cudaMemcpyDeviceToHost, stream));
    if (P > 0) d_Vh.download(Vh.data(), (size_t)P * 10, stream);
    if (nbt > 0) PTZ_CUDA(cudaMemcpyAsync(gb.data(), p_gb, nbt * 8, cudaMemcpyDeviceToHost, stream));
    if (A > 0) PTZ_CUDA(cudaMemcpyAsync(cp.data(), p_cost_pts, 16, cudaMemcpyDeviceToHost, stream));
    PTZ_CUDA(cudaStreamSynchronize(stream));
    double cost = cp[0];
    for (int v = 0; v < V; ++v) cost += cv[v];
    out->cost = cost;
    out->num_tangent = V * ncv + 3 * P + dd + (A > 0 ? 6 : 0);
    if (out->gradient) {
      for (int i = 0; i < out->num_tangent; ++i) out->gradient[i] = 0.0;
      for (int v = 0; v < V; ++v)
        for (int a = 0; a < NCL; ++a) out->gradient[(size_t)v * ncv + live2col[a]] = g[(size_t)v * NCL + a];
      for (int p = 0; p < P; ++p)
        for (int j = 0; j < 3; ++j) out->gradient[(size_t)V * ncv + 3 * (size_t)p + j] = Vh[(size_t)p * 10 + 6 + j];
      for (int j = 0; j < dd; ++j) out->gradient[(size_t)V * ncv + 3 * (size_t)P + j] = gb[bo_disp + j];
      if (A > 0) {
        for (int j = 0; j < 6; ++j) out->gradient[(size_t)V * ncv + 3 * (size_t)P + dd + j] = gb[bo_tlw + j];
        for (int k = 0; k < nf; ++k) out->gradient[(size_t)h_ann_view[k] * ncv + 1] = gb[nb + k];
      }
    }
  }

  void set_stage_timing(bool on) override { clk.per_kernel = on; }
  void stage_times(ptzba_stage_times* t) override {
    for (int i = 0; i < PTZ_K_COUNT; ++i) { t->ms_kernel[i] = clk.ms[i]; t->launches[i] = clk.launches[i]; }
    t->ms_run = clk.ms_run;
    t->lm_iterations = (int)life_lm; t->pcg_iterations = (int)life_pcg; t->jacobian_evals = jac_evals; t->cost_evals = cost_evals;
    t->nnz_blocks = ds.nnzb; t->deflated_solves = defl_solves; t->deflation_vectors = have_W ? defl_kd : 0; t->reserved_ = 0; t->num_pairs = ds.npairs;
  }
};

static int check_problem(const ptzba_problem* p) {
  if (!p) return PTZ_ERR_INVALID;
  if (p->num_views <= 0 || p->num_tracks < 0 || p->num_obs < 0 || p->num_pts3d < 0) return PTZ_ERR_INVALID;  // CheckValid (:515-535)
  if (!p->intr || !p->ext) return PTZ_ERR_INVALID;
  if (p->num_obs > 0 && (!p->obs_uv || !p->obs_view || !p->obs_track || !p->track_weight)) return PTZ_ERR_INVALID;
  if (p->num_pts3d > 0 && (!p->pt_uv || !p->pt_xyz || !p->pt_view)) return PTZ_ERR_INVALID;
  if (p->factor_type < 0 || p->factor_type > 3) return PTZ_ERR_INVALID;
  for (int k = 0; k < p->num_obs; ++k)
    if (p->obs_view[k] < 0 || p->obs_view[k] >= p->num_views || p->obs_track[k] < 0 || p->obs_track[k] >= p->num_tracks) return PTZ_ERR_INVALID;
  for (int k = 0; k < p->num_pts3d; ++k)
    if (p->pt_view[k] < 0 || p->pt_view[k] >= p->num_views) return PTZ_ERR_INVALID;
  if (p->shared_ic_id && p->num_pts3d > 0) {  // shared intrinsics blocks are built for the 2d-2d terms only (ba_border.cuh)
    for (int i = 0; i < p->num_views; ++i)
      for (int j = 0; j < i; ++j)
        if (p->shared_ic_id[i] == p->shared_ic_id[j]) { set_last_error("shared intrinsics together with 2d-3d terms are not built"); return PTZ_ERR_UNSUPPORTED; }
  }
  return PTZ_OK;
}

static BaSolverBase* make_solver(const ptzba_problem* p, const ptz_solver_options* o) {
  switch (p->factor_type) {
    case PTZ_BA_PTZRAY: return new BaSolver<BA_PTZRAY>(p, o);
    case PTZ_BA_PTZRAY_DIST: return new BaSolver<BA_PTZRAY_DIST>(p, o);
    case PTZ_BA_PTZRAY_FXFY_DIST: return new BaSolver<BA_PTZRAY_FXFY_DIST>(p, o);
    default: return new BaSolver<BA_PTZRAY_DIST_DISP>(p, o);
  }
}

static int ensure_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) { set_last_error("no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e)); return PTZ_ERR_NO_DEVICE; }
  return PTZ_OK;
}

template <class F>
static int guarded(F&& f) {
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

struct ptzba_handle {
  std::unique_ptr<BaSolverBase> s;
};

extern "C" {

void ptz_solver_options_default(ptz_solver_options* o) {
  o->max_num_iterations = 50;
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
  o->linear_solver = 1;
  o->num_threads = 1;
  o->verbose = 0;
}

const char* ptz_last_error(void) { return g_err; }

int ptz_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int ptzba_create(const ptzba_problem* prob, const ptz_solver_options* opt, ptzba_handle** h) {
  if (!h || !opt) return PTZ_ERR_INVALID;
  *h = nullptr;
  int rc = check_problem(prob);
  if (rc != PTZ_OK) return rc;
  if (opt->max_num_iterations <= 0) return PTZ_ERR_INVALID;
  rc = ensure_device();
  if (rc != PTZ_OK) return rc;
  return guarded([&]() {
    ptzba_handle* hh = new ptzba_handle();
    try {
      hh->s.reset(make_solver(prob, opt));
    } catch (...) {
      delete hh;
      throw;
    }
    *h = hh;
    return (int)PTZ_OK;
  });
}

int ptzba_reset(ptzba_handle* h) {
  if (!h) return PTZ_ERR_INVALID;
  return guarded([&]() { h->s->reset(); return (int)PTZ_OK; });
}

int ptzba_run(ptzba_handle* h, int max_new_iterations, ptzba_result* out) {
  if (!h) return PTZ_ERR_INVALID;
  return guarded([&]() { h->s->run(max_new_iterations, out); return (int)PTZ_OK; });
}

int ptzba_get_stage_times(ptzba_handle* h, ptzba_stage_times* t) {
  if (!h || !t) return PTZ_ERR_INVALID;
  h->s->stage_times(t);
  return PTZ_OK;
}

int ptzba_set_stage_timing(ptzba_handle* h, int per_kernel) {
  if (!h || !h->s) return PTZ_ERR_INVALID;
  h->s->set_stage_timing(per_kernel != 0);
  return PTZ_OK;
}

int ptzba_destroy(ptzba_handle* h) {
  delete h;
  return PTZ_OK;
}

int ptzba_solve(const ptzba_problem* prob, const ptz_solver_options* opt, ptzba_result* out) {
  if (!out) return PTZ_ERR_INVALID;
  ptzba_handle* h = nullptr;
  const bool timing = getenv("PTZ_TIMING") != nullptr;  // phase wall-clock on stderr (diagnostics)
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  int rc = ptzba_create(prob, opt, &h);
  if (rc != PTZ_OK) return rc;
  if (!getenv("PTZ_STAGE_TIMES")) h->s->set_stage_timing(false);  // nobody can read the per-kernel times of a one-shot solve
  const auto t1 = now();
  rc = ptzba_run(h, opt->max_num_iterations + 1, out);
  const auto t2 = now();
  ptzba_destroy(h);
  if (timing) fprintf(stderr, "[ptzba_solve rank %d] create %.2f ms  run %.2f ms  destroy %.2f ms\n", g_nccl.rank, ms(t0, t1), ms(t1, t2), ms(t2, now()));
  return rc;
}

int ptzba_eval(const ptzba_problem* prob, const double* disp, ptzba_eval_out* out) {
  if (!out) return PTZ_ERR_INVALID;
  ptz_solver_options o;
  ptz_solver_options_default(&o);
  ptzba_handle* h = nullptr;
  int rc = ptzba_create(prob, &o, &h);
  if (rc != PTZ_OK) return rc;
  rc = guarded([&]() { h->s->eval(disp, out); return (int)PTZ_OK; });
  ptzba_destroy(h);
  return rc;
}

int ptz_nccl_unique_id(void* id_bytes128) {
  if (!id_bytes128) return PTZ_ERR_INVALID;
  return guarded([&]() {
    ncclUniqueId id;
    PTZ_NCCL(ncclGetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id_bytes128, &id, 128);
    return (int)PTZ_OK;
  });
}

int ptz_nccl_init(const void* id_bytes128, int rank, int world_size) {
  if (!id_bytes128 || rank < 0 || world_size <= 0 || rank >= world_size) return PTZ_ERR_INVALID;
  return guarded([&]() {
    ncclUniqueId id;
    memcpy(&id, id_bytes128, 128);
    arena_release();
    if (g_nccl.comm) { ncclCommDestroy(g_nccl.comm); g_nccl.comm = nullptr; }
    PTZ_NCCL(ncclCommInitRank(&g_nccl.comm, world_size, id, rank));
    g_nccl.rank = rank;
    g_nccl.world = world_size;
    return (int)PTZ_OK;
  });
}

int ptz_nccl_finalize(void) {
  arena_release();
  if (g_nccl.comm) { ncclCommDestroy(g_nccl.comm); g_nccl.comm = nullptr; }
  g_nccl.rank = 0;
  g_nccl.world = 1;
  return PTZ_OK;
}

}  // extern "C"
