/*
 * ptzcalib_b200.h — C ABI of the B200-native PTZ-Calib nonlinear-least-squares hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI layer: its boundary is the
 * two C++ classes `ptzcalib::PTZRayOptimizer` (src/core/ptzray_optimizer.h:112-177) and
 * `ptzcalib::KRTOptimizer` (src/core/krt_optimizer.h:108-145), which hand a `ceres::Problem` to
 * `ceres::Solve` (ptzray_optimizer.cc:469-475, krt_optimizer.cc:387-394).  Every entry point below
 * replaces one of those calls with flat arrays; `include/ptzcalib_b200.hpp` rebuilds the two classes
 * on top of it.  Plain pointers and sizes only; no torch / CUDA types in any signature.
 *
 * The CPU oracle (`oracle/ptz_oracle.cpp`, test infrastructure only) exports the same functions with
 * the prefix `orc_` instead of `ptz`, on the same structs, so parity tests read one layout.
 *
 * Conventions
 *   - all arithmetic is fp64; pixel observations are fp32 (cv::Point2f, types.h:20,37) promoted to fp64
 *   - intr[9]  = fx, fy, cx, cy, d0..d4      (ptzray_optimizer.cc:647;  d = Camera::dist, types.cc:50-54)
 *   - ext[6]   = rvec(3), t(3)               (ptzray_optimizer.cc:651)
 *   - cam15    = fx, fy, cx, cy, rvec(3), t(3), d0..d4                  (Camera::ToVector, types.cc:32-57)
 *   - krt21    = fx, fy, cx, cy, R(9 row-major, camera<-world), t(3), d0..d4   (K,R,t,dist of Camera)
 *   - every function returns PTZ_OK (0) or a negative PTZ_ERR_* code; nothing throws across the ABI.
 */
#ifndef PTZCALIB_B200_H
#define PTZCALIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum FACTOR_TYPE, ptzray_optimizer.h:110 */
enum { PTZ_BA_PTZRAY = 0, PTZ_BA_PTZRAY_DIST = 1, PTZ_BA_PTZRAY_FXFY_DIST = 2, PTZ_BA_PTZRAY_DIST_DISP = 3 };
/* enum KRTOptimizer::FACTOR_TYPE, krt_optimizer.h:110 */
enum { PTZ_KRT_F = 0, PTZ_KRT_FDIST = 1, PTZ_KRT_FXFY = 2, PTZ_KRT_FXFYDIST = 3 };
/* ceres::TerminationType values the wrappers branch on (ptzray_optimizer.cc:482, krt_optimizer.cc:513) */
enum { PTZ_CONVERGENCE = 0, PTZ_NO_CONVERGENCE = 1, PTZ_FAILURE = 2 };

enum {
  PTZ_OK = 0,
  PTZ_ERR_INVALID = -1,     /* CheckValid()==false (ptzray_optimizer.cc:515-535) or a null/negative argument */
  PTZ_ERR_UNSUPPORTED = -2, /* a reference feature this build does not cover */
  PTZ_ERR_CUDA = -3,        /* CUDA runtime error; ptz_last_error() has the text */
  PTZ_ERR_NCCL = -4,
  PTZ_ERR_NO_DEVICE = -5    /* no CUDA device: there is NO CPU fallback */
};

/* Ceres 1.14.0 Solver::Options defaults that the reference leaves untouched, plus the
 * knobs of the GPU linear solver.  max_num_iterations is the only one the reference sets per call. */
typedef struct ptz_solver_options {
  int max_num_iterations;              /* ptzray_optimizer.cc:470, krt_optimizer.cc:388 */
  double function_tolerance;           /* 1e-6  */
  double gradient_tolerance;           /* 1e-10 */
  double parameter_tolerance;          /* 1e-8  */
  double initial_trust_region_radius;  /* 1e4   */
  double max_trust_region_radius;      /* 1e16  */
  double min_trust_region_radius;      /* 1e-32 */
  double min_relative_decrease;        /* 1e-3  */
  double min_lm_diagonal;              /* 1e-6  */
  double max_lm_diagonal;              /* 1e32  */
  int max_num_consecutive_invalid_steps; /* 5 */
  int jacobi_scaling;                  /* 1 */
  /* reduced-camera-system solver: block-Jacobi PCG, run until |r|<=tol*|b| or the iteration cap */
  int pcg_max_iterations;              /* 2000 */
  double pcg_rel_tolerance;            /* 1e-13: "exact" in the sense of SPARSE_SCHUR */
  /* oracle only: 0 = exact (forward-mode) Jacobian, 1 = Ceres CENTRAL numeric differentiation (A13) */
  int jacobian_mode;
  /* oracle only: 0 = dense Cholesky / QR, 1 = block-Jacobi PCG on the reduced system */
  int linear_solver;
  int num_threads;                     /* oracle only (OpenMP); the reference asks for 32 */
  int verbose;
} ptz_solver_options;

/* one row of Ceres' per-iteration table (minimizer_progress_to_stdout, ptzray_optimizer.cc:472) */
typedef struct ptz_iter_log {
  double cost;            /* cost of the iterate kept after this iteration (candidate cost if rejected) */
  double cost_change;
  double gradient_max_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  int linear_solver_iterations;
  int step_is_successful; /* -1 = invalid step */
} ptz_iter_log;

/* ---------------------------------------------------------------------------------------------------
 * PTZ bundle adjustment  (PTZRayOptimizer::Solve, ptzray_optimizer.cc:454-489)
 * ------------------------------------------------------------------------------------------------- */
typedef struct ptzba_problem {
  int factor_type;   /* PTZ_BA_* */
  int num_views;     /* V: candidate views only (isCandidate, ptzray_optimizer.cc:554-560) */
  int num_tracks;    /* P: tracks kept by FindTracks (:537-552) */
  int num_obs;       /* M: (track, candidate view) pairs, AddConstraints2d2d :811-848 */
  int num_pts3d;     /* A: annotated 2d-3d points, AddConstraints2d3d :887-925 */
  const double* intr;         /* [V*9]  initial intrinsics blocks  (:647) */
  const double* ext;          /* [V*6]  initial extrinsics blocks  (:651) */
  const float* obs_uv;        /* [M*2]  keypoint pixel */
  const int32_t* obs_view;    /* [M]    0..V-1 */
  const int32_t* obs_track;   /* [M]    0..P-1 */
  const double* track_weight; /* [P]    ScaledLoss weight = track.size() incl. non-candidates (:805-806) */
  const double* ray0;         /* [P*3]  initial rays, or NULL: computed as Pix2Ray (:768-797) */
  const float* pt_uv;         /* [A*2]  annotated pixel   (may be NULL when A==0) */
  const double* pt_xyz;       /* [A*3]  annotated world point */
  const int32_t* pt_view;     /* [A] */
  const double* tlw0;         /* [6]    initial T_l_w (:562-633) or NULL = zeros */
  const int32_t* shared_ic_id;/* [V]    SetSharedIntrinsics (:497-505): views with equal ids use ONE intrinsics block, that of the
                                        first such view (:640-650); NULL = every view its own.  Groups of >= 2 views put
                                        their 1-3 free intrinsics into the dense border of the reduced system: at most 16
                                        such unknowns in total (minus 3 for PTZRayDistDisp); not together with 2d-3d points */
} ptzba_problem;

typedef struct ptzba_result {
  int termination;           /* PTZ_CONVERGENCE / PTZ_NO_CONVERGENCE / PTZ_FAILURE */
  int num_iterations;        /* rows in Ceres' iteration table minus one */
  int num_successful_steps;  /* incl. iteration 0, as Ceres counts them */
  int num_unsuccessful_steps;
  int num_residuals;         /* 2*(M+A) */
  int linear_solver_iterations; /* total PCG iterations */
  double initial_cost;
  double final_cost;
  double init_reproj_error_all;  /* sqrt(2)*sqrt(2*cost/num_residuals), :962-963 */
  double final_reproj_error_all;
  double final_reproj_error_2d2d; /* unweighted RMS, :972-1028 (NaN when M==0) */
  double final_reproj_error_2d3d; /* :1030-1072 (NaN when A==0) */
  /* refined parameter blocks (caller-allocated, written for every termination type; the C++ adaptor
   * applies the reference's "only on CONVERGENCE" rule, :482-487) */
  double* intr;       /* [V*9] */
  double* ext;        /* [V*6] */
  double* ray;        /* [P*3]  local frame */
  double* disp;       /* [3] */
  double* tlw;        /* [6] */
  double* cams_world; /* [V*21] krt21 after ObtainRefinedCameraParams (:672-741), or NULL */
  double* rays_world; /* [P*3]  rays in the world frame (:743-754), or NULL */
  ptz_iter_log* log;  /* optional per-iteration table, or NULL */
  int log_capacity;
  int log_count;
  double seconds_setup; /* host+device set-up (sorting, structure) */
  double seconds_solve; /* LM loop */
} ptzba_result;

/* tangent-space layout used by ptzba_eval and the logs:
 *   per view i (nc = 5 for PTZRay, 6 otherwise):  [fx, fy, (k1), w1, w2, w3]   at i*nc
 *   per track p:                                   [x, y, z]                    at V*nc + 3p
 *   disp[3] (PTZRayDistDisp only), then tlw[6] (A>0 only)
 * which is Ceres' SubsetParameterization column selection (:861-882). */
typedef struct ptzba_eval_out {
  double cost;           /* 1/2 sum w |r|^2 */
  double* residuals;     /* [2*(M+A)] raw functor output (unweighted), 2d-2d first */
  double* jac_obs;       /* [M][2][nc+3(+3 disp)] raw d r/d(view cols, ray, disp), or NULL */
  double* jac_pts;       /* [A][2][nc+6(+3 disp)] raw d r/d(view cols, tlw, disp), or NULL */
  double* gradient;      /* [num_tangent] J^T r with the sqrt(w) loss scaling, or NULL */
  int num_tangent;
} ptzba_eval_out;

void ptz_solver_options_default(ptz_solver_options* opt);
/* measured fp64 FMA throughput of the current device in GFLOP/s (bench.py: denominator of the fp64-bound reloc kernel's roofline) */
int ptz_measure_fp64_gflops(double* gflops);
const char* ptz_last_error(void);
int ptz_device_count(void);

/* one-shot solve with HOST buffers: copies in, runs the LM loop on the current CUDA device, copies out */
int ptzba_solve(const ptzba_problem* prob, const ptz_solver_options* opt, ptzba_result* out);
/* stage 1 only, at the given parameters (prob->intr/ext/ray0/tlw0 are the evaluation point):
 * residuals, analytic Jacobian, cost and gradient — the hook for "Ceres evaluation at the same x" parity */
int ptzba_eval(const ptzba_problem* prob, const double* disp, ptzba_eval_out* out);

/* persistent-handle variant for the many BA calls of one IBA run and for benchmarking:
 * create = upload + structure set-up; iterate = run LM iterations on the resident problem. */
typedef struct ptzba_handle ptzba_handle;
int ptzba_create(const ptzba_problem* prob, const ptz_solver_options* opt, ptzba_handle** h);
int ptzba_reset(ptzba_handle* h);                               /* back to the initial parameters */
int ptzba_run(ptzba_handle* h, int max_new_iterations, ptzba_result* out); /* out arrays may be NULL */
/* device time (CUDA events on the solver's stream) of every kernel of the LM loop since create/reset, by kernel id */
enum {
  PTZ_K_VIEW_PREP = 0, PTZ_K_RESJAC, PTZ_K_VIEW_FINALIZE, PTZ_K_TRACK_ACCUM, PTZ_K_PTS,       /* stage 1 */
  PTZ_K_TRACK_SOLVE, PTZ_K_SCHUR_DIAG, PTZ_K_SCHUR_OFFDIAG, PTZ_K_PRECOND,                     /* stage 2 */
  PTZ_K_PCG,                                                                                   /* stage 3 */
  PTZ_K_TRACK_BACKSUB, PTZ_K_CAM_UPDATE, PTZ_K_COST, PTZ_K_SCALARS,                            /* stage 4 */
  PTZ_K_ALLREDUCE,
  PTZ_K_DEFLATE, /* stage 3: per-solve set-up of the deflated CG (basis scaling, S~W, S~S~W, Gram matrix, start vector) */
  PTZ_K_COUNT = 16
};
typedef struct ptzba_stage_times {
  float ms_kernel[16];   /* summed device time per kernel id */
  int launches[16];      /* launches per kernel id */
  float ms_run;          /* device span of the ptzba_run calls (first launch to last completion, host gaps included) */
  int lm_iterations, pcg_iterations, jacobian_evals, cost_evals;
  /* structure of the reduced camera system (bench.py: algorithmic bytes of stage 3 and of the off-diagonal Schur kernel) */
  int nnz_blocks;          /* blocks of S, both triangles + diagonal */
  int deflated_solves;     /* linear solves that ran the deflated kernel */
  int deflation_vectors;   /* size of the Ritz basis in use (0: none) */
  int reserved_;
  int64_t num_pairs;       /* observation pairs summed into the off-diagonal blocks */
} ptzba_stage_times;
int ptzba_get_stage_times(ptzba_handle* h, ptzba_stage_times* t);
/* per_kernel = 0: stop bracketing every kernel with CUDA events (ms_kernel[] stops accumulating; launches[] and ms_run keep counting).
 * The events cost two cudaEventRecord per launch; bench.py times its headline with them off and the per-kernel table in a second pass. */
int ptzba_set_stage_timing(ptzba_handle* h, int per_kernel);
int ptzba_destroy(ptzba_handle* h);

/* multi-GPU: the caller shards tracks (hence observations) across ranks and hands every rank the same
 * views AND the same annotated 2d-3d points (they are evaluated on rank 0 and reach the other ranks inside
 * the all-reduce); the library all-reduces camera blocks over its own NCCL communicator.  id_bytes is the 128-byte
 * ncclUniqueId made on rank 0 by ptz_nccl_unique_id and broadcast by the caller (torch.distributed). */
int ptz_nccl_unique_id(void* id_bytes128);
int ptz_nccl_init(const void* id_bytes128, int rank, int world_size);
int ptz_nccl_finalize(void);

/* ---------------------------------------------------------------------------------------------------
 * Batched PTZ relocalisation (KRTOptimizer, krt_optimizer.cc:257-404,504-567; one query = one
 * iteration of the loop at src/app/run_ptz_reloc.cc:68-118)
 * ------------------------------------------------------------------------------------------------- */
typedef struct ptzreloc_batch {
  int factor_type;              /* PTZ_KRT_* */
  int num_queries;              /* B */
  const int64_t* match_offset;  /* [B+1] rows of query b are match_offset[b] .. match_offset[b+1] */
  const float* uv_ref;          /* [N*2] kpts_ref[match.queryIdx].pt  (krt_optimizer.cc:289) */
  const float* uv_cur;          /* [N*2] kpts_curr[match.trainIdx].pt (:290) */
  const double* ref_cam;        /* [B*21] krt21 of cam_ref (Add2d2dConstraints, :265) */
  const double* init_cam;       /* [B*21] krt21 given to SetInitParams (:257-263) */
  int max_iter;                 /* KRTOptimizer ctor, :251 */
  double max_reproj_error;
  /* optional 2d-3d terms, Add2d3dConstraints (:350-383): Factor2d3dDist / Factor2d3dFxfyDist through cv::projectPoints
   * (uses the camera t and reads v[10..14] in OpenCV order k1,k2,p1,p2,k3).  All three NULL = none. */
  const int64_t* pt_offset;     /* [B+1] */
  const float* pt_uv;           /* [Np*2] pts2d */
  const double* pt_xyz;         /* [Np*3] pts3d in the WORLD frame (moved to the reference-local frame as at :357-362) */
} ptzreloc_batch;

typedef struct ptzreloc_result {
  double* cam;           /* [B*21] krt21 in the world frame (ObtainRefinedCameraParams :535-567);
                            written for every query, valid when success[b]!=0 */
  int32_t* success;      /* [B] KRTOptimizer::Solve return value (:398-403) */
  int32_t* termination;  /* [B] */
  int32_t* num_iter;     /* [B] num_iter_ = summary.num_successful_steps (:396) */
  int32_t* iterations;   /* [B] LM iterations, or NULL */
  double* initial_cost;  /* [B] or NULL */
  double* final_cost;    /* [B] or NULL */
  double* final_rms;     /* [B] sqrt(2)*sqrt(2*final_cost/num_residuals) (:507), or NULL */
  double* local_cam15;   /* [B*15] refined cam_curr_local_param_ (ref-local frame), or NULL */
} ptzreloc_result;

int ptzreloc_solve_batch(const ptzreloc_batch* batch, const ptz_solver_options* opt, ptzreloc_result* out);

/* stage-level hook: residuals + analytic Jacobian of one query at local_cam15, free columns in ascending
 * parameter index (F: fx,w | Fxfy: fx,fy,w | FDist: fx,w,k1 | FxfyDist: fx,fy,w,k1) */
int ptzreloc_eval(int factor_type, int num_matches, const float* uv_ref, const float* uv_cur,
                  const double* ref_cam21, const double* local_cam15,
                  double* residuals /*[2N]*/, double* jac /*[N][2][nfree]*/, double* cost, double* gradient);

/* KRTOptimizer::Cal2d2dReprojError (krt_optimizer.cc:406-455) and Cal2d3dReprojError (:457-500): unweighted RMS of the functor
 * residuals, sqrt(sum |r|^2 / count), at local_cam15 = cam_curr_local_param_ (the initial local parameters before Solve, the refined
 * ones -- ptzreloc_result.local_cam15 -- after).  pt_xyz are WORLD points (moved to the reference-local frame as at :466-470).
 * err_2d3d = -1 without points (:459-460); err_2d2d is NaN for an empty match list (0/0, as the reference). */
/* cam_curr_local_param_ right after Add2d2dConstraints (krt_optimizer.cc:269-286): the initial camera moved to the frame of the
 * reference camera and flattened by Camera::ToVector.  Plain host arithmetic (no device needed). */
int ptzreloc_local_params(const double* ref_cam21, const double* init_cam21, double* local_cam15);
int ptzreloc_reproj_error(int factor_type, const double* ref_cam21, const double* local_cam15, int num_matches, const float* uv_ref,
                          const float* uv_cur, int num_pts, const float* pt_uv, const double* pt_xyz, double* err_2d2d, double* err_2d3d);

/* device-resident variant used by bench.py's kernel-only timing: all pointers are DEVICE pointers */
int ptzreloc_solve_batch_dev(const ptzreloc_batch* dev_batch, const ptz_solver_options* opt,
                             ptzreloc_result* dev_out, void* cuda_stream);


/* ------------------------------------------------------------------------------------------------------------------
 * Track building: the step right before every BA, PTZRayOptimizer::FindTracks
 * (ptzray_optimizer.cc:537-552) = TracksBuilder::Build / Filter(4) / ExportToSTL (src/core/tracks.cc:19-113), and the
 * flattening of the tracks into the observation arrays of ptzba_problem (AddConstraints2d2d, ptzray_optimizer.cc:799-848).
 * On the device: radix sort + unique of the (image, feature) nodes, lock-free union-find over the matches, one more sort
 * by component, scans.  Integer work, results identical to the reference's as SETS; the reference's track ids are the
 * union-by-rank roots of its sequential UnionFind (union_find.h:66-92), which only order the tracks — here a track's id is
 * canonical: the flat index of its smallest (image, feature) node, and tracks come in ascending id.  ptztracks_reference_ids
 * (below) turns them into the reference's ids and order; the C++ adaptor does so by default.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ptztracks_matches {
  int32_t num_pairs;            /* matches_info.size() (struct MatchesInfo, types.h:24-35) */
  const int32_t* pair_src;      /* [num_pairs] MatchesInfo::src_img_idx */
  const int32_t* pair_dst;      /* [num_pairs] MatchesInfo::dst_img_idx */
  const int64_t* match_offset;  /* [num_pairs+1] rows of pair k are match_offset[k] .. match_offset[k+1]-1 */
  const int32_t* query_idx;     /* [N] cv::DMatch::queryIdx: feature of the src image (tracks.cc:29) */
  const int32_t* train_idx;     /* [N] cv::DMatch::trainIdx: feature of the dst image (tracks.cc:30) */
  int32_t min_track_length;     /* TracksBuilder::Filter argument; FindTracks passes 4 (ptzray_optimizer.cc:541) */
} ptztracks_matches;

typedef struct ptztracks_result {
  int32_t num_nodes;       /* out: distinct (image, feature) pairs = map_node_to_index_.size() (tracks.cc:35-43) */
  int32_t num_components;  /* out: connected components before Filter */
  int32_t num_tracks;      /* out: tracks ExportToSTL emits (no image listed twice, length >= min_track_length, > 1 node) */
  int64_t num_elems;       /* out: sum of their lengths */
  int64_t cap_tracks;      /* in: capacity of track_id (track_offset holds cap_tracks+1); N is always enough */
  int64_t cap_elems;       /* in: capacity of elem_img / elem_feat; 2N is always enough */
  int32_t* track_id;       /* [num_tracks] ascending */
  int64_t* track_offset;   /* [num_tracks+1] */
  int32_t* elem_img;       /* [num_elems] per track in ascending image id (Track = std::map<int,int>, tracks.h:31) */
  int32_t* elem_feat;      /* [num_elems] */
} ptztracks_result;

/* PTZ_ERR_INVALID with the counts filled in when a capacity is too small, or for a negative image / feature index */
int ptztracks_build(const ptztracks_matches* matches, ptztracks_result* out);
/* all pointers (of both structs) are DEVICE pointers; counts come back in the struct (bench.py's kernel-only timing).
 * Scratch memory is cached per stream: pass a stream you created (with the legacy default stream, NULL, every call pays cudaMalloc). */
int ptztracks_build_dev(const ptztracks_matches* dev_matches, int64_t num_matches, ptztracks_result* dev_out, void* cuda_stream);

/* Reference-id mode: relabels the tracks of ptztracks_build with the ids the reference's sequential union-by-rank forest gives them
 * (union_find.h:66-92 replayed over the matches in order, on the host) and re-sorts them in ascending id, i.e. into the iteration
 * order of the reference's Tracks map.  After this call track ids, track order and therefore Ray::id_ and the order of the residual
 * blocks are the reference's, not only the tracks as sets.  Host code: O(N log N) for N matches, no device needed. */
int ptztracks_reference_ids(const ptztracks_matches* matches, ptztracks_result* tracks);

/* tracks -> observation rows of ptzba_problem, as the loop at ptzray_optimizer.cc:801-848 adds residual blocks: tracks in
 * ascending id, inside a track ascending image id, candidate views only (isCandidate, :554-560); a track without any
 * candidate view gets no row; track_weight = number of images of the WHOLE track (:805); views are renumbered densely
 * in ascending image id. */
typedef struct ptztracks_views {
  int32_t num_images;
  const uint8_t* is_candidate;  /* [num_images] cam_ids_ membership */
  const int64_t* kp_offset;     /* [num_images+1] keypoints of image i are rows kp_offset[i] .. kp_offset[i+1]-1 */
  const float* kp_uv;           /* [.. * 2] features_[i].keypoints[j].pt (struct ImageFeatures, types.h:17-22) */
} ptztracks_views;

typedef struct ptztracks_obs {
  int32_t num_rows;       /* out: tracks with at least one candidate view = ptzba_problem.num_tracks */
  int32_t num_obs;        /* out: ptzba_problem.num_obs */
  int64_t cap_rows, cap_obs; /* in: num_tracks and num_elems of the tracks are always enough */
  int32_t* row_track;     /* [num_rows] index into the input tracks */
  double* track_weight;   /* [num_rows] */
  float* obs_uv;          /* [num_obs*2] */
  int32_t* obs_view;      /* [num_obs] dense candidate-view index */
  int32_t* obs_track;     /* [num_obs] row */
} ptztracks_obs;

int ptztracks_flatten(const ptztracks_result* tracks, const ptztracks_views* views, ptztracks_obs* out);


/* ------------------------------------------------------------------------------------------------------------------
 * Initial local->world transform of a georeferencing BA: PTZRayOptimizer::SetInitTransLocalToWorld
 * (ptzray_optimizer.cc:562-633).  For each view in order that has annotated points: cv::solvePnP(SOLVEPNP_EPNP) (:572, restated
 * in ptzcalib_epnp.hpp), the gates of :582-604, and T_l_w = T_i_l^-1 T_i_w (:606-616); the first view that passes gives tlw =
 * [rvec | t] (what ptzba_problem.tlw0 takes) and *view_used; when none passes tlw = 0, *view_used = -1 (the reference then
 * starts from zeros too).  Host code, a few tens of points: no device needed.  cams21 holds the CANDIDATE views only, in the
 * order of ptzba_problem's views, in the krt21 layout of ptzreloc_* (fx fy cx cy | R row-major | t | k1 k2 p1 p2 k3).
 * ------------------------------------------------------------------------------------------------------------------ */
int ptzgeo_init_tlw(int32_t num_views, const double* cams21, const int64_t* pt_offset /* [num_views+1] */,
                    const float* pt_uv /* [.. * 2] */, const double* pt_xyz /* [.. * 3] */, double tlw[6], int32_t* view_used);

#ifdef __cplusplus
}
#endif
#endif /* PTZCALIB_B200_H */
