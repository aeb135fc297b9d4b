// iba_oracle.cpp — CPU restatement of the PTZ-IBA driver, ptzcalib::PtzIncrementalOptimizer
// (/root/reference/src/core/ptz_incremental_optimizer.cc:39-441), on top of the oracle's own solvers.
//
// TEST INFRASTRUCTURE ONLY (see ptz_oracle.cpp's header): it exists so that the GPU driver in include/ptzcalib_b200.hpp can be
// checked decision by decision — seed pair, order of the registration attempts, which registered neighbour an image was
// registered from, when a global BA runs and whether it is accepted, the un-register rule — and not only against ground truth.
// It follows the reference's SEQUENTIAL control flow (one KRT solve per candidate neighbour, first success wins, :384-415), where
// the product batches those solves; every numerical step goes to the oracle: tracks (orc_tracks_build / orc_tracks_flatten =
// TracksBuilder, tracks.cc:19-118), BA (orc_ptzba_solve = PTZRayOptimizer::Solve, ptzray_optimizer.cc:454-489) and registration
// (orc_ptzreloc_solve_batch with one query = KRTOptimizer, krt_optimizer.cc:257-404).
//
// Parity unpinned against the reference itself (it cannot be built here: Ceres / OpenCV C++ are absent); the decision logic below is
// a line-by-line restatement, each function citing the lines it follows.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/ptzcalib_b200.h"
#include "iba_oracle.h"

extern "C" {
void orc_solver_options_default(ptz_solver_options* o);
void orc_rodrigues_inv(const double* R, double* r);
void orc_inv3(const double* S, double* D);
int orc_ptzba_solve(const ptzba_problem* prob, const ptz_solver_options* opt, ptzba_result* out);
int orc_ptzreloc_solve_batch(const ptzreloc_batch* b, const ptz_solver_options* opt, ptzreloc_result* out);
int orc_tracks_build(const ptztracks_matches* m, ptztracks_result* out);
int orc_tracks_flatten(const ptztracks_result* tr, const ptztracks_views* v, ptztracks_obs* out);
}

namespace {

struct Cam {  // Camera, types.h:49-100: K (fx, fy, cx, cy; no skew), R, t, dist
  double K[9], R[9], t[3], d[5];
  void from21(const double* c) {
    const double k[9] = {c[0], 0, c[2], 0, c[1], c[3], 0, 0, 1};
    std::memcpy(K, k, sizeof(K)); std::memcpy(R, c + 4, sizeof(R)); std::memcpy(t, c + 13, sizeof(t)); std::memcpy(d, c + 16, sizeof(d));
  }
  void to21(double* c) const {
    c[0] = K[0]; c[1] = K[4]; c[2] = K[2]; c[3] = K[5];
    std::memcpy(c + 4, R, sizeof(R)); std::memcpy(c + 13, t, sizeof(t)); std::memcpy(c + 16, d, sizeof(d));
  }
};

void mul33(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}

struct Driver {
  const orc_iba_input& in;
  orc_iba_output& out;
  std::vector<Cam> cameras;
  std::unordered_set<long> init_image_pairs;          // ptz_incremental_optimizer.h: init_image_pairs_
  std::unordered_map<long, size_t> num_reg_trials;    // num_reg_trials_
  std::unordered_set<long> reg;                       // reg_image_ids_
  double last_reproj = 0;
  int last_ba_iterations = 0;
  static constexpr long kMaxNumImages = 100000;       // .cc:24
  static constexpr float kBaGlobalImagesRatio = 1.1f; // .cc:25

  Driver(const orc_iba_input& i, orc_iba_output& o) : in(i), out(o), cameras(i.num_images) {
    for (int k = 0; k < in.num_images; ++k) cameras[k].from21(in.cams21 + 21 * (size_t)k);
    out.num_events = 0;
  }
  void event(long kind, long a, long b) {
    if (out.events && out.num_events < out.cap_events) {
      out.events[3 * (size_t)out.num_events] = kind; out.events[3 * (size_t)out.num_events + 1] = a; out.events[3 * (size_t)out.num_events + 2] = b;
    }
    ++out.num_events;
  }
  bool is_reg(long id) const { return reg.find(id) != reg.end(); }
  const float* kp(long img, int f) const { return in.kp_uv + 2 * (in.kp_offset[img] + f); }

  // the tail shared by FindFirstInitialImage / FindSecondInitialImage / FindNextImages (.cc:196-205, 236-246, 279-289)
  static std::vector<long> ranked(const std::vector<float>& confidences_rank) {
    std::vector<long> indices(confidences_rank.size());
    std::iota(indices.begin(), indices.end(), 0);
    std::sort(indices.begin(), indices.end(), [&](int A, int B) -> bool { return confidences_rank[A] > confidences_rank[B]; });
    std::vector<long> sorted_ids;
    for (long index : indices) {
      if (confidences_rank[index] <= 0.0f) break;
      sorted_ids.push_back(index);
    }
    return sorted_ids;
  }
  // .cc:292-304: mean pixel distance of the matched keypoints; float accumulator, cv::norm(Point2f) returns double
  float cal_pixel_diff(long id1, long id2, int pair) const {
    const int64_t m0 = in.match_offset[pair], m1 = in.match_offset[pair + 1];
    float total_dists = 0.0f;
    for (int64_t m = m0; m < m1; ++m) {
      const float* p1 = kp(id1, in.query_idx[m]);
      const float* p2 = kp(id2, in.train_idx[m]);
      const float dx = p1[0] - p2[0], dy = p1[1] - p2[1];
      total_dists += std::sqrt((double)dx * dx + (double)dy * dy);
    }
    return total_dists * 1.0f / (size_t)(m1 - m0);
  }
  long pair_id(long a, long b) const { return a < b ? a * kMaxNumImages + b : b * kMaxNumImages + a; }  // .cc:306-312

  std::vector<long> find_first_initial_image() const {  // .cc:179-206
    std::vector<float> rank(in.num_images, 0.0f);
    for (int k = 0; k < in.num_pairs; ++k) { rank[in.pair_src[k]] += in.confidence[k]; rank[in.pair_dst[k]] += in.confidence[k]; }
    return ranked(rank);
  }
  std::vector<long> find_second_initial_image(long image_id1) const {  // .cc:208-247
    std::vector<float> rank(in.num_images, 0.0f);
    const float kMinPixelDiff = 50;
    for (int k = 0; k < in.num_pairs; ++k) {
      const long src = in.pair_src[k], dst = in.pair_dst[k];
      if (in.match_offset[k + 1] == in.match_offset[k]) continue;
      if (image_id1 != src && image_id1 != dst) continue;
      if (image_id1 == src && image_id1 == dst) continue;
      if (cal_pixel_diff(src, dst, k) < kMinPixelDiff) continue;
      if (image_id1 == src && image_id1 != dst) rank[dst] += in.confidence[k];
      if (image_id1 != src && image_id1 == dst) rank[src] += in.confidence[k];
    }
    return ranked(rank);
  }
  std::vector<long> find_next_images() const {  // .cc:249-290
    std::vector<float> rank(in.num_images, 0.0f);
    const size_t kMaxRegTrials = 4;
    for (int k = 0; k < in.num_pairs; ++k) {
      const long src = in.pair_src[k], dst = in.pair_dst[k];
      if (src == dst) continue;
      if (!in.has_H[k]) continue;
      if (num_reg_trials.count(src) != 0 && num_reg_trials.at(src) > kMaxRegTrials) continue;
      if (num_reg_trials.count(dst) != 0 && num_reg_trials.at(dst) > kMaxRegTrials) continue;
      const bool rs = is_reg(src), rd = is_reg(dst);
      if (rs && rd) continue;          // already registered
      else if (!rs && !rd) continue;   // not a neighbour of the model
      else if (rs && !rd) rank[dst] += in.confidence[k];
      else rank[src] += in.confidence[k];
    }
    return ranked(rank);
  }
  bool find_initial_image_pair(long& id1, long& id2) {  // .cc:142-177
    std::vector<long> ids1;
    if (in.num_seeds == 0) ids1 = find_first_initial_image();
    else ids1.assign(in.seed_ids, in.seed_ids + in.num_seeds);
    for (size_t i1 = 0; i1 < ids1.size(); ++i1) {
      id1 = ids1[i1];
      const std::vector<long> ids2 = find_second_initial_image(id1);
      for (size_t i2 = 0; i2 < ids2.size(); ++i2) {
        id2 = ids2[i2];
        const long pid = pair_id(id1, id2);
        if (init_image_pairs.count(pid) > 0) continue;  // every pair only once
        init_image_pairs.insert(pid);
        return true;
      }
    }
    id1 = id2 = std::numeric_limits<long>::max();
    return false;
  }
  // R_j = K_j^-1 H_ji K_i R_i (.cc:343-346, 391-393)
  void rotation_from_h(const double* Kj, const double* H, const double* Ki, const double* Ri, double* Rj) const {
    double Kinv[9], a[9], b[9];
    orc_inv3(Kj, Kinv);
    mul33(Kinv, H, a);
    mul33(a, Ki, b);
    mul33(b, Ri, Rj);
  }
  void set_initial_image_pair_parameters(long id1, long id2) {  // .cc:314-350
    const double ratio = 1.2;
    for (long id : {id1, id2}) {
      const double focal = ratio * std::max(in.img_w[id], in.img_h[id]);
      cameras[id].K[0] = cameras[id].K[4] = focal;
      cameras[id].K[2] = 0.5 * in.img_w[id];
      cameras[id].K[5] = 0.5 * in.img_h[id];
      if (id == id1) { const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; std::memcpy(cameras[id].R, I, sizeof(I)); }
    }
    for (int k = 0; k < in.num_pairs; ++k)
      if (in.pair_src[k] == id1 && in.pair_dst[k] == id2) {
        double Rj[9];
        rotation_from_h(cameras[id2].K, in.H + 9 * (size_t)k, cameras[id1].K, cameras[id1].R, Rj);
        std::memcpy(cameras[id2].R, Rj, sizeof(Rj));
        break;
      }
  }

  // PTZRayOptimizer(features, matches, cameras, cam_ids, max_iter, PTZRay).Solve(cameras): ptzray_optimizer.cc:454-489 with
  // FindTracks (:537-552) over the WHOLE table, rows for candidate views only (:811-848), outputs on CONVERGENCE only (:482-487)
  bool ptzray_solve(const std::unordered_set<long>& cam_ids) {
    last_ba_iterations = 0;
    const int n = in.num_images;
    if (n == 0 || in.max_iter <= 0) return false;  // CheckValid
    const int64_t N = in.match_offset[in.num_pairs];
    ptztracks_matches mm{in.num_pairs, in.pair_src, in.pair_dst, in.match_offset, in.query_idx, in.train_idx, 4};
    const size_t capT = std::max<int64_t>(N, 1), capE = std::max<int64_t>(2 * N, 1);
    std::vector<int32_t> tid(capT), eimg(capE), efeat(capE);
    std::vector<int64_t> toff(capT + 1);
    ptztracks_result tr{};
    tr.cap_tracks = (int64_t)capT; tr.cap_elems = (int64_t)capE;
    tr.track_id = tid.data(); tr.track_offset = toff.data(); tr.elem_img = eimg.data(); tr.elem_feat = efeat.data();
    if (orc_tracks_build(&mm, &tr) != PTZ_OK) return false;
    std::vector<uint8_t> cand(n, 0);
    std::vector<long> view_of;
    for (int i = 0; i < n; ++i)
      if (cam_ids.find(i) != cam_ids.end()) { cand[i] = 1; view_of.push_back(i); }
    ptztracks_views tv{n, cand.data(), in.kp_offset, in.kp_uv};
    const size_t nt = std::max<size_t>(tr.num_tracks, 1), ne = std::max<size_t>((size_t)tr.num_elems, 1);
    std::vector<int32_t> row_track(nt), oview(ne), otrack(ne);
    std::vector<double> weight(nt);
    std::vector<float> uv(2 * ne);
    ptztracks_obs ob{};
    ob.cap_rows = (int64_t)nt; ob.cap_obs = (int64_t)ne;
    ob.row_track = row_track.data(); ob.track_weight = weight.data(); ob.obs_uv = uv.data(); ob.obs_view = oview.data(); ob.obs_track = otrack.data();
    if (orc_tracks_flatten(&tr, &tv, &ob) != PTZ_OK) return false;
    // SetUpInitialCameraParams (:635-670): intr = [fx, fy, cx, cy, d0..d4], ext = [rvec, t] from Camera::ToVector
    std::vector<double> intr, ext;
    for (long id : view_of) {
      const Cam& c = cameras[id];
      const double in9[9] = {c.K[0], c.K[4], c.K[2], c.K[5], c.d[0], c.d[1], c.d[2], c.d[3], c.d[4]};
      double rv[3];
      orc_rodrigues_inv(c.R, rv);
      const double ex6[6] = {rv[0], rv[1], rv[2], c.t[0], c.t[1], c.t[2]};
      intr.insert(intr.end(), in9, in9 + 9);
      ext.insert(ext.end(), ex6, ex6 + 6);
    }
    ptzba_problem p{};
    p.factor_type = PTZ_BA_PTZRAY;
    p.num_views = (int)view_of.size(); p.num_tracks = ob.num_rows; p.num_obs = ob.num_obs; p.num_pts3d = 0;
    p.intr = intr.data(); p.ext = ext.data(); p.obs_uv = uv.data(); p.obs_view = oview.data(); p.obs_track = otrack.data(); p.track_weight = weight.data();
    ptz_solver_options o;
    orc_solver_options_default(&o);
    o.max_num_iterations = in.max_iter;
    std::vector<double> cams_w(21 * std::max<size_t>(view_of.size(), 1));
    ptzba_result r{};
    r.cams_world = cams_w.data();
    if (orc_ptzba_solve(&p, &o, &r) != PTZ_OK) return false;
    last_reproj = r.final_reproj_error_all;
    last_ba_iterations = r.num_iterations;
    if (r.termination != PTZ_CONVERGENCE) return false;
    for (size_t k = 0; k < view_of.size(); ++k) cameras[view_of[k]].from21(&cams_w[21 * k]);
    return true;
  }

  bool register_initial_image_pair(long id1, long id2) {  // .cc:352-375
    num_reg_trials[id1] += 1;
    num_reg_trials[id2] += 1;
    init_image_pairs.insert(pair_id(id1, id2));
    set_initial_image_pair_parameters(id1, id2);
    const bool ok = ptzray_solve(std::unordered_set<long>{id1, id2});
    if (ok) { reg.insert(id1); reg.insert(id2); }
    return ok;
  }

  bool register_next_image(long image_id, long* via) {  // .cc:377-419, sequential as there
    num_reg_trials[image_id] += 1;
    *via = -1;
    for (int k = 0; k < in.num_pairs; ++k) {
      const long i = in.pair_src[k], j = in.pair_dst[k];
      if (!in.has_H[k]) continue;
      if (reg.count(i) == 1 && j == image_id) {
        std::memcpy(cameras[j].K, cameras[i].K, sizeof(cameras[j].K));
        double Rj[9];
        rotation_from_h(cameras[j].K, in.H + 9 * (size_t)k, cameras[i].K, cameras[i].R, Rj);
        std::memcpy(cameras[j].R, Rj, sizeof(Rj));
        // KRTOptimizer(100, 100, F): SetInitParams(cameras_[j]) + Add2d2dConstraints(cameras_[i], kpts_i, kpts_j, matches) + Solve
        const int64_t m0 = in.match_offset[k], m1 = in.match_offset[k + 1];
        std::vector<float> uv1, uv2;
        for (int64_t m = m0; m < m1; ++m) {
          const float* a = kp(i, in.query_idx[m]);
          const float* b = kp(j, in.train_idx[m]);
          uv1.push_back(a[0]); uv1.push_back(a[1]); uv2.push_back(b[0]); uv2.push_back(b[1]);
        }
        double ref[21], init[21], res[21];
        cameras[i].to21(ref);
        cameras[j].to21(init);
        const int64_t off[2] = {0, m1 - m0};
        ptzreloc_batch b{};
        b.factor_type = PTZ_KRT_F; b.num_queries = 1; b.match_offset = off; b.uv_ref = uv1.data(); b.uv_cur = uv2.data();
        b.ref_cam = ref; b.init_cam = init; b.max_iter = 100; b.max_reproj_error = 100;  // .cc:395-396
        int32_t ok = 0, term = 0, nit = 0;
        ptzreloc_result r{};
        r.cam = res; r.success = &ok; r.termination = &term; r.num_iter = &nit;
        ptz_solver_options o;
        orc_solver_options_default(&o);
        if (orc_ptzreloc_solve_batch(&b, &o, &r) != PTZ_OK) continue;
        if (ok) {
          Cam c;
          c.from21(res);
          std::memcpy(cameras[j].K, c.K, sizeof(c.K));
          std::memcpy(cameras[j].R, c.R, sizeof(c.R));
          reg.insert(j);
          *via = i;
          return true;
        }
      }
    }
    return false;
  }

  bool adjust_global_bundle() {  // .cc:421-439
    const bool ok = ptzray_solve(reg);
    event(ORC_IBA_GLOBAL_BA, ok ? 1 : 0, (long)reg.size());
    return ok;
  }

  bool solve() {  // .cc:39-125
    if (in.num_images == 0 || in.max_iter <= 0) return false;  // CheckValid (.cc:134-140)
    const int kInitNumTrials = 50;
    for (int num_trials = 0; num_trials < kInitNumTrials; ++num_trials) {
      long id1, id2;
      if (!find_initial_image_pair(id1, id2)) return false;
      event(ORC_IBA_SEED_PAIR, id1, id2);
      const bool init_ok = register_initial_image_pair(id1, id2);
      event(ORC_IBA_INIT_RESULT, init_ok ? 1 : 0, last_ba_iterations);
      if (!init_ok) continue;
      adjust_global_bundle();
      size_t ba_prev_num_reg_images = reg.size();
      bool reg_next_success = true;
      while (reg_next_success) {
        reg_next_success = false;
        const std::vector<long> next_image_ids = find_next_images();
        event(ORC_IBA_NEXT_LIST, (long)next_image_ids.size(), next_image_ids.empty() ? -1 : next_image_ids[0]);
        if (next_image_ids.empty()) break;
        for (size_t reg_trial = 0; reg_trial < next_image_ids.size(); ++reg_trial) {
          const long image_id = next_image_ids[reg_trial];
          long via = -1;
          reg_next_success = register_next_image(image_id, &via);
          event(ORC_IBA_REGISTER, image_id, via);
          if (reg_next_success) {
            if (reg.size() >= kBaGlobalImagesRatio * ba_prev_num_reg_images) {
              const bool gba_success = adjust_global_bundle();
              if (gba_success) {
                ba_prev_num_reg_images = reg.size();
                break;
              } else {
                reg.erase(image_id);
                event(ORC_IBA_UNREGISTER, image_id, 0);
                reg_next_success = false;
              }
            }
          }
          if (!reg_next_success) {
            const long kMinNumInitialRegTrials = 30;
            const int kMinModelSize = 3;
            if ((long)reg_trial >= kMinNumInitialRegTrials && reg.size() < (size_t)kMinModelSize) break;
          }
        }
      }
      adjust_global_bundle();
      return true;
    }
    return false;
  }
};

}  // namespace

extern "C" int orc_iba_solve(const orc_iba_input* in, orc_iba_output* out) {
  if (!in || !out || in->num_images < 0) return PTZ_ERR_INVALID;
  Driver d(*in, *out);
  out->ok = d.solve() ? 1 : 0;
  out->num_registered = (int32_t)d.reg.size();
  out->last_reproj_error = d.last_reproj;
  for (int i = 0; i < in->num_images; ++i) {
    if (out->registered) out->registered[i] = d.reg.count(i) ? 1 : 0;
    if (out->cams21) d.cameras[i].to21(out->cams21 + 21 * (size_t)i);
  }
  return PTZ_OK;
}
