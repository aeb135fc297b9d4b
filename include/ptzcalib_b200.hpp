// ptzcalib_b200.hpp — header-only C++ adaptor: the reference's two solver classes on top of the C ABI.
//
// Same class names, constructor / method signatures, argument meaning and error behaviour as
//   ptzcalib::PTZRayOptimizer   src/core/ptzray_optimizer.h:112-177, .cc:405-766
//   ptzcalib::KRTOptimizer      src/core/krt_optimizer.h:108-145,    .cc:251-567
// and the value types of src/core/types.h (Camera, ImageFeatures, MatchesInfo, Ray) and src/core/tracks.h (Tracks,
// TracksBuilder) they take — with plain-array stand-ins for the OpenCV types (cv::Mat 3x3 -> Mat33, cv::KeyPoint ->
// KeyPoint{pt}, cv::DMatch -> DMatch{queryIdx, trainIdx}), because OpenCV's C++ headers are not part of this build.
// A caller of the reference switches by including this header instead of ptzray_optimizer.h / krt_optimizer.h and
// linking libptzcalib_b200.so; all numerics run on the GPU behind ptzba_solve / ptzreloc_solve_batch.
//
// PTZRayOptimizer::SetInitTransLocalToWorld (.cc:562-633) runs on the host as in the reference; its cv::solvePnP(EPNP) is restated in
// ptzcalib_epnp.hpp.  SetInitTransLocalToWorld(const double[6]) is an extra overload for callers that already know T_l_w.
#ifndef PTZCALIB_B200_HPP
#define PTZCALIB_B200_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <map>
#include <numeric>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../ptz-calib_b200/csrc/ptz_math.cuh"  // rodrigues_jac / rodrigues_inv / mul33 (plain C++ when not compiled by nvcc)
#include "ptzcalib_b200.h"
#include "ptzcalib_epnp.hpp"

namespace ptzcalib {

struct Point2f { float x = 0, y = 0; };
struct Point3d { double x = 0, y = 0, z = 0; };
struct Size { int width = 0, height = 0; };
struct KeyPoint { Point2f pt; };
struct DMatch { int queryIdx = 0, trainIdx = 0; };
typedef std::array<double, 9> Mat33;  // row-major
typedef std::array<double, 3> Vec3;
typedef std::array<double, 5> Vec5;

// types.h:17-45
struct ImageFeatures { long img_idx = 0; Size img_size; std::vector<KeyPoint> keypoints; };
struct MatchesInfo {
  long src_img_idx = 0, dst_img_idx = 0;
  std::vector<DMatch> matches;
  std::vector<unsigned char> inliers_mask;
  int num_inliers = 0;
  Mat33 H{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
  bool has_H = false;  // stands for !H.empty() of the cv::Mat (data_io.cc fills H for image pairs that passed RANSAC)
  double confidence = 0;
};
struct Ray {
  int id_;
  Point3d pt3d_;
  Point2f uv_;
  Ray(int id, const Vec3& p, const Point2f& uv) : id_(id), uv_(uv) { pt3d_.x = p[0]; pt3d_.y = p[1]; pt3d_.z = p[2]; }
};

// types.h:47-97, types.cc:16-73
class Camera {
 public:
  Camera() : K_{{1, 0, 0, 0, 1, 0, 0, 0, 1}}, R_{{1, 0, 0, 0, 1, 0, 0, 0, 1}}, t_{{0, 0, 0}}, dist_{{0, 0, 0, 0, 0}} {}
  Camera(const Mat33& K, const Mat33& R, const Vec3& t, const Vec5& dist) : K_(K), R_(R), t_(t), dist_(dist) {}
  const Mat33& K() const { return K_; }
  Mat33& K() { return K_; }
  const Mat33& R() const { return R_; }
  Mat33& R() { return R_; }
  const Vec3& t() const { return t_; }
  Vec3& t() { return t_; }
  const Vec5& dist() const { return dist_; }
  Vec5& dist() { return dist_; }
  Vec3 rvec() const { Vec3 r; ptz::rodrigues_inv(R_.data(), r.data()); return r; }
  std::vector<double> ToVector() const {  // fx, fy, cx, cy, rvec, t, dist
    std::vector<double> v(15);
    v[0] = K_[0]; v[1] = K_[4]; v[2] = K_[2]; v[3] = K_[5];
    ptz::rodrigues_inv(R_.data(), &v[4]);
    for (int i = 0; i < 3; ++i) v[7 + i] = t_[i];
    for (int i = 0; i < 5; ++i) v[10 + i] = dist_[i];
    return v;
  }
  void FromVector(const std::vector<double>& v) {
    if (v.size() != 15) throw std::invalid_argument("Expected camera vector size: 15, actual size :" + std::to_string(v.size()));
    K_ = Mat33{{v[0], 0, v[2], 0, v[1], v[3], 0, 0, 1}};
    ptz::rodrigues_jac(&v[4], R_.data(), nullptr);
    for (int i = 0; i < 3; ++i) t_[i] = v[7 + i];
    for (int i = 0; i < 5; ++i) dist_[i] = v[10 + i];
  }
  void ToKrt21(double* o) const {
    o[0] = K_[0]; o[1] = K_[4]; o[2] = K_[2]; o[3] = K_[5];
    for (int i = 0; i < 9; ++i) o[4 + i] = R_[i];
    for (int i = 0; i < 3; ++i) o[13 + i] = t_[i];
    for (int i = 0; i < 5; ++i) o[16 + i] = dist_[i];
  }
  void FromKrt21(const double* c) {
    K_ = Mat33{{c[0], 0, c[2], 0, c[1], c[3], 0, 0, 1}};
    for (int i = 0; i < 9; ++i) R_[i] = c[4 + i];
    for (int i = 0; i < 3; ++i) t_[i] = c[13 + i];
    for (int i = 0; i < 5; ++i) dist_[i] = c[16 + i];
  }

 private:
  Mat33 K_, R_;
  Vec3 t_;
  Vec5 dist_;
};

// tracks.h:24-60, tracks.cc:19-113 (openMVG-style union-find over (image, feature) nodes)
typedef std::pair<int, int> IndexedFeaturePair;
typedef std::map<int, int> Track;    // image id -> feature id
typedef std::map<int, Track> Tracks;  // track id -> track

class TracksBuilder {
 public:
  void Build(const std::vector<MatchesInfo>& matches_info) {
    std::set<IndexedFeaturePair> all;
    for (const auto& mi : matches_info)
      for (const auto& m : mi.matches) { all.emplace((int)mi.src_img_idx, m.queryIdx); all.emplace((int)mi.dst_img_idx, m.trainIdx); }
    nodes_.assign(all.begin(), all.end());  // sorted: node -> flat index by position
    parent_.resize(nodes_.size());
    std::iota(parent_.begin(), parent_.end(), 0);
    rank_.assign(nodes_.size(), 0);
    size_.assign(nodes_.size(), 1);
    for (const auto& mi : matches_info)
      for (const auto& m : mi.matches) Union(IndexOf({(int)mi.src_img_idx, m.queryIdx}), IndexOf({(int)mi.dst_img_idx, m.trainIdx}));
  }
  // remove tracks that list an image twice or are shorter than min_track_length
  void Filter(int min_track_length = 2) {
    std::map<int, std::set<int>> tracks;
    std::set<int> bad;
    for (int k = 0; k < (int)nodes_.size(); ++k) {
      const int id = Find(k);
      if (!tracks[id].insert(nodes_[k].first).second) bad.insert(id);
    }
    for (const auto& t : tracks) if ((int)t.second.size() < min_track_length) bad.insert(t.first);
    for (int k = 0; k < (int)nodes_.size(); ++k) Find(k);  // full path compression: parent_ == root
    for (int& root : parent_)
      if (bad.count(root) > 0) { size_[root] = 1; root = std::numeric_limits<int>::max(); }
  }
  void ExportToSTL(Tracks& tracks) {
    tracks.clear();
    for (int k = 0; k < (int)nodes_.size(); ++k) {
      const int id = parent_[k];
      if (id != std::numeric_limits<int>::max() && size_[id] > 1) tracks[id].insert(nodes_[k]);
    }
  }

 private:
  int IndexOf(const IndexedFeaturePair& p) const { return (int)(std::lower_bound(nodes_.begin(), nodes_.end(), p) - nodes_.begin()); }
  int Find(int i) {
    while (parent_[i] != i) { parent_[i] = parent_[parent_[i]]; i = parent_[i]; }
    return i;
  }
  void Union(int a, int b) {
    a = Find(a); b = Find(b);
    if (a == b) return;
    if (rank_[a] < rank_[b]) std::swap(a, b);
    parent_[b] = a; size_[a] += size_[b];
    if (rank_[a] == rank_[b]) ++rank_[a];
  }
  std::vector<IndexedFeaturePair> nodes_;
  std::vector<int> parent_, rank_, size_;
};

// tracks.cc:120-136: total / longest / shortest track length
inline void Length(const Tracks& tracks, int& total_length, int& max_length, int& min_length) {
  total_length = 0; max_length = 0; min_length = std::numeric_limits<int>::max();
  for (const auto& t : tracks) {
    const int n = (int)t.second.size();
    total_length += n; max_length = std::max(max_length, n); min_length = std::min(min_length, n);
  }
}
// tracks.cc:150-202: the largest set of images connected through tracks.  The reference merges std::sets track by track; the result is
// the largest connected component of the image graph, found here with a union-find over the images (ties: the component holding the
// smallest image id).
inline void FindMaxCoVisible(const Tracks& tracks, int num_images, std::set<int>& max_connect_imgs) {
  std::vector<int> parent(std::max(num_images, 0));
  std::iota(parent.begin(), parent.end(), 0);
  std::vector<char> seen(parent.size(), 0);
  auto find = [&](int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
  for (const auto& t : tracks) {
    int first = -1;
    for (const auto& e : t.second) {
      if (e.first < 0 || e.first >= num_images) continue;
      seen[e.first] = 1;
      if (first < 0) first = e.first; else parent[find(e.first)] = find(first);
    }
  }
  std::map<int, std::vector<int>> comps;
  for (int i = 0; i < num_images; ++i) if (seen[i]) comps[find(i)].push_back(i);
  max_connect_imgs.clear();
  size_t best = 0;
  int best_min = std::numeric_limits<int>::max();
  for (const auto& c : comps)
    if (c.second.size() > best || (c.second.size() == best && c.second.front() < best_min)) {
      best = c.second.size(); best_min = c.second.front();
      max_connect_imgs = std::set<int>(c.second.begin(), c.second.end());
    }
}

enum FACTOR_TYPE { PTZRay, PTZRayDist, PTZRayFxfyDist, PTZRayDistDisp };  // ptzray_optimizer.h:110

// What FindTracks produces.  It depends on the match table alone (ptzray_optimizer.cc:537-552 builds the tracks from ALL of
// matches_info_, whatever the candidate set), so a caller that solves many BAs over one table -- the incremental driver -- builds it
// once and hands it to every optimizer (PTZRayOptimizer::SetTrackCache) instead of re-running the build per BA as the reference does.
struct TrackCache {
  Tracks tracks;
  std::vector<int32_t> flat_id, flat_img, flat_feat;
  std::vector<int64_t> flat_off;
};

class PTZRayOptimizer {
 public:
  PTZRayOptimizer(const std::vector<ImageFeatures>& features, const std::vector<MatchesInfo>& matches_info, const std::vector<Camera>& cameras,
                  const std::vector<std::vector<Point2f>>& pixels, const std::vector<std::vector<Point3d>>& pts3d,
                  const std::unordered_set<long>& cam_ids, int max_iter, FACTOR_TYPE type)
      : cameras_(cameras), features_(features), matches_info_(matches_info), pixels_(pixels), pts3d_(pts3d), num_cams_(cameras.size()), type_(type),
        max_iter_(max_iter) { InitIds(cam_ids); }
  PTZRayOptimizer(const std::vector<ImageFeatures>& features, const std::vector<MatchesInfo>& matches_info, const std::vector<Camera>& cameras,
                  const std::unordered_set<long>& cam_ids, int max_iter, FACTOR_TYPE type)
      : cameras_(cameras), features_(features), matches_info_(matches_info), num_cams_(cameras.size()), type_(type), max_iter_(max_iter) { InitIds(cam_ids); }

  bool Solve(std::vector<Camera>& cameras) { std::vector<std::vector<Ray>> rays; return Solve(cameras, rays); }

  bool Solve(std::vector<Camera>& cameras, std::vector<std::vector<Ray>>& rays) {
    if (!CheckValid()) return false;
    if (!FindTracks()) return false;
    if (!tlw_given_) SetInitTransLocalToWorld();
    // candidate views get dense ids in ascending image id; one row per (track, candidate view): AddConstraints2d2d, .cc:799-848,
    // flattened on the device by ptztracks_flatten
    std::vector<long> view_of;
    std::unordered_map<long, int> dense;
    for (size_t i = 0; i < num_cams_; ++i) if (isCandidate((long)i)) { dense[(long)i] = (int)view_of.size(); view_of.push_back((long)i); }
    std::vector<double> intr, ext, weight, pxyz;
    std::vector<float> uv, puv;
    std::vector<int32_t> oview, otrack, pview, row_track;
    std::vector<int> track_ids;
    for (long id : view_of) {
      const std::vector<double> v = cameras_[id].ToVector();
      const double in[9] = {v[0], v[1], v[2], v[3], v[10], v[11], v[12], v[13], v[14]};
      intr.insert(intr.end(), in, in + 9);
      ext.insert(ext.end(), v.begin() + 4, v.begin() + 10);
    }
    {
      std::vector<uint8_t> cand(num_cams_, 0);
      std::vector<int64_t> kp_off(num_cams_ + 1, 0);
      for (size_t i = 0; i < num_cams_; ++i) { cand[i] = isCandidate((long)i) ? 1 : 0; kp_off[i + 1] = kp_off[i] + (int64_t)features_[i].keypoints.size(); }
      std::vector<float> kp(2 * (size_t)kp_off[num_cams_] + 2);
      for (size_t i = 0; i < num_cams_; ++i)
        for (size_t j = 0; j < features_[i].keypoints.size(); ++j) { kp[2 * (kp_off[i] + j)] = features_[i].keypoints[j].pt.x; kp[2 * (kp_off[i] + j) + 1] = features_[i].keypoints[j].pt.y; }
      ptztracks_views tv{(int32_t)num_cams_, cand.data(), kp_off.data(), kp.data()};
      ptztracks_result tr{};
      tr.num_tracks = (int32_t)flat_id_.size(); tr.num_elems = (int64_t)flat_img_.size();
      tr.track_id = flat_id_.data(); tr.track_offset = flat_off_.data(); tr.elem_img = flat_img_.data(); tr.elem_feat = flat_feat_.data();
      const size_t nt = std::max<size_t>(flat_id_.size(), 1), ne = std::max<size_t>(flat_img_.size(), 1);
      row_track.resize(nt); weight.resize(nt); uv.resize(2 * ne); oview.resize(ne); otrack.resize(ne);
      ptztracks_obs ob{};
      ob.cap_rows = (int64_t)nt; ob.cap_obs = (int64_t)ne;
      ob.row_track = row_track.data(); ob.track_weight = weight.data(); ob.obs_uv = uv.data(); ob.obs_view = oview.data(); ob.obs_track = otrack.data();
      last_status_ = ptztracks_flatten(&tr, &tv, &ob);
      if (last_status_ != PTZ_OK) return false;
      row_track.resize(ob.num_rows); weight.resize(ob.num_rows); uv.resize(2 * (size_t)ob.num_obs); oview.resize(ob.num_obs); otrack.resize(ob.num_obs);
      for (int32_t r : row_track) track_ids.push_back(flat_id_[r]);
    }
    for (long id : view_of)
      if (!pixels_.empty())
        for (size_t j = 0; j < pixels_[id].size(); ++j) {
          puv.push_back(pixels_[id][j].x); puv.push_back(pixels_[id][j].y); pview.push_back(dense[id]);
          pxyz.push_back(pts3d_[id][j].x); pxyz.push_back(pts3d_[id][j].y); pxyz.push_back(pts3d_[id][j].z);
        }
    ptzba_problem p{};
    p.factor_type = (int)type_;
    p.num_views = (int)view_of.size(); p.num_tracks = (int)track_ids.size(); p.num_obs = (int)oview.size(); p.num_pts3d = (int)pview.size();
    p.intr = intr.data(); p.ext = ext.data(); p.obs_uv = uv.data(); p.obs_view = oview.data(); p.obs_track = otrack.data(); p.track_weight = weight.data();
    p.pt_uv = puv.data(); p.pt_xyz = pxyz.data(); p.pt_view = pview.data(); p.tlw0 = tlw_param_.data();
    std::vector<int32_t> shared;  // SetSharedIntrinsics: the ids of the candidate views, in their dense order
    for (long id : view_of) shared.push_back((int32_t)shared_ic_ids_[id]);
    bool identity = true;
    for (size_t k = 0; k < view_of.size(); ++k) if (shared_ic_ids_[view_of[k]] != view_of[k]) identity = false;
    if (!identity) p.shared_ic_id = shared.data();
    ptz_solver_options o;
    ptz_solver_options_default(&o);
    o.max_num_iterations = max_iter_;
    std::vector<double> cams_w(21 * view_of.size()), rays_w(3 * std::max<size_t>(track_ids.size(), 1));
    ptzba_result r{};
    r.cams_world = cams_w.data(); r.rays_world = rays_w.data();
    last_status_ = ptzba_solve(&p, &o, &r);
    if (last_status_ != PTZ_OK) return false;
    init_reproj_error_all_ = r.init_reproj_error_all; final_reproj_error_all_ = r.final_reproj_error_all;
    final_reproj_error_2d2d_ = r.final_reproj_error_2d2d; final_reproj_error_2d3d_ = r.final_reproj_error_2d3d;
    num_iterations_ = r.num_iterations;
    if (r.termination != PTZ_CONVERGENCE) return false;  // .cc:482-487: outputs untouched
    for (size_t k = 0; k < view_of.size(); ++k) cameras[view_of[k]].FromKrt21(&cams_w[21 * k]);
    rays.clear();
    rays.resize(num_cams_);
    for (size_t t = 0; t < track_ids.size(); ++t) {
      const Vec3 rw{{rays_w[3 * t], rays_w[3 * t + 1], rays_w[3 * t + 2]}};
      for (const auto& it : tracks_.at(track_ids[t])) rays[it.first].emplace_back(track_ids[t], rw, features_[it.first].keypoints[it.second].pt);
    }
    return true;
  }

  double final_reproj_error_all() const { return final_reproj_error_all_; }
  double final_reproj_error_2d2d() const { return final_reproj_error_2d2d_; }
  double final_reproj_error_2d3d() const { return final_reproj_error_2d3d_; }
  void SetSharedIntrinsics(const std::vector<long>& shared_ic_ids) {
    if (shared_ic_ids.size() != cameras_.size()) return;  // .cc:499-502
    shared_ic_ids_ = shared_ic_ids;
  }
  // filled by the first Solve that sees it empty, reused by every later optimizer it is given to (same match table!)
  void SetTrackCache(TrackCache* cache) { track_cache_ = cache; }
  // false: keep the device's canonical track ids (skips the host pass over the matches; the tracks are the same sets)
  void SetReferenceTrackIds(bool on) { reference_track_ids_ = on; }
  void SetInitTransLocalToWorld(const double tlw[6]) { tlw_param_.assign(tlw, tlw + 6); tlw_given_ = true; }  // see the header comment
  // .cc:562-633: T_l_w from EPnP on the first annotated candidate view that passes the gates; zeros and false when none does
  bool SetInitTransLocalToWorld() {
    for (size_t i = 0; i < num_cams_; ++i) {
      if (!isCandidate((long)i) || pixels_.empty() || pixels_[i].empty()) continue;
      const size_t n = pts3d_[i].size();
      std::vector<double> obj(3 * n);
      std::vector<float> pix(2 * n);
      for (size_t j = 0; j < n; ++j) {
        obj[3 * j] = pts3d_[i][j].x; obj[3 * j + 1] = pts3d_[i][j].y; obj[3 * j + 2] = pts3d_[i][j].z;
        pix[2 * j] = pixels_[i][j].x; pix[2 * j + 1] = pixels_[i][j].y;
      }
      double tlw[6];  // EPnP, the gates of .cc:582-604 and T_l_w = T_i_l^-1 T_i_w: epnp::init_tlw_from_view (shared with ptzgeo_init_tlw)
      if (!epnp::init_tlw_from_view((int)n, obj.data(), pix.data(), cameras_[i].K().data(), cameras_[i].dist().data(), cameras_[i].R().data(),
                                    cameras_[i].t().data(), tlw))
        continue;
      tlw_param_.assign(tlw, tlw + 6);
      return true;
    }
    tlw_param_.assign(6, 0.0);
    return false;
  }
  static void T_l_w(const double* tlw, Mat33& R_l_w, Vec3& t_l_w) {                         // .cc:507-513
    ptz::rodrigues_jac(tlw, R_l_w.data(), nullptr);
    t_l_w = Vec3{{tlw[3], tlw[4], tlw[5]}};
  }
  const Tracks& tracks() const { return tracks_; }
  const std::vector<double>& tlw_init() const { return tlw_param_; }  // T_l_w the solve started from
  int num_iterations() const { return num_iterations_; }
  int last_status() const { return last_status_; }

 private:
  void InitIds(const std::unordered_set<long>& cam_ids) {
    if (cam_ids.empty()) for (size_t i = 0; i < cameras_.size(); ++i) cam_ids_.insert((long)i);
    else cam_ids_ = cam_ids;
    shared_ic_ids_.resize(cameras_.size());
    std::iota(shared_ic_ids_.begin(), shared_ic_ids_.end(), 0);
  }
  bool CheckValid() const {  // .cc:515-535
    if (num_cams_ == 0 || features_.size() != num_cams_ || max_iter_ <= 0) return false;
    if (!pixels_.empty()) {
      if (pixels_.size() != num_cams_ || pts3d_.size() != num_cams_) return false;
      for (size_t i = 0; i < num_cams_; ++i) if (pixels_[i].size() != pts3d_[i].size()) return false;
    }
    return true;
  }
  // .cc:537-552: TracksBuilder::Build / Filter(4) / ExportToSTL, on the device (ptztracks_build); the class above stays as the
  // host-side mirror of the reference's TracksBuilder API.  By default the device's canonical ids (smallest node) are turned into
  // the reference's union-by-rank root ids, so tracks(), Ray::id_ and the order of the residual blocks are the reference's.
  bool FindTracks() {
    if (track_cache_ && !track_cache_->flat_off.empty()) {  // built by an earlier optimizer over the same match table
      tracks_ = track_cache_->tracks; flat_id_ = track_cache_->flat_id; flat_off_ = track_cache_->flat_off;
      flat_img_ = track_cache_->flat_img; flat_feat_ = track_cache_->flat_feat;
      return true;
    }
    std::vector<int32_t> src, dst, q, t;
    std::vector<int64_t> off(1, 0);
    for (const auto& mi : matches_info_) {
      src.push_back((int32_t)mi.src_img_idx); dst.push_back((int32_t)mi.dst_img_idx);
      for (const auto& m : mi.matches) { q.push_back(m.queryIdx); t.push_back(m.trainIdx); }
      off.push_back((int64_t)q.size());
    }
    const size_t N = q.size();
    ptztracks_matches mm{(int32_t)src.size(), src.data(), dst.data(), off.data(), q.data(), t.data(), 4};
    flat_id_.assign(std::max<size_t>(N, 1), 0); flat_off_.assign(std::max<size_t>(N, 1) + 1, 0);
    flat_img_.assign(std::max<size_t>(2 * N, 1), 0); flat_feat_.assign(std::max<size_t>(2 * N, 1), 0);
    ptztracks_result tr{};
    tr.cap_tracks = (int64_t)N; tr.cap_elems = (int64_t)(2 * N);
    tr.track_id = flat_id_.data(); tr.track_offset = flat_off_.data(); tr.elem_img = flat_img_.data(); tr.elem_feat = flat_feat_.data();
    last_status_ = ptztracks_build(&mm, &tr);
    if (last_status_ != PTZ_OK) return false;
    if (reference_track_ids_) {  // ids and order of the reference's sequential UnionFind (host pass over the matches)
      last_status_ = ptztracks_reference_ids(&mm, &tr);
      if (last_status_ != PTZ_OK) return false;
    }
    flat_id_.resize(tr.num_tracks); flat_off_.resize((size_t)tr.num_tracks + 1); flat_img_.resize(tr.num_elems); flat_feat_.resize(tr.num_elems);
    tracks_.clear();
    for (int32_t k = 0; k < tr.num_tracks; ++k) {
      Track& trk = tracks_[flat_id_[k]];
      for (int64_t i = flat_off_[k]; i < flat_off_[k + 1]; ++i) trk.emplace_hint(trk.end(), flat_img_[i], flat_feat_[i]);
    }
    if (track_cache_) {
      track_cache_->tracks = tracks_; track_cache_->flat_id = flat_id_; track_cache_->flat_off = flat_off_;
      track_cache_->flat_img = flat_img_; track_cache_->flat_feat = flat_feat_;
    }
    return true;
  }
  bool isCandidate(long id) const { return cam_ids_.find(id) != cam_ids_.end(); }

  std::vector<Camera> cameras_;
  std::vector<ImageFeatures> features_;
  std::vector<MatchesInfo> matches_info_;
  std::vector<std::vector<Point2f>> pixels_;
  std::vector<std::vector<Point3d>> pts3d_;
  size_t num_cams_ = 0;
  std::unordered_set<long> cam_ids_;
  std::vector<long> shared_ic_ids_;
  FACTOR_TYPE type_;
  std::vector<double> tlw_param_ = std::vector<double>(6, 0.0);
  bool tlw_given_ = false, reference_track_ids_ = true;
  TrackCache* track_cache_ = nullptr;
  Tracks tracks_;
  std::vector<int32_t> flat_id_, flat_img_, flat_feat_;  // tracks_ as ptztracks_result arrays
  std::vector<int64_t> flat_off_;
  int max_iter_ = 100, num_iterations_ = 0, last_status_ = 0;
  double init_reproj_error_all_ = 0, final_reproj_error_all_ = 0, final_reproj_error_2d2d_ = 0, final_reproj_error_2d3d_ = 0;
};

class KRTOptimizer {
 public:
  enum FACTOR_TYPE { F, FDist, Fxfy, FxfyDist };  // krt_optimizer.h:110 (order differs from the C enum: mapped below)

  KRTOptimizer(int max_iter, double max_reproj_error, FACTOR_TYPE factor_type)
      : factor_type_(factor_type), max_iter_(max_iter), max_reproj_error_(max_reproj_error) {}

  void SetInitParams(const Mat33& K, const Mat33& R, const Vec3& t, const Vec5& dist) { cam_curr_world_ = Camera(K, R, t, dist); }  // .cc:257-263
  void Add2d2dConstraints(const Camera& cam_ref, const std::vector<KeyPoint>& kpts_ref, const std::vector<KeyPoint>& kpts_curr,
                          const std::vector<DMatch>& matches) {  // .cc:265-348
    cam_ref_ = cam_ref;
    uv1_.clear(); uv2_.clear();
    for (const auto& m : matches) {
      uv1_.push_back(kpts_ref[m.queryIdx].pt.x); uv1_.push_back(kpts_ref[m.queryIdx].pt.y);
      uv2_.push_back(kpts_curr[m.trainIdx].pt.x); uv2_.push_back(kpts_curr[m.trainIdx].pt.y);
    }
    double ref[21], init[21];
    cam_ref_.ToKrt21(ref);
    cam_curr_world_.ToKrt21(init);
    ptzreloc_local_params(ref, init, cam_curr_local_param_);  // .cc:269-286
  }
  // .cc:406-455: RMS of the 2d-2d functor residuals at cam_curr_local_param_ (initial before Solve, refined after), for ANY match
  // list against the camera passed to Add2d2dConstraints' frame
  double Cal2d2dReprojError(const Camera& cam_ref, const std::vector<KeyPoint>& kpts_ref, const std::vector<KeyPoint>& kpts_curr,
                            const std::vector<DMatch>& matches) {
    std::vector<float> u1, u2;
    for (const auto& m : matches) {
      u1.push_back(kpts_ref[m.queryIdx].pt.x); u1.push_back(kpts_ref[m.queryIdx].pt.y);
      u2.push_back(kpts_curr[m.trainIdx].pt.x); u2.push_back(kpts_curr[m.trainIdx].pt.y);
    }
    double ref[21], e22 = 0, e23 = 0;
    Camera local = cam_ref;  // cam_ref_local: K and dist of the reference, R = I, t = 0 (.cc:409-413)
    local.ToKrt21(ref);
    for (int i = 0; i < 9; ++i) ref[4 + i] = (i % 4 == 0) ? 1.0 : 0.0;
    ref[13] = ref[14] = ref[15] = 0.0;
    if (ptzreloc_reproj_error(kMap()[(int)factor_type_], ref, cam_curr_local_param_, (int)matches.size(), u1.data(), u2.data(), 0, nullptr, nullptr, &e22,
                              &e23) != PTZ_OK)
      return -1;
    return e22;
  }
  // .cc:457-500: world points go through R_local_world_, t_local_world_ = the reference camera's R, t
  double Cal2d3dReprojError(const std::vector<Point2f>& pts2d, const std::vector<Point3d>& pts3d) {
    if (pts2d.size() != pts3d.size() || pts2d.empty()) return -1;
    std::vector<float> pu;
    std::vector<double> px;
    for (size_t i = 0; i < pts2d.size(); ++i) {
      pu.push_back(pts2d[i].x); pu.push_back(pts2d[i].y);
      px.push_back(pts3d[i].x); px.push_back(pts3d[i].y); px.push_back(pts3d[i].z);
    }
    double ref[21], e22 = 0, e23 = 0;
    cam_ref_.ToKrt21(ref);
    if (ptzreloc_reproj_error(kMap()[(int)factor_type_], ref, cam_curr_local_param_, 0, nullptr, nullptr, (int)pts2d.size(), pu.data(), px.data(), &e22, &e23) !=
        PTZ_OK)
      return -1;
    return e23;
  }
  // .cc:350-383; the world->local transform of the points (.cc:357-362) is applied on the device with cam_ref's R, t
  void Add2d3dConstraints(const std::vector<Point2f>& pts2d, const std::vector<Point3d>& pts3d) {
    if (pts2d.size() != pts3d.size() || pts2d.empty()) return;
    for (size_t i = 0; i < pts2d.size(); ++i) {
      puv_.push_back(pts2d[i].x); puv_.push_back(pts2d[i].y);
      pxyz_.push_back(pts3d[i].x); pxyz_.push_back(pts3d[i].y); pxyz_.push_back(pts3d[i].z);
    }
  }
  bool Solve(Mat33& K, Mat33& R, Vec3& t, Vec5& dist) {  // .cc:385-404
    double ref[21], init[21], out[21];
    cam_ref_.ToKrt21(ref);
    cam_curr_world_.ToKrt21(init);
    const int64_t off[2] = {0, (int64_t)(uv1_.size() / 2)};
    ptzreloc_batch b{};
    b.factor_type = kMap()[(int)factor_type_]; b.num_queries = 1; b.match_offset = off; b.uv_ref = uv1_.data(); b.uv_cur = uv2_.data();
    b.ref_cam = ref; b.init_cam = init; b.max_iter = max_iter_; b.max_reproj_error = max_reproj_error_;
    const int64_t poff[2] = {0, (int64_t)(puv_.size() / 2)};
    if (!puv_.empty()) { b.pt_offset = poff; b.pt_uv = puv_.data(); b.pt_xyz = pxyz_.data(); }
    int32_t ok = 0, term = 0, nit = 0;
    ptzreloc_result r{};
    double local[15];
    r.cam = out; r.success = &ok; r.termination = &term; r.num_iter = &nit; r.local_cam15 = local;
    ptz_solver_options o;
    ptz_solver_options_default(&o);
    if (ptzreloc_solve_batch(&b, &o, &r) != PTZ_OK) return false;
    num_iter_ = nit;
    for (int j = 0; j < 15; ++j) cam_curr_local_param_[j] = local[j];  // Ceres refines the block in place, converged or not
    if (!ok) return false;  // CheckResults (.cc:504-533) ran on the device
    Camera c;
    c.FromKrt21(out);
    K = c.K(); R = c.R(); t = c.t(); dist = c.dist();
    return true;
  }
  void SetFixedFocal() { set_fixed_focal_ = true; }  // a flag nothing reads, as in the reference (.cc:502)
  int num_iter_ = 0;

 private:
  static const int* kMap() {  // krt_optimizer.h:110 order -> the C enum
    static const int m[4] = {PTZ_KRT_F, PTZ_KRT_FDIST, PTZ_KRT_FXFY, PTZ_KRT_FXFYDIST};
    return m;
  }
  double cam_curr_local_param_[15] = {0};
  Camera cam_curr_world_, cam_ref_;
  std::vector<float> uv1_, uv2_, puv_;
  std::vector<double> pxyz_;
  bool set_fixed_focal_ = false;
  FACTOR_TYPE factor_type_ = F;
  int max_iter_ = 100;
  double max_reproj_error_ = 50;
};

// ptz_incremental_optimizer.h:24-124, .cc:39-441 — PTZ-IBA: greedy incremental registration around global bundle adjustments.
// Same control flow, thresholds and state as the reference; the numerical work goes to the GPU: the tracks are built ONCE per run
// (ptztracks_build; they depend on the match table alone, the reference rebuilds them for each of its ~20-40 BAs), every global BA
// is then one ptztracks_flatten + ptzba_solve (through PTZRayOptimizer above), and RegisterNextImage solves ALL the
// candidate (registered neighbour, new image) KRT problems of one image in ONE ptzreloc_solve_batch call — one CTA each —
// then takes the first success in matches_info order, which is what the reference's sequential loop (.cc:384-415) returns.
class PtzIncrementalOptimizer {
 public:
  PtzIncrementalOptimizer(const std::vector<ImageFeatures>& features, const std::vector<MatchesInfo>& matches_info, const std::vector<Camera>& cameras,
                          int max_iter)
      : cameras_(cameras), features_(features), matches_info_(matches_info), max_iter_(max_iter) {}
  PtzIncrementalOptimizer(const std::vector<ImageFeatures>& features, const std::vector<MatchesInfo>& matches_info, const std::vector<Camera>& cameras,
                          const std::vector<std::string>& names, int max_iter)
      : cameras_(cameras), features_(features), matches_info_(matches_info), names_(names), max_iter_(max_iter) {}

  static long& kMaxNumImages() { static long v = 100000; return v; }            // .cc:24
  static float& kBaGlobalImagesRatio() { static float v = 1.1f; return v; }     // .cc:25

  void SetSeedImageId(const std::vector<long>& image_ids) { seed_image_ids_ = image_ids; }  // .cc:127-132

  bool Solve(std::vector<Camera>& cameras, std::unordered_set<long>& reg_image_ids) {  // .cc:39-125
    if (features_.empty() || features_.size() != cameras_.size() || max_iter_ <= 0) return false;  // CheckValid, .cc:134-140
    const int kInitNumTrials = 50;
    for (int trial = 0; trial < kInitNumTrials; ++trial) {
      long id1, id2;
      if (!FindInitialImagePair(id1, id2)) return false;
      Trace(kTraceSeedPair, id1, id2);
      const bool init_ok = RegisterInitialImagePair(id1, id2);
      Trace(kTraceInitResult, init_ok ? 1 : 0, last_ba_iterations_);
      if (!init_ok) continue;
      AdjustGlobalBundle();
      size_t ba_prev = reg_image_ids_.size();
      bool reg_next = true;
      while (reg_next) {
        reg_next = false;
        const std::vector<long> next = FindNextImages();
        Trace(kTraceNextList, (long)next.size(), next.empty() ? -1 : next[0]);
        if (next.empty()) break;
        for (size_t reg_trial = 0; reg_trial < next.size(); ++reg_trial) {
          const long image_id = next[reg_trial];
          reg_next = RegisterNextImage(image_id);
          if (reg_next && reg_image_ids_.size() >= kBaGlobalImagesRatio() * ba_prev) {
            if (AdjustGlobalBundle()) { ba_prev = reg_image_ids_.size(); break; }
            reg_image_ids_.erase(image_id);
            Trace(kTraceUnregister, image_id, 0);
            reg_next = false;
          }
          if (!reg_next) {
            // an initial pair that cannot grow is abandoned for another one
            const size_t kMinNumInitialRegTrials = 30, kMinModelSize = 3;
            if (reg_trial >= kMinNumInitialRegTrials && reg_image_ids_.size() < kMinModelSize) break;
          }
        }
      }
      AdjustGlobalBundle();
      reg_image_ids = reg_image_ids_;
      cameras = cameras_;
      return true;
    }
    return false;
  }

  // bookkeeping a caller may want to look at (not in the reference's public interface)
  // the decisions of Solve in the order they were taken, as (kind, a, b): seed pair (a, b); result of the two-view BA (ok, LM
  // iterations); list FindNextImages returned (length, first id); RegisterNextImage (image, registered neighbour it succeeded
  // from or -1); global BA (ok, registered images); image erased again after a failed global BA.  tests/ compare this log with
  // the CPU restatement of the reference's driver (oracle/iba_oracle.cpp), which numbers the kinds the same way.
  enum { kTraceSeedPair = 1, kTraceInitResult = 2, kTraceGlobalBa = 3, kTraceRegister = 4, kTraceUnregister = 5, kTraceNextList = 6 };
  const std::vector<std::array<long, 3>>& trace() const { return trace_; }
  int num_global_bundles() const { return num_global_bundles_; }
  int num_reloc_batches() const { return num_reloc_batches_; }
  int num_reloc_queries() const { return num_reloc_queries_; }
  double last_reproj_error() const { return last_reproj_error_; }

 private:
  bool isReg(long id) const { return reg_image_ids_.find(id) != reg_image_ids_.end(); }
  void Trace(long kind, long a, long b) { trace_.push_back(std::array<long, 3>{{kind, a, b}}); }
  static std::vector<long> RankedIds(const std::vector<float>& rank) {  // descending confidence, images without any dropped
    std::vector<long> idx(rank.size());
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](int A, int B) -> bool { return rank[A] > rank[B]; });
    std::vector<long> out;
    for (long i : idx) {
      if (rank[i] <= 0.0f) break;
      out.push_back(i);
    }
    return out;
  }
  long PairId(long a, long b) const { return a < b ? a * kMaxNumImages() + b : b * kMaxNumImages() + a; }  // .cc:306-312

  bool FindInitialImagePair(long& id1, long& id2) {  // .cc:142-177
    const std::vector<long> first = seed_image_ids_.empty() ? FindFirstInitialImage() : seed_image_ids_;
    for (long a : first)
      for (long b : FindSecondInitialImage(a)) {
        const long pid = PairId(a, b);
        if (init_image_pairs_.count(pid) > 0) continue;  // every pair is tried once
        init_image_pairs_.insert(pid);
        id1 = a; id2 = b;
        return true;
      }
    id1 = id2 = std::numeric_limits<long>::max();
    return false;
  }
  std::vector<long> FindFirstInitialImage() const {  // .cc:179-206
    std::vector<float> rank(features_.size(), 0.0f);
    // (float += double, as the reference: the sum is formed in double and rounded back)
    for (const auto& mi : matches_info_) { rank[mi.src_img_idx] += mi.confidence; rank[mi.dst_img_idx] += mi.confidence; }
    return RankedIds(rank);
  }
  std::vector<long> FindSecondInitialImage(long id1) const {  // .cc:208-247
    std::vector<float> rank(features_.size(), 0.0f);
    const float kMinPixelDiff = 50;
    for (const auto& mi : matches_info_) {
      const long s = mi.src_img_idx, d = mi.dst_img_idx;
      if (mi.matches.empty() || (id1 != s && id1 != d) || (id1 == s && id1 == d)) continue;
      if (CalPixelDiff(s, d, mi.matches) < kMinPixelDiff) continue;
      rank[id1 == s ? d : s] += mi.confidence;
    }
    return RankedIds(rank);
  }
  std::vector<long> FindNextImages() const {  // .cc:249-290
    std::vector<float> rank(features_.size(), 0.0f);
    const size_t kMaxRegTrials = 4;
    auto tired = [&](long id) { auto it = num_reg_trials_.find(id); return it != num_reg_trials_.end() && it->second > kMaxRegTrials; };
    for (const auto& mi : matches_info_) {
      const long s = mi.src_img_idx, d = mi.dst_img_idx;
      if (s == d || !mi.has_H || tired(s) || tired(d)) continue;
      const bool rs = isReg(s), rd = isReg(d);
      if (rs == rd) continue;  // both registered already, or neither a neighbour of the model
      rank[rs ? d : s] += mi.confidence;
    }
    return RankedIds(rank);
  }
  float CalPixelDiff(long id1, long id2, const std::vector<DMatch>& matches) const {  // .cc:292-304 (float accumulation as there)
    float total = 0.0f;
    for (const DMatch& m : matches) {
      const Point2f a = features_[id1].keypoints[m.queryIdx].pt, b = features_[id2].keypoints[m.trainIdx].pt;
      const float dx = a.x - b.x, dy = a.y - b.y;  // Point2f difference, then cv::norm in double, added into the float total
      total += std::sqrt((double)dx * dx + (double)dy * dy);
    }
    return total * 1.0f / matches.size();
  }
  // R_j = K_j^-1 H_ji K_i R_i (.cc:343-346, 391-393)
  static Mat33 RotationFromHomography(const Mat33& Kj, const Mat33& H, const Mat33& Ki, const Mat33& Ri) {
    Mat33 Kinv, a, b, c;
    ptz::inv3(Kj.data(), Kinv.data());
    ptz::mul33(Kinv.data(), H.data(), a.data());
    ptz::mul33(a.data(), Ki.data(), b.data());
    ptz::mul33(b.data(), Ri.data(), c.data());
    return c;
  }
  void SetInitialImagePairParameters(long id1, long id2) {  // .cc:314-350
    const double ratio = 1.2;  // sets the initial field of view
    for (long id : {id1, id2}) {
      const double focal = ratio * std::max(features_[id].img_size.width, features_[id].img_size.height);
      Mat33& K = cameras_[id].K();
      K[0] = K[4] = focal; K[2] = 0.5 * features_[id].img_size.width; K[5] = 0.5 * features_[id].img_size.height;
    }
    cameras_[id1].R() = Mat33{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
    for (const auto& mi : matches_info_)
      if (mi.src_img_idx == id1 && mi.dst_img_idx == id2) {
        cameras_[id2].R() = RotationFromHomography(cameras_[id2].K(), mi.H, cameras_[id1].K(), cameras_[id1].R());
        break;
      }
  }
  bool RegisterInitialImagePair(long id1, long id2) {  // .cc:352-375
    num_reg_trials_[id1] += 1; num_reg_trials_[id2] += 1;
    init_image_pairs_.insert(PairId(id1, id2));
    SetInitialImagePairParameters(id1, id2);
    PTZRayOptimizer optimizer(features_, matches_info_, cameras_, std::unordered_set<long>{id1, id2}, max_iter_, PTZRay);
    optimizer.SetTrackCache(&track_cache_);
    const bool ok = optimizer.Solve(cameras_);
    last_ba_iterations_ = optimizer.num_iterations();
    if (ok) { reg_image_ids_.insert(id1); reg_image_ids_.insert(id2); }
    return ok;
  }
  bool RegisterNextImage(long image_id) {  // .cc:377-419, batched
    num_reg_trials_[image_id] += 1;
    const long j = image_id;
    std::vector<const MatchesInfo*> cand;
    for (const auto& mi : matches_info_)
      if (mi.has_H && isReg(mi.src_img_idx) && mi.dst_img_idx == j) cand.push_back(&mi);
    if (cand.empty()) { Trace(kTraceRegister, j, -1); return false; }
    const int B = (int)cand.size();
    std::vector<int64_t> off(1, 0);
    std::vector<float> uv1, uv2;
    std::vector<double> ref(21 * (size_t)B), init(21 * (size_t)B), out(21 * (size_t)B);
    std::vector<Camera> init_cam(B);
    for (int q = 0; q < B; ++q) {
      const MatchesInfo& mi = *cand[q];
      const long i = mi.src_img_idx;
      Camera cj = cameras_[j];
      cj.K() = cameras_[i].K();
      cj.R() = RotationFromHomography(cj.K(), mi.H, cameras_[i].K(), cameras_[i].R());
      init_cam[q] = cj;
      cameras_[i].ToKrt21(&ref[21 * (size_t)q]);
      cj.ToKrt21(&init[21 * (size_t)q]);
      for (const DMatch& m : mi.matches) {
        const Point2f a = features_[i].keypoints[m.queryIdx].pt, b = features_[j].keypoints[m.trainIdx].pt;
        uv1.push_back(a.x); uv1.push_back(a.y); uv2.push_back(b.x); uv2.push_back(b.y);
      }
      off.push_back((int64_t)(uv1.size() / 2));
    }
    ptzreloc_batch b{};
    b.factor_type = PTZ_KRT_F; b.num_queries = B; b.match_offset = off.data(); b.uv_ref = uv1.data(); b.uv_cur = uv2.data();
    b.ref_cam = ref.data(); b.init_cam = init.data();
    b.max_iter = 100; b.max_reproj_error = 100;  // .cc:395-396
    std::vector<int32_t> ok(B, 0), term(B, 0), nit(B, 0);
    ptzreloc_result r{};
    r.cam = out.data(); r.success = ok.data(); r.termination = term.data(); r.num_iter = nit.data();
    ptz_solver_options o;
    ptz_solver_options_default(&o);
    ++num_reloc_batches_; num_reloc_queries_ += B;
    if (ptzreloc_solve_batch(&b, &o, &r) != PTZ_OK) { Trace(kTraceRegister, j, -1); return false; }
    for (int q = 0; q < B; ++q)
      if (ok[q]) {
        Camera c;
        c.FromKrt21(&out[21 * (size_t)q]);
        cameras_[j].K() = c.K();
        cameras_[j].R() = c.R();
        reg_image_ids_.insert(j);
        Trace(kTraceRegister, j, cand[q]->src_img_idx);
        return true;
      }
    Trace(kTraceRegister, j, -1);
    // every trial failed: the reference leaves the last trial's initial K, R in cameras_[j]
    cameras_[j].K() = init_cam[B - 1].K();
    cameras_[j].R() = init_cam[B - 1].R();
    return false;
  }
  bool AdjustGlobalBundle() {  // .cc:421-439
    PTZRayOptimizer optimizer(features_, matches_info_, cameras_, reg_image_ids_, max_iter_, PTZRay);
    optimizer.SetTrackCache(&track_cache_);  // the tracks are a function of the match table: built by the first BA of the run only
    const bool ok = optimizer.Solve(cameras_);
    last_reproj_error_ = optimizer.final_reproj_error_all();
    ++num_global_bundles_;
    Trace(kTraceGlobalBa, ok ? 1 : 0, (long)reg_image_ids_.size());
    return ok;
  }

  std::vector<Camera> cameras_;
  std::vector<ImageFeatures> features_;
  std::vector<MatchesInfo> matches_info_;
  std::vector<std::string> names_;
  int max_iter_;
  std::unordered_set<long> init_image_pairs_;            // image pairs already tried as the seed
  std::unordered_map<long, size_t> num_reg_trials_;      // registration attempts per image (bounded by kMaxRegTrials)
  std::unordered_set<long> reg_image_ids_;
  std::vector<long> seed_image_ids_;
  int num_global_bundles_ = 0, num_reloc_batches_ = 0, num_reloc_queries_ = 0, last_ba_iterations_ = 0;
  double last_reproj_error_ = 0;
  std::vector<std::array<long, 3>> trace_;
  TrackCache track_cache_;
};

}  // namespace ptzcalib
#endif  // PTZCALIB_B200_HPP
