// Drives PtzIncrementalOptimizer (include/ptzcalib_b200.hpp, mirror of src/core/ptz_incremental_optimizer.h) the way
// run_ptz_ba.cc does: features + a table of pairwise matches with homographies and confidences, cameras unknown.
// Input: a synthetic scene (ground-truth cameras, observations by view/track); output: registered ids and refined cameras.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "../../include/ptzcalib_b200.hpp"

using namespace ptzcalib;

template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  auto hdr = rd<int>(f, 4);  // V, M, max_iter, both_directions
  const int V = hdr[0], M = hdr[1], max_iter = hdr[2], both = hdr[3];
  auto gt21 = rd<double>(f, 21 * (size_t)V);
  auto oview = rd<int>(f, M), otrack = rd<int>(f, M);
  auto ouv = rd<float>(f, 2 * (size_t)M);
  fclose(f);

  std::vector<Camera> gt(V), cameras(V);  // `cameras` start as default objects: the driver initialises what it registers
  std::vector<ImageFeatures> features(V);
  for (int i = 0; i < V; ++i) { gt[i].FromKrt21(&gt21[21 * (size_t)i]); features[i].img_idx = i; features[i].img_size = Size{1920, 1080}; }
  std::map<int, std::vector<std::pair<int, int>>> by_track;  // track -> (view, feature id)
  for (int k = 0; k < M; ++k) {
    KeyPoint kp; kp.pt.x = ouv[2 * (size_t)k]; kp.pt.y = ouv[2 * (size_t)k + 1];
    features[oview[k]].keypoints.push_back(kp);
    by_track[otrack[k]].push_back({oview[k], (int)features[oview[k]].keypoints.size() - 1});
  }
  // every pair of a track's observations is a match (exhaustive matching), stored under (src, dst) and, if asked, (dst, src)
  std::map<std::pair<int, int>, MatchesInfo> pairs;
  for (auto& t : by_track) {
    std::sort(t.second.begin(), t.second.end());
    for (size_t a = 0; a < t.second.size(); ++a)
      for (size_t b = a + 1; b < t.second.size(); ++b) {
        for (int dir = 0; dir < (both ? 2 : 1); ++dir) {
          const auto& s = dir ? t.second[b] : t.second[a];
          const auto& d = dir ? t.second[a] : t.second[b];
          MatchesInfo& mi = pairs[std::make_pair(s.first, d.first)];
          mi.src_img_idx = s.first; mi.dst_img_idx = d.first;
          DMatch m; m.queryIdx = s.second; m.trainIdx = d.second;
          mi.matches.push_back(m);
        }
      }
  }
  std::vector<MatchesInfo> matches_info;
  for (auto& p : pairs) {
    MatchesInfo mi = p.second;
    if (mi.matches.size() < 8) continue;  // pairs a matcher would not have kept
    // H_j_i = K_j R_j R_i^T K_i^-1 (what cv::findHomography estimates for a rotating camera; data_io.cc:339-354)
    const Camera &ci = gt[mi.src_img_idx], &cj = gt[mi.dst_img_idx];
    Mat33 Rit, Kinv, a, b, c;
    for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) Rit[3 * r + q] = ci.R()[3 * q + r];
    ptz::inv3(ci.K().data(), Kinv.data());
    ptz::mul33(cj.K().data(), cj.R().data(), a.data());
    ptz::mul33(a.data(), Rit.data(), b.data());
    ptz::mul33(b.data(), Kinv.data(), c.data());
    mi.H = c; mi.has_H = true;
    mi.num_inliers = (int)mi.matches.size();
    mi.confidence = mi.matches.size() >= 100 ? 1.0 : mi.matches.size() / 100.0;  // CalMatchingScore, data_io.cc:356-364
    matches_info.push_back(mi);
  }

  PtzIncrementalOptimizer iba(features, matches_info, cameras, max_iter);
  std::unordered_set<long> reg;
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = iba.Solve(cameras, reg);
  const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  FILE* g = fopen(argv[2], "wb");
  if (!g) return 2;
  double head[8] = {(double)ok, (double)reg.size(), (double)iba.num_global_bundles(), (double)iba.num_reloc_batches(), (double)iba.num_reloc_queries(),
                    iba.last_reproj_error(), (double)matches_info.size(), seconds};
  fwrite(head, sizeof(double), 8, g);
  std::vector<double> out(22 * (size_t)V);
  for (int i = 0; i < V; ++i) { out[22 * (size_t)i] = reg.count(i) ? 1.0 : 0.0; cameras[i].ToKrt21(&out[22 * (size_t)i + 1]); }
  fwrite(out.data(), sizeof(double), out.size(), g);
  fclose(g);
  size_t nm = 0;
  for (auto& mi : matches_info) nm += mi.matches.size();
  printf("iba ok=%d registered=%zu/%d global BAs=%d reloc batches=%d (%d queries) reproj=%.4f pairs=%zu matches=%zu  %.3f s\n", (int)ok, reg.size(), V,
         iba.num_global_bundles(), iba.num_reloc_batches(), iba.num_reloc_queries(), iba.last_reproj_error(), matches_info.size(), nm, seconds);
  return 0;
}
