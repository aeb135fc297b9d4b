// Georeferencing through the adaptor as run_ptz_ba.cc:142-145 does it: PTZRayOptimizer(features, matches_info, cameras, pixels, pts3d,
// cam_ids, max_iter, type).Solve(cameras) — T_l_w initialised by EPnP inside Solve (SetInitTransLocalToWorld, ptzray_optimizer.cc:562-633).
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
  auto hdr = rd<int>(f, 5);  // V, M, A, type, max_iter
  const int V = hdr[0], M = hdr[1], A = hdr[2], type = hdr[3], max_iter = hdr[4];
  auto cams21 = rd<double>(f, 21 * (size_t)V);
  auto oview = rd<int>(f, M), otrack = rd<int>(f, M);
  auto ouv = rd<float>(f, 2 * (size_t)M);
  auto pview = rd<int>(f, A);
  auto puv = rd<float>(f, 2 * (size_t)A);
  auto pxyz = rd<double>(f, 3 * (size_t)A);
  fclose(f);

  std::vector<Camera> cameras(V);
  std::vector<ImageFeatures> features(V);
  for (int i = 0; i < V; ++i) { cameras[i].FromKrt21(&cams21[21 * (size_t)i]); features[i].img_idx = i; features[i].img_size = Size{1920, 1080}; }
  std::map<int, std::vector<std::pair<int, int>>> by_track;
  for (int k = 0; k < M; ++k) {
    KeyPoint kp; kp.pt.x = ouv[2 * (size_t)k]; kp.pt.y = ouv[2 * (size_t)k + 1];
    features[oview[k]].keypoints.push_back(kp);
    by_track[otrack[k]].push_back({oview[k], (int)features[oview[k]].keypoints.size() - 1});
  }
  std::map<std::pair<int, int>, MatchesInfo> pairs;
  for (auto& t : by_track) {
    std::sort(t.second.begin(), t.second.end());
    for (size_t i = 0; i + 1 < t.second.size(); ++i) {
      auto key = std::make_pair(t.second[i].first, t.second[i + 1].first);
      MatchesInfo& mi = pairs[key];
      mi.src_img_idx = key.first; mi.dst_img_idx = key.second;
      DMatch m; m.queryIdx = t.second[i].second; m.trainIdx = t.second[i + 1].second;
      mi.matches.push_back(m);
    }
  }
  std::vector<MatchesInfo> matches_info;
  for (auto& p : pairs) matches_info.push_back(p.second);
  std::vector<std::vector<Point2f>> pixels(V);
  std::vector<std::vector<Point3d>> pts3d(V);
  for (int a = 0; a < A; ++a) {
    Point2f p; p.x = puv[2 * (size_t)a]; p.y = puv[2 * (size_t)a + 1];
    Point3d q; q.x = pxyz[3 * (size_t)a]; q.y = pxyz[3 * (size_t)a + 1]; q.z = pxyz[3 * (size_t)a + 2];
    pixels[pview[a]].push_back(p); pts3d[pview[a]].push_back(q);
  }

  PTZRayOptimizer ba(features, matches_info, cameras, pixels, pts3d, std::unordered_set<long>(), max_iter, (FACTOR_TYPE)type);
  const bool ok = ba.Solve(cameras);
  FILE* g = fopen(argv[2], "wb");
  if (!g) return 2;
  double head[12] = {(double)ok, (double)ba.num_iterations(), ba.final_reproj_error_all(), ba.final_reproj_error_2d2d(), ba.final_reproj_error_2d3d(), 0};
  for (int i = 0; i < 6; ++i) head[6 + i] = ba.tlw_init()[i];
  fwrite(head, sizeof(double), 12, g);
  std::vector<double> cw(21 * (size_t)V);
  for (int i = 0; i < V; ++i) cameras[i].ToKrt21(&cw[21 * (size_t)i]);
  fwrite(cw.data(), sizeof(double), cw.size(), g);
  fclose(g);
  printf("georef ok=%d it=%d err all %.4f 2d2d %.4f 2d3d %.4f  tlw0 = %.4f %.4f %.4f | %.3f %.3f %.3f\n", (int)ok, ba.num_iterations(), ba.final_reproj_error_all(),
         ba.final_reproj_error_2d2d(), ba.final_reproj_error_2d3d(), head[6], head[7], head[8], head[9], head[10], head[11]);
  return 0;
}
