// Drives the header-only C++ adaptor (include/ptzcalib_b200.hpp) the way the reference's callers drive its classes:
//   PTZRayOptimizer(features, matches_info, cameras, cam_ids, max_iter, type).Solve(cameras, rays)   (ptz_incremental_optimizer.cc:424-426)
//   KRTOptimizer(max_iter, thr, type); SetInitParams; Add2d2dConstraints; Solve                       (run_ptz_reloc.cc:94-108)
// Input / output are flat binary files written / read by tests/test_gpu_cpp_adaptor.py.
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
  auto hdr = rd<int>(f, 6);  // V, P, M, type, max_iter, N (reloc matches)
  const int V = hdr[0], M = hdr[2], type = hdr[3], max_iter = hdr[4], N = hdr[5];
  auto cams21 = rd<double>(f, 21 * (size_t)V);
  auto oview = rd<int>(f, M), otrack = rd<int>(f, M);
  auto ouv = rd<float>(f, 2 * (size_t)M);
  auto ref21 = rd<double>(f, 21), init21 = rd<double>(f, 21);
  auto uv1 = rd<float>(f, 2 * (size_t)N), uv2 = rd<float>(f, 2 * (size_t)N);
  fclose(f);

  // features: the observations of a view are its keypoints; matches chain the consecutive views of every track
  std::vector<Camera> cameras(V);
  std::vector<ImageFeatures> features(V);
  for (int i = 0; i < V; ++i) { cameras[i].FromKrt21(&cams21[21 * (size_t)i]); features[i].img_idx = i; features[i].img_size = Size{1920, 1080}; }
  std::map<int, std::vector<std::pair<int, int>>> by_track;  // track -> (view, feature id)
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

  PTZRayOptimizer ba(features, matches_info, cameras, std::unordered_set<long>(), max_iter, (FACTOR_TYPE)type);
  std::vector<std::vector<Ray>> rays;
  const bool ok = ba.Solve(cameras, rays);
  size_t nrays = 0;
  for (auto& r : rays) nrays += r.size();

  // an iteration cap of 1 must report failure and leave the cameras untouched (ptzray_optimizer.cc:482-487)
  std::vector<Camera> untouched(V);
  for (int i = 0; i < V; ++i) untouched[i].FromKrt21(&cams21[21 * (size_t)i]);
  PTZRayOptimizer capped(features, matches_info, untouched, std::unordered_set<long>(), 1, (FACTOR_TYPE)type);
  const bool ok_capped = capped.Solve(untouched);
  bool same = true;
  for (int i = 0; i < V; ++i) { double a[21]; untouched[i].ToKrt21(a); for (int j = 0; j < 21; ++j) same = same && a[j] == cams21[21 * (size_t)i + j]; }

  // reloc
  KRTOptimizer krt(200, 100.0, KRTOptimizer::F);
  Camera cref, cinit;
  cref.FromKrt21(ref21.data()); cinit.FromKrt21(init21.data());
  krt.SetInitParams(cinit.K(), cinit.R(), cinit.t(), cinit.dist());
  std::vector<KeyPoint> k1(N), k2(N);
  std::vector<DMatch> mm(N);
  for (int i = 0; i < N; ++i) { k1[i].pt.x = uv1[2 * i]; k1[i].pt.y = uv1[2 * i + 1]; k2[i].pt.x = uv2[2 * i]; k2[i].pt.y = uv2[2 * i + 1]; mm[i].queryIdx = i; mm[i].trainIdx = i; }
  krt.Add2d2dConstraints(cref, k1, k2, mm);
  Mat33 K, R; Vec3 t; Vec5 dist;
  const double err_before = krt.Cal2d2dReprojError(cref, k1, k2, mm);  // krt_optimizer.cc:406-455, at the initial parameters
  const bool ok_krt = krt.Solve(K, R, t, dist);
  const double err_after = krt.Cal2d2dReprojError(cref, k1, k2, mm);   // ... and at the refined ones
  const double err_pts = krt.Cal2d3dReprojError(std::vector<Point2f>(), std::vector<Point3d>());  // -1 without points (:459-460)

  FILE* g = fopen(argv[2], "wb");
  double head[10] = {(double)ok, (double)ba.num_iterations(), ba.final_reproj_error_all(), ba.final_reproj_error_2d2d(), (double)ba.tracks().size(), (double)nrays,
                     (double)ok_capped, (double)same, (double)ok_krt, (double)krt.num_iter_};
  fwrite(head, sizeof(double), 10, g);
  for (int i = 0; i < V; ++i) { double a[21]; cameras[i].ToKrt21(a); fwrite(a, sizeof(double), 21, g); }
  Camera out(K, R, t, dist);
  double a[21]; out.ToKrt21(a); fwrite(a, sizeof(double), 21, g);
  double errs[3] = {err_before, err_after, err_pts};
  fwrite(errs, sizeof(double), 3, g);
  fclose(g);
  printf("adaptor_check: ba ok=%d iters=%d err=%.6f tracks=%zu rays=%zu | capped ok=%d untouched=%d | krt ok=%d iters=%d fx=%.4f\n", (int)ok, ba.num_iterations(),
         ba.final_reproj_error_all(), ba.tracks().size(), nrays, (int)ok_capped, (int)same, (int)ok_krt, krt.num_iter_, K[0]);
  return 0;
}
