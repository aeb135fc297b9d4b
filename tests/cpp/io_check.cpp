// Reads the reference's text formats with include/ptzcalib_io.hpp and echoes what it found; round-trips the camera JSON.
#include <cstdio>
#include "../../include/ptzcalib_io.hpp"
using namespace ptzcalib;
int main(int argc, char** argv) {
  if (argc < 5) return 2;
  std::vector<KeyPoint> kp; std::vector<float> desc; int dim = 0;
  ReadColmapFeatures(argv[1], kp, desc, dim);
  printf("features %zu %d", kp.size(), dim);
  for (size_t i = 0; i < kp.size(); ++i) printf(" %.9g %.9g", kp[i].pt.x, kp[i].pt.y);
  if (!desc.empty()) printf(" d %.9g %.9g", desc[0], desc.back());
  printf("\n");
  std::vector<std::vector<DMatch>> pm; std::vector<std::pair<std::string, std::string>> names;
  ReadColmapMatches(argv[2], pm, names);
  printf("pairs %zu", pm.size());
  for (size_t i = 0; i < pm.size(); ++i) { printf(" | %s %s %zu", names[i].first.c_str(), names[i].second.c_str(), pm[i].size()); for (auto& m : pm[i]) printf(" %d:%d", m.queryIdx, m.trainIdx); }
  printf("\n");
  std::vector<Camera> cams; std::vector<std::string> cn; std::vector<std::vector<Point2f>> pix; std::vector<std::vector<Point3d>> pts; std::vector<Size> sizes;
  const bool ok = ReadFromJson(argv[3], cams, cn, pix, pts, sizes);
  printf("json %d %zu\n", (int)ok, cams.size());
  if (!ok) return 1;
  std::vector<std::string> files;
  for (auto& n : cn) files.push_back(n + ".jpg");
  if (!SaveToJson(cams, files, pix, pts, argv[4])) return 1;
  std::vector<Camera> again;
  const bool ok2 = ReadCamFromJson(argv[4], files, again);
  double worst = 0;
  for (size_t i = 0; i < cams.size() && ok2; ++i) {
    for (int k = 0; k < 9; ++k) { worst = std::max(worst, std::fabs(cams[i].K()[k] - again[i].K()[k])); worst = std::max(worst, std::fabs(cams[i].R()[k] - again[i].R()[k])); }
    for (int k = 0; k < 5; ++k) worst = std::max(worst, std::fabs(cams[i].dist()[k] - again[i].dist()[k]));
  }
  // Length / FindMaxCoVisible (tracks.cc:120-202): two components {0,1,2} and {5,6}; image 9 unseen
  Tracks tr;
  tr[0] = Track{{0, 1}, {1, 4}}; tr[3] = Track{{1, 2}, {2, 7}, {0, 3}}; tr[8] = Track{{5, 1}, {6, 1}};
  int tot, mx, mn; Length(tr, tot, mx, mn);
  std::set<int> co; FindMaxCoVisible(tr, 10, co);
  printf("tracks %d %d %d covis", tot, mx, mn);
  for (int i : co) printf(" %d", i);
  printf("\n");
  // FindImgIndex compares root names (data_io.cc:460-474); FindBestMatch takes the pair with the most matches whose second image is the query
  std::vector<std::string> fn{"img001.jpg", "dir.v2/img002.png", "img003"};
  printf("find %ld %ld %ld %ld %ld\n", FindImgIndex(fn, "img001"), FindImgIndex(fn, "img001.png"), FindImgIndex(fn, "dir.v2/img002"), FindImgIndex(fn, "img003.jpg"),
         FindImgIndex(fn, "img004.jpg"));
  BestMatchT bm = FindBestMatch("a.jpg", names, pm), none_bm = FindBestMatch("zzz.jpg", names, pm);
  printf("best %s %zu | %s %zu | order", bm.first.c_str(), bm.second.size(), none_bm.first.c_str(), none_bm.second.size());
  for (auto& n : cn) printf(" %s", n.c_str());
  printf("\n");
  std::vector<std::string> missing{"nope.jpg"};
  std::vector<Camera> none;
  printf("roundtrip %d %.3g missing %d pix0 %.9g %.9g size %d %d\n", (int)ok2, worst, (int)ReadCamFromJson(argv[4], missing, none), pix[0].empty() ? -1.0 : pix[0][0].x,
         pix[0].empty() ? -1.0 : pix[0][0].y, sizes[0].width, sizes[0].height);
  return 0;
}
