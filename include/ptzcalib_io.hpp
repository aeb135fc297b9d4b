// ptzcalib_io.hpp — the reference's on-disk formats for the types of ptzcalib_b200.hpp (host-side C++17, header only).
//
// Mirrors the plain-data part of src/core/data_io.h:
//   ReadColmapFeatures   data_io.cc:24-52    "<n> <dim>" then n lines "x y scale orientation d_0 .. d_dim-1"
//   ReadColmapMatches    data_io.cc:64-108   blocks "<name1> <name2>" + lines "i j", separated by empty lines
//   SaveToJson           data_io.cc:110-164  {"cameras": {<rootname>: {name,pos,res,K,R,t,dist,distType,marker{pix,pos},version}}}
//   ReadFromJson         data_io.cc:180-247  (marker pixels are stored divided by the resolution)
//   ReadCamFromJson      data_io.cc:249-292
//   FindImgIndex         data_io.cc:460-474 (by root name, extension ignored)
//   FindBestMatch        run_ptz_reloc.cc:147-166 (reference image with the most matches for a query image)
// JSON is read and written by a small parser here (the reference uses nlohmann::ordered_json, an un-vendored dependency).
// Not built: LoadImgsAndFeatures (cv::imread on every image) and LoadMatchesInfo's cv::findHomography(RANSAC); callers that
// already hold homographies fill MatchesInfo::H / has_H themselves (tests/cpp/iba_check.cpp does).
#ifndef PTZCALIB_IO_HPP
#define PTZCALIB_IO_HPP

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "ptzcalib_b200.hpp"

namespace ptzcalib {

// ---------------------------------------------------------------------------------------------------------- COLMAP text files
inline void ReadColmapFeatures(const std::string& filepath, std::vector<KeyPoint>& kpts, std::vector<float>& desc, int& desc_dim) {
  kpts.clear();
  desc.clear();
  desc_dim = 0;
  std::ifstream fin(filepath);
  if (!fin.good()) return;
  int num = 0, dim = 0;
  if (!(fin >> num >> dim) || num < 0 || dim < 0) return;
  kpts.resize(num);
  desc.assign((size_t)num * dim, 0.0f);
  desc_dim = dim;
  for (int i = 0; i < num; ++i) {
    float scale, orientation;
    if (!(fin >> kpts[i].pt.x >> kpts[i].pt.y >> scale >> orientation)) { kpts.clear(); desc.clear(); desc_dim = 0; return; }
    for (int j = 0; j < dim; ++j)
      if (!(fin >> desc[(size_t)i * dim + j])) { kpts.clear(); desc.clear(); desc_dim = 0; return; }
  }
}

inline void ReadColmapMatches(const std::string& filepath, std::vector<std::vector<DMatch>>& pairs_matches,
                              std::vector<std::pair<std::string, std::string>>& img_pairs_name) {
  pairs_matches.clear();
  img_pairs_name.clear();
  auto is_image = [](const std::string& s) {
    for (const char* e : {".png", ".jpg", ".jpeg"}) {
      const std::string ext(e);
      if (s.size() >= ext.size() && s.compare(s.size() - ext.size(), ext.size(), ext) == 0) return true;
    }
    return false;
  };
  std::ifstream fin(filepath);
  std::string line;
  std::vector<DMatch> matches;
  std::pair<std::string, std::string> names;
  while (std::getline(fin, line)) {
    if (line.empty()) {  // a block ends at an empty line; blocks without matches are dropped (data_io.cc:77-87)
      if (!matches.empty()) { pairs_matches.push_back(matches); img_pairs_name.push_back(names); matches.clear(); names = {}; }
      continue;
    }
    std::istringstream iss(line);
    std::string a, b;
    iss >> a >> b;
    if (is_image(a)) names = {a, b};
    else { DMatch m; m.queryIdx = std::atoi(a.c_str()); m.trainIdx = std::atoi(b.c_str()); matches.push_back(m); }
  }
}

// root name: the file name without its extension (utils/os_path splitext: the last '.' of the last path component, not a leading one)
inline std::string NameWithoutExt(const std::string& fname) {
  const size_t slash = fname.find_last_of('/');
  const size_t base = slash == std::string::npos ? 0 : slash + 1;
  const size_t dot = fname.find_last_of('.');
  if (dot == std::string::npos || dot <= base) return fname;
  return fname.substr(0, dot);
}
// data_io.cc:460-474: images are matched by ROOT name -- JSON camera keys carry no extension, match files may name "a.png"
// where the feature file was "a.jpg"
inline long FindImgIndex(const std::vector<std::string>& fnames, const std::string& fname) {
  const std::string want = NameWithoutExt(fname);
  for (size_t i = 0; i < fnames.size(); ++i) if (NameWithoutExt(fnames[i]) == want) return (long)i;
  return -1;
}

// run_ptz_reloc.cc:147-166: among the pairs whose SECOND image is `fname`, the one with the most matches (first wins ties);
// an empty name / list when there is none
typedef std::pair<std::string, std::vector<DMatch>> BestMatchT;
inline BestMatchT FindBestMatch(const std::string& fname, const std::vector<std::pair<std::string, std::string>>& img_pairs_name,
                                const std::vector<std::vector<DMatch>>& pairs_matches) {
  BestMatchT best;
  if (img_pairs_name.size() != pairs_matches.size()) return best;
  for (size_t i = 0; i < img_pairs_name.size(); ++i) {
    if (img_pairs_name[i].second != fname) continue;
    if (pairs_matches[i].size() > best.second.size()) best = {img_pairs_name[i].first, pairs_matches[i]};
  }
  return best;
}

// ---------------------------------------------------------------------------------------------------------- a small JSON value
namespace json {
struct Value {
  enum Kind { Null, Number, String, Array, Object } kind = Null;
  double num = 0;
  std::string str;
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;  // insertion order kept, as nlohmann::ordered_json
  const Value* find(const std::string& k) const { for (const auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
  const Value& at(const std::string& k) const { const Value* v = find(k); if (!v) throw std::runtime_error("json: missing key " + k); return *v; }
  std::vector<double> numbers() const { std::vector<double> v; for (const auto& e : arr) v.push_back(e.num); return v; }
};
struct Parser {
  const std::string& s;
  size_t i = 0;
  explicit Parser(const std::string& text) : s(text) {}
  void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
  bool lit(const char* w) { size_t n = std::string(w).size(); if (s.compare(i, n, w) == 0) { i += n; return true; } return false; }
  Value parse() {
    ws();
    if (i >= s.size()) throw std::runtime_error("json: unexpected end");
    Value v;
    const char c = s[i];
    if (c == '{') {
      v.kind = Value::Object; ++i; ws();
      if (s[i] == '}') { ++i; return v; }
      for (;;) {
        ws();
        Value k = parse();
        if (k.kind != Value::String) throw std::runtime_error("json: key is not a string");
        ws();
        if (s[i++] != ':') throw std::runtime_error("json: ':' expected");
        v.obj.emplace_back(k.str, parse());
        ws();
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == '}') { ++i; return v; }
        throw std::runtime_error("json: ',' or '}' expected");
      }
    }
    if (c == '[') {
      v.kind = Value::Array; ++i; ws();
      if (s[i] == ']') { ++i; return v; }
      for (;;) {
        v.arr.push_back(parse());
        ws();
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == ']') { ++i; return v; }
        throw std::runtime_error("json: ',' or ']' expected");
      }
    }
    if (c == '"') {
      v.kind = Value::String; ++i;
      while (i < s.size() && s[i] != '"') {
        if (s[i] == '\\' && i + 1 < s.size()) { ++i; const char e = s[i]; v.str.push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e); }
        else v.str.push_back(s[i]);
        ++i;
      }
      ++i;
      return v;
    }
    if (lit("null")) return v;
    if (lit("true")) { v.kind = Value::Number; v.num = 1; return v; }
    if (lit("false")) { v.kind = Value::Number; v.num = 0; return v; }
    char* end = nullptr;
    v.num = std::strtod(s.c_str() + i, &end);
    if (end == s.c_str() + i) throw std::runtime_error("json: value expected");
    v.kind = Value::Number;
    i = (size_t)(end - s.c_str());
    return v;
  }
};
inline std::string num(double x) { char b[40]; std::snprintf(b, sizeof(b), "%.17g", x); return b; }
inline std::string list(const std::vector<double>& v) { std::string s = "["; for (size_t i = 0; i < v.size(); ++i) s += (i ? ", " : "") + num(v[i]); return s + "]"; }
}  // namespace json

// ---------------------------------------------------------------------------------------------------------- camera JSON
inline bool SaveToJson(const std::vector<Camera>& cameras, const std::vector<std::string>& names, const std::vector<std::vector<Point2f>>& pixels_gt,
                       const std::vector<std::vector<Point3d>>& pts3d_gt, const std::string& filepath) {
  std::ofstream fout(filepath);
  if (!fout.good()) return false;
  fout << "{\n    \"cameras\": {\n";
  for (size_t i = 0; i < cameras.size(); ++i) {
    const Camera& c = cameras[i];
    const size_t dot = names[i].find_last_of('.');
    const std::string root = dot == std::string::npos ? names[i] : names[i].substr(0, dot);
    // t_wc = -R^-1 t = -R^T t for a rotation (Camera::t_wc, types.h:93)
    std::vector<double> pos(3);
    for (int r = 0; r < 3; ++r) pos[r] = -(c.R()[r] * c.t()[0] + c.R()[3 + r] * c.t()[1] + c.R()[6 + r] * c.t()[2]);
    const int width = (int)(2 * c.K()[2]), height = (int)(2 * c.K()[5]);
    const std::vector<double> K(c.K().begin(), c.K().end()), R(c.R().begin(), c.R().end()), t(c.t().begin(), c.t().end()), d(c.dist().begin(), c.dist().end());
    const std::string in = "            ";
    fout << "        \"" << root << "\": {\n";
    fout << in << "\"name\": \"" << root << "\",\n" << in << "\"pos\": " << json::list(pos) << ",\n";
    fout << in << "\"res\": [" << width << ", " << height << "],\n";
    fout << in << "\"K\": " << json::list(K) << ",\n" << in << "\"R\": " << json::list(R) << ",\n" << in << "\"t\": " << json::list(t) << ",\n";
    fout << in << "\"dist\": " << json::list(d) << ",\n" << in << "\"distType\": \"" << (d[0] < 1e-5 ? "" : "k1") << "\",\n";  // data_io.cc:143-146
    fout << in << "\"marker\": {\n" << in << "    \"pix\": [";
    const bool have = i < pixels_gt.size() && i < pts3d_gt.size();
    if (have)
      for (size_t k = 0; k < pixels_gt[i].size(); ++k)
        fout << (k ? ", " : "") << "[" << json::num((double)(pixels_gt[i][k].x / width)) << ", " << json::num((double)(pixels_gt[i][k].y / height)) << "]";
    fout << "],\n" << in << "    \"pos\": [";
    if (have)
      for (size_t k = 0; k < pts3d_gt[i].size(); ++k)
        fout << (k ? ", " : "") << "[" << json::num(pts3d_gt[i][k].x) << ", " << json::num(pts3d_gt[i][k].y) << ", " << json::num(pts3d_gt[i][k].z) << "]";
    fout << "]\n" << in << "},\n" << in << "\"version\": \"2.0\"\n";
    fout << "        }" << (i + 1 < cameras.size() ? "," : "") << "\n";
  }
  fout << "    }\n}\n";
  return fout.good();
}

namespace detail {
inline bool load_json(const std::string& filepath, json::Value& root) {
  std::ifstream fin(filepath);
  if (!fin.good()) return false;
  std::stringstream ss;
  ss << fin.rdbuf();
  const std::string text = ss.str();
  try {
    json::Parser p(text);
    root = p.parse();
  } catch (const std::exception&) {
    return false;
  }
  return root.kind == json::Value::Object;
}
inline void camera_from(const json::Value& v, Camera& cam) {
  const std::vector<double> K = v.at("K").numbers(), R = v.at("R").numbers(), t = v.at("t").numbers(), d = v.at("dist").numbers();
  for (size_t i = 0; i < 9 && i < K.size(); ++i) cam.K()[i] = K[i];
  for (size_t i = 0; i < 9 && i < R.size(); ++i) cam.R()[i] = R[i];
  for (size_t i = 0; i < 3 && i < t.size(); ++i) cam.t()[i] = t[i];
  for (size_t i = 0; i < 5 && i < d.size(); ++i) cam.dist()[i] = d[i];
}
}  // namespace detail

inline bool ReadFromJson(const std::string& filepath, std::vector<Camera>& cameras, std::vector<std::string>& names, std::vector<std::vector<Point2f>>& pixels,
                         std::vector<std::vector<Point3d>>& pts3d, std::vector<Size>& sizes) {
  cameras.clear(); names.clear(); pixels.clear(); pts3d.clear(); sizes.clear();
  json::Value root;
  if (!detail::load_json(filepath, root)) return false;
  try {
    // the reference parses with nlohmann::json (std::map): cameras come out sorted by key, whatever the order in the file
    std::vector<std::pair<std::string, json::Value>> sorted_cams = root.at("cameras").obj;
    std::stable_sort(sorted_cams.begin(), sorted_cams.end(), [](const std::pair<std::string, json::Value>& a, const std::pair<std::string, json::Value>& b) { return a.first < b.first; });
    for (const auto& el : sorted_cams) {
      Camera cam;
      detail::camera_from(el.second, cam);
      Size size;
      size.width = (int)el.second.at("res").arr.at(0).num;
      size.height = (int)el.second.at("res").arr.at(1).num;
      std::vector<Point2f> pix;
      std::vector<Point3d> pos;
      for (const auto& p : el.second.at("marker").at("pix").arr) {  // stored divided by the resolution (data_io.cc:223-229)
        Point2f q; q.x = (float)(size.width * p.arr.at(0).num); q.y = (float)(size.height * p.arr.at(1).num);
        pix.push_back(q);
      }
      for (const auto& p : el.second.at("marker").at("pos").arr) { Point3d q; q.x = p.arr.at(0).num; q.y = p.arr.at(1).num; q.z = p.arr.at(2).num; pos.push_back(q); }
      names.push_back(el.first); pixels.push_back(pix); pts3d.push_back(pos); cameras.push_back(cam); sizes.push_back(size);
    }
  } catch (const std::exception&) {
    return false;
  }
  return true;
}

inline bool ReadCamFromJson(const std::string& filepath, const std::vector<std::string>& names, std::vector<Camera>& cameras) {
  cameras.assign(names.size(), Camera());
  json::Value root;
  if (!detail::load_json(filepath, root)) return false;
  try {
    const json::Value& cams = root.at("cameras");
    for (size_t i = 0; i < names.size(); ++i) {
      const size_t dot = names[i].find_last_of('.');
      const std::string rootname = dot == std::string::npos ? names[i] : names[i].substr(0, dot);
      const json::Value* v = cams.find(rootname);
      if (!v) return false;  // "Cannot find camera parameters in json file" (data_io.cc:279-282)
      detail::camera_from(*v, cameras[i]);
    }
  } catch (const std::exception&) {
    return false;
  }
  return true;
}

}  // namespace ptzcalib
#endif  // PTZCALIB_IO_HPP
