/*
 * tracks_oracle.cpp — CPU restatement of the reference's track building.  TEST INFRASTRUCTURE ONLY (same rules as
 * ptz_oracle.cpp: loaded by tests/, smoke() and bench.py's cpu_baseline leg, never by the product).
 *
 * Follows, container for container and in the same order of operations (paths under /root/reference):
 *   - TracksBuilder::Build        src/core/tracks.cc:19-62    std::set of (image, feature), flat index by ascending pair,
 *                                                              one Union per match in matches_info order
 *   - UnionFind                   src/core/union_find.h:28-106 union by rank, Find with full path compression
 *   - TracksBuilder::Filter       src/core/tracks.cc:64-101   image listed twice / fewer than min_track_length images
 *   - TracksBuilder::ExportToSTL  src/core/tracks.cc:103-118  std::map<int, std::map<int,int>>, 1-node sets dropped
 *   - the residual-block loop of PTZRayOptimizer::AddConstraints2d2d, ptzray_optimizer.cc:801-848 (orc_tracks_flatten)
 *
 * Track ids here ARE the reference's (the union-by-rank root of each set); tracks come out in std::map order of those
 * ids.  The CUDA path labels a track by its smallest node instead (include/ptzcalib_b200.h); tests compare the two as
 * sets of tracks and re-label before comparing the flattened rows.
 *
 * PINNING STATUS: the reference ships no fixtures for this path.  The restatement is pinned by tests/test_tracks_oracle.py
 * against scipy.sparse.csgraph.connected_components + an independent numpy Filter on random match graphs, and against
 * hand-worked cases (collision, short track, chain across pairs).
 */
#include "../include/ptzcalib_b200.h"

#include <algorithm>
#include <climits>
#include <limits>
#include <map>
#include <numeric>
#include <set>
#include <utility>
#include <vector>

namespace {

struct UnionFind {  // union_find.h:28-106
  std::vector<int> parent, rank, size;
  void InitSets(int n) {
    size.assign(n, 1);
    parent.resize(n);
    std::iota(parent.begin(), parent.end(), 0);
    rank.assign(n, 0);
  }
  int Find(int i) {  // the recursion of union_find.h:52-62 unrolled: every node of the path ends up pointing at the root
    int root = i;
    while (parent[root] != root) root = parent[root];
    while (parent[i] != root) { const int next = parent[i]; parent[i] = root; i = next; }
    return root;
  }
  void Union(int i, int j) {  // union_find.h:66-92
    const int ri = Find(i), rj = Find(j);
    if (ri == rj) return;
    if (rank[ri] < rank[rj]) {
      parent[ri] = rj;
      size[rj] += size[ri];
    } else {
      parent[rj] = ri;
      size[ri] += size[rj];
      if (rank[ri] == rank[rj]) ++rank[ri];
    }
  }
};

typedef std::pair<int, int> Node;  // (image, feature), tracks.h:25
typedef std::map<int, int> Track;  // tracks.h:31
typedef std::map<int, Track> Tracks;

struct Builder {
  std::vector<Node> nodes;  // flat_pair_map after sort(): index = position (tracks.cc:35-43)
  UnionFind uf;
  int IndexOf(const Node& n) const { return (int)(std::lower_bound(nodes.begin(), nodes.end(), n) - nodes.begin()); }

  void Build(const ptztracks_matches* m) {  // tracks.cc:19-62
    std::set<Node> all;
    for (int k = 0; k < m->num_pairs; ++k)
      for (int64_t r = m->match_offset[k]; r < m->match_offset[k + 1]; ++r) {
        all.emplace(m->pair_src[k], m->query_idx[r]);
        all.emplace(m->pair_dst[k], m->train_idx[r]);
      }
    nodes.assign(all.begin(), all.end());
    uf.InitSets((int)nodes.size());
    for (int k = 0; k < m->num_pairs; ++k)
      for (int64_t r = m->match_offset[k]; r < m->match_offset[k + 1]; ++r)
        uf.Union(IndexOf(Node(m->pair_src[k], m->query_idx[r])), IndexOf(Node(m->pair_dst[k], m->train_idx[r])));
  }
  void Filter(int min_track_length) {  // tracks.cc:64-101
    std::map<int, std::set<int>> tracks;
    std::set<int> problematic;
    for (int k = 0; k < (int)nodes.size(); ++k) {
      const int id = uf.Find(k);
      if (!tracks[id].insert(nodes[k].first).second) problematic.insert(id);
    }
    for (const auto& t : tracks)
      if ((int)t.second.size() < min_track_length) problematic.insert(t.first);
    for (int& root : uf.parent)
      if (problematic.count(root) > 0) {
        uf.size[root] = 1;
        root = std::numeric_limits<int>::max();
      }
  }
  void Export(Tracks& out) const {  // tracks.cc:103-118
    out.clear();
    for (int k = 0; k < (int)nodes.size(); ++k) {
      const int id = uf.parent[k];
      if (id != std::numeric_limits<int>::max() && uf.size[id] > 1) out[id].insert(nodes[k]);
    }
  }
};

}  // namespace

extern "C" {

int orc_tracks_build(const ptztracks_matches* m, ptztracks_result* out) {
  if (!m || !out) return PTZ_ERR_INVALID;
  Builder b;
  b.Build(m);
  std::set<int> comps;
  {
    Builder probe = b;
    for (int k = 0; k < (int)probe.nodes.size(); ++k) comps.insert(probe.uf.Find(k));
  }
  b.Filter(m->min_track_length);
  Tracks tracks;
  b.Export(tracks);
  out->num_nodes = (int32_t)b.nodes.size();
  out->num_components = (int32_t)comps.size();
  out->num_tracks = (int32_t)tracks.size();
  int64_t ne = 0;
  for (const auto& t : tracks) ne += (int64_t)t.second.size();
  out->num_elems = ne;
  if ((int64_t)tracks.size() > out->cap_tracks || ne > out->cap_elems) return PTZ_ERR_INVALID;
  int64_t pos = 0;
  int ti = 0;
  for (const auto& t : tracks) {
    out->track_id[ti] = t.first;
    out->track_offset[ti] = pos;
    for (const auto& e : t.second) { out->elem_img[pos] = e.first; out->elem_feat[pos] = e.second; ++pos; }
    ++ti;
  }
  out->track_offset[ti] = pos;
  return PTZ_OK;
}

/* the loop of AddConstraints2d2d (ptzray_optimizer.cc:801-848) over tracks given in the caller's order */
int orc_tracks_flatten(const ptztracks_result* tr, const ptztracks_views* v, ptztracks_obs* out) {
  if (!tr || !v || !out) return PTZ_ERR_INVALID;
  std::vector<int> dense(v->num_images, -1);
  int nd = 0;
  for (int i = 0; i < v->num_images; ++i)
    if (v->is_candidate[i]) dense[i] = nd++;
  int rows = 0;
  int64_t nobs = 0;
  for (int t = 0; t < tr->num_tracks; ++t) {
    int row = -1;
    for (int64_t i = tr->track_offset[t]; i < tr->track_offset[t + 1]; ++i) {
      const int img = tr->elem_img[i], f = tr->elem_feat[i];
      if (img < 0 || img >= v->num_images || f < 0 || f >= v->kp_offset[img + 1] - v->kp_offset[img]) return PTZ_ERR_INVALID;
      if (!v->is_candidate[img]) continue;  // isCandidate, :814-815
      if (row < 0) {
        row = rows++;
        if (row < out->cap_rows) { out->row_track[row] = t; out->track_weight[row] = (double)(tr->track_offset[t + 1] - tr->track_offset[t]); }  // :805
      }
      if (nobs < out->cap_obs) {
        out->obs_uv[2 * nobs] = v->kp_uv[2 * (v->kp_offset[img] + f)];
        out->obs_uv[2 * nobs + 1] = v->kp_uv[2 * (v->kp_offset[img] + f) + 1];
        out->obs_view[nobs] = dense[img];
        out->obs_track[nobs] = row;
      }
      ++nobs;
    }
  }
  out->num_rows = rows;
  out->num_obs = (int32_t)nobs;
  return (rows > out->cap_rows || nobs > out->cap_obs) ? PTZ_ERR_INVALID : PTZ_OK;
}

}  // extern "C"
