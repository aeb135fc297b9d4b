// ba_structure.hpp — host-side, one-off per solve: turns the flat observation list of a PTZ-BA problem into the
// orderings and the block-sparsity pattern the kernels stream over.  Pure C++ (unit-tested on the build host).
//
//   * observations in VIEW-MAJOR order (sorted by view, then track): a CTA works on one view's slice, so the view's
//     rotation/intrinsics table is CTA-uniform and per-view blocks reduce over contiguous segments
//   * fixed-size CHUNKS that never straddle a view (one CTA each, deterministic two-level reductions)
//   * a by-track index list (each ray touches its L observations through it)
//   * the reduced camera system S: block CSR over views (both triangles stored), and for every upper off-diagonal
//     block the list of observation PAIRS (o in row view, o' in column view, same track) whose outer products make it
//
// This is the structure Ceres derives inside SchurEliminator/BlockSparseMatrix from the residual blocks the reference
// adds at ptzray_optimizer.cc:811-848; it is static across LM iterations.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace ptz {

struct BaStructure {
  int V = 0, P = 0, M = 0, chunk = 128;
  std::vector<int> perm;  // view-major position -> caller's observation index
  std::vector<int> o_view, o_track;
  std::vector<float> o_uv;  // 2*M
  std::vector<int> view_off;  // V+1
  std::vector<int> chunk_view, chunk_begin, chunk_cnt, view_chunk_off;  // view_chunk_off: V+1
  std::vector<int> t_off, t_obs;  // P+1, M
  // reduced camera system pattern
  std::vector<int> ub_row, ub_col;     // upper off-diagonal blocks, sorted by (row, col), row < col
  std::vector<int64_t> ub_pair_off;    // nub+1
  std::vector<int> pair_a, pair_b;     // view-major observation positions: a in the row view, b in the column view
  std::vector<int> s_rowptr, s_col;    // full block CSR (diagonal included), columns ascending
  std::vector<int> diag_pos, ub_pos, ub_pos_t;  // CSR slot of (c,c), (row,col), (col,row)
  int nub() const { return (int)ub_row.size(); }
  int nnzb() const { return (int)s_col.size(); }
  int nchunks() const { return (int)chunk_view.size(); }
};

inline int64_t ub_key(int r, int c) { return ((int64_t)r << 32) | (uint32_t)c; }

// extra_upper_keys: sorted unique (row<col) block keys to include even when no local observation pair touches them
// (multi-GPU: every rank must hold the same pattern so that the camera blocks can be all-reduced in place).
inline void build_structure(int V, int P, int M, const float* uv, const int32_t* view, const int32_t* track, int chunk, BaStructure& s,
                            const std::vector<int64_t>* extra_upper_keys = nullptr) {
  s.V = V; s.P = P; s.M = M; s.chunk = chunk;
  // ---- view-major order: counting sort by track, then stable counting sort by view
  std::vector<int> by_track(M), cnt(std::max(V, P) + 1, 0);
  for (int k = 0; k < M; ++k) ++cnt[track[k] + 1];
  for (int p = 0; p < P; ++p) cnt[p + 1] += cnt[p];
  {
    std::vector<int> pos(cnt.begin(), cnt.begin() + P + 1);
    for (int k = 0; k < M; ++k) by_track[pos[track[k]]++] = k;
  }
  s.view_off.assign(V + 1, 0);
  for (int k = 0; k < M; ++k) ++s.view_off[view[k] + 1];
  for (int v = 0; v < V; ++v) s.view_off[v + 1] += s.view_off[v];
  s.perm.resize(M);
  {
    std::vector<int> pos(s.view_off.begin(), s.view_off.end() - 1);
    for (int i = 0; i < M; ++i) { int k = by_track[i]; s.perm[pos[view[k]]++] = k; }
  }
  s.o_view.resize(M); s.o_track.resize(M); s.o_uv.resize(2 * (size_t)M);
  for (int i = 0; i < M; ++i) {
    int k = s.perm[i];
    s.o_view[i] = view[k]; s.o_track[i] = track[k];
    s.o_uv[2 * (size_t)i] = uv[2 * (size_t)k]; s.o_uv[2 * (size_t)i + 1] = uv[2 * (size_t)k + 1];
  }
  // ---- chunks
  s.chunk_view.clear(); s.chunk_begin.clear(); s.chunk_cnt.clear();
  s.view_chunk_off.assign(V + 1, 0);
  for (int v = 0; v < V; ++v) {
    for (int b = s.view_off[v]; b < s.view_off[v + 1]; b += chunk) {
      s.chunk_view.push_back(v); s.chunk_begin.push_back(b); s.chunk_cnt.push_back(std::min(chunk, s.view_off[v + 1] - b));
    }
    s.view_chunk_off[v + 1] = (int)s.chunk_view.size();
  }
  // ---- by-track lists (positions in view-major order, ascending => ascending view within a track)
  s.t_off.assign(P + 1, 0);
  for (int i = 0; i < M; ++i) ++s.t_off[s.o_track[i] + 1];
  for (int p = 0; p < P; ++p) s.t_off[p + 1] += s.t_off[p];
  s.t_obs.resize(M);
  {
    std::vector<int> pos(s.t_off.begin(), s.t_off.end() - 1);
    for (int i = 0; i < M; ++i) s.t_obs[pos[s.o_track[i]]++] = i;
  }
  // ---- observation pairs of every track, keyed by (row view < col view)
  int64_t npairs = 0;
  for (int p = 0; p < P; ++p) { int64_t L = s.t_off[p + 1] - s.t_off[p]; npairs += L * (L - 1) / 2; }
  std::vector<int> pa(npairs), pb(npairs), pr(npairs), pc(npairs);
  {
    int64_t q = 0;
    for (int p = 0; p < P; ++p)
      for (int i = s.t_off[p]; i < s.t_off[p + 1]; ++i)
        for (int j = i + 1; j < s.t_off[p + 1]; ++j) {
          int a = s.t_obs[i], b = s.t_obs[j];
          int va = s.o_view[a], vb = s.o_view[b];
          if (va == vb) continue;  // a track never lists an image twice (tracks.cc:63-82); tolerate it anyway
          pa[q] = a; pb[q] = b; pr[q] = va; pc[q] = vb; ++q;
        }
    npairs = q;
  }
  // stable counting sort by column then by row => sorted by (row, col)
  std::vector<int64_t> ord(npairs), ord2(npairs);
  {
    std::vector<int64_t> c2(V + 1, 0);
    for (int64_t q = 0; q < npairs; ++q) ++c2[pc[q] + 1];
    for (int v = 0; v < V; ++v) c2[v + 1] += c2[v];
    for (int64_t q = 0; q < npairs; ++q) ord[c2[pc[q]]++] = q;
    std::fill(c2.begin(), c2.end(), 0);
    for (int64_t q = 0; q < npairs; ++q) ++c2[pr[q] + 1];
    for (int v = 0; v < V; ++v) c2[v + 1] += c2[v];
    for (int64_t i = 0; i < npairs; ++i) { int64_t q = ord[i]; ord2[c2[pr[q]]++] = q; }
  }
  // merge the local block keys with the extra (global) ones
  std::vector<int64_t> keys;
  keys.reserve(npairs / 8 + 16);
  for (int64_t i = 0; i < npairs; ++i) {
    int64_t k = ub_key(pr[ord2[i]], pc[ord2[i]]);
    if (keys.empty() || keys.back() != k) keys.push_back(k);
  }
  if (extra_upper_keys && !extra_upper_keys->empty()) {
    std::vector<int64_t> merged(keys.size() + extra_upper_keys->size());
    auto e = std::set_union(keys.begin(), keys.end(), extra_upper_keys->begin(), extra_upper_keys->end(), merged.begin());
    merged.resize(e - merged.begin());
    keys.swap(merged);
  }
  const int nub = (int)keys.size();
  s.ub_row.resize(nub); s.ub_col.resize(nub); s.ub_pair_off.assign(nub + 1, 0);
  for (int b = 0; b < nub; ++b) { s.ub_row[b] = (int)(keys[b] >> 32); s.ub_col[b] = (int)(keys[b] & 0xffffffff); }
  s.pair_a.resize(npairs); s.pair_b.resize(npairs);
  {
    // pairs are sorted by key and `keys` is a sorted superset: walk both, count per block, prefix-sum
    int b = 0;
    for (int64_t i = 0; i < npairs; ++i) {
      int64_t q = ord2[i];
      int64_t k = ub_key(pr[q], pc[q]);
      while (keys[b] != k) ++b;
      ++s.ub_pair_off[b + 1];
      s.pair_a[i] = pa[q]; s.pair_b[i] = pb[q];
    }
    for (int bb = 0; bb < nub; ++bb) s.ub_pair_off[bb + 1] += s.ub_pair_off[bb];
  }
  // ---- full block CSR
  s.s_rowptr.assign(V + 1, 0);
  for (int v = 0; v < V; ++v) ++s.s_rowptr[v + 1];
  for (int b = 0; b < nub; ++b) { ++s.s_rowptr[s.ub_row[b] + 1]; ++s.s_rowptr[s.ub_col[b] + 1]; }
  for (int v = 0; v < V; ++v) s.s_rowptr[v + 1] += s.s_rowptr[v];
  s.s_col.assign(s.s_rowptr[V], 0);
  s.diag_pos.assign(V, 0); s.ub_pos.assign(nub, 0); s.ub_pos_t.assign(nub, 0);
  {
    // columns of a row ascend if we insert: lower entries (from blocks with col == row, ascending in row index),
    // then the diagonal, then upper entries (ascending col).  Blocks are sorted by (row, col), so one pass per kind works.
    std::vector<int> pos(s.s_rowptr.begin(), s.s_rowptr.end() - 1);
    for (int b = 0; b < nub; ++b) { int r = s.ub_col[b]; s.ub_pos_t[b] = pos[r]; s.s_col[pos[r]++] = s.ub_row[b]; }  // (col,row): lower
    for (int v = 0; v < V; ++v) { s.diag_pos[v] = pos[v]; s.s_col[pos[v]++] = v; }
    for (int b = 0; b < nub; ++b) { int r = s.ub_row[b]; s.ub_pos[b] = pos[r]; s.s_col[pos[r]++] = s.ub_col[b]; }
  }
}

}  // namespace ptz
