// Host harness around ptz-calib_b200/csrc/ba_structure.hpp (pure C++): exposes the structure arrays to the CPU tests.
#include "../ptz-calib_b200/csrc/ba_structure.hpp"
#include <cstring>
static ptz::BaStructure g;
extern "C" {
int hs_build(int V, int P, int M, const float* uv, const int* view, const int* track, int chunk, const long long* extra, int nextra) {
  std::vector<int64_t> ex(extra, extra + nextra);
  ptz::build_structure(V, P, M, uv, view, track, chunk, g, nextra ? &ex : nullptr);
  return 0;
}
int hs_size(int which) {
  switch (which) {
    case 0: return g.nchunks(); case 1: return g.nub(); case 2: return g.nnzb(); case 3: return (int)g.pair_a.size();
  }
  return -1;
}
void hs_get(int which, int* out) {
  const std::vector<int>* v = nullptr;
  switch (which) {
    case 0: v = &g.perm; break; case 1: v = &g.o_view; break; case 2: v = &g.o_track; break; case 3: v = &g.view_off; break;
    case 4: v = &g.chunk_view; break; case 5: v = &g.chunk_begin; break; case 6: v = &g.chunk_cnt; break; case 7: v = &g.view_chunk_off; break;
    case 8: v = &g.t_off; break; case 9: v = &g.t_obs; break; case 10: v = &g.ub_row; break; case 11: v = &g.ub_col; break;
    case 12: v = &g.pair_a; break; case 13: v = &g.pair_b; break; case 14: v = &g.s_rowptr; break; case 15: v = &g.s_col; break;
    case 16: v = &g.diag_pos; break; case 17: v = &g.ub_pos; break; case 18: v = &g.ub_pos_t; break;
  }
  if (v && !v->empty()) std::memcpy(out, v->data(), v->size() * sizeof(int));
}
void hs_get_pair_off(long long* out) { for (size_t i = 0; i < g.ub_pair_off.size(); ++i) out[i] = g.ub_pair_off[i]; }
}
