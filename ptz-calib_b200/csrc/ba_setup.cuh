// ba_setup.cuh — device-side construction of the orderings and the block pattern of ba_structure.hpp.
//
// Same outputs, array for array and in the same order, as the host builder (which stays as the unit-tested
// reference and as a fallback selectable with ptz_solver_options.verbose & 4): radix sorts (CUB), prefix scans and
// binary-search kernels.  At cfg 4 (2e6 observations, 4.2e6 observation pairs) this replaces ~0.3 s of single-threaded
// host work per solve by a few milliseconds on the device; it is set-up plumbing, not one of the LM stage kernels.
#pragma once
#include <cub/cub.cuh>

#include "ba_structure.hpp"
#include "common.cuh"

namespace ptz {

struct DevStructure {
  int V = 0, P = 0, M = 0, nchunks = 0, nub = 0, nnzb = 0;
  int64_t npairs = 0;
  DevBuf<float2> o_uv;
  DevBuf<int> perm, o_view, o_track, view_off, chunk_view, chunk_begin, chunk_cnt, view_chunk_off, t_off, t_obs;
  DevBuf<int64_t> pair_off;
  DevBuf<int> pair_a, pair_b, s_rowptr, s_col, diag_pos, ub_pos, ub_pos_t, blk_row, view_active;
  DevBuf<unsigned long long> ub_keys;  // sorted unique (row << 32 | col)
  std::vector<int> h_rowptr;           // host copy (CG shared-memory sizing)
  // between the two build phases: sorted pair keys and the local unique block keys
  DevBuf<unsigned long long> pk_sorted, uniq_local;
  int n_local = 0;
  std::vector<int64_t> local_keys_host(cudaStream_t s) const {
    std::vector<int64_t> k(n_local);
    if (n_local) { PTZ_CUDA(cudaMemcpyAsync(k.data(), uniq_local.p, (size_t)n_local * 8, cudaMemcpyDeviceToHost, s)); PTZ_CUDA(cudaStreamSynchronize(s)); }
    return k;
  }
};

namespace setup {

__global__ void k_make_obs_keys(int M, int tbits, const int* __restrict__ view, const int* __restrict__ track, unsigned long long* __restrict__ keys,
                                int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  keys[i] = ((unsigned long long)(unsigned)view[i] << tbits) | (unsigned)track[i];
  idx[i] = i;
}
__global__ void k_gather_obs(int M, int tbits, const unsigned long long* __restrict__ keys, const int* __restrict__ perm, const float2* __restrict__ uv_in,
                             float2* __restrict__ o_uv, int* __restrict__ o_view, int* __restrict__ o_track, int* __restrict__ pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const unsigned long long k = keys[i];
  o_view[i] = (int)(k >> tbits);
  o_track[i] = (int)(k & ((1ull << tbits) - 1ull));
  o_uv[i] = uv_in[perm[i]];
  pos[i] = i;
}
// off[v] = first index i with sorted[i] >= v, for v = 0..n (off[n] = M)
__global__ void k_lower_bounds_int(int n, int M, const int* __restrict__ sorted, int* __restrict__ off) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > n) return;
  int lo = 0, hi = M;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted[mid] < v) lo = mid + 1; else hi = mid; }
  off[v] = lo;
}
__global__ void k_lower_bounds_u64(int n, long long M, const unsigned long long* __restrict__ sorted, const unsigned long long* __restrict__ probe,
                                   long long* __restrict__ off) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n) return;
  if (b == n) { off[n] = M; return; }
  const unsigned long long key = probe[b];
  long long lo = 0, hi = M;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (sorted[mid] < key) lo = mid + 1; else hi = mid; }
  off[b] = lo;
}
__global__ void k_chunk_counts(int V, int chunk, const int* __restrict__ view_off, int* __restrict__ nch, int* __restrict__ active) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int c = view_off[v + 1] - view_off[v];
  nch[v] = (c + chunk - 1) / chunk;
  active[v] = c > 0 ? 1 : 0;
}
__global__ void k_fill_chunks(int V, int chunk, const int* __restrict__ view_off, const int* __restrict__ view_chunk_off, int* __restrict__ chunk_view,
                              int* __restrict__ chunk_begin, int* __restrict__ chunk_cnt) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int c = view_chunk_off[v];
  for (int b = view_off[v]; b < view_off[v + 1]; b += chunk, ++c) {
    chunk_view[c] = v; chunk_begin[c] = b; chunk_cnt[c] = min(chunk, view_off[v + 1] - b);
  }
}
__global__ void k_pair_counts(int P, const int* __restrict__ t_off, long long* __restrict__ cnt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long L = t_off[p + 1] - t_off[p];
  cnt[p] = L * (L - 1) / 2;
}
__global__ void k_make_pairs(int P, const int* __restrict__ t_off, const int* __restrict__ t_obs, const int* __restrict__ o_view,
                             const long long* __restrict__ base, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ payload) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  long long q = base[p];
  for (int i = t_off[p]; i < t_off[p + 1]; ++i)
    for (int j = i + 1; j < t_off[p + 1]; ++j) {
      const int a = t_obs[i], b = t_obs[j];
      keys[q] = ((unsigned long long)(unsigned)o_view[a] << 32) | (unsigned)o_view[b];
      payload[q] = ((unsigned long long)(unsigned)b << 32) | (unsigned)a;
      ++q;
    }
}
__global__ void k_split_pairs(long long n, const unsigned long long* __restrict__ payload, int* __restrict__ pa, int* __restrict__ pb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pa[i] = (int)(payload[i] & 0xffffffffull);
  pb[i] = (int)(payload[i] >> 32);
}
// head flags of the sorted keys -> compaction handled by cub::DeviceSelect::Unique
__global__ void k_swap_keys(int nub, const unsigned long long* __restrict__ keys, unsigned long long* __restrict__ keys2, int* __restrict__ idx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nub) return;
  keys2[b] = (keys[b] << 32) | (keys[b] >> 32);  // (col << 32 | row)
  idx[b] = b;
}
__global__ void k_row_bounds(int V, int nub, const unsigned long long* __restrict__ keys /* row-major sorted */,
                             const unsigned long long* __restrict__ keys2 /* col-major sorted */, int* __restrict__ first_ub, int* __restrict__ first_lb) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > V) return;
  const unsigned long long probe = (unsigned long long)(unsigned)v << 32;
  int lo = 0, hi = nub;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < probe) lo = mid + 1; else hi = mid; }
  first_ub[v] = lo;
  lo = 0; hi = nub;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys2[mid] < probe) lo = mid + 1; else hi = mid; }
  first_lb[v] = lo;
}
__global__ void k_row_len(int V, const int* __restrict__ first_ub, const int* __restrict__ first_lb, int* __restrict__ len) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V) len[v] = (first_lb[v + 1] - first_lb[v]) + 1 + (first_ub[v + 1] - first_ub[v]);
}
__global__ void k_fill_csr(int V, int nub, const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ keys2, const int* __restrict__ idx2,
                           const int* __restrict__ first_ub, const int* __restrict__ first_lb, const int* __restrict__ rowptr, int* __restrict__ s_col,
                           int* __restrict__ blk_row, int* __restrict__ diag_pos, int* __restrict__ ub_pos, int* __restrict__ ub_pos_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V) {
    const int pos = rowptr[i] + (first_lb[i + 1] - first_lb[i]);
    diag_pos[i] = pos; s_col[pos] = i; blk_row[pos] = i;
  }
  if (i < nub) {
    // upper entry of block i = (r, c)
    const int r = (int)(keys[i] >> 32), c = (int)(keys[i] & 0xffffffffull);
    const int pos = rowptr[r] + (first_lb[r + 1] - first_lb[r]) + 1 + (i - first_ub[r]);
    ub_pos[i] = pos; s_col[pos] = c; blk_row[pos] = r;
    // lower entry: position i of the col-major order is block idx2[i] = (r2, c2), stored in row c2 among its lower entries
    const int c2 = (int)(keys2[i] >> 32), r2 = (int)(keys2[i] & 0xffffffffull);
    const int pos2 = rowptr[c2] + (i - first_lb[c2]);
    ub_pos_t[idx2[i]] = pos2; s_col[pos2] = r2; blk_row[pos2] = c2;
  }
}

struct CubTemp {
  DevBuf<char> buf;
  cudaStream_t s = nullptr;
  void* get(size_t bytes) {
    if (bytes > buf.n) buf.alloc(bytes + bytes / 4 + 256, s);
    return buf.p;
  }
};

template <class K, class Vv>
inline void sort_pairs(CubTemp& tmp, K* kin, K* kout, Vv* vin, Vv* vout, long long n, int begin_bit, int end_bit, cudaStream_t s) {
  size_t bytes = 0;
  PTZ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, begin_bit, end_bit, s));
  void* t = tmp.get(bytes);
  PTZ_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, kin, kout, vin, vout, n, begin_bit, end_bit, s));
}
template <class T>
inline void exclusive_scan(CubTemp& tmp, const T* in, T* out, int n, cudaStream_t s) {
  size_t bytes = 0;
  PTZ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
  void* t = tmp.get(bytes);
  PTZ_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, n, s));
}
inline int bits_for(long long n) { int b = 1; while ((1ll << b) < n) ++b; return b; }

}  // namespace setup

// phase A: orderings, chunks, by-track lists, sorted observation pairs, local block keys
inline void build_structure_device_obs(int V, int P, int M, const float* h_uv, const int32_t* h_view, const int32_t* h_track, int chunk, DevStructure& d,
                                       cudaStream_t s) {
  using namespace setup;
  d.V = V; d.P = P; d.M = M;
  CubTemp tmp;
  tmp.s = s;
  auto grid = [](long long n, int b) { return (unsigned)std::max<long long>((n + b - 1) / b, 1); };
  // ---- raw inputs
  DevBuf<float2> uv_in;
  DevBuf<int> view_in, track_in, idx0, idx1, pos0;
  DevBuf<unsigned long long> k0, k1;
  const int Mx = std::max(M, 1);
  uv_in.upload(reinterpret_cast<const float2*>(h_uv), M, s);
  view_in.upload(h_view, M, s);
  track_in.upload(h_track, M, s);
  k0.alloc(Mx, s); k1.alloc(Mx, s); idx0.alloc(Mx, s); idx1.alloc(Mx, s); pos0.alloc(Mx, s);
  d.o_uv.alloc(Mx, s); d.o_view.alloc(Mx, s); d.o_track.alloc(Mx, s); d.perm.alloc(Mx, s);
  d.view_off.alloc(V + 1, s); d.view_chunk_off.alloc(V + 1, s); d.t_off.alloc(P + 1, s); d.t_obs.alloc(Mx, s); d.view_active.alloc(V, s);
  const int tbits = bits_for(std::max(P, 2)), vbits = bits_for(std::max(V, 2));
  // ---- view-major order: sort by (view, track)
  if (M > 0) {
    k_make_obs_keys<<<grid(M, 256), 256, 0, s>>>(M, tbits, view_in.p, track_in.p, k0.p, idx0.p);
    sort_pairs(tmp, k0.p, k1.p, idx0.p, d.perm.p, M, 0, tbits + vbits, s);
    k_gather_obs<<<grid(M, 256), 256, 0, s>>>(M, tbits, k1.p, d.perm.p, uv_in.p, d.o_uv.p, d.o_view.p, d.o_track.p, pos0.p);
  }
  k_lower_bounds_int<<<grid(V + 1, 256), 256, 0, s>>>(V, M, d.o_view.p, d.view_off.p);
  // ---- chunks
  DevBuf<int> nch;
  nch.alloc(V + 1, s);
  nch.zero(s);
  k_chunk_counts<<<grid(V, 256), 256, 0, s>>>(V, chunk, d.view_off.p, nch.p, d.view_active.p);
  exclusive_scan(tmp, nch.p, d.view_chunk_off.p, V + 1, s);
  // ---- by-track lists: stable sort of positions by track
  DevBuf<int> tk0, tk1;
  tk0.alloc(Mx, s); tk1.alloc(Mx, s);
  if (M > 0) {
    PTZ_CUDA(cudaMemcpyAsync(tk0.p, d.o_track.p, (size_t)M * 4, cudaMemcpyDeviceToDevice, s));
    sort_pairs(tmp, tk0.p, tk1.p, pos0.p, d.t_obs.p, M, 0, tbits, s);
  }
  k_lower_bounds_int<<<grid(P + 1, 256), 256, 0, s>>>(P, M, tk1.p, d.t_off.p);
  // ---- observation pairs
  DevBuf<long long> pcnt, pbase;
  pcnt.alloc(P + 1, s); pbase.alloc(P + 1, s);
  pcnt.zero(s);
  if (P > 0) k_pair_counts<<<grid(P, 256), 256, 0, s>>>(P, d.t_off.p, pcnt.p);
  exclusive_scan(tmp, pcnt.p, pbase.p, P + 1, s);
  int h_nchunks = 0;
  long long h_npairs = 0;
  PTZ_CUDA(cudaMemcpyAsync(&h_nchunks, d.view_chunk_off.p + V, 4, cudaMemcpyDeviceToHost, s));
  PTZ_CUDA(cudaMemcpyAsync(&h_npairs, pbase.p + P, 8, cudaMemcpyDeviceToHost, s));
  PTZ_CUDA(cudaStreamSynchronize(s));
  d.nchunks = h_nchunks; d.npairs = h_npairs;
  d.chunk_view.alloc(std::max(h_nchunks, 1), s); d.chunk_begin.alloc(std::max(h_nchunks, 1), s); d.chunk_cnt.alloc(std::max(h_nchunks, 1), s);
  k_fill_chunks<<<grid(V, 128), 128, 0, s>>>(V, chunk, d.view_off.p, d.view_chunk_off.p, d.chunk_view.p, d.chunk_begin.p, d.chunk_cnt.p);
  const long long NPx = std::max<long long>(h_npairs, 1);
  DevBuf<unsigned long long> pk0, pv0, pv1;
  DevBuf<unsigned long long>& pk1 = d.pk_sorted;
  DevBuf<unsigned long long>& uniq = d.uniq_local;
  pk0.alloc(NPx, s); pk1.alloc(NPx, s); pv0.alloc(NPx, s); pv1.alloc(NPx, s);
  d.pair_a.alloc(NPx, s); d.pair_b.alloc(NPx, s);
  uniq.alloc(NPx, s);
  DevBuf<int> d_nuniq;
  d_nuniq.alloc(1, s);
  int h_nuniq = 0;
  if (h_npairs > 0) {
    k_make_pairs<<<grid(P, 128), 128, 0, s>>>(P, d.t_off.p, d.t_obs.p, d.o_view.p, pbase.p, pk0.p, pv0.p);
    sort_pairs(tmp, pk0.p, pk1.p, pv0.p, pv1.p, h_npairs, 0, 32 + vbits, s);
    k_split_pairs<<<grid(h_npairs, 256), 256, 0, s>>>(h_npairs, pv1.p, d.pair_a.p, d.pair_b.p);
    size_t bytes = 0;
    PTZ_CUDA(cub::DeviceSelect::Unique(nullptr, bytes, pk1.p, uniq.p, d_nuniq.p, h_npairs, s));
    void* t = tmp.get(bytes);
    PTZ_CUDA(cub::DeviceSelect::Unique(t, bytes, pk1.p, uniq.p, d_nuniq.p, h_npairs, s));
    PTZ_CUDA(cudaMemcpyAsync(&h_nuniq, d_nuniq.p, 4, cudaMemcpyDeviceToHost, s));
    PTZ_CUDA(cudaStreamSynchronize(s));
  }
  d.n_local = h_nuniq;
}

// phase B: block keys (local, or their union with the global ones on a sharded problem), pair ranges, block CSR
// sorted union of `count` keys that are already on the device (duplicates and the ~0 padding removed): radix sort + unique
inline int union_keys_device(unsigned long long* keys, long long count, DevBuf<unsigned long long>& out, cudaStream_t s) {
  using namespace setup;
  CubTemp tmp;
  tmp.s = s;
  DevBuf<unsigned long long> sorted;
  DevBuf<int> d_n;
  sorted.alloc((size_t)count, s);
  out.alloc((size_t)count, s);
  d_n.alloc(1, s);
  size_t bytes = 0;
  PTZ_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys, sorted.p, count, 0, 64, s));
  void* t = tmp.get(bytes);
  PTZ_CUDA(cub::DeviceRadixSort::SortKeys(t, bytes, keys, sorted.p, count, 0, 64, s));
  PTZ_CUDA(cub::DeviceSelect::Unique(nullptr, bytes, sorted.p, out.p, d_n.p, count, s));
  t = tmp.get(bytes);
  PTZ_CUDA(cub::DeviceSelect::Unique(t, bytes, sorted.p, out.p, d_n.p, count, s));
  int n = 0;
  unsigned long long last = 0;
  d_n.download(&n, 1, s);
  PTZ_CUDA(cudaStreamSynchronize(s));
  if (n > 0) {
    PTZ_CUDA(cudaMemcpyAsync(&last, out.p + (n - 1), 8, cudaMemcpyDeviceToHost, s));
    PTZ_CUDA(cudaStreamSynchronize(s));
    if (last == ~0ull) --n;  // the padding of the all-gather
  }
  return n;
}

// extra_upper_keys / dev_union: the block pattern every rank must share (sharded problems); dev_union is already the sorted
// union INCLUDING the local keys.
inline void build_structure_device_blocks(DevStructure& d, const std::vector<int64_t>* extra_upper_keys, cudaStream_t s,
                                          const unsigned long long* dev_union = nullptr, int n_union = 0) {
  using namespace setup;
  CubTemp tmp;
  tmp.s = s;
  auto grid = [](long long n, int b) { return (unsigned)std::max<long long>((n + b - 1) / b, 1); };
  const int V = d.V;
  const long long h_npairs = d.npairs;
  const int h_nuniq = d.n_local;
  const int vbits = bits_for(std::max(V, 2));
  DevBuf<unsigned long long>& pk1 = d.pk_sorted;
  DevBuf<unsigned long long>& uniq = d.uniq_local;
  if (dev_union) {
    d.nub = n_union;
    d.ub_keys.alloc(std::max(n_union, 1), s);
    if (n_union) PTZ_CUDA(cudaMemcpyAsync(d.ub_keys.p, dev_union, (size_t)n_union * 8, cudaMemcpyDeviceToDevice, s));
  } else if (extra_upper_keys && !extra_upper_keys->empty()) {
    std::vector<unsigned long long> local(h_nuniq);
    if (h_nuniq) { uniq.download(local.data(), h_nuniq, s); PTZ_CUDA(cudaStreamSynchronize(s)); }
    std::vector<unsigned long long> merged(local.size() + extra_upper_keys->size());
    std::vector<unsigned long long> extra(extra_upper_keys->begin(), extra_upper_keys->end());
    auto e = std::set_union(local.begin(), local.end(), extra.begin(), extra.end(), merged.begin());
    merged.resize(e - merged.begin());
    d.ub_keys.upload(merged, s);
    d.nub = (int)merged.size();
  } else {
    d.nub = h_nuniq;
    d.ub_keys.alloc(std::max(h_nuniq, 1), s);
    if (h_nuniq) PTZ_CUDA(cudaMemcpyAsync(d.ub_keys.p, uniq.p, (size_t)h_nuniq * 8, cudaMemcpyDeviceToDevice, s));
  }
  const int nub = d.nub, nubx = std::max(nub, 1);
  d.pair_off.alloc(nub + 1, s);
  {
    DevBuf<long long> off;
    off.alloc(nub + 1, s);
    k_lower_bounds_u64<<<grid(nub + 1, 256), 256, 0, s>>>(nub, h_npairs, pk1.p, d.ub_keys.p, off.p);
    PTZ_CUDA(cudaMemcpyAsync(d.pair_off.p, off.p, (size_t)(nub + 1) * 8, cudaMemcpyDeviceToDevice, s));
    PTZ_CUDA(cudaStreamSynchronize(s));
  }
  // ---- block CSR (both triangles)
  d.nnzb = V + 2 * nub;
  DevBuf<unsigned long long> c0, c1;
  DevBuf<int> ci0, ci1, first_ub, first_lb, rowlen;
  c0.alloc(nubx, s); c1.alloc(nubx, s); ci0.alloc(nubx, s); ci1.alloc(nubx, s); first_ub.alloc(V + 1, s); first_lb.alloc(V + 1, s); rowlen.alloc(V + 1, s);
  rowlen.zero(s);
  d.s_rowptr.alloc(V + 1, s); d.s_col.alloc(d.nnzb, s); d.blk_row.alloc(d.nnzb, s); d.diag_pos.alloc(V, s); d.ub_pos.alloc(nubx, s); d.ub_pos_t.alloc(nubx, s);
  if (nub > 0) {
    k_swap_keys<<<grid(nub, 256), 256, 0, s>>>(nub, d.ub_keys.p, c0.p, ci0.p);
    sort_pairs(tmp, c0.p, c1.p, ci0.p, ci1.p, nub, 0, 32 + vbits, s);
  }
  k_row_bounds<<<grid(V + 1, 256), 256, 0, s>>>(V, nub, d.ub_keys.p, c1.p, first_ub.p, first_lb.p);
  k_row_len<<<grid(V, 256), 256, 0, s>>>(V, first_ub.p, first_lb.p, rowlen.p);
  exclusive_scan(tmp, rowlen.p, d.s_rowptr.p, V + 1, s);
  k_fill_csr<<<grid(std::max(V, nub), 256), 256, 0, s>>>(V, nub, d.ub_keys.p, c1.p, ci1.p, first_ub.p, first_lb.p, d.s_rowptr.p, d.s_col.p, d.blk_row.p,
                                                         d.diag_pos.p, d.ub_pos.p, d.ub_pos_t.p);
  PTZ_CUDA(cudaGetLastError());
  d.h_rowptr.resize(V + 1);
  d.s_rowptr.download(d.h_rowptr.data(), V + 1, s);
  PTZ_CUDA(cudaStreamSynchronize(s));
  d.pk_sorted.release();
  d.uniq_local.release();
}

// upload of a host-built structure into the same device representation
inline void upload_structure(const BaStructure& st, DevStructure& d, cudaStream_t s) {
  d.V = st.V; d.P = st.P; d.M = st.M; d.nchunks = st.nchunks(); d.nub = st.nub(); d.nnzb = st.nnzb(); d.npairs = (int64_t)st.pair_a.size();
  d.o_uv.upload(reinterpret_cast<const float2*>(st.o_uv.data()), st.M, s);
  if (st.M == 0) d.o_uv.alloc(1, s);
  auto up = [&](DevBuf<int>& b, const std::vector<int>& v) { if (v.empty()) b.alloc(1, s); else b.upload(v, s); };
  up(d.perm, st.perm); up(d.o_view, st.o_view); up(d.o_track, st.o_track); up(d.view_off, st.view_off);
  up(d.chunk_view, st.chunk_view); up(d.chunk_begin, st.chunk_begin); up(d.chunk_cnt, st.chunk_cnt); up(d.view_chunk_off, st.view_chunk_off);
  up(d.t_off, st.t_off); up(d.t_obs, st.t_obs); up(d.pair_a, st.pair_a); up(d.pair_b, st.pair_b);
  d.pair_off.upload(st.ub_pair_off, s);
  up(d.s_rowptr, st.s_rowptr); up(d.s_col, st.s_col); up(d.diag_pos, st.diag_pos); up(d.ub_pos, st.ub_pos); up(d.ub_pos_t, st.ub_pos_t);
  std::vector<int> blk_row(st.nnzb()), active(st.V, 0);
  for (int v = 0; v < st.V; ++v) {
    for (int k = st.s_rowptr[v]; k < st.s_rowptr[v + 1]; ++k) blk_row[k] = v;
    active[v] = st.view_off[v + 1] > st.view_off[v];
  }
  up(d.blk_row, blk_row); up(d.view_active, active);
  std::vector<unsigned long long> keys(st.nub());
  for (int b = 0; b < st.nub(); ++b) keys[b] = ((unsigned long long)(unsigned)st.ub_row[b] << 32) | (unsigned)st.ub_col[b];
  if (keys.empty()) d.ub_keys.alloc(1, s); else d.ub_keys.upload(keys, s);
  d.h_rowptr = st.s_rowptr;
}

}  // namespace ptz
