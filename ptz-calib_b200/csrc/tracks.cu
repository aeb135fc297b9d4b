// tracks.cu — feature tracks from pairwise matches, and their flattening into the observation arrays of ptzba_problem.
//
// Replaces PTZRayOptimizer::FindTracks (ptzray_optimizer.cc:537-552): TracksBuilder::Build / Filter / ExportToSTL
// (src/core/tracks.cc:19-113, a sequential std::set + union-by-rank UnionFind, union_find.h:28-106) and the residual-block
// loop of AddConstraints2d2d (ptzray_optimizer.cc:801-848) that turns tracks into (uv, view, ray) rows.  SURVEY.md §8f row 1.
//
// Device algorithm (integer work, HBM/L2-bound; every step is a radix sort, a scan or a one-thread-per-item kernel):
//   1. every match contributes two endpoint keys (image << fbits | feature, packed to the bits in use); sort them carrying the endpoint number
//   2. heads of equal-key runs -> inclusive scan = flat node index of every endpoint.  Because the keys are sorted this IS
//      the index the reference's flat_pair_map assigns (tracks.cc:35-43 iterates a std::set of pairs, i.e. ascending)
//   3. lock-free union-find over the matches: roots are hooked larger-under-smaller with atomicCAS, so the final root of a
//      component is its smallest node whatever the interleaving (deterministic labels); then one pass of full compression
//   4. stable sort of the nodes by root: a component becomes a run, inside it nodes ascend = (image, feature) ascends
//   5. Filter (tracks.cc:63-101): a run with two neighbouring nodes of the same image lists that image twice -> rejected;
//      fewer than min_track_length nodes -> rejected; ExportToSTL (tracks.cc:103-118) also drops 1-node sets
//   6. scans over the surviving runs -> track offsets; one more pass emits (image, feature) per element
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

namespace ptz {
namespace trk {

struct Temp {
  DevBuf<char> buf;
  cudaStream_t s = nullptr;
  void* get(size_t bytes) {
    if (bytes > buf.n) buf.alloc(bytes + bytes / 4 + 256, s);
    return buf.p;
  }
};
template <class K, class Vv>
static void sort_pairs(Temp& tmp, const K* kin, K* kout, const Vv* vin, Vv* vout, long long n, int end_bit, cudaStream_t s) {
  size_t bytes = 0;
  PTZ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
  void* t = tmp.get(bytes);
  PTZ_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
}
template <class Tin, class Tout>
static void inclusive_scan(Temp& tmp, const Tin* in, Tout* out, long long n, cudaStream_t s) {
  size_t bytes = 0;
  PTZ_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, s));
  void* t = tmp.get(bytes);
  PTZ_CUDA(cub::DeviceScan::InclusiveSum(t, bytes, in, out, n, s));
}
template <class Tin, class Tout>
static void exclusive_scan(Temp& tmp, const Tin* in, Tout* out, long long n, cudaStream_t s) {
  size_t bytes = 0;
  PTZ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
  void* t = tmp.get(bytes);
  PTZ_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, n, s));
}
static inline unsigned grid_for(long long n, int b) { return (unsigned)std::max<long long>((n + b - 1) / b, 1); }
static inline int bits_for(long long n) { int b = 1; while ((1ll << b) < n) ++b; return b; }

// ---- 1. endpoint keys.  Endpoint e = 2*match + side.  The pair of a match: largest k with match_offset[k] <= match.
// Keys are packed as image << fbits | feature with fbits = bits of the largest feature index, so that the radix sort only
// walks the bits in use (a 1000-image, 2000-keypoint problem sorts 21 bits instead of 42).
__global__ void k_index_ranges(long long N, int npairs, const int* __restrict__ pair_src, const int* __restrict__ pair_dst, const int* __restrict__ query_idx,
                               const int* __restrict__ train_idx, unsigned int* __restrict__ maxv /* [0] image, [1] feature */, int* __restrict__ bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int f = 0, g = 0;
  if (i < N) { const int q = query_idx[i], t = train_idx[i]; if (q < 0 || t < 0) *bad = 1; f = max(max(q, t), 0); }
  if (i < npairs) { const int a = pair_src[i], c = pair_dst[i]; if (a < 0 || c < 0) *bad = 1; g = max(max(a, c), 0); }
  // warp maximum -> block maximum in shared memory -> one global atomic per block
  __shared__ unsigned int sm[2];
  if (threadIdx.x < 2) sm[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int fm = __reduce_max_sync(0xffffffffu, (unsigned int)f), gm = __reduce_max_sync(0xffffffffu, (unsigned int)g);
  if ((threadIdx.x & 31) == 0) { if (fm) atomicMax(&sm[1], fm); if (gm) atomicMax(&sm[0], gm); }
  __syncthreads();
  if (threadIdx.x < 2 && sm[threadIdx.x]) atomicMax(maxv + threadIdx.x, sm[threadIdx.x]);
}
__device__ __forceinline__ int pair_of(const long long* __restrict__ match_offset, int npairs, long long m) {
  int lo = 0, up = npairs;  // invariant: match_offset[lo] <= m < match_offset[up]
  while (up - lo > 1) {
    const int mid = (lo + up) >> 1;
    if (match_offset[mid] <= m) lo = mid; else up = mid;
  }
  return lo;
}
__global__ void k_endpoint_keys(long long N, int npairs, int fbits, const int* __restrict__ pair_src, const int* __restrict__ pair_dst,
                                const long long* __restrict__ match_offset, const int* __restrict__ query_idx, const int* __restrict__ train_idx,
                                unsigned long long* __restrict__ key, int* __restrict__ val) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long mc = m < N ? m : N - 1;
  // consecutive matches mostly belong to the same image pair: the first and last lane search, the others only if those differ
  const int lane = threadIdx.x & 31;
  int pr = (lane == 0 || lane == 31) ? pair_of(match_offset, npairs, mc) : 0;
  const int p0 = __shfl_sync(0xffffffffu, pr, 0), p1 = __shfl_sync(0xffffffffu, pr, 31);
  if (p0 == p1) pr = p0;
  else if (lane != 0 && lane != 31) pr = pair_of(match_offset, npairs, mc);
  if (m >= N) return;
  const unsigned long long a = ((unsigned long long)(unsigned int)pair_src[pr] << fbits) | (unsigned int)query_idx[m];
  const unsigned long long c = ((unsigned long long)(unsigned int)pair_dst[pr] << fbits) | (unsigned int)train_idx[m];
  reinterpret_cast<ulonglong2*>(key)[m] = make_ulonglong2(a, c);
  reinterpret_cast<int2*>(val)[m] = make_int2((int)(2 * m), (int)(2 * m + 1));
}

// ---- 2. run heads of the sorted keys
__global__ void k_key_heads(long long n, const unsigned long long* __restrict__ key, int* __restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}
__global__ void k_assign_nodes(long long n, const unsigned long long* __restrict__ key, const int* __restrict__ val, const int* __restrict__ head,
                               const int* __restrict__ incl, int* __restrict__ endpoint_node, unsigned long long* __restrict__ node_key) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int node = incl[i] - 1;
  endpoint_node[val[i]] = node;
  if (head[i]) node_key[node] = key[i];
}

// ---- 3. union-find.  Invariant: parent[x] <= x.  Racy path shortening only ever writes an ancestor, so it is safe.
__device__ __forceinline__ int uf_root(volatile int* parent, int x) {
  int cur = parent[x];
  if (cur != x) {
    int prev = x, next;
    while (cur > (next = parent[cur])) { parent[prev] = next; prev = cur; cur = next; }
  }
  return cur;
}
__global__ void k_iota(int n, int* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
__global__ void k_union(long long N, const int* __restrict__ endpoint_node, int* parent) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= N) return;
  int a = uf_root(parent, endpoint_node[2 * m]), b = uf_root(parent, endpoint_node[2 * m + 1]);
  // hook the larger root under the smaller one; a failed CAS means the root moved: follow it and retry
  while (a != b) {
    if (a < b) { const int t = a; a = b; b = t; }  // a > b
    const int old = atomicCAS(&parent[a], a, b);
    if (old == a) break;
    a = old;
  }
}
__global__ void k_compress(int n, int* parent, int* __restrict__ root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) root[i] = uf_root(parent, i);
}

// ---- 5. runs of equal root (nodes stably sorted by root): heads, image listed twice
__global__ void k_run_heads(int K, int fbits, const int* __restrict__ sroot, const int* __restrict__ snode, const unsigned long long* __restrict__ node_key,
                            int* __restrict__ head, int* __restrict__ twice /* [K] by root, zeroed */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const bool h = (i == 0 || sroot[i] != sroot[i - 1]);
  head[i] = h ? 1 : 0;
  if (!h && (node_key[snode[i]] >> fbits) == (node_key[snode[i - 1]] >> fbits)) twice[sroot[i]] = 1;  // every writer stores the same value
}
__global__ void k_run_starts(int K, const int* __restrict__ head, const int* __restrict__ incl, int* __restrict__ run_start) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K && head[i]) run_start[incl[i] - 1] = i;
}
__global__ void k_run_valid(int C, int K, int min_len, const int* __restrict__ run_start, const int* __restrict__ sroot, const int* __restrict__ twice,
                            int* __restrict__ valid, long long* __restrict__ vlen) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int b = run_start[c], e = (c + 1 < C) ? run_start[c + 1] : K;
  const int len = e - b;
  // Filter: image listed twice, or fewer than min_track_length images (tracks.cc:76-90); ExportToSTL: size > 1 (tracks.cc:112)
  const bool ok = !twice[sroot[b]] && len >= min_len && len > 1;
  valid[c] = ok ? 1 : 0;
  vlen[c] = ok ? len : 0;
}
__global__ void k_emit(int K, int C, int fbits, const int* __restrict__ incl, const int* __restrict__ run_start, const int* __restrict__ valid,
                       const int* __restrict__ trank, const long long* __restrict__ eoff, const int* __restrict__ sroot, const int* __restrict__ snode,
                       const unsigned long long* __restrict__ node_key, long long cap_tracks, long long cap_elems, int* __restrict__ track_id,
                       long long* __restrict__ track_offset, int* __restrict__ elem_img, int* __restrict__ elem_feat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const int c = incl[i] - 1;
  if (!valid[c]) return;
  const int b = run_start[c];
  const long long pos = eoff[c] + (i - b);
  if (pos < cap_elems) {
    const unsigned long long k = node_key[snode[i]];
    elem_img[pos] = (int)(k >> fbits);
    elem_feat[pos] = (int)(k & ((1ull << fbits) - 1ull));
  }
  if (i == b && trank[c] < cap_tracks) { track_id[trank[c]] = sroot[i]; track_offset[trank[c]] = eoff[c]; }
}
__global__ void k_finish(int C, const int* __restrict__ valid, const int* __restrict__ trank, const long long* __restrict__ vlen,
                         const long long* __restrict__ eoff, long long cap_tracks, long long* __restrict__ track_offset, long long* __restrict__ counts) {
  // counts: [0] nodes (set by the host side), [1] components, [2] tracks, [3] elements
  const long long nt = C > 0 ? trank[C - 1] + valid[C - 1] : 0;
  const long long ne = C > 0 ? eoff[C - 1] + vlen[C - 1] : 0;
  counts[1] = C; counts[2] = nt; counts[3] = ne;
  if (nt <= cap_tracks) track_offset[nt] = ne;
}

struct DevCounts { long long nodes = 0, components = 0, tracks = 0, elems = 0; int bad = 0; };

// all pointers are device pointers
static DevCounts build_device(long long N, int npairs, const int* pair_src, const int* pair_dst, const long long* match_offset, const int* query_idx,
                              const int* train_idx, int min_len, long long cap_tracks, long long cap_elems, int* track_id, long long* track_offset,
                              int* elem_img, int* elem_feat, cudaStream_t s) {
  DevCounts out;
  Temp tmp;
  tmp.s = s;
  DevBuf<long long> d_counts;
  d_counts.alloc(4, s);
  d_counts.zero(s);
  if (N == 0 || npairs == 0) {
    if (cap_tracks >= 0 && track_offset) PTZ_CUDA(cudaMemsetAsync(track_offset, 0, 8, s));
    PTZ_CUDA(cudaStreamSynchronize(s));
    return out;
  }
  const long long E = 2 * N;
  DevBuf<unsigned long long> key0, key1, node_key;
  DevBuf<int> val0, val1, head, incl, endpoint_node, d_bad;
  DevBuf<unsigned int> d_max;
  key0.alloc(E, s); key1.alloc(E, s); val0.alloc(E, s); val1.alloc(E, s); head.alloc(E, s); incl.alloc(E, s); endpoint_node.alloc(E, s);
  d_bad.alloc(1, s); d_bad.zero(s); d_max.alloc(2, s); d_max.zero(s);
  k_index_ranges<<<grid_for(std::max<long long>(N, npairs), 256), 256, 0, s>>>(N, npairs, pair_src, pair_dst, query_idx, train_idx, d_max.p, d_bad.p);
  unsigned int h_max[2] = {0, 0};
  PTZ_CUDA(cudaMemcpyAsync(h_max, d_max.p, 8, cudaMemcpyDeviceToHost, s));
  PTZ_CUDA(cudaMemcpyAsync(&out.bad, d_bad.p, 4, cudaMemcpyDeviceToHost, s));
  PTZ_CUDA(cudaStreamSynchronize(s));
  if (out.bad) return out;
  const int fbits = bits_for((long long)h_max[1] + 1), ibits = bits_for((long long)h_max[0] + 1);
  k_endpoint_keys<<<grid_for(N, 256), 256, 0, s>>>(N, npairs, fbits, pair_src, pair_dst, match_offset, query_idx, train_idx, key0.p, val0.p);
  sort_pairs(tmp, key0.p, key1.p, val0.p, val1.p, E, fbits + ibits, s);
  k_key_heads<<<grid_for(E, 256), 256, 0, s>>>(E, key1.p, head.p);
  inclusive_scan(tmp, head.p, incl.p, E, s);
  int K = 0;
  PTZ_CUDA(cudaMemcpyAsync(&K, incl.p + (E - 1), 4, cudaMemcpyDeviceToHost, s));
  PTZ_CUDA(cudaStreamSynchronize(s));
  out.nodes = K;
  node_key.alloc(K, s);
  k_assign_nodes<<<grid_for(E, 256), 256, 0, s>>>(E, key1.p, val1.p, head.p, incl.p, endpoint_node.p, node_key.p);
  // union-find
  DevBuf<int> parent, root, ids, sroot, snode, twice, run_start, valid, trank;
  DevBuf<long long> vlen, eoff;
  parent.alloc(K, s); root.alloc(K, s); ids.alloc(K, s); sroot.alloc(K, s); snode.alloc(K, s); twice.alloc(K, s);
  k_iota<<<grid_for(K, 256), 256, 0, s>>>(K, parent.p);
  k_iota<<<grid_for(K, 256), 256, 0, s>>>(K, ids.p);
  k_union<<<grid_for(N, 256), 256, 0, s>>>(N, endpoint_node.p, parent.p);
  k_compress<<<grid_for(K, 256), 256, 0, s>>>(K, parent.p, root.p);
  // components as runs
  sort_pairs(tmp, root.p, sroot.p, ids.p, snode.p, K, bits_for(K), s);
  twice.zero(s);
  k_run_heads<<<grid_for(K, 256), 256, 0, s>>>(K, fbits, sroot.p, snode.p, node_key.p, head.p, twice.p);
  inclusive_scan(tmp, head.p, incl.p, K, s);
  int Cn = 0;
  PTZ_CUDA(cudaMemcpyAsync(&Cn, incl.p + (K - 1), 4, cudaMemcpyDeviceToHost, s));
  PTZ_CUDA(cudaStreamSynchronize(s));
  run_start.alloc(Cn, s); valid.alloc(Cn, s); trank.alloc(Cn, s); vlen.alloc(Cn, s); eoff.alloc(Cn, s);
  k_run_starts<<<grid_for(K, 256), 256, 0, s>>>(K, head.p, incl.p, run_start.p);
  k_run_valid<<<grid_for(Cn, 256), 256, 0, s>>>(Cn, K, min_len, run_start.p, sroot.p, twice.p, valid.p, vlen.p);
  exclusive_scan(tmp, valid.p, trank.p, Cn, s);
  exclusive_scan(tmp, vlen.p, eoff.p, Cn, s);
  k_emit<<<grid_for(K, 256), 256, 0, s>>>(K, Cn, fbits, incl.p, run_start.p, valid.p, trank.p, eoff.p, sroot.p, snode.p, node_key.p, cap_tracks, cap_elems, track_id,
                                          track_offset, elem_img, elem_feat);
  k_finish<<<1, 1, 0, s>>>(Cn, valid.p, trank.p, vlen.p, eoff.p, cap_tracks, track_offset, d_counts.p);
  long long h_counts[4];
  PTZ_CUDA(cudaMemcpyAsync(h_counts, d_counts.p, 32, cudaMemcpyDeviceToHost, s));
  PTZ_CUDA(cudaStreamSynchronize(s));
  PTZ_CUDA(cudaGetLastError());
  out.components = h_counts[1]; out.tracks = h_counts[2]; out.elems = h_counts[3];
  return out;
}

// ---- flattening: one thread per track (tracks are a handful of elements long)
__global__ void k_track_cand(int T, const long long* __restrict__ toff, const int* __restrict__ eimg, const int* __restrict__ efeat, int num_images,
                             const unsigned char* __restrict__ cand, const long long* __restrict__ kp_off, int* __restrict__ has_row,
                             int* __restrict__ ncand, int* __restrict__ bad) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int n = 0;
  for (long long i = toff[t]; i < toff[t + 1]; ++i) {
    const int img = eimg[i], f = efeat[i];
    if (img < 0 || img >= num_images || f < 0 || f >= kp_off[img + 1] - kp_off[img]) { *bad = 1; continue; }
    n += cand[img] ? 1 : 0;
  }
  ncand[t] = n;
  has_row[t] = n > 0 ? 1 : 0;
}
__global__ void k_cand_flags(int n, const unsigned char* __restrict__ cand, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = cand[i] ? 1 : 0;
}
__global__ void k_emit_obs(int T, const long long* __restrict__ toff, const int* __restrict__ eimg, const int* __restrict__ efeat,
                           const unsigned char* __restrict__ cand, const int* __restrict__ dense, const long long* __restrict__ kp_off,
                           const float2* __restrict__ kp_uv, const int* __restrict__ has_row, const int* __restrict__ row_of, const int* __restrict__ obs_off,
                           long long cap_rows, long long cap_obs, int* __restrict__ row_track, double* __restrict__ weight, float2* __restrict__ obs_uv,
                           int* __restrict__ obs_view, int* __restrict__ obs_track) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T || !has_row[t]) return;
  const int row = row_of[t];
  if (row < cap_rows) { row_track[row] = t; weight[row] = (double)(toff[t + 1] - toff[t]); }
  long long o = obs_off[t];
  for (long long i = toff[t]; i < toff[t + 1]; ++i) {
    const int img = eimg[i];
    if (!cand[img]) continue;
    if (o < cap_obs) { obs_uv[o] = kp_uv[kp_off[img] + efeat[i]]; obs_view[o] = dense[img]; obs_track[o] = row; }
    ++o;
  }
}

template <class F>
static int guarded_t(F&& f) {
  try {
    return f();
  } catch (const CudaError& e) {
    set_last_error("%s", e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error("%s", e.what());
    return PTZ_ERR_CUDA;
  }
}
static int need_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) { set_last_error("no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e)); return PTZ_ERR_NO_DEVICE; }
  return PTZ_OK;
}

}  // namespace trk
}  // namespace ptz

using namespace ptz;
using namespace ptz::trk;

extern "C" {

int ptztracks_build_dev(const ptztracks_matches* m, int64_t num_matches, ptztracks_result* out, void* cuda_stream) {
  if (!m || !out || m->num_pairs < 0 || num_matches < 0 || num_matches > (1ll << 30)) return PTZ_ERR_INVALID;
  int rc = need_device();
  if (rc != PTZ_OK) return rc;
  return guarded_t([&]() {
    enable_memory_pool();  // keep the scratch of repeated calls in the stream-ordered pool
    const DevCounts c = build_device(num_matches, m->num_pairs, m->pair_src, m->pair_dst, reinterpret_cast<const long long*>(m->match_offset), m->query_idx,
                                     m->train_idx, m->min_track_length, out->cap_tracks, out->cap_elems, out->track_id,
                                     reinterpret_cast<long long*>(out->track_offset), out->elem_img, out->elem_feat, (cudaStream_t)cuda_stream);
    if (c.bad) { set_last_error("negative image or feature index in the matches"); return (int)PTZ_ERR_INVALID; }
    out->num_nodes = (int32_t)c.nodes; out->num_components = (int32_t)c.components; out->num_tracks = (int32_t)c.tracks; out->num_elems = c.elems;
    if (c.tracks > out->cap_tracks || c.elems > out->cap_elems) { set_last_error("track capacity too small: need %lld tracks, %lld elements", c.tracks, c.elems); return (int)PTZ_ERR_INVALID; }
    return (int)PTZ_OK;
  });
}

int ptztracks_build(const ptztracks_matches* m, ptztracks_result* out) {
  if (!m || !out || m->num_pairs < 0 || out->cap_tracks < 0 || out->cap_elems < 0) return PTZ_ERR_INVALID;
  if (m->num_pairs > 0 && (!m->pair_src || !m->pair_dst || !m->match_offset)) return PTZ_ERR_INVALID;
  const int64_t N = m->num_pairs > 0 ? m->match_offset[m->num_pairs] : 0;
  if (N < 0 || (N > 0 && (!m->query_idx || !m->train_idx))) return PTZ_ERR_INVALID;
  if (N > (1ll << 30) || (m->num_pairs > 0 && m->match_offset[0] != 0)) return PTZ_ERR_INVALID;
  for (int k = 0; k < m->num_pairs; ++k)
    if (m->match_offset[k] > m->match_offset[k + 1]) return PTZ_ERR_INVALID;
  if (!out->track_offset || (out->cap_tracks > 0 && !out->track_id) || (out->cap_elems > 0 && (!out->elem_img || !out->elem_feat))) return PTZ_ERR_INVALID;
  int rc = need_device();
  if (rc != PTZ_OK) return rc;
  out->num_nodes = out->num_components = out->num_tracks = 0;
  out->num_elems = 0;
  out->track_offset[0] = 0;
  if (N == 0) return PTZ_OK;  // no matches: no nodes, no tracks (tracks.cc:19-62 on empty input)
  return guarded_t([&]() {
    StreamHolder sh;
    sh.create();
    cudaStream_t s = sh.s;
    enable_memory_pool();
    DevBuf<int> d_src, d_dst, d_q, d_t, d_tid, d_eimg, d_efeat;
    DevBuf<long long> d_off, d_toff;
    d_src.upload(m->pair_src, m->num_pairs, s); d_dst.upload(m->pair_dst, m->num_pairs, s);
    d_off.upload(reinterpret_cast<const long long*>(m->match_offset), (size_t)m->num_pairs + 1, s);
    d_q.upload(m->query_idx, N, s); d_t.upload(m->train_idx, N, s);
    d_tid.alloc(std::max<int64_t>(out->cap_tracks, 1), s); d_toff.alloc(out->cap_tracks + 1, s);
    d_eimg.alloc(std::max<int64_t>(out->cap_elems, 1), s); d_efeat.alloc(std::max<int64_t>(out->cap_elems, 1), s);
    const DevCounts c = build_device(N, m->num_pairs, d_src.p, d_dst.p, d_off.p, d_q.p, d_t.p, m->min_track_length, out->cap_tracks, out->cap_elems, d_tid.p,
                                     d_toff.p, d_eimg.p, d_efeat.p, s);
    if (c.bad) { set_last_error("negative image or feature index in the matches"); return (int)PTZ_ERR_INVALID; }
    out->num_nodes = (int32_t)c.nodes; out->num_components = (int32_t)c.components; out->num_tracks = (int32_t)c.tracks; out->num_elems = c.elems;
    if (c.tracks > out->cap_tracks || c.elems > out->cap_elems) { set_last_error("track capacity too small: need %lld tracks, %lld elements", c.tracks, c.elems); return (int)PTZ_ERR_INVALID; }
    d_toff.download(reinterpret_cast<long long*>(out->track_offset), (size_t)c.tracks + 1, s);
    d_tid.download(out->track_id, (size_t)c.tracks, s);
    d_eimg.download(out->elem_img, (size_t)c.elems, s); d_efeat.download(out->elem_feat, (size_t)c.elems, s);
    PTZ_CUDA(cudaStreamSynchronize(s));
    return (int)PTZ_OK;
  });
}

// Reference track ids.  The device build labels a track by the flat index of its smallest (image, feature) node; the reference's id
// is the root its SEQUENTIAL union-by-rank forest ends up with (union_find.h:66-92: equal ranks keep the first argument's root,
// i.e. the source node of the match), which depends on the order of the matches and cannot be formed in parallel.  It only orders
// the tracks (std::map<int, Track>, tracks.h:32) and fills Ray::id_ (types.h:35), which nothing in the reference reads -- but a
// drop-in should hand back the same numbers, so this host pass replays the unions in match order (one sort of the 2N nodes, then
// near-linear time), relabels the tracks and re-sorts them by the reference id.  Plain host code: no device needed.
int ptztracks_reference_ids(const ptztracks_matches* m, ptztracks_result* tr) {
  if (!m || !tr || m->num_pairs < 0 || tr->num_tracks < 0) return PTZ_ERR_INVALID;
  if (tr->num_tracks == 0) return PTZ_OK;
  if (!m->pair_src || !m->pair_dst || !m->match_offset || !tr->track_id || !tr->track_offset || !tr->elem_img || !tr->elem_feat) return PTZ_ERR_INVALID;
  const int64_t N = m->match_offset[m->num_pairs];
  if (N > 0 && (!m->query_idx || !m->train_idx)) return PTZ_ERR_INVALID;
  auto key = [](int img, int feat) { return ((uint64_t)(uint32_t)img << 32) | (uint32_t)feat; };
  std::vector<uint64_t> nodes;
  nodes.reserve(2 * (size_t)N);
  for (int k = 0; k < m->num_pairs; ++k)
    for (int64_t i = m->match_offset[k]; i < m->match_offset[k + 1]; ++i) {
      nodes.push_back(key(m->pair_src[k], m->query_idx[i]));
      nodes.push_back(key(m->pair_dst[k], m->train_idx[i]));
    }
  std::sort(nodes.begin(), nodes.end());
  nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());  // flat index = position (flat_pair_map, tracks.cc:35-43)
  auto index_of = [&](uint64_t kx) { return (int)(std::lower_bound(nodes.begin(), nodes.end(), kx) - nodes.begin()); };
  std::vector<int> parent(nodes.size()), rank(nodes.size(), 0);
  for (size_t i = 0; i < parent.size(); ++i) parent[i] = (int)i;
  auto find = [&](int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
  for (int k = 0; k < m->num_pairs; ++k)
    for (int64_t i = m->match_offset[k]; i < m->match_offset[k + 1]; ++i) {
      const int a = find(index_of(key(m->pair_src[k], m->query_idx[i]))), b = find(index_of(key(m->pair_dst[k], m->train_idx[i])));
      if (a == b) continue;
      if (rank[a] < rank[b]) parent[a] = b;
      else { parent[b] = a; if (rank[a] == rank[b]) ++rank[a]; }
    }
  const int T = tr->num_tracks;
  std::vector<int> ref(T), perm(T);
  for (int t = 0; t < T; ++t) {
    const int64_t o = tr->track_offset[t];
    const uint64_t kx = key(tr->elem_img[o], tr->elem_feat[o]);
    const int idx = index_of(kx);
    if (idx >= (int)nodes.size() || nodes[idx] != kx) { ptz::set_last_error("track %d does not come from these matches", t); return PTZ_ERR_INVALID; }
    ref[t] = find(idx);
    perm[t] = t;
  }
  std::sort(perm.begin(), perm.end(), [&](int a, int b) { return ref[a] < ref[b]; });
  const int64_t E = tr->track_offset[T];
  std::vector<int32_t> img(tr->elem_img, tr->elem_img + E), feat(tr->elem_feat, tr->elem_feat + E);
  std::vector<int64_t> off(tr->track_offset, tr->track_offset + T + 1);
  int64_t pos = 0;
  for (int q = 0; q < T; ++q) {
    const int t = perm[q];
    tr->track_id[q] = ref[t];
    tr->track_offset[q] = pos;
    for (int64_t i = off[t]; i < off[t + 1]; ++i, ++pos) { tr->elem_img[pos] = img[i]; tr->elem_feat[pos] = feat[i]; }
  }
  tr->track_offset[T] = pos;
  return PTZ_OK;
}

int ptztracks_flatten(const ptztracks_result* tr, const ptztracks_views* v, ptztracks_obs* out) {
  if (!tr || !v || !out || tr->num_tracks < 0 || tr->num_elems < 0 || v->num_images < 0 || out->cap_rows < 0 || out->cap_obs < 0) return PTZ_ERR_INVALID;
  if (tr->num_tracks > 0 && (!tr->track_offset || !tr->elem_img || !tr->elem_feat || !v->is_candidate || !v->kp_offset || !v->kp_uv)) return PTZ_ERR_INVALID;
  if (tr->num_tracks > 0 && tr->track_offset[tr->num_tracks] != tr->num_elems) return PTZ_ERR_INVALID;
  int rc = need_device();
  if (rc != PTZ_OK) return rc;
  out->num_rows = 0; out->num_obs = 0;
  if (tr->num_tracks == 0) return PTZ_OK;
  return guarded_t([&]() {
    StreamHolder sh;
    sh.create();
    cudaStream_t s = sh.s;
    enable_memory_pool();
    Temp tmp;
    tmp.s = s;
    const int T = tr->num_tracks, NI = v->num_images;
    const long long Ne = tr->num_elems, nkp = v->kp_offset[NI];
    DevBuf<long long> d_toff, d_kpoff;
    DevBuf<int> d_eimg, d_efeat, d_flag, d_dense, d_has, d_ncand, d_row, d_ooff, d_bad, d_rowtrack, d_oview, d_otrack;
    DevBuf<unsigned char> d_cand;
    DevBuf<float2> d_kp, d_ouv;
    DevBuf<double> d_w;
    d_toff.upload(reinterpret_cast<const long long*>(tr->track_offset), (size_t)T + 1, s);
    d_eimg.upload(tr->elem_img, Ne, s); d_efeat.upload(tr->elem_feat, Ne, s);
    d_cand.upload(v->is_candidate, NI, s);
    d_kpoff.upload(reinterpret_cast<const long long*>(v->kp_offset), (size_t)NI + 1, s);
    d_kp.upload(reinterpret_cast<const float2*>(v->kp_uv), std::max<long long>(nkp, 0), s);
    d_flag.alloc(NI + 1, s); d_dense.alloc(NI + 1, s); d_has.alloc(T + 1, s); d_ncand.alloc(T + 1, s); d_row.alloc(T + 1, s); d_ooff.alloc(T + 1, s);
    d_bad.alloc(1, s); d_bad.zero(s); d_flag.zero(s); d_has.zero(s); d_ncand.zero(s);
    k_cand_flags<<<grid_for(NI, 256), 256, 0, s>>>(NI, d_cand.p, d_flag.p);
    exclusive_scan(tmp, d_flag.p, d_dense.p, NI + 1, s);
    k_track_cand<<<grid_for(T, 256), 256, 0, s>>>(T, d_toff.p, d_eimg.p, d_efeat.p, NI, d_cand.p, d_kpoff.p, d_has.p, d_ncand.p, d_bad.p);
    exclusive_scan(tmp, d_has.p, d_row.p, T + 1, s);
    exclusive_scan(tmp, d_ncand.p, d_ooff.p, T + 1, s);
    int h_rows = 0, h_obs = 0, h_bad = 0;
    PTZ_CUDA(cudaMemcpyAsync(&h_rows, d_row.p + T, 4, cudaMemcpyDeviceToHost, s));
    PTZ_CUDA(cudaMemcpyAsync(&h_obs, d_ooff.p + T, 4, cudaMemcpyDeviceToHost, s));
    PTZ_CUDA(cudaMemcpyAsync(&h_bad, d_bad.p, 4, cudaMemcpyDeviceToHost, s));
    PTZ_CUDA(cudaStreamSynchronize(s));
    if (h_bad) { set_last_error("track element outside the images / keypoints given"); return (int)PTZ_ERR_INVALID; }
    out->num_rows = h_rows; out->num_obs = h_obs;
    if (h_rows > out->cap_rows || h_obs > out->cap_obs) { set_last_error("observation capacity too small: need %d rows, %d observations", h_rows, h_obs); return (int)PTZ_ERR_INVALID; }
    if (h_rows == 0) return (int)PTZ_OK;
    if (!out->row_track || !out->track_weight || !out->obs_uv || !out->obs_view || !out->obs_track) return (int)PTZ_ERR_INVALID;
    d_rowtrack.alloc(h_rows, s); d_w.alloc(h_rows, s); d_ouv.alloc(std::max(h_obs, 1), s); d_oview.alloc(std::max(h_obs, 1), s); d_otrack.alloc(std::max(h_obs, 1), s);
    k_emit_obs<<<grid_for(T, 256), 256, 0, s>>>(T, d_toff.p, d_eimg.p, d_efeat.p, d_cand.p, d_dense.p, d_kpoff.p, d_kp.p, d_has.p, d_row.p, d_ooff.p, h_rows, h_obs,
                                                d_rowtrack.p, d_w.p, d_ouv.p, d_oview.p, d_otrack.p);
    PTZ_CUDA(cudaGetLastError());
    d_rowtrack.download(out->row_track, h_rows, s); d_w.download(out->track_weight, h_rows, s);
    d_ouv.download(reinterpret_cast<float2*>(out->obs_uv), h_obs, s); d_oview.download(out->obs_view, h_obs, s); d_otrack.download(out->obs_track, h_obs, s);
    PTZ_CUDA(cudaStreamSynchronize(s));
    return (int)PTZ_OK;
  });
}

}  // extern "C"
