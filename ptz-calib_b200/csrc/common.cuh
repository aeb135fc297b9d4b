// common.cuh — error handling, device buffers and reduction helpers shared by the BA and reloc translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ptzcalib_b200.h"

namespace ptz {

void set_last_error(const char* fmt, ...);

struct CudaError : std::runtime_error {
  int code;
  CudaError(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};

#define PTZ_CUDA(expr)                                                                                        \
  do {                                                                                                        \
    cudaError_t _e = (expr);                                                                                  \
    if (_e != cudaSuccess) {                                                                                  \
      char _b[512];                                                                                           \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));         \
      throw ::ptz::CudaError(PTZ_ERR_CUDA, _b);                                                               \
    }                                                                                                         \
  } while (0)

// Owns a stream; declare it BEFORE any DevBuf that allocates on it so that it is released after them.  Streams come from a
// small per-device free list and go back to it: the many solver objects of one IBA run (and NCCL, which keeps per-stream
// state) see the same few streams instead of a fresh one per call.
struct StreamHolder {
  cudaStream_t s = nullptr;
  int dev = 0;
  StreamHolder() {}
  StreamHolder(const StreamHolder&) = delete;
  StreamHolder& operator=(const StreamHolder&) = delete;
  static std::vector<cudaStream_t>& free_list(int device) {
    static std::vector<cudaStream_t> lists[64];
    return lists[device & 63];
  }
  static std::mutex& lock() { static std::mutex m; return m; }
  void create() {
    PTZ_CUDA(cudaGetDevice(&dev));
    {
      std::lock_guard<std::mutex> g(lock());
      auto& fl = free_list(dev);
      if (!fl.empty()) { s = fl.back(); fl.pop_back(); return; }
    }
    PTZ_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  }
  ~StreamHolder() {
    if (!s) return;
    cudaStreamSynchronize(s);
    std::lock_guard<std::mutex> g(lock());
    free_list(dev).push_back(s);
  }
};

// keep freed device memory in the default pool: the many BA calls of one IBA run (and bench.py's repeated solves)
// then allocate in microseconds instead of paying cudaMalloc/cudaFree every time
inline void enable_memory_pool() {
  static bool done = false;
  if (done) return;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thresh = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
  }
  done = true;
}

// Device buffer.  alloc(count, stream) with a non-null stream is stream-ordered (cudaMallocAsync / cudaFreeAsync on that
// stream); without a stream it is a plain cudaMalloc.
// Exact-size cache of device blocks in front of cudaMallocAsync, keyed by (stream, bytes).  A solve allocates ~80 buffers (1.3 GB at
// cfg 4) and the next solve of an IBA run / of bench.py's end-to-end leg asks for the very same sizes; going back to the CUDA pool
// each time was measured to stall now and then for tens to hundreds of milliseconds (the pool re-growing after fragmentation,
// PTZ_SETUP_DEBUG).  A block is only ever handed back to work on the SAME stream it was released on, so reuse is stream-ordered
// exactly like cudaFreeAsync / cudaMallocAsync.  Above the limit (PTZ_BLOCK_CACHE_MB, default 32 GB) everything cached is released.
struct BlockCache {
  typedef std::pair<cudaStream_t, size_t> Key;
  static std::mutex& lock() { static std::mutex m; return m; }
  static std::map<Key, std::vector<void*>>& blocks() { static std::map<Key, std::vector<void*>> b; return b; }
  static size_t& cached() { static size_t c = 0; return c; }
  static size_t limit() {
    static size_t l = []() { const char* e = getenv("PTZ_BLOCK_CACHE_MB"); return (size_t)(e ? atoll(e) : 32768) << 20; }();
    return l;
  }
  static size_t round(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
  static void* get(cudaStream_t s, size_t bytes) {
    std::lock_guard<std::mutex> g(lock());
    auto it = blocks().find(Key(s, bytes));
    if (it == blocks().end() || it->second.empty()) return nullptr;
    void* p = it->second.back();
    it->second.pop_back();
    cached() -= bytes;
    return p;
  }
  static void put(cudaStream_t s, void* p, size_t bytes) {
    std::lock_guard<std::mutex> g(lock());
    if (cached() + bytes > limit()) {
      for (auto& kv : blocks())
        for (void* q : kv.second) cudaFree(q);  // (rare; synchronising, and safe even if the caller's stream no longer exists)
      blocks().clear();
      cached() = 0;
    }
    if (bytes > limit()) { cudaFreeAsync(p, s); return; }
    blocks()[Key(s, bytes)].push_back(p);
    cached() += bytes;
  }
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t as = nullptr;
  bool async = false;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) { if (async) BlockCache::put(as, p, BlockCache::round(n * sizeof(T))); else cudaFree(p); }
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count, cudaStream_t s = nullptr) {
    release();
    n = count;
    as = s;
    async = (s != nullptr);
    if (count) {
      if (async) {
        const size_t bytes = BlockCache::round(count * sizeof(T));
        p = reinterpret_cast<T*>(BlockCache::get(s, bytes));
        if (!p) PTZ_CUDA(cudaMallocAsync((void**)&p, bytes, s));
      } else {
        PTZ_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
      }
    }
  }
  void zero(cudaStream_t s) {
    if (n) PTZ_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    if (count > n) alloc(count, s);
    if (count) PTZ_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void download(T* h, size_t count, cudaStream_t s) const {
    if (count) PTZ_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

#if defined(__CUDACC__)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;  // valid in lane 0
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
// Warp reduction of NP values per lane (NP a power of two <= 32) that SCATTERS the totals: at every step a lane keeps one
// half of its values and adds the partner's copy of that half, so NP-1 (+ log2(32/NP)) shuffles do what 5*NP would in
// NP separate butterflies.  Returns the warp total of element  lane / (32/NP)  (every element ends up in 32/NP lanes).
// Fixed order -> bit-reproducible.
template <int C, int O>
struct WarpRS {
  static __device__ __forceinline__ void run(double* v, int lane) {
    if (C > 1) {
      constexpr int H = C > 1 ? C / 2 : 1;
      const bool up = (lane & O) != 0;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const double send = up ? v[i] : v[i + H];
        const double keep = up ? v[i + H] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
      }
      WarpRS<H, O / 2>::run(v, lane);
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], O);
      WarpRS<1, O / 2>::run(v, lane);
    }
  }
};
template <int C>
struct WarpRS<C, 0> {
  static __device__ __forceinline__ void run(double*, int) {}
};
template <int NP>
__device__ __forceinline__ double warp_reduce_scatter(double (&v)[NP], int lane) {
  static_assert(NP >= 1 && NP <= 32 && (NP & (NP - 1)) == 0, "NP must be a power of two <= 32");
  WarpRS<NP, 16>::run(v, lane);
  return v[0];
}
// block-wide sum of NV values per thread; result valid in thread 0.  sm must hold NV * (blockDim/32) doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sm) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = warp_sum(v[i]);
    if (lane == 0) sm[wid * NV + i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0;
      for (int w = 0; w < nw; ++w) s += sm[w * NV + i];
      v[i] = s;
    }
  }
  __syncthreads();
}
// Ampere-style asynchronous copies global -> shared (LDGSTS): no register staging, completion tracked per thread in groups
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {  // 16 bytes, L2 only (streamed / gathered once)
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- Blackwell/Hopper bulk asynchronous copies (TMA engine, 1-D: cp.async.bulk) and the mbarrier that tracks them.  One elected
// thread issues a copy of a whole contiguous block (16-byte aligned, a multiple of 16 bytes); nothing goes through the LSU pipe or
// through registers.  Loads complete on an mbarrier (expect_tx bytes), stores in bulk groups (commit / wait_group.read).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
// shared -> global; the shared-memory image must have been written before a fence.proxy.async (bulk_store_fence) + barrier
__device__ __forceinline__ void bulk_store(void* gmem, const void* smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes) : "memory");
}
// L2 residency: data that is streamed once (record blocks) is marked evict-first, so that it does not push the small gathered
// arrays of the same kernel (track records, per-track Cholesky factors: 25-32 MB) out of the 126 MB L2
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_store_hint(void* gmem, const void* smem, unsigned bytes, unsigned long long pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async16_hint(void* smem, const void* gmem, unsigned long long pol) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }  // sources may be reused
// 256-bit global loads (sm_100: LDG.E.256): one instruction and one L1 wavefront per lane for a 32-byte, 32-byte-aligned piece of a
// gathered record instead of two.  nc: read-only data of an earlier kernel.
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// block-wide sum of NV values per thread through shared memory (NT threads): NV stores, NT/8 loads per partial and 3 shuffle
// steps instead of 10 shuffles per value.  Result valid in thread 0.  sm must hold NV*(NT+4)+NV doubles.  Fixed order.
template <int NV, int NT>
__device__ __forceinline__ void block_sum_sm(double (&v)[NV], double* sm) {
  constexpr int LD = NT + 4;  // row padding: the four rows a warp sums at once fall into different banks
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NV; ++i) sm[i * LD + t] = v[i];
  __syncthreads();
  double* res = sm + NV * LD;
  for (int u0 = 0; u0 < NV * 8; u0 += NT) {  // trip count uniform over the block: every lane reaches the shuffles
    const int u = u0 + t;
    const bool valid = u < NV * 8;
    const int val = valid ? (u >> 3) : 0, seg = u & 7;
    double sacc = 0;
    if (valid) {
#pragma unroll
      for (int k = 0; k < NT / 8; ++k) sacc += sm[val * LD + k * 8 + seg];
    }
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
    if (valid && seg == 0) res[val] = sacc;
  }
  __syncthreads();
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = res[i];
  }
  __syncthreads();
}
#endif

}  // namespace ptz
