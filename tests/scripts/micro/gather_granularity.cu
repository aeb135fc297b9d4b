// How many DRAM bytes does a random 64-byte gather cost on B200?  (decides whether splitting the 128-byte residual/Jacobian record
// into [r,E | F] arrays would pay for k_track_accum).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_granularity
// gather_granularity.cu ; run under: ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_granularity
#include <cstdio>
#include <cuda_runtime.h>

template <int STRIDE_BYTES, int READ_BYTES>
__global__ void k_gather(const char* __restrict__ base, const int* __restrict__ idx, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2* p = reinterpret_cast<const double2*>(base + (size_t)idx[i] * STRIDE_BYTES);
  double s = 0;
#pragma unroll
  for (int k = 0; k < READ_BYTES / 16; ++k) { const double2 v = p[k]; s += v.x + v.y; }
  out[i] = s;
}

int main() {
  const int n = 2000000;
  char* base; int* idx; double* out;
  cudaMalloc(&base, (size_t)n * 128);
  cudaMalloc(&idx, n * 4);
  cudaMalloc(&out, n * 8);
  cudaMemset(base, 0, (size_t)n * 128);
  int* h = new int[n];
  unsigned long long x = 88172645463325252ull;
  for (int i = 0; i < n; ++i) h[i] = i;
  for (int i = n - 1; i > 0; --i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; int j = (int)(x % (unsigned long long)(i + 1)); int t = h[i]; h[i] = h[j]; h[j] = t; }
  cudaMemcpy(idx, h, n * 4, cudaMemcpyHostToDevice);
  char* flush; cudaMalloc(&flush, 512u << 20);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(flush, rep, 512u << 20);
    k_gather<128, 64><<<(n + 127) / 128, 128>>>(base, idx, n, out);   // 64 B of a 128-byte record (today's r+E gather)
    cudaMemset(flush, rep + 2, 512u << 20);
    k_gather<64, 64><<<(n + 127) / 128, 128>>>(base, idx, n, out);    // 64-byte records packed (split layout)
    cudaMemset(flush, rep + 4, 512u << 20);
    k_gather<128, 128><<<(n + 127) / 128, 128>>>(base, idx, n, out);  // whole 128-byte record
    cudaMemset(flush, rep + 6, 512u << 20);
    k_gather<96, 96><<<(n + 127) / 128, 128>>>(base, idx, n, out);    // 96-byte records (What, NCL = 4)
    cudaMemset(flush, rep + 8, 512u << 20);
    k_gather<128, 96><<<(n + 127) / 128, 128>>>(base, idx, n, out);   // 96 B of a 128-byte-aligned slot (padded What)
  }
  cudaDeviceSynchronize();
  printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
