// Micro-benchmark: round-trip latency of a warp-wide 512-byte ld.global.cg (L2-miss, random row)
// with W warps per SM doing the same, and of the rolling 4-deep prefetch pattern the sweep uses.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void k_lat(const double* buf, size_t nrows512, int n, double* out, long long* cyc) {
  const int lane = threadIdx.x & 31;
  unsigned h = hash(blockIdx.x * 1024 + threadIdx.x / 32 + 1);
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    h = hash(h + (unsigned)(acc != 12345.0));  // dependent on the previous load
    const double2 v = __ldcg((const double2*)(buf + (size_t)(h % nrows512) * 64) + lane);
    acc += v.x + v.y;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / n;
}
// D independent 512B loads in flight per warp (rolling), like the sweep's item loop
template <int D>
__global__ void k_roll(const double* buf, size_t nrows512, int n, double* out, long long* cyc) {
  const int lane = threadIdx.x & 31;
  unsigned h = hash(blockIdx.x * 1024 + threadIdx.x / 32 + 1);
  double2 v[D];
  for (int d = 0; d < D; ++d) { h = hash(h); v[d] = __ldcg((const double2*)(buf + (size_t)(h % nrows512) * 64) + lane); }
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; i += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      acc += v[d].x * v[d].y;
      h = hash(h);
      v[d] = __ldcg((const double2*)(buf + (size_t)(h % nrows512) * 64) + lane);
    }
  }
  long long t1 = clock64();
  for (int d = 0; d < D; ++d) acc += v[d].x;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0);
}
int main() {
  const size_t bytes = 400ull << 20;
  double* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
  double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  const size_t nrows = bytes / 512;
  for (int threads : {32, 128, 512, 1024}) {
    k_lat<<<148, threads>>>(buf, nrows, 2000, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps/SM %2d: dependent 512B ld.cg round trip %lld cycles\n", threads / 32, h[0]);
  }
  const int n = 4096;
  for (int threads : {512, 1024}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    cudaEventRecord(e0); k_roll<4><<<148, threads>>>(buf, nrows, n, out, cyc); cudaEventRecord(e1); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("warps/SM %2d rolling depth 4: %.0f cycles per 512B load per warp, %.0f GB/s\n", threads / 32, (double)h[0] / n, 148.0 * (threads / 32) * n * 512 / ms / 1e6);
    cudaEventRecord(e0); k_roll<8><<<148, threads>>>(buf, nrows, n, out, cyc); cudaEventRecord(e1); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("warps/SM %2d rolling depth 8: %.0f cycles per 512B load per warp, %.0f GB/s\n", threads / 32, (double)h[0] / n, 148.0 * (threads / 32) * n * 512 / ms / 1e6);
    cudaEventRecord(e0); k_roll<16><<<148, threads>>>(buf, nrows, n, out, cyc); cudaEventRecord(e1); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("warps/SM %2d rolling depth 16: %.0f cycles per 512B load per warp, %.0f GB/s\n", threads / 32, (double)h[0] / n, 148.0 * (threads / 32) * n * 512 / ms / 1e6);
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
