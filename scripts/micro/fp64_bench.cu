// Micro-benchmark: FP64 latency / throughput per SM on this GPU (decides how much arithmetic the
// sweep can afford per byte).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dep(double* out, int n, long long* cyc) {
  double a = threadIdx.x * 1e-9 + 1.0, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, b, c);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void k_ilp(double* out, int n, long long* cyc) {
  double a[ILP];
  for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x * 1e-9 + j;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], b, c);
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < ILP; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_log(double* out, int n, long long* cyc) {
  double a = threadIdx.x + 2.0, s = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { s += log(a); a += 1.0; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_f32(float* out, int n, long long* cyc) {
  float a[8];
  for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-9f + j;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], 1.0000001f, 1e-9f);
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  const int n = 4096;
  for (int threads : {32, 128, 512, 1024}) {
    k_dep<<<148, threads>>>(out, n, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %4d dependent DFMA: %.2f cycles/op per warp\n", threads, (double)h[0] / n);
    k_ilp<8><<<148, threads>>>(out, n, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double per_sm = (double)threads * 8 * n / h[0];
    printf("threads %4d ILP8 DFMA: %.2f cycles per warp-instr; %.1f DFMA lanes/clk/SM\n", threads, (double)h[0] / (n * 8), per_sm);
    k_log<<<148, threads>>>(out, n, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %4d log(): %.1f cycles per call per warp (%.2f calls/clk/SM)\n", threads, (double)h[0] / n, (double)threads * n / h[0]);
    k_f32<<<148, threads>>>((float*)out, n, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %4d ILP8 FFMA: %.1f FFMA lanes/clk/SM\n", threads, (double)threads * 8 * n / h[0]);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
