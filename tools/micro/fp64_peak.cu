// micro-benchmark: fp64 FMA throughput / latency and shared-memory fp64 read-modify-write rate on one CTA per SM
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k_fma(double* out, int iters, int chains) {
  double a[8];
  for (int c = 0; c < 8; ++c) a[c] = 1.0 + threadIdx.x * 1e-9 + c;
  const double m = 1.0000001, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c < chains) a[c] = fma(a[c], m, b);
  }
  long long t1 = clock64();
  double s = 0; for (int c = 0; c < 8; ++c) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
__global__ void __launch_bounds__(512, 1) k_smem(double* out, int iters) {
  __shared__ double sm[4096];
  for (int e = threadIdx.x; e < 4096; e += blockDim.x) sm[e] = e;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    for (int e = threadIdx.x; e < 4096; e += blockDim.x) sm[e] += 1e-3 * sm[(e * 7 + i) & 4095];
    __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 512 * 8); double h;
  for (int nt : {32, 128, 512}) for (int chains : {1, 2, 4, 8}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 100000;
    k_fma<<<148, nt>>>(out, 1000, chains);
    cudaEventRecord(e0); k_fma<<<148, nt>>>(out, iters, chains); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    double flops = 2.0 * 148 * nt * (double)iters * chains;
    printf("threads %3d chains %d: %.1f cycles/iter  (%.2f cycles per dependent FMA)  %.2f TFLOP/s  clock %.0f MHz\n", nt, chains, h / iters, h / iters, flops / (ms * 1e-3) / 1e12, h / (ms * 1e-3) / 1e6);
  }
  k_smem<<<148, 512>>>(out, 100); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("smem RMW of 4096 doubles by 512 threads: %.0f cycles per pass (8 per thread)\n", h / 100);
  return 0;
}
