// micro-benchmark: symmetric packed rank-1 update as in scp_device.inl (sp_rank1), shared vs global memory
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ size_t sp_row(int i) { return ((size_t)i * (size_t)(i + 1)) / 2; }
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* Ag, int n, int reps, long long* out) {
  extern __shared__ double sm[];
  double* A = MODE == 0 ? sm + 1024 : Ag + (size_t)blockIdx.x * 20000;
  double* v = sm;
  const int nt = blockDim.x;
  const long long total = (long long)sp_row(n);
  for (long long e = threadIdx.x; e < total; e += nt) A[e] = 1.0 + 1e-3 * e;
  for (int i = threadIdx.x; i < n; i += nt) v[i] = 0.01 * i;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    const int skip = r % n;
    const double s = -1e-6;
    if (MODE <= 1) {
      long long e = threadIdx.x;
      if (e < total) {
        int i = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
        while ((long long)sp_row(i) > e) --i;
        while ((long long)sp_row(i + 1) <= e) ++i;
        int j = (int)(e - (long long)sp_row(i));
        for (; e < total; e += nt) {
          if (i != skip && j != skip) A[e] += s * v[i] * v[j];
          j += nt;
          while (j > i) { j -= i + 1; ++i; }
        }
      }
    } else {
      // row-per-warp variant: warp w handles rows w, w+16, ...; lanes stride over j
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nt >> 5;
      for (int i = w; i < n; i += nw) {
        if (i == skip) continue;
        const double vi = s * v[i];
        double* row = A + sp_row(i);
        for (int j = lane; j <= i; j += 32) if (j != skip) row[j] += vi * v[j];
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / reps;
}
int main() {
  double* Ag; long long* out; long long h[4];
  cudaMalloc(&Ag, 148 * 20000 * 8); cudaMalloc(&out, 148 * 8);
  const int smem = (1024 + 18000) * 8;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n : {64, 110, 150, 186}) {
    k<0><<<148, 512, smem>>>(Ag, n, 200, out); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); long long a = h[0];
    k<1><<<148, 512, smem>>>(Ag, n, 200, out); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); long long b = h[0];
    k<2><<<148, 512, smem>>>(Ag, n, 200, out); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); long long c = h[0];
    printf("n=%d cycles per rank-1: smem %lld  global %lld  row-per-warp(smem) %lld  err=%s\n", n, a, b, c, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
