// Micro-test: setmaxnreg.dec/inc for a 640-thread CTA launched with 96 registers per thread
// (512 threads -> 88, one warpgroup -> 128); prints the launch register count.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__device__ __forceinline__ double heavy(const double* in, int base) {
  double v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = in[base + i * 7];
#pragma unroll 1
  for (int r = 0; r < 50; ++r) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = v[i] * 1.0000001 + v[(i + 1) % N] * 1e-9;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) s += v[i];
  return s;
}
__global__ void __launch_bounds__(640, 1) k(const double* in, double* out) {
  const int warp = threadIdx.x >> 5;
  if (warp >= 16) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    out[blockIdx.x * 640 + threadIdx.x] = heavy<48>(in, threadIdx.x);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    out[blockIdx.x * 640 + threadIdx.x] = heavy<30>(in, threadIdx.x);
  }
}
int main() {
  double *in, *out;
  cudaMalloc(&in, 8 * 4096); cudaMemset(in, 0, 8 * 4096);
  cudaMalloc(&out, 8 * 640 * 148);
  k<<<148, 640>>>(in, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  cudaFuncAttributes a; cudaFuncGetAttributes(&a, k);
  printf("numRegs %d localSize %zu\n", a.numRegs, a.localSizeBytes);
  return 0;
}
