// How long does the shared-memory CDF search of the particle filter take per quad? (B200)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ int ub(uint32_t ta, double pos, int len, int top) {
  int lo = 0;
  for (int step = top; step > 0; step >>= 1) { const int probe = lo + step; if (probe <= len && lds_f64(ta + 8u * probe) <= pos) lo = probe; }
  return lo;
}
__device__ __forceinline__ int np_(uint32_t ta, double pos, int lo, int len, int top) {
  const int wend = min(lo + 8, len);
  if (lds_f64(ta + 8u * wend) <= pos) return (wend == len) ? len : ub(ta, pos, len, top);
  int r = lo;
#pragma unroll
  for (int step = 4; step > 0; step >>= 1) { const int probe = r + step; if (probe <= wend && lds_f64(ta + 8u * probe) <= pos) r = probe; }
  return r;
}
// variant B: clamped probe index instead of a predicated load; variant C: + integer compare
__device__ __forceinline__ int ubB(uint32_t ta, double pos, int len, int top) {
  int lo = 0;
  for (int step = top; step > 0; step >>= 1) { const int probe = min(lo + step, len); const double v = lds_f64(ta + 8u * probe); lo = (v <= pos) ? probe : lo; }
  return lo;
}
__device__ __forceinline__ long long lds_s64(uint32_t a) { long long v; asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ int ubC(uint32_t ta, double pos, int len, int top) {
  int lo = 0;
  const long long ipos = __double_as_longlong(pos);
  for (int step = top; step > 0; step >>= 1) { const int probe = min(lo + step, len); const long long v = lds_s64(ta + 8u * probe); lo = (v <= ipos) ? probe : lo; }
  return lo;
}
__device__ __forceinline__ int ubD(uint32_t ta, double pos, int len, int top) {  // fully unrolled 13 steps
  int lo = 0;
#pragma unroll
  for (int step = 4096; step > 0; step >>= 1) { const int probe = min(lo + step, len); const double v = lds_f64(ta + 8u * probe); lo = (v <= pos) ? probe : lo; }
  return lo;
}
template <int V>
__global__ void k2(int len, int quads_per_thread, long long* cyc, int* out) {
  extern __shared__ double tile[];
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    unsigned h = (i * 2654435761u) >> 8;
    tile[i] = (i + 0.5 + 0.9 * ((h & 0xffff) / 65536.0 - 0.5)) / len;
  }
  __syncthreads();
  const uint32_t ta = (uint32_t)__cvta_generic_to_shared(tile) - 8u;
  int top = 1; while (top * 2 <= len) top *= 2;
  int acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < quads_per_thread; ++it) {
    const int q = it * blockDim.x + threadIdx.x;
    const double base = (4.0 * q + 0.3) / (4.0 * quads_per_thread * blockDim.x);
    acc += (V == 0) ? ubB(ta, base, len, top) : (V == 1) ? ubC(ta, base, len, top) : ubD(ta, base, len, top);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
__global__ void k(int len, int quads_per_thread, long long* cyc, int* out) {
  extern __shared__ double tile[];
  // CDF with random-ish increments (mean 1/len)
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    unsigned h = (i * 2654435761u) >> 8;
    tile[i] = (i + 0.5 + 0.9 * ((h & 0xffff) / 65536.0 - 0.5)) / len;
  }
  __syncthreads();
  const uint32_t ta = (uint32_t)__cvta_generic_to_shared(tile) - 8u;
  int top = 1; while (top * 2 <= len) top *= 2;
  int acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < quads_per_thread; ++it) {
    const int q = it * blockDim.x + threadIdx.x;
    const double base = (4.0 * q + 0.3) / (4.0 * quads_per_thread * blockDim.x);
    const double d = 1.0 / (4.0 * quads_per_thread * blockDim.x);
    int lo = ub(ta, base, len, top);
    acc += lo;
    if (MODE == 1) {
      for (int e = 1; e < 4; ++e) { lo = np_(ta, base + e * d, min(lo, len - 1), len, top); acc += lo; }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  long long* cyc; int* out; long long h[2];
  CK(cudaMalloc(&cyc, 64)); CK(cudaMalloc(&out, 148 * 1024 * 4));
  const int len = 7085;
  CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
  CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
  for (int threads = 32; threads <= 512; threads *= 4) {
    const int qpt = len / 4 / threads + 1;
    for (int rep = 0; rep < 2; ++rep) { k<0><<<148, threads, len * 8>>>(len, qpt, cyc, out); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("%3d threads: full descent only: %6.0f cycles per search (thread 0), %6.0f incl. slowest warp\n", threads, (double)h[0] / qpt, (double)h[1] / qpt);
    for (int rep = 0; rep < 2; ++rep) { k<1><<<148, threads, len * 8>>>(len, qpt, cyc, out); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("%3d threads: descent + 3 merges  : %6.0f cycles per quad   (thread 0), %6.0f incl. slowest warp\n", threads, (double)h[0] / qpt, (double)h[1] / qpt);
  }
  CK(cudaFuncSetAttribute(k2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
  CK(cudaFuncSetAttribute(k2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
  CK(cudaFuncSetAttribute(k2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
  const int qpt = len / 4 / 512 + 1;
  for (int rep = 0; rep < 2; ++rep) { k2<0><<<148, 512, len * 8>>>(len, qpt, cyc, out); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost)); printf("clamped probe            : %6.0f cycles per search\n", (double)h[0] / qpt);
  for (int rep = 0; rep < 2; ++rep) { k2<1><<<148, 512, len * 8>>>(len, qpt, cyc, out); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost)); printf("clamped + integer compare: %6.0f cycles per search\n", (double)h[0] / qpt);
  for (int rep = 0; rep < 2; ++rep) { k2<2><<<148, 512, len * 8>>>(len, qpt, cyc, out); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost)); printf("clamped + fully unrolled : %6.0f cycles per search\n", (double)h[0] / qpt);
  return 0;
}
