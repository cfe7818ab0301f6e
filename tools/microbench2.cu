// Load-path latency comparison for the lineage walk (B200).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <int MODE>
__device__ __forceinline__ int ld(const int* p, unsigned long long pol) {
  int v;
  if (MODE == 0) v = __ldcg(p);
  else if (MODE == 1) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  else if (MODE == 2) asm volatile("ld.relaxed.gpu.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
  else if (MODE == 3) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
  else v = *p;
  return v;
}
template <int MODE>
__global__ void k_chase(const int* __restrict__ next, int* out, long long* cyc, int iters) {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  int j = threadIdx.x + blockIdx.x * blockDim.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) j = ld<MODE>(next + j, pol);
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x + blockIdx.x * blockDim.x] = j;
}
template <int MODE>
int run(const char* name, const int* nxt, int* iout, long long* cyc) {
  long long hc;
  for (int cfg = 0; cfg < 3; ++cfg) {
    int blocks = cfg == 0 ? 1 : 148, threads = cfg == 2 ? 512 : 32;
    for (int rep = 0; rep < 2; ++rep) { k_chase<MODE><<<blocks, threads>>>(nxt, iout, cyc, 500); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    printf("%-44s %3d x %3d threads: %6.0f cycles/load\n", name, blocks, threads, hc / 500.0);
  }
  return 0;
}
int main() {
  long long* cyc; int* nxt; int* iout;
  CK(cudaMalloc(&cyc, 64));
  const int M = 1 << 23;  // 32 MB table (L2 resident)
  CK(cudaMalloc(&nxt, M * 4ull)); CK(cudaMalloc(&iout, 1 << 22));
  int* h = (int*)malloc(M * 4ull);
  for (int i = 0; i < M; ++i) h[i] = (int)(((long long)i * 7919 + 12345) % M);
  CK(cudaMemcpy(nxt, h, M * 4ull, cudaMemcpyHostToDevice));
  run<0>("ld.global.cg", nxt, iout, cyc);
  run<1>("ld.relaxed.gpu", nxt, iout, cyc);
  run<2>("ld.relaxed.gpu.L2::cache_hint(evict_last)", nxt, iout, cyc);
  run<3>("ld.L1::no_allocate.L2::cache_hint", nxt, iout, cyc);
  run<4>("ld.global (default, L1 allocating)", nxt, iout, cyc);
  return 0;
}
