// Latency / throughput micro-benchmarks used to size the particle-filter kernel (B200, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_dep(double* out, long long* cyc, int iters) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = fma(a, b, 1.0); a = fma(a, b, 1.0); a = fma(a, b, 1.0); a = fma(a, b, 1.0); }
  long long t1 = clock64();
  float f = (float)a;
  for (int i = 0; i < iters; ++i) { f = fmaf(f, 0.5f, 1.0f); f = fmaf(f, 0.5f, 1.0f); f = fmaf(f, 0.5f, 1.0f); f = fmaf(f, 0.5f, 1.0f); }
  long long t2 = clock64();
  double d = a;
  for (int i = 0; i < iters; ++i) { d = 1.0 / (d + 1.5); }
  long long t3 = clock64();
  double e = d;
  for (int i = 0; i < iters; ++i) { e = exp(-e * 0.5); }
  long long t4 = clock64();
  double l = e + 0.3;
  for (int i = 0; i < iters; ++i) { l = log(l + 1.1); }
  long long t5 = clock64();
  double s = l, c;
  for (int i = 0; i < iters; ++i) { sincospi(s * 0.3, &s, &c); s += c; }
  long long t6 = clock64();
  double q = s;
  for (int i = 0; i < iters; ++i) { q = sqrt(q + 2.0); }
  long long t7 = clock64();
  double sh = q;
  for (int i = 0; i < iters; ++i) { sh += __shfl_xor_sync(0xffffffffu, sh, 1); }
  long long t8 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7;
  }
  out[2 + threadIdx.x + blockIdx.x * blockDim.x] = a + f + d + e + l + s + q + sh;
}

// throughput: many independent DFMA chains per thread, many warps
__global__ void k_tput(double* out, long long* cyc, int iters) {
  double a0 = out[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, b = out[1];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, 1.0); a1 = fma(a1, b, 1.0); a2 = fma(a2, b, 1.0); a3 = fma(a3, b, 1.0);
    a4 = fma(a4, b, 1.0); a5 = fma(a5, b, 1.0); a6 = fma(a6, b, 1.0); a7 = fma(a7, b, 1.0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[2 + threadIdx.x + blockIdx.x * blockDim.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// pointer chase through L2 (ld.cg) / smem
__global__ void k_chase(const int* __restrict__ next, int* out, long long* cyc, int iters) {
  __shared__ int sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) j = __ldcg(next + j);
  long long t1 = clock64();
  int k = j & 1023;
  for (int i = 0; i < iters; ++i) k = sm[k];
  long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
  out[threadIdx.x + blockIdx.x * blockDim.x] = j + k;
}

__global__ void k_sync(long long* cyc, int iters, unsigned* ctr) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  long long t1 = clock64();
  for (int i = 0; i < iters; ++i) { __threadfence(); }
  long long t2 = clock64();
  unsigned v = 0;
  for (int i = 0; i < iters; ++i) { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); }
  long long t3 = clock64();
  for (int i = 0; i < iters; ++i) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr + 32), "r"(1u) : "memory"); }
  long long t4 = clock64();
  for (int i = 0; i < iters; ++i) { v += atomicAdd(ctr + 64, 1u); }
  long long t5 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = v; }
}

int main() {
  double* out; long long* cyc; int* nxt; int* iout; unsigned* ctr;
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 1024)); CK(cudaMemset(out, 0, 1 << 24));
  CK(cudaMalloc(&ctr, 4096)); CK(cudaMemset(ctr, 0, 4096));
  const int M = 1 << 24;  // 64 MB of ints: L2 resident after warm-up
  CK(cudaMalloc(&nxt, M * 4ull)); CK(cudaMalloc(&iout, 1 << 22));
  int* h = (int*)malloc(M * 4ull);
  for (int i = 0; i < M; ++i) h[i] = (int)(((long long)i * 7919 + 12345) % M);
  CK(cudaMemcpy(nxt, h, M * 4ull, cudaMemcpyHostToDevice));
  double init[2] = {0.5, 0.25}; CK(cudaMemcpy(out, init, 16, cudaMemcpyHostToDevice));
  long long hc[16]; const int it = 1000;
  k_dep<<<1, 32>>>(out, cyc, it); CK(cudaDeviceSynchronize());
  k_dep<<<1, 32>>>(out, cyc, it); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost));
  printf("dependent-chain latency (cycles/op): DFMA %.1f  FFMA %.1f  1/x(dbl) %.1f  exp %.1f  log %.1f  sincospi %.1f  sqrt %.1f  shfl+dadd %.1f\n",
         hc[0] / (4.0 * it), hc[1] / (4.0 * it), hc[2] / (double)it, hc[3] / (double)it, hc[4] / (double)it, hc[5] / (double)it, hc[6] / (double)it, hc[7] / (double)it);
  for (int threads = 128; threads <= 1024; threads *= 2) {
    k_tput<<<148, threads>>>(out, cyc, it); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DFMA throughput, %4d threads/SM: %.2f DFMA/clk/SM\n", threads, 8.0 * it * threads / hc[0]);
  }
  for (int rep = 0; rep < 2; ++rep) { k_chase<<<1, 32>>>(nxt, iout, cyc, 2000); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost));
  printf("pointer chase: ld.cg (64 MB table) %.0f cycles/load, smem %.1f cycles/load\n", hc[0] / 2000.0, hc[1] / 2000.0);
  for (int threads = 256; threads <= 1024; threads *= 2) {
    k_chase<<<148, threads>>>(nxt, iout, cyc, 200); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost));
    printf("loaded chase, 148 x %4d threads: ld.cg %.0f cycles/level\n", threads, hc[0] / 200.0);
  }
  for (int threads = 32; threads <= 1024; threads *= 4) {
    k_sync<<<1, threads>>>(cyc, 200, ctr); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hc, cyc, 48, cudaMemcpyDeviceToHost));
    printf("%4d threads: __syncthreads %.0f  __threadfence %.0f  ld.acquire.gpu %.0f  red.release %.0f  atomicAdd(ret) %.0f cycles\n",
           threads, hc[0] / 200.0, hc[1] / 200.0, hc[2] / 200.0, hc[3] / 200.0, hc[4] / 200.0);
  }
  k_sync<<<148, 32>>>(cyc, 200, ctr); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hc, cyc, 48, cudaMemcpyDeviceToHost));
  printf("148 CTAs x 32 thr: __threadfence %.0f  ld.acquire.gpu %.0f  red.release %.0f  atomicAdd(ret) %.0f cycles\n",
         hc[1] / 200.0, hc[2] / 200.0, hc[3] / 200.0, hc[4] / 200.0);
  return 0;
}
