// Latency of the primitives the cluster kernel's time step is made of (B200): one warp of one CTA
// (cluster of 2) runs a dependent chain of each primitive; cycles per repetition by clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench4 tools/microbench4.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void st_async(uint32_t ra, uint32_t rb, double a, double b) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(ra), "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(rb) : "memory");
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void cl_sync() { asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ double dmax(double a, double b) { return (b > a) ? b : a; }
__device__ __forceinline__ void nbar(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }

constexpr int REP = 64;

__global__ void __cluster_dims__(2, 1, 1) bench(long long* out, double* gbuf, double seed) {
  __shared__ double tile[1024];
  __shared__ double xbuf[8];
  __shared__ unsigned long long bars[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t sm_tile = (uint32_t)__cvta_generic_to_shared(tile);
  const uint32_t sm_x = (uint32_t)__cvta_generic_to_shared(xbuf);
  const uint32_t sm_b = (uint32_t)__cvta_generic_to_shared(bars);
  for (int i = tid; i < 1024; i += blockDim.x) tile[i] = (i + 0.5) / 1024.0;
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(sm_b + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  cl_sync();
  long long t0, t1;
  int slot = 0;
  double v = seed + lane;
#define REPORT() do { if (tid == 0 && rank == 0) out[slot] = (t1 - t0) / REP; ++slot; } while (0)
  if (warp == 0) {
    // 0: dependent DFMA
    t0 = clock64();
    for (int i = 0; i < REP; ++i) v = fma(v, 1.0000001, 1e-9);
    t1 = clock64(); REPORT();
    // 1: dependent DSETP + select (dmax against a changing value)
    double w = v * 0.5;
    t0 = clock64();
    for (int i = 0; i < REP; ++i) { w = dmax(w, v - (double)i); v = w + 1e-3; }
    t1 = clock64(); REPORT();
    // 2: 64-bit shuffle + add (one scan level)
    t0 = clock64();
    for (int i = 0; i < REP; ++i) v += __shfl_xor_sync(0xffffffffu, v, 1 + (i & 15));
    t1 = clock64(); REPORT();
    // 3: dependent shared-memory load (binary-search level: min, address, LDS.64, compare, select)
    int lo = 0;
    const double pos = 0.37 + 1e-3 * lane;
    t0 = clock64();
    for (int i = 0; i < REP; ++i) {
      const int step = 512 >> (i % 10);
      if (step == 512) lo = 0;
      const int probe = min(lo + step, 1024);
      lo = (lds_f64(sm_tile - 8u + 8u * probe) <= pos) ? probe : lo;
    }
    t1 = clock64(); REPORT();
    v += lo;
    // 4: IEEE fp64 division
    t0 = clock64();
    for (int i = 0; i < REP; ++i) v = 1.0 / (v + 3.0);
    t1 = clock64(); REPORT();
    // 5: mbarrier: expect_tx + st.async to SELF + try_wait
    uint32_t par = 0;
    t0 = clock64();
    for (int i = 0; i < REP; ++i) {
      if (lane == 0) {
        mbar_expect(sm_b, 16);
        st_async(mapa(sm_x, rank), mapa(sm_b, rank), v, v);
        while (!mbar_try(sm_b, par)) {}
      }
      par ^= 1;
      __syncwarp();
    }
    t1 = clock64(); REPORT();
    // 6: the same with test_wait polling
    t0 = clock64();
    for (int i = 0; i < REP; ++i) {
      if (lane == 0) {
        mbar_expect(sm_b, 16);
        st_async(mapa(sm_x, rank), mapa(sm_b, rank), v, v);
        while (!mbar_test(sm_b, par)) {}
      }
      par ^= 1;
      __syncwarp();
    }
    t1 = clock64(); REPORT();
    // 7: fence.acq_rel.cluster with nothing outstanding
    t0 = clock64();
    for (int i = 0; i < REP; ++i) { fence_cluster(); v += 1.0; }
    t1 = clock64(); REPORT();
    // 8: global store + fence.acq_rel.cluster
    t0 = clock64();
    for (int i = 0; i < REP; ++i) { gbuf[tid + 32 * (i & 7)] = v; fence_cluster(); v += 1.0; }
    t1 = clock64(); REPORT();
    // 9: global store + fence.acq_rel.gpu
    t0 = clock64();
    for (int i = 0; i < REP; ++i) { gbuf[tid + 32 * (i & 7)] = v; fence_gpu(); v += 1.0; }
    t1 = clock64(); REPORT();
  } else {
    slot = 10;
  }
  __syncthreads();
  cl_sync();
  // 10: ping-pong between the two CTAs of the cluster: st.async to the PEER + try_wait on own barrier
  //     (one round trip = two one-way exchanges)
  if (warp == 0) {
    uint32_t par = 0;
    const uint32_t peer = rank ^ 1u;
    t0 = clock64();
    for (int i = 0; i < REP; ++i) {
      if (lane == 0) {
        if (rank == 0) {
          mbar_expect(sm_b + 8, 16);
          st_async(mapa(sm_x, peer), mapa(sm_b + 8, peer), v, v);
          while (!mbar_try(sm_b + 8, par)) {}
        } else {
          mbar_expect(sm_b + 8, 16);
          while (!mbar_try(sm_b + 8, par)) {}
          st_async(mapa(sm_x, peer), mapa(sm_b + 8, peer), v, v);
        }
      }
      par ^= 1;
      __syncwarp();
    }
    t1 = clock64(); REPORT();
  } else ++slot;
  __syncthreads();
  cl_sync();
  // 11: named barrier of all warps of the CTA, back to back
  t0 = clock64();
  for (int i = 0; i < REP; ++i) nbar(blockDim.x);
  t1 = clock64(); REPORT();
  // 12: named barrier when warp 0 does ~200 cycles of work between barriers and the others idle
  t0 = clock64();
  for (int i = 0; i < REP; ++i) {
    if (warp == 0) { for (int k = 0; k < 25; ++k) v = fma(v, 1.0000001, 1e-9); }
    nbar(blockDim.x);
  }
  t1 = clock64(); REPORT();
  // 13: 5-level fp64 warp max + 5-level inclusive scan (phase A reductions of one warp)
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < REP; ++i) {
      double m = v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = dmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      double s = v - m;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const double n = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += n; }
      v = s * 1e-3 + m;
    }
    t1 = clock64(); REPORT();
  } else ++slot;
  // 14: cluster barrier (all threads of both CTAs)
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < REP; ++i) cl_sync();
  t1 = clock64(); REPORT();
  if (v == 123.456) out[63] = 1;
}

int main() {
  long long* out; double* gbuf;
  CK(cudaMalloc(&out, 64 * 8)); CK(cudaMalloc(&gbuf, 4096 * 8));
  CK(cudaMemset(out, 0, 64 * 8));
  const char* names[] = {"dependent DFMA", "dependent DSETP+SEL+DADD", "64-bit SHFL + DADD", "search level (LDS.64 + compare + select)",
                         "fp64 division (1/x)", "mbarrier expect + st.async(self) + try_wait", "... with test_wait polling",
                         "fence.acq_rel.cluster (idle)", "st.global + fence.acq_rel.cluster", "st.global + fence.acq_rel.gpu",
                         "DSMEM ping-pong round trip (st.async peer + try_wait) x2 one-way", "bar.sync all warps back to back",
                         "bar.sync, warp 0 works 25 DFMA, others idle", "warp max (5 lvl) + inclusive scan (5 lvl), fp64", "barrier.cluster arrive+wait"};
  for (int threads : {64, 640}) {
    bench<<<2, threads>>>(out, gbuf, 1.5);
    CK(cudaDeviceSynchronize());
    long long h[64];
    CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("== %d threads per CTA, cluster of 2\n", threads);
    for (int i = 0; i < 15; ++i) printf("  %-70s %6lld cycles\n", names[i], h[i]);
  }
  return 0;
}
