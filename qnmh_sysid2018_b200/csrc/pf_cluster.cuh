// Small and medium filters (N <= 16 * 1024) on ONE thread-block cluster: the per-time-step critical
// path never leaves the cluster's shared memory.
//
// The persistent kernel of pf_kernels.cuh hands a generation from the CTAs that produce it to the
// CTAs that own it through L2 (global rings) and synchronises a filter's CTAs with two counter
// exchanges per time step; at N = 1000 .. 65536 that is all latency (7.4 us per time step at
// N = 1000, of which none is bandwidth).  Here a filter is run by a cluster of C = 1..16 CTAs and
//
//   * the parents' states and log-weights live in the OWNER's shared memory (two generations,
//     by parity); a child is delivered by st.async into the owner's buffers with a transaction-count
//     completion on the owner's mbarrier: the owner waits for exactly own * bytes_per_child bytes,
//     so the hand-off is a neighbour-only synchronisation, not a group barrier;
//   * there is ONE all-gather per time step: every CTA sends (local max log-weight, local sum of
//     shifted weights) to every CTA's mbarrier-guarded slot (st.async, 16 bytes).  Weights are
//     shifted by the CTA's own maximum, exp(lw - m_g), and the tile totals are rescaled by
//     exp(m_g - m) when the prefix is formed -- the second exchange of the large-filter kernel (the
//     global maximum before the weights can be formed) disappears;
//   * the ancestor ring of the fixed-lag smoother is distributed shared memory too: a lineage walk is
//     a chain of ld.shared::cluster instead of L2 round trips.  The smoother warps take one job (lag
//     time) per WARP, four jobs in flight per CTA, up to kLag time steps behind the filter warps.
//
// Same results as the large-filter kernel up to the rounding of the per-CTA shift (bit-identical for
// C = 1 up to the scan tree); same Philox counters, so ancestors agree with the oracle away from CDF
// ties.  Reference behaviour restated: standard.py:42-189, resampling.pyx:65-85 (systematic only: the
// stratified / multinomial / fully adapted / distributed variants stay on the general kernel).
#pragma once
#include "pf_kernels.cuh"

namespace pfb200 {

constexpr int kClLag = 6;        // a smoother job may finish up to kClLag time steps behind the filter warps
constexpr int kClMaxPer = 1024;  // parents per CTA (two per filter thread)
constexpr int kClMaxCluster = 16;
constexpr int kClSpecPairs = 256;  // normal pairs the idle filter warps pre-draw per time step (cap <= 256)

// ---- cluster / DSMEM / mbarrier primitives (PTX ISA 8.x, sm_90+) --------------------------------
__device__ __forceinline__ uint32_t cl_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cl_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void cl_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void cl_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool cl_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cl_fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
// remote stores that complete `bytes` on the destination CTA's mbarrier
__device__ __forceinline__ void cl_st_async(uint32_t raddr, uint32_t rbar, double a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(raddr), "l"(__double_as_longlong(a)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cl_st_async(uint32_t raddr, uint32_t rbar, double a, double b) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cl_st_async(uint32_t raddr, uint32_t rbar, float a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(__float_as_uint(a)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cl_st_async(uint32_t raddr, uint32_t rbar, float a, float b) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cl_st_async(uint32_t raddr, uint32_t rbar, int a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(a), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cl_st_async(uint32_t raddr, uint32_t rbar, int a, int b) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "r"(a), "r"(b), "r"(rbar) : "memory");
}
__device__ __forceinline__ int cl_ld_remote_s32(uint32_t raddr) {
  int v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(raddr) : "memory");
  return v;
}
__device__ __forceinline__ void cl_st_release_remote_u32(uint32_t raddr, unsigned v) {
  asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(raddr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned cl_ld_acquire_local_u32(uint32_t addr) {
  unsigned v;
  asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double cl_lds(uint32_t addr, double) { return lds_f64(addr); }
__device__ __forceinline__ float cl_lds(uint32_t addr, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Dynamic shared memory of a cluster CTA (cap = parents per CTA, a power of two):
//   double   tile[cap]              normalised CDF of the CTA's parents
//   Real     xs[2][cap], lws[2][cap] parents' states / log-weights by generation parity
//   int      anc[RAs][cap]          ancestors of the CTA's own index range, ring of RAs generations
//   double   xbuf[2][2 * C]         all-gather slots (m_g, A_g), by exchange parity
//   double   wscale[8]              exp(m_g - m) of the last generations (filter warps -> smoother warps)
//   uint64   bars[4]                mbarriers: all-gather [2], children [2]
//   unsigned prog[C][4]             smoother jobs finished by warp w of CTA c (cumulative)
//   Real     zb[2][256][2]          normals of the children pairs around the CTA's nominal range, pre-drawn by
//                                   the idle filter warps one time step ahead (by parity of the time step)
//   unsigned vis[C]                 last generation (sequence number) whose global ring rows, as written by
//                                   CTA c, are visible cluster-wide
template <typename Real>
struct ClLayout {
  uint32_t tile, xs, lws, anc, xbuf, wscale, bars, prog, vis, zb, bytes;
  __host__ __device__ ClLayout(int cap, int RAs, int C) {
    uint32_t o = 0;
    tile = o; o += 8u * cap;
    xs = o; o += 2u * (uint32_t)sizeof(Real) * cap;
    lws = o; o += 2u * (uint32_t)sizeof(Real) * cap;
    anc = o; o += 4u * (uint32_t)RAs * cap;
    o = (o + 15u) & ~15u;
    xbuf = o; o += 8u * 2u * 2u * C;
    wscale = o; o += 8u * 8u;
    bars = o; o += 8u * 4u;
    prog = o; o += 4u * 4u * C;
    vis = o; o += 4u * C;
    o = (o + 15u) & ~15u;
    zb = o; o += 2u * (uint32_t)kClSpecPairs * 2u * (uint32_t)sizeof(Real);
    bytes = (o + 15u) & ~15u;
  }
};

template <int MODEL, typename Real, bool GENERAL>
struct ClKernel : PfKernel<MODEL, Real, PFB_BLOCK, PFB_ITEMS_SMALL, GENERAL> {
  using Base = PfKernel<MODEL, Real, PFB_BLOCK, PFB_ITEMS_SMALL, GENERAL>;
  using Ops = typename Base::Ops;
  using Base::p; using Base::tid; using Base::lane; using Base::warp; using Base::g; using Base::q;
  using Base::sh_red; using Base::sh_c; using Base::sh_flags; using Base::aborted; using Base::fsync;
  using Base::fsync_or; using Base::dmax; using Base::position; using Base::lower_bound_child;
  using Base::opaque;
#ifdef PFB_PROFILE
  using Base::prof_acc; using Base::prof_t;
#endif
  static constexpr int BLOCK = PFB_BLOCK;
  static constexpr int kWarps = BLOCK / 32;
  static constexpr int kSW = PFB_SMOOTH_THREADS / 32;  // smoother warps = jobs in flight per CTA
  static constexpr int kWalk = 8;                      // lineages a smoother lane walks at once
  static constexpr uint32_t kChildBytes = 2u * (uint32_t)sizeof(Real) + 4u;
  static_assert(kSW == 4, "progress slots are laid out for four smoother warps");

  const int C, cap, lg, RAs;
  const ClLayout<Real> lay;
  const uint32_t sm0;  // shared-window address of the dynamic shared memory
  unsigned xepoch = 0;        // all-gathers issued so far
  unsigned cuse0 = 0, cuse1 = 0;  // completed uses of the two children barriers
  unsigned done_base = 0;     // smoother jobs per warp completed by earlier evaluations (same in all warps: see jobs_of)

  __device__ ClKernel(const DeviceProblem& prob, double* smem, double* red, double* cst, double* sred, unsigned* flags)
      : Base(prob, smem, red, cst, sred, flags), C(prob.cluster), cap(prob.per_cta), lg(prob.lg_per),
        RAs(prob.RA_s), lay(prob.per_cta, prob.RA_s, prob.cluster), sm0(smem_base(smem)) {}

  // jobs of one evaluation taken by smoother warp w: regular jobs k = 0 .. n_reg-1 round robin, and a
  // share of the tail job for every warp
  __device__ __forceinline__ unsigned jobs_of(int w, int n) const { return n > w ? (unsigned)((n - w + kSW - 1) / kSW) : 0u; }

  // ---- all-gather of (a, b) over the cluster (warp 0): lanes r < C send to CTA r.  The payload travels
  //      in the st.async itself: nothing else is ordered by it (see publish_ring_rows).
  __device__ __forceinline__ void xarrive(double a, double b) {
    ++xepoch;
    if (warp == 0) {
      const uint32_t par = xepoch & 1u;
      if (lane == 0) cl_mbar_arrive_expect_tx(sm0 + lay.bars + 8u * par, 16u * (uint32_t)C);
      if (lane < C) {
        cl_st_async(cl_mapa(sm0 + lay.xbuf + 16u * (par * (uint32_t)C + (uint32_t)g), (uint32_t)lane),
                    cl_mapa(sm0 + lay.bars + 8u * par, (uint32_t)lane), a, b);
      }
    }
  }
  // Off the critical path (the last filter warp, after a block barrier that follows the stores): this
  // CTA's global ring rows up to sequence number `seq` are visible to the whole cluster -- a fence
  // (330-480 cycles) and a flag in every CTA, read by the smoother warps before they follow lineages
  // into rows other CTAs wrote.
  __device__ __forceinline__ void publish_ring_rows(unsigned seq) const {
    if (warp == kWarps - 1 && lane < C) {
      cl_fence_cluster();
      cl_st_release_remote_u32(cl_mapa(sm0 + lay.vis + 4u * (uint32_t)g, (uint32_t)lane), seq);
    }
  }
  // lanes < C of a warp wait until every CTA has published `seq`; returns 1 when abandoned
  __device__ __forceinline__ int wait_ring_rows(unsigned seq) const {
    int ab = 0;
    if (lane < C) {
      long long spins = 0;
      while ((int)(cl_ld_acquire_local_u32(sm0 + lay.vis + 4u * (uint32_t)lane) - seq) < 0) {
        if ((++spins & 255) == 0 && (spins > p.spin_limit || ld_relaxed_i32(p.abort_flag) != 0)) {
          atomicExch(p.abort_flag, 1);
          ab = 1;
          break;
        }
      }
    }
    return __any_sync(0xffffffffu, ab) ? 1 : 0;
  }

  // exp(x) for x <= 0 as in exp_nonpos(), without the early-out branch (two of them interleave)
  static __device__ __forceinline__ double exp_nonpos_nb(double x) {
    const double t = x * kExpLn2[0];
    const bool tiny = !(t > -1021.0);
    const double tc = tiny ? 0.0 : t;
    const int k = __double2int_rn(tc);
    const double kf = (double)k;
    double r = fma(kf, kExpLn2[1], tiny ? 0.0 : x);
    r = fma(kf, kExpLn2[2], r);
    double qv = kExpCoef[0];
#pragma unroll
    for (int i = 1; i < 14; ++i) qv = fma(qv, r, kExpCoef[i]);
    const double e = qv * __hiloint2double((k + 1023) << 20, 0);
    return tiny ? ((t == t) ? 0.0 : t) : e;
  }
  static __device__ __forceinline__ double shifted_weight_nb(double logw, double m) { return exp_nonpos_nb(logw - m); }
  static __device__ __forceinline__ double shifted_weight_nb(float logw, double m) { return (double)__expf(logw - (float)m); }

  // exact fp64 maximum over a warp with two integer warp reductions (REDUX) on an order-preserving
  // key instead of five shuffle + compare levels; NaN entries are ignored like dmax() does
  static __device__ __forceinline__ double warp_max_exact(double v) {
    unsigned long long k = (unsigned long long)__double_as_longlong((v == v) ? v : -INFINITY);
    k = (k >> 63) ? ~k : (k | 0x8000000000000000ull);
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
    k = ((unsigned long long)hmax << 32) | lmax;
    k = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)k);
  }

  // A warp waits for a phase of one of its CTA's mbarriers (lane 0 polls, with the dead-man limit).
  // Returns 1 when the wait was abandoned; callers feed that into the next block-wide barrier.
  __device__ __forceinline__ int warp_mbar_wait(uint32_t bar, uint32_t parity) const {
    int ab = 0;
    if (lane == 0) {
      long long spins = 0;
      while (!cl_mbar_try_wait(bar, parity)) {
        if ((++spins & 63) == 0 && (spins > (p.spin_limit >> 4) || ld_relaxed_i32(p.abort_flag) != 0)) {
          atomicExch(p.abort_flag, 1);
          ab = 1;
          break;
        }
      }
    }
    __syncwarp();
    return ab;
  }
  // warp 0: every smoother warp of every CTA has finished its share of the first n_jobs jobs of this
  // evaluation (+ `tail` tail shares): their reads of the ring slots about to be overwritten are done
  __device__ __forceinline__ int wait_smoother_progress(int n_jobs, unsigned tail) const {
    int ab = 0;
    long long spins = 0;
    for (int s = lane; s < kSW * C; s += 32) {
      const unsigned need = done_base + jobs_of(s & (kSW - 1), n_jobs) + tail;
      while ((int)(cl_ld_acquire_local_u32(sm0 + lay.prog + 4u * (uint32_t)s) - need) < 0) {
        if ((++spins & 255) == 0 && (spins > p.spin_limit || ld_relaxed_i32(p.abort_flag) != 0)) {
          atomicExch(p.abort_flag, 1);
          ab = 1;
          break;
        }
      }
    }
    return __any_sync(0xffffffffu, ab) ? 1 : 0;
  }

  // (n + u) / N correctly rounded (resampling.pyx:72).  N a power of two: one multiplication.  Otherwise
  // Markstein's sequence on the correctly rounded reciprocal: q0 = a * (1/N), r = a - q0 * N (exact in the
  // FMA), q = q0 + r * (1/N) -- the IEEE quotient without the slow path of the general division.
  __device__ __forceinline__ double pos_sys(int n, double u) const {
    const double a = __dadd_rn((double)n, u);
    if (this->pow2N) return __dmul_rn(a, this->invN);
    const double q0 = __dmul_rn(a, this->invN);
    const double r = __fma_rn(-q0, (double)p.N, a);
    return __fma_rn(r, this->invN, q0);
  }
  // smallest n in [0, N] with NOT (pos_sys(n) < c)
  __device__ __forceinline__ int first_child_not_below(double c, double u) const {
    const int N = p.N;
    if (!(c == c)) return N;  // NaN CDF (all weights zero): nothing to assign
    const double est = ceil(c * (double)N - u);
    int n = est < 0.0 ? 0 : (est > (double)N ? N : (int)est);
    // the estimate is almost always exact: test its two neighbours side by side before looping
    const double pl = pos_sys(max(n - 1, 0), u), ph = pos_sys(min(n, N - 1), u);
    if ((n == 0 || pl < c) && (n == N || !(ph < c))) return n;
    while (n > 0 && !(pos_sys(n - 1, u) < c)) --n;
    while (n < N && pos_sys(n, u) < c) ++n;
    return n;
  }

  // upper bound in the normalised tile: #{ jl < len : tile[jl] <= pos }, two positions side by side
  __device__ __forceinline__ void search2(uint32_t ta, double pa, double pb, int len, int top, int& la, int& lb) const {
    la = 0;
    lb = 0;
    for (int step = top; step > 0; step >>= 1) {
      const int qa = min(la + step, len), qb = min(lb + step, len);
      const double va = lds_f64(ta + 8u * qa), vb = lds_f64(ta + 8u * qb);
      la = (va <= pa) ? qa : la;
      lb = (vb <= pb) ? qb : lb;
    }
  }

  // ---- one filter evaluation on the filter warps --------------------------------------------------
  // Only the warps that hold parents (two per thread) take part in the reductions of phase A, only the
  // warps that hold children pairs in phase B; the others go from block barrier to block barrier (three
  // per time step).  Warp 0 always takes part: it runs the exchange.
  __device__ void run_filter(int b, int b_idx) {
    const int N = p.N, T = p.T, L = p.L;
    const int ws = p.ws_per_filter ? b : q;
    const size_t NS = (size_t)p.NS;
    Real* xr = reinterpret_cast<Real*>(p.x_ring) + (size_t)(p.xs_per_filter ? b : ws) * p.RX * NS;
    Real* lwr = reinterpret_cast<Real*>(p.logw) + (size_t)ws * p.RW * NS;
    Real* swr = reinterpret_cast<Real*>(p.s_ring) + (size_t)ws * p.RS * NS;
    int* ar = p.anc_ring + (size_t)ws * p.RA * NS;
    const double* obs = p.obs + (p.obs_per_filter ? (size_t)b * (T + 1) : 0);
    const unsigned long long eval_id = p.eval_ids[b];
    const size_t ib = p.inject_per_filter ? (size_t)b : 0;
    const bool inject = GENERAL && p.inject;
    const double* inj_z0 = inject ? p.inj_z0 + ib * N : nullptr;
    const double* inj_u = inject ? p.inj_u + ib * (T + 1) : nullptr;
    const double* inj_z = inject ? p.inj_z + ib * (size_t)(T + 1) * N : nullptr;

    const int j_begin = min(g * cap, N);
    const int j_end = min(j_begin + cap, N);
    const int own = j_end - j_begin;
    const bool last_cta = (j_end == N);
    const int n_act = max(1, (own + 63) >> 6);  // warps that hold parents
    int npow = 1;                               // reduction width over the active warps' partials
    while (npow < n_act) npow *= 2;
    int cpow = 1;
    while (cpow < C) cpow *= 2;
    const bool act = warp < n_act;
    const bool bw = warp == (n_act < kWarps ? n_act : 0);  // the warp that computes the CTA's children range
    int* sh_bounds = reinterpret_cast<int*>(sh_red + 96);  // n_begin, n_end of the step (warp 0 -> everybody)

    fsync();
    if (tid < kNumConsts) sh_c[tid] = p.consts[(size_t)b * kNumConsts + tid];
    if (tid < kWarps) {  // partial slots of the warps without parents: identities, never rewritten
      sh_red[tid] = -INFINITY;
      sh_red[32 + tid] = 0.0;
      sh_red[64 + tid] = 0.0;
    }
    fsync();

    const uint32_t a_tile = sm0 + lay.tile;
    constexpr uint32_t RS = (uint32_t)sizeof(Real);
    const uint32_t a_xs = sm0 + lay.xs, a_lws = sm0 + lay.lws, a_anc = sm0 + lay.anc;
    int top = 1;
    while (top * 2 <= own) top *= 2;

    // ---- generation 0 (standard.py:62-67): the owner draws its own range, uniform weights
    for (int pp = (j_begin >> 1) + tid; pp < ((j_end + 1) >> 1); pp += BLOCK) {
      Real z[2] = {0, 0};
      if (p.gen_init && !inject) philox_normal_pair<Real>(p.seed, eval_id, 0u, (uint32_t)pp, z[0], z[1]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * pp + e;
        if (n >= j_begin && n < j_end) {
          if (p.gen_init && inject) z[e] = (Real)ld_cg(inj_z0 + n);
          const Real x0 = p.gen_init ? Ops::init(sh_c, z[e]) : (Real)p.initial_state;
          const uint32_t off = (uint32_t)(n - j_begin);
          *reinterpret_cast<Real*>(reinterpret_cast<char*>(this->sh_s) + lay.xs + RS * off) = x0;
          *reinterpret_cast<Real*>(reinterpret_cast<char*>(this->sh_s) + lay.lws + RS * off) = (Real)0;
          xr[n] = x0;
          lwr[n] = (Real)0;
          if (p.ws_per_filter) ar[n] = 0;  // generation-0 ancestors are never followed
        }
      }
    }
    // Normals one time step ahead: a draw depends only on (seed, eval_id, t, pair index), and the children
    // of this CTA are, up to the drift of the weight mass, the pairs around its own parent range.  The upper
    // half of the filter warps (idle when a CTA holds <= 256 parents) draws that window while the lower half
    // runs the time step; a child pair outside the window is drawn in place.
    const bool spec = !inject && cap <= kClSpecPairs;
    const int zbase = max(0, ((j_begin + (own >> 1)) >> 1) - kClSpecPairs / 2);
    const uint32_t a_zb = sm0 + lay.zb;
    auto predraw = [&](int tt) {
      if (spec && warp >= kWarps / 2 && tt <= T) {
        const int k = tid - BLOCK / 2;
        Real z0, z1;
        philox_normal_pair_keys<Real>(p.pkeys, eval_id, (uint32_t)tt, (uint32_t)(zbase + k), z0, z1);
        Real* dst = reinterpret_cast<Real*>(reinterpret_cast<char*>(this->sh_s) + lay.zb) + 2 * (((tt & 1) * kClSpecPairs) + k);
        Base::store2(dst, z0, z1);
      }
    };
    predraw(1);
    fsync();
    int sx = 0, sw = 0, sa = 0, sas = 0;  // ring slots of the generation being written: global x / lw / anc, shared anc
    const unsigned seq0 = (unsigned)b_idx * (unsigned)(T + 2) + 1u;
    const int n_reg_jobs = (p.do_smoother && T > L) ? T - L : 0;

    for (int t = 1; t <= T + 1; ++t) {
      const int tau = t - 1;
      const uint32_t par = (uint32_t)tau & 1u, npar = par ^ 1u;
      sx = (sx + 1 == p.RX) ? 0 : sx + 1;
      sw = (sw + 1 == p.RW) ? 0 : sw + 1;
      sa = (sa + 1 == p.RA) ? 0 : sa + 1;
      sas = (sas + 1 == RAs) ? 0 : sas + 1;
      const bool final_step = (t == T + 1);

      double u = 0.0, y_prop = 0.0, y_obs = 0.0;
      const double* inj_z_t = nullptr;
      if (!final_step) {
        if (inject) {
          inj_z_t = inj_z + (size_t)t * N;
          u = ld_cg(inj_u + t);
        } else {
          u = philox_uniform(p.seed, eval_id, kStreamU, (uint32_t)t, 0u);
        }
        y_obs = __ldg(obs + t);
        y_prop = p.svl_obs_model ? __ldg(obs + t - 1) : y_obs;  // Q10
      } else {
        u = inject ? ld_cg(p.inj_u_final + ib) : philox_uniform(p.seed, eval_id, kStreamUFinal, 0u, 0u);
      }
#ifdef PFB_PROFILE
      prof_t = clock64();
#endif
      // ---------------- children of generation tau delivered?  (generation 0 is local)
      int ab = 0;
      const unsigned use = par ? cuse1 : cuse0;
      if (t >= 2) {
        if (act) ab = warp_mbar_wait(sm0 + lay.bars + 16u + 8u * par, use & 1u);
        if (par) ++cuse1; else ++cuse0;
      }
      // arm the barrier of the generation this step produces (its previous phase completed one step ago)
      if (!final_step && tid == 0) cl_mbar_arrive_expect_tx(sm0 + lay.bars + 16u + 8u * npar, (uint32_t)own * kChildBytes);
      PFB_TICK(0);

      // ---------------- phase A: local maximum, shifted weights, local scan (two parents per thread)
      const int j0 = 2 * tid;
      Real l[2] = {(Real)(-INFINITY), (Real)(-INFINITY)}, xv[2] = {(Real)0, (Real)0};
      if (act) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (j0 + k < own) {
            l[k] = cl_lds(a_lws + RS * (par * (uint32_t)cap + (uint32_t)(j0 + k)), Real());
            xv[k] = cl_lds(a_xs + RS * (par * (uint32_t)cap + (uint32_t)(j0 + k)), Real());
          }
        }
        const double mw = warp_max_exact(dmax((double)l[0], (double)l[1]));
        if (lane == 0) sh_red[warp] = mw;
      }
      aborted = fsync_or(ab != 0);  // (1)
      if (aborted) return;
      publish_ring_rows(seq0 + (unsigned)tau);  // rows of generation tau were stored before this barrier
      double mg = -INFINITY, A = 0.0, F = 0.0, pre0 = 0.0, pre1 = 0.0, woff = 0.0, winc = 0.0;
      double s[2] = {0.0, 0.0};
      if (act) {
        mg = (lane < kWarps) ? sh_red[lane] : -INFINITY;
        for (int o = npow >> 1; o > 0; o >>= 1) mg = dmax(mg, __shfl_xor_sync(0xffffffffu, mg, o));
        mg = __shfl_sync(0xffffffffu, mg, 0);
        const double shift = (mg == -INFINITY) ? 0.0 : mg;  // a CTA whose parents all have zero weight
        // both exponentials side by side (out-of-range parents carry -inf: exp gives an exact 0)
        s[0] = shifted_weight_nb(l[0], shift);
        s[1] = shifted_weight_nb(l[1], shift);
        double f = s[0] * (double)xv[0] + s[1] * (double)xv[1];
        const double run = s[0] + s[1];
        const double incl = warp_inclusive_scan(run, lane);
        f = warp_sum(f);
        if (lane == 31) sh_red[32 + warp] = incl;
        if (lane == 0) sh_red[64 + warp] = f;
        pre0 = (incl - run) + s[0];  // inclusive prefixes inside the warp
        pre1 = pre0 + s[1];
      }
      fsync();  // (2)
      if (act) {
        // inclusive scan of the active warps' totals in lanes 0 .. npow-1 (fixed tree)
        const double wt = (lane < kWarps) ? sh_red[32 + lane] : 0.0;
        double wi = wt, fs = (lane < kWarps) ? sh_red[64 + lane] : 0.0;
        for (int o = 1; o < npow; o <<= 1) {
          const double nb = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= o) wi += nb;
        }
        for (int o = npow >> 1; o > 0; o >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, o);
        // exclusive offset of this warp := inclusive prefix of the previous one (the SAME bits both warps use
        // for their common boundary)
        woff = __shfl_sync(0xffffffffu, wi, max(warp - 1, 0));
        if (warp == 0) woff = 0.0;
        winc = __shfl_sync(0xffffffffu, wi, warp);
        A = __shfl_sync(0xffffffffu, wi, npow - 1);
        F = __shfl_sync(0xffffffffu, fs, 0);
        pre0 += woff;
        pre1 += woff;
      }
      PFB_TICK(1);
      xarrive(mg, A);
      // behind the arrival: the shifted weights the smoother warps of THIS CTA re-read (lag time tau)
      if (act && p.do_smoother) {
        Real* s_row = swr + (size_t)(tau % p.RS) * NS + j_begin;
        if (j0 + 1 < own) Base::store2(s_row + j0, (Real)s[0], (Real)s[1]);
        else if (j0 < own) s_row[j0] = (Real)s[0];
      }
      // ---------------- all-gather complete: prefix over the cluster (every active warp on its own)
      ab = 0;
      if (act || bw) {
        ab = warp_mbar_wait(sm0 + lay.bars + 8u * (xepoch & 1u), ((xepoch - 1u) >> 1) & 1u);
        PFB_TICK(2);
        double mr = -INFINITY, Ar = 0.0;
        if (lane < C) {
          const uint32_t a = sm0 + lay.xbuf + 16u * ((xepoch & 1u) * (uint32_t)C + (uint32_t)lane);
          mr = lds_f64(a);
          Ar = lds_f64(a + 8u);
        }
        double m = mr;
        for (int o = cpow >> 1; o > 0; o >>= 1) m = dmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        m = __shfl_sync(0xffffffffu, m, 0);
        const double sc = (mr == -INFINITY) ? 0.0 : exp_nonpos(mr - m);  // exp(0) == 1 exactly
        const double v = __dmul_rn(Ar, sc);
        double vin = v;
        for (int o = 1; o < cpow; o <<= 1) {
          const double nb = __shfl_up_sync(0xffffffffu, vin, o);
          if (lane >= o) vin += nb;
        }
        // exclusive prefix of CTA g := inclusive prefix of CTA g-1: neighbours agree bit for bit on their boundary
        const double Pg1 = __shfl_sync(0xffffffffu, vin, g);
        double Pg = __shfl_sync(0xffffffffu, vin, max(g - 1, 0));
        if (g == 0) Pg = 0.0;
        const double S = __shfl_sync(0xffffffffu, vin, C - 1);
        const double scale = __shfl_sync(0xffffffffu, sc, g);
        const double invS = 1.0 / S;
        // normalised CDF of the CTA's parents: (group prefix + rescaled local prefix) * (1/S)
        const double c0 = __dmul_rn(__dadd_rn(Pg, __dmul_rn(pre0, scale)), invS);
        const double c1 = __dmul_rn(__dadd_rn(Pg, __dmul_rn(pre1, scale)), invS);
        if (j0 < own) this->sh_s[j0] = c0;
        if (j0 + 1 < own) this->sh_s[j0 + 1] = c1;
        if (bw && !final_step && lane < 2) {
          // children range of this CTA: lanes 0 / 1 run the SAME code on the two boundaries.  The warp that
          // does this holds no parents (when there is one): it runs beside the normalisation above.
          const bool fixed = lane == 0 ? (g == 0) : last_cta;
          const int nb = first_child_not_below(__dmul_rn(lane == 0 ? Pg : Pg1, invS), u);
          sh_bounds[lane] = fixed ? (lane == 0 ? 0 : N) : nb;
        }
        if (warp == 0) {
          if (lane == 0) {
            p.filt_part[((size_t)b * (T + 1) + tau) * p.G_loc + g] = F * scale;
            *reinterpret_cast<double*>(reinterpret_cast<char*>(this->sh_s) + lay.wscale + 8u * ((uint32_t)tau & 7u)) = scale;
            if (g == 0) {
              p.stat_m[(size_t)b * (T + 1) + tau] = m;
              p.stat_S[(size_t)b * (T + 1) + tau] = S;
            }
          }
          PFB_TICK(3);
        }
      }
      // jobs whose ring slots phase B of this step overwrites: lag times <= t - kClLag (checked by the last
      // filter warp, off the critical path of the warps that hold parents whenever it holds none)
      if (warp == kWarps - 1 && p.do_smoother && !final_step)
        ab |= wait_smoother_progress(min(max(t - kClLag - L + 1, 0), n_reg_jobs), 0u);
      aborted = fsync_or(ab != 0);  // (3)
      if (aborted) return;
      // generation tau is complete (every CTA has passed its all-gather arrival) and this CTA's shifted
      // weights / scale are stored; the smoother warps also wait for every CTA's ring rows (wait_ring_rows)
      if (tid == 0) st_release_cta_shared_u32(sh_flags, seq0 + (unsigned)tau);
      PFB_TICK(4);

      if (final_step) {
        // trajectory draw (standard.py:104): k = #{ j : cdf[j] <= u_final }, summed over the cluster
        int cnt = 0;
        if (j0 < own && this->sh_s[j0] <= u) ++cnt;
        if (j0 + 1 < own && this->sh_s[j0 + 1] <= u) ++cnt;
        cnt = this->block_sum_int(cnt);
        cnt = __shfl_sync(0xffffffffu, cnt, 0);  // valid in warp 0, which sends it
        xarrive((double)cnt, 0.0);
        ab = 0;
        if (warp == 0) {
          ab = warp_mbar_wait(sm0 + lay.bars + 8u * (xepoch & 1u), ((xepoch - 1u) >> 1) & 1u);
          // + every smoother job of this evaluation, on every CTA (the next evaluation reuses the rings)
          if (p.do_smoother) ab |= wait_smoother_progress(n_reg_jobs, 1u);
          ab |= wait_ring_rows(seq0 + (unsigned)T);  // the ancestor row of generation T, whoever wrote it
          double c = (lane < C) ? lds_f64(sm0 + lay.xbuf + 16u * ((xepoch & 1u) * (uint32_t)C + (uint32_t)lane)) : 0.0;
          c = warp_sum(c);
          if (g == 0 && lane == 0 && !ab) {
            int k = (int)c;
            if (k > N - 1) k = N - 1;
            p.traj_idx[b] = ld_cg(ar + (size_t)(sa == 0 ? p.RA - 1 : sa - 1) * NS + k);
          }
        }
        aborted = fsync_or(ab != 0);
        if (p.do_smoother) done_base += jobs_of(0, n_reg_jobs) + 1u;
        break;
      }

      predraw(t + 1);
      // ---------------- phase B: a thread takes one pair of children (one Philox block)
      const int n_begin = sh_bounds[0], n_end = sh_bounds[1];
      const int pp0 = n_begin >> 1, pp1 = (n_end + 1) >> 1;
      if (pp0 + warp * 32 < pp1) {
        Real* x_new = xr + (size_t)sx * NS;
        Real* lw_new = lwr + (size_t)sw * NS;
        int* anc_new = ar + (size_t)sa * NS;
        double cst[kNumConsts];
        load_step_consts<MODEL>(smem_base(sh_c), cst);
        for (int pp = pp0 + tid; pp < pp1; pp += BLOCK) {
          const int n0 = 2 * pp;
          const bool v0 = n0 >= n_begin, v1 = n0 + 1 < n_end;
          int ja, jb;
          const double pa = pos_sys(n0, u);
          const double pb = (n0 + 1 < N) ? pos_sys(n0 + 1, u) : INFINITY;
          search2(a_tile - 8u, pa, pb, own, top, ja, jb);
          ja = min(ja, own - 1);
          jb = min(jb, own - 1);
          PFB_TICK(5);
          const Real xpa = cl_lds(a_xs + RS * (par * (uint32_t)cap + (uint32_t)ja), Real());
          const Real xpb = cl_lds(a_xs + RS * (par * (uint32_t)cap + (uint32_t)jb), Real());
          Real z0, z1;
          if (inject) {
            z0 = v0 ? (Real)ld_cg(inj_z_t + n0) : (Real)0;
            z1 = v1 ? (Real)ld_cg(inj_z_t + n0 + 1) : (Real)0;
          } else if (spec && (unsigned)(pp - zbase) < (unsigned)kClSpecPairs) {
            const uint32_t za = a_zb + 2u * RS * (uint32_t)((t & 1) * kClSpecPairs + (pp - zbase));
            z0 = cl_lds(za, Real());
            z1 = cl_lds(za + RS, Real());
          } else {
            philox_normal_pair_keys<Real>(p.pkeys, eval_id, (uint32_t)t, (uint32_t)pp, z0, z1);
          }
          const Real xa = Ops::propagate(cst, xpa, z0, y_prop);
#ifdef PFB_PROFILE
          if (xa == (Real)123456.789) u = 0.0;
          PFB_TICK(6);
#endif
          const Real xb = Ops::propagate(cst, xpb, z1, y_prop);
          const Real la = Ops::logweight(cst, xa, y_obs);
          const Real lb = Ops::logweight(cst, xb, y_obs);
          const int aa = j_begin + ja, abq = j_begin + jb;
          // owner of the pair's index range (cap is even: a pair never straddles two owners)
          const uint32_t owner = (uint32_t)n0 >> lg;
          const uint32_t off = (uint32_t)n0 & (uint32_t)(cap - 1);
          const uint32_t rbar = cl_mapa(sm0 + lay.bars + 16u + 8u * npar, owner);
          const uint32_t rx = cl_mapa(a_xs + RS * (npar * (uint32_t)cap + off), owner);
          const uint32_t rl = cl_mapa(a_lws + RS * (npar * (uint32_t)cap + off), owner);
          const uint32_t ra = cl_mapa(a_anc + 4u * ((uint32_t)sas * (uint32_t)cap + off), owner);
          if (v0 && v1) {
            cl_st_async(rx, rbar, xa, xb);
            cl_st_async(rl, rbar, la, lb);
            cl_st_async(ra, rbar, aa, abq);
            Base::store2(x_new + n0, xa, xb);
            Base::store2(lw_new + n0, la, lb);
            *reinterpret_cast<int2*>(anc_new + n0) = make_int2(aa, abq);
          } else if (v0) {
            cl_st_async(rx, rbar, xa);
            cl_st_async(rl, rbar, la);
            cl_st_async(ra, rbar, aa);
            x_new[n0] = xa; lw_new[n0] = la; anc_new[n0] = aa;
          } else if (v1) {
            cl_st_async(rx + RS, rbar, xb);
            cl_st_async(rl + RS, rbar, lb);
            cl_st_async(ra + 4u, rbar, abq);
            x_new[n0 + 1] = xb; lw_new[n0 + 1] = lb; anc_new[n0 + 1] = abq;
          }
          PFB_TICK(7);
        }
      }
      PFB_TICK(8);
    }
  }

  // ---- fixed-lag smoother (standard.py:146-169), ONE job per smoother warp ---------------------------
  // lineages of the CTA's own range at lag time tau (shared ancestor-ring slot of generation tau: slot_a)
  // walked back nlev generations to smoothing time i; partial sums scaled to the global shift
  __device__ void smooth_job(int b, const Real* xr, const Real* swr, const double* obs, int tau, int i,
                             int j_begin, int own) {
    const int T = p.T;
    const size_t NS = (size_t)p.NS;
    const int nlev = tau - i;
    const uint32_t a_anc = sm0 + lay.anc;
    const uint32_t capm = (uint32_t)(cap - 1);
    const uint64_t once = p.pol_once;
    const Real* sw_row = swr + (size_t)(tau % p.RS) * NS + j_begin;
    const int si = i % p.RX, sn = (i + 1) % p.RX;
    const Real* xi = opaque(xr + (size_t)si * NS);
    const Real* xn = opaque(xr + (size_t)sn * NS);
    const double y = __ldg(obs + i);
    double acc[kMaxParams + 1] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int c0 = 0; c0 < own; c0 += 32 * kWalk) {
      int cur[kWalk], nxt[kWalk];
#pragma unroll
      for (int k = 0; k < kWalk; ++k) {
        const int jl = c0 + k * 32 + lane;
        cur[k] = j_begin + ((jl < own) ? jl : c0);
        nxt[k] = cur[k];
      }
      const int kmax = min(kWalk, (own - c0 + 31) / 32);  // warp-uniform number of live items
      int slot = tau % RAs;
      for (int lvl = 1; lvl <= nlev; ++lvl) {
        if (lvl == nlev) {
#pragma unroll
          for (int k = 0; k < kWalk; ++k) nxt[k] = cur[k];
        }
#pragma unroll
        for (int k = 0; k < kWalk; ++k) {
          if (k < kmax) {
            const uint32_t cu = (uint32_t)cur[k];
            cur[k] = cl_ld_remote_s32(cl_mapa(a_anc + 4u * ((uint32_t)slot * (uint32_t)cap + (cu & capm)), cu >> lg));
          }
        }
        slot = (slot == 0) ? RAs - 1 : slot - 1;
      }
      Real wv[kWalk], xc[kWalk], xnext[kWalk];
#pragma unroll
      for (int k = 0; k < kWalk; ++k) {
        const int jl = c0 + k * 32 + lane;
        wv[k] = (k < kmax && jl < own) ? ld_hint(sw_row + jl, once) : (Real)0;
        if (k < kmax) {
          xc[k] = ld_hint(xi + (unsigned)cur[k], once);
          xnext[k] = ld_hint(xn + (unsigned)nxt[k], once);
        } else {
          xc[k] = (Real)0;
          xnext[k] = (Real)0;
        }
      }
#pragma unroll
      for (int k = 0; k < kWalk; ++k) {
        if (k < kmax) {
          const Real xck = xc[k], xnk = xnext[k];
          const double sw = (double)wv[k];
          double gr[kMaxParams];
          Ops::gradient(sh_c, xnk, xck, y, gr);
          const double t0 = (double)xck * sw;
          if (isfinite((double)xck) && isfinite((double)xnk)) {  // np.nansum: NaN terms are skipped
            acc[0] += t0;
#pragma unroll
            for (int pp = 0; pp < kMaxParams; ++pp) acc[1 + pp] += gr[pp] * sw;
          } else {
            if (!isnan(t0)) acc[0] += t0;
#pragma unroll
            for (int pp = 0; pp < kMaxParams; ++pp) {
              const double tp = gr[pp] * sw;
              if (!isnan(tp)) acc[1 + pp] += tp;
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k <= kMaxParams; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
      const double scale = *reinterpret_cast<const volatile double*>(reinterpret_cast<char*>(this->sh_s) + lay.wscale + 8u * ((uint32_t)tau & 7u));
      double* dst = p.smo_part + (((size_t)b * (T + 1) + i) * p.G_loc + g) * (p.P + 1);
#pragma unroll
      for (int k = 0; k <= kMaxParams; ++k)
        if (k <= p.P) dst[k] = acc[k] * scale;
    }
  }

  // lane 0 waits until the filter warps have published generation `seq`; returns true when abandoned
  __device__ bool warp_wait_generation(unsigned seq) const {
    int ab = 0;
    if (lane == 0) ab = this->template wait_counter<0>(sh_flags, seq);
    ab = __shfl_sync(0xffffffffu, ab, 0);
    if (!ab) ab = wait_ring_rows(seq);
    return ab != 0;
  }
  // this warp has finished `count` jobs (cumulative): tell every CTA of the cluster
  __device__ void warp_report(int sw_id, unsigned count) const {
    __syncwarp();
    if (lane < C) cl_st_release_remote_u32(cl_mapa(sm0 + lay.prog + 4u * (uint32_t)(g * kSW + sw_id), (uint32_t)lane), count);
  }

  __device__ void run_smoother(int b, int b_idx) {
    const int N = p.N, T = p.T, L = p.L;
    const int sw_id = (tid - BLOCK) >> 5;
    const int ws = p.ws_per_filter ? b : q;
    const size_t NS = (size_t)p.NS;
    const Real* xr = reinterpret_cast<const Real*>(p.x_ring) + (size_t)(p.xs_per_filter ? b : ws) * p.RX * NS;
    const Real* swr = reinterpret_cast<const Real*>(p.s_ring) + (size_t)ws * p.RS * NS;
    const double* obs = p.obs + (p.obs_per_filter ? (size_t)b * (T + 1) : 0);
    const int j_begin = min(g * cap, N);
    const int own = min(j_begin + cap, N) - j_begin;
    const unsigned seq0 = (unsigned)b_idx * (unsigned)(T + 2) + 1u;
    const int n_reg = (T > L) ? T - L : 0;
    unsigned done = done_base;
    // the model constants of evaluation b are in sh_c once generation 0 .. L is published (the filter
    // warps load them before generation 0); every job below waits for a generation first
    for (int k = sw_id; k < n_reg; k += kSW) {
      const int tau = L + k;
      if (warp_wait_generation(seq0 + (unsigned)tau)) { aborted = true; return; }
      smooth_job(b, xr, swr, obs, tau, tau - L, j_begin, own);
      warp_report(sw_id, ++done);
    }
    if (warp_wait_generation(seq0 + (unsigned)T)) { aborted = true; return; }
    for (int i = T - 1 - sw_id; i >= ((T - L > 0) ? T - L : 0); i -= kSW) smooth_job(b, xr, swr, obs, T, i, j_begin, own);
    warp_report(sw_id, ++done);
    done_base += jobs_of(0, n_reg) + 1u;
    // warps with fewer regular jobs than warp 0 pad their count: `done_base + jobs_of(w, n)` is what the
    // filter warps wait for, so the cumulative base must advance identically in every warp
    if (done != done_base) warp_report(sw_id, done_base);
  }
};

// Cluster launch: grid = n_groups * C CTAs, cluster dimension C (launch attribute); BLOCK filter threads +
// one warpgroup of smoother threads per CTA, register file re-split as in pf_persistent_kernel.
template <int MODEL, typename Real, bool GENERAL>
__global__ void __launch_bounds__(PFB_BLOCK + PFB_SMOOTH_THREADS, 1) pf_cluster_kernel(const DeviceProblem prob) {
  extern __shared__ double pf_smem[];
  __shared__ double sh_red[9 * 32 + 8];
  __shared__ double sh_c[kNumConsts];
  __shared__ double sh_sred[(PFB_SMOOTH_THREADS / 32) * (kMaxParams + 1)];
  __shared__ unsigned sh_flags[4];
  ClKernel<MODEL, Real, GENERAL> k(prob, pf_smem, sh_red, sh_c, sh_sred, sh_flags);
  if (threadIdx.x == 0) {
    sh_flags[0] = 0u;
    sh_flags[1] = 0u;
    for (int i = 0; i < 4; ++i) cl_mbar_init(k.sm0 + k.lay.bars + 8u * i, 1u);
    cl_fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 5 * prob.cluster; i += blockDim.x)  // prog[C][4] and vis[C] are contiguous
    reinterpret_cast<unsigned*>(reinterpret_cast<char*>(pf_smem) + k.lay.prog)[i] = 0u;
  __syncthreads();
  cl_sync_all();  // barriers and progress slots of every CTA exist before any remote access
  if (threadIdx.x >= PFB_BLOCK) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    if (prob.do_smoother) {
      int b_idx = 0;
      for (int b = k.q; b < prob.B && !k.aborted; b += prob.n_groups, ++b_idx) k.run_smoother(b, b_idx);
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    int b_idx = 0;
    for (int b = k.q; b < prob.B && !k.aborted; b += prob.n_groups, ++b_idx) k.run_filter(b, b_idx);
#ifdef PFB_PROFILE
    if (threadIdx.x == 0)
      for (int i = 0; i < kProfSlots; ++i) prob.prof[(size_t)blockIdx.x * kProfSlots + i] = k.prof_acc[i];
#endif
  }
  cl_sync_all();  // no CTA leaves while a peer may still address its shared memory
}

}  // namespace pfb200
