// The persistent particle-filter / fixed-lag-smoother kernel.
//
// One launch runs the whole time loop t = 1..T of one or many filters.  A filter is worked on by a
// GROUP of G cooperating CTAs (G = 1 for the batched small-N case, G = all resident CTAs for one
// large filter); a group loops over the filters assigned to it.  Per time step a CTA
//
//   phase A (generation tau = t-1, parents j in the CTA's contiguous range)
//     s[j]   = exp(logw[j] - m_tau)                       weight normalisation, standard.py:90-93
//     tile-local inclusive scan of s  -> CDF tile in shared memory         resampling.pyx:75
//     partial sums  sum s*x  (filtered mean, standard.py:101)
//     fixed-lag smoother increment for time i = tau - L: chase L ancestor links through the
//     ancestor ring, accumulate s * (x_i, d log p / d theta)                standard.py:146-169
//   exchange 1: every CTA publishes its tile total; all CTAs form the same exclusive prefix
//   phase B (generation t, children n whose position (n+u)/N falls into the CTA's CDF range --
//            contiguous because systematic resampling is monotone)
//     a[n]   = first j with (n+u)/N < c[j]   (search in the shared-memory CDF tile) resampling.pyx:78-83
//     x_t[n] = f(x_{t-1}[a[n]]) + noise(t, n),  logw_t[n] = log g(y_t | x_t[n])   standard.py:85-88
//   exchange 2: per-CTA max of the new log-weights -> m_t
//
// so neither the CDF nor the normalised weights ever touch global memory.  Generations live in
// time-major rings: x (L+2 deep), ancestors (L+1), log-weights (2).
#pragma once
#include "pf_device.cuh"

namespace pfb200 {

#ifdef PFB_PROFILE
#define PFB_TICK(slot) do { const long long now__ = clock64(); prof_acc[slot] += now__ - prof_t; prof_t = now__; } while (0)
#else
#define PFB_TICK(slot) do { } while (0)
#endif

struct DeviceProblem {
  int N, T, L, P;
  int NS;        // row stride of the rings (N rounded up to 32 elements: aligned vector stores)
  int B;         // number of filters in the batch
  int G;         // CTAs per filter group
  int n_groups;  // groups resident in this launch
  int per_cta;   // parents owned by one CTA
  int RX, RA, RW;  // ring depths: particles, ancestors, log-weights
  int do_smoother, gen_init, resampling, svl_obs_model;
  int inject, inject_per_filter, obs_per_filter, ws_per_filter;
  double initial_state;
  unsigned long long seed;
  long long spin_limit;
  const double* obs;
  const double* consts;              // [B][kNumConsts]
  const unsigned long long* eval_ids;  // [B]
  const double* inj_z0;
  const double* inj_u;
  const double* inj_z;
  const double* inj_u_final;
  void* x_ring;    // [n_ws][RX][N] Real
  void* logw;      // [n_ws][RW][N] Real
  int* anc_ring;   // [n_ws][RA][N]
  double* xval;    // [n_groups][2][G]
  unsigned* xflag; // [n_groups][2][G]
  int* abort_flag;
  double* stat_m;     // [B][T+1]
  double* stat_S;     // [B][T+1]
  double* filt_part;  // [B][T+1][G]
  double* smo_part;   // [B][T+1][G][P+1]
  int* traj_k;        // [B]
  int* traj_idx;      // [B]
  long long* prof;    // [grid][8] per-phase cycle counters (PFB_PROFILE builds only)
};

template <int BLOCK, int ITEMS>
struct Cfg {
  static constexpr int kBlock = BLOCK;
  static constexpr int kItems = ITEMS;           // odd: the blocked smem pass is bank-conflict free
  static constexpr int kChunk = BLOCK * ITEMS;   // parents per shared-memory CDF tile
  static constexpr int kWarps = BLOCK / 32;
};

// ---------------------------------------------------------------------------------------------
template <int MODEL, typename Real, int BLOCK, int ITEMS>
struct PfKernel {
  using C = Cfg<BLOCK, ITEMS>;
  using Ops = ModelOps<MODEL, Real>;
  static constexpr int kRedK = kMaxParams + 2;

  const DeviceProblem& p;
  double* sh_s;    // [kChunk] CDF tile
  double* sh_x;    // [G] exchange staging
  double* sh_red;  // [kWarps][kRedK]
  double* sh_c;    // [kNumConsts]
  const int tid, lane, warp, g, q;
  const bool pow2N;
  const double invN;
  unsigned epoch;
  bool aborted;
#ifdef PFB_PROFILE
  long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_t = 0;
#endif

  __device__ PfKernel(const DeviceProblem& prob, double* smem, double* red, double* cst)
      : p(prob), sh_s(smem), sh_x(smem + C::kChunk), sh_red(red), sh_c(cst), tid(threadIdx.x),
        lane(threadIdx.x & 31), warp(threadIdx.x >> 5), g(blockIdx.x % prob.G),
        q(blockIdx.x / prob.G), pow2N((prob.N & (prob.N - 1)) == 0), invN(1.0 / (double)prob.N),
        epoch(0), aborted(false) {}

  // ---- inter-CTA exchange: every CTA of the group contributes one double; on return
  //      sh_x[0..G) holds all contributions (same bits in every CTA).
  __device__ void exchange(double v) {
    const int G = p.G;
    if (G == 1) {
      __syncthreads();
      if (tid == 0) sh_x[0] = v;
      __syncthreads();
      return;
    }
    ++epoch;
    double* val = p.xval + ((size_t)q * 2 + (epoch & 1u)) * G;
    unsigned* flg = p.xflag + ((size_t)q * 2 + (epoch & 1u)) * G;
    __syncthreads();  // all global writes of this CTA precede the release below
    if (tid == 0) {
      val[g] = v;
      __threadfence();
      st_release_u32(flg + g, epoch);
    }
    for (int i = tid; i < G; i += BLOCK) {
      long long spins = 0;
      while ((int)(ld_acquire_u32(flg + i) - epoch) < 0) {
        if (++spins > p.spin_limit || ld_relaxed_i32(p.abort_flag) != 0) {
          atomicExch(p.abort_flag, 1);
          aborted = true;
          break;
        }
      }
      sh_x[i] = ld_cg(val + i);
    }
    aborted = __syncthreads_or(aborted ? 1 : 0) != 0;  // block-uniform
  }

  __device__ double xchg_max() {  // order independent
    double m = -INFINITY;
    for (int i = tid; i < p.G; i += BLOCK) m = fmax(m, sh_x[i]);
    return block_max(m);
  }

  // Prefix sums of sh_x in a fixed order.  incl(i) = offset of the lane segment holding i (a
  // shuffle scan of the segment totals) + the running sum inside the segment; every CTA evaluates
  // the same function on the same array, so neighbouring CTAs agree bit for bit on their common
  // boundary: excl(g) := incl(g-1).
  __device__ void xchg_prefix(double& excl, double& incl, double& total) {
    const int G = p.G;
    if (warp == 0) {
      const int W = (G + 31) / 32;
      const int b0 = lane * W;
      double run = 0.0;
      for (int k = 0; k < W; ++k) {
        const int i = b0 + k;
        if (i < G) run += sh_x[i];
      }
      const double lane_incl = warp_inclusive_scan(run, lane);
      double r = __shfl_up_sync(0xffffffffu, lane_incl, 1);
      if (lane == 0) r = 0.0;
      double vp = 0.0, vc = 0.0, vt = 0.0;
      bool hp = false, hc = false, ht = false;
      for (int k = 0; k < W; ++k) {
        const int i = b0 + k;
        if (i < G) {
          r += sh_x[i];
          if (i == g - 1) { vp = r; hp = true; }
          if (i == g) { vc = r; hc = true; }
          if (i == G - 1) { vt = r; ht = true; }
        }
      }
      const unsigned wp = __ballot_sync(0xffffffffu, hp);
      const unsigned wc = __ballot_sync(0xffffffffu, hc);
      const unsigned wt = __ballot_sync(0xffffffffu, ht);
      vp = wp ? __shfl_sync(0xffffffffu, vp, __ffs(wp) - 1) : 0.0;
      vc = __shfl_sync(0xffffffffu, vc, __ffs(wc) - 1);
      vt = __shfl_sync(0xffffffffu, vt, __ffs(wt) - 1);
      if (lane == 0) { sh_red[0] = vp; sh_red[1] = vc; sh_red[2] = vt; }
    }
    __syncthreads();
    excl = sh_red[0]; incl = sh_red[1]; total = sh_red[2];
    __syncthreads();
  }

  __device__ double block_max(double v) {
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) sh_red[warp] = v;
    __syncthreads();
    double r = -INFINITY;
    if (warp == 0) {
      for (int i = lane; i < C::kWarps; i += 32) r = fmax(r, sh_red[i]);
      r = warp_max(r);
      if (lane == 0) sh_red[0] = r;
    }
    __syncthreads();
    r = sh_red[0];
    __syncthreads();
    return r;
  }

  // sums K accumulators over the block; result valid in thread 0 (fixed tree)
  template <int K>
  __device__ void block_sum(double (&v)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) sh_red[warp * kRedK + k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double r = 0.0;
        for (int i = lane; i < C::kWarps; i += 32) r += sh_red[i * kRedK + k];
        v[k] = warp_sum(r);
      }
    }
    __syncthreads();
  }

  __device__ int block_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int* red = reinterpret_cast<int*>(sh_red);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    int r = 0;
    if (warp == 0) {
      for (int i = lane; i < C::kWarps; i += 32) r += red[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;  // valid in warp 0
  }

  // ---- tile load: sh_s[jl] = exp(logw[j] - m) for the chunk [c0, c0+len); optional sum s*x
  __device__ void load_chunk(const Real* lw, const Real* xg, int c0, int len, double m, double* accF) {
    double f = 0.0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int jl = k * BLOCK + tid;
      double s = 0.0;
      if (jl < len) {
        s = shifted_weight<Real>(ld_cg(lw + c0 + jl), m);
        if (accF) f += s * (double)ld_cg(xg + c0 + jl);
      }
      sh_s[jl] = s;
    }
    if (accF) *accF += f;
    __syncthreads();
  }

  // ---- in-place inclusive scan of sh_s[0..kChunk): thread-blocked runs (odd ITEMS: no bank
  //      conflicts), warp shuffle scan of run totals, warp totals through shared memory.
  __device__ double scan_chunk() {
    double v[ITEMS];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      run += sh_s[tid * ITEMS + k];
      v[k] = run;
    }
    const double incl = warp_inclusive_scan(run, lane);
    if (lane == 31) sh_red[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      double w = (lane < C::kWarps) ? sh_red[lane] : 0.0;
      const double wi = warp_inclusive_scan(w, lane);
      sh_red[lane] = wi - w;  // exclusive offsets of the warps (kWarps <= 32)
      if (lane == 31) sh_red[32] = wi;
    }
    __syncthreads();
    const double off = sh_red[warp] + (incl - run);
    const double total = sh_red[32];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) sh_s[tid * ITEMS + k] = off + v[k];
    __syncthreads();
    return total;
  }

  // ---- fixed-lag smoother increment for smoothing time i with weights of generation tau
  //      (standard.py:146-169); s[j] is read from the (unscanned) tile in sh_s.
  __device__ void smooth_chunk(const Real* xr, const int* ar, const double* obs, int tau, int i,
                               int c0, int len, double (&acc)[kMaxParams + 1]) {
    const size_t NS = (size_t)p.NS;
    int cur[ITEMS], nxt[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int jl = k * BLOCK + tid;
      cur[k] = (jl < len) ? c0 + jl : c0;
      nxt[k] = cur[k];
    }
    int slot = tau % p.RA;
    for (int lvl = tau; lvl > i; --lvl) {
      const int* a = ar + (size_t)slot * NS;
      slot = (slot == 0) ? p.RA - 1 : slot - 1;
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        nxt[k] = cur[k];
        cur[k] = ld_cg(a + cur[k]);
      }
    }
    const Real* xi = xr + (size_t)(i % p.RX) * NS;
    const Real* xn = xr + (size_t)((i + 1) % p.RX) * NS;
    const double y = __ldg(obs + i);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int jl = k * BLOCK + tid;
      const double sw = (jl < len) ? sh_s[jl] : 0.0;
      const Real xc = ld_cg(xi + cur[k]);
      const Real xnext = ld_cg(xn + nxt[k]);
      double gr[kMaxParams];
      Ops::gradient(sh_c, xnext, xc, y, gr);
      const double t0 = (double)xc * sw;
      if (!isnan(t0)) acc[0] += t0;  // np.nansum
#pragma unroll
      for (int pp = 0; pp < kMaxParams; ++pp) {
        const double tp = gr[pp] * sw;
        if (!isnan(tp)) acc[1 + pp] += tp;
      }
    }
  }

  // position of child n on the unit interval (resampling.pyx:72 systematic, :51 stratified)
  __device__ __forceinline__ double position(int n, double u, int t, const double* inj_u_t,
                                             unsigned long long eval_id) const {
    if (p.resampling == 1) {
      u = p.inject ? ld_cg(inj_u_t + n) : philox_uniform(p.seed, eval_id, kStreamU, (uint32_t)t, (uint32_t)n);
      const double a = __dadd_rn(u, (double)n);
      return pow2N ? __dmul_rn(a, invN) : __ddiv_rn(a, (double)p.N);
    }
    const double a = __dadd_rn((double)n, u);
    return pow2N ? __dmul_rn(a, invN) : __ddiv_rn(a, (double)p.N);  // x * 2^-k == x / 2^k exactly
  }

  // smallest n in [0, N] with NOT (position(n) < c)
  __device__ int lower_bound_child(double c, double u, int t, const double* inj_u_t,
                                   unsigned long long eval_id) const {
    const int N = p.N;
    if (!(c == c)) return N;  // NaN CDF (all weights zero): nothing to assign
    double est = ceil(c * (double)N - (p.resampling == 1 ? 0.5 : u));
    int n = est < 0.0 ? 0 : (est > (double)N ? N : (int)est);
    while (n > 0 && !(position(n - 1, u, t, inj_u_t, eval_id) < c)) --n;
    while (n < N && position(n, u, t, inj_u_t, eval_id) < c) ++n;
    return n;
  }

  // #{ jl in [lo, len) : sh_s[jl] <= pos } + lo   (upper bound)
  __device__ __forceinline__ int upper_bound_tile(double pos, int lo, int len) const {
    int hi = len;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sh_s[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    return lo;
  }

  // parent (tile-local) of a child at `pos`, knowing that it is >= lo (children are monotone)
  __device__ __forceinline__ int next_parent(double pos, int lo, int len) const {
    if (sh_s[lo] > pos) return lo;
    if (lo + 1 < len && sh_s[lo + 1] > pos) return lo + 1;
    return upper_bound_tile(pos, (lo + 2 < len) ? lo + 2 : len, len);
  }

  // ---- children of the CDF tile in sh_s: resample + propagate + weight (phase B).  A thread owns
  //      4 consecutive children: one binary search, then short forward merges; 32-byte stores.
  __device__ void children_chunk(const Real* x_prev, Real* x_new, Real* lw_new, int* anc_new,
                                 const double* inj_z_t, const double* inj_u_t, int t, double u,
                                 unsigned long long eval_id, double y_prop, double y_obs, int c0,
                                 int len, int n_lo, int n_hi, double& mx) {
    const int q0 = n_lo >> 2, q1 = (n_hi + 3) >> 2;
    for (int qq = q0 + tid; qq < q1; qq += BLOCK) {
      const int nb = 4 * qq;
      Real z[4];
      if (p.inject) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = nb + e;
          z[e] = (n >= n_lo && n < n_hi) ? (Real)ld_cg(inj_z_t + n) : (Real)0;
        }
      } else {
        philox_normal_pair<Real>(p.seed, eval_id, (uint32_t)t, (uint32_t)(2 * qq), z[0], z[1]);
        philox_normal_pair<Real>(p.seed, eval_id, (uint32_t)t, (uint32_t)(2 * qq + 1), z[2], z[3]);
      }
      int j[4];
      int lo = -1;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = nb + e;
        j[e] = 0;
        if (n >= n_lo && n < n_hi) {
          const double pos = position(n, u, t, inj_u_t, eval_id);
          lo = (lo < 0) ? upper_bound_tile(pos, 0, len) : next_parent(pos, lo, len);
          if (lo > len - 1) lo = len - 1;
          j[e] = lo;
        }
      }
      Real xv[4], lv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = nb + e;
        xv[e] = 0; lv[e] = 0;
        if (n >= n_lo && n < n_hi) {
          xv[e] = Ops::propagate(sh_c, ld_cg(x_prev + c0 + j[e]), z[e], y_prop);
          lv[e] = Ops::logweight(sh_c, xv[e], y_obs);
          mx = fmax(mx, (double)lv[e]);
        }
      }
      if (nb >= n_lo && nb + 3 < n_hi) {
        if (sizeof(Real) == 8) {
          double2* xd = reinterpret_cast<double2*>(x_new + nb);
          double2* ld = reinterpret_cast<double2*>(lw_new + nb);
          xd[0] = make_double2((double)xv[0], (double)xv[1]);
          xd[1] = make_double2((double)xv[2], (double)xv[3]);
          ld[0] = make_double2((double)lv[0], (double)lv[1]);
          ld[1] = make_double2((double)lv[2], (double)lv[3]);
        } else {
          *reinterpret_cast<float4*>(x_new + nb) = make_float4((float)xv[0], (float)xv[1], (float)xv[2], (float)xv[3]);
          *reinterpret_cast<float4*>(lw_new + nb) = make_float4((float)lv[0], (float)lv[1], (float)lv[2], (float)lv[3]);
        }
        *reinterpret_cast<int4*>(anc_new + nb) = make_int4(c0 + j[0], c0 + j[1], c0 + j[2], c0 + j[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = nb + e;
          if (n >= n_lo && n < n_hi) { x_new[n] = xv[e]; lw_new[n] = lv[e]; anc_new[n] = c0 + j[e]; }
        }
      }
    }
  }

  // ---- one filter evaluation (all T steps) -------------------------------------------------
  __device__ void run_filter(int b) {
    const int N = p.N, T = p.T, L = p.L, G = p.G, P = p.P;
    const int ws = p.ws_per_filter ? b : q;
    const size_t NS = (size_t)p.NS;
    Real* xr = reinterpret_cast<Real*>(p.x_ring) + (size_t)ws * p.RX * NS;
    Real* lwr = reinterpret_cast<Real*>(p.logw) + (size_t)ws * p.RW * NS;
    int* ar = p.anc_ring + (size_t)ws * p.RA * NS;
    const double* obs = p.obs + (p.obs_per_filter ? (size_t)b * (T + 1) : 0);
    const unsigned long long eval_id = p.eval_ids[b];
    const size_t ib = p.inject_per_filter ? (size_t)b : 0;
    const double* inj_z0 = p.inject ? p.inj_z0 + ib * N : nullptr;
    const size_t u_row = (p.resampling == 1) ? (size_t)N : 1;
    const double* inj_u = p.inject ? p.inj_u + ib * (T + 1) * u_row : nullptr;
    const double* inj_z = p.inject ? p.inj_z + ib * (size_t)(T + 1) * N : nullptr;

    __syncthreads();
    if (tid < kNumConsts) sh_c[tid] = p.consts[(size_t)b * kNumConsts + tid];
    __syncthreads();

    const int j_begin = min(g * p.per_cta, N);
    const int j_end = min(j_begin + p.per_cta, N);
    const int n_chunks = (j_end - j_begin + C::kChunk - 1) / C::kChunk;
    const bool last_cta = (j_end == N);

    // ---- generation 0 (standard.py:62-67): stationary draw or fixed value, uniform weights
    for (int pp = (j_begin >> 1) + tid; pp < ((j_end + 1) >> 1); pp += BLOCK) {
      Real z[2] = {0, 0};
      if (p.gen_init && !p.inject) philox_normal_pair<Real>(p.seed, eval_id, 0u, (uint32_t)pp, z[0], z[1]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * pp + e;
        if (n >= j_begin && n < j_end) {
          if (p.gen_init && p.inject) z[e] = (Real)ld_cg(inj_z0 + n);
          xr[n] = p.gen_init ? Ops::init(sh_c, z[e]) : (Real)p.initial_state;
          lwr[n] = (Real)0;
          if (p.ws_per_filter) ar[n] = 0;  // generation-0 ancestors are never followed
        }
      }
    }
    exchange(0.0);
    double m = xchg_max();

    for (int t = 1; t <= T + 1; ++t) {
      if (aborted) return;
      const int tau = t - 1;
      const Real* x_prev = xr + (size_t)(tau % p.RX) * NS;
      const Real* lw_prev = lwr + (size_t)(tau % p.RW) * NS;
      const bool final_step = (t == T + 1);
      const int i_sm = (p.do_smoother && !final_step && tau >= L) ? tau - L : -1;

#ifdef PFB_PROFILE
      prof_t = clock64();
#endif
      // ---------------- phase A
      double accF = 0.0, A = 0.0, chunk_total = 0.0;
      double acc[kMaxParams + 1] = {0.0, 0.0, 0.0, 0.0, 0.0};
      for (int c = 0; c < n_chunks; ++c) {
        const int c0 = j_begin + c * C::kChunk;
        const int len = min(C::kChunk, j_end - c0);
        load_chunk(lw_prev, x_prev, c0, len, m, &accF);
        PFB_TICK(8);
        if (i_sm >= 0) {
          smooth_chunk(xr, ar, obs, tau, i_sm, c0, len, acc);
          __syncthreads();
        }
        PFB_TICK(9);
        chunk_total = scan_chunk();
        A += chunk_total;
      }
      PFB_TICK(0);
      {
        double red[kMaxParams + 2] = {accF, acc[0], acc[1], acc[2], acc[3], acc[4]};
        block_sum(red);
        if (tid == 0) {
          p.filt_part[((size_t)b * (T + 1) + tau) * G + g] = red[0];
          if (i_sm >= 0) {
            double* dst = p.smo_part + (((size_t)b * (T + 1) + i_sm) * G + g) * (P + 1);
            for (int k = 0; k <= P; ++k) dst[k] = red[1 + k];
          }
        }
      }
      PFB_TICK(1);
      exchange(A);
      PFB_TICK(2);
      double Pg, Pg1, S;
      xchg_prefix(Pg, Pg1, S);
      if (g == 0 && tid == 0) {
        p.stat_m[(size_t)b * (T + 1) + tau] = m;
        p.stat_S[(size_t)b * (T + 1) + tau] = S;
      }

      // ---------------- smoother tail (standard.py:146: lag = T for the last L times)
      if (final_step && p.do_smoother) {
        const int i_first = (T - L > 0) ? T - L : 0;
        for (int i = T - 1; i >= i_first; --i) {
          double ta[kMaxParams + 1] = {0.0, 0.0, 0.0, 0.0, 0.0};
          for (int c = 0; c < n_chunks; ++c) {
            const int c0 = j_begin + c * C::kChunk;
            const int len = min(C::kChunk, j_end - c0);
            __syncthreads();
            load_chunk(lw_prev, x_prev, c0, len, m, nullptr);
            smooth_chunk(xr, ar, obs, tau, i, c0, len, ta);
          }
          block_sum(ta);
          if (tid == 0) {
            double* dst = p.smo_part + (((size_t)b * (T + 1) + i) * G + g) * (P + 1);
            for (int k = 0; k <= P; ++k) dst[k] = ta[k];
          }
        }
        // restore the scanned tile of the (single) chunk for the trajectory draw below
        if (n_chunks == 1) {
          __syncthreads();
          load_chunk(lw_prev, x_prev, j_begin, j_end - j_begin, m, nullptr);
          chunk_total = scan_chunk();
        }
      }

      // ---------------- phase B
      double u = 0.0, y_prop = 0.0, y_obs = 0.0;
      const double* inj_u_t = nullptr;
      const double* inj_z_t = nullptr;
      Real* x_new = nullptr; Real* lw_new = nullptr; int* anc_new = nullptr;
      if (!final_step) {
        if (p.inject) {
          inj_u_t = inj_u + (size_t)t * u_row;
          inj_z_t = inj_z + (size_t)t * N;
          u = (p.resampling == 1) ? 0.0 : ld_cg(inj_u_t);
        } else if (p.resampling == 0) {
          u = philox_uniform(p.seed, eval_id, kStreamU, (uint32_t)t, 0u);
        }
        y_obs = __ldg(obs + t);
        y_prop = p.svl_obs_model ? __ldg(obs + t - 1) : y_obs;  // Q10
        x_new = xr + (size_t)(t % p.RX) * NS;
        lw_new = lwr + (size_t)(t % p.RW) * NS;
        anc_new = ar + (size_t)(t % p.RA) * NS;
      } else {
        u = p.inject ? ld_cg(p.inj_u_final + ib) : philox_uniform(p.seed, eval_id, kStreamUFinal, 0u, 0u);
      }
      const double invS = 1.0 / S;  // the CDF is (prefix sum) * (1/S) everywhere, boundaries included
      const int n_begin = final_step ? 0 : ((g == 0) ? 0 : lower_bound_child(__dmul_rn(Pg, invS), u, t, inj_u_t, eval_id));
      const int n_end = final_step ? 0 : (last_cta ? N : lower_bound_child(__dmul_rn(Pg1, invS), u, t, inj_u_t, eval_id));
      PFB_TICK(3);
      double mx = -INFINITY;
      int n_cursor = n_begin, traj_cnt = 0;
      double base = Pg;
      for (int c = 0; c < n_chunks; ++c) {
        const int c0 = j_begin + c * C::kChunk;
        const int len = min(C::kChunk, j_end - c0);
        if (n_chunks > 1) {
          __syncthreads();
          load_chunk(lw_prev, x_prev, c0, len, m, nullptr);
          chunk_total = scan_chunk();
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
          const int jl = k * BLOCK + tid;
          if (jl < len) sh_s[jl] = __dmul_rn(__dadd_rn(base, sh_s[jl]), invS);
        }
        __syncthreads();
        PFB_TICK(10);
        if (final_step) {
          // trajectory draw (standard.py:104): k = #{ j : cdf[j] <= u_final }
#pragma unroll
          for (int k = 0; k < ITEMS; ++k) {
            const int jl = k * BLOCK + tid;
            if (jl < len && sh_s[jl] <= u) ++traj_cnt;
          }
        } else {
          const double next_base = base + chunk_total;
          int n_hi = n_end;
          if (c + 1 < n_chunks) {
            n_hi = lower_bound_child(__dmul_rn(next_base, invS), u, t, inj_u_t, eval_id);
            n_hi = max(n_cursor, min(n_hi, n_end));
          }
          children_chunk(x_prev, x_new, lw_new, anc_new, inj_z_t, inj_u_t, t, u, eval_id, y_prop,
                         y_obs, c0, len, n_cursor, n_hi, mx);
          n_cursor = n_hi;
        }
        base += chunk_total;
      }
      if (final_step) {
        const int cnt = block_sum_int(traj_cnt);
        if (tid == 0 && cnt) atomicAdd(p.traj_k + b, cnt);
        exchange(0.0);  // all counts are in; also fences the workspace before the next filter
        if (g == 0 && tid == 0) {
          int k = ld_cg(p.traj_k + b);
          if (k > N - 1) k = N - 1;
          p.traj_idx[b] = ld_cg(ar + (size_t)(T % p.RA) * NS + k);
        }
      } else {
        PFB_TICK(4);
        mx = block_max(mx);
        PFB_TICK(5);
        exchange(mx);
        PFB_TICK(6);
        m = xchg_max();
        PFB_TICK(7);
      }
    }
  }
};

template <int MODEL, typename Real, int BLOCK, int ITEMS, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) pf_persistent_kernel(const DeviceProblem prob) {
  extern __shared__ double pf_smem[];
  __shared__ double sh_red[(BLOCK / 32) * (kMaxParams + 2) + 40];
  __shared__ double sh_c[kNumConsts];
  PfKernel<MODEL, Real, BLOCK, ITEMS> k(prob, pf_smem, sh_red, sh_c);
  for (int b = k.q; b < prob.B; b += prob.n_groups) {
    k.run_filter(b);
    if (k.aborted) return;
  }
#ifdef PFB_PROFILE
  if (threadIdx.x == 0)
    for (int i = 0; i < 12; ++i) prob.prof[(size_t)blockIdx.x * 12 + i] = k.prof_acc[i];
#endif
}

// ---------------------------------------------------------------------------------------------
// Finalisation: combines the per-CTA partial sums in a fixed order into the outputs the reference
// returns (standard.py:96-101, 140-141, 156-169).  One block per filter.
// ---------------------------------------------------------------------------------------------
struct FinalizeArgs {
  int N, T, L, P, B, G, do_smoother;
  const double* stat_m; const double* stat_S; const double* filt_part; const double* smo_part;
  double* ll_terms;   // [B][T+1]
  double* log_like;   // [B]
  double* filt;       // [B][T+1]
  double* smo;        // [B][T+1]
  double* grad;       // [B][P][T+1]
};

__global__ void __launch_bounds__(256) pf_finalize_kernel(const FinalizeArgs a) {
  const int b = blockIdx.x;
  const int T1 = a.T + 1;
  __shared__ double sh[256];
  const double logN = log((double)a.N);
  double ll_acc = 0.0;
  for (int t = threadIdx.x; t < T1; t += blockDim.x) {
    const size_t bt = (size_t)b * T1 + t;
    const double S = a.stat_S[bt];
    double f = 0.0;
    for (int g = 0; g < a.G; ++g) f += a.filt_part[bt * a.G + g];
    const double filt = (t == 0) ? 0.0 : f / S;
    double ll = 0.0;
    if (t > 0) { ll = a.stat_m[bt]; ll += log(S); ll -= logN; }
    a.ll_terms[bt] = ll;
    a.filt[bt] = filt;
    ll_acc += ll;
    if (a.do_smoother) {
      if (t < a.T) {
        const int lag = min(t + a.L, a.T);
        const double Sl = a.stat_S[(size_t)b * T1 + lag];
        for (int k = 0; k <= a.P; ++k) {
          double s = 0.0;
          for (int g = 0; g < a.G; ++g) s += a.smo_part[(bt * a.G + g) * (a.P + 1) + k];
          s /= Sl;
          if (k == 0) a.smo[bt] = s;
          else a.grad[((size_t)b * a.P + (k - 1)) * T1 + t] = s;
        }
      } else {
        a.smo[bt] = filt;
        for (int k = 0; k < a.P; ++k) a.grad[((size_t)b * a.P + k) * T1 + t] = 0.0;
      }
    }
  }
  sh[threadIdx.x] = ll_acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.log_like[b] = sh[0];
}

// trajectory rows from the stored history: traj[b][t] = x[b][t][traj_idx[b]]
template <typename Real>
__global__ void pf_trajectory_kernel(const Real* x_hist, const int* traj_idx, double* traj, int NS, int T1) {
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < T1; t += blockDim.x)
    traj[(size_t)b * T1 + t] = (double)x_hist[((size_t)b * T1 + t) * NS + traj_idx[b]];
}

template <typename Real>
__global__ void pf_to_double_kernel(const Real* src, double* dst, size_t rows, int N, int NS) {
  const size_t n = rows * (size_t)N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (double)src[(i / N) * NS + (i % N)];
}

__global__ void pf_philox_normals_kernel(unsigned long long seed, unsigned long long eval_id, int t, int n, double* z) {
  const int pp = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * pp < n) {
    double a, b;
    philox_normal_pair<double>(seed, eval_id, (uint32_t)t, (uint32_t)pp, a, b);
    z[2 * pp] = a;
    if (2 * pp + 1 < n) z[2 * pp + 1] = b;
  }
}

__global__ void pf_philox_uniforms_kernel(unsigned long long seed, unsigned long long eval_id, int t, int count, double* u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    u[i] = (t < 0) ? philox_uniform(seed, eval_id, kStreamUFinal, 0u, 0u)
                   : philox_uniform(seed, eval_id, kStreamU, (uint32_t)t, (uint32_t)i);
  }
}

}  // namespace pfb200
