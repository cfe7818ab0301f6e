// Device-side building blocks of the B200 particle filter / fixed-lag smoother:
// Philox4x32-10 counter RNG, the per-particle model arithmetic of the three reference models,
// inter-CTA exchange primitives and block-level scan / reduce helpers.
//
// Reference behaviour restated here (paths relative to the reference repository root):
//   models/linear_gaussian_model.py:77-107,142-154,208-244          (LGSS callbacks)
//   models/stochastic_volatility_model.py:74-104,137-150,200-230     (SV callbacks)
//   models/stochastic_volatility_model_leverage.py:77-108,143-156,212-259 (SV with leverage)
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace pfb200 {

constexpr int kMaxParams = 4;
constexpr int kNumConsts = 16;
constexpr int kModelLGSS = 0, kModelSV = 1, kModelSVL = 2;
// Fully adapted (optimal-proposal) auxiliary particle filter for the LGSS model: resampling weights
// p(y_{t+1} | x_t), proposal p(x_{t+1} | x_t, y_{t+1}).  Not in the reference; named by the
// north star; pinned to the Kalman log-likelihood and to a NumPy restatement in the test tree.
constexpr int kModelLGSSFA = 3;
__host__ __device__ constexpr bool is_fully_adapted(int model) { return model == kModelLGSSFA; }

// Derived per-filter constants, computed on the host with the reference's own expressions
// (so that e.g. sigma_v**(-2) has the bits Python gives it) -- see host_consts() in pfb200.cu.
enum ConstIdx {
  C_MU = 0, C_PHI, C_SIGV, C_SIGE_OR_RHO, C_SD0, C_LOGSIGE, C_Q, C_R, C_1MPHI, C_1MPHI2,
  C_SVRHO, C_STDEV, C_RHOTERM, C_RHOINV_RHO, C_FA_A, C_FA_B
};
// fully adapted LGSS reuses: C_STDEV = sqrt(sv^2 se^2 / (sv^2+se^2)) (proposal std), C_FA_A = se^2/(sv^2+se^2),
// C_FA_B = sv^2/(sv^2+se^2) (proposal mean = A*pred + B*y), C_SVRHO = sqrt(sv^2+se^2) (predictive std),
// C_RHOTERM = log(sqrt(sv^2+se^2))

#define PFB_NORM_LOGC 0.91893853320467267  // log(sqrt(2 pi)), scipy _norm_pdf_logC

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Counter = (index, time step, eval_id lo, stream|eval_id hi),
// key = seed.  A draw is a pure function of (seed, eval_id, stream, t, index).
// ------------------------------------------------------------------------------------------
enum PhiloxStream : unsigned { kStreamZ = 0u, kStreamU = 1u, kStreamUFinal = 2u };

// hi:lo = a * M as ONE mul.wide.u32 (the C++ form `(uint64_t)M * a` reaches ptxas as a 64 x 64 multiply whose
// zero high-word terms survive as additions of a zero uniform register: 2 dead instructions per round)
__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t m, uint32_t a, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(m));
#else
  const uint64_t p = (uint64_t)m * a;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  uint32_t h0, l0, h1, l1;
  philox_mulhilo(0xD2511F53u, c[0], h0, l0);
  philox_mulhilo(0xCD9E8D57u, c[2], h1, l1);
  const uint32_t n0 = h1 ^ c[1] ^ k0;
  const uint32_t n2 = h0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t eval_id, unsigned stream,
                                                       uint32_t t, uint32_t index, uint32_t (&out)[4]) {
  uint32_t c[4] = {index, t, (uint32_t)eval_id, (uint32_t)(eval_id >> 32) ^ (stream << 28)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// Same function with the ten round keys precomputed on the host (kernel parameters live in the
// constant bank, so a key is an operand of the round's XOR instead of two additions per round).
__device__ __forceinline__ void philox4x32_10_keys(const unsigned* keys, uint64_t eval_id, unsigned stream,
                                                   uint32_t t, uint32_t index, uint32_t (&out)[4]) {
  uint32_t c[4] = {index, t, (uint32_t)eval_id, (uint32_t)(eval_id >> 32) ^ (stream << 28)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, keys[2 * r], keys[2 * r + 1]);
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// uniform in [0,1) with 53 random bits (numpy's uniform() is also [0,1))
__host__ __device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t eval_id, unsigned stream,
                                                 uint32_t t, uint32_t index) {
  uint32_t r[4];
  philox4x32_10(seed, eval_id, stream, t, index, r);
  return u01_53(r[0], r[1]);
}

// Two standard normals for the particle pair (2*pair, 2*pair+1) at time step t (Box-Muller on
// one Philox block, both branches used).
template <typename Real>
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t eval_id, uint32_t t,
                                                   uint32_t pair, Real& z_even, Real& z_odd);

// ---- fp64 Box-Muller building blocks.  The general-purpose log / sqrt / sincospi of the CUDA math
// library cost ~180 instructions per normal pair (argument classes that cannot occur here: huge,
// denormal, negative, non-finite).  These versions only cover the ranges Box-Muller produces and
// stay within a few ulp (checked against NumPy in tests/test_gpu_parity.py).

// ln(u) for u in [2^-53, 1]: fdlibm's scheme -- u = 2^e m, m in [sqrt(1/2), sqrt(2)), f = m - 1,
// s = f / (2 + f), ln(1+f) = f - (f^2/2 - s (f^2/2 + R(s^2))).  The quotient uses a float-seeded
// reciprocal refined by two Newton steps and one residual correction.
__device__ __forceinline__ double log_unit_interval(double u) {
  int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = __hiloint2double(hi, lo);  // [1, 2)
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  const double f = m - 1.0;
  const double d = 2.0 + f;
  double r = (double)__frcp_rn((float)d);
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  double sq = f * r;
  sq = fma(fma(-d, sq, f), r, sq);
  const double z = sq * sq;
  double R = 1.479819860511658591e-01;
  R = fma(R, z, 1.531383769920937332e-01);
  R = fma(R, z, 1.818357216161805012e-01);
  R = fma(R, z, 2.222219843214978396e-01);
  R = fma(R, z, 2.857142874366239149e-01);
  R = fma(R, z, 3.999999999940941908e-01);
  R = fma(R, z, 6.666666666666735130e-01);
  R *= z;
  const double hfsq = 0.5 * f * f;
  const double ef = (double)e;
  return ef * 6.93147180369123816490e-01 - ((hfsq - (sq * (hfsq + R) + ef * 1.90821492927058770002e-10)) - f);
}

// sqrt(x) for 0 <= x < 2^100: float-seeded reciprocal square root, two Newton steps, one correction
__device__ __forceinline__ double sqrt_nonneg(double x) {
  double y = (double)rsqrtf((float)x);
  const double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  double r = x * y;
  r = fma(fma(-r, r, x), 0.5 * y, r);
  return x > 0.0 ? r : 0.0;
}

// sin(pi a), cos(pi a) for a in [0, 2): quadrant reduction a = k/2 + r, |r| <= 1/4, Taylor
// polynomials in (pi r) (truncation < 5e-17)
__device__ __forceinline__ void sincospi_0_2(double a, double& sn, double& cs) {
  const double kf = rint(2.0 * a);
  const int k = (int)kf;
  const double r = fma(kf, -0.5, a);
  const double r2 = r * r;
  double ps = 7.95205400147550838e-07;
  ps = fma(ps, r2, -2.19153534478302037e-05);
  ps = fma(ps, r2, 4.66302805767612337e-04);
  ps = fma(ps, r2, -7.37043094571434784e-03);
  ps = fma(ps, r2, 8.21458866111281910e-02);
  ps = fma(ps, r2, -5.99264529320791883e-01);
  ps = fma(ps, r2, 2.55016403987734508e+00);
  ps = fma(ps, r2, -5.16771278004996937e+00);
  ps = fma(ps, r2, 3.14159265358979312e+00);
  ps *= r;
  double pc = -1.38789524622137608e-07;
  pc = fma(pc, r2, 4.30306958703294391e-06);
  pc = fma(pc, r2, -1.04638104924845650e-04);
  pc = fma(pc, r2, 1.92957430940392206e-03);
  pc = fma(pc, r2, -2.58068913900140508e-02);
  pc = fma(pc, r2, 2.35330630358893123e-01);
  pc = fma(pc, r2, -1.33526276885458928e+00);
  pc = fma(pc, r2, 4.05871212641676760e+00);
  pc = fma(pc, r2, -4.93480220054467900e+00);
  pc = fma(pc, r2, 1.0);
  const bool swap = k & 1;
  const double s0 = swap ? pc : ps, c0 = swap ? ps : pc;
  sn = (k & 2) ? -s0 : s0;
  cs = (((k + 1) & 2) != 0) ? -c0 : c0;
}

__device__ __forceinline__ void normal_pair_from_bits(const uint32_t (&r)[4], double& z_even, double& z_odd);
__device__ __forceinline__ void normal_pair_from_bits(const uint32_t (&r)[4], float& z_even, float& z_odd);
// hot-path variant: round keys from the kernel parameters
template <typename Real>
__device__ __forceinline__ void philox_normal_pair_keys(const unsigned* keys, uint64_t eval_id, uint32_t t,
                                                        uint32_t pair, Real& z_even, Real& z_odd) {
  uint32_t r[4];
  philox4x32_10_keys(keys, eval_id, kStreamZ, t, pair, r);
  normal_pair_from_bits(r, z_even, z_odd);
}

template <>
__device__ __forceinline__ void philox_normal_pair<double>(uint64_t seed, uint64_t eval_id, uint32_t t,
                                                           uint32_t pair, double& z_even, double& z_odd) {
  uint32_t r[4];
  philox4x32_10(seed, eval_id, kStreamZ, t, pair, r);
  normal_pair_from_bits(r, z_even, z_odd);
}
__device__ __forceinline__ void normal_pair_from_bits(const uint32_t (&r)[4], double& z_even, double& z_odd) {
  const double u1 = ((double)((((uint64_t)r[0] << 32) | r[1]) >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0,1]
  const double u2 = u01_53(r[2], r[3]);
  const double rad = sqrt_nonneg(-2.0 * log_unit_interval(u1));
  double s, c;
  sincospi_0_2(2.0 * u2, s, c);
  z_even = rad * c;
  z_odd = rad * s;
}

template <>
__device__ __forceinline__ void philox_normal_pair<float>(uint64_t seed, uint64_t eval_id, uint32_t t,
                                                          uint32_t pair, float& z_even, float& z_odd) {
  uint32_t r[4];
  philox4x32_10(seed, eval_id, kStreamZ, t, pair, r);
  normal_pair_from_bits(r, z_even, z_odd);
}
__device__ __forceinline__ void normal_pair_from_bits(const uint32_t (&r)[4], float& z_even, float& z_odd) {
  const float u1 = ((float)(r[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
  const float u2 = (float)(r[2] >> 8) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * __logf(u1));
  float s, c;
  // angle 2 pi u2 - pi in [-pi, pi): the range in which the fast sine / cosine keep their 2^-21 absolute error
  // (the same order as __logf above; any uniform angle serves Box-Muller)
  __sincosf(fmaf(u2, 6.28318530717958648f, -3.14159265358979324f), &s, &c);
  z_even = rad * c;
  z_odd = rad * s;
}

// Shared-memory access through a 32-bit shared-window address held in a register.  Indexing a
// generic pointer to shared memory makes the compiler rematerialise the window base
// (S2R SR_CgaCtaId + LEA) in front of every load -- in the binary-search loop that costs more
// than the load itself -- so hot loops keep an opaque base address and use ld.shared directly.
__device__ __forceinline__ uint32_t smem_base(const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("mov.u32 %0, %0;" : "+r"(a));  // opaque: must live in a register
  return a;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// ------------------------------------------------------------------------------------------
// Per-particle model arithmetic.  The fp64 propagate / weight paths pin the reference's operation
// order with round-to-nearest intrinsics (never contracted to FMA), so that with injected noise a
// particle has the same bits as in the NumPy reference.  fp32 uses fast intrinsics.
// ------------------------------------------------------------------------------------------
template <int MODEL, typename Real>
struct ModelOps;

// constants read by propagate() / logweight(): loaded once per phase into registers
template <int MODEL>
__device__ __forceinline__ void load_step_consts(uint32_t c_addr, double (&c)[kNumConsts]) {
#pragma unroll
  for (int i = 0; i < kNumConsts; ++i) c[i] = 0.0;
  c[C_MU] = lds_f64(c_addr + 8 * C_MU);
  c[C_PHI] = lds_f64(c_addr + 8 * C_PHI);
  if (MODEL == kModelSVL) {
    c[C_SVRHO] = lds_f64(c_addr + 8 * C_SVRHO);
    c[C_STDEV] = lds_f64(c_addr + 8 * C_STDEV);
  } else {
    c[C_SIGV] = lds_f64(c_addr + 8 * C_SIGV);
  }
  if (MODEL == kModelLGSS) {
    c[C_SIGE_OR_RHO] = lds_f64(c_addr + 8 * C_SIGE_OR_RHO);
    c[C_LOGSIGE] = lds_f64(c_addr + 8 * C_LOGSIGE);
    c[C_FA_A] = lds_f64(c_addr + 8 * C_FA_A);  // LGSS: 1 / sigma_e
  }
  if (MODEL == kModelLGSSFA) {
    c[C_STDEV] = lds_f64(c_addr + 8 * C_STDEV);
    c[C_FA_A] = lds_f64(c_addr + 8 * C_FA_A);
    c[C_FA_B] = lds_f64(c_addr + 8 * C_FA_B);
    c[C_SVRHO] = lds_f64(c_addr + 8 * C_SVRHO);
    c[C_RHOTERM] = lds_f64(c_addr + 8 * C_RHOTERM);
  }
}

template <int MODEL>
struct ModelOps<MODEL, double> {
  // generate_initial_state: mu + (sigma_v / sqrt(1 - phi**2)) * z
  static __device__ __forceinline__ double init(const double* c, double z) {
    return __dadd_rn(c[C_MU], __dmul_rn(c[C_SD0], z));
  }
  // generate_state; y is the observation the model reads for the leverage term
  static __device__ __forceinline__ double propagate(const double* c, double x, double z, double y) {
    double mean = __dadd_rn(c[C_MU], __dmul_rn(c[C_PHI], __dsub_rn(x, c[C_MU])));
    if (MODEL == kModelLGSSFA) {  // x' ~ p(x' | x, y) = N(A pred + B y, s^2)
      const double pm = __dadd_rn(__dmul_rn(c[C_FA_A], mean), __dmul_rn(c[C_FA_B], y));
      return __dadd_rn(pm, __dmul_rn(c[C_STDEV], z));
    }
    if (MODEL == kModelSVL) {
      const double lev = __dmul_rn(__dmul_rn(c[C_SVRHO], exp(__dmul_rn(-0.5, x))), y);
      mean = __dadd_rn(mean, lev);
      return __dadd_rn(mean, __dmul_rn(c[C_STDEV], z));
    }
    return __dadd_rn(mean, __dmul_rn(c[C_SIGV], z));
  }
  // evaluate_obs = scipy.stats.norm.logpdf
  static __device__ __forceinline__ double logweight(const double* c, double x, double y) {
    if (MODEL == kModelLGSS) {
      // (y - x) * (1 / sigma_e): one rounding more than the reference's division (<= 1 ulp on the
      // log-weight; particles are untouched), 12 instructions less on the filter warps' critical path
      const double zz = __dmul_rn(__dsub_rn(y, x), c[C_FA_A]);
      return __dsub_rn(__dsub_rn(__dmul_rn(-__dmul_rn(zz, zz), 0.5), PFB_NORM_LOGC), c[C_LOGSIGE]);
    } else if (MODEL == kModelLGSSFA) {  // log p(y_next | x) = N(y; mu + phi (x - mu), sv^2 + se^2)
      const double pred = __dadd_rn(c[C_MU], __dmul_rn(c[C_PHI], __dsub_rn(x, c[C_MU])));
      const double zz = __ddiv_rn(__dsub_rn(y, pred), c[C_SVRHO]);
      return __dsub_rn(__dsub_rn(__dmul_rn(-__dmul_rn(zz, zz), 0.5), PFB_NORM_LOGC), c[C_RHOTERM]);
    } else {
      // norm.logpdf(y, 0, exp(x / 2)) with 1 / vol = exp(-x / 2) and log(vol) = x / 2: ONE exp instead of
      // exp -> division -> log (a dependent chain of ~150 instructions on the filter warps' critical path; 10 % of
      // the step at C4's shape).  Differs from the reference's expression by a few ulp of the terms (tests: 1e-13).
      const double zz = __dmul_rn(y, exp(__dmul_rn(-0.5, x)));
      return __dsub_rn(__dsub_rn(__dmul_rn(-__dmul_rn(zz, zz), 0.5), PFB_NORM_LOGC), __dmul_rn(0.5, x));
    }
  }
  // log_joint_gradient(next_state, cur_state, time_index); out[p] in parameter order
  static __device__ __forceinline__ void gradient(const double* c, double xn, double xc, double y,
                                                  double (&out)[kMaxParams]) {
    double q = xn - c[C_MU];
    q = q - c[C_PHI] * (xc - c[C_MU]);
    double e = 0.0;
    if (MODEL == kModelSVL) {
      e = exp(-0.5 * xc);
      q = q - c[C_SVRHO] * e * y;
    }
    const double Q = c[C_Q];
    out[0] = Q * q * c[C_1MPHI];
    out[1] = Q * q * (xc - c[C_MU]) * c[C_1MPHI2];
    out[2] = Q * (q * q) - 1.0;
    if (MODEL == kModelLGSS || MODEL == kModelLGSSFA) {
      const double oq = y - xc;
      out[3] = c[C_R] * (oq * oq) - 1.0;
    } else if (MODEL == kModelSVL) {
      const double rho = c[C_SIGE_OR_RHO], sv = c[C_SIGV];
      out[2] = out[2] + q * rho * sv * y * e * Q * sv;
      double gr = c[C_RHOINV_RHO];
      gr = gr + q * sv * y * e * Q;
      gr = gr - (-rho * (q * q) * Q / c[C_RHOTERM]);
      out[3] = gr * c[C_RHOTERM];
    } else {
      out[3] = 0.0;
    }
  }
};

template <int MODEL>
struct ModelOps<MODEL, float> {
  static __device__ __forceinline__ float init(const double* c, float z) {
    return (float)c[C_MU] + (float)c[C_SD0] * z;
  }
  static __device__ __forceinline__ float propagate(const double* c, float x, float z, double y) {
    const float mu = (float)c[C_MU];
    float mean = mu + (float)c[C_PHI] * (x - mu);
    if (MODEL == kModelLGSSFA) return (float)c[C_FA_A] * mean + (float)c[C_FA_B] * (float)y + (float)c[C_STDEV] * z;
    if (MODEL == kModelSVL) {
      mean += (float)c[C_SVRHO] * __expf(-0.5f * x) * (float)y;
      return mean + (float)c[C_STDEV] * z;
    }
    return mean + (float)c[C_SIGV] * z;
  }
  static __device__ __forceinline__ float logweight(const double* c, float x, double y) {
    if (MODEL == kModelLGSS) {
      const float zz = ((float)y - x) * (float)(1.0 / c[C_SIGE_OR_RHO]);
      return -0.5f * zz * zz - (float)PFB_NORM_LOGC - (float)c[C_LOGSIGE];
    } else if (MODEL == kModelLGSSFA) {
      const float mu = (float)c[C_MU];
      const float zz = ((float)y - (mu + (float)c[C_PHI] * (x - mu))) * (float)(1.0 / c[C_SVRHO]);
      return -0.5f * zz * zz - (float)PFB_NORM_LOGC - (float)c[C_RHOTERM];
    } else {
      // -0.5 y^2 exp(-x) - logC - x/2   (log(exp(x/2)) folded analytically in fp32)
      const float yy = (float)y;
      return -0.5f * yy * yy * __expf(-x) - (float)PFB_NORM_LOGC - 0.5f * x;
    }
  }
  static __device__ __forceinline__ void gradient(const double* c, float xn, float xc, double y,
                                                  double (&out)[kMaxParams]) {
    // score terms are accumulated in fp64 from the fp32 states
    ModelOps<MODEL, double>::gradient(c, (double)xn, (double)xc, y, out);
  }
};

// ------------------------------------------------------------------------------------------
// exp(x) for x <= 0 (shifted log-weights): round-to-nearest range reduction x = k ln2 + r,
// |r| <= ln2/2, degree-13 Taylor polynomial (truncation < 5e-18), scaling by 2^k through the
// exponent bits.  Results below 2^-1021 (weights 1e-307 times the maximum) flush to zero; NaN
// propagates.  About 1 ulp, ~4x fewer instructions than the general-purpose exp().
// ------------------------------------------------------------------------------------------
// Taylor coefficients 1/13! ... 1/0! in constant memory: DFMA takes them as constant-bank operands
// (a 64-bit immediate would cost two extra instructions per Horner step).
static __constant__ double kExpCoef[14] = {
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.3888888888888889e-03,
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0};
static __constant__ double kExpLn2[3] = {1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10};

__device__ __forceinline__ double exp_nonpos(double x) {
  const double t = x * kExpLn2[0];
  if (!(t > -1021.0)) return (t == t) ? 0.0 : t;  // underflow -> 0, NaN -> NaN
  const int k = __double2int_rn(t);
  const double kf = (double)k;
  double r = fma(kf, kExpLn2[1], x);
  r = fma(kf, kExpLn2[2], r);
  double q = kExpCoef[0];
#pragma unroll
  for (int i = 1; i < 14; ++i) q = fma(q, r, kExpCoef[i]);
  return q * __hiloint2double((k + 1023) << 20, 0);
}

template <typename Real>
__device__ __forceinline__ double shifted_weight(Real logw, double m);
template <>
__device__ __forceinline__ double shifted_weight<double>(double logw, double m) {
  return exp_nonpos(logw - m);
}
template <>
__device__ __forceinline__ double shifted_weight<float>(float logw, double m) {
  return (double)__expf(logw - (float)m);
}

// ------------------------------------------------------------------------------------------
// Memory-ordering helpers for the inter-CTA exchange (gpu scope).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// CTA-scope flags in shared memory (filter warps <-> smoother warps of one CTA)
__device__ __forceinline__ unsigned ld_acquire_cta_shared_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// named barriers: the filter warps and the smoother warps of a CTA synchronise separately
template <int ID, int COUNT>
__device__ __forceinline__ void named_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
template <int ID, int COUNT>
__device__ __forceinline__ bool named_sync_or(bool pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
      : "=r"(r) : "r"((int)pred), "n"(ID), "n"(COUNT) : "memory");
  return r != 0;
}
// system-scope variants for the exchange between the GPUs of a distributed filter (peer memory)
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_sys(const int* p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_i32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// L2-coherent loads for data produced by other CTAs of the same launch (never served from a
// stale L1 line; the rings reuse addresses every few steps).
template <typename T>
__device__ __forceinline__ T ld_cg(const T* p) {
  return __ldcg(p);
}

// L2 eviction-priority hints.  The ancestor ring is re-read on every time step for L steps while
// particles / log-weights stream through once or twice, and the whole ring set is larger than what
// L2 keeps: ancestors are tagged evict_last, one-shot reads of old generations evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ int ld_hint(const int* p, uint64_t pol) {
  int v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ double ld_hint(const double* p, uint64_t pol) {
  double v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ float ld_hint(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
// the same loads at a compile-time byte offset from one base pointer (address = register pair + immediate:
// no per-load 64-bit address arithmetic)
template <int OFF>
__device__ __forceinline__ double ld_hint_off(const double* p, uint64_t pol) {
  double v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.f64 %0, [%1+%3], %2;" : "=d"(v) : "l"(p), "l"(pol), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ float ld_hint_off(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.f32 %0, [%1+%3], %2;" : "=f"(v) : "l"(p), "l"(pol), "n"(OFF) : "memory");
  return v;
}
__device__ __forceinline__ void st_hint_int2(int* p, int a, int b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.s32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(a), "r"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint_int4(int* p, int a, int b, int c, int d, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint_int(int* p, int a, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(a), "l"(pol) : "memory");
}

// ------------------------------------------------------------------------------------------
// Warp / block helpers (fixed reduction trees: results do not depend on scheduling)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_inclusive_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

}  // namespace pfb200
