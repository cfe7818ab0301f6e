// C-ABI host side of libpfb200.so (see include/pfb200.h): device workspace management, launch
// geometry, H2D/D2H staging.  No PyTorch types; plain CUDA runtime.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pfb200.h"
#include "pf_config.h"
#define PFB_UTILITY_KERNELS
#include "pf_kernels.cuh"
#include "pf_cluster.cuh"

using namespace pfb200;

typedef void (*KernelFn)(const pfb200::DeviceProblem);
// getters of the 16 instantiation objects (pf_instance.cu): pfb200_kernel_<model>_<dtype>_<general>(items)
#define PFB_DECLARE_MODEL(m)                                                     \
  KernelFn pfb200_kernel_##m##_0_0(int); KernelFn pfb200_kernel_##m##_0_1(int);   \
  KernelFn pfb200_kernel_##m##_1_0(int); KernelFn pfb200_kernel_##m##_1_1(int);   \
  KernelFn pfb200_cluster_kernel_##m##_0_0(); KernelFn pfb200_cluster_kernel_##m##_0_1(); \
  KernelFn pfb200_cluster_kernel_##m##_1_0(); KernelFn pfb200_cluster_kernel_##m##_1_1();
#define PFB_PICK_CLUSTER(m)                                                                        \
  return dtype == 0 ? (general ? pfb200_cluster_kernel_##m##_0_1() : pfb200_cluster_kernel_##m##_0_0())  \
                    : (general ? pfb200_cluster_kernel_##m##_1_1() : pfb200_cluster_kernel_##m##_1_0());
#define PFB_PICK_MODEL(m)                                                                        \
  return dtype == 0 ? (general ? pfb200_kernel_##m##_0_1(items) : pfb200_kernel_##m##_0_0(items))  \
                    : (general ? pfb200_kernel_##m##_1_1(items) : pfb200_kernel_##m##_1_0(items));
PFB_DECLARE_MODEL(0)
#ifndef PFB_LGSS_ONLY
PFB_DECLARE_MODEL(1)
PFB_DECLARE_MODEL(2)
PFB_DECLARE_MODEL(3)
#endif

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                    \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess)                                                             \
      return fail(err__ == cudaErrorMemoryAllocation ? PFB200_ENOMEM : PFB200_ECUDA,      \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

// arrival counters: 32 words per group; a distributed filter (one group) spreads its system-scope counters over
// 2 x 32 slots of one 128-byte line each behind them (pf_kernels.cuh: kDistX / kDistS)
static size_t counter_bytes(int n_groups) { return (size_t)(n_groups * 32 > 2080 ? n_groups * 32 : 2080) * 4; }

struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    if (bytes == 0) return cudaSuccess;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&ptr, want);
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(ptr); }
};

constexpr int kBlock = PFB_BLOCK;                       // filter threads per CTA
constexpr int kThreads = PFB_BLOCK + PFB_SMOOTH_THREADS;  // + the smoother warpgroup
constexpr int kItems = PFB_ITEMS;        // default tile granularity
constexpr int kItemsAlt = PFB_ITEMS_ALT;  // alternative, chosen when it saves a scan sub-chunk
constexpr int kItemsSmall = PFB_ITEMS_SMALL;  // small filters: a CTA's parent range fits 2 per thread
constexpr int kChunk = kBlock * kItems;


// The persistent kernel is instantiated in pf_instance.cu, compiled once per (model, dtype, flavour) into
// separate objects (built in parallel).
KernelFn kernel_for_model(int model, int dtype, bool general, int items) {
  switch (model) {
    case kModelLGSS: PFB_PICK_MODEL(0)
#ifndef PFB_LGSS_ONLY
    case kModelSV: PFB_PICK_MODEL(1)
    case kModelSVL: PFB_PICK_MODEL(2)
    case kModelLGSSFA: PFB_PICK_MODEL(3)
#endif
  }
  return nullptr;
}

// the cluster kernel (pf_cluster.cuh) of a model; nullptr: the model has none (fully adapted filter)
KernelFn cluster_kernel_for_model(int model, int dtype, bool general) {
  switch (model) {
    case kModelLGSS: PFB_PICK_CLUSTER(0)
#ifndef PFB_LGSS_ONLY
    case kModelSV: PFB_PICK_CLUSTER(1)
    case kModelSVL: PFB_PICK_CLUSTER(2)
#endif
  }
  return nullptr;
}

// Derived constants with the reference's own expressions (Python float arithmetic == C double):
//   linear_gaussian_model.py:87-90,223-237; stochastic_volatility_model_leverage.py:103-108,227-252
void host_consts(int model, const double* th, double* c) {
  for (int i = 0; i < kNumConsts; ++i) c[i] = 0.0;
  const double mu = th[0], phi = th[1], sv = th[2];
  c[C_MU] = mu;
  c[C_PHI] = phi;
  c[C_SIGV] = sv;
  c[C_SD0] = sv / std::sqrt(1.0 - std::pow(phi, 2.0));
  c[C_Q] = std::pow(sv, -2.0);
  c[C_1MPHI] = 1.0 - phi;
  c[C_1MPHI2] = 1.0 - std::pow(phi, 2.0);
  if (model == kModelLGSS || model == kModelLGSSFA) {
    const double se = th[3];
    c[C_SIGE_OR_RHO] = se;
    c[C_LOGSIGE] = std::log(se);
    c[C_R] = std::pow(se, -2.0);
    if (model == kModelLGSS) c[C_FA_A] = 1.0 / se;  // reciprocal used by the device log-weight
    if (model == kModelLGSSFA) {
      const double tot = sv * sv + se * se;
      c[C_STDEV] = std::sqrt(sv * sv * se * se / tot);
      c[C_FA_A] = se * se / tot;
      c[C_FA_B] = sv * sv / tot;
      c[C_SVRHO] = std::sqrt(tot);
      c[C_RHOTERM] = std::log(std::sqrt(tot));
    }
  } else if (model == kModelSVL) {
    const double rho = th[3];
    const double rho_term = 1.0 - std::pow(rho, 2.0);
    c[C_SIGE_OR_RHO] = rho;
    c[C_SVRHO] = sv * rho;
    c[C_STDEV] = std::sqrt(1.0 - std::pow(rho, 2.0)) * sv;
    c[C_RHOTERM] = rho_term;
    c[C_Q] = std::pow(sv, -2.0) * std::pow(rho_term, -1.0);
    c[C_RHOINV_RHO] = std::pow(rho_term, -1.0) * rho;
  }
}

}  // namespace

struct pfb200_handle {
  int device = 0;
  int model = 0;
  int dtype = 0;
  int P = 0;
  int num_sms = 0;
  unsigned long long pol_keep = 0, pol_once = 0;  // L2 cache-policy descriptors (device-created)
  int blocks_per_sm = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
  bool timed = false;
  KernelFn kernel = nullptr;          // flavour selected by the configured problem
  int items = PFB_ITEMS;              // tile granularity of the selected kernel
  // options
  long long opt_ctas_per_filter = 0;
  long long opt_max_groups = 0;
  long long opt_spin_limit = 20000000LL;
  long long opt_l2_persist = 0;
  long long opt_anc_smem = 1;       // shared-memory ancestor ring for one-CTA filters
  long long opt_cluster = 0;        // cluster kernel: 0 = automatic, -1 = never, 1..16 = CTAs per filter (power of two)
  long long opt_cluster_per = 128;  // automatic choice: parents per CTA the cluster size aims at
  int cluster = 0;                  // CTAs per cluster of the configured problem (0: persistent kernel)
  long long opt_items = 0;          // 0 = automatic tile granularity, else PFB_ITEMS / PFB_ITEMS_ALT   // pin the ancestor ring in L2 (access policy window)
  size_t l2_persist_max = 0, l2_window_max = 0;
  // distributed single filter (one process per GPU, peer memory over NVLink)
  long long opt_dist_world = 1, opt_dist_rank = 0;
  bool dist_connected = false;
  bool dist_same_device = false;  // ranks of one process on ONE device (single-GPU test of the peer paths):
                                  // plain launches, co-residency by construction (R * G_loc <= SM count)
  unsigned dist_epoch = 0;
  unsigned dist_sjobs = 0;
  void* peer_opened[8][6] = {};  // IPC mappings to close
  long long dist_geom[8] = {};   // geometry the peer mappings were made for (see dist_geometry)
  // configured problem
  bool configured = false;
  bool executed = false;
  DeviceProblem prob{};
  unsigned flags = 0;
  size_t smem_bytes = 0;
  bool cooperative = false;
  int n_ws = 0;
  // device buffers
  DevBuf obs, consts, evals, inj_z0, inj_u, inj_z, inj_uf;
  DevBuf x_ring, logw, anc, s_ring, cdf_g, xval, xcount, abortf;
  DevBuf stat_m, stat_S, filt_part, smo_part, traj_k, traj_idx;
  DevBuf ll_terms, log_like, filt, smo, grad, traj, scratch, prof;
  std::vector<double> h_consts;
  std::vector<unsigned long long> h_evals;
};

constexpr int kDistBufs = 6;  // exported per rank: particles, log-weights, ancestors, exchange values, counters, CDF copy

static void dist_close(pfb200_handle* h) {
  for (int r = 0; r < 8; ++r)
    for (int k = 0; k < kDistBufs; ++k)
      if (h->peer_opened[r][k]) { cudaIpcCloseMemHandle(h->peer_opened[r][k]); h->peer_opened[r][k] = nullptr; }
  h->dist_connected = false;
  h->dist_same_device = false;
}

extern "C" {

int pfb200_abi_version(void) { return PFB200_ABI_VERSION; }

const char* pfb200_last_error(void) { return g_last_error.c_str(); }

int pfb200_model_no_params(int model_id) {
  switch (model_id) {
    case PFB200_MODEL_LGSS: return 4;
    case PFB200_MODEL_SV: return 3;
    case PFB200_MODEL_SVL: return 4;
    case PFB200_MODEL_LGSS_FA: return 4;
  }
  return PFB200_EINVAL;
}

int pfb200_create(int device, int model_id, int dtype, pfb200_handle** out) {
  if (!out) return fail(PFB200_EINVAL, "pfb200_create: out is NULL");
  *out = nullptr;
  const int P = pfb200_model_no_params(model_id);
  if (P < 0) return fail(PFB200_EINVAL, "pfb200_create: unknown model_id %d", model_id);
  if (dtype != PFB200_F64 && dtype != PFB200_F32) return fail(PFB200_EINVAL, "pfb200_create: unknown dtype %d", dtype);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    return fail(PFB200_ENODEVICE, "pfb200_create: no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= count) return fail(PFB200_EINVAL, "pfb200_create: device %d out of range (%d devices)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  pfb200_handle* h = new pfb200_handle();
  h->device = device;
  h->model = model_id;
  h->dtype = dtype;
  h->P = P;
  h->kernel = kernel_for_model(model_id, dtype, false, kItems);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  h->num_sms = prop.multiProcessorCount;
  h->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
  h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  CUDA_TRY(cudaEventCreate(&h->ev0));
  CUDA_TRY(cudaEventCreate(&h->ev1));
  CUDA_TRY(cudaEventCreate(&h->evk0));
  CUDA_TRY(cudaEventCreate(&h->evk1));
  {
    unsigned long long* dpol = nullptr;
    unsigned long long hpol[2] = {0, 0};
    CUDA_TRY(cudaMalloc(&dpol, sizeof(hpol)));
    pf_policy_kernel<<<1, 1, 0, h->stream>>>(dpol);
    CUDA_TRY(cudaMemcpyAsync(hpol, dpol, sizeof(hpol), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaFree(dpol));
    h->pol_keep = hpol[0];
    h->pol_once = hpol[1];
  }
  *out = h;
  return PFB200_OK;
}

int pfb200_destroy(pfb200_handle* h) {
  if (!h) return PFB200_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  dist_close(h);
  DevBuf* bufs[] = {&h->obs, &h->consts, &h->evals, &h->inj_z0, &h->inj_u, &h->inj_z, &h->inj_uf, &h->x_ring,
                    &h->logw, &h->anc, &h->s_ring, &h->cdf_g, &h->xval, &h->xcount, &h->abortf, &h->stat_m, &h->stat_S, &h->filt_part,
                    &h->smo_part, &h->traj_k, &h->traj_idx, &h->ll_terms, &h->log_like, &h->filt, &h->smo,
                    &h->grad, &h->traj, &h->scratch, &h->prof};
  for (DevBuf* b : bufs) b->release();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->evk0) cudaEventDestroy(h->evk0);
  if (h->evk1) cudaEventDestroy(h->evk1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return PFB200_OK;
}

int pfb200_set_stream(pfb200_handle* h, void* cuda_stream) {
  if (!h) return fail(PFB200_EINVAL, "pfb200_set_stream: NULL handle");
  h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return PFB200_OK;
}

int pfb200_set_option(pfb200_handle* h, const char* name, long long value) {
  if (!h || !name) return fail(PFB200_EINVAL, "pfb200_set_option: NULL argument");
  if (!strcmp(name, "ctas_per_filter")) h->opt_ctas_per_filter = value;
  else if (!strcmp(name, "max_groups")) h->opt_max_groups = value;
  else if (!strcmp(name, "spin_limit")) h->opt_spin_limit = value;
  else if (!strcmp(name, "l2_persist")) h->opt_l2_persist = value;
  else if (!strcmp(name, "items")) {
    if (value != 0 && value != kItems && value != kItemsAlt && value != kItemsSmall)
      return fail(PFB200_EINVAL, "pfb200_set_option: items must be 0 (automatic), %d, %d or %d", kItems, kItemsAlt, kItemsSmall);
    h->opt_items = value;
  }
  else if (!strcmp(name, "anc_smem")) h->opt_anc_smem = value;
  else if (!strcmp(name, "cluster")) {
    if (value < -1 || value > kClMaxCluster || (value > 0 && (value & (value - 1))))
      return fail(PFB200_EINVAL, "pfb200_set_option: cluster must be -1 (never), 0 (automatic) or a power of two <= %d", kClMaxCluster);
    h->opt_cluster = value;
  } else if (!strcmp(name, "cluster_per")) {
    if (value < 2 || value > kClMaxPer) return fail(PFB200_EINVAL, "pfb200_set_option: cluster_per must be 2..%d", kClMaxPer);
    h->opt_cluster_per = value;
  }
  else if (!strcmp(name, "dist_world")) {
    if (value < 1 || value > 8) return fail(PFB200_EINVAL, "pfb200_set_option: dist_world must be 1..8");
    h->opt_dist_world = value;
  } else if (!strcmp(name, "dist_rank")) h->opt_dist_rank = value;
  else return fail(PFB200_EINVAL, "pfb200_set_option: unknown option '%s'", name);
  h->configured = false;
  return PFB200_OK;
}

int pfb200_get_info(pfb200_handle* h, const char* name, long long* value) {
  if (!h || !name || !value) return fail(PFB200_EINVAL, "pfb200_get_info: NULL argument");
  if (!strcmp(name, "num_sms")) *value = h->num_sms;
  else if (!strcmp(name, "ctas_per_filter")) *value = h->prob.G;
  else if (!strcmp(name, "n_groups")) *value = h->prob.n_groups;
  else if (!strcmp(name, "per_cta")) *value = h->prob.per_cta;
  else if (!strcmp(name, "block_threads")) *value = kThreads;  // filter warps + the smoother warpgroup
  else if (!strcmp(name, "filter_threads")) *value = kBlock;
  else if (!strcmp(name, "chunk")) *value = kBlock * h->items;
  else if (!strcmp(name, "items")) *value = h->items;
  else if (!strcmp(name, "anc_smem")) *value = h->prob.anc_smem;
  else if (!strcmp(name, "cluster")) *value = h->cluster;
  else if (!strcmp(name, "tile_cap")) *value = h->prob.tile_cap;
  else if (!strcmp(name, "ctas_per_rank")) *value = h->prob.G_loc;
  else if (!strcmp(name, "span")) *value = h->prob.span;
  else if (!strcmp(name, "l2_persist_max")) *value = (long long)h->l2_persist_max;
  else if (!strcmp(name, "l2_window_max")) *value = (long long)h->l2_window_max;
  else if (!strcmp(name, "blocks_per_sm")) *value = h->blocks_per_sm;
  else if (!strcmp(name, "cooperative")) *value = h->cooperative ? 1 : 0;
  else if (!strcmp(name, "smem_bytes")) *value = (long long)h->smem_bytes;
  else if (!strcmp(name, "kernel_launches_per_execute")) *value = 3 + ((h->flags & (PFB200_FLAG_STORE_HISTORY | PFB200_FLAG_STORE_STATES)) ? 1 : 0);
  else if (!strcmp(name, "workspace_bytes")) {
    *value = (long long)(h->x_ring.cap + h->logw.cap + h->anc.cap + h->filt_part.cap + h->smo_part.cap);
  } else return fail(PFB200_EINVAL, "pfb200_get_info: unknown key '%s'", name);
  return PFB200_OK;
}

// seed + its Philox4x32 key schedule: key_r = key_0 + r * (0x9E3779B9, 0xBB67AE85)
static void set_seed(DeviceProblem& p, uint64_t seed) {
  p.seed = seed;
  for (int r = 0; r < 10; ++r) {
    p.pkeys[2 * r] = (unsigned)seed + (unsigned)r * 0x9E3779B9u;
    p.pkeys[2 * r + 1] = (unsigned)(seed >> 32) + (unsigned)r * 0xBB67AE85u;
  }
}

static int upload_params(pfb200_handle* h, const double* theta, uint64_t seed, const uint64_t* eval_ids) {
  const int B = h->prob.B;
  h->h_consts.resize((size_t)B * kNumConsts);
  h->h_evals.resize(B);
  for (int b = 0; b < B; ++b) {
    host_consts(h->model, theta + (size_t)b * h->P, h->h_consts.data() + (size_t)b * kNumConsts);
    h->h_evals[b] = eval_ids ? eval_ids[b] : (unsigned long long)b;
  }
  set_seed(h->prob, seed);
  CUDA_TRY(cudaMemcpyAsync(h->consts.ptr, h->h_consts.data(), h->h_consts.size() * sizeof(double),
                           cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->evals.ptr, h->h_evals.data(), h->h_evals.size() * sizeof(unsigned long long),
                           cudaMemcpyHostToDevice, h->stream));
  // the staging vectors are pageable: the copies above have completed on return
  return PFB200_OK;
}

int pfb200_configure(pfb200_handle* h, int n_filters, int no_particles, int no_obs, int fixed_lag,
                     int do_smoother, int generate_initial_state, double initial_state,
                     int resampling, unsigned flags, const double* obs, const double* theta,
                     uint64_t seed, const uint64_t* eval_ids, const double* inj_z0,
                     const double* inj_u, const double* inj_z, const double* inj_u_final) {
  if (!h) return fail(PFB200_EINVAL, "pfb200_configure: NULL handle");
  h->configured = false;
  h->executed = false;
  const int B = n_filters, N = no_particles, T = no_obs, L = fixed_lag, P = h->P;
  if (B < 1 || N < 1 || T < 1 || L < 0)
    return fail(PFB200_EINVAL, "pfb200_configure: need n_filters>=1, no_particles>=1, no_obs>=1, fixed_lag>=0 (got %d,%d,%d,%d)", B, N, T, L);
  if (N > (1 << 30)) return fail(PFB200_EINVAL, "pfb200_configure: no_particles %d exceeds 2^30", N);
  // fixed_lag = 0 with the smoother: zero hops, smo == filt (standard.py:146-157); the score terms of such a
  // run are undefined (the reference raises NameError at standard.py:161 when it is asked for them) and are
  // returned as zeros
  if (resampling != PFB200_RESAMPLE_SYSTEMATIC && resampling != PFB200_RESAMPLE_STRATIFIED &&
      resampling != PFB200_RESAMPLE_MULTINOMIAL)
    return fail(PFB200_EINVAL, "pfb200_configure: unsupported resampling method %d", resampling);
  if (!obs || !theta) return fail(PFB200_EINVAL, "pfb200_configure: obs and theta are required");
  const int n_inj = (inj_z0 != nullptr) + (inj_u != nullptr) + (inj_z != nullptr) + (inj_u_final != nullptr);
  if (n_inj != 0 && n_inj != 4)
    return fail(PFB200_EINVAL, "pfb200_configure: injected randomness needs all of inj_z0, inj_u, inj_z, inj_u_final");
  const bool inject = n_inj == 4;
  const bool history = (flags & PFB200_FLAG_STORE_HISTORY) != 0;
  const bool states = !history && (flags & PFB200_FLAG_STORE_STATES) != 0;
  CUDA_TRY(cudaSetDevice(h->device));

  bool general = inject || resampling != PFB200_RESAMPLE_SYSTEMATIC;  // + the distributed filter (below)

  // ---- launch geometry
  const size_t real_sz = h->dtype == PFB200_F64 ? 8 : 4;
  const size_t smem_limit = 200 * 1024;  // dynamic shared memory we are willing to use per CTA
  const int R = (int)h->opt_dist_world, rank = (int)h->opt_dist_rank;
  if (R > 1) {
    if (rank < 0 || rank >= R) return fail(PFB200_EINVAL, "pfb200_configure: dist_rank %d out of range for dist_world %d", rank, R);
    if (B != 1 || history || states)
      return fail(PFB200_EINVAL, "pfb200_configure: a distributed filter takes n_filters = 1 and no STORE_HISTORY / STORE_STATES");
    general = true;  // the peer-memory paths live in the general flavour
  }
  int G = 1, G_loc = 1, per = N, n_groups = 1, tile_cap = kChunk, capacity = 0, bps = 0, n_chunks = 1;
  size_t smem = 0;
  // Two tile granularities are compiled (ITEMS = 7 / 9 parents per thread and scan sub-chunk): take
  // the alternative when it covers a CTA's parent range with fewer sub-chunks (e.g. N = 4096 per CTA
  // in the batched mode: one sub-chunk of 4608 instead of 3584 + 512).
  int items = kItems;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int chunk = kBlock * items;
    h->kernel = kernel_for_model(h->model, h->dtype, general, items);
    CUDA_TRY(cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
    G = 1; G_loc = 1; per = N; tile_cap = chunk;
    for (int pass = 0; pass < 3; ++pass) {
      // shared memory of the current guess -> co-resident CTAs -> group size
      smem = (size_t)(tile_cap + (G > 1 ? G : 1)) * sizeof(double);
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, h->kernel, kThreads, smem));
      if (bps < 1) return fail(PFB200_ECUDA, "pfb200_configure: kernel does not fit on an SM (%zu B smem)", smem);
      capacity = bps * h->num_sms;
      if (h->opt_ctas_per_filter > 0) G_loc = (int)h->opt_ctas_per_filter;
      else if (B == 1 && R == 1 && N <= kBlock * kItemsSmall) G_loc = 1;  // one CTA, no group barrier: 7.5 vs 10.2 us per step at N = 1000
      else if (B == 1) G_loc = ((N + R - 1) / R + kBlock - 1) / kBlock;
      else G_loc = (N <= 4 * kChunk) ? 1 : (N + kChunk - 1) / kChunk;
      if (G_loc > capacity) G_loc = capacity;
      {  // group size limit (exchange staging, fused reductions): 1184 = 8 GPUs x 148 CTAs
        const int g_max = 1184;
        if (G_loc * R > g_max) G_loc = g_max / R;
      }
      if (G_loc < 1) G_loc = 1;
      G = G_loc * R;
      per = (N + G - 1) / G;
      if (G > 1) per = (per + 31) & ~31;  // tile boundaries on 32-particle multiples (aligned stores on
                                          // every rank; same tiling for any split of G over GPUs)
      if (R == 1) { G = (N + per - 1) / per; G_loc = G; }  // no empty CTAs on a single GPU
      n_chunks = (per + chunk - 1) / chunk;
      const size_t want = (size_t)n_chunks * chunk;
      tile_cap = (n_chunks <= 8 && (want + G) * sizeof(double) <= smem_limit) ? (int)want : chunk;
    }
    const int alt_chunks = (per + kBlock * kItemsAlt - 1) / (kBlock * kItemsAlt);
    if (attempt == 0 && (h->opt_items == kItemsSmall || (h->opt_items == 0 && per <= kBlock * kItemsSmall))) items = kItemsSmall;
    else if (attempt == 0 && h->opt_items == 0 && alt_chunks < n_chunks) items = kItemsAlt;
    else if (attempt == 0 && h->opt_items == kItemsAlt) items = kItemsAlt;
    else break;
  }
  h->items = items;
  // Shared-memory ancestor ring: one CTA per filter, indices fit 16 bits, ring + CDF tile fit.
  const int RA_ring = L + 2;
  const int anc_stride = (N + 7) & ~7;
  const size_t anc_bytes = (size_t)RA_ring * anc_stride * sizeof(unsigned short);
  const bool anc_smem = h->opt_anc_smem && G == 1 && R == 1 && !history && do_smoother && N <= 65535 &&
                        (size_t)(tile_cap + 2) * sizeof(double) + anc_bytes <= smem_limit;
  smem = (size_t)(tile_cap + (G > 1 ? G : 1)) * sizeof(double) + (anc_smem ? anc_bytes + 8 : 0);
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, h->kernel, kThreads, smem));
  if (bps < 1) return fail(PFB200_ECUDA, "pfb200_configure: kernel does not fit on an SM (%zu B smem)", smem);
  capacity = bps * h->num_sms;
  if (G_loc > capacity)
    return fail(PFB200_ECUDA, "pfb200_configure: %d cooperating CTAs exceed the %d co-resident CTAs", G_loc, capacity);
  h->blocks_per_sm = bps;
  if (history) {
    n_groups = B;
    if (G > 1 && (long long)B * G > capacity)
      return fail(PFB200_EINVAL, "pfb200_configure: STORE_HISTORY with %d filters x %d CTAs exceeds the %d co-resident CTAs", B, G, capacity);
  } else {
    n_groups = capacity / G_loc;
    if (n_groups > B) n_groups = B;
    if (h->opt_max_groups > 0 && n_groups > h->opt_max_groups) n_groups = (int)h->opt_max_groups;
    if (n_groups < 1) n_groups = 1;
  }
  h->cooperative = G > 1;
  h->smem_bytes = smem;
  h->n_ws = history ? B : n_groups;

  // ---- small / medium filters: one thread-block cluster per filter (pf_cluster.cuh) instead of the
  //      persistent kernel.  Systematic resampling on one GPU; C CTAs of cap = 2^lg parents each.
  h->cluster = 0;
  int cl_lg = 0, cl_RAs = 0;
  const bool cl_auto_ok = h->opt_ctas_per_filter == 0 && h->opt_items == 0;  // explicit persistent-kernel geometry wins
  if ((h->opt_cluster > 0 || (h->opt_cluster == 0 && cl_auto_ok)) && R == 1 && resampling == PFB200_RESAMPLE_SYSTEMATIC &&
      N <= kClMaxCluster * kClMaxPer) {
    KernelFn ck = cluster_kernel_for_model(h->model, h->dtype, general);
    int C = 1;
    if (h->opt_cluster > 0) C = (int)h->opt_cluster;
    else {
      while (C < kClMaxCluster && (long long)C * h->opt_cluster_per < N) C *= 2;
      while (C > 1 && (long long)B * C > h->num_sms) C /= 2;  // a batch keeps every filter resident
    }
    bool ok = ck != nullptr && (h->opt_cluster > 0 || (long long)B * C <= h->num_sms);
    int cap = 0, lgc = 0, RAs = (L < T ? L : T) + kClLag;
    size_t csmem = 0;
    while (ok) {
      cap = 32; lgc = 5;
      while ((long long)cap * C < N) { cap *= 2; ++lgc; }
      csmem = h->dtype == PFB200_F64 ? ClLayout<double>(cap, RAs, C).bytes : ClLayout<float>(cap, RAs, C).bytes;
      if (cap <= kClMaxPer && csmem <= smem_limit) break;
      if (C == kClMaxCluster || h->opt_cluster > 0) ok = false;
      else C *= 2;
    }
    int max_clusters = 0;
    if (ok) {
      CUDA_TRY(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
      if (C > 8) CUDA_TRY(cudaFuncSetAttribute(ck, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)(C * (B < h->num_sms ? B : h->num_sms)));
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = csmem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, (const void*)ck, &cfg) != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        ok = false;
        if (h->opt_cluster > 0)
          return fail(PFB200_ECUDA, "pfb200_configure: a cluster of %d CTAs with %zu B of shared memory cannot be scheduled", C, csmem);
      }
    } else if (h->opt_cluster > 0) {
      return fail(PFB200_EINVAL, "pfb200_configure: cluster = %d does not fit this problem (N = %d, model %d)", C, N, h->model);
    }
    if (ok) {
      h->cluster = C;
      h->kernel = ck;
      h->cooperative = false;
      G = G_loc = C; per = cap; tile_cap = cap; cl_lg = lgc; cl_RAs = RAs;
      n_groups = history ? B : (B < max_clusters ? B : max_clusters);
      if (!history && h->opt_max_groups > 0 && n_groups > h->opt_max_groups) n_groups = (int)h->opt_max_groups;
      h->smem_bytes = csmem;
      h->n_ws = history ? B : n_groups;
      h->items = kItemsSmall;
    }
  }

  // A connected distributed handle keeps its peer mappings (and the running exchange counters) as long
  // as the geometry the mappings were made for is unchanged: an MH iteration re-configures / updates
  // parameters without touching the IPC handles.  Any other change needs pfb200_dist_disconnect on every
  // rank first (peers must not keep mappings of buffers this rank is about to re-allocate).
  const long long geom[8] = {N, T, L, G_loc, R, rank, (long long)h->dtype * 4 + (do_smoother ? 2 : 0) + (resampling == PFB200_RESAMPLE_MULTINOMIAL ? 1 : 0), items};
  bool keep_peers = false;
  if (h->dist_connected) {
    keep_peers = R > 1 && !memcmp(geom, h->dist_geom, sizeof(geom));
    if (!keep_peers)
      return fail(PFB200_ESTATE, "pfb200_configure: this distributed handle is connected to its peers; call "
                                 "pfb200_dist_disconnect on every rank (then a host barrier) before changing its geometry");
  }
  memcpy(h->dist_geom, geom, sizeof(geom));
  const DeviceProblem old = h->prob;

  DeviceProblem& p = h->prob;
  p = DeviceProblem{};
  p.N = N; p.NS = (N + 31) & ~31; p.T = T; p.L = L; p.P = P; p.B = B; p.G = G; p.n_groups = n_groups; p.per_cta = per; p.tile_cap = tile_cap;
  p.R = R; p.rank = R > 1 ? rank : 0; p.G_loc = G_loc; p.span = R > 1 ? G_loc * per : p.NS; p.epoch0 = 0; p.sjobs0 = 0;
  // distributed filter: ring rows are [halo | own span | halo] (mirrors of both neighbours' nearest
  // states and ancestors: lineage walks across a rank boundary stay in local memory)
  p.halo = R > 1 ? (p.span < (1 << 17) ? p.span : (1 << 17)) : 0;
  if (R > 1) p.NS = p.span + 2 * p.halo;
  p.RX = (history || states) ? T + 1 : L + 3;  // +1: the smoother job of lag time t-2 overlaps the writes of generation t
  p.RS = 3;
  if (h->cluster) {  // smoother jobs run up to kClLag time steps behind the filter warps
    p.cluster = h->cluster; p.lg_per = cl_lg; p.RA_s = cl_RAs; p.RS = 8;
    if (!(history || states)) p.RX = L + kClLag;
  }
  p.xs_per_filter = states ? 1 : 0;
  p.RA = history ? T + 1 : L + 2;
  p.RW = history ? T + 1 : 3;
  p.do_smoother = do_smoother ? 1 : 0;
  p.gen_init = generate_initial_state ? 1 : 0;
  p.resampling = resampling;
  p.svl_obs_model = (flags & PFB200_FLAG_SVL_OBS_MODEL) ? 1 : 0;
  p.inject = inject ? 1 : 0;
  p.inject_per_filter = (flags & PFB200_FLAG_INJECT_PER_FILTER) ? 1 : 0;
  p.obs_per_filter = (flags & PFB200_FLAG_OBS_PER_FILTER) ? 1 : 0;
  p.ws_per_filter = history ? 1 : 0;
  p.anc_smem = (anc_smem && !h->cluster) ? 1 : 0;
  p.anc_stride = anc_stride;
  p.initial_state = initial_state;
  set_seed(p, seed);
  p.spin_limit = h->opt_spin_limit;
  p.pol_keep = h->pol_keep;
  p.pol_once = h->pol_once;
  h->flags = flags;

  // ---- buffers
  const size_t T1 = (size_t)T + 1;
  const size_t n_obs_rows = p.obs_per_filter ? B : 1;
  const size_t n_injr = p.inject_per_filter ? B : 1;
  const size_t u_row = resampling != PFB200_RESAMPLE_SYSTEMATIC ? (size_t)N : 1;
  CUDA_TRY(h->obs.reserve(n_obs_rows * T1 * 8));
  CUDA_TRY(h->consts.reserve((size_t)B * kNumConsts * 8));
  CUDA_TRY(h->evals.reserve((size_t)B * 8));
  if (inject) {
    CUDA_TRY(h->inj_z0.reserve(n_injr * N * 8));
    CUDA_TRY(h->inj_u.reserve(n_injr * T1 * u_row * 8));
    CUDA_TRY(h->inj_z.reserve(n_injr * T1 * N * 8));
    CUDA_TRY(h->inj_uf.reserve(n_injr * 8));
  }
  CUDA_TRY(h->x_ring.reserve((size_t)(states ? B : h->n_ws) * p.RX * p.NS * real_sz));
  CUDA_TRY(h->logw.reserve((size_t)h->n_ws * p.RW * p.NS * real_sz));
  CUDA_TRY(h->anc.reserve((size_t)h->n_ws * p.RA * p.NS * 4));
  if (do_smoother) CUDA_TRY(h->s_ring.reserve((size_t)h->n_ws * p.RS * p.NS * real_sz));
  if (resampling == PFB200_RESAMPLE_MULTINOMIAL)  // distributed: every rank keeps a copy of the WHOLE CDF
    CUDA_TRY(h->cdf_g.reserve(R > 1 ? ((size_t)N + 32) * 8 : (size_t)h->n_ws * p.NS * 8));
  CUDA_TRY(h->xval.reserve((size_t)n_groups * 2 * G * 8));
  CUDA_TRY(h->xcount.reserve(counter_bytes(n_groups)));
  CUDA_TRY(h->abortf.reserve(16));
  CUDA_TRY(h->stat_m.reserve((size_t)B * T1 * 8));
  CUDA_TRY(h->stat_S.reserve((size_t)B * T1 * 8));
  CUDA_TRY(h->filt_part.reserve((size_t)B * T1 * G_loc * 8));
  CUDA_TRY(h->smo_part.reserve((size_t)B * T1 * G_loc * (P + 1) * 8));
  CUDA_TRY(h->traj_k.reserve((size_t)B * 4));
  CUDA_TRY(h->traj_idx.reserve((size_t)B * 4));
  CUDA_TRY(h->ll_terms.reserve((size_t)B * T1 * 8));
  CUDA_TRY(h->log_like.reserve((size_t)B * 8));
  CUDA_TRY(h->filt.reserve((size_t)B * T1 * 8));
  CUDA_TRY(h->smo.reserve((size_t)B * T1 * 8));
  CUDA_TRY(h->grad.reserve((size_t)B * P * T1 * 8));
  CUDA_TRY(h->traj.reserve((size_t)B * T1 * 8));
  CUDA_TRY(h->prof.reserve((size_t)n_groups * G_loc * kProfSlots * 8));

  p.obs = h->obs.as<double>();
  p.consts = h->consts.as<double>();
  p.eval_ids = h->evals.as<unsigned long long>();
  p.inj_z0 = h->inj_z0.as<double>();
  p.inj_u = h->inj_u.as<double>();
  p.inj_z = h->inj_z.as<double>();
  p.inj_u_final = h->inj_uf.as<double>();
  p.x_ring = h->x_ring.ptr;
  p.logw = h->logw.ptr;
  p.anc_ring = h->anc.as<int>();
  p.xval = h->xval.as<double>();
  p.xcount = h->xcount.as<unsigned>();
  p.abort_flag = h->abortf.as<int>();
  p.stat_m = h->stat_m.as<double>();
  p.stat_S = h->stat_S.as<double>();
  p.filt_part = h->filt_part.as<double>();
  p.smo_part = h->smo_part.as<double>();
  p.traj_k = h->traj_k.as<int>();
  p.cdf_g = h->cdf_g.as<double>();
  p.s_ring = h->s_ring.ptr;
  p.traj_idx = h->traj_idx.as<int>();
  p.prof = h->prof.as<long long>();
  for (int r = 0; r < 8; ++r) { p.peer_x[r] = p.x_ring; p.peer_lw[r] = p.logw; p.peer_anc[r] = p.anc_ring; p.peer_xval[r] = p.xval; p.peer_xcount[r] = p.xcount; p.peer_cdf[r] = p.cdf_g; }
  if (keep_peers) {
    if (p.x_ring != old.x_ring || p.logw != old.logw || p.anc_ring != old.anc_ring || p.xval != old.xval || p.xcount != old.xcount || p.cdf_g != old.cdf_g)
      return fail(PFB200_ESTATE, "pfb200_configure: internal error: a connected distributed handle re-allocated exported buffers");
    for (int r = 0; r < 8; ++r) { p.peer_x[r] = old.peer_x[r]; p.peer_lw[r] = old.peer_lw[r]; p.peer_anc[r] = old.peer_anc[r]; p.peer_xval[r] = old.peer_xval[r]; p.peer_xcount[r] = old.peer_xcount[r]; p.peer_cdf[r] = old.peer_cdf[r]; }
  }

  // ---- uploads
  CUDA_TRY(cudaMemcpyAsync(h->obs.ptr, obs, n_obs_rows * T1 * 8, cudaMemcpyHostToDevice, h->stream));
  if (inject) {
    CUDA_TRY(cudaMemcpyAsync(h->inj_z0.ptr, inj_z0, n_injr * N * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->inj_u.ptr, inj_u, n_injr * T1 * u_row * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->inj_z.ptr, inj_z, n_injr * T1 * N * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->inj_uf.ptr, inj_u_final, n_injr * 8, cudaMemcpyHostToDevice, h->stream));
  }
  int rc = upload_params(h, theta, seed, eval_ids);
  if (rc) return rc;
  h->configured = true;
  return PFB200_OK;
}

int pfb200_update_params(pfb200_handle* h, const double* theta, uint64_t seed, const uint64_t* eval_ids) {
  if (!h || !h->configured) return fail(PFB200_ESTATE, "pfb200_update_params: handle is not configured");
  if (!theta) return fail(PFB200_EINVAL, "pfb200_update_params: theta is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  return upload_params(h, theta, seed, eval_ids);
}

int pfb200_execute(pfb200_handle* h) {
  if (!h || !h->configured) return fail(PFB200_ESTATE, "pfb200_execute: handle is not configured");
  CUDA_TRY(cudaSetDevice(h->device));
  DeviceProblem& p = h->prob;
  const size_t T1 = (size_t)p.T + 1;
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  if (p.R > 1) {
    if (!h->dist_connected) return fail(PFB200_ESTATE, "pfb200_execute: distributed filter is not connected (pfb200_dist_connect)");
    p.epoch0 = h->dist_epoch;               // peer counters keep counting across launches
    // exchanges per launch: init + 2 per step + 2 at the end (+ 1 per step: the CDF barrier of multinomial resampling)
    h->dist_epoch += 2u * (unsigned)p.T + 3u + (p.resampling == PFB200_RESAMPLE_MULTINOMIAL ? (unsigned)p.T : 0u);
    p.sjobs0 = h->dist_sjobs;               // smoother jobs per CTA and launch: lag times L..T-1 + the tail
    if (p.do_smoother) h->dist_sjobs += (unsigned)(p.T > p.L ? p.T - p.L : 0) + 1u;
  } else {
    CUDA_TRY(cudaMemsetAsync(h->xcount.ptr, 0, counter_bytes(p.n_groups), h->stream));
  }
  CUDA_TRY(cudaMemsetAsync(h->abortf.ptr, 0, 16, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->traj_k.ptr, 0, (size_t)p.B * 4, h->stream));
  const dim3 grid(p.n_groups * p.G_loc), block(kThreads);
  if (h->opt_l2_persist && h->stream == h->own_stream && p.do_smoother && h->l2_persist_max > 0 && h->l2_window_max > 0) {
    // Tuning option (off by default, and never applied to a caller-supplied stream whose own access
    // policy window it would clobber): pin the ancestor ring -- re-read every time step for L steps by
    // the lineage walks -- in the persisting part of L2.
    const size_t anc_bytes = (size_t)h->n_ws * p.RA * p.NS * 4;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    {
      size_t win = anc_bytes < h->l2_window_max ? anc_bytes : h->l2_window_max;
      size_t set_aside = win < h->l2_persist_max ? win : h->l2_persist_max;
      CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside));
      attr.accessPolicyWindow.base_ptr = h->anc.ptr;
      attr.accessPolicyWindow.num_bytes = win;
      attr.accessPolicyWindow.hitRatio = win <= set_aside ? 1.0f : (float)set_aside / (float)win;
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    CUDA_TRY(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  }
  CUDA_TRY(cudaEventRecord(h->evk0, h->stream));
  if (h->cluster) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = h->smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)h->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, h->kernel, p));
  } else if (h->cooperative && !h->dist_same_device) {
    void* args[] = {&p};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)h->kernel, grid, block, args, h->smem_bytes, h->stream));
  } else {
    h->kernel<<<grid, block, h->smem_bytes, h->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaEventRecord(h->evk1, h->stream));
  FinalizeArgs fa{};
  fa.N = p.N; fa.T = p.T; fa.L = p.L; fa.P = p.P; fa.B = p.B; fa.G = p.G_loc; fa.do_smoother = p.do_smoother; fa.fully_adapted = is_fully_adapted(h->model) ? 1 : 0;
  fa.stat_m = p.stat_m; fa.stat_S = p.stat_S; fa.filt_part = p.filt_part; fa.smo_part = p.smo_part;
  fa.ll_terms = h->ll_terms.as<double>(); fa.log_like = h->log_like.as<double>();
  fa.filt = h->filt.as<double>(); fa.smo = h->smo.as<double>(); fa.grad = h->grad.as<double>();
  pf_finalize_kernel<<<dim3((unsigned)p.B, (unsigned)((T1 + 255) / 256)), 256, 0, h->stream>>>(fa);
  CUDA_TRY(cudaGetLastError());
  pf_loglike_kernel<<<p.B, 256, 0, h->stream>>>(fa.ll_terms, fa.log_like, (int)T1);
  CUDA_TRY(cudaGetLastError());
  if (h->flags & (PFB200_FLAG_STORE_HISTORY | PFB200_FLAG_STORE_STATES)) {  // state ring = all T+1 generations
    if (h->dtype == PFB200_F64)
      pf_trajectory_kernel<double><<<p.B, 256, 0, h->stream>>>(h->x_ring.as<double>(), p.traj_idx, h->traj.as<double>(), p.NS, (int)T1);
    else
      pf_trajectory_kernel<float><<<p.B, 256, 0, h->stream>>>(h->x_ring.as<float>(), p.traj_idx, h->traj.as<double>(), p.NS, (int)T1);
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  h->executed = true;
  h->timed = true;
  return PFB200_OK;
}

int pfb200_synchronize(pfb200_handle* h) {
  if (!h) return fail(PFB200_EINVAL, "pfb200_synchronize: NULL handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->executed) {
    int flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, h->abortf.ptr, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
      // the peer counters of a distributed filter are out of step after an abort: force a fresh
      // export / connect (dist_export resets the counters)
      if (h->prob.R > 1) dist_close(h);
      return fail(PFB200_ETIMEOUT, "pfb200: device-side abort: an inter-CTA exchange exceeded spin_limit");
    }
  }
  return PFB200_OK;
}

int pfb200_last_execute_ms(pfb200_handle* h, float* ms) {
  if (!h || !ms) return fail(PFB200_EINVAL, "pfb200_last_execute_ms: NULL argument");
  if (!h->timed) return fail(PFB200_ESTATE, "pfb200_last_execute_ms: nothing executed yet");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  CUDA_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return PFB200_OK;
}

int pfb200_last_kernel_ms(pfb200_handle* h, float* ms) {
  if (!h || !ms) return fail(PFB200_EINVAL, "pfb200_last_kernel_ms: NULL argument");
  if (!h->timed) return fail(PFB200_ESTATE, "pfb200_last_kernel_ms: nothing executed yet");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaEventSynchronize(h->evk1));
  CUDA_TRY(cudaEventElapsedTime(ms, h->evk0, h->evk1));
  return PFB200_OK;
}

static int d2h(pfb200_handle* h, void* dst, const void* src, size_t bytes) {
  if (!dst) return PFB200_OK;
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  return PFB200_OK;
}

int pfb200_fetch(pfb200_handle* h, double* log_like, double* filt, double* smo, double* grad_terms,
                 int32_t* traj_index, double* traj) {
  if (!h || !h->executed) return fail(PFB200_ESTATE, "pfb200_fetch: nothing executed yet");
  CUDA_TRY(cudaSetDevice(h->device));
  const DeviceProblem& p = h->prob;
  const size_t T1 = (size_t)p.T + 1, B = p.B;
  int rc;
  if ((rc = d2h(h, log_like, h->log_like.ptr, B * 8))) return rc;
  if ((rc = d2h(h, filt, h->filt.ptr, B * T1 * 8))) return rc;
  if (p.do_smoother) {
    if ((rc = d2h(h, smo, h->smo.ptr, B * T1 * 8))) return rc;
    if ((rc = d2h(h, grad_terms, h->grad.ptr, B * p.P * T1 * 8))) return rc;
  }
  if ((rc = d2h(h, traj_index, h->traj_idx.ptr, B * 4))) return rc;
  const bool history = (h->flags & (PFB200_FLAG_STORE_HISTORY | PFB200_FLAG_STORE_STATES)) != 0;
  if (history && (rc = d2h(h, traj, h->traj.ptr, B * T1 * 8))) return rc;
  rc = pfb200_synchronize(h);
  if (rc) return rc;
  if (!p.do_smoother) {
    if (smo) for (size_t i = 0; i < B * T1; ++i) smo[i] = 0.0;
    if (grad_terms) for (size_t i = 0; i < B * p.P * T1; ++i) grad_terms[i] = 0.0;
  }
  if (p.do_smoother && p.L == 0 && grad_terms)  // zero hops: score terms are undefined (see pfb200_configure)
    for (size_t i = 0; i < B * p.P * T1; ++i) grad_terms[i] = 0.0;
  if (!history && traj) for (size_t i = 0; i < B * T1; ++i) traj[i] = NAN;
  return PFB200_OK;
}

int pfb200_fetch_steps(pfb200_handle* h, double* ll_terms, double* max_logw, double* sum_w) {
  if (!h || !h->executed) return fail(PFB200_ESTATE, "pfb200_fetch_steps: nothing executed yet");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t n = (size_t)h->prob.B * (h->prob.T + 1) * 8;
  int rc;
  if ((rc = d2h(h, ll_terms, h->ll_terms.ptr, n))) return rc;
  if ((rc = d2h(h, max_logw, h->stat_m.ptr, n))) return rc;
  if ((rc = d2h(h, sum_w, h->stat_S.ptr, n))) return rc;
  return pfb200_synchronize(h);
}

int pfb200_fetch_history(pfb200_handle* h, double* particles, double* log_weights, int32_t* ancestors) {
  if (!h || !h->executed) return fail(PFB200_ESTATE, "pfb200_fetch_history: nothing executed yet");
  if (!(h->flags & PFB200_FLAG_STORE_HISTORY))
    return fail(PFB200_ESTATE, "pfb200_fetch_history: configure with PFB200_FLAG_STORE_HISTORY");
  CUDA_TRY(cudaSetDevice(h->device));
  const DeviceProblem& p = h->prob;
  const size_t rows = (size_t)p.B * (p.T + 1);
  const size_t n = rows * p.N;
  if (h->dtype == PFB200_F64) {
    if (particles)
      CUDA_TRY(cudaMemcpy2DAsync(particles, (size_t)p.N * 8, h->x_ring.ptr, (size_t)p.NS * 8, (size_t)p.N * 8, rows,
                                 cudaMemcpyDeviceToHost, h->stream));
    if (log_weights)
      CUDA_TRY(cudaMemcpy2DAsync(log_weights, (size_t)p.N * 8, h->logw.ptr, (size_t)p.NS * 8, (size_t)p.N * 8, rows,
                                 cudaMemcpyDeviceToHost, h->stream));
  } else {
    CUDA_TRY(h->scratch.reserve(n * 8));
    if (particles) {
      pf_to_double_kernel<float><<<1024, 256, 0, h->stream>>>(h->x_ring.as<float>(), h->scratch.as<double>(), rows, p.N, p.NS);
      CUDA_TRY(cudaMemcpyAsync(particles, h->scratch.ptr, n * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    if (log_weights) {
      pf_to_double_kernel<float><<<1024, 256, 0, h->stream>>>(h->logw.as<float>(), h->scratch.as<double>(), rows, p.N, p.NS);
      CUDA_TRY(cudaMemcpyAsync(log_weights, h->scratch.ptr, n * 8, cudaMemcpyDeviceToHost, h->stream));
    }
  }
  if (ancestors)
    CUDA_TRY(cudaMemcpy2DAsync(ancestors, (size_t)p.N * 4, h->anc.ptr, (size_t)p.NS * 4, (size_t)p.N * 4, rows,
                               cudaMemcpyDeviceToHost, h->stream));
  return pfb200_synchronize(h);
}

int pfb200_run(pfb200_handle* h, int n_filters, int no_particles, int no_obs, int fixed_lag,
               int do_smoother, int generate_initial_state, double initial_state, int resampling,
               unsigned flags, const double* obs, const double* theta, uint64_t seed,
               const uint64_t* eval_ids, const double* inj_z0, const double* inj_u,
               const double* inj_z, const double* inj_u_final, double* log_like, double* filt,
               double* smo, double* grad_terms, int32_t* traj_index, double* traj) {
  int rc = pfb200_configure(h, n_filters, no_particles, no_obs, fixed_lag, do_smoother, generate_initial_state,
                            initial_state, resampling, flags, obs, theta, seed, eval_ids, inj_z0, inj_u, inj_z,
                            inj_u_final);
  if (rc) return rc;
  if ((rc = pfb200_execute(h))) return rc;
  return pfb200_fetch(h, log_like, filt, smo, grad_terms, traj_index, traj);
}

// ---- distributed single filter: exchange of CUDA IPC handles between the per-GPU processes ----

int pfb200_dist_handle_bytes(void) { return kDistBufs * (int)sizeof(cudaIpcMemHandle_t); }

int pfb200_dist_export(pfb200_handle* h, void* out, int bytes) {
  if (!h || !out || !h->configured) return fail(PFB200_ESTATE, "pfb200_dist_export: configure first");
  if (h->prob.R < 2) return fail(PFB200_ESTATE, "pfb200_dist_export: set dist_world/dist_rank before configure");
  if (bytes < pfb200_dist_handle_bytes()) return fail(PFB200_EINVAL, "pfb200_dist_export: buffer too small");
  CUDA_TRY(cudaSetDevice(h->device));
  dist_close(h);
  cudaIpcMemHandle_t* hs = reinterpret_cast<cudaIpcMemHandle_t*>(out);
  memset(hs, 0, (size_t)pfb200_dist_handle_bytes());
  void* ptrs[kDistBufs] = {h->x_ring.ptr, h->logw.ptr, h->anc.ptr, h->xval.ptr, h->xcount.ptr,
                           h->prob.resampling == PFB200_RESAMPLE_MULTINOMIAL ? h->cdf_g.ptr : nullptr};
  for (int k = 0; k < kDistBufs; ++k)
    if (ptrs[k]) CUDA_TRY(cudaIpcGetMemHandle(&hs[k], ptrs[k]));
  // fresh counters; the caller synchronises all ranks (host barrier) between connect and execute
  CUDA_TRY(cudaMemset(h->xcount.ptr, 0, counter_bytes(h->prob.n_groups)));
  h->dist_epoch = 0;
  h->dist_sjobs = 0;
  return PFB200_OK;
}

int pfb200_dist_connect(pfb200_handle* h, const void* all_handles, int world) {
  if (!h || !all_handles || !h->configured) return fail(PFB200_ESTATE, "pfb200_dist_connect: configure first");
  DeviceProblem& p = h->prob;
  if (world != p.R) return fail(PFB200_EINVAL, "pfb200_dist_connect: world %d != configured dist_world %d", world, p.R);
  CUDA_TRY(cudaSetDevice(h->device));
  const cudaIpcMemHandle_t* hs = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
  const bool cdf = p.resampling == PFB200_RESAMPLE_MULTINOMIAL;
  void* own[kDistBufs] = {h->x_ring.ptr, h->logw.ptr, h->anc.ptr, h->xval.ptr, h->xcount.ptr, h->cdf_g.ptr};
  for (int r = 0; r < world; ++r) {
    void* ptr[kDistBufs];
    for (int k = 0; k < kDistBufs; ++k) {
      if (r == p.rank || (k == 5 && !cdf)) { ptr[k] = own[k]; continue; }
      CUDA_TRY(cudaIpcOpenMemHandle(&ptr[k], hs[r * kDistBufs + k], cudaIpcMemLazyEnablePeerAccess));
      h->peer_opened[r][k] = ptr[k];
    }
    p.peer_x[r] = ptr[0]; p.peer_lw[r] = ptr[1]; p.peer_anc[r] = reinterpret_cast<int*>(ptr[2]);
    p.peer_xval[r] = reinterpret_cast<double*>(ptr[3]); p.peer_xcount[r] = reinterpret_cast<unsigned*>(ptr[4]);
    p.peer_cdf[r] = reinterpret_cast<double*>(ptr[5]);
  }
  h->dist_connected = true;
  return PFB200_OK;
}

int pfb200_dist_connect_local(pfb200_handle* const* handles, int world) {
  if (!handles || world < 2 || world > 8) return fail(PFB200_EINVAL, "pfb200_dist_connect_local: need 2..8 handles");
  for (int r = 0; r < world; ++r) {
    pfb200_handle* h = handles[r];
    if (!h || !h->configured) return fail(PFB200_ESTATE, "pfb200_dist_connect_local: configure every handle first");
    if (h->prob.R != world || h->prob.rank != r)
      return fail(PFB200_EINVAL, "pfb200_dist_connect_local: handle %d is configured as rank %d of %d", r, h->prob.rank, h->prob.R);
    if (memcmp(h->dist_geom, handles[0]->dist_geom, 5 * sizeof(long long)) || h->dtype != handles[0]->dtype)
      return fail(PFB200_EINVAL, "pfb200_dist_connect_local: handle %d has a different geometry", r);
  }
  bool same_device = true;
  for (int r = 1; r < world; ++r) same_device = same_device && handles[r]->device == handles[0]->device;
  if (same_device && (long long)world * handles[0]->prob.G_loc > (long long)handles[0]->blocks_per_sm * handles[0]->num_sms)
    return fail(PFB200_EINVAL, "pfb200_dist_connect_local: %d ranks x %d CTAs on one device are not co-resident", world, handles[0]->prob.G_loc);
  for (int r = 0; r < world; ++r) {
    pfb200_handle* h = handles[r];
    CUDA_TRY(cudaSetDevice(h->device));
    dist_close(h);
    for (int q = 0; q < world; ++q) {
      pfb200_handle* o = handles[q];
      if (o->device != h->device) {
        int can = 0;
        CUDA_TRY(cudaDeviceCanAccessPeer(&can, h->device, o->device));
        if (!can) return fail(PFB200_ECUDA, "pfb200_dist_connect_local: device %d cannot access device %d", h->device, o->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
        (void)cudaGetLastError();
      }
      DeviceProblem& p = h->prob;
      p.peer_x[q] = o->x_ring.ptr; p.peer_lw[q] = o->logw.ptr; p.peer_anc[q] = o->anc.as<int>();
      p.peer_xval[q] = o->xval.as<double>(); p.peer_xcount[q] = o->xcount.as<unsigned>(); p.peer_cdf[q] = o->cdf_g.as<double>();
    }
    CUDA_TRY(cudaMemset(h->xcount.ptr, 0, counter_bytes(h->prob.n_groups)));
    h->dist_epoch = 0;
    h->dist_sjobs = 0;
  }
  for (int r = 0; r < world; ++r) {
    CUDA_TRY(cudaSetDevice(handles[r]->device));
    if (same_device) {
      // Ranks on ONE device are kernels that wait for each other, launched one after the other by this thread:
      // nothing pfb200_execute enqueues behind the first rank's kernel may still have to be LOADED (CUDA loads
      // kernels lazily at their first launch, and that load can wait for the running -- spinning -- kernel, so
      // the second rank would never be launched).  Touch everything execute uses before the first launch.
      cudaFuncAttributes fa;
      CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)handles[r]->kernel));
      CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)pf_finalize_kernel));
      CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)pf_loglike_kernel));
      CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)pf_trajectory_kernel<double>));
      CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)pf_trajectory_kernel<float>));
      CUDA_TRY(cudaMemsetAsync(handles[r]->abortf.ptr, 0, 16, handles[r]->stream));
    }
    CUDA_TRY(cudaDeviceSynchronize());
    handles[r]->dist_connected = true;
    handles[r]->dist_same_device = same_device;
  }
  return PFB200_OK;
}

int pfb200_dist_disconnect(pfb200_handle* h) {
  if (!h) return fail(PFB200_EINVAL, "pfb200_dist_disconnect: NULL handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  dist_close(h);
  return PFB200_OK;
}

int pfb200_dist_traj_count(pfb200_handle* h, int32_t* count) {
  if (!h || !count || !h->executed) return fail(PFB200_ESTATE, "pfb200_dist_traj_count: nothing executed yet");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(count, h->traj_k.ptr, sizeof(int32_t), cudaMemcpyDeviceToHost));
  return PFB200_OK;
}

int pfb200_dist_traj_lookup(pfb200_handle* h, int32_t k, int32_t* index) {
  if (!h || !index || !h->executed) return fail(PFB200_ESTATE, "pfb200_dist_traj_lookup: nothing executed yet");
  const DeviceProblem& p = h->prob;
  *index = -1;
  if (k < 0) return fail(PFB200_EINVAL, "pfb200_dist_traj_lookup: k < 0");
  if (k > p.N - 1) k = p.N - 1;
  const long long lo = (long long)p.rank * p.span;
  if (k < lo || k >= lo + p.span) return PFB200_OK;  // another rank owns particle k
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  // generation T lives in slot T mod RA of the ancestor ring; rows are [halo | own span | halo]
  const size_t off = (size_t)(p.T % p.RA) * p.NS + p.halo + (size_t)(k - lo);
  CUDA_TRY(cudaMemcpy(index, h->anc.as<int>() + off, sizeof(int32_t), cudaMemcpyDeviceToHost));
  return PFB200_OK;
}

int pfb200_debug_profile(pfb200_handle* h, long long* out, int max_ctas) {
  if (!h || !out || !h->executed) return fail(PFB200_ESTATE, "pfb200_debug_profile: nothing executed yet");
  CUDA_TRY(cudaSetDevice(h->device));
  int n = h->prob.n_groups * h->prob.G_loc;
  if (n > max_ctas) n = max_ctas;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(out, h->prof.ptr, (size_t)n * kProfSlots * 8, cudaMemcpyDeviceToHost));
  return n;
}

int pfb200_philox_normals(pfb200_handle* h, uint64_t seed, uint64_t eval_id, int t, int n, double* z) {
  if (!h || !z || n < 1 || t < 0) return fail(PFB200_EINVAL, "pfb200_philox_normals: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->scratch.reserve((size_t)n * 8));
  const int pairs = (n + 1) / 2;
  pf_philox_normals_kernel<<<(pairs + 255) / 256, 256, 0, h->stream>>>(seed, eval_id, t, n, h->scratch.as<double>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(z, h->scratch.ptr, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PFB200_OK;
}

int pfb200_philox_uniforms(pfb200_handle* h, uint64_t seed, uint64_t eval_id, int t, int count, double* u) {
  if (!h || !u || count < 1) return fail(PFB200_EINVAL, "pfb200_philox_uniforms: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->scratch.reserve((size_t)count * 8));
  pf_philox_uniforms_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(seed, eval_id, t, count, h->scratch.as<double>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(u, h->scratch.ptr, (size_t)count * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PFB200_OK;
}

}  // extern "C"
