/*
 * pfb200.h -- C ABI of the B200-native (sm_100a) sequential Monte Carlo state estimator.
 *
 * Drop-in boundary for the hot path of compops/qnmh-sysid2018: the bootstrap particle filter and
 * fixed-lag particle smoother of python/state/particle_methods.  The entry points below are what
 * the reference's native boundary for this path binds (citations are relative to the reference
 * repository root):
 *
 *   reference interface                                               replaced by
 *   ----------------------------------------------------------------  ---------------------------
 *   bpf_lgss(obs, mu, phi, sigmav, sigmae) -> (filt, ll, traj)         pfb200_run(do_smoother=0)
 *     python/state/particle_methods/cython_lgss_helper.pyx:35,140
 *   flps_lgss(obs, mu, phi, sigmav, sigmae)                            pfb200_run(do_smoother=1)
 *     -> (filt, smo, ll, gradient[4][NoObs], traj)
 *     python/state/particle_methods/cython_lgss_helper.pyx:144,335
 *   bpf_sv / flps_sv (obs, mu, phi, sigmav)                            pfb200_run(model=PFB200_MODEL_SV)
 *     python/state/particle_methods/cython_sv_helper.pyx:35,144
 *   bpf_sv / flps_sv (obs, mu, phi, sigmav, rho)                       pfb200_run(model=PFB200_MODEL_SVL)
 *     python/state/particle_methods/cython_sv_leverage_helper.pyx:35,146
 *   ParticleMethods.filter / .smoother (NumPy twin, run-time N/T/L)    same calls; semantics follow
 *     python/state/particle_methods/standard.py:42,121                 the NumPy path
 *   systematic / stratified / multinomial resampling                   `resampling` argument
 *     python/state/particle_methods/resampling.pyx:33,44,65
 *
 * Conventions: every pointer is a HOST pointer owned by the caller unless stated otherwise; arrays
 * are C-contiguous doubles in time-major order; the library owns all device memory inside the
 * handle.  No exceptions cross the boundary: every call returns 0 on success or a negative
 * PFB200_E* code, and pfb200_last_error() returns a human readable message for the calling
 * thread's most recent failure.  A handle is bound to one CUDA device and one stream and is not
 * thread-safe; different handles may be used from different threads.  There is no CPU fallback:
 * without a usable CUDA device pfb200_create fails with PFB200_ENODEVICE.
 */
#ifndef PFB200_H_
#define PFB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFB200_ABI_VERSION 1

/* model_id: which per-particle model arithmetic runs on the device */
#define PFB200_MODEL_LGSS 0 /* models/linear_gaussian_model.py:            theta = mu, phi, sigma_v, sigma_e */
#define PFB200_MODEL_SV 1   /* models/stochastic_volatility_model.py:      theta = mu, phi, sigma_v          */
#define PFB200_MODEL_SVL 2  /* models/stochastic_volatility_model_leverage.py: theta = mu, phi, sigma_v, rho */
#define PFB200_MODEL_LGSS_FA 3 /* LGSS with the FULLY ADAPTED (optimal proposal) auxiliary particle filter: not in the
                                  reference, named by the north star; same theta as LGSS.  filt = equally
                                  weighted means, log_like = sum_t log mean_n p(y_t | x_{t-1}[n]); the smoother weights
                                  of lag time tau are the look-ahead weights p(y_{tau+1} | x_tau) */

/* dtype: storage / per-particle arithmetic type (CDF, reductions and outputs are always fp64) */
#define PFB200_F64 0
#define PFB200_F32 1

/* resampling (resampling.pyx) */
#define PFB200_RESAMPLE_SYSTEMATIC 0
#define PFB200_RESAMPLE_STRATIFIED 1
#define PFB200_RESAMPLE_MULTINOMIAL 2 /* resampling.pyx:33-42; per-particle uniforms like stratified */

/* flags (bit mask) */
#define PFB200_FLAG_STORE_HISTORY 1u    /* keep all T+1 generations (particles, log-weights, ancestors) */
#define PFB200_FLAG_OBS_PER_FILTER 2u   /* obs is [B][T+1] instead of one shared [T+1] series */
#define PFB200_FLAG_SVL_OBS_MODEL 4u    /* SV-leverage: propagate with obs[t-1] (model-consistent, Cython twin)
                                           instead of obs[t] (NumPy twin) */
#define PFB200_FLAG_INJECT_PER_FILTER 8u /* injected randomness is given per filter ([B][...]) */
#define PFB200_FLAG_STORE_STATES 16u    /* keep all T+1 generations of the PARTICLES only (the other rings stay
                                           L+2 deep): enough for `traj` = particles[traj_index, :] (standard.py:104-106)
                                           at 1/3 of the memory of STORE_HISTORY and without its slower lineage walks */

/* error codes */
#define PFB200_OK 0
#define PFB200_EINVAL -1
#define PFB200_ENODEVICE -2
#define PFB200_ECUDA -3
#define PFB200_ENOMEM -4
#define PFB200_ESTATE -5
#define PFB200_ETIMEOUT -6

typedef struct pfb200_handle pfb200_handle;

/* Number of parameters P of a model (4, 3, 4); negative on an unknown model. */
int pfb200_model_no_params(int model_id);
int pfb200_abi_version(void);
const char* pfb200_last_error(void);

/* Creates an estimator bound to CUDA device `device` (-1: current device). */
int pfb200_create(int device, int model_id, int dtype, pfb200_handle** out);
int pfb200_destroy(pfb200_handle* h);

/* Uses `cuda_stream` (a cudaStream_t, e.g. torch's current stream) for all work; NULL restores
 * the handle's own stream. */
int pfb200_set_stream(pfb200_handle* h, void* cuda_stream);

/* Tuning knobs (defaults are chosen from the problem size):
 *   "ctas_per_filter"  G: CTAs cooperating on one filter (0 = automatic)
 *   "max_groups"       upper bound on concurrently resident filter groups (0 = automatic)
 *   "spin_limit"       abort a launch whose inter-CTA exchange spins longer than this many polls
 *   "items"            parents per thread and scan sub-chunk: 7, 9 or 2 (0 = automatic)
 *   "anc_smem"         1 (default): filters of one CTA and N <= 65535 keep the ancestor ring in
 *                      shared memory; 0: always in global memory
 *   "cluster"          which kernel runs filters of up to 16 384 particles with systematic resampling on one
 *                      GPU: 0 (default) = automatic -- the cluster kernel (one filter per thread-block
 *                      cluster of C CTAs, parents / ancestors in distributed shared memory, one all-gather
 *                      per time step; csrc/pf_cluster.cuh) whenever all B filters are resident at once
 *                      (B * C <= SM count), else the persistent kernel; -1 = never; C = 1, 2, 4, 8, 16 =
 *                      always, with C CTAs per filter (configure fails if the problem does not fit).  An
 *                      explicit "ctas_per_filter" or "items" selects the persistent kernel.
 *   "cluster_per"      parents per CTA the automatic cluster size aims at (default 128)
 *   "dist_world", "dist_rank"  shard ONE filter over several GPUs (see pfb200_dist_* below)
 * pfb200_get_info keys: "ctas_per_filter", "n_groups", "per_cta", "cluster" (CTAs per cluster, 0 = persistent
 * kernel), "items", "anc_smem", "tile_cap", "smem_bytes", "cooperative", "block_threads", "filter_threads",
 * "kernel_launches_per_execute", "workspace_bytes", "num_sms", ... */
int pfb200_set_option(pfb200_handle* h, const char* name, long long value);
int pfb200_get_info(pfb200_handle* h, const char* name, long long* value);

/*
 * Describes the next evaluation(s) and uploads the inputs to the device (H2D on the handle's
 * stream).  B = n_filters independent evaluations are run side by side (MH proposals, Monte Carlo
 * replicates).
 *
 *   obs        [T+1] (or [B][T+1] with PFB200_FLAG_OBS_PER_FILTER); obs[0] is the dummy initial row
 *   theta      [B][P] model parameters in the reference's order (get_all_params())
 *   seed, eval_ids[B] (NULL -> 0..B-1): Philox4x32-10 stream = (seed, eval_id); a draw depends only
 *              on (seed, eval_id, time step, particle index), never on the launch geometry
 *   inj_z0 [N], inj_u [T+1] ([T+1][N] for stratified / multinomial), inj_z [T+1][N], inj_u_final [1]:
 *              injected randomness replacing Philox (all four or none; NULL = Philox); with
 *              PFB200_FLAG_INJECT_PER_FILTER each carries a leading [B] dimension.  Row t of
 *              inj_u / inj_z is consumed at time step t (row 0 is unused).
 */
int pfb200_configure(pfb200_handle* h, int n_filters, int no_particles, int no_obs, int fixed_lag,
                     int do_smoother, int generate_initial_state, double initial_state,
                     int resampling, unsigned flags, const double* obs, const double* theta,
                     uint64_t seed, const uint64_t* eval_ids, const double* inj_z0,
                     const double* inj_u, const double* inj_z, const double* inj_u_final);

/* Re-uploads only theta / eval_ids for the configured problem (what an MH iteration changes). */
int pfb200_update_params(pfb200_handle* h, const double* theta, uint64_t seed,
                         const uint64_t* eval_ids);

/* Enqueues the filter (+ smoother) kernels on the handle's stream.  Asynchronous. */
int pfb200_execute(pfb200_handle* h);

/* Waits for the stream; reports a device-side abort (exchange time-out) as PFB200_ETIMEOUT. */
int pfb200_synchronize(pfb200_handle* h);

/* Device time of the most recent pfb200_execute in ms (CUDA events on the handle's stream). */
int pfb200_last_execute_ms(pfb200_handle* h, float* ms);
/* Device time of the persistent filter/smoother kernel alone within that execute. */
int pfb200_last_kernel_ms(pfb200_handle* h, float* ms);

/*
 * Copies results to host memory (D2H on the stream, then waits).  Any pointer may be NULL.
 *   log_like   [B]            sum_t ( m_t + log S_t - log N )           standard.py:96-98,110
 *   filt       [B][T+1]       filtered means (entry 0 is 0)             standard.py:101
 *   smo        [B][T+1]       fixed-lag smoothed means                  standard.py:140-157
 *   grad_terms [B][P][T+1]    G[p][i] = sum_n w_lag[n] d_p log p(x_{i+1}, y_i | x_i)   standard.py:161-169
 *   traj_index [B]            row index of the trajectory draw          standard.py:104-105
 *   traj       [B][T+1]       particles[traj_index, :]; needs PFB200_FLAG_STORE_STATES or _HISTORY (else NaN)
 * fixed_lag = 0 with do_smoother: smo == filt (zero hops) and grad_terms is returned as zeros (the reference
 * raises NameError at standard.py:161 when it is asked for score terms with fixed_lag = 0).
 */
int pfb200_fetch(pfb200_handle* h, double* log_like, double* filt, double* smo, double* grad_terms,
                 int32_t* traj_index, double* traj);

/* Per-step diagnostics: log-likelihood terms [B][T+1], max log-weight m_t and sum S_t. */
int pfb200_fetch_steps(pfb200_handle* h, double* ll_terms, double* max_logw, double* sum_w);

/* Full history (PFB200_FLAG_STORE_HISTORY only), time-major: particles/log_weights [B][T+1][N]
 * (converted to double), ancestors [B][T+1][N] int32. */
int pfb200_fetch_history(pfb200_handle* h, double* particles, double* log_weights,
                         int32_t* ancestors);

/* Convenience: configure + execute + fetch in one call with host buffers (the end-to-end path
 * the reference's flps_* / bpf_* calls map to). */
int pfb200_run(pfb200_handle* h, int n_filters, int no_particles, int no_obs, int fixed_lag,
               int do_smoother, int generate_initial_state, double initial_state, int resampling,
               unsigned flags, const double* obs, const double* theta, uint64_t seed,
               const uint64_t* eval_ids, const double* inj_z0, const double* inj_u,
               const double* inj_z, const double* inj_u_final, double* log_like, double* filt,
               double* smo, double* grad_terms, int32_t* traj_index, double* traj);

/* Writes the standard normals the Philox path uses at time step t (0 = initial state) for
 * particles [0, N) of evaluation eval_id into host array z[N] (for replaying a Philox run through
 * a CPU checker). */
int pfb200_philox_normals(pfb200_handle* h, uint64_t seed, uint64_t eval_id, int t, int n,
                          double* z);
/* The resampling uniform(s) the Philox path uses at time step t (count = 1 systematic, N stratified /
 * multinomial; t = -1: the trajectory-draw uniform). */
int pfb200_philox_uniforms(pfb200_handle* h, uint64_t seed, uint64_t eval_id, int t, int count,
                           double* u);

/*
 * One filter sharded over the GPUs of a box (one process per GPU).  Rank r holds the particles with
 * global indices [r*span, (r+1)*span); children are written to the owner of their index range with
 * NVLink peer stores, the per-step exchange (tile totals, max log-weight) is an all-gather through
 * peer memory inside the persistent kernel.  Protocol, identical on every rank:
 *   pfb200_set_option(h, "dist_world", R); pfb200_set_option(h, "dist_rank", r);
 *   pfb200_configure(h, 1, N_global, ...)            -- same arguments on all ranks
 *   pfb200_dist_export(h, mine, bytes)               -- bytes = pfb200_dist_handle_bytes()
 *   (all-gather the R handle blobs between the processes, e.g. torch.distributed)
 *   pfb200_dist_connect(h, all, R); (host barrier); pfb200_execute(h) on all ranks; pfb200_fetch
 * filt / smo / grad_terms of a rank are its PARTIAL sums (add them over ranks); log_like is complete
 * and identical on every rank; traj_index is -1 (see pfb200_dist_traj_* below).
 *
 * A connected handle stays connected across pfb200_update_params and across pfb200_configure calls that
 * keep the geometry (N, T, L, CTAs per rank, world, dtype, smoother, resampling): an MH iteration only
 * re-uploads theta.  Any other re-configuration needs pfb200_dist_disconnect on EVERY rank (then a host
 * barrier) first -- a rank must not re-allocate buffers its peers still map.  A device-side abort
 * (PFB200_ETIMEOUT) disconnects the handle: export / connect again.
 */
int pfb200_dist_handle_bytes(void);
int pfb200_dist_export(pfb200_handle* h, void* handles_out, int bytes);
int pfb200_dist_connect(pfb200_handle* h, const void* all_handles, int world);
int pfb200_dist_disconnect(pfb200_handle* h);
/* The same for ranks that live in ONE process (one handle per GPU, peer access instead of IPC; also
 * several ranks on one device for single-GPU tests of the peer paths): handles[r] must be configured as
 * rank r of `world`.  Then pfb200_execute(handles[0..world)) in any order, pfb200_fetch on each. */
int pfb200_dist_connect_local(pfb200_handle* const* handles, int world);
/* Trajectory draw of a distributed filter (standard.py:104-105): `count` = this rank's part of
 * k = #{ j : cdf[j] <= u_final }; sum the counts over ranks, then the rank that owns particle
 * min(k, N-1) returns traj_index = a_T[k] from pfb200_dist_traj_lookup (every other rank: -1). */
int pfb200_dist_traj_count(pfb200_handle* h, int32_t* count);
int pfb200_dist_traj_lookup(pfb200_handle* h, int32_t k, int32_t* index);

/* Debug builds (-DPFB_PROFILE) only: per-CTA cycle counters of 12 phases of a time step, summed
 * over the launch; out[cta][12].  Returns the number of CTAs written (or a negative error). */
int pfb200_debug_profile(pfb200_handle* h, long long* out, int max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* PFB200_H_ */
