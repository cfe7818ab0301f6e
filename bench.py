#!/usr/bin/env python
"""Benchmark of the B200 particle filter + fixed-lag smoother (BASELINE.json metric:
particle-timesteps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--workload c2|c3|c4]
    python bench.py --impl reference ...        # CPU arm: the oracle port on the host cores

A "step" is one pass of the hot path over one batch of synthetic input: one complete evaluation
(bootstrap PF + fixed-lag smoother over all T time steps) of the workload's filter(s).
Default workload (N GPUs = 1..8): BASELINE config C2 -- one LGSS filter, N = 2^20 particles,
T = 1000 -- per GPU; with --gpus N every rank runs an independent replicate of it (Monte Carlo
replicates shard trivially, no data-path collective: weak scaling).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM, CUDA events,
max over ranks); `e2e` goes through the host-buffer C ABI call per step (H2D of the inputs from
pinned memory + D2H of the results inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

THETA_LGSS = np.array([0.2, 0.5, 1.0, 0.5])  # helper_linear_gaussian.py:37-40
THETA_SV = np.array([0.5, 0.95, 0.25])
FIXED_LAG = 10  # example2_lgss_particles.py:30-37
SEED = 87655678

WORKLOADS = {
    # name: (model, B filters per GPU, N, T)
    "c2": ("lgss", 1, 1 << 20, 1000),
    "c3": ("lgss", 4096, 4096, 500),
    "c4": ("sv", 1, 1 << 16, 1000),
    "c1": ("lgss", 1, 1000, 500),
    # one filter sharded over ALL ranks (peer-memory exchange inside the kernel); B/N/T: whole job
    "c5": ("lgss", 1, 1 << 24, 100),
}


def synthetic_obs(model, T, seed=12345):
    """S-LGSS / S-SV data by the model recursion (helpers/data_handling.py:107-113)."""
    rs = np.random.RandomState(seed)
    th = THETA_LGSS if model == "lgss" else THETA_SV
    x = 0.0 if model == "lgss" else th[0]
    obs = np.zeros(T + 1)
    for t in range(1, T + 1):
        x = th[0] + th[1] * (x - th[0]) + th[2] * rs.randn()
        obs[t] = x + th[3] * rs.randn() if model == "lgss" else np.exp(0.5 * x) * rs.randn()
    return obs


def algorithmic_bytes_per_particle_step(dtype, L, smoother=True):
    """filter 6s+4, smoother increment 4L+3s  ->  9s + 4 + 4L."""
    s = 8 if dtype == "f64" else 4
    return 6 * s + 4 + ((4 * L + 3 * s) if smoother else 0)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region: NVML in-process every 20 ms
    (nvidia-smi every 200 ms as fallback)."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
            ("sw_power_cap", 0x4))

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []  # (sm_mhz, set of active reasons)
        self.sm_max = None
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            if uuid:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, handle)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        pynvml, handle = self.nvml
        mhz = float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
        try:
            mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle))
        except Exception:
            mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
        self.samples.append((mhz, {n for n, b in self.BITS if mask & b}))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        if out:
            f = [x.strip() for x in out.split(",")]
            self.sm_max = float(f[1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.samples.append((float(f[0]), {n for i, n in enumerate(names) if f[2 + i] == "Active"}))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted(set().union(*[s[1] for s in self.samples]))
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (NumPy restatement of the reference) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_one(args):
    model, N, T_s, L, seed = args
    from oracle import pf_oracle as orc
    mid = {"lgss": orc.MODEL_LGSS, "sv": orc.MODEL_SV}[model]
    theta = THETA_LGSS if model == "lgss" else THETA_SV
    obs = synthetic_obs(model, T_s)
    draws = orc.random_draws(np.random.default_rng(seed), N, T_s)
    t0 = time.perf_counter()
    out = orc.smoother(mid, theta, obs, N, L, draws)
    return time.perf_counter() - t0, out["log_like"]


def cpu_sample_size(model, B, N, T):
    """Bounded sample of the workload: same N, fewer time steps / filters (10-30 s of CPU work)."""
    budget = 2.6e7  # particle-timesteps one NumPy process gets through in roughly 10 s
    N_s = min(N, 1 << 20)  # the NumPy port keeps (T+1, N) arrays: cap the sample's particle count
    T_s = int(max(2 * FIXED_LAG + 5, min(T, budget // N_s)))
    return N_s, T_s


def run_cpu_baseline(model, B, N, T, procs=1, reps=1):
    """Times the oracle port.  procs > 1: independent filters (replicates) on several processes."""
    N_s, T_s = cpu_sample_size(model, B, N, T)
    jobs = [(model, N_s, T_s, FIXED_LAG, 1000 + i) for i in range(procs * reps)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_one(j) for j in jobs]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_one, jobs)
    wall = time.perf_counter() - t0
    units = len(jobs) * N_s * T_s
    value = units / wall if procs > 1 else units / sum(r[0] for r in res)
    sample = ("NumPy oracle port (pf_oracle.smoother, O(N T L), no O(N T^2) ancestry copy), %s N=%d, "
              "first %d of T=%d time steps, L=%d, %d evaluation(s) on %d process(es)"
              % (model, N_s, T_s, T, FIXED_LAG, len(jobs), procs))
    return {"value": value, "unit": "particle-timesteps/s", "cores": procs, "kind": "port",
            "sample": sample, "seconds": wall}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, B, N, T = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 16))
    times = []
    last = None
    for i in range(args.warmup + args.steps):
        last = run_cpu_baseline(model, B, N, T, procs=procs)
        if i >= args.warmup:
            times.append(last)
    value = float(np.mean([t["value"] for t in times]))
    ms = float(np.mean([t["seconds"] for t in times])) * 1e3
    line = {
        "impl": "reference", "metric": "particle-timesteps/sec (PF+fixed-lag smoother)", "value": value,
        "unit": "particle-timesteps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, model, B, N, T),
        "cpu_baseline": {"value": value, "unit": "particle-timesteps/s", "cores": procs, "kind": "port",
                         "sample": last["sample"] + " per step; host has %d cores" % cores},
        "e2e": {"value": value, "unit": "particle-timesteps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, model, B, N, T):
    if getattr(args, "particles_log2", 0):
        N = 1 << args.particles_log2
    if args.workload == "c5":
        return {"workload": "C5: ONE %s filter sharded over all GPUs, N=%d particles x T=%d, L=%d, PF + fixed-lag "
                            "smoother; children pushed to the owner rank over NVLink peer memory, per-step "
                            "all-gather of CDF tile totals and max log-weight inside the persistent kernel"
                            % (model.upper(), N, T, FIXED_LAG),
                "no_particles": N, "no_obs": T, "fixed_lag": FIXED_LAG,
                "parallelism": "particles sharded by index range over ranks",
                "l2": "no flush: rings (%.0f MB per GPU at 1 GPU) exceed L2" % (N * 176 / 1e6)}
    return {"workload": "%s: %s, %d filter(s)/GPU x N=%d particles x T=%d, L=%d, systematic resampling, "
                        "PF + fixed-lag smoother (ll, filtered/smoothed means, score terms)"
                        % (args.workload.upper(), model.upper(), B, N, T, FIXED_LAG),
            "filters_per_gpu": B, "no_particles": N, "no_obs": T, "fixed_lag": FIXED_LAG,
            "parallelism": "independent replicates per GPU (no collective)",
            "l2": "no flush between steps: each step streams ~%.0f GB through rings larger than L2"
                  % (N * T * B * 92 / 1e9)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from qnmh_sysid2018_b200.engine import Engine

    model, B, N, T = WORKLOADS[args.workload]
    dtype = {"f64": "float64", "f32": "float32"}[args.dtype]
    theta0 = THETA_LGSS if model == "lgss" else THETA_SV
    obs = synthetic_obs(model, T)
    rs = np.random.RandomState(777)
    thetas = np.tile(theta0, (B, 1))
    if B > 1:  # MH-proposal-like perturbations around the true parameters
        thetas = thetas + 0.02 * rs.randn(B, len(theta0))
        thetas[:, 1] = np.clip(thetas[:, 1], -0.95, 0.95)
        thetas[:, 2:] = np.abs(thetas[:, 2:])
    eval_ids = np.arange(B, dtype=np.uint64) + np.uint64(rank * B)
    B_total = B
    if args.workload == "c3" and world > 1:
        # the 4096 evaluations are one fixed batch sharded over the ranks (strong scaling); the
        # Philox stream of an evaluation is keyed by its global index
        from qnmh_sysid2018_b200.batch import shard_range
        lo, hi = shard_range(B, rank, world)
        thetas, eval_ids, B = thetas[lo:hi], np.arange(lo, hi, dtype=np.uint64), hi - lo

    stream = torch.cuda.Stream()  # a real (non-default) stream: the library and the events share it
    torch.cuda.set_stream(stream)
    sharded = args.workload == "c5" and world > 1  # ONE filter over all ranks (strong scaling)
    if args.particles_log2:
        N = 1 << args.particles_log2
    if sharded:
        from qnmh_sysid2018_b200.distributed import DistributedParticleFilter
        dpf = DistributedParticleFilter(model, dtype, local_rank)
        eng = dpf.engine
        eng.set_stream(stream.cuda_stream)
        dpf.prepare(obs, theta0, N, FIXED_LAG, seed=SEED, eval_id=0)
    else:
        eng = Engine(model, dtype, local_rank)
        eng.set_stream(stream.cuda_stream)
        eng.configure(obs, thetas, N, FIXED_LAG, True, True, 0.0, "systematic", seed=SEED, eval_ids=eval_ids)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(args.warmup):
        eng.execute()
    barrier()
    try:
        gpu_uuid = torch.cuda.get_device_properties(local_rank).uuid
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        eng.execute()
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    kernel_ms.append(eng.last_kernel_ms())
    clocks = sampler.summary()
    eng.synchronize()
    ll = eng.fetch_log_like()

    # kernel-only duration, averaged over a few separately timed launches
    for _ in range(min(4, args.steps)):
        eng.execute()
        kernel_ms.append(eng.last_kernel_ms())
    kernel_ms = float(np.mean(kernel_ms))

    # ---- end to end through the host-buffer call
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    h_obs, h_theta, h_ev = pin(obs), pin(thetas), pin(eval_ids)
    P = eng.no_params
    outs = {"log_like": pin(np.empty(B)), "filt": pin(np.empty((B, T + 1))), "smo": pin(np.empty((B, T + 1))),
            "grad_terms": pin(np.empty((B, P, T + 1))), "traj_index": pin(np.empty(B, dtype=np.int32)),
            "traj": np.empty((B, T + 1))}

    def e2e_step():
        if sharded:  # inputs are re-uploaded and the peer memory re-connected on every call
            dpf.prepare(h_obs, theta0, N, FIXED_LAG, seed=SEED, eval_id=0)
            dpf.execute()
            return dpf.fetch()
        eng.configure(h_obs, h_theta, N, FIXED_LAG, True, True, 0.0, "systematic", seed=SEED, eval_ids=h_ev)
        eng.execute()
        return eng.fetch(outs)

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = h_obs.nbytes + B * 16 * 8 + h_ev.nbytes  # obs + derived per-filter constants + Philox stream ids
    d2h = B * 8 + 2 * B * (T + 1) * 8 + B * P * (T + 1) * 8 + B * 4

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])

    fixed_job = sharded or (args.workload == "c3" and world > 1)
    units_per_step = B_total * N * T if fixed_job else world * B * N * T
    value = units_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = units_per_step * args.steps / e2e_s

    if rank == 0:
        bytes_unit = algorithmic_bytes_per_particle_step(args.dtype, FIXED_LAG)
        peak, peak_src = measured_peak_gbs()
        per_gpu_units = B * N * T / (world if sharded else 1)
        achieved = bytes_unit * per_gpu_units / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as fh:
                    traffic = json.load(fh).get("%s_%s" % (args.workload, args.dtype))
            except Exception:
                traffic = None
        cpu = run_cpu_baseline(model, B, N, T, procs=1) if world == 1 and not args.no_cpu else None
        line = {
            "metric": "particle-timesteps/sec (PF+fixed-lag smoother)", "value": value,
            "unit": "particle-timesteps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if fixed_job else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, model, B_total, N, T),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-timesteps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": eng.info("kernel_launches_per_execute") * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "pf_persistent_kernel", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_particle_step": bytes_unit,
                         "algorithmic_bytes_per_launch": bytes_unit * per_gpu_units,
                         "peak_source": peak_src},
            "geometry": {"ctas_per_filter": eng.info("ctas_per_filter"), "groups": eng.info("n_groups"),
                         "block_threads": eng.info("block_threads"), "filter_threads": eng.info("filter_threads"),
                         "cooperative": eng.info("cooperative")},
            "log_like_sample": float(ll[0]),
        }
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--particles-log2", type=int, default=0, help="override N = 2^k (c5 sweep)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
