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
def measure(workload, dtype_key, steps, warmup, ctx, particles_log2=0, sample_clocks=True):
    """One workload on the ranks of this job: device-resident throughput (CUDA events, max over ranks),
    kernel-only duration, and the end-to-end figure through the host-buffer call.  Returns a dict on
    every rank (the reductions are collective)."""
    import torch
    import torch.distributed as dist
    from qnmh_sysid2018_b200.engine import Engine

    rank, world, local_rank, stream = ctx["rank"], ctx["world"], ctx["local_rank"], ctx["stream"]
    model, B, N, T = WORKLOADS[workload]
    dtype = {"f64": "float64", "f32": "float32"}[dtype_key]
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
    if workload == "c3" and world > 1:
        # the 4096 evaluations are one fixed batch sharded over the ranks (strong scaling); the
        # Philox stream of an evaluation is keyed by its global index
        from qnmh_sysid2018_b200.batch import shard_range
        lo, hi = shard_range(B, rank, world)
        thetas, eval_ids, B = thetas[lo:hi], np.arange(lo, hi, dtype=np.uint64), hi - lo
    sharded = workload == "c5" and world > 1  # ONE filter over all ranks (strong scaling)
    if particles_log2:
        N = 1 << particles_log2
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
    for _ in range(warmup):
        eng.execute()
    barrier()
    sampler = None
    if sample_clocks:
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
    for _ in range(steps):
        eng.execute()
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    kernel_ms.append(eng.last_kernel_ms())
    clocks = sampler.summary() if sampler else None
    eng.synchronize()
    ll = eng.fetch_log_like()

    # kernel-only duration, averaged over a few separately timed launches
    for _ in range(min(4, steps)):
        eng.execute()
        kernel_ms.append(eng.last_kernel_ms())
    kernel_ms = float(np.mean(kernel_ms))

    # ---- end to end through the host-buffer call: every step uploads its inputs from pinned host memory
    # (observations, parameters, stream ids) and reads its results back
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    h_obs, h_theta, h_ev = pin(obs), pin(thetas), pin(eval_ids)
    P = eng.no_params
    outs = {"log_like": pin(np.empty(B)), "filt": pin(np.empty((B, T + 1))), "smo": pin(np.empty((B, T + 1))),
            "grad_terms": pin(np.empty((B, P, T + 1))), "traj_index": pin(np.empty(B, dtype=np.int32)),
            "traj": np.empty((B, T + 1))}

    def e2e_step(i):
        if sharded:
            # an MH iteration on a connected sharded filter: new parameters and stream id go up (the peer
            # mappings and exchange counters stay alive), the per-rank partial results come back and are summed
            dpf.prepare(h_obs, theta0, N, FIXED_LAG, seed=SEED, eval_id=i)
            dpf.execute()
            return dpf.fetch()
        eng.configure(h_obs, h_theta, N, FIXED_LAG, True, True, 0.0, "systematic", seed=SEED, eval_ids=h_ev)
        eng.execute()
        return eng.fetch(outs)

    for i in range(max(1, warmup // 2)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        e2e_step(100 + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if sharded:
        h2d = P * 8 + 16 * 8 + 8  # parameters, derived constants, stream id (the observations stay resident)
    else:
        h2d = h_obs.nbytes + B * 16 * 8 + h_ev.nbytes  # obs + derived per-filter constants + Philox stream ids
    d2h = B * 8 + 2 * B * (T + 1) * 8 + B * P * (T + 1) * 8 + B * 4

    t = torch.tensor([total_ms, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, kernel_ms = float(t[0]), float(t[1]), float(t[2])

    fixed_job = sharded or (workload == "c3" and world > 1)
    units_per_step = B_total * N * T if fixed_job else world * B * N * T
    bytes_unit = algorithmic_bytes_per_particle_step(dtype_key, FIXED_LAG)
    peak, peak_src = measured_peak_gbs()
    per_gpu_units = B * N * T / (world if sharded else 1)
    achieved = bytes_unit * per_gpu_units / (kernel_ms * 1e-3) / 1e9
    res = {"model": model, "B": B, "B_total": B_total, "N": N, "T": T, "fixed_job": fixed_job,
           "value": units_per_step * steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps,
           "e2e_value": units_per_step * steps / e2e_s, "e2e_ms_per_step": e2e_s / steps * 1e3,
           "h2d": int(h2d), "d2h": int(d2h), "kernel_ms": kernel_ms, "achieved": achieved, "peak": peak,
           "peak_src": peak_src, "bytes_unit": bytes_unit, "per_gpu_units": per_gpu_units, "clocks": clocks,
           "log_like": float(ll[0]), "launches": eng.info("kernel_launches_per_execute") * steps,
           "geometry": {"kernel": "pf_cluster_kernel" if eng.info("cluster") > 0 else "pf_persistent_kernel",
                        "cluster": eng.info("cluster"),
                        "ctas_per_filter": eng.info("ctas_per_filter"), "groups": eng.info("n_groups"),
                        "block_threads": eng.info("block_threads"), "filter_threads": eng.info("filter_threads"),
                        "cooperative": eng.info("cooperative")}}
    if sharded:
        dpf.close()
    else:
        eng.close()
    return res


def traffic_entry(workload, dtype_key):
    """DRAM bytes per launch of the persistent kernel from the committed ncu capture (NOT measured in this
    run; null for workloads that were not captured)."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(tpath) as fh:
            data = json.load(fh)
        key = "%s_%s" % (workload, dtype_key)
        return data.get(key), data.get(key + "_source")
    except Exception:
        return None, None


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
    stream = torch.cuda.Stream()  # a real (non-default) stream: the library and the events share it
    torch.cuda.set_stream(stream)
    ctx = {"rank": rank, "world": world, "local_rank": local_rank, "stream": stream}

    m = measure(args.workload, args.dtype, args.steps, args.warmup, ctx, args.particles_log2)

    # ---- multi-GPU runs also record the two workloads that EXCHANGE something (the default one is independent
    # replicates): the C3 batch sharded over the ranks (strong scaling, one gather at the end) and ONE filter
    # sharded over the ranks (C5: children pushed over NVLink, per-step all-gathers inside the kernel)
    extra = {}
    if world > 1 and args.workload == "c2" and not args.no_extra:
        def brief(r):
            return {"value": r["value"], "unit": "particle-timesteps/s", "ms_per_step": r["ms_per_step"],
                    "e2e": r["e2e_value"], "e2e_ms_per_step": r["e2e_ms_per_step"], "kernel_ms": r["kernel_ms"],
                    "roofline_frac_per_gpu": r["achieved"] / r["peak"], "scaling": "strong",
                    "no_particles": r["N"], "no_obs": r["T"], "filters": r["B_total"], "log_like_sample": r["log_like"],
                    "ctas_per_filter": r["geometry"]["ctas_per_filter"]}
        extra["c3_strong"] = brief(measure("c3", args.dtype, 3, 3, ctx, sample_clocks=False))
        for k in (24, 26):
            extra["c5_2p%d" % k] = brief(measure("c5", args.dtype, 3, 3, ctx, particles_log2=k, sample_clocks=False))

    if rank == 0:
        model, N, T = m["model"], m["N"], m["T"]
        traffic, traffic_src = traffic_entry(args.workload, args.dtype)
        cpu = run_cpu_baseline(model, m["B"], N, T, procs=1) if world == 1 and not args.no_cpu else None
        line = {
            "metric": "particle-timesteps/sec (PF+fixed-lag smoother)", "value": m["value"],
            "unit": "particle-timesteps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if m["fixed_job"] else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, model, m["B_total"], N, T),
            "clocks": m["clocks"],
            "e2e": {"value": m["e2e_value"], "unit": "particle-timesteps/s", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_ms_per_step"]},
            "gpu_launches": m["launches"],
            "roofline": {"bound": "hbm", "achieved": m["achieved"], "peak": m["peak"], "unit": "GB/s",
                         "frac": m["achieved"] / m["peak"], "traffic": traffic,
                         "traffic_source": (traffic_src or "not captured for this workload") +
                                           " -- ncu dram__bytes of one launch, copied from profiles/traffic.json, not measured in this run",
                         "kernel": m["geometry"]["kernel"], "kernel_ms": m["kernel_ms"],
                         "algorithmic_bytes_per_particle_step": m["bytes_unit"],
                         "algorithmic_bytes_per_launch": m["bytes_unit"] * m["per_gpu_units"],
                         "peak_source": m["peak_src"]},
            "geometry": m["geometry"],
            "log_like_sample": m["log_like"],
        }
        if extra:
            line["extra"] = extra
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
    ap.add_argument("--no-extra", action="store_true", help="multi-GPU: skip the c3_strong / c5 records")
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
