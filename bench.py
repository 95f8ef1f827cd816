#!/usr/bin/env python
"""bench.py -- train-points/sec of the projected-process statistics hot path
(`getMatrixKmnKnmAndVectorKmny`, PGPH:20-36) on N B200s, next to the reference's CPU path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" = one pass of the hot path over one batch of synthetic input: begin(active set) ->
accumulate(shard) -> finish (all-reduce of [G;b] across ranks).  Workload at N GPUs: BASELINE configs[1]
per GPU (synthetic 1M x 16 fp32, active=1000, `1*ARDRBFKernel(16) + 1.const*EyeKernel`, sigma2=1e-4,
theta fixed at C=1, beta_k=sqrt(18/d)) -- weak scaling, value = (N * 1M) / max-over-ranks device time.

  value : inputs resident in HBM before the timed region (sgp_stats_accumulate_device), CUDA events on the
          library's stream, L2 flushed between steps (the 64 MB shard is smaller than the 126 MB L2).
  e2e   : the same step through the public host-buffer entry (sgp_stats_accumulate from pinned host
          memory, G and b copied back), host<->device copies inside the timed region.
"""
from __future__ import annotations

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

# dram__bytes_read.sum + dram__bytes_write.sum of one kmn_gram_i8_kernel launch over the 1M-point shard, from the
# committed `ncu --set full` capture (profiles/); None until measured
TRAFFIC_BYTES_PER_LAUNCH = 136_726_016   # profiles/r01_i8_gram_ncu_summary.txt: 132.35 MB read + 4.37 MB write

METRIC = "train_points_per_sec"
UNIT = "points/s"
N_PER_GPU, D, M, N_E = 1_000_000, 16, 1000, 100
SIGMA2 = 1e-4


def workload_config(n_gpus: int) -> dict:
    return {"workload": "synthetic %dx%d fp32 regression per GPU, active=%d, expert=%d, "
                        "1*ARDRBFKernel(%d)+1.const*EyeKernel, sigma2=1e-4, C=1, beta=sqrt(18/d)"
                        % (N_PER_GPU, D, M, N_E, D),
            "stage": "stats (K_mn + K_mn K_nm + K_mn y, all-reduced)", "n_per_gpu": N_PER_GPU, "d": D, "m": M,
            "n_total": N_PER_GPU * n_gpus, "parallelism": "points sharded over %d GPU(s), one ncclAllReduce of [G;b]" % n_gpus,
            "l2": "flushed between timed steps (256 MiB memset)",
            "precision_mode": "SGP_PREC_AUTO -> tcgen05 int8 exact-accumulation kernel (fp16-split distance contraction, "
                              "23-bit fixed-point elements, int32 accumulators folded into fp64)"}


def make_shard(rank: int):
    """X ~ U[0,1) generated in fp32 (the oracle consumes the same values up-cast), y = sin(sum x) + 0.1 eps."""
    rng = np.random.default_rng(13 + rank)
    X = rng.random((N_PER_GPU, D), dtype=np.float32)
    y = np.sin(X.astype(np.float64).sum(1)) + 0.1 * rng.standard_normal(N_PER_GPU)
    return X, y


def active_set():
    rng = np.random.default_rng(7)
    X0 = np.random.default_rng(13).random((N_PER_GPU, D), dtype=np.float32)
    return X0[rng.permutation(N_PER_GPU)[:M]].astype(np.float64)


def algorithmic_flops_per_point(m=M, d=D) -> float:
    """2*m*d (distance contraction) + m*(m+1) (symmetric Gram) + 2*m (K_mn y)."""
    return 2.0 * m * d + m * (m + 1.0) + 2.0 * m


def cpu_sample_points(cores: int) -> int:
    """Bounded CPU sample: ~80 experts (8000 points) per worker so per-expert work, not the final sum of the
    per-worker m x m partials, dominates -- capped at the whole 1M-point shard."""
    return int(min(N_PER_GPU, max(40_000, cores * 80 * N_E)))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")) -> dict:
        """Median SM clock over the samples taken inside [t_begin, t_end] (host perf_counter), i.e. under load;
        throttle reasons over the same window.  The sampler runs from before the warm-up (20 ms period)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, all_sm = [], None, set(), []
        for row in self.rows:
            ts, r = row[0], row[1:]
            try:
                v_sm = float(r[0]); mx = float(r[1])
            except Exception:
                continue
            all_sm.append(v_sm)
            if ts < t_begin - 0.02 or ts > t_end + 0.02:
                continue
            sm.append(v_sm)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else (float(np.max(all_sm)) if all_sm else None), "sm_max_mhz": mx,
                "samples": len(sm), "samples_total": len(all_sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path (here: the oracle port -- the reference is Scala
    and there is no JVM in this image), all host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from oracle.cpu_baseline import stats_parallel
    from oracle.cpu_baseline import usable_cores
    cores = usable_cores()
    sample = cpu_sample_points(cores)
    X, y = make_shard(0)
    X, y = X[:sample].astype(np.float64), y[:sample]
    Z = active_set()
    beta = np.full(D, np.sqrt(18.0 / D))
    fac = lambda: (1 * oracle.ARDRBFKernel(beta) + oracle.const(1) * oracle.EyeKernel()
                   + oracle.const(SIGMA2) * oracle.EyeKernel())
    theta = fac().get_hyperparameters()
    for _ in range(args.warmup):
        stats_parallel(X[:4000], y[:4000], Z, fac, theta, N_E, cores)
    t = 0.0
    for _ in range(args.steps):
        _, _, dt = stats_parallel(X, y, Z, fac, theta, N_E, cores)
        t += dt
    value = sample * args.steps / t
    cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d points of the 1M-point shard per step (path is linear in N at fixed m,d,n_e)" % sample}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": workload_config(args.gpus), "cpu_baseline": cpu,
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import spark_gp_b200 as sg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = sg.ProjectedProcessEngine(local_rank)
    if world > 1:
        ids = [sg.ProjectedProcessEngine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0], rank, world)

    Xh, yh = make_shard(rank)
    Z = active_set()
    beta = np.full(D, np.sqrt(18.0 / D))
    kernel = 1 * sg.ARDRBFKernel(beta) + sg.const(1) * sg.EyeKernel() + sg.const(SIGMA2) * sg.EyeKernel()
    # device-resident copies (value leg) and pinned host copies (e2e leg)
    Xd = torch.from_numpy(Xh).to(dev)
    yd = torch.from_numpy(yh).to(dev)
    Xp = torch.from_numpy(Xh).pin_memory()
    yp = torch.from_numpy(yh).pin_memory()
    Gp = torch.empty((M, M), dtype=torch.float64).pin_memory()
    bp = torch.empty(M, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        eng.event_record(0)
        eng.begin(kernel, Z)
        eng.accumulate_ptr(Xd.data_ptr(), True, yd.data_ptr(), N_PER_GPU, device=True)
        eng.finish(copy_out=False)                    # all-reduce of [G;b]; statistics stay on the device
        eng.event_record(1)
        return eng.event_elapsed_ms(0, 1)

    def step_e2e():
        t0 = time.perf_counter()
        eng.begin(kernel, Z)
        eng.accumulate_ptr(Xp.data_ptr(), True, yp.data_ptr(), N_PER_GPU, device=False)
        eng._check(eng._lib.sgp_stats_finish(eng._h, Gp.data_ptr(), bp.data_ptr()))
        return 1e3 * (time.perf_counter() - t0)

    def l2_flush():
        flush.zero_()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    l2_flush()
    launches0 = eng.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    kern_ms, kern_n = 0.0, 0
    for _ in range(args.steps):
        dev_ms += step_device()
        kms, kn = eng.gram_kernel_time()
        kern_ms += kms; kern_n += kn
        l2_flush()
        # every step ends in an all-reduce, so a rank that starts its step early spends the skew waiting INSIDE its
        # event-timed region: line the ranks up again (outside the timed region) before the next step
        barrier()
    t_wall1 = time.perf_counter()
    wall_ms = 1e3 * (t_wall1 - t_wall0)
    launches = eng.launch_count() - launches0

    # e2e leg: host buffers, copies inside the timed region (wall clock around the blocking API calls)
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        e2e_ms += step_e2e()
        barrier()

    # tail (m x m, fp64, rank 0 does it in a fit) -- reported beside the stats number
    eng.begin(kernel, Z)
    eng.accumulate_ptr(Xd.data_ptr(), True, yd.data_ptr(), N_PER_GPU, device=True)
    eng.finish(copy_out=False)
    eng.magic(copy_out=False)                       # first call pays cuSOLVER's lazy initialisation
    t0 = time.perf_counter()
    eng.magic(copy_out=False)
    tail_ms = 1e3 * (time.perf_counter() - t0)

    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None   # stopped after the e2e leg: more samples to fall back on
    t = torch.tensor([dev_ms, e2e_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, kern_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        n_total = N_PER_GPU * world
        value = n_total * args.steps / (dev_ms / 1e3)
        e2e_value = n_total * args.steps / (e2e_ms / 1e3)
        peak_tf, peak_src = measured_peaks()
        launch_ms = kern_ms / max(kern_n, 1)
        achieved_tf = algorithmic_flops_per_point() * N_PER_GPU / (launch_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "kmn_gram_i8_kernel", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "peak_source": peak_src,
                "traffic": TRAFFIC_BYTES_PER_LAUNCH, "launch_ms": launch_ms, "launches_timed": kern_n,
                "algorithmic_flops_per_launch": algorithmic_flops_per_point() * N_PER_GPU,
                "note": "algorithmic flops = N*(2md + m(m+1) + 2m); the kernel executes 6 int8 products per Gram "
                        "tile pair (exact 23-bit arithmetic) + recomputed distance tiles, none of which is credited"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            from oracle.cpu_baseline import stats_parallel
            from oracle.cpu_baseline import usable_cores
            cores = usable_cores()
            sample = cpu_sample_points(cores)
            obeta = beta
            fac = lambda: (1 * oracle.ARDRBFKernel(obeta) + oracle.const(1) * oracle.EyeKernel()
                           + oracle.const(SIGMA2) * oracle.EyeKernel())
            stats_parallel(Xh[:4000].astype(np.float64), yh[:4000], Z, fac, fac().get_hyperparameters(), N_E, cores)
            _, _, dt = stats_parallel(Xh[:sample].astype(np.float64), yh[:sample], Z, fac,
                                      fac().get_hyperparameters(), N_E, cores)
            cpu = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d points of the same shard, oracle restatement, %d worker processes" % (sample, cores)}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": workload_config(world), "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                       "h2d_bytes_per_step": int(N_PER_GPU * D * 4 + N_PER_GPU * 8 + M * D * 8),
                       "d2h_bytes_per_step": int((M * M + M) * 8)},
               "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
               "tail_ms": tail_ms, "wall_ms_timed_region": wall_ms}
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
