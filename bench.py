#!/usr/bin/env python
"""bench.py - neg2loglik evaluations/sec (assembly + Cholesky + solves) at n = 50 000 (BASELINE.json).

    python bench.py --gpus 1 --steps K --warmup W                 our arm, one B200
    torchrun --nproc-per-node N ... bench.py --gpus N ...          our arm, N replicas (weak scaling)
    python bench.py --impl reference --steps K --warmup W          the reference's CPU path, host cores

A step is one full objective evaluation (GetNeg2loglikelihood, R/neg2loglikelihood.R:183-222) of
the synthetic nonstationary model of SURVEY.md §8(d) config 3 at a fresh theta (the optimiser's
finite-difference neighbours, R/optim.R:237-259).  With N GPUs every rank evaluates its own
theta on its own replica of the data - the reference's own fan-out, no data-path collective.
Prints ONE JSON line (rank 0).
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

SEED = 20261018
THETA = {"mean": np.zeros(5), "std.dev": np.array([0.2, 0.15, 0.10, -0.05, 0.05]),
         "scale": np.array([-1.6, 0.2, -0.15, 0.1, -0.1]), "aniso": np.array([0.1, 0.2, -0.1, 0.05, 0]),
         "tilt": np.array([0.3, -0.2, 0.1, 0.1, -0.1]), "smooth": np.array([0.2, 0.3, -0.2, 0.1, 0.1]),
         "nugget": np.array([-4, 0.1, 0.1, 0, 0])}
LIMITS = [0.5, 2.5]
DERIVED_FP64_PEAK_TFLOPS = 37.2  # 148 SM x 128 flop/clk x 1.965 GHz


def synthetic(n):
    """locs ~ U(-1,1)^2, four covariates, z ~ N(0,1), numpy PCG64 seed 20261018."""
    rng = np.random.default_rng(SEED)
    locs = rng.uniform(-1, 1, (n, 2))
    c1, c2 = (locs[:, 0] + 1) / 2, (locs[:, 1] + 1) / 2
    raw = np.column_stack([np.ones(n), c1, c2, c1 * c2,
                           0.5 + 0.5 * np.sin(np.pi * locs[:, 0]) * np.cos(np.pi * locs[:, 1])])
    mean, sd = raw.mean(axis=0), raw.std(axis=0, ddof=1)
    mean[0], sd[0] = 0.0, 1.0
    X = np.asfortranarray((raw - mean) / sd)  # getScale, R/getFunctions.R:376-436
    z = rng.standard_normal(n)
    return np.asfortranarray(locs), X, z


def theta_at(step, rank):
    """A distinct evaluation point per (step, rank): the base theta nudged the way a finite-difference
    gradient would nudge it (ndeps = eps^(1/4), R/profile.R:14)."""
    th = {k: v.copy() for k, v in THETA.items()}
    keys = ("std.dev", "scale", "aniso", "tilt", "smooth", "nugget")
    idx = (step * 131 + rank * 17) % 30
    th[keys[idx // 5]][idx % 5] += np.finfo(float).eps ** 0.25
    return th


def flops_chol(n):
    return n ** 3 / 3 + n ** 2 / 2 + n / 6


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])), smax.append(float(r[2])), power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# the reference's CPU path on the host cores (oracle/_ref = its own source compiled, + LAPACK)
# ---------------------------------------------------------------------------------------------
def _asm_sample(args):
    n_s, kind = args
    from oracle import cov
    locs, X, _ = synthetic(n_s)
    t0 = time.perf_counter()
    cov.cov_rns(THETA, locs, X, LIMITS, kind=kind)
    return time.perf_counter() - t0


def cpu_reference_sample(n, n_asm=2200, n_chol=7000):
    """Bounded sample of one evaluation at size n on the host cores.

    assembly : the reference's single-threaded pair loop (src/cocons_full.cpp:257-313) on n_asm
               sites -> ns/pair, scaled to n(n-1)/2 pairs; run on every core at once (one sample
               per core) the way optimParallel's workers would run it (R/optim.R:117-121)
    cholesky : LAPACK dpotrf on an n_chol matrix -> flop/s, scaled to n^3/3; (a) one thread per
               worker on every core at once, (b) one worker with every thread
    Returns evals/s for the better of the two layouts.
    """
    import multiprocessing as mpc

    import scipy.linalg as sla
    from threadpoolctl import threadpool_limits

    from oracle import cov
    cores = os.cpu_count() or 1
    kind = "reference" if cov.have_reference() else "restatement"
    with mpc.get_context("fork").Pool(cores) as pool:
        t_asm = float(np.mean(pool.map(_asm_sample, [(n_asm, kind)] * cores)))
    ns_pair = t_asm / (n_asm * (n_asm - 1) / 2) * 1e9
    t_asm_full = ns_pair * 1e-9 * n * (n - 1) / 2
    rng = np.random.default_rng(1)
    A = rng.standard_normal((n_chol, 64))
    S = A @ A.T + n_chol * np.eye(n_chol)
    with threadpool_limits(limits=cores):
        t0 = time.perf_counter()
        sla.cholesky(S, lower=True, check_finite=False)
        t_all = time.perf_counter() - t0
    gflops_all = flops_chol(n_chol) / t_all / 1e9
    with threadpool_limits(limits=1):
        m1 = 2500
        t0 = time.perf_counter()
        sla.cholesky(S[:m1, :m1], lower=True, check_finite=False)
        t_one = time.perf_counter() - t0
    gflops_one = flops_chol(m1) / t_one / 1e9
    # (a) `cores` workers, each single-threaded end to end (the reference's layout)
    thr_workers = cores / (t_asm_full + flops_chol(n) / (gflops_one * 1e9))
    # (b) one worker, assembly single-threaded (it has no threads), LAPACK on every core
    thr_single = 1.0 / (t_asm_full + flops_chol(n) / (gflops_all * 1e9))
    best = max(thr_workers, thr_single)
    return {"value": best, "unit": "evals/s", "cores": cores, "kind": "reference" if kind == "reference" else "port",
            "sample": ("extrapolated from a bounded sample: reference pair loop (general Bessel branch) on %d sites "
                       "per core = %.0f ns/pair; LAPACK dpotrf n=%d all threads = %.0f GFLOP/s, n=%d one thread = "
                       "%.1f GFLOP/s; layouts: %d single-threaded workers %.3e evals/s, one worker + threaded "
                       "LAPACK %.3e evals/s" % (n_asm, ns_pair, n_chol, gflops_all, m1, gflops_one, cores,
                                                thr_workers, thr_single)),
            "ns_per_pair": ns_pair, "dpotrf_gflops_all_threads": gflops_all}


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals, t_steps = [], []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_reference_sample(args.n)
        if s >= args.warmup:
            vals.append(res["value"])
            t_steps.append(time.perf_counter() - t0)
    v = float(np.mean(vals))
    res["value"] = v
    line = {"impl": "reference", "metric": "neg2loglik evals/sec (assembly+Cholesky) at n=50k", "value": v,
            "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic n=%d nonstationary Matern, 4 covariates (p=5), r=1, ML objective"
                                   % args.n, "n": args.n},
            "cpu_baseline": res,
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "sample_seconds_per_step": float(np.mean(t_steps))}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def dgemm_ceiling(torch, dev, n=8192, reps=4):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2 * n ** 3 / e0.elapsed_time(e1) / 1e9)
    del a, b
    torch.cuda.empty_cache()
    return best


def run_ours(args, rank, world, local_rank):
    import torch

    import cocons_b200 as cb
    from cocons_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - cocons_b200 has no CPU path to fall back to")
    torch.cuda.set_device(local_rank)
    os.environ["COCONS_DEVICE"] = str(local_rank)  # device of the one-shot (host-buffer) entry point
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its banner / debug lines to stdout by default; stdout carries the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    locs, X, z = synthetic(n)
    p = X.shape[1]
    par_pos = {k: np.ones(p, dtype=bool) for k in ("mean", "std.dev", "scale", "aniso", "tilt", "smooth", "nugget")}
    lam = (0.0, 0.0, 0.0)
    L = _lib.lib()

    peak = dgemm_ceiling(torch, dev) if (rank == 0 and not args.profile) else 0.0

    # ---- device-resident arm: inputs in HBM before the timed region --------------------------
    stream = torch.cuda.current_stream().cuda_stream
    ctx = cb.DenseLikelihood(locs, X, z, device=local_rank, stream=stream)
    values = []
    for s in range(args.warmup):
        t = ctx.terms(_lib.ML, theta_at(s, rank), LIMITS, THETA["mean"])
    phases = {"assembly_ms": [], "factor_ms": [], "solve_ms": [], "kernel_ms": []}
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = L.cocons_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        t = ctx.terms(_lib.ML, theta_at(args.warmup + s, rank), LIMITS, THETA["mean"])
        values.append(n * np.log(2 * np.pi) + 2 * t["logdet"] + float(t["quad"][0]))
        tm = ctx.timings()
        for k in phases:
            phases[k].append(tm[k])
    e1.record()
    barrier()
    launches = L.cocons_launch_count() - launches0
    elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    ctx.close()

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": elapsed / args.steps * 1e3,
                              "phases_ms": {k: float(np.mean(v)) for k, v in phases.items()},
                              "gpu_launches": int(launches)}))
        return
    # ---- end to end: the call an R user makes, host buffers in, scalar out, every step -------
    for s in range(min(args.warmup, 2)):
        cb.GetNeg2loglikelihood(theta_vec(theta_at(s, rank), par_pos), par_pos, locs, X, LIMITS, z, n, lam)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        v = cb.GetNeg2loglikelihood(theta_vec(theta_at(args.warmup + s, rank), par_pos), par_pos, locs, X, LIMITS,
                                    z, n, lam)
        assert abs(v - values[s]) <= 1e-9 * abs(v), (v, values[s])
    torch.cuda.synchronize()
    e2e_elapsed = max_over_ranks(time.perf_counter() - t0)
    L.cocons_release_workspace()
    barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    value = world * args.steps / elapsed
    f_ms = float(np.mean(phases["factor_ms"]))
    phase_tflops = flops_chol(n) / (f_ms * 1e-3) / 1e12
    # dominant kernel: the largest trailing-update launch of each factorisation (C -= P P^T on the
    # (n_pad - 2K)^2 lower triangle, K = 768 at n = 50 000), bracketed by CUDA events inside the timed region
    n_pad = (n + 127) // 128 * 128
    # outer panel width, the rule of chol_outer() in csrc/chol.cu: 768 from 16 384 sites up, else 512
    kk = 128 * max(1, min(int(os.environ.get("COCONS_CHOL_OUTER", "6" if n_pad >= 16384 else "4")), 16))
    rest = n_pad - 2 * kk
    kernel_flops = rest * (rest + 1) / 2 * 2 * kk  # algorithmic: lower triangle incl. diagonal
    k_ms = float(np.mean(phases["kernel_ms"]))
    achieved = kernel_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_tma_ncu_traffic.json")) as f:
            tr = json.load(f)
        if tr.get("n") == n:
            traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
    except (OSError, ValueError, KeyError):
        pass
    cpu = cpu_reference_sample(n) if world == 1 and not args.no_cpu_baseline else None
    line = {
        "metric": "neg2loglik evals/sec (assembly+Cholesky) at n=50k", "value": value, "unit": "evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic n=%d nonstationary Matern, 4 covariates (p=5), r=1, ML objective" % n,
                   "n": n, "p": p, "r": 1, "parallelism": "replicas x%d (one theta per GPU, no collective)" % world,
                   "l2": "working set %.1f GB per evaluation >> 126 MB L2; no flush needed" % (8e-9 * n * n),
                   "objective_value_step0": values[0]},
        "phases_ms": {k: float(np.mean(v)) for k, v in phases.items() if k != "kernel_ms"},
        "assembly_pairs_per_s": n * (n - 1) / 2 / (float(np.mean(phases["assembly_ms"])) * 1e-3),
        "roofline": {"bound": "tensor",
                     "kernel": "gemm_nt_tma_kernel<64,2,0> (DMMA.8x8x4 trailing update, bulk-copy fed): the largest "
                               "launch of each factorisation, %d^2 lower triangle x K=%d" % (rest, kk),
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if (peak and achieved) else None,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_flops_per_launch": kernel_flops, "launch_ms": k_ms,
                     "peak_source": "same-run cuBLAS dgemm 8192^3 (torch.matmul float64), burst; "
                                    "MEASURED_PEAKS.json has no FP64 entry",
                     "derived_peak": DERIVED_FP64_PEAK_TFLOPS,
                     "frac_of_derived": achieved / DERIVED_FP64_PEAK_TFLOPS if achieved else None,
                     "dmma_pipe_peak": 36.9,  # tools/micro/dmma_peak.cu on this pool (register-only DMMA loop)
                     "cholesky_phase": {"achieved": phase_tflops, "frac": phase_tflops / peak if peak else None,
                                        "frac_of_derived": phase_tflops / DERIVED_FP64_PEAK_TFLOPS,
                                        "algorithmic_flops_per_eval": flops_chol(n), "ms": f_ms}},
        "e2e": {"value": world * args.steps / e2e_elapsed, "unit": "evals/s",
                "h2d_bytes_per_step": int(8 * (n * 2 + n * p + n + 7 * p)), "d2h_bytes_per_step": int(8 * 2 + 4)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def theta_vec(tl, par_pos):
    """Optimiser-level theta vector whose getModelLists(type='diff') image is tl (all aspects free)."""
    a, b = tl["std.dev"] + tl["scale"], tl["std.dev"] - tl["scale"]
    parts = {"mean": tl["mean"], "std.dev": a, "scale": b, "aniso": tl["aniso"], "tilt": tl["tilt"],
             "smooth": tl["smooth"], "nugget": tl["nugget"]}
    return np.concatenate([parts[k] for k in ("mean", "std.dev", "scale", "aniso", "tilt", "smooth", "nugget")])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--sites", dest="n", type=int, default=50000, help="sites (the metric is quoted at 50 000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run: device-resident arm only (no DGEMM probe, no e2e leg, no CPU sample)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
