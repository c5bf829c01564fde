"""torchrun --nproc-per-node N tools/dist_check.py [--n 100000]
Parity of the multi-GPU path against the committed golden value (holes, n = 5570, general model)
and, optionally, timing of one evaluation at a large n (Cholesky-phase TFLOP/s over all GPUs)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cocons_b200 as cb  # noqa: E402
from cocons_b200 import _lib  # noqa: E402
from cocons_b200.distributed import DistributedDenseLikelihood  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", dest="n", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    out = {"world": world}
    gold = json.load(open("tests/golden/n2ll_cases.json"))["cases"]
    c = [g for g in gold if g["name"] == "holes_full_general"][0]
    D = np.load("tests/golden/datasets.npz")["holes_training"]
    n = c["n"]
    X = cb.getScale(np.column_stack([np.ones(n), D[:n, 2], D[:n, 3]]))["std.covs"]
    pp = {k: (np.array(v, dtype=bool) if isinstance(v, list) else v) for k, v in c["par_pos"].items()}
    tl = cb.getModelLists(np.array(c["theta"]), pp, "diff")
    with DistributedDenseLikelihood(D[:n, :2], X, D[:n, 4]) as d:
        t = d.terms(_lib.ML, tl, c["limits"], tl["mean"])
    v = n * np.log(2 * np.pi) + 2 * t["logdet"] + t["quad"][0]
    out["holes_full_general"] = {"value": v, "golden": c["values"]["ml"], "rel": abs(v - c["values"]["ml"]) / abs(v)}
    if args.n:
        sys.path.insert(0, ".")
        import bench
        locs, X, z = bench.synthetic(args.n)
        sampler = bench.ClockSampler(local) if rank == 0 else None
        with DistributedDenseLikelihood(locs, X, z) as d:
            for rep in range(args.reps):
                if sampler is not None and rep == args.reps - 1:
                    sampler.start()
                t0 = time.perf_counter()
                t = d.terms(_lib.ML, bench.theta_at(rep, 0), bench.LIMITS, bench.THETA["mean"])
                wall = time.perf_counter() - t0
                ph = dict(d.last_phase_s)
            out["large"] = {"n": args.n, "wall_s": wall, **ph,
                            "chol_tflops_all_gpus": bench.flops_chol(args.n) / ph["assemble_factor_s"] / 1e12,
                            "value": args.n * np.log(2 * np.pi) + 2 * t["logdet"] + float(t["quad"][0])}
            if sampler is not None:
                out["large"]["clocks_rank0"] = sampler.stop()
            if d.last_wait_ms is not None:
                out["large"]["bcast_wait_ms_rank0"] = d.last_wait_ms
    if rank == 0:
        print("DIST_CHECK " + json.dumps(out))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
