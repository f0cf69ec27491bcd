#!/usr/bin/env python3
"""Parameter sweep (BASELINE config 5): (h, m, step_length, foot_spread, F_ext) over many scenarios,
sharded across the GPUs of one box, NCCL all-gather of the per-scenario outcomes.

    python scripts/run_param_sweep.py --scenarios 65536
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_param_sweep.py --scenarios 1048576
"""

import argparse
import json
import os
import time

import numpy as np

from _common import base_config, write_json
from mpc_bipedal import parallel, sweeps


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str)
    ap.add_argument("--scenarios", type=int, default=65536)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--precision", choices=["fp64", "f16x3"], default="f16x3")
    ap.add_argument("--non-strict", action="store_true")
    ap.add_argument("--output-dir", type=str, default="results")
    args = ap.parse_args(argv)

    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    config = base_config(args.config)
    table = sweeps.make_sweep(args.scenarios, seed=args.seed)
    precision = 1 if args.precision == "f16x3" else 0
    # warm-up: CUDA context, kernel images, plans of every CoM height (each rank runs all of them), the first collective
    warm = sweeps.make_sweep(min(len(table), 64 * world), seed=args.seed)
    w = sweeps.run_shard(config, warm, rank, world, strict=not args.non_strict, precision=precision)
    parallel.gather_by_index(w["record"], w["index"], len(warm), world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    local = sweeps.run_shard(config, table, rank, world, strict=not args.non_strict, precision=precision)
    gathered = parallel.gather_by_index(local["record"], local["index"], len(table), world)
    torch.cuda.synchronize()
    seconds = time.perf_counter() - t0
    solves = torch.tensor([local["solves"]], dtype=torch.float64, device=gathered.device)
    if world > 1:
        dist.all_reduce(solves)
    if rank == 0:
        out = parallel.unpack_outcomes(gathered.cpu())
        summary = {"scenarios": len(table), "gpus": world, "seconds": seconds, "qp_solves": int(solves.item()),
                   "qp_solves_per_s": solves.item() / seconds, "precision": args.precision,
                   "fallen": int(out["fell"].sum()), "not_converged": int((out["status"] != 0).sum()),
                   "max_zmp_violation_m": float(out["max_violation"].max()),
                   "note": "qp_solves counts the walks that ran to their end (a walk that stops at an unconverged QP counts for nothing)"}
        print(json.dumps(summary))
        os.makedirs(args.output_dir, exist_ok=True)
        order = table.scenario_id
        np.savez_compressed(os.path.join(args.output_dir, "param_sweep.npz"), scenario_id=order,
                            h=table.h_values[table.h_index], variant=table.variant, mass=table.mass, F_ext=table.F_ext,
                            fell=out["fell"].numpy(), max_violation=out["max_violation"].numpy(),
                            lean_max=out["lean_max"].numpy(), status=out["status"].numpy(), iters=out["iters"].numpy())
        write_json(os.path.join(args.output_dir, "param_sweep.json"), summary)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
