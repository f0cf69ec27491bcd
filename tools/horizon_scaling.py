#!/usr/bin/env python3
"""Horizon scaling (BASELINE configs[2] / SURVEY 8(d) config 3): one control tick for a batch of
independent axis problems, N in {100,150,300,600}, unconstrained vs strict, CUDA-event timed.

Protocol of the reference's scripts/run_compare_runtime.py:59-73 (3 warm-ups, >= 10 timed calls of
one tick) with NON-trivial inputs -- the reference's x = 0 / +-0.1 bounds give the trivial optimum
u = 0: x ~ N(0, diag(0.05, 0.1, 0.5)^2), bounds = centre c ~ U(-0.05, 0.05) +- 0.055, piecewise
constant with a step every 25-30 samples, seed 1.  Prints one JSON object per (N, mode).
"""

import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def tick_inputs(N, B, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(B, 3)) * np.array([0.05, 0.1, 0.5])
    centre = np.empty((B, N))
    pos = np.zeros(B, dtype=np.int64)
    c = rng.uniform(-0.05, 0.05, B)
    filled = np.zeros(B, dtype=bool)
    col = np.arange(N)[None, :]
    start = np.zeros(B, dtype=np.int64)
    while not filled.all():
        run = rng.integers(25, 31, B)
        mask = (col >= start[:, None]) & (col < (start + run)[:, None])
        centre = np.where(mask, c[:, None], centre)
        start = start + run
        c = np.clip(c + rng.uniform(-0.05, 0.05, B), -0.1, 0.1)
        filled = start >= N
    return x, centre - 0.055, centre + 0.055


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--horizons", type=int, nargs="+", default=[100, 150, 300, 600])
    ap.add_argument("--timed", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--precision", choices=["fp64", "f16x3"], default="fp64",
                    help="strict-mode arithmetic (f16x3: tcgen05 paths, horizons <= 637)")
    args = ap.parse_args()

    import torch
    from mpc_bipedal import engine

    dev = torch.device("cuda", 0)
    rows = []
    for N in args.horizons:
        reduced = args.precision == "f16x3"
        if reduced and N > engine.load_library().lipm_max_reduced_horizon():
            print(json.dumps({"horizon": N, "mode": "strict", "precision": "f16x3", "skipped": "beyond the reduced-precision kernel's horizon limit"}), flush=True)
            continue
        plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, precision=engine.PRECISION_F16X3 if reduced else engine.PRECISION_FP64)
        x, lo, hi = tick_inputs(N, args.batch)
        xd, lod, hid = (torch.as_tensor(a).to(dev) for a in (x, lo, hi))
        for mode, name in ((engine.MODE_UNCONSTRAINED, "unconstrained"), (engine.MODE_UNCONSTRAINED, "unconstrained_full_preview"),
                           (engine.MODE_STRICT, "strict")):
            full = name == "unconstrained_full_preview"      # whole jerk sequence: dense H^-1 product on the tensor pipe
            for _ in range(3):
                res = plan.solve_tick(mode, xd, lod, hid, want_plan=full)
            torch.cuda.synchronize()
            # the ticks are enqueued back to back and timed as one block on the device: the host only has to
            # stay ahead of the GPU (a synchronise per tick would put the launch latency of a busy host on the clock)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.timed):
                res = plan.solve_tick(mode, xd, lod, hid, want_plan=full)
            e1.record()
            torch.cuda.synchronize()
            times = [e0.elapsed_time(e1) / args.timed]
            iters = int(res["iters"].sum().item())
            ms = float(times[0])
            row = {"horizon": N, "mode": name, "precision": args.precision if mode == engine.MODE_STRICT else "fp64", "batch": args.batch, "ms_per_tick": ms,
                   "solves_per_s": args.batch / (ms * 1e-3), "cold_start": True,
                   "mean_iters_per_solve": iters / args.batch,
                   "not_converged": int((res["status"] != 0).sum().item())}
            if full:
                row["algorithmic_tflops"] = 2.0 * N * N * args.batch / (ms * 1e-3) / 1e12      # the H^-1 product alone
            if mode == engine.MODE_STRICT:
                row["algorithmic_tflops"] = 2.0 * N * N * iters / (ms * 1e-3) / 1e12
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
