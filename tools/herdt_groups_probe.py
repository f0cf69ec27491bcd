#!/usr/bin/env python3
"""Per-timeline-group timing of BASELINE configs[3] (8192 randomized Herdt plans): each group alone, then all on their
own streams.  Shows where the concurrent run loses against the sum of its parts."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal import engine
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.herdt import randomized_plans, run_plan_groups

engine.keep_pool_memory(0)
cfg = MPCConfig(); cfg.method = "herdt"
plans = randomized_plans(cfg, 8192, seed=2, on_device=True)
warm = dict(plans); warm["groups"] = [dict(g, ids=g["ids"][:2]) for g in plans["groups"]]
_, controllers, _ = run_plan_groups(cfg, warm)
torch.cuda.synchronize()
total = 0.0
for k, g in enumerate(plans["groups"]):
    one = dict(plans); one["groups"] = [g]
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res, _, streams = run_plan_groups(cfg, one, [controllers[k]])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ticks = len(g["states"]) - 1
    it = int(res[0]["iters"].sum())
    fb = int(res[0]["status"].sum())
    total += dt
    print(f"group {k}: {len(g['ids']):5d} plans x {ticks} ticks: {dt*1e3:7.1f} ms = {dt/ticks*1e6:6.0f} us/tick, {it/(len(g['ids'])*ticks):5.1f} it/solve, "
          f"fallback ticks/plan {fb/len(g['ids']):.1f}", flush=True)
print(f"sum of groups alone: {total*1e3:.1f} ms")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res, _, streams = run_plan_groups(cfg, plans, controllers)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"all groups on their own streams: {dt*1e3:.1f} ms")
