#!/usr/bin/env python3
"""One timeline group of the randomized Herdt plans, for ncu launch lists: python tools/herdt_one_group.py GROUP [PLANS]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.herdt import randomized_plans, run_plan_groups
cfg = MPCConfig(); cfg.method = "herdt"
plans = randomized_plans(cfg, 8192, seed=2, on_device=True)
k = int(sys.argv[1])
one = dict(plans); one["groups"] = [plans["groups"][k]]
if len(sys.argv) > 2:
    one["groups"] = [dict(plans["groups"][k], ids=plans["groups"][k]["ids"][: int(sys.argv[2])])]
res, ctrl, streams = run_plan_groups(cfg, one)
torch.cuda.synchronize()
print("group", k, "plans", len(one["groups"][0]["ids"]), "ticks", len(one["groups"][0]["states"]) - 1, "iters", int(res[0]["iters"].sum()))
