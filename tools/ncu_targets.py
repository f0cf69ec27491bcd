#!/usr/bin/env python3
"""Small driver for `ncu --set full` captures: one launch of a chosen kernel on a chosen workload.
    python tools/ncu_targets.py f16 4096 | fp64 4096 (ADMM) | riccati 4096 (active set) | free 4096 | free 262144 | herdt 8192"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal import engine
from bench import sweep_inputs

what, B = sys.argv[1], int(sys.argv[2])
cfg, z_max, z_min, F, tick = sweep_inputs(B)
precision = engine.PRECISION_F16X3 if what == "f16" else engine.PRECISION_FP64
plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, precision=precision, solver="admm" if what == "fp64" else "active_set")
dev = plan.device
x0 = torch.zeros((B, 3), dtype=torch.float64, device=dev)
zmin, zmax = torch.as_tensor(z_min).to(dev), torch.as_tensor(z_max).to(dev)
imp, tk = torch.as_tensor(cfg.dt * F / cfg.m).to(dev), torch.as_tensor(tick).to(dev)
mode = engine.MODE_UNCONSTRAINED if what == "free" else engine.MODE_STRICT
out = {}
for _ in range(2):
    plan.rollout(mode, x0, x0, zmin, zmax, impulse=imp, force_tick=tk, out=out, want_readout=(what == "free"))
torch.cuda.synchronize()
print(what, B, "iters", int(out["iters"].sum()))
