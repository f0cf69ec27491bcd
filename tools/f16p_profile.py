#!/usr/bin/env python3
"""Cycle breakdown of the pair kernel (needs tools/_prof_liblipm.so built with -DLIPM_F16P_PROFILE)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
from mpc_bipedal import engine
engine.LIBRARY_PATH = os.path.join(REPO, "tools", "_prof_liblipm.so")
import torch
from bench import sweep_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg, z_max, z_min, F, tick = sweep_inputs(B)
plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, precision=engine.PRECISION_F16X3)
dev = plan.device
x0 = torch.zeros((B, 3), dtype=torch.float64, device=dev)
zmin, zmax = torch.as_tensor(z_min).to(dev), torch.as_tensor(z_max).to(dev)
imp, tk = torch.as_tensor(cfg.dt * F / cfg.m).to(dev), torch.as_tensor(tick).to(dev)
plan.rollout(engine.MODE_STRICT, x0, x0, zmin, zmax, impulse=imp, force_tick=tk)
torch.cuda.synchronize()
