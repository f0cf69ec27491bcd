#!/usr/bin/env python3
"""Iterations and device time of the FP64 push sweep against the ADMM penalty rho and the relaxation alpha
(design probe: what the library defaults were chosen from)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mpc_bipedal import engine

B = 4096
cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
dev = torch.device("cuda", 0)
zero = torch.zeros((B, 3), dtype=torch.float64, device=dev)
lo, hi = torch.as_tensor(z_min).to(dev), torch.as_tensor(z_max).to(dev)
imp = torch.as_tensor(cfg.dt * F / cfg.m).to(dev)
tk = torch.as_tensor(tick).to(dev)
base = None
for rho in (1.5, 2.0, 2.5, 3.0):
    for alpha in (1.7, 1.75, 1.8, 1.85):
        plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, rho=rho, alpha=alpha)
        for _ in range(2):
            res = plan.rollout(engine.MODE_STRICT, zero, zero, lo, hi, impulse=imp, force_tick=tk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            res = plan.rollout(engine.MODE_STRICT, zero, zero, lo, hi, impulse=imp, force_tick=tk)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        y = res["y_hist"][..., 0].cpu().numpy()
        if base is None:
            base = y
        print(f"rho {rho:4.2f} alpha {alpha:4.2f}: {ms:7.2f} ms, {res['iters'].sum().item() / (B * 838):5.2f} iterations per solve, "
              f"not converged {int((res['status'] != 0).sum())}, max |dy| vs first {np.abs(y - base).max():.1e}", flush=True)
