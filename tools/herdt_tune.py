#!/usr/bin/env python3
"""Iterations and time of the Herdt walk batch against the ADMM penalty rho and the relaxation (design probe)."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator

cfg = MPCConfig(); cfg.method = "herdt"
vx, vy, states = SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
v = np.stack([vx, vy], 1); st = np.array(states)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
x0 = rng.normal(size=(B, 3)) * [0.01, 0.02, 0.05]
vv = v[None] * rng.uniform(0.8, 1.2, (B, 1, 1))
F = rng.uniform(0, 300, B)
base = None
for rho in (4.0, 6.0, 8.0, 12.0, 16.0):
    for relax in (1.6, 1.7, 1.8):
        ctrl = ZMPController(cfg)
        ctrl.herdt_options = {"rho": rho, "relax": relax}
        ctrl.generate_com_trajectory_herdt_batched(x0[:8], x0[:8], vv[:8], st, F_ext=F[:8], to_numpy=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = ctrl.generate_com_trajectory_herdt_batched(x0, x0, vv, st, F_ext=F, to_numpy=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        y = res["y_hist"][:, :250].cpu().numpy()
        if base is None:
            base = y
        print(f"rho {rho:5.1f} relax {relax:.2f}: {dt:.3f} s, {res['iters'].sum().item() / (B * 301):.1f} iterations per joint solve, "
              f"failed ticks per walk {res['status'].sum().item() / B:.1f}, max |dy| vs first setting {np.abs(y - base).max():.2e}", flush=True)
        ctrl._herdt_cache[1].plan.close() if hasattr(ctrl._herdt_cache[1].plan, "close") else None
