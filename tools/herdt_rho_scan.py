#!/usr/bin/env python3
"""Iterations per joint solve and time of the class-default Herdt walk (2048 scenarios) against rho / relaxation."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator
cfg = MPCConfig(); cfg.method = "herdt"; cfg.speed_generation = "classic"
vx, vy, states = SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
v, st = np.stack([vx, vy], 1), np.array(states)
B = 2048
rng = np.random.default_rng(0)
scale = rng.uniform(0.8, 1.2, (B, 1, 1))
zero = np.zeros((B, 3))
for rho in [float(a) for a in sys.argv[1:]] or [2.0, 4.0, 6.0, 10.0, 16.0]:
    for relax in (1.8, 1.6):
        ctrl = ZMPController(cfg); ctrl.herdt_options = {"rho": rho, "relax": relax}
        ctrl.generate_com_trajectory_herdt_batched(zero[:8], zero[:8], v, st, to_numpy=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = ctrl.generate_com_trajectory_herdt_batched(zero, zero, v[None] * scale, st, to_numpy=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ticks = len(st) - 1
        print(f"rho {rho:5.1f} relax {relax}: {dt*1e3:7.1f} ms, {int(res['iters'].sum())/(B*ticks):6.1f} it/solve, fallback ticks/plan {int(res['status'].sum())/B:.1f}", flush=True)
        ctrl._herdt_cache[1].plan.close()
