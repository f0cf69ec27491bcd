import os, sys
import numpy as np
REPO = "/root/repo"
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator
cfg = MPCConfig(); cfg.method = "herdt"
vx, vy, states = SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
v = np.stack([vx, vy], 1); st = np.array(states)
ctrl = ZMPController(cfg)
B = 8192
rng = np.random.default_rng(0)
x0 = rng.normal(size=(B, 3)) * [0.01, 0.02, 0.05]
vv = v[None] * rng.uniform(0.8, 1.2, (B, 1, 1))
res = ctrl.generate_com_trajectory_herdt_batched(x0, x0, vv, st, F_ext=rng.uniform(0, 300, B), to_numpy=False)
torch.cuda.synchronize()
print("iters per joint solve", res["iters"].sum().item() / (B * 301), "failed ticks", res["status"].sum().item() / B)
