"""Break down one end-to-end call of ZMPController.generate_com_trajectory_batched (NumPy in, NumPy out)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import cProfile, pstats
import numpy as np, torch
import bench
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal import engine
B = 4096
cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
ctrl = ZMPController(cfg)
hx0 = np.zeros((B, 3)); hy0 = np.zeros((B, 3))
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("full call", i, round((t1 - t0) * 1e3, 1), "ms")
pr = cProfile.Profile(); pr.enable()
res = ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
