"""Per-call wall and device time of the end-to-end push sweep (NumPy in, NumPy out), optionally at a raised
scheduling priority -- to separate what the engine costs from what a busy shared host adds."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from mpc_bipedal.controllers import ZMPController
if len(sys.argv) > 1 and sys.argv[1] == "nice":
    try:
        print("nice ->", os.nice(-15))
    except OSError as exc:
        print("cannot raise priority:", exc)
print("loadavg", os.getloadavg(), "cpus", len(os.sched_getaffinity(0)))
B = 4096
cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
ctrl = ZMPController(cfg)
hx0 = np.zeros((B, 3)); hy0 = np.zeros((B, 3))
for _ in range(3):
    ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
wall, gpu = [], []
for i in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    res = ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
    e1.record(); wall.append(round((time.perf_counter() - t0) * 1e3, 1))
    e1.synchronize(); gpu.append(round(e0.elapsed_time(e1), 1))
print("wall ms:", wall)
print("gpu  ms:", gpu)
print("median wall", np.median(wall), "mean wall", np.mean(wall))
