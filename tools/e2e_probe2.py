"""Per-call wall time of ZMPController.generate_com_trajectory_batched, with a host profile of the slowest call."""
import cProfile, io, os, pstats, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from mpc_bipedal.controllers import ZMPController
B = 4096
cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
ctrl = ZMPController(cfg)
hx0 = np.zeros((B, 3)); hy0 = np.zeros((B, 3))
for _ in range(3):
    ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
worst = (0, None)
times = []
gpu_times = []
for i in range(16):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
    e0.record()
    res = ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
    e1.record()
    pr.disable(); dt = time.perf_counter() - t0
    e1.synchronize()
    gpu_times.append(round(e0.elapsed_time(e1), 1))
    times.append(round(dt * 1e3, 1))
    if dt > worst[0]:
        worst = (dt, pr)
print("wall ms:", times)
print("gpu  ms:", gpu_times)
s = io.StringIO(); pstats.Stats(worst[1], stream=s).sort_stats("tottime").print_stats(6); print(s.getvalue()[:1800])
