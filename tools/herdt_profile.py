"""Where does a Herdt rollout spend its time?  (host profile + device time of the tick launches)"""
import cProfile, os, pstats, sys, time
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
ctrl = ZMPController(cfg)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x0 = np.zeros((B, 3))
t0 = time.perf_counter(); res = ctrl.generate_com_trajectory_herdt_batched(x0, x0, v, st); torch.cuda.synchronize(); t1 = time.perf_counter()
print("first call (plan + timeline build + rollout)", round(t1 - t0, 3), "s")
t0 = time.perf_counter(); res = ctrl.generate_com_trajectory_herdt_batched(x0, x0, v, st); torch.cuda.synchronize(); t1 = time.perf_counter()
print("second call (rollout only)", round(t1 - t0, 3), "s; iters per joint solve", res["iters"].sum() / (B * 301))
pr = cProfile.Profile(); pr.enable()
res = ctrl.generate_com_trajectory_herdt_batched(x0, x0, v, st); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

# ---- build cost of a fresh batch object (what every new timeline group pays)
from mpc_bipedal import engine, herdt
padded = np.concatenate([st, np.repeat(st[-1:], cfg.horizon)])
for rep in range(2):
    t0 = time.perf_counter(); lay = herdt.timeline_layout(st, cfg.horizon); t1 = time.perf_counter()
    nb = ctrl.find_nb_steps(padded); t2 = time.perf_counter()
    plan = engine.HerdtPlan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.alpha, cfg.beta, cfg.gamma, cfg.foot_length, cfg.foot_width, cfg.foot_spread)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    tl = plan.timeline(lay.phase_of_row, lay.row_kind, lay.m); torch.cuda.synchronize(); t4 = time.perf_counter()
    pv = herdt.polygon_vertices(cfg.left_foot_polytope); t5 = time.perf_counter()
    print(f"rep {rep}: layout {t1-t0:.3f}  find_nb_steps {t2-t1:.3f}  plan {t3-t2:.3f}  timeline (device build) {t4-t3:.3f}  polygon {t5-t4:.3f}")
    plan.close()
