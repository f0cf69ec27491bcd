"""Device-side phase timing (CUDA events) of the end-to-end push sweep: upload + rollout / post-processing /
device-to-host.  This probe found the 0.1-1 s stalls of the result read-back that came from the stream-ordered
memory pool being trimmed at every synchronisation (fixed in lipm_plan_create: release threshold raised)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from mpc_bipedal import engine
B = 4096
cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
hx0 = np.zeros((B, 3))
pool = engine.HostPool(); cache = None
c = torch.as_tensor(np.array([1.0, 0.0, -cfg.h / cfg.g])).cuda()
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for i in range(24):
    impulse = cfg.dt * F / cfg.m
    e0 = ev()
    if cache is not None:
        for name in ("x_hist", "y_hist", "flags"): cache[name].zero_()
    out = plan.rollout(engine.MODE_STRICT, hx0, hx0, z_min, z_max, impulse=impulse, force_tick=tick, out=cache)
    cache = out
    e1 = ev()
    x_hist, y_hist = out["x_hist"], out["y_hist"]
    com = torch.stack([x_hist[..., 0], y_hist[..., 0]], dim=-1)
    zmp = torch.stack([x_hist @ c, y_hist @ c], dim=-1)
    lean = (com - zmp).abs().amax(dim=(1, 2))
    res = {"com": com, "x_hist": x_hist, "y_hist": y_hist, "zmp": zmp, "flags": out["flags"],
           "max_violation": out["max_violation"].amax(dim=1), "fell": lean > 0.4, "status": out["status"].sum(dim=1), "iters": out["iters"]}
    e2 = ev()
    host = engine.to_host(res, pool)
    e3 = ev(); e3.synchronize()
    print(f"call {i:2d}: upload+rollout {e0.elapsed_time(e1):7.1f} | post-processing {e1.elapsed_time(e2):6.1f} | to_host {e2.elapsed_time(e3):6.1f} | iters {int(host['iters'].sum())}")
