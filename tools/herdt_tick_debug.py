"""Print GPU-vs-oracle errors of single Herdt ticks (debug aid for tests/test_herdt.py)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal import engine, herdt
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.models.lipm_model import discrete_model
from oracle import herdt_oracle as ho
g = np.load(os.path.join(REPO, "tests/golden/ref_herdt_helpers.npz"))
cfg = MPCConfig(); N = cfg.horizon
A, Bm, _ = discrete_model(cfg.dt, cfg.h, cfg.g)
states = g["class_defaults/states"]; v_ref = np.stack([g["class_defaults/vx"], g["class_defaults/vy"]], 1)
polys = {s: (g[f"polytope/{s}/A"], g[f"polytope/{s}/b"]) for s in ("left", "right")}
lay = herdt.timeline_layout(states, N)
plan = engine.HerdtPlan(N, cfg.h, cfg.g, cfg.dt, cfg.alpha, cfg.beta, cfg.gamma, cfg.foot_length, cfg.foot_width, cfg.foot_spread)
mats = ho.prediction_matrices(N, cfg.dt, cfg.h, cfg.g)
ticks = [0, 60, 100, 251, 260]
tl = plan.timeline(lay.phase_of_row[ticks], lay.row_kind[ticks], lay.m[ticks])
polygons = (herdt.polygon_vertices(cfg.left_foot_polytope), herdt.polygon_vertices(cfg.right_foot_polytope))
v_ext = np.vstack([v_ref, np.repeat(v_ref[-1:], N, 0)])
rng = np.random.default_rng(0); dev = plan.device
for slot, i in enumerate(ticks):
    side = int(lay.foot_side[i])
    fc = np.array([[0.3 * (i // 50), 0.1 if side == 0 else -0.1]] * 3)
    x = rng.normal(size=(3, 3)) * [0.02, 0.05, 0.1] + [fc[0, 0], 0.2 if 50 < i < 250 else 0.0, 0.0]
    y = rng.normal(size=(3, 3)) * [0.01, 0.03, 0.1] + [fc[0, 1] * 0.3, 0.0, 0.0]
    vref = torch.as_tensor(np.vstack([np.zeros((slot + 1, 2)), v_ext[i + 1: i + 1 + N]])).to(dev)
    out = tl.tick(slot, torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev), torch.as_tensor(fc).to(dev), vref, side,
                  polygons[side] if lay.m[i] > 0 else None, want_solution=True)
    for b in range(3):
        xn, yn, fx1, fy1, u, lam, info = ho.predict_joint(cfg, A, Bm, x[b], y[b], v_ext[i + 1: i + 1 + N], fc[b, 0], fc[b, 1],
                                                          int(lay.current[i]), lay.codes[i + 1: i + 1 + N],
                                                          "left" if side == 0 else "right", polys["left" if side == 0 else "right"], mats)
        ex = np.abs(out["next_x"][b].cpu().numpy() - xn); ey = np.abs(out["next_y"][b].cpu().numpy() - yn)
        f = out["f1"][b].cpu().numpy()
        print(f"tick {i} b{b} m={lay.m[i]} side={side} status={out['status'][2*b:2*b+2].tolist()} iters={out['iters'][2*b:2*b+2].tolist()} "
              f"ex={ex.max():.2e} ey={ey.max():.2e} f1err={(abs(f[0]-fx1) if fx1 is not None else 0):.2e},{(abs(f[1]-fy1) if fy1 is not None else 0):.2e} kkt={info['kkt']['ok']}")
