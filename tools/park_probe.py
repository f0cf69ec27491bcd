"""What bounds the parked stage of the active-set solver on the push sweep: warp shape x chunks, and the latency floor
(the heaviest scenarios alone, one column per warp).

    python tools/park_probe.py [B]
"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"))
import bench  # noqa: E402
from mpc_bipedal import engine  # noqa: E402
from tools.riccati_probe import timed  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
    n = len(z_min)
    dev = torch.device("cuda", 0)
    zmin = torch.tensor(z_min, device=dev); zmax = torch.tensor(z_max, device=dev)
    plan = engine.LipmPlan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
    out = {"B": B}

    def run(sel, r):
        b = len(sel)
        x0 = torch.zeros(b, 3, dtype=torch.float64, device=dev); y0 = torch.zeros_like(x0)
        imp = torch.tensor(cfg.dt * F[sel] / cfg.m, device=dev); ft = torch.tensor(tick[sel], device=dev)
        return lambda: plan.rollout(1, x0, y0, zmin, zmax, impulse=imp, force_tick=ft, out=r, want_readout=True)

    allsel = np.arange(B)
    for chunks in (1, 4):
        plan.set_walk_chunks(chunks)
        for lanes in (0, 1, 2, 4, 8):
            plan.set_riccati_lanes(lanes)
            r = {}
            ms = timed(run(allsel, r), reps=3)
            out[f"chunks{chunks}_lanes{lanes}"] = ms
            print(json.dumps({f"chunks{chunks}_lanes{lanes}": ms}), flush=True)
    plan.set_walk_chunks(0); plan.set_riccati_lanes(0)
    r = {}
    run(allsel, r)(); torch.cuda.synchronize()
    it = r["iters"].cpu().numpy().reshape(B, -1).sum(axis=1)
    order = np.argsort(-it)
    for k in (1, 16, 64, 256, 1024):
        sel = np.sort(order[:k])
        for lanes in (1, 4):
            plan.set_riccati_lanes(lanes)
            rr = {}
            ms = timed(run(sel, rr), reps=3)
            out[f"heaviest{k}_lanes{lanes}"] = ms
            print(json.dumps({f"heaviest{k}_lanes{lanes}": ms, "sweeps": int(it[sel].sum())}), flush=True)
    sel = np.sort(order[-1024:])
    plan.set_riccati_lanes(0)
    out["lightest1024"] = timed(run(sel, {}), reps=3)
    print(json.dumps({"lightest1024": out["lightest1024"]}), flush=True)
    with open(os.path.join(REPO, "gpurun_out", f"park_probe_b{B}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
