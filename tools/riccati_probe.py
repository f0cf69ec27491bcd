"""Active-set (Riccati) kernel against the ADMM kernel on the BASELINE push sweep: agreement, time per warp shape.

    python tools/riccati_probe.py [B] [--lanes 4,8,16,32] [--no-admm]
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


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    lanes = [0, 4, 8, 16, 32]
    for i, a in enumerate(sys.argv):
        if a == "--lanes":
            lanes = [int(v) for v in sys.argv[i + 1].split(",")]
    cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
    n = len(z_min)
    dev = torch.device("cuda", 0)
    imp = torch.tensor(cfg.dt * F / cfg.m, device=dev)
    ft = torch.tensor(tick, device=dev)
    zmin = torch.tensor(z_min, device=dev); zmax = torch.tensor(z_max, device=dev)
    x0 = torch.zeros(B, 3, dtype=torch.float64, device=dev); y0 = torch.zeros_like(x0)
    solves = B * 2 * (n - 1)
    out = {"B": B, "n": n, "N": cfg.horizon}
    plan_r = engine.LipmPlan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
    res_r = {}
    run_r = lambda: plan_r.rollout(1, x0, y0, zmin, zmax, impulse=imp, force_tick=ft, out=res_r, want_readout=True)
    for l in lanes:
        plan_r.set_riccati_lanes(l)
        ms = timed(run_r)
        out[f"active_set_lanes{l}"] = {"ms": ms, "M_solves_per_s": solves / ms / 1e3}
        print(json.dumps({f"lanes{l}": out[f"active_set_lanes{l}"]}), flush=True)
    plan_r.set_riccati_lanes(0)
    run_r(); torch.cuda.synchronize()
    it = res_r["iters"].cpu().numpy()
    out["sweeps_per_solve"] = {"x": float(it[:, 0].sum() / (B * (n - 1))), "y": float(it[:, 1].sum() / (B * (n - 1)))}
    out["not_converged"] = int((res_r["status"].cpu().numpy() != 0).sum())
    if "--no-admm" not in sys.argv:
        plan_a = engine.LipmPlan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, solver="admm")
        res_a = {}
        run_a = lambda: plan_a.rollout(1, x0, y0, zmin, zmax, impulse=imp, force_tick=ft, out=res_a, want_readout=True)
        ms = timed(run_a, reps=1)
        out["admm"] = {"ms": ms, "M_solves_per_s": solves / ms / 1e3}
        xa, ya = res_a["x_hist"].cpu().numpy(), res_a["y_hist"].cpu().numpy()
        xr, yr = res_r["x_hist"].cpu().numpy(), res_r["y_hist"].cpu().numpy()
        d = np.maximum(np.abs(xa - xr).max(axis=(1, 2)), np.abs(ya - yr).max(axis=(1, 2)))
        scale = np.maximum(1.0, np.abs(ya).max(axis=(1, 2)))
        out["vs_admm"] = {"max_abs": float(d.max()), "max_rel": float((d / scale).max()), "median": float(np.median(d)),
                          "flags_differ": int((res_a["flags"].cpu().numpy() != res_r["flags"].cpu().numpy()).sum()),
                          "zmp_max_abs": float(np.abs(res_a["zmp"].cpu().numpy() - res_r["zmp"].cpu().numpy()).max()),
                          "viol_active_set": float(res_r["max_violation"].max()), "viol_admm": float(res_a["max_violation"].max())}
    print(json.dumps(out))
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", f"riccati_probe_b{B}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
