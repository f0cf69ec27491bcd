"""Row-parallel first stage of the active-set solver: time and agreement with the thread-per-column kernel.

    python tools/rowpar_probe.py [B]
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
    x0 = torch.zeros(B, 3, dtype=torch.float64, device=dev); y0 = torch.zeros_like(x0)
    solves = B * 2 * (n - 1)
    plan = engine.LipmPlan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
    out = {"B": B}
    for name, Fs in (("push_sweep", F), ("no_push", np.zeros_like(F))):
        imp = torch.tensor(cfg.dt * Fs / cfg.m, device=dev); ft = torch.tensor(tick, device=dev)
        res = {}
        for on in (1, 0):
            plan.set_row_parallel(bool(on))
            r = {}
            ms = timed(lambda: plan.rollout(1, x0, y0, zmin, zmax, impulse=imp, force_tick=ft, out=r, want_readout=True))
            st = plan.stats()
            res[on] = {k: v.clone() for k, v in r.items() if torch.is_tensor(v)}
            out[f"{name}_rowpar{on}"] = {"ms": ms, "M_solves_per_s": solves / ms / 1e3, "sweeps": st["group_iterations"], "columns": st["units"],
                                             "parked": st["debug_cycles_contraction"], "handed_back": st["debug_cycles_iteration"]} if on else {"ms": ms}
            print(json.dumps({f"{name}_rowpar{on}": out[f"{name}_rowpar{on}"]}), flush=True)
        a, b = res[1], res[0]
        d = max(float((a["x_hist"] - b["x_hist"]).abs().max()), float((a["y_hist"] - b["y_hist"]).abs().max()))
        out[f"{name}_agreement"] = {"max_abs_hist": d, "flag_bytes_differ": int((a["flags"] != b["flags"]).sum()),
                                    "zmp_max_abs": float((a["zmp"] - b["zmp"]).abs().max()),
                                    "iters_rowpar": int(a["iters"].sum()), "iters_riccati": int(b["iters"].sum()),
                                    "status_nonzero": int((a["status"] != 0).sum()), "viol": float(a["max_violation"].max())}
        print(json.dumps({f"{name}_agreement": out[f"{name}_agreement"]}), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", f"rowpar_probe_b{B}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
