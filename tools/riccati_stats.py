"""Active-set kernel on the BASELINE push sweep: where the stages go (tables / general recursion / leading block),
and the serial time of the heaviest columns alone (the critical path of the launch).

    python tools/riccati_stats.py [B]
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
    cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
    n = len(z_min)
    dev = torch.device("cuda", 0)
    zmin = torch.tensor(z_min, device=dev); zmax = torch.tensor(z_max, device=dev)
    plan = engine.LipmPlan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
    N = cfg.horizon

    def run(sel, lanes=0):
        imp = torch.tensor(cfg.dt * F[sel] / cfg.m, device=dev)
        ft = torch.tensor(tick[sel], device=dev)
        b = len(sel)
        x0 = torch.zeros(b, 3, dtype=torch.float64, device=dev); y0 = torch.zeros_like(x0)
        res = {}
        plan.set_riccati_lanes(lanes)
        fn = lambda: plan.rollout(1, x0, y0, zmin, zmax, impulse=imp, force_tick=ft, out=res, want_readout=True)
        ms = timed(fn)
        st = plan.stats()
        it = res["iters"].cpu().numpy()
        return ms, st, it, res

    allsel = np.arange(B)
    ms, st, it, res = run(allsel)
    sweeps = st["group_iterations"]
    gen, tab = st["debug_cycles_contraction"], st["debug_cycles_iteration"]
    out = {"B": B, "ms": ms, "sweep_pairs": sweeps, "stage_share": {"general": gen / (sweeps * N), "table": tab / (sweeps * N),
                                                                  "leading_block": 1 - (gen + tab) / (sweeps * N)}}
    print(json.dumps(out), flush=True)
    fell = res["outcome"]["fell"].cpu().numpy() if isinstance(res.get("outcome"), dict) else None
    # heaviest / lightest columns alone
    order = np.argsort(-it[:, 1])
    for name, sel in (("heaviest64", order[:64]), ("heaviest8", order[:8]), ("lightest64", order[-64:]), ("median64", order[B // 2 - 32:B // 2 + 32])):
        for lanes in (1, 8):
            ms_s, st_s, it_s, _ = run(np.sort(sel), lanes)
            sw = st_s["group_iterations"]
            rec = {"subset": name, "lanes": lanes, "ms": ms_s, "sweeps_y_mean": float(it_s[:, 1].mean()), "sweeps_x_mean": float(it_s[:, 0].mean()),
                   "general_share": st_s["debug_cycles_contraction"] / (sw * N), "table_share": st_s["debug_cycles_iteration"] / (sw * N)}
            print(json.dumps(rec), flush=True)
    # x axis only / y axis only cannot be separated in a walk; no-push walks give the x-like cost on both axes
    Fz = np.zeros_like(F)
    imp = torch.tensor(cfg.dt * Fz / cfg.m, device=dev); ft = torch.tensor(tick, device=dev)
    x0 = torch.zeros(B, 3, dtype=torch.float64, device=dev); y0 = torch.zeros_like(x0)
    res = {}
    plan.set_riccati_lanes(0)
    ms0 = timed(lambda: plan.rollout(1, x0, y0, zmin, zmax, impulse=imp, force_tick=ft, out=res, want_readout=True))
    st0 = plan.stats()
    print(json.dumps({"no_push_ms": ms0, "sweeps": st0["group_iterations"], "general_share": st0["debug_cycles_contraction"] / (st0["group_iterations"] * N)}))


if __name__ == "__main__" and "--wide" not in sys.argv:
    main()


def wide_box():
    """Both axes without any active row (y box widened): the cost of a walk that only runs table stages."""
    B = 4096
    cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
    dev = torch.device("cuda", 0)
    z_min = z_min.copy(); z_max = z_max.copy()
    z_min[:, 1] -= 10.0; z_max[:, 1] += 10.0
    zmin = torch.tensor(z_min, device=dev); zmax = torch.tensor(z_max, device=dev)
    plan = engine.LipmPlan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
    for b, lanes in ((8, 1), (64, 1), (4096, 8), (4096, 16), (4096, 32), (32768, 32)):
        x0 = torch.zeros(b, 3, dtype=torch.float64, device=dev); y0 = torch.zeros_like(x0)
        res = {}
        plan.set_riccati_lanes(lanes)
        ms = timed(lambda: plan.rollout(1, x0, y0, zmin, zmax, out=res, want_readout=True))
        st = plan.stats()
        sw = st["group_iterations"]
        print(json.dumps({"wide_box": True, "B": b, "lanes": lanes, "ms": ms, "sweeps": sw, "general_share": st["debug_cycles_contraction"] / (sw * cfg.horizon),
                          "M_solves_per_s": b * 2 * (len(z_min) - 1) / ms / 1e3}), flush=True)


if __name__ == "__main__" and "--wide" in sys.argv:
    wide_box()
