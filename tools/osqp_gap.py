#!/usr/bin/env python3
"""How far does the reference's ACTUAL output (OSQP stopped at cvxpy's eps 1e-5) sit from the exact minimiser the
oracle computes?  CPU probe, not product (imports oracle/).

Runs the default.json strict walk (BASELINE configs[0]) twice in closed loop: per-tick QPs solved exactly
(oracle.lipm_oracle.strict_axis_admm to 1e-13, the golden trajectory) and by the NumPy restatement of OSQP as cvxpy
calls it (oracle/osqp_restatement.py: non-condensed form with the auxiliary variable, eps_abs = eps_rel = 1e-5,
max_iter 10000, cold start every tick as the reference).  Reports the per-tick and the trajectory gaps.

    python tools/osqp_gap.py [--ticks 419] [--interval 50] [--out profiles/r02_osqp_restatement_gap.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, _p)
from oracle import lipm_oracle as orc  # noqa: E402
from oracle import osqp_restatement as osq  # noqa: E402
from mpc_bipedal.config import MPCConfig  # noqa: E402
from mpc_bipedal.generators import CoPGenerator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ticks", type=int, default=419)
    ap.add_argument("--interval", type=int, default=50)
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r02_osqp_restatement_gap.json"))
    args = ap.parse_args()
    cfg = MPCConfig.from_json(os.path.join(REPO, "configs", "default.json"))
    z_max, z_min, _ = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
    n, N = len(z_min), cfg.horizon
    p = orc.Problem.make(N, cfg.h, cfg.g, cfg.Q, cfg.R)
    hi = np.vstack([z_max, np.tile(z_max[-1:], (N, 1))])
    lo = np.vstack([z_min, np.tile(z_min[-1:], (N, 1))])
    settings = osq.Settings(adaptive_rho_interval=args.interval)
    impulse, force_tick = cfg.dt * cfg.F_ext / cfg.m, n // 2
    exact = [np.zeros(3), np.zeros(3)]
    approx = [np.zeros(3), np.zeros(3)]
    gap_state, gap_tick, iters, statuses, box = [], [], [], {}, 0.0
    t0 = time.time()
    ticks = min(args.ticks, n - 1)
    for i in range(ticks):
        for ax in (0, 1):
            hw, lw = hi[i + 1:i + 1 + N, ax], lo[i + 1:i + 1 + N, ax]
            exact[ax] = orc.strict_axis_admm(p, exact[ax], hw, lw)[0]
            # one-tick gap: both solvers from the SAME state (the OSQP walk's)
            same = orc.strict_axis_admm(p, approx[ax], hw, lw)[0]
            nxt, info = osq.wieber_strict_axis(p, approx[ax], hw, lw, settings)
            statuses[info["status"]] = statuses.get(info["status"], 0) + 1
            if nxt is None:
                raise SystemExit(f"tick {i} axis {ax}: restated OSQP reports {info['status']}")
            gap_tick.append(np.abs(nxt - same))
            iters.append(info["iters"])
            approx[ax] = nxt
            zmp = nxt @ p.C
            box = max(box, zmp - hw[0], lw[0] - zmp)
        if i == force_tick:
            exact[1] = exact[1] - np.array([0.0, impulse, 0.0])
            approx[1] = approx[1] - np.array([0.0, impulse, 0.0])
        gap_state.append([np.abs(approx[0] - exact[0]), np.abs(approx[1] - exact[1])])
        if i % 50 == 49:
            print(f"tick {i + 1}/{ticks}: {time.time() - t0:.0f} s, trajectory gap so far {np.max(np.array(gap_state)[:, :, 0]):.2e} m",
                  flush=True)
    gap_state, gap_tick = np.array(gap_state), np.array(gap_tick)
    out = {
        "what": "default.json strict walk in closed loop: restated OSQP (cvxpy defaults eps_abs = eps_rel = 1e-5, max_iter 10000, "
                "cold start per tick) against the exact per-tick minimiser",
        "ticks": ticks, "adaptive_rho_interval": args.interval, "statuses": statuses,
        "osqp_iterations_per_solve": {"mean": float(np.mean(iters)), "max": int(np.max(iters))},
        "one_tick_gap_next_state": {"pos_m": float(gap_tick[:, 0].max()), "vel": float(gap_tick[:, 1].max()),
                                    "acc": float(gap_tick[:, 2].max())},
        "trajectory_gap": {"x_pos_m": float(gap_state[:, 0, 0].max()), "y_pos_m": float(gap_state[:, 1, 0].max()),
                           "y_vel": float(gap_state[:, 1, 1].max()), "y_acc": float(gap_state[:, 1, 2].max())},
        "max_box_violation_of_realised_zmp_m": float(max(box, 0.0)),
        "seconds": time.time() - t0,
        "note": "parity unpinned: cvxpy / OSQP are not installed; this is the published algorithm restated, not the solver",
    }
    print(json.dumps(out, indent=1))
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
