#!/usr/bin/env python3
"""What does the reference do on the final-stance ticks of a Herdt walk?  CPU probe, not product (imports oracle/).

On the last ~50 ticks of every walk (the final STANDING phase enters the preview window, `zmp_controller.py:715-771`)
neither the oracle's exact solver nor the GPU ADMM converge: the QP the reference builds there is infeasible in
effect.  The reference hands it to cvxpy -> OSQP (`:792-794`) and acts on OSQP's verdict: "solved" or "maximum
iterations reached" -> `u.value` is set and applied; "primal infeasible" -> `u.value is None` -> fallback to zero jerk
and the interpolated footstep (`:796-804`).  This probe walks the class-default timeline (run_herdt_method.py:53) with
the oracle and asks the NumPy restatement of OSQP (oracle/osqp_restatement.py, cvxpy defaults) for its verdict on every
tick from `--from-tick` on, applying the reference's rule, and compares with what the CUDA engine's rule would do
(converged -> apply; stalled / max_iter -> fallback).

    python tools/herdt_final_stance.py [--from-tick 230] [--out profiles/r02_herdt_final_stance.json]
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
from oracle import herdt_oracle as ho  # noqa: E402
from oracle import osqp_restatement as osq  # noqa: E402
from mpc_bipedal.config import MPCConfig  # noqa: E402
from mpc_bipedal.models.lipm_model import discrete_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from-tick", type=int, default=230)
    ap.add_argument("--intervals", type=str, default="50")
    ap.add_argument("--max-ticks", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r02_herdt_final_stance.json"))
    args = ap.parse_args()
    g = np.load(os.path.join(REPO, "tests", "golden", "ref_herdt_helpers.npz"))
    cfg = MPCConfig()
    cfg.method = "herdt"
    N = cfg.horizon
    A_model, B_model, _ = discrete_model(cfg.dt, cfg.h, cfg.g)
    B_vec = np.asarray(B_model, float).reshape(3)
    states = g["class_defaults/states"]
    v_ref = np.stack([g["class_defaults/vx"], g["class_defaults/vy"]], 1)
    nb = g["class_defaults/nb_steps"]
    polys = {s: (g[f"polytope/{s}/A"], g[f"polytope/{s}/b"]) for s in ("left", "right")}
    n = len(v_ref)
    v_ext = np.vstack([v_ref, np.repeat(v_ref[-1:], N, axis=0)])
    s_ext = np.concatenate([states, np.repeat(states[-1:], N)])
    mats = ho.prediction_matrices(N, cfg.dt, cfg.h, cfg.g)
    intervals = [int(s) for s in args.intervals.split(",")]

    x, y = np.zeros(3), np.zeros(3)
    x_fc, y_fc = 0.0, float(cfg.foot_spread)
    x_air, y_air = x_fc, y_fc
    side, current = "left", int(s_ext[0])
    rows = []
    t0 = time.time()
    ticks = n - 1 if not args.max_ticks else min(n - 1, args.max_ticks)
    for i in range(ticks):
        qp = ho.build_qp(cfg, x, y, v_ext[i + 1: i + 1 + N], x_fc, y_fc, current, s_ext[i + 1: i + 1 + N], side, polys[side], mats)
        u, lam, info = ho.solve_qp(qp)
        certified = bool(info["kkt"]["ok"])
        m = qp["m"]
        row = {"tick": i, "state": current, "m": m, "oracle_certified": certified}
        applied = u
        if i >= args.from_tick:
            verdicts = {}
            for interval in intervals:
                res = osq.solve(*osq.herdt_problem(qp), osq.Settings(adaptive_rho_interval=interval))
                verdicts[str(interval)] = {"status": res["status"], "iters": res["iters"], "prim_res": res["prim_res"],
                                           "dual_res": res["dual_res"]}
                if interval == intervals[0]:
                    first = res
            row["osqp"] = verdicts
            status = first["status"]
            row["reference_action"] = "apply" if status in ("solved", "max_iter") else "fallback"
            row["engine_action"] = "apply" if certified else "fallback"
            if certified and status == "solved":
                row["osqp_vs_exact_J0"] = [abs(float(first["x"][0] - u[0])), abs(float(first["x"][N + m] - u[N + m]))]
            if not certified:
                # follow the reference: apply OSQP's iterate when it has one, else the fallback of :796-804
                if status in ("solved", "max_iter"):
                    applied = first["x"]
                else:
                    applied = np.zeros_like(u)
                    if m > 0:
                        applied[N], applied[N + m + N] = x_air, y_air
        elif not certified:
            row["note"] = "oracle did not certify before --from-tick"
        rows.append(row)
        fx1 = applied[N] if m > 0 else None
        fy1 = applied[N + m + N] if m > 0 else None
        x = A_model @ x + B_vec * applied[0]
        y = A_model @ y + B_vec * applied[N + m]
        if fx1 is not None:
            x_air += (1 / nb[i][0]) * (fx1 - x_air)
            y_air += (1 / nb[i][0]) * (fy1 - y_air)
        if s_ext[i + 1] != current and current == ho.SINGLE_SUPPORT:
            side = "left" if side == "right" else "right"
            x_fc, y_fc = (float(fx1), float(fy1)) if fx1 is not None else (x_air, y_air)
            x_air, y_air = x_fc, y_fc
        if s_ext[i + 1] != current:
            current = int(s_ext[i + 1])
        if i >= args.from_tick:
            print(f"tick {i}: certified={certified} osqp={ {k: (v['status'], v['iters']) for k, v in row['osqp'].items()} } "
                  f"[{time.time() - t0:.0f} s]", flush=True)
    tail = [r for r in rows if "osqp" in r]
    summary = {
        "what": "class-default Herdt walk (run_herdt_method.py:53), oracle-driven; OSQP restatement verdict per tick from --from-tick on",
        "ticks": ticks, "from_tick": args.from_tick, "adaptive_rho_intervals": intervals,
        "uncertified_ticks": [r["tick"] for r in rows if not r["oracle_certified"]],
        "reference_fallback_ticks": [r["tick"] for r in tail if r["reference_action"] == "fallback"],
        "engine_fallback_ticks": [r["tick"] for r in tail if r["engine_action"] == "fallback"],
        "disagreements": [r["tick"] for r in tail if r["reference_action"] != r["engine_action"]],
        "verdict_depends_on_interval": [r["tick"] for r in tail if len({v["status"] for v in r["osqp"].values()}) > 1],
        "final_com": [float(x[0]), float(y[0])],
        "rows": tail, "seconds": time.time() - t0,
        "note": "parity unpinned: cvxpy / OSQP are not installed; the verdicts come from the published algorithm restated",
    }
    print(json.dumps({k: v for k, v in summary.items() if k != "rows"}, indent=1))
    with open(args.out, "w") as fh:
        json.dump(summary, fh, indent=1)


if __name__ == "__main__":
    main()
