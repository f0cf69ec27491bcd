#!/usr/bin/env python3
"""Push resistance, strict vs non-strict (reference scripts/run_compare_resistance.py), batched.

The reference runs two walks (strict / non-strict) with one push of --F-ext at tick n//2 and plots the
ZMP-y against the bounds.  Here the same CoP bounds feed a SWEEP of B scenarios (BASELINE config 2:
F_ext ~ U(0, F_max), push tick ~ U{100..319}, seed 0) in one launch per mode; --sweep 1 reproduces the
reference's single pair of walks exactly.

    python scripts/run_compare_resistance.py                       # default.json, 4096 scenarios
    python scripts/run_compare_resistance.py --F-ext 600 --sweep 1 # the reference's experiment
"""

import argparse
import os
import time

import numpy as np

from _common import base_config, write_json
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal.generators import CoPGenerator


def main(argv=None):
    parser = argparse.ArgumentParser(description="strict vs non-strict push resistance (batched sweep)")
    parser.add_argument("--config", type=str)
    parser.add_argument("--F-ext", type=float, dest="F_ext")
    parser.add_argument("--output-dir", type=str, default="results")
    parser.add_argument("--sweep", type=int, default=4096, help="scenarios (1 = the reference's single push)")
    parser.add_argument("--F-max", type=float, dest="F_max", default=1000.0)
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--precision", choices=["fp64", "f16x3"], default="fp64")
    args = parser.parse_args(argv)

    config = base_config(args.config)
    if args.F_ext is not None:
        config.F_ext = args.F_ext
    config.add_force = True
    config.method = "wieber"
    z_max, z_min, _ = CoPGenerator(config).generate_cop_trajectory(save_footsteps=False, output_dir=args.output_dir)
    n = len(z_min)
    B = args.sweep
    if B == 1:
        F, tick = np.array([config.F_ext]), np.array([n // 2], np.int32)
    else:
        rng = np.random.default_rng(args.seed)
        F = rng.uniform(0.0, args.F_max, B)
        tick = rng.integers(100, min(320, n - 1), B).astype(np.int32)
    zero = np.zeros((B, 3))

    summary = {"scenarios": B, "n_steps": n, "horizon": config.horizon, "qp_solves_per_mode": int(B * 2 * (n - 1))}
    results = {}
    for label, strict in (("strict", True), ("non_strict", False)):
        config.strict = strict
        controller = ZMPController(config)
        if strict and args.precision == "f16x3":
            controller.solver_options = {"precision": 1}
        controller.generate_com_trajectory_batched(zero[:1], zero[:1], z_max, z_min, F_ext=F[:1], force_tick=tick[:1])
        t0 = time.perf_counter()
        res = controller.generate_com_trajectory_batched(zero, zero, z_max, z_min, F_ext=F, force_tick=tick)
        seconds = time.perf_counter() - t0
        results[label] = res
        summary[label] = {
            "seconds": seconds, "qp_solves_per_s": summary["qp_solves_per_mode"] / seconds,
            "max_zmp_violation_m": float(res["max_violation"].max()),
            "scenarios_with_violation": int((res["max_violation"] > 1e-9).sum()),
            "scenarios_fallen": int(res["fell"].sum()),
            "not_converged": int((res["status"] != 0).sum()),
        }
        print(f"{label:>10}: {seconds * 1e3:8.1f} ms  {summary[label]['qp_solves_per_s']:.3e} solves/s  "
              f"violations {summary[label]['scenarios_with_violation']}/{B}  fallen {summary[label]['scenarios_fallen']}/{B}")

    os.makedirs(args.output_dir, exist_ok=True)
    out = os.path.join(args.output_dir, "compare_resistance")
    np.savez_compressed(out + ".npz", F_ext=F, force_tick=tick, z_max=z_max, z_min=z_min,
                        zmp_y_strict=results["strict"]["zmp"][..., 1], zmp_y_non_strict=results["non_strict"]["zmp"][..., 1],
                        max_violation_strict=results["strict"]["max_violation"],
                        max_violation_non_strict=results["non_strict"]["max_violation"],
                        fell_strict=results["strict"]["fell"], fell_non_strict=results["non_strict"]["fell"])
    write_json(out + ".json", summary)
    try:
        import plotly.graph_objects as go
        t = np.arange(n) * config.dt
        worst = int(np.argmax(results["non_strict"]["max_violation"]))
        fig = go.Figure()
        for name, series in (("z_max y", z_max[:, 1]), ("z_min y", z_min[:, 1]),
                             ("ZMP y strict", results["strict"]["zmp"][worst, :, 1]),
                             ("ZMP y non-strict", results["non_strict"]["zmp"][worst, :, 1])):
            fig.add_trace(go.Scatter(x=t, y=series, name=name))
        fig.update_layout(title=f"scenario {worst}: F = {F[worst]:.0f} N at t = {tick[worst] * config.dt:.2f} s")
        fig.write_html(out + ".html")
    except ImportError:
        pass
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
