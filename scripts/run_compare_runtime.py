#!/usr/bin/env python3
"""Runtime vs horizon: Wieber strict, Wieber non-strict and Herdt (reference scripts/run_compare_runtime.py).

The reference loops N = 10, 20, ..., 300 (`:139`) and times ONE control tick per method -- x-solve + y-solve through
`predict_wieber_axis` (`:36-73`: 3 warm-ups, 10 timed calls, mean) and one `predict_herdt_joint` (`:76-128`: median;
that leg raises TypeError in the reference because the call lacks the `idx` argument) -- then plots the
three curves.  This driver keeps the loop, the three legs and the protocol, and reports per horizon

* `api_ms`    : the reference's own measurement on the mirrored API: wall clock around the SAME calls
                (`ZMPController.predict_wieber_axis` twice / `predict_herdt_joint` once, B = 1, NumPy in and out);
* `batch_ms`  : one tick for `--batch` independent scenarios in one launch (CUDA events), with NON-trivial inputs
                (the reference's x = 0, +-0.1 bounds give the trivial optimum u = 0): BASELINE configs[2];
* `cpu_ms`    : the non-strict tick on ONE host core with NumPy, matrices cached ("fair NumPy"); `cpu_as_reference_ms`
                rebuilds Px / Pu by the Python double loop and inverts per call as the reference does (BASELINE.md
                section 4, items 1 and 3).  The reference's strict / Herdt legs need cvxpy + OSQP (not installed).

    python scripts/run_compare_runtime.py                       # N = 10..300 step 10, batch 16384
    python scripts/run_compare_runtime.py --horizons 100 150 300 600 --batch 16384 --precision f16x3
"""

import argparse
import json
import os
import sys
import time

import numpy as np

from _common import REPO, write_json

sys.path.insert(0, os.path.join(REPO, "tools"))
import horizon_scaling  # noqa: E402  (tick_inputs: the seeded non-trivial states / bounds of BASELINE configs[2])
from mpc_bipedal.config import MPCConfig  # noqa: E402
from mpc_bipedal.controllers import ZMPController  # noqa: E402
from mpc_bipedal.generators.cop_generator import State  # noqa: E402

WARMUPS, TIMED = 3, 10          # reference :61-66


def _prepare_dummy_bounds(horizon: int, n_steps: int = 100):
    """The reference's +-0.1 bounds, padded by `horizon` rows (`:21-33`)."""
    z_max = np.vstack([np.ones((n_steps, 2)) * 0.1, np.ones((horizon, 2)) * 0.1])
    return z_max, -z_max


def measure_wieber_iteration_time(config: MPCConfig, n_steps: int = 100) -> float:
    """Reference protocol (`:36-73`): mean wall-clock time of one tick = two `predict_wieber_axis` calls, B = 1."""
    controller = ZMPController(config)
    z_max_ext, z_min_ext = _prepare_dummy_bounds(config.horizon, n_steps)
    N = config.horizon

    def run_once(x_state, y_state):
        x_state = controller.predict_wieber_axis(x_state, N, z_max_ext[1:1 + N, 0:1], z_min_ext[1:1 + N, 0:1])
        y_state = controller.predict_wieber_axis(y_state, N, z_max_ext[1:1 + N, 1:2], z_min_ext[1:1 + N, 1:2])
        return x_state, y_state

    x_state, y_state = np.zeros((3, 1)), np.zeros((3, 1))
    for _ in range(WARMUPS):
        x_state, y_state = run_once(x_state, y_state)
    times = []
    for _ in range(TIMED):
        start = time.perf_counter()
        run_once(np.zeros((3, 1)), np.zeros((3, 1)))
        times.append(time.perf_counter() - start)
    return float(np.mean(times))


def measure_herdt_iteration_time(config: MPCConfig, n_steps: int = 100) -> float:
    """Reference protocol (`:76-128`): median wall-clock time of one `predict_herdt_joint`, all-STANDING timeline,
    zero reference speed, right foot -- with the `idx` argument the reference's own call forgets."""
    controller = ZMPController(config)
    N = config.horizon
    v_ref_ext = np.zeros((n_steps + N, 2))
    state_ref_ext = np.array([State.STANDING] * (n_steps + N))
    nb = controller.find_nb_steps(state_ref_ext)
    x_fc, y_fc = np.array([[0.0]]), np.array([[config.foot_spread]])

    def run_once():
        controller.predict_herdt_joint(np.zeros((3, 1)), np.zeros((3, 1)), v_ref_ext[1:1 + N], x_fc, y_fc, state_ref_ext[0],
                                       state_ref_ext[1:1 + N], N, nb[0], x_fc.copy(), y_fc.copy(), "right", 0)

    for _ in range(WARMUPS):
        run_once()
    times = []
    for _ in range(TIMED):
        start = time.perf_counter()
        run_once()
        times.append(time.perf_counter() - start)
    return float(np.median(times))


def batched_wieber_tick(config: MPCConfig, batch: int, precision: int):
    """One tick (x and y solve) of `batch` scenarios = 2 * batch axis problems in one `lipm_solve_tick`, CUDA events."""
    import torch

    from mpc_bipedal import engine

    N = config.horizon
    plan = engine.get_plan(N, config.h, config.g, config.dt, config.Q, config.R, precision=precision if config.strict else 0)
    x, lo, hi = horizon_scaling.tick_inputs(N, 2 * batch)
    dev = plan.device
    xd, lod, hid = (torch.as_tensor(a).to(dev) for a in (x, lo, hi))
    mode = engine.MODE_STRICT if config.strict else engine.MODE_UNCONSTRAINED
    for _ in range(WARMUPS):
        res = plan.solve_tick(mode, xd, lod, hid)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(TIMED):
        res = plan.solve_tick(mode, xd, lod, hid)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / TIMED
    return ms, float(res["iters"].sum().item()) / (2 * batch), int((res["status"] != 0).sum().item())


def batched_herdt_tick(config: MPCConfig, batch: int):
    """One joint tick of `batch` scenarios in the middle of a walk (tick 60 of the gait timeline: single support,
    free footsteps in the window) through `herdt_tick`, CUDA events, median like the reference's Herdt leg."""
    import torch

    from mpc_bipedal import engine, herdt
    from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator

    N = config.horizon
    walk = MPCConfig(horizon=N, method="herdt", speed_generation="classic")
    vx, vy, states = SpeedTrajectoryGenerator(walk).generate_speed_and_state(save_footsteps=False)
    lay = herdt.timeline_layout(states, N)
    i = min(60 * N // 150 + 1, len(lay.m) - 1)
    plan = engine.HerdtPlan(N, walk.h, walk.g, walk.dt, walk.alpha, walk.beta, walk.gamma, walk.foot_length, walk.foot_width,
                            walk.foot_spread)
    timeline = plan.timeline(lay.phase_of_row[i:i + 1], lay.row_kind[i:i + 1], lay.m[i:i + 1])
    dev = plan.device
    rng = np.random.default_rng(5)
    side = int(lay.foot_side[i])
    fc = np.tile([[0.0, walk.foot_spread if side == 0 else -walk.foot_spread]], (batch, 1))
    x = rng.normal(size=(batch, 3)) * [0.02, 0.05, 0.1] + [0.0, 0.2, 0.0]
    y = rng.normal(size=(batch, 3)) * [0.01, 0.03, 0.1] + [fc[0, 1] * 0.3, 0.0, 0.0]
    v = np.stack([vx, vy], 1)
    v_ext = np.vstack([v, np.repeat(v[-1:], N, 0)])
    vref = torch.as_tensor(np.vstack([np.zeros((1, 2)), v_ext[i + 1: i + 1 + N]])).to(dev)
    xd, yd, fcd = (torch.as_tensor(a).to(dev) for a in (x, y, fc))
    poly = herdt.polygon_vertices(walk.left_foot_polytope if side == 0 else walk.right_foot_polytope) if lay.m[i] > 0 else None
    for _ in range(WARMUPS):
        out = timeline.tick(0, xd, yd, fcd, vref, side, poly)
    torch.cuda.synchronize(dev)
    times = []
    for _ in range(TIMED):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = timeline.tick(0, xd, yd, fcd, vref, side, poly)
        e1.record()
        torch.cuda.synchronize(dev)
        times.append(e0.elapsed_time(e1))
    iters = float(out["iters"].view(batch, 2).amax(dim=1).double().mean().item())
    bad = int((out["status"] != 0).sum().item())
    plan.close()
    return float(np.median(times)), iters, bad


def cpu_tick_ms(config: MPCConfig, rebuild: bool):
    """One NON-STRICT tick (x + y) on one host core with NumPy, the reference's own protocol (`:59-73`) and arithmetic
    (`zmp_controller.py:162-171,196-199`): `rebuild=True` rebuilds Px / Pu by the Python double loop and inverts per call
    exactly as the reference does ("reference as-is"), `rebuild=False` caches the matrices
    and the inverse ("fair NumPy", item 3).  Timing yardstick only: nothing of the product runs through it.  The strict
    and Herdt legs of the reference need cvxpy + OSQP; when `import cvxpy` fails (as in this image) they have no CPU
    number here (bench.py reports the C port of the strict path instead)."""
    if config.strict:
        try:
            import cvxpy  # noqa: F401
        except Exception:
            return None
        return None                                     # cvxpy present: time the reference itself, not a restatement
    N, T, hg = config.horizon, config.dt, config.h / config.g
    x, lo, hi = horizon_scaling.tick_inputs(N, 2, seed=1)
    A = np.array([[1, T, T * T / 2], [0, 1, T], [0, 0, 1]])
    B = np.array([T ** 3 / 6, T * T / 2, T])

    def matrices():
        Px, Pu = np.zeros((N, 3)), np.zeros((N, N))
        for i in range(N):
            Px[i] = 1, T * (i + 1), (T ** 2) / 2 * (i + 1) ** 2 - hg
            for j in range(i + 1):
                Pu[i, j] = (T ** 3) / 6 * (1 + 3 * (i - j) + 3 * (i - j) ** 2) - T * hg
        return Px, Pu

    cached = None
    if not rebuild:
        Px, Pu = matrices()
        cached = (Px, Pu, np.linalg.inv(Pu.T @ Pu + config.R / config.Q * np.eye(N)) @ Pu.T)

    def tick():
        for k in (0, 1):
            zref = (hi[k] + lo[k]) / 2
            if rebuild:
                Px, Pu = matrices()
                jerk = -np.linalg.inv(Pu.T @ Pu + config.R / config.Q * np.eye(N)) @ Pu.T @ (Px @ x[k] - zref)
            else:
                jerk = -cached[2] @ (cached[0] @ x[k] - zref)
            _ = A @ x[k] + B * jerk[0]

    for _ in range(1 if rebuild else WARMUPS):
        tick()
    times = []
    for _ in range(3 if rebuild else TIMED):
        start = time.perf_counter()
        tick()
        times.append(time.perf_counter() - start)
    return 1e3 * float(np.mean(times))


def main(argv=None):
    ap = argparse.ArgumentParser(description="Runtime comparison: Wieber (strict vs non-strict) vs Herdt")
    ap.add_argument("--horizons", type=int, nargs="+", default=list(range(10, 301, 10)), help="reference: 10..300 step 10 (:139)")
    ap.add_argument("--batch", type=int, default=16384, help="scenarios per batched tick (BASELINE configs[2])")
    ap.add_argument("--precision", choices=["fp64", "f16x3"], default="fp64", help="arithmetic of the batched strict leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the one-core NumPy restatement beside the GPU numbers")
    ap.add_argument("--no-api", action="store_true", help="skip the B = 1 reference-protocol legs")
    ap.add_argument("--output-dir", type=str, default="results")
    args = ap.parse_args(argv)

    import torch
    from mpc_bipedal import engine

    if not torch.cuda.is_available():
        raise SystemExit("run_compare_runtime.py needs a CUDA device: the product has no CPU path")
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)                   # the CPU numbers are ONE core
    except Exception:
        pass
    precision = engine.PRECISION_F16X3 if args.precision == "f16x3" else engine.PRECISION_FP64
    max_herdt = 152                                    # N + footstep slots must fit the resident-matrix kernel
    print("=" * 60)
    print("Runtime Comparison: Wieber (strict vs non-strict) vs Herdt")
    print("=" * 60)
    rows = []
    for horizon in args.horizons:
        row = {"horizon": int(horizon), "batch": args.batch, "precision": args.precision}
        for tag, strict in (("strict", True), ("non_strict", False)):
            cfg = MPCConfig(horizon=int(horizon), strict=strict, add_force=False, method="wieber")      # reference :152-168
            if strict and precision == engine.PRECISION_F16X3 and horizon > engine.load_library().lipm_max_reduced_horizon():
                row[tag] = {"skipped": "beyond the reduced-precision kernels' horizon limit"}
                continue
            if strict and horizon > engine.load_library().lipm_max_strict_horizon():
                row[tag] = {"skipped": "beyond the strict kernels' horizon limit"}
                continue
            ms, iters, bad = batched_wieber_tick(cfg, args.batch, precision)
            leg = {"batch_ms": ms, "batch_ticks_per_s": args.batch / (ms * 1e-3), "batch_qp_solves_per_s": 2 * args.batch / (ms * 1e-3),
                   "mean_iters_per_solve": iters, "not_converged": bad}
            if not args.no_api:
                leg["api_ms"] = 1e3 * measure_wieber_iteration_time(cfg)
            if not args.no_cpu:
                cpu = cpu_tick_ms(cfg, rebuild=False)
                if cpu is None:
                    leg["cpu_note"] = "the reference's strict leg needs cvxpy + OSQP (not installed): no CPU number here"
                else:
                    leg["cpu_ms"] = cpu
                    leg["cpu_as_reference_ms"] = cpu_tick_ms(cfg, rebuild=True)
                    leg["speedup_per_tick_vs_one_core_as_reference"] = leg["cpu_as_reference_ms"] * args.batch / ms
            row[tag] = leg
        if horizon <= max_herdt and horizon >= 20:
            cfg = MPCConfig(horizon=int(horizon), strict=True, add_force=False, method="herdt")           # reference :170-176
            ms, iters, bad = batched_herdt_tick(cfg, args.batch)
            leg = {"batch_ms": ms, "batch_ticks_per_s": args.batch / (ms * 1e-3), "mean_iters_per_joint_solve": iters, "not_converged": bad}
            if not args.no_api:
                leg["api_ms"] = 1e3 * measure_herdt_iteration_time(cfg)
            row["herdt"] = leg
        else:
            row["herdt"] = {"skipped": "the Herdt kernel serves 20 <= N <= 152 (N + footstep slots fit one resident tile)"}
        rows.append(row)

        def fmt(leg):
            return "skipped" if "skipped" in leg else f"{leg['batch_ms']:.3f} ms / {args.batch} scenarios" + (
                f", API {leg['api_ms']:.3f} ms" if "api_ms" in leg else "") + (f", 1 core {leg['cpu_ms']:.3f} ms" if "cpu_ms" in leg else "")
        print(f"N = {horizon}: strict {fmt(row['strict'])} | non-strict {fmt(row['non_strict'])} | herdt {fmt(row['herdt'])}", flush=True)
    out = os.path.join(args.output_dir, "runtime_comparison")
    write_json(out + ".json", {"protocol": {"warmups": WARMUPS, "timed": TIMED, "wieber": "mean", "herdt": "median"}, "rows": rows})

    def column(leg, key):
        return np.array([r[leg].get(key, np.nan) if isinstance(r.get(leg), dict) else np.nan for r in rows], float)
    np.savez(out + ".npz", horizon=np.array([r["horizon"] for r in rows]),
             **{f"{leg}_{key}": column(leg, key) for leg in ("strict", "non_strict", "herdt") for key in ("batch_ms", "api_ms", "cpu_ms")})
    print(f"Results written to {out}.json / .npz")
    try:                                               # the reference's figure (:205-300), when plotly is installed
        import plotly.graph_objects as go
        fig = go.Figure()
        for leg, name in (("strict", "Wieber strict"), ("non_strict", "Wieber non-strict"), ("herdt", "Herdt")):
            fig.add_trace(go.Scatter(x=[r["horizon"] for r in rows], y=column(leg, "batch_ms") / args.batch, mode="lines+markers",
                                     name=f"{name} (ms per scenario tick, batch {args.batch})"))
        fig.update_layout(title="Runtime vs horizon", xaxis_title="Horizon N", yaxis_title="ms", yaxis_type="log")
        fig.write_html(out + ".html")
    except Exception:
        pass
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
