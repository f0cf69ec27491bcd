#!/usr/bin/env python3
"""Herdt free-footstep MPC driver (reference scripts/run_herdt_method.py), single walk or batched plans.

    python scripts/run_herdt_method.py                       # the reference's experiment (class defaults)
    python scripts/run_herdt_method.py --F-ext 200
    python scripts/run_herdt_method.py --plans 8192          # BASELINE config 4: randomized footstep plans

--plans B draws step_length ~ U(0.2, 0.4), foot_spread ~ U(0.08, 0.12), distance ~ U(1.5, 2.5) (seed 2,
speed_generation "classic"); plans are grouped by their state timeline (scenarios of a group share the
per-tick Hessians) and every group runs as one batch; foot_spread (standing hull, initial foot position)
is a per-scenario input of the engine.
"""

import argparse
import os
import time
from dataclasses import replace

import numpy as np

from _common import load_config_from_json, write_json
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator


def main(argv=None):
    parser = argparse.ArgumentParser(description="Run Herdt MPC with speed/state generation.")
    parser.add_argument("--config", type=str)
    parser.add_argument("--output-dir", type=str, default="results")
    parser.add_argument("--F-ext", type=float, dest="F_ext")
    parser.add_argument("--plans", type=int, default=1, help="randomized footstep plans (1 = the reference's single walk)")
    parser.add_argument("--seed", type=int, default=2)
    args = parser.parse_args(argv)

    config = load_config_from_json(args.config) if args.config and os.path.exists(args.config) else MPCConfig()
    if args.F_ext is not None:
        config.F_ext, config.add_force = args.F_ext, True
    config.method = "herdt"
    os.makedirs(args.output_dir, exist_ok=True)
    print(f"Herdt MPC: horizon {config.horizon}, dt {config.dt:.4f}, alpha {config.alpha}, beta {config.beta}, "
          f"gamma {config.gamma}, F_ext {config.F_ext} (add_force={config.add_force})")

    out = os.path.join(args.output_dir, "herdt")
    if args.plans <= 1:
        v_x, v_y, states = SpeedTrajectoryGenerator(config).generate_speed_and_state(save_footsteps=False,
                                                                                     output_dir=args.output_dir)
        v_ref, state_ref = np.stack([v_x, v_y], axis=1), np.array(states)
        controller = ZMPController(config)
        t0 = time.perf_counter()
        com, y_hist, foot_hist = controller.generate_com_trajectory(x_init=np.zeros((3, 1)), y_init=np.zeros((3, 1)),
                                                                    v_ref=v_ref, state_ref=state_ref)
        seconds = time.perf_counter() - t0
        zmp_y = np.tensordot(y_hist[:, :, 0], controller.C, axes=([1], [0]))       # reference run_herdt_method.py:200
        np.savez(out + ".npz", com=com, y_hist=y_hist, foot_hist=foot_hist, zmp_y=zmp_y, v_ref=v_ref)
        write_json(out + ".json", {"n_steps": len(v_ref), "seconds": seconds, "final_com": com[-1].tolist(),
                                   "footsteps": np.unique(foot_hist, axis=0).tolist()})
        print(f"COM trajectory shape: {com.shape}; final CoM ({com[-1, 0]:.4f}, {com[-1, 1]:.4f}); {seconds:.2f} s")
        return 0

    from mpc_bipedal import engine
    from mpc_bipedal.herdt import randomized_plans, run_plan_groups

    B = args.plans
    plans = randomized_plans(config, B, seed=args.seed)
    groups = plans["groups"]
    step_length, spread, distance, speed = (plans[k] for k in ("step_length", "foot_spread", "distance", "speed"))
    print(f"{B} plans in {len(groups)} timeline groups")
    final = np.zeros((B, 2))
    status = np.zeros(B, np.int32)
    solves = 0
    # warm-up: CUDA context, kernel images, page-locked staging buffers
    ZMPController(config).generate_com_trajectory_herdt_batched(np.zeros((2, 3)), np.zeros((2, 3)), groups[0]["v"], groups[0]["states"])
    import torch

    t0 = time.perf_counter()
    results, controllers, streams = run_plan_groups(config, plans)
    t_built = 0.0
    for entry, res, stream in zip(groups, results, streams):
        ids = entry["ids"]
        with torch.cuda.stream(stream):
            res = engine.to_host(res)
        print(f"  group of {len(ids):4d} plans, {len(entry['states'])} samples: {time.perf_counter() - t0:.2f} s elapsed, "
              f"{res['iters'].sum() / (len(ids) * (len(entry['states']) - 1)):.0f} iterations per joint solve")
        final[ids] = res["com"][:, -1]
        status[ids] = res["status"]
        solves += len(ids) * (len(entry["states"]) - 1)
    seconds = time.perf_counter() - t0
    summary = {"plans": B, "groups": len(groups), "seconds": seconds, "seconds_building_timelines": t_built, "joint_qp_solves": solves,
               "joint_qp_solves_per_s": solves / seconds, "not_converged_plans": int((status != 0).sum()),
               "mean_final_x": float(final[:, 0].mean())}
    print(summary)
    np.savez_compressed(out + "_plans.npz", step_length=step_length, foot_spread=spread, distance=distance, speed=speed,
                        final_com=final, status=status)
    write_json(out + "_plans.json", summary)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
