#!/usr/bin/env python3
"""Main driver: one walk (reference scripts/run_mpc.py) -- JSON config + CLI overrides -> CoP bounds ->
receding-horizon rollout on the GPU -> results/ (npz + summary json; a plotly figure when plotly exists).

    python scripts/run_mpc.py --config configs/default.json
    python scripts/run_mpc.py --distance 3.0 --horizon 100 --no-strict
    python scripts/run_mpc.py --create-config configs/my_config.json
"""

import argparse
import os

import numpy as np

from _common import add_field_flags, apply_field_flags, base_config, write_json
from mpc_bipedal.config import MPCConfig
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal.generators import CoPGenerator


def create_default_config_file(path: str) -> None:
    """Writes {"mpc": {...class defaults...}} (reference run_mpc.py:42-78; the loader only reads "mpc")."""
    write_json(path, {"mpc": MPCConfig().to_dict()})
    print(f"Configuration file written: {path}")


def main(argv=None):
    parser = argparse.ArgumentParser(description="LIPM-MPC walk (batched B200 engine, B = 1)")
    parser.add_argument("--config", type=str)
    parser.add_argument("--create-config", type=str, metavar="FILE")
    add_field_flags(parser)
    parser.add_argument("--no-visualization", action="store_true")
    parser.add_argument("--save-animation", action="store_true")     # accepted for CLI parity; no 3-D animation here
    parser.add_argument("--output-dir", type=str, default="results")
    args = parser.parse_args(argv)

    if args.create_config:
        create_default_config_file(args.create_config)
        return 0
    config = apply_field_flags(base_config(args.config), args)
    if config.method.lower() != "wieber":
        raise SystemExit("run_mpc.py drives the wieber method; use scripts/run_herdt_method.py for herdt "
                         "(the reference's own `run_mpc.py --method herdt` raises: no v_ref/state_ref are passed)")
    os.makedirs(args.output_dir, exist_ok=True)
    print("=" * 60)
    print(f"horizon={config.horizon} dt={config.dt:.5f} strict={config.strict} add_force={config.add_force} "
          f"F_ext={config.F_ext} distance={config.distance} step_length={config.step_length}")

    z_max, z_min, states = CoPGenerator(config).generate_cop_trajectory(save_footsteps=False, output_dir=args.output_dir)
    controller = ZMPController(config)
    x_init, y_init = np.zeros((3, 1)), np.zeros((3, 1))
    com, y_hist = controller.generate_com_trajectory(x_init, y_init, z_max, z_min)
    zmp_y = np.tensordot(y_hist[:, :, 0], controller.C, axes=([1], [0]))          # reference run_mpc.py:294
    violation = float(np.maximum(z_min[:, 1] - zmp_y, zmp_y - z_max[:, 1]).max())

    out = os.path.join(args.output_dir, "run_mpc")
    np.savez(out + ".npz", com=com, y_hist=y_hist, zmp_y=zmp_y, z_max=z_max, z_min=z_min)
    write_json(out + ".json", {"n_steps": int(len(z_min)), "qp_solves": int(2 * (len(z_min) - 1)),
                               "final_com": com[-1].tolist(), "max_zmp_y_violation_m": max(violation, 0.0),
                               "config": config.to_dict()})
    print(f"final CoM = ({com[-1, 0]:.4f}, {com[-1, 1]:.4f}); max ZMP-y violation = {max(violation, 0.0):.3e} m")
    if not args.no_visualization:
        try:
            import plotly.graph_objects as go
            t = np.arange(len(z_min)) * config.dt
            fig = go.Figure()
            for name, series in (("z_max y", z_max[:, 1]), ("z_min y", z_min[:, 1]), ("CoM y", com[:, 1]), ("ZMP y", zmp_y)):
                fig.add_trace(go.Scatter(x=t, y=series, name=name))
            fig.write_html(out + ".html")
        except ImportError:
            print("plotly not installed: skipping the figure (results are in the .npz / .json files)")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
