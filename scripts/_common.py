"""Shared plumbing of the drivers: import path, JSON loading, CLI overrides of MPCConfig fields.

The drivers keep the reference's command lines and JSON schema (scripts/run_mpc.py:108-147,
run_compare_resistance.py:59-61, run_herdt_method.py:44-46); plotting is optional and lazy (plotly /
matplotlib are not needed to compute anything), results are also written as .npz / .json.
"""

import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACKAGE_DIR = os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")
for _p in (REPO, PACKAGE_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from mpc_bipedal.config import MPCConfig  # noqa: E402

# (flag, MPCConfig field, type) -- the reference's per-field overrides, run_mpc.py:112-141
FIELD_FLAGS = [
    ("--distance", "distance", float), ("--step-length", "step_length", float), ("--foot-spread", "foot_spread", float),
    ("--ssp-duration", "ssp_duration", float), ("--dsp-duration", "dsp_duration", float),
    ("--standing-duration", "standing_duration", float), ("--Q", "Q", float), ("--R", "R", float), ("--S", "S", float),
    ("--h", "h", float), ("--m", "m", float), ("--F-ext", "F_ext", float), ("--alpha", "alpha", float),
    ("--beta", "beta", float), ("--gamma", "gamma", float), ("--vx-ref", "vx_ref", float), ("--vy-ref", "vy_ref", float),
    ("--foot-length", "foot_length", float), ("--foot-width", "foot_width", float),
]


def load_config_from_json(config_file: str) -> MPCConfig:
    """Same contract as the reference helper duplicated in its drivers (run_mpc.py:23-39)."""
    return MPCConfig.from_json(config_file)


def base_config(path=None) -> MPCConfig:
    """--config, else configs/default.json when present, else the class defaults (run_mpc.py:156-166)."""
    if path:
        return load_config_from_json(path)
    default = os.path.join(REPO, "configs", "default.json")
    return load_config_from_json(default) if os.path.exists(default) else MPCConfig()


def add_field_flags(parser):
    for flag, field, kind in FIELD_FLAGS:
        parser.add_argument(flag, dest=field, type=kind, default=None)
    parser.add_argument("--horizon", type=int, default=None)
    parser.add_argument("--dt", type=float, default=None)
    parser.add_argument("--strict", action="store_true", default=None)
    parser.add_argument("--no-strict", action="store_true", dest="no_strict")
    parser.add_argument("--add-force", action="store_true", default=None)
    parser.add_argument("--no-add-force", action="store_true", dest="no_add_force")
    parser.add_argument("--method", choices=["wieber", "herdt"], default=None)


def apply_field_flags(config: MPCConfig, args) -> MPCConfig:
    for _, field, _ in FIELD_FLAGS:
        value = getattr(args, field, None)
        if value is not None:
            setattr(config, field, value)
    # horizon wins over dt; dt is always 1.5 / horizon (run_mpc.py:182-193)
    if args.horizon is not None:
        config.horizon = args.horizon
    elif args.dt is not None:
        config.horizon = int(1.5 / args.dt)
    config.dt = 1.5 / config.horizon
    if args.strict:
        config.strict = True
    elif args.no_strict:
        config.strict = False
    if args.add_force:
        config.add_force = True
    elif args.no_add_force:
        config.add_force = False
    if args.method is not None:
        config.method = args.method
    return config


def write_json(path, payload):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as handle:
        json.dump(payload, handle, indent=1)
