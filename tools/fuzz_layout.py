#!/usr/bin/env python3
"""Randomized layout-independence check: a scenario's walk must come out bit-identical whether it is solved inside a
random batch (random size, ragged lengths, shared or per-scenario bounds, random pushes) or alone -- for the FP64
strict solver, the reduced-precision kernel and the unconstrained mode.  The default FP64 solver has two formulations
of the same optimum (row-parallel kernel for walks on a shared series of uniform length, Riccati kernel otherwise): the
check is bit for bit with the row-parallel stage switched off, and with it on whenever batch and single walk use the
same formulation (neither ragged nor own bounds); across the two formulations the walks must agree to 1e-9 relative and
in every flag.  Prints one line per trial; exits non-zero on the first mismatch."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mpc_bipedal import engine

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
cfg, z_max_full, z_min_full, _, _ = bench.sweep_inputs(1)
plans = {"fp64": engine.get_plan(150, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R),
         "f16x3": engine.get_plan(150, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, precision=engine.PRECISION_F16X3)}
for trial in range(trials):
    B = int(rng.choice([1, 3, 9, 130, 700, 2500, 5000]))
    n = int(rng.integers(3, 90))
    start = int(rng.integers(0, 300))
    z_max, z_min = z_max_full[start:start + n], z_min_full[start:start + n]
    ragged = rng.random() < 0.5
    own_bounds = rng.random() < 0.3
    n_steps = rng.integers(2, n + 1, B).astype(np.int32) if ragged else None
    if own_bounds:
        shift = rng.uniform(-0.01, 0.01, (B, 1, 2))
        zmax_b, zmin_b = z_max[None] + shift, z_min[None] + shift
    x0 = rng.normal(size=(B, 3)) * [0.005, 0.01, 0.02]
    y0 = rng.normal(size=(B, 3)) * [0.005, 0.01, 0.02]
    impulse = rng.uniform(0, 0.25, B)
    tick = rng.integers(0, n, B).astype(np.int32)
    for name, mode, plan in (("fp64", engine.MODE_STRICT, plans["fp64"]), ("fp64_one_kernel", engine.MODE_STRICT, plans["fp64"]),
                             ("f16x3", engine.MODE_STRICT, plans["f16x3"]), ("free", engine.MODE_UNCONSTRAINED, plans["fp64"])):
        plans["fp64"].set_row_parallel(name != "fp64_one_kernel")
        bitwise = not (name == "fp64" and (ragged or own_bounds))
        zm, zM = (zmin_b, zmax_b) if own_bounds else (z_min, z_max)
        out = plan.rollout(mode, x0, y0, zm, zM, n_steps=n_steps, impulse=impulse, force_tick=tick)
        out = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
        for b in rng.choice(B, size=min(B, 4), replace=False):
            nb = int(n_steps[b]) if ragged else n
            zm1, zM1 = (zmin_b[b, :nb], zmax_b[b, :nb]) if own_bounds else (z_min[:nb], z_max[:nb])
            one = plan.rollout(mode, x0[b:b + 1], y0[b:b + 1], zm1, zM1, impulse=impulse[b:b + 1], force_tick=tick[b:b + 1])
            for key in ("x_hist", "y_hist", "flags"):
                got, want = one[key][0].cpu().numpy(), out[key][b, :nb]
                same = np.array_equal(got, want) if (bitwise or key == "flags") else \
                    float(np.abs(got - want).max()) <= 1e-9 * max(1.0, float(np.abs(want).max()))
                if not same:
                    print(f"MISMATCH trial {trial} {name} B={B} n={n} ragged={ragged} own_bounds={own_bounds} scenario {b} {key}")
                    sys.exit(1)
            if out["y_hist"][b, nb:].any():
                print(f"WRITE PAST END trial {trial} {name} scenario {b}")
                sys.exit(1)
    print(f"trial {trial}: B={B} n={n} ragged={ragged} own_bounds={own_bounds} ok", flush=True)
plans["fp64"].set_row_parallel(True)
print("all trials ok")
