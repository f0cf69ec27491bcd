#!/usr/bin/env python3
"""Reduced-precision push sweep (4096 scenarios) against the FP64 path for several stop tolerances: accuracy vs throughput."""
import os, sys, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal import engine
from bench import sweep_inputs
B = 4096
cfg, z_max, z_min, F, tick = sweep_inputs(B)
dev = torch.device("cuda", 0)
x0 = torch.zeros((B, 3), dtype=torch.float64, device=dev)
zmin, zmax = torch.as_tensor(z_min).to(dev), torch.as_tensor(z_max).to(dev)
imp, tk = torch.as_tensor(cfg.dt * F / cfg.m).to(dev), torch.as_tensor(tick).to(dev)
ref_plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
ref = ref_plan.rollout(engine.MODE_STRICT, x0, x0, zmin, zmax, impulse=imp, force_tick=tk)
fell = engine.unpack_outcome(ref["outcome"])["fell"]
rows = []
for eps in (1e-5, 5e-6, 2e-6, 1e-6, 5e-7):
    plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, precision=engine.PRECISION_F16X3, eps=eps)
    out = {}
    plan.rollout(engine.MODE_STRICT, x0, x0, zmin, zmax, impulse=imp, force_tick=tk, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan.rollout(engine.MODE_STRICT, x0, x0, zmin, zmax, impulse=imp, force_tick=tk, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    diff = torch.maximum((out["y_hist"][..., 0] - ref["y_hist"][..., 0]).abs().amax(dim=1), (out["x_hist"][..., 0] - ref["x_hist"][..., 0]).abs().amax(dim=1))
    f16 = engine.unpack_outcome(out["outcome"])["fell"]
    row = {"eps": eps, "ms": ms, "M_solves_per_s": B * 838 / ms / 1e3, "iters_per_solve": int(out["iters"].sum()) / (B * 838),
           "median_diff_m": diff.median().item(), "within_1e-4_all": float((diff <= 1e-4).double().mean()),
           "within_1e-4_not_falling": float((diff[~fell] <= 1e-4).double().mean()), "max_diff_not_falling_m": diff[~fell].max().item(),
           "fall_mismatches": int((fell != f16).sum()), "not_converged": int((out["status"] != 0).sum())}
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(REPO, "gpurun_out", "r02_reduced_eps_scan.json"), "w"), indent=1)
