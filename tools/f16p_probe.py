#!/usr/bin/env python3
"""Pair kernel (cta_group::2) against the single-CTA tcgen05 kernel: bit-identity and timing on the push sweep."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, p)
import torch
from mpc_bipedal import engine
sys.path.insert(0, REPO)
from bench import sweep_inputs

def run(B, pair, reps=3):
    cfg, z_max, z_min, F, tick = sweep_inputs(B)
    plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, precision=engine.PRECISION_F16X3)
    plan.set_single_cta_f16(not pair)
    dev = plan.device
    x0 = torch.zeros((B, 3), dtype=torch.float64, device=dev)
    zmin, zmax = torch.as_tensor(z_min).to(dev), torch.as_tensor(z_max).to(dev)
    imp, tk = torch.as_tensor(cfg.dt * F / cfg.m).to(dev), torch.as_tensor(tick).to(dev)
    out = {}
    plan.rollout(engine.MODE_STRICT, x0, x0, zmin, zmax, impulse=imp, force_tick=tk, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.rollout(engine.MODE_STRICT, x0, x0, zmin, zmax, impulse=imp, force_tick=tk, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = len(z_min)
    return ms, B * 2 * (n - 1) / (ms * 1e-3), {k: v.clone() for k, v in out.items() if not k.startswith("_")}, plan.stats()

for B in [int(a) for a in sys.argv[1:]] or [64, 4096, 32768]:
    ms1, r1, o1, s1 = run(B, False)
    t0 = time.time()
    ms2, r2, o2, s2 = run(B, True)
    same = {k: bool(torch.equal(o1[k], o2[k])) for k in ("y_hist", "x_hist", "flags", "iters", "status", "max_violation")}
    print(f"B={B}: single-CTA {ms1:.2f} ms {r1/1e6:.1f} M solves/s | pair {ms2:.2f} ms {r2/1e6:.1f} M solves/s | identical {same} "
          f"| iters {int(o1['iters'].sum())} vs {int(o2['iters'].sum())} | status {int(o2['status'].sum())}", flush=True)
    if not all(same.values()):
        d = (o1["y_hist"] - o2["y_hist"]).abs()
        print("   max |dy|", d.max().item(), "first differing scenario", int((d.amax(dim=(1, 2)) > 0).nonzero()[0]) if d.max() > 0 else None)
