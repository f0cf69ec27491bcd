#!/usr/bin/env python3
"""Streamed reduced-precision strict kernel (N > 152) against the FP64 path: one tick, error and time."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"), os.path.join(REPO, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch
from mpc_bipedal import engine
from horizon_scaling import tick_inputs

for N in [int(a) for a in sys.argv[1:]] or [300]:
    B = 1024
    x, lo, hi = tick_inputs(N, B)
    fp64 = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6)
    f16 = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, precision=engine.PRECISION_F16X3)
    want = fp64.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    torch.cuda.synchronize()
    t0 = time.time()
    got = f16.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    z, zw = got["zmp_plan"].cpu().numpy(), want["zmp_plan"].cpu().numpy()
    err = np.abs(z - zw).max(axis=1)
    print(N, "max err", err.max(), "median err", np.median(err), "bad cols", int((err > 1e-4).sum()), "status", int((got["status"].cpu().numpy() != 0).sum()),
          "iters f16", got["iters"].float().mean().item(), "fp64", want["iters"].float().mean().item(), "first call s", round(dt, 4), flush=True)
    if err.max() > 1e-4:
        b = int(err.argmax()); print(" worst col", b, "rows", np.nonzero(np.abs(z[b] - zw[b]) > 1e-4)[0][:20], z[b][:6], zw[b][:6])
