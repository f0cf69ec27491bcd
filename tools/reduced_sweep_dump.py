"""Dump FP64 vs F16X3 push-sweep trajectories (positions + ZMP) for offline analysis of the parity bound."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
from mpc_bipedal import engine
from test_cuda_parity import push_sweep_inputs
B = 4096
cfg, z_max, z_min, impulse, tick = push_sweep_inputs(B)
zero = np.zeros((B, 3))
res = {}
for name, prec, eps in (("fp64", engine.PRECISION_FP64, 0.0), ("f16", engine.PRECISION_F16X3, 0.0), ("f16_eps1e-6", engine.PRECISION_F16X3, 1e-6),
                        ("fp64_eps1e-9", engine.PRECISION_FP64, 1e-9)):
    plan = engine.get_plan(150, 0.75, 9.81, 0.01, 1.0, 1e-6, precision=prec, eps=eps)
    out = plan.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, impulse=impulse, force_tick=tick)
    y = out["y_hist"].cpu().numpy()
    if name == "fp64":
        base = y
        res["fp64_pos"] = y[:, :, 0]
        res["fp64_zmp"] = y[:, :, 0] - 0.75 / 9.81 * y[:, :, 2]
    else:
        res[name + "_dpos"] = (y[:, :, 0] - base[:, :, 0]).astype(np.float32)
    res[name + "_iters"] = out["iters"].cpu().numpy()
np.savez_compressed(os.path.join(REPO, "gpurun_out", "reduced_sweep.npz"), impulse=impulse, tick=tick, **res)
print("saved")
