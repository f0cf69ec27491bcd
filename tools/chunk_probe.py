"""Active-set walks cut into chunks on side streams (LIPM_OPT_WALK_CHUNKS): time per setting, bit-identity with the serial run.

    python tools/chunk_probe.py [B]
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"))
import bench  # noqa: E402
from mpc_bipedal import engine  # noqa: E402
from tools.riccati_probe import timed  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    cfg, z_max, z_min, F, tick = bench.sweep_inputs(B)
    n = len(z_min)
    dev = torch.device("cuda", 0)
    zmin = torch.tensor(z_min, device=dev); zmax = torch.tensor(z_max, device=dev)
    x0 = torch.zeros(B, 3, dtype=torch.float64, device=dev); y0 = torch.zeros_like(x0)
    imp = torch.tensor(cfg.dt * F / cfg.m, device=dev); ft = torch.tensor(tick, device=dev)
    solves = B * 2 * (n - 1)
    plan = engine.LipmPlan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
    out = {"B": B}
    ref = None
    for chunks in (1, 2, 3, 4, 0):
        plan.set_walk_chunks(chunks)
        r = {}
        ms = timed(lambda: plan.rollout(1, x0, y0, zmin, zmax, impulse=imp, force_tick=ft, out=r, want_readout=True), reps=5)
        st = plan.stats()
        cur = {k: v.clone() for k, v in r.items() if torch.is_tensor(v)}
        rec = {"ms": ms, "M_solves_per_s": solves / ms / 1e3, "sweeps": st["group_iterations"], "columns": st["units"]}
        if ref is None:
            ref = cur
        else:
            rec["tensors_differing_from_serial"] = [k for k in ref if not torch.equal(ref[k], cur[k])]
        out[f"chunks{chunks}"] = rec
        print(json.dumps({f"chunks{chunks}": rec}), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", f"chunk_probe_b{B}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
