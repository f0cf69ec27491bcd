"""Cycles per ADMM group-iteration with 1, 2, 3 active solver groups per SM (forced iteration count)."""
import os, sys, subprocess, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"))

def child(groups, max_iter, ticks):
    import torch
    from mpc_bipedal import engine
    plan = engine.LipmPlan(150, 0.75, 9.81, None, 1.0, 1e-6, eps=1e-300, max_iter=max_iter)
    rounds = 2
    units = 148 * groups * rounds
    B = units * 8 // 2
    n = ticks + 1
    lo = np.full((n, 2), -0.05); hi = np.full((n, 2), 0.05)
    x0 = np.random.default_rng(0).normal(size=(B, 3)) * [0.01, 0.02, 0.05]
    args = [torch.as_tensor(a).cuda() for a in (x0, x0, lo, hi)]
    out = {}
    for _ in range(2):
        plan.rollout(engine.MODE_STRICT, *args, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.rollout(engine.MODE_STRICT, *args, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = plan.stats()
    per_slot = st["group_iterations"] / (148 * groups)
    print(json.dumps({"groups": groups, "ms": ms, "group_iterations": st["group_iterations"],
                      "cycles_per_group_iteration_at_1965MHz": ms * 1e-3 * 1.965e9 / per_slot,
                      "pipe_cycles_needed_per_iteration": 722 * 16 / 4,
                      "t0_contraction_cycles_per_iter": st["debug_cycles_contraction"] / (2 * ticks * max_iter),
                      "t0_iteration_cycles_per_iter": st["debug_cycles_iteration"] / (2 * ticks * max_iter)}))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
    else:
        for skip in (24,):
            for g in (1, 2, 3):
                env = dict(os.environ, LIPM_STRICT_GROUPS=str(g), LIPM_STRICT_SKIP=str(skip))
                print("skip", skip, end=" ", flush=True)
                subprocess.run([sys.executable, __file__, str(g), "100", "12"], env=env, check=True)
