"""Debug helper: strict single ticks on the GPU vs the NumPy ADMM oracle for several horizons."""
import sys, os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"))
from mpc_bipedal import engine
from oracle import lipm_oracle as orc

rng = np.random.default_rng(11)
for N in [int(a) for a in sys.argv[1:]] or [16, 24, 40, 50, 56, 64, 96, 100, 120, 128, 150, 160]:
    plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6)
    p = orc.Problem.make(N)
    cases = 24
    x = rng.normal(size=(cases, 3)) * np.array([0.05, 0.1, 0.5])
    centre = rng.uniform(-0.05, 0.05, size=(cases, 1)) + np.cumsum(
        (rng.random((cases, N)) < 0.04) * rng.uniform(-0.1, 0.1, size=(cases, N)), axis=1)
    lo, hi = centre - 0.055, centre + 0.055
    res = plan.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    z = res["zmp_plan"].cpu().numpy(); it = res["iters"].cpu().numpy(); st = res["status"].cpu().numpy()
    M = orc.admm_matrix(p, 2.0)
    worst = 0
    for k in range(cases):
        _, zo, _, ito = orc.strict_axis_admm(p, x[k], hi[k], lo[k], eps=1e-12, M=M)
        err = np.abs(z[k] - zo)
        worst = max(worst, err.max())
        if err.max() > 1e-9:
            bad = np.flatnonzero(err > 1e-9)
            print(f"  N={N} case {k}: max err {err.max():.2e} rows {bad[:12]}{'...' if len(bad)>12 else ''} ({len(bad)} rows) iters gpu {it[k]} oracle {ito} status {st[k]}")
    print(f"N={N}: worst {worst:.2e}  mean iters gpu {it.mean():.1f}")
