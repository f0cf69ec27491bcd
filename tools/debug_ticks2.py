import sys, os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"))
from mpc_bipedal import engine
from oracle import lipm_oracle as orc
np.set_printoptions(linewidth=200, precision=4, suppress=True)
rng = np.random.default_rng(11)
for N in [16, 24, 40]:
    plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6)
    p = orc.Problem.make(N)
    cases = 24
    x = rng.normal(size=(cases, 3)) * np.array([0.05, 0.1, 0.5])
    centre = rng.uniform(-0.05, 0.05, size=(cases, 1)) + np.cumsum(
        (rng.random((cases, N)) < 0.04) * rng.uniform(-0.1, 0.1, size=(cases, N)), axis=1)
    lo, hi = centre - 0.055, centre + 0.055
    if N != 40: continue
    M = orc.admm_matrix(p, 2.0)
    def run(idx):
        res = plan.solve_tick(engine.MODE_STRICT, x[idx], lo[idx], hi[idx], want_plan=True)
        return res["zmp_plan"].cpu().numpy(), res["iters"].cpu().numpy()
    z_all, it_all = run(np.arange(cases))
    zo = np.array([orc.strict_axis_admm(p, x[k], hi[k], lo[k], eps=1e-12, M=M)[1] for k in range(cases)])
    err = np.abs(z_all - zo).max(1)
    print("all 24: bad cases", np.flatnonzero(err > 1e-9), "iters", it_all)
    for k in np.flatnonzero(err > 1e-9)[:2]:
        z1, it1 = run(np.array([k]))
        print(f"case {k} alone: err {np.abs(z1[0]-zo[k]).max():.2e} iters {it1}")
        z8, it8 = run(np.array([k]*8))
        print(f"case {k} x8: errs {np.abs(z8-zo[k]).max(1)} iters {it8}")
        print(" lo ", lo[k]); print(" hi ", hi[k]); print(" zo ", zo[k]); print(" zg ", z_all[k]); print(" x", x[k])
        act = (zo[k] <= lo[k] + 1e-12) * -1 + (zo[k] >= hi[k] - 1e-12) * 1
        print(" act", act)
    perm = rng.permutation(cases)
    zp, itp = run(perm)
    errp = np.abs(zp - zo[perm]).max(1)
    print("permuted: bad original cases", np.sort(perm[errp > 1e-9]))
    # repeat the same launch: deterministic?
    z_b, it_b = run(np.arange(cases))
    print("repeat identical:", np.array_equal(z_b, z_all), np.array_equal(it_b, it_all))
