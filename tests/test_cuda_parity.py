"""Parity of the CUDA path (through the C ABI) with the reference's own outputs and the oracle.

All tests need a B200 (``-m gpu``).  Tolerances (north_star): FP64 path <= 1e-9 relative on
trajectories; flag bits bit-exact, except at ties, which are defined explicitly where used.
``ref_*`` golden files are outputs of the reference itself; ``oracle_*`` of the exact oracle.
"""

import os

import numpy as np
import pytest

from oracle import lipm_oracle as orc

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL = 1e-9


@pytest.fixture(scope="module")
def engine():
    from mpc_bipedal import engine as eng
    return eng


@pytest.fixture(scope="module", params=["active_set", "admm"])
def solver(request):
    """Both FP64 strict solvers go through the same tests: the active-set / Riccati kernel (default) and the ADMM
    tensor-pipe kernel share no arithmetic, the oracle they are compared with is a third implementation."""
    return request.param


@pytest.fixture(scope="module")
def plan150(engine, solver):
    return engine.get_plan(150, 0.75, 9.81, 1.5 / 150, 1.0, 1e-6, solver=solver)


def rel_err(got, want):
    return float(np.max(np.abs(np.asarray(got) - np.asarray(want))) / max(np.max(np.abs(want)), 1e-300))


def rel_err_mm(got, want):
    """Relative error of a trajectory with a millimetre floor on the scale (walks of a few samples stay near zero)."""
    return float(np.max(np.abs(np.asarray(got) - np.asarray(want))) / max(np.max(np.abs(want)), 1e-3))


# ----------------------------------------------------------------------------- plan matrices
@pytest.mark.parametrize("N,h", [(150, 0.75), (100, 0.62), (16, 0.9), (37, 0.75)])
def test_plan_matrices(engine, N, h):
    plan = engine.get_plan(N, h, 9.81, 1.5 / N, 1.0, 1e-6)
    p = orc.Problem.make(N, h=h)
    # Px, Pu, A, B, C: bit-identical to the reference's float64 loop (zmp_controller.py:15-21,162-171)
    assert np.array_equal(plan.export(engine.EXPORT_PX), p.Px)
    assert np.array_equal(plan.export(engine.EXPORT_PU), p.Pu)
    abc = plan.export(engine.EXPORT_ABC)
    assert np.array_equal(abc[:9].reshape(3, 3), p.A) and np.array_equal(abc[9:12], p.B) and np.array_equal(abc[12:], p.C)
    c = p._cache()
    M = np.linalg.inv(c["Hz"] + 2.0 * np.eye(N))
    assert rel_err(plan.export(engine.EXPORT_ADMM_M), M) < 1e-11
    assert rel_err(plan.export(engine.EXPORT_HINV), np.linalg.inv(c["Hz"])) < 1e-11
    assert rel_err(plan.export(engine.EXPORT_GPX), c["GPx"]) < 1e-11
    # first-jerk gain row against the reference's closed form -inv(Pu'Pu + R/Q I) Pu' (Px x - zref)
    K = (c["Hu_inv"] @ p.Pu.T)[0]
    gain = plan.export(engine.EXPORT_GAIN)
    assert rel_err(gain[:N], K) < 1e-9
    assert rel_err(gain[N:], -K @ p.Px) < 1e-9


# ----------------------------------------------------------------------------- unconstrained mode
def test_nonstrict_ticks_match_reference(engine, golden):
    g = golden("ref_nonstrict_ticks")
    for N in (10, 50, 100, 150):
        for h in (0.75, 0.62):
            key = f"N{N}_h{h}"
            plan = engine.get_plan(N, h, 9.81, 1.5 / N, 1.0, 1e-6)
            res = plan.solve_tick(engine.MODE_UNCONSTRAINED, g[f"{key}/x"], g[f"{key}/z_min"], g[f"{key}/z_max"],
                                  want_plan=True)
            got = res["next_state"].cpu().numpy()
            assert rel_err(got, g[f"{key}/next"]) < REL, key
            # the full jerk sequence is the reference's X_burst: check it through Pu J + Px x = zmp plan
            p = orc.Problem.make(N, h=h)
            zmp, jerk = res["zmp_plan"].cpu().numpy(), res["jerk_plan"].cpu().numpy()
            for k in range(len(got)):
                x = g[f"{key}/x"][k]
                zref = (g[f"{key}/z_max"][k] + g[f"{key}/z_min"][k]) / 2
                want_jerk = -p._cache()["Hu_inv"] @ p.Pu.T @ (p.Px @ x - zref)
                assert rel_err(jerk[k], want_jerk) < 1e-8, key
                assert np.max(np.abs(p.Pu @ jerk[k] + p.Px @ x - zmp[k])) < 1e-11


def test_nonstrict_default_walk_matches_reference(engine, plan150, golden):
    g = golden("ref_nonstrict_walk")
    zero = np.zeros((1, 3))
    out = plan150.rollout(engine.MODE_UNCONSTRAINED, zero, zero, g["z_min"], g["z_max"],
                          impulse=[float(g["impulse"])], force_tick=[int(g["force_tick"])])
    y_hist = out["y_hist"][0].cpu().numpy()
    x_hist = out["x_hist"][0].cpu().numpy()
    assert rel_err(y_hist, g["y_hist"][:, :, 0]) < REL
    assert rel_err(np.stack([x_hist[:, 0], y_hist[:, 0]], 1), g["com"]) < REL
    # max ZMP-y violation 0.038 m after the push
    viol = out["max_violation"][0].cpu().numpy()
    assert 0.03 < viol[1] < 0.045 and viol[0] < 1e-3
    flags = out["flags"][0].cpu().numpy()
    assert (flags[:, 1] & (engine.FLAG_VIOL_LO | engine.FLAG_VIOL_HI)).any()
    assert not (flags & (engine.FLAG_ACTIVE_LO | engine.FLAG_ACTIVE_HI | engine.FLAG_NOT_CONVERGED)).any()
    # no push: generate_state_trajectory_wieber
    out = plan150.rollout(engine.MODE_UNCONSTRAINED, zero, zero, g["z_min"], g["z_max"])
    assert rel_err(out["x_hist"][0].cpu().numpy(), g["nopush_x_hist"][:, :, 0]) < REL
    assert rel_err(out["y_hist"][0].cpu().numpy(), g["nopush_y_hist"][:, :, 0]) < REL


# ----------------------------------------------------------------------------- strict mode
def test_strict_gpu_plans_are_kkt_points_of_the_reference_qp(engine, golden, solver):
    """The ZMP plans the FP64 kernel returns, mapped back to jerks and checked in the reference's literal QP
    (zmp_controller.py:174-189: H = R I, A_ineq = [Pu; -Pu]): primal feasible, multipliers vanish on free rows and have
    the right sign on active ones.  Independent of both oracle solvers and of the z-coordinates the kernel works in."""
    g = golden("oracle_strict_ticks")
    rng = np.random.default_rng(17)
    for N in (16, 50, 150):
        plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, solver=solver)
        p = orc.Problem.make(N)
        x, lo, hi = g[f"N{N}/x"], g[f"N{N}/z_min"], g[f"N{N}/z_max"]
        # golden ticks plus fresh random ones the oracle never saw (narrow boxes: many active rows)
        xr = rng.normal(size=(16, 3)) * [0.05, 0.1, 0.5]
        centre = rng.uniform(-0.05, 0.05, size=(16, 1)) + 0.03 * np.sign(rng.normal(size=(16, N))).cumsum(axis=1) / np.sqrt(N)
        x, lo, hi = np.vstack([x, xr]), np.vstack([lo, centre - 0.02]), np.vstack([hi, centre + 0.02])
        res = plan.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
        assert (res["status"].cpu().numpy() == 0).all()
        z, nxt = res["zmp_plan"].cpu().numpy(), res["next_state"].cpu().numpy()
        active = 0
        for k in range(len(x)):
            cert = orc.kkt_certificate_jerk(p, x[k], hi[k], lo[k], z[k], bound_tol=0.0)
            assert cert["primal"] == 0.0, (N, k, cert)                       # the kernel clips: feasible to the last bit
            assert cert["free_multiplier"] <= REL and cert["wrong_sign"] <= REL, (N, k, cert)     # north_star: 1e-9 relative
            assert np.max(np.abs(p.A @ x[k] + p.B * cert["jerk0"] - nxt[k])) <= 1e-9 * max(1.0, np.abs(nxt[k]).max())
            active += cert["active_rows"]
        assert active > 10 * N // 16


@pytest.mark.parametrize("N", [16, 50, 150])
def test_strict_ticks_match_exact_oracle(engine, golden, N, solver):
    g = golden("oracle_strict_ticks")
    plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, solver=solver)
    x, lo, hi = g[f"N{N}/x"], g[f"N{N}/z_min"], g[f"N{N}/z_max"]
    res = plan.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    assert (res["status"].cpu().numpy() == 0).all()
    z = res["zmp_plan"].cpu().numpy()
    assert np.max(np.abs(z - g[f"N{N}/z"])) < 1e-10
    assert np.max(np.abs(res["next_state"].cpu().numpy() - g[f"N{N}/next"])) < 1e-9
    assert (z >= lo).all() and (z <= hi).all()                      # exactly feasible: clipped values
    # active set by equality; tie = oracle row free but within 1e-9 of a bound
    act = np.where(z == lo, -1, 0) + np.where(z == hi, 1, 0)
    want = g[f"N{N}/active"]
    free_margin = np.minimum(g[f"N{N}/z"] - lo, hi - g[f"N{N}/z"])
    decided = (want != 0) | (free_margin > 1e-9)
    assert np.array_equal(act[decided], want[decided])
    # jerk plan consistent with the ZMP plan
    p = orc.Problem.make(N)
    jerk = res["jerk_plan"].cpu().numpy()
    for k in range(0, len(x), 5):
        assert np.max(np.abs(p.Pu @ jerk[k] + p.Px @ x[k] - z[k])) < 1e-11


def check_flags(flags, want_flags, margin_of_got, engine, tie=1e-9):
    """Flags must be bit-exact except at ties: samples whose realised ZMP is within `tie` of a
    bound, where 'on the bound' vs 'just inside' is decided by the last bits of the solve."""
    diff = flags != want_flags
    assert not (diff & (np.abs(margin_of_got) > tie)).any()
    return int(diff.sum())


def test_strict_default_walk_matches_exact_oracle(engine, plan150, golden, solver):
    g = golden("oracle_strict_walk")
    zero = np.zeros((1, 3))
    out = plan150.rollout(engine.MODE_STRICT, zero, zero, g["z_min"], g["z_max"],
                          impulse=[float(g["impulse"])], force_tick=[int(g["force_tick"])])
    assert (out["status"].cpu().numpy() == 0).all()
    x_hist, y_hist = out["x_hist"][0].cpu().numpy(), out["y_hist"][0].cpu().numpy()
    assert rel_err(x_hist, g["x_hist"]) < REL
    assert rel_err(y_hist, g["y_hist"]) < REL
    p = orc.Problem.make(150)
    zmp = np.stack([x_hist @ p.C, y_hist @ p.C], 1)
    margin = np.minimum(zmp - g["z_min"], g["z_max"] - zmp)
    assert margin.min() > -1e-12                                    # strict: ZMP stays in the box
    flags = out["flags"][0].cpu().numpy()
    ties = check_flags(flags, g["flags"], margin, engine)
    assert ties <= 8
    assert out["max_violation"].max().item() < 1e-12
    iters = out["iters"][0].cpu().numpy()
    if solver == "admm":          # ADMM iterations
        assert 5 * 419 < iters[0] < 80 * 419 and 5 * 419 < iters[1] < 80 * 419
    else:                         # sweep pairs of the active-set iteration: the guess is right on most ticks
        assert 419 <= iters[0] < 1.2 * 419 and 419 <= iters[1] < 2 * 419


def test_strict_equals_unconstrained_when_box_is_wide(engine, plan150):
    rng = np.random.default_rng(0)
    B, n = 40, 60
    x0 = rng.normal(size=(B, 3)) * [0.02, 0.05, 0.1]
    y0 = rng.normal(size=(B, 3)) * [0.02, 0.05, 0.1]
    lo, hi = np.full((n, 2), -50.0), np.full((n, 2), 50.0)
    a = plan150.rollout(engine.MODE_STRICT, x0, y0, lo, hi)
    a = {k: v.cpu().numpy() for k, v in a.items() if not k.startswith("_")}
    b = plan150.rollout(engine.MODE_UNCONSTRAINED, x0, y0, lo, hi)
    b = {k: v.cpu().numpy() for k, v in b.items() if not k.startswith("_")}
    assert (a["status"] == 0).all() and not a["flags"].any()
    assert rel_err(a["x_hist"], b["x_hist"]) < REL and rel_err(a["y_hist"], b["y_hist"]) < REL


def push_sweep_inputs(B, seed=0):
    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.generators import CoPGenerator
    cfg = MPCConfig.from_json(os.path.join(REPO, "configs", "default.json"))
    z_max, z_min, _ = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
    rng = np.random.default_rng(seed)
    F = rng.uniform(0.0, 1000.0, B)
    tick = rng.integers(100, 320, B).astype(np.int32)
    return cfg, z_max, z_min, cfg.dt * F / cfg.m, tick


def test_batch_results_do_not_depend_on_batch_layout(engine, plan150):
    """A scenario's output is bit-identical whatever batch / tile / position it is solved in
    (multi-GPU row: sharding must not change results)."""
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(37)
    n = 90                                     # first 90 samples: standing + first steps, short test
    z_max, z_min = z_max[:n], z_min[:n]
    tick = (tick % 60 + 10).astype(np.int32)
    zero = np.zeros((37, 3))
    full = plan150.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, impulse=impulse, force_tick=tick)
    full = {k: v.cpu().numpy() for k, v in full.items() if not k.startswith("_")}
    pick = np.array([36, 5, 20])
    sub = plan150.rollout(engine.MODE_STRICT, zero[pick], zero[pick], z_min, z_max, impulse=impulse[pick],
                          force_tick=tick[pick])
    sub = {k: v.cpu().numpy() for k, v in sub.items() if not k.startswith("_")}
    for key in ("x_hist", "y_hist", "flags", "iters", "max_violation", "status"):
        assert np.array_equal(sub[key], full[key][pick]), key
    # per-scenario bounds (B,n,2) give the same bits as the shared (n,2) series
    per = plan150.rollout(engine.MODE_STRICT, zero[pick], zero[pick], np.repeat(z_min[None], 3, 0),
                          np.repeat(z_max[None], 3, 0), impulse=impulse[pick], force_tick=tick[pick])
    # (the default FP64 solver starts shared series on its row-parallel kernel and per-scenario series on the Riccati
    # kernel: two formulations of the same optimum, equal to rounding amplified by the walk, not to the bit)
    if plan150.solver == engine.SOLVER_ADMM:
        assert np.array_equal(per["y_hist"].cpu().numpy(), sub["y_hist"])
    else:
        assert rel_err_mm(per["y_hist"].cpu().numpy(), sub["y_hist"]) < REL
        plan150.set_row_parallel(False)
        try:
            sub2 = plan150.rollout(engine.MODE_STRICT, zero[pick], zero[pick], z_min, z_max, impulse=impulse[pick], force_tick=tick[pick])
            assert np.array_equal(per["y_hist"].cpu().numpy(), sub2["y_hist"].cpu().numpy())
        finally:
            plan150.set_row_parallel(True)


def test_tick_ranges_do_not_change_results(engine):
    """Load balancing hands a unit's walk out as consecutive tick ranges that different solver groups
    continue (iterate, axis state and running outcomes travel through device memory).  Every split gives
    the bits of the unsplit walk, ragged lengths and an aborted column included."""
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(29, seed=5)
    n = 140
    z_max, z_min = z_max[:n], z_min[:n]
    n_steps = np.full(29, n, np.int32)
    n_steps[[3, 11, 12]] = [35, 2, 97]
    tick = (tick % 100 + 10).astype(np.int32)
    zero = np.zeros((29, 3))
    # 12 iterations are enough before the push and too few after a strong one: those columns stop at their
    # first QP without a solution, in whatever range that happens
    plan = engine.get_plan(150, 0.75, 9.81, 0.01, 1.0, 1e-6, max_iter=12, eps=1e-9, solver="admm")
    results = {}
    for segments in (1, 2, 3, 4):
        plan.set_tick_ranges(segments)                  # lipm_set_option(LIPM_OPT_TICK_RANGES)
        out = plan.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, n_steps=n_steps, impulse=impulse, force_tick=tick,
                           want_readout=True)
        results[segments] = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
    plan.set_tick_ranges(0)
    failed = results[1]["status"].any(axis=1)
    assert failed.any() and not failed.all(), failed
    for segments in (2, 3, 4):
        for key in ("x_hist", "y_hist", "flags", "iters", "max_violation", "status", "com", "zmp", "lean_max", "outcome"):
            assert np.array_equal(results[segments][key], results[1][key]), (segments, key)


def test_readout_and_outcomes_come_from_the_walk_kernels(engine, plan150, plan150_f16):
    """com, the realised ZMP  C . state  (scripts/run_mpc.py:294), the per-axis maxima and the packed per-scenario
    record are written by the walk kernels (every mode / precision), ragged walks included: equal to the read-out
    the reference's drivers compute on the host from the state histories."""
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(21, seed=9)
    n = 130
    z_max, z_min = z_max[:n], z_min[:n]
    n_steps = np.full(21, n, np.int32)
    n_steps[[2, 7]] = [50, 2]
    tick = (tick % 90 + 10).astype(np.int32)
    impulse[20] *= 6.0                                             # one scenario that falls
    zero = np.zeros((21, 3))
    C = np.array([1.0, 0.0, -cfg.h / cfg.g])
    for plan, mode in ((plan150, engine.MODE_STRICT), (plan150_f16, engine.MODE_STRICT), (plan150, engine.MODE_UNCONSTRAINED)):
        out = plan.rollout(mode, zero, zero, z_min, z_max, n_steps=n_steps, impulse=impulse, force_tick=tick, want_readout=True,
                           fall_lean=0.12)
        out = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
        rec = engine.unpack_outcome_host(out["outcome"])
        for b in range(21):
            nb = int(n_steps[b])
            xh, yh = out["x_hist"][b, :nb], out["y_hist"][b, :nb]
            assert np.array_equal(out["com"][b, :nb], np.stack([xh[:, 0], yh[:, 0]], 1))
            zmp = np.stack([np.tensordot(xh, C, axes=([1], [0])), np.tensordot(yh, C, axes=([1], [0]))], 1)
            # the push changes the y velocity after the sample's ZMP was taken; velocity has weight 0 in C
            assert np.max(np.abs(out["zmp"][b, :nb] - zmp)) <= 1e-15 * max(1.0, np.abs(zmp).max())
            assert not out["com"][b, nb:].any() and not out["zmp"][b, nb:].any()
            lean = np.abs(out["com"][b, :nb] - out["zmp"][b, :nb]).max(axis=0) if nb > 1 else np.zeros(2)
            assert np.array_equal(out["lean_max"][b], lean), (b, out["lean_max"][b], lean)
            assert rec["max_violation"][b] == out["max_violation"][b].max()
            assert rec["lean_max"][b] == lean.max() and bool(rec["fell"][b]) == bool(lean.max() > 0.12)
            assert rec["status"][b] == out["status"][b].sum() and rec["iters"][b] == out["iters"][b].sum()
        assert rec["fell"][20] and not rec["fell"][:8].all()


def test_one_plan_on_two_streams(engine, plan150):
    """A plan is re-entrant: two walks enqueued on different streams of the same plan (every launch carries its own
    work counter) give the bits of the same walks run one after the other."""
    import torch

    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(1200, seed=11)
    n = 120
    z_max, z_min = z_max[:n], z_min[:n]
    tick = (tick % 90 + 10).astype(np.int32)
    dev = plan150.device
    halves = []
    for lo, hi in ((0, 600), (600, 1200)):
        halves.append(dict(x0=torch.zeros((hi - lo, 3), dtype=torch.float64, device=dev),
                           zmin=torch.as_tensor(z_min).to(dev), zmax=torch.as_tensor(z_max).to(dev),
                           imp=torch.as_tensor(impulse[lo:hi]).to(dev), tick=torch.as_tensor(tick[lo:hi]).to(dev)))
    serial = [plan150.rollout(engine.MODE_STRICT, h["x0"], h["x0"], h["zmin"], h["zmax"], impulse=h["imp"], force_tick=h["tick"])
              for h in halves]
    torch.cuda.synchronize()
    serial = [{k: v.clone() for k, v in s.items() if not k.startswith("_")} for s in serial]
    for attempt in range(3):
        streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        both = []
        for h, st in zip(halves, streams):
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                both.append(plan150.rollout(engine.MODE_STRICT, h["x0"], h["x0"], h["zmin"], h["zmax"], impulse=h["imp"],
                                            force_tick=h["tick"]))
        torch.cuda.synchronize()
        for got, want in zip(both, serial):
            for key in ("x_hist", "y_hist", "flags", "iters", "max_violation", "status", "outcome"):
                assert torch.equal(got[key], want[key]), (attempt, key)


def test_row_parallel_stage_agrees_with_riccati_kernel(engine):
    """The default FP64 solver starts walks on shared series on the row-parallel kernel (dense corrections with the
    inverse Hessian), parks columns whose set outgrows it on the Riccati kernel and takes them back when the push is
    absorbed.  The same sweep with that stage switched off runs on the Riccati kernel alone: both are the optimum of the
    same QPs, so the walks agree to 1e-9 relative, the flags bit for bit, and the outcome verdicts are the same --
    including scenarios that fall (parked for good) and ones that recover (parked and handed back)."""
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(384, seed=5)
    plan = engine.get_plan(150, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
    zero = np.zeros((384, 3))
    runs = {}
    for on in (True, False):
        plan.set_row_parallel(on)
        try:
            out = plan.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, impulse=impulse, force_tick=tick, want_readout=True)
            runs[on] = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
            if on:
                st = plan.stats()
        finally:
            plan.set_row_parallel(True)
    a, b = runs[True], runs[False]
    assert st["columns_parked"] > 50 and st["columns_handed_back"] > 10
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.array_equal(a["flags"], b["flags"])
    for key in ("x_hist", "y_hist", "zmp", "com"):
        scale = np.maximum(1.0, np.abs(b[key]).reshape(384, -1).max(axis=1))
        diff = np.abs(a[key] - b[key]).reshape(384, -1).max(axis=1)
        assert (diff / scale).max() < REL, key
    fell_a, fell_b = engine.unpack_outcome_host(a["outcome"])["fell"], engine.unpack_outcome_host(b["outcome"])["fell"]
    assert np.array_equal(fell_a, fell_b) and 0 < fell_a.sum() < 384
    assert a["max_violation"].max() < 1e-9
    # indexed series of different lengths (the parameter sweep's layout): same agreement
    nA, nB = 300, 260
    zmin2 = np.stack([z_min[:nA], np.concatenate([z_min[40:40 + nB], np.repeat(z_min[39 + nB:40 + nB], nA - nB, 0)])])
    zmax2 = np.stack([z_max[:nA], np.concatenate([z_max[40:40 + nB], np.repeat(z_max[39 + nB:40 + nB], nA - nB, 0)])])
    idx = (np.arange(96) % 2).astype(np.int32)
    n_steps = np.where(idx == 0, nA, nB).astype(np.int32)
    n_steps[5] = 200                                   # a walk shorter than its series: falls back to the Riccati kernel
    zero = np.zeros((96, 3))
    runs = {}
    for on in (True, False):
        plan.set_row_parallel(on)
        try:
            out = plan.rollout(engine.MODE_STRICT, zero, zero, zmin2, zmax2, n_steps=n_steps, bounds_index=idx,
                               impulse=impulse[:96] * 0.6, force_tick=(tick[:96] % 150).astype(np.int32))
            runs[on] = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
        finally:
            plan.set_row_parallel(True)
    a, b = runs[True], runs[False]
    assert np.array_equal(a["flags"], b["flags"]) and (a["status"] == 0).all()
    assert rel_err_mm(a["y_hist"], b["y_hist"]) < REL and rel_err_mm(a["x_hist"], b["x_hist"]) < REL


def test_walk_chunks_on_side_streams_are_bit_identical(engine):
    """LIPM_OPT_WALK_CHUNKS: batches of 4096 scenarios or more are cut into four chunks of the processing order whose
    Riccati / hand-back stages run on side streams of the library beside the row-parallel passes of the later chunks.
    A column does the same arithmetic wherever it runs: every output equals the serial run (one chunk, caller's stream
    only) bit for bit: four chunks on 1024 scenarios, forced chunk counts on 97 scenarios (chunks of unequal size), and three
    chunks on the parameter sweep's layout (indexed series of different lengths)."""
    import torch
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(1024, seed=11)
    n = 160
    z_max, z_min = z_max[:n], z_min[:n]
    tick = (tick % 90 + 30).astype(np.int32)
    plan = engine.get_plan(150, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)

    def run(b):
        zero = np.zeros((b, 3))
        out = plan.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, impulse=impulse[:b], force_tick=tick[:b], want_readout=True)
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}, plan.stats()

    try:
        plan.set_walk_chunks(1)
        base, st1 = run(1024)
        assert (base["status"] == 0).all() and st1["columns_parked"] > 100
        plan.set_walk_chunks(4)
        got, st0 = run(1024)
        assert st0 == st1
        for key in base:
            assert np.array_equal(got[key], base[key]), key
        plan.set_walk_chunks(1)
        base, _ = run(97)
        for chunks in (2, 3, 4):
            plan.set_walk_chunks(chunks)
            got, _ = run(97)
            for key in base:
                assert np.array_equal(got[key], base[key]), (chunks, key)
        # indexed series of different lengths (the parameter sweep's layout)
        nA, nB = 160, 130
        zmin2 = np.stack([z_min[:nA], np.concatenate([z_min[20:20 + nB], np.repeat(z_min[19 + nB:20 + nB], nA - nB, 0)])])
        zmax2 = np.stack([z_max[:nA], np.concatenate([z_max[20:20 + nB], np.repeat(z_max[19 + nB:20 + nB], nA - nB, 0)])])
        idx = (np.arange(203) % 2).astype(np.int32)
        n_steps = np.where(idx == 0, nA, nB).astype(np.int32)
        zero = np.zeros((203, 3))

        def run_indexed():
            out = plan.rollout(engine.MODE_STRICT, zero, zero, zmin2, zmax2, n_steps=n_steps, bounds_index=idx,
                               impulse=impulse[:203] * 0.8, force_tick=(tick[:203] % 80 + 20).astype(np.int32), want_readout=True)
            torch.cuda.synchronize()
            return {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}, plan.stats()

        plan.set_walk_chunks(1)
        base, st1 = run_indexed()
        assert st1["columns_parked"] > 10 and (base["status"] == 0).all()
        plan.set_walk_chunks(3)
        got, st3 = run_indexed()
        assert st3 == st1
        for key in base:
            assert np.array_equal(got[key], base[key]), ("indexed", key)
    finally:
        plan.set_walk_chunks(0)


def test_riccati_kernel_options_do_not_change_the_optimum(engine):
    """Launch-shape and placement options of the thread-per-column kernel (columns per warp, stage tables and
    feed-forward terms in shared memory or through L1 / global scratch) do not change a bit; few-pin sweeps (table solve +
    dense correction instead of the Riccati recursion when at most 16 rows are pinned: LIPM_OPT_FEW_PIN_ROWS, off by
    default) are another formulation of the same optimum: flags bit for bit, walks to 1e-9."""
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(48, seed=9)
    n = 200
    z_max, z_min = z_max[:n], z_min[:n]
    tick = (tick % 100 + 40).astype(np.int32)
    plan = engine.get_plan(150, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R)
    zero = np.zeros((48, 3))

    def run():
        out = plan.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, impulse=impulse * 0.7, force_tick=tick)
        return {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}

    plan.set_row_parallel(False)
    try:
        base = run()
        assert (base["status"] == 0).all()
        for lanes, scratch in ((4, 0), (32, 0), (8, 2), (8, 1)):
            plan.set_riccati_lanes(lanes)
            plan.set_riccati_scratch(scratch)
            got = run()
            for key in ("x_hist", "y_hist", "flags", "iters"):
                assert np.array_equal(got[key], base[key]), (lanes, scratch, key)
        plan.set_riccati_lanes(0)
        plan.set_riccati_scratch(0)
        plan.set_few_pin_rows(16)
        few = run()
        assert np.array_equal(few["flags"], base["flags"]) and (few["status"] == 0).all()
        assert rel_err_mm(few["y_hist"], base["y_hist"]) < REL and rel_err_mm(few["x_hist"], base["x_hist"]) < REL
    finally:
        plan.set_few_pin_rows(0)
        plan.set_riccati_lanes(0)
        plan.set_riccati_scratch(0)
        plan.set_row_parallel(True)


def test_ragged_walks(engine, plan150):
    """n_steps per scenario: each walk equals the same walk run alone on its truncated series;
    the last bound row is what gets repeated past the end (zmp_controller.py:81-88)."""
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(5, seed=3)
    n_steps = np.array([70, 2, 64, 33, 70], np.int32)
    tick = (n_steps // 2).astype(np.int32)
    zero = np.zeros((5, 3))
    for mode in (engine.MODE_STRICT, engine.MODE_UNCONSTRAINED):
        out = plan150.rollout(mode, zero, zero, z_min[:70], z_max[:70], n_steps=n_steps, impulse=impulse, force_tick=tick)
        out = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
        for b, n in enumerate(n_steps):
            one = plan150.rollout(mode, zero[:1], zero[:1], z_min[:n], z_max[:n], impulse=impulse[b:b + 1],
                                  force_tick=tick[b:b + 1])
            if mode == engine.MODE_STRICT and plan150.solver == engine.SOLVER_ACTIVE_SET:
                # ragged walks on a shared series run on the Riccati kernel, the single walk on the row-parallel one:
                # two formulations of the same optimum (bit for bit with the row-parallel stage off, below)
                assert rel_err_mm(one["y_hist"][0].cpu().numpy(), out["y_hist"][b, :n]) < REL, (mode, b)
                assert rel_err_mm(one["x_hist"][0].cpu().numpy(), out["x_hist"][b, :n]) < REL
            else:
                assert np.array_equal(one["y_hist"][0].cpu().numpy(), out["y_hist"][b, :n]), (mode, b)
                assert np.array_equal(one["x_hist"][0].cpu().numpy(), out["x_hist"][b, :n])
            assert np.array_equal(one["flags"][0].cpu().numpy(), out["flags"][b, :n])
            assert not out["y_hist"][b, n:].any()              # untouched past the end
    if plan150.solver == engine.SOLVER_ACTIVE_SET:
        plan150.set_row_parallel(False)
        try:
            out = plan150.rollout(engine.MODE_STRICT, zero, zero, z_min[:70], z_max[:70], n_steps=n_steps, impulse=impulse, force_tick=tick)
            for b, n in enumerate(n_steps):
                one = plan150.rollout(engine.MODE_STRICT, zero[:1], zero[:1], z_min[:n], z_max[:n], impulse=impulse[b:b + 1],
                                      force_tick=tick[b:b + 1])
                assert np.array_equal(one["y_hist"][0].cpu().numpy(), out["y_hist"][b, :n].cpu().numpy()), b
        finally:
            plan150.set_row_parallel(True)


def test_unconstrained_large_ragged_batch(engine, plan150):
    """Unconstrained mode with per-scenario lengths: the preview term is computed per column, in chunks of 65,535
    columns -- scenarios on both sides of a chunk boundary equal the same walk run alone (shared-series path)."""
    cfg, z_max, z_min, _, _ = push_sweep_inputs(4, seed=1)
    n, B = 40, 40000                                  # 80,000 columns: two chunks
    z_max, z_min = z_max[100:100 + n], z_min[100:100 + n]
    rng = np.random.default_rng(8)
    x0 = rng.normal(size=(B, 3)) * [0.01, 0.02, 0.05]
    y0 = rng.normal(size=(B, 3)) * [0.01, 0.02, 0.05]
    n_steps = rng.integers(2, n + 1, B).astype(np.int32)
    impulse = rng.uniform(0, 0.2, B)
    tick = (n_steps // 2).astype(np.int32)
    out = plan150.rollout(engine.MODE_UNCONSTRAINED, x0, y0, z_min, z_max, n_steps=n_steps, impulse=impulse, force_tick=tick)
    out = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
    for b in (0, 17, 25534, 32767, 32768, 39999):      # x columns 0..39999, y columns 40000..79999: chunk edge at 65535
        nb = int(n_steps[b])
        one = plan150.rollout(engine.MODE_UNCONSTRAINED, x0[b:b + 1], y0[b:b + 1], z_min[:nb], z_max[:nb],
                              impulse=impulse[b:b + 1], force_tick=tick[b:b + 1])
        for key in ("x_hist", "y_hist", "flags"):
            assert np.array_equal(one[key][0].cpu().numpy(), out[key][b, :nb]), (b, key)
        assert not out["y_hist"][b, nb:].any()
        assert np.array_equal(one["max_violation"][0].cpu().numpy(), out["max_violation"][b])


def test_layout_fuzz(engine):
    """Randomized batches (size 1..5000, ragged lengths, shared or per-scenario bounds, random pushes): every sampled
    scenario equals the same walk solved alone, bit for bit, in all three modes (tools/fuzz_layout.py)."""
    import subprocess
    import sys as _sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_layout.py")
    run = subprocess.run([_sys.executable, tool, "8", "3"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "all trials ok" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]


def test_push_sweep_full_size_properties(engine, plan150):
    """BASELINE config 2 at full size (4096 scenarios, N=150, strict): size-independent properties."""
    B = 4096
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(B)
    impulse[:8] = 0.0                                            # eight scenarios without push
    zero = np.zeros((B, 3))
    out = plan150.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, impulse=impulse, force_tick=tick)
    status = out["status"].cpu().numpy()
    assert (status == 0).all()
    assert out["max_violation"].max().item() < 1e-9             # strict keeps the ZMP in the box
    y_hist = out["y_hist"].cpu().numpy()
    x_hist = out["x_hist"].cpu().numpy()
    assert np.isfinite(y_hist).all() and np.isfinite(x_hist).all()
    # the push acts on y only: every x trajectory is the same walk, bit for bit
    assert (x_hist == x_hist[0]).all()
    # scenarios without push coincide; identical (F, tick) pairs coincide
    assert (y_hist[:8] == y_hist[0]).all()
    # before its push every scenario follows the push-free walk exactly
    for b in (8, 100, 4095):
        assert np.array_equal(y_hist[b, : tick[b] + 1], y_hist[0, : tick[b] + 1])
        assert not np.array_equal(y_hist[b, tick[b] + 1], y_hist[0, tick[b] + 1])
    # spot-check two scenarios against the oracle (ADMM, FP64)
    p = orc.Problem.make(150)
    b = 100
    want = orc.rollout_wieber(p, np.zeros(3), np.zeros(3), z_max, z_min, strict=True, impulse=impulse[b],
                              force_tick=int(tick[b]), solver="admm", n_ticks=int(tick[b]) + 12)
    m = len(want["y_hist"])
    assert rel_err(y_hist[b, :m], want["y_hist"]) < REL


def test_errors_are_loud(engine):
    plan = engine.get_plan(700, 0.75, 9.81, 1.5 / 700, 1.0, 1e-6)
    zero = np.zeros((2, 3))
    lo, hi = -np.ones((400, 2)), np.ones((400, 2))
    with pytest.raises(engine.LipmError, match="horizon 700"):
        plan.rollout(engine.MODE_STRICT, zero, zero, lo, hi)
    out = plan.rollout(engine.MODE_UNCONSTRAINED, zero, zero, lo, hi)       # unconstrained has no such limit
    assert np.isfinite(out["x_hist"].cpu().numpy()).all()
    with pytest.raises(ValueError):
        plan.rollout(engine.MODE_UNCONSTRAINED, zero, zero, lo[:, :1], hi[:, :1])
    with pytest.raises(engine.LipmError):
        engine.LipmPlan(150, 0.75, 9.81, None, 1.0, -1.0)
    with pytest.raises(engine.LipmError):
        engine.LipmPlan(1, 0.75, 9.81, None, 1.0, 1e-6)


def test_reference_python_surface_on_gpu(golden):
    """The reference's own call sequence (scripts/run_mpc.py:273-294) through the mirrored API."""
    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.controllers import ZMPController
    from mpc_bipedal.generators import CoPGenerator

    cfg = MPCConfig.from_json(os.path.join(REPO, "configs", "default.json"))
    z_max, z_min, _ = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
    ctrl = ZMPController(cfg)
    zero = np.zeros((3, 1))
    com, y_hist = ctrl.generate_com_trajectory(zero, zero.copy(), z_max, z_min)
    g = golden("oracle_strict_walk")
    assert com.shape == (420, 2) and y_hist.shape == (420, 3, 1)
    assert rel_err(y_hist[:, :, 0], g["y_hist"]) < REL
    zmp_y = np.tensordot(y_hist[:, :, 0], ctrl.C, axes=([1], [0]))
    assert (zmp_y >= z_min[:, 1] - 1e-12).all() and (zmp_y <= z_max[:, 1] + 1e-12).all()
    # one tick, both modes, reference signature: (3,1), N, (N,1), (N,1) -> (3,1)
    nxt = ctrl.predict_wieber_axis(zero, 150, z_max[1:151, 1:2], z_min[1:151, 1:2])
    assert nxt.shape == (3, 1) and np.max(np.abs(nxt[:, 0] - g["y_hist"][1])) < 1e-10
    cfg.strict = False
    gn = golden("ref_nonstrict_walk")
    nxt = ctrl.predict_wieber_axis(zero, 150, z_max[1:151, 1:2], z_min[1:151, 1:2])
    assert np.max(np.abs(nxt[:, 0] - gn["y_hist"][1, :, 0])) < 1e-12
    x_state, y_state = ctrl.generate_state_trajectory_wieber(zero, zero.copy(), z_max, z_min)
    assert x_state.shape == (420, 3, 1)
    assert rel_err(x_state, gn["nopush_x_hist"]) < REL and rel_err(y_state, gn["nopush_y_hist"]) < REL


@pytest.mark.parametrize("N", [24, 40, 100, 128, 145, 153, 160, 161, 200, 300, 600])
def test_strict_ticks_other_horizons_and_determinism(engine, N, solver):
    """Every kernel variant (row blocks per warp 1..5, partial last row group, padded rows; resident
    matrix up to N=160, L2-streamed matrix with 8- and 16-warp groups beyond -- BASELINE config 3):
    result equals the oracle's ADMM and repeated launches are bit-identical (no races)."""
    rng = np.random.default_rng(11 + N)
    plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, solver=solver)
    p = orc.Problem.make(N)
    cases = 40
    x = rng.normal(size=(cases, 3)) * np.array([0.05, 0.1, 0.5])
    centre = rng.uniform(-0.05, 0.05, size=(cases, 1)) + np.cumsum(
        (rng.random((cases, N)) < 0.04) * rng.uniform(-0.1, 0.1, size=(cases, N)), axis=1)
    lo, hi = centre - 0.055, centre + 0.055
    first = plan.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    z = first["zmp_plan"].cpu().numpy()
    assert (first["status"].cpu().numpy() == 0).all()
    M = orc.admm_matrix(p, 2.0)
    for k in range(0, cases, 3):
        _, want, _, _ = orc.strict_axis_admm(p, x[k], hi[k], lo[k], eps=1e-13, M=M)
        assert np.max(np.abs(z[k] - want)) < 1e-10, (N, k)
    for _ in range(3):
        again = plan.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
        assert np.array_equal(again["zmp_plan"].cpu().numpy(), z)
        assert np.array_equal(again["iters"].cpu().numpy(), first["iters"].cpu().numpy())
    # a permuted batch gives each problem the same bits
    perm = rng.permutation(cases)
    shuffled = plan.solve_tick(engine.MODE_STRICT, x[perm], lo[perm], hi[perm], want_plan=True)
    assert np.array_equal(shuffled["zmp_plan"].cpu().numpy(), z[perm])


# ----------------------------------------------------------------------------- reduced precision (tcgen05, FP16 x3)
TOL_REDUCED_M = 1e-4     # north_star: reduced-precision path within 1e-4 m of the reference trajectories


@pytest.fixture(scope="module")
def plan150_f16(engine):
    return engine.get_plan(150, 0.75, 9.81, 1.5 / 150, 1.0, 1e-6, precision=engine.PRECISION_F16X3)


@pytest.mark.parametrize("N", [16, 50, 150])
def test_reduced_ticks_match_exact_oracle(engine, golden, N):
    g = golden("oracle_strict_ticks")
    plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, precision=engine.PRECISION_F16X3)
    x, lo, hi = g[f"N{N}/x"], g[f"N{N}/z_min"], g[f"N{N}/z_max"]
    res = plan.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    assert (res["status"].cpu().numpy() == 0).all()
    z = res["zmp_plan"].cpu().numpy()
    assert np.max(np.abs(z - g[f"N{N}/z"])) < TOL_REDUCED_M
    assert (z >= lo).all() and (z <= hi).all()                      # clamped to the FP64 box
    # the state update is FP64 on the FP32 solution: position / velocity / acceleration of the next sample
    nxt = res["next_state"].cpu().numpy()
    assert np.max(np.abs(nxt[:, 0] - g[f"N{N}/next"][:, 0])) < TOL_REDUCED_M
    # active set, ties = oracle row free but within 1e-5 of a bound (FP32 iterate, eps 1e-5)
    act = np.where(z == lo, -1, 0) + np.where(z == hi, 1, 0)
    want = g[f"N{N}/active"]
    free_margin = np.minimum(g[f"N{N}/z"] - lo, hi - g[f"N{N}/z"])
    decided = ((want != 0) & (np.abs(z - g[f"N{N}/z"]) < 1e-12)) | (free_margin > 1e-5)
    assert np.array_equal(act[decided], want[decided])
    # repeated launches are bit-identical, and a permuted batch gives each problem the same bits
    again = plan.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    assert np.array_equal(again["zmp_plan"].cpu().numpy(), z)
    perm = np.random.default_rng(5).permutation(len(x))
    shuffled = plan.solve_tick(engine.MODE_STRICT, x[perm], lo[perm], hi[perm], want_plan=True)
    assert np.array_equal(shuffled["zmp_plan"].cpu().numpy(), z[perm])
    assert np.array_equal(shuffled["iters"].cpu().numpy(), res["iters"].cpu().numpy()[perm])


def test_reduced_default_walk_matches_exact_oracle(engine, plan150_f16, golden):
    g = golden("oracle_strict_walk")
    zero = np.zeros((1, 3))
    out = plan150_f16.rollout(engine.MODE_STRICT, zero, zero, g["z_min"], g["z_max"],
                              impulse=[float(g["impulse"])], force_tick=[int(g["force_tick"])])
    assert (out["status"].cpu().numpy() == 0).all()
    x_hist, y_hist = out["x_hist"][0].cpu().numpy(), out["y_hist"][0].cpu().numpy()
    assert np.max(np.abs(x_hist[:, 0] - g["x_hist"][:, 0])) < TOL_REDUCED_M      # CoM trajectory
    assert np.max(np.abs(y_hist[:, 0] - g["y_hist"][:, 0])) < TOL_REDUCED_M
    p = orc.Problem.make(150)
    zmp = np.stack([x_hist @ p.C, y_hist @ p.C], 1)
    want_zmp = np.stack([g["x_hist"] @ p.C, g["y_hist"] @ p.C], 1)
    assert np.max(np.abs(zmp - want_zmp)) < TOL_REDUCED_M                        # ZMP trajectory
    margin = np.minimum(zmp - g["z_min"], g["z_max"] - zmp)
    assert margin.min() > -1e-12                   # the applied sample is clamped in FP64: still strictly feasible
    assert out["max_violation"].max().item() < 1e-12
    flags = out["flags"][0].cpu().numpy()
    viol_bits = engine.FLAG_VIOL_LO | engine.FLAG_VIOL_HI | engine.FLAG_NOT_CONVERGED
    assert np.array_equal(flags & viol_bits, g["flags"] & viol_bits)             # box-violation flags bit-exact
    # active-set bits: exact except where the oracle's margin to the bound is below the FP32 resolution
    diff = (flags != g["flags"])
    assert not (diff & (np.abs(margin) > 1e-5)).any()
    iters = out["iters"][0].cpu().numpy()
    assert iters.sum() < 30 * 838                  # far fewer iterations than the FP64 path's eps = 1e-12


def test_reduced_push_sweep_full_size(engine, plan150, plan150_f16):
    """BASELINE config 2 at full size on the reduced-precision path, against the FP64 path of the same library."""
    B = 4096
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(B)
    impulse[:8] = 0.0
    zero = np.zeros((B, 3))
    ref = plan150.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, impulse=impulse, force_tick=tick)
    ref = {k: v.cpu().numpy() for k, v in ref.items() if not k.startswith("_")}
    out = plan150_f16.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, impulse=impulse, force_tick=tick)
    out = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
    assert (out["status"] == 0).all()
    assert out["max_violation"].max() < 1e-9
    # Conditioning.  While a push keeps the ZMP saturated on the edge of the sole the loop is open and the
    # CoM is the unstable pendulum (time constant sqrt(h/g) = 0.28 s): ANY difference between two solvers --
    # the FP64 path's own stop tolerance included -- is amplified by e^(t/0.28 s) until the walk has
    # recovered or fallen.  The 1e-4 m bound is therefore asserted (a) over the whole walk for every scenario
    # that absorbs its push with margin (pendulum lean |CoM - ZMP| < 0.12 m on the FP64 path; > half of the
    # sweep) and (b) for EVERY scenario up to 1.5 s after its push; beyond that, borderline and falling
    # scenarios are compared through their outcome (fall / no fall).
    p = orc.Problem.make(150)
    lean = np.abs(ref["y_hist"][:, :, 0] - ref["y_hist"] @ p.C)
    err = np.abs(out["y_hist"][:, :, 0] - ref["y_hist"][:, :, 0])
    calm = lean.max(axis=1) < 0.12
    assert calm.mean() > 0.5
    assert err[calm].max() < TOL_REDUCED_M
    window = np.arange(lean.shape[1])[None, :] <= (tick[:, None] + 150)
    assert (err * window).max() < TOL_REDUCED_M
    fall_lean = 0.75 * np.tan(np.radians(30.0))
    lean_out = np.abs(out["y_hist"][:, :, 0] - out["y_hist"] @ p.C)
    decided = np.abs(lean.max(axis=1) - fall_lean) > 0.05 * fall_lean
    assert np.array_equal((lean_out.max(axis=1) > fall_lean)[decided], (lean.max(axis=1) > fall_lean)[decided])
    assert 0.2 < (lean.max(axis=1) > fall_lean).mean() < 0.8    # the sweep has both outcomes
    assert np.abs(out["x_hist"][:, :, 0] - ref["x_hist"][:, :, 0]).max() < TOL_REDUCED_M
    assert (out["x_hist"] == out["x_hist"][0]).all()            # the push acts on y only
    assert (out["y_hist"][:8] == out["y_hist"][0]).all()
    viol_bits = engine.FLAG_VIOL_LO | engine.FLAG_VIOL_HI | engine.FLAG_NOT_CONVERGED
    assert np.array_equal(out["flags"] & viol_bits, ref["flags"] & viol_bits)
    # a sub-batch in another order gives the same bits (sharding invariance)
    pick = np.array([4095, 17, 2048, 9])
    sub = plan150_f16.rollout(engine.MODE_STRICT, zero[pick], zero[pick], z_min, z_max, impulse=impulse[pick],
                              force_tick=tick[pick])
    assert np.array_equal(sub["y_hist"].cpu().numpy(), out["y_hist"][pick])
    assert np.array_equal(sub["iters"].cpu().numpy(), out["iters"][pick])


def test_reduced_ragged_and_limits(engine, plan150_f16):
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(5, seed=3)
    n_steps = np.array([70, 2, 64, 33, 70], np.int32)
    tick = (n_steps // 2).astype(np.int32)
    zero = np.zeros((5, 3))
    out = plan150_f16.rollout(engine.MODE_STRICT, zero, zero, z_min[:70], z_max[:70], n_steps=n_steps,
                              impulse=impulse, force_tick=tick)
    out = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
    for b, n in enumerate(n_steps):
        one = plan150_f16.rollout(engine.MODE_STRICT, zero[:1], zero[:1], z_min[:n], z_max[:n],
                                  impulse=impulse[b:b + 1], force_tick=tick[b:b + 1])
        assert np.array_equal(one["y_hist"][0].cpu().numpy(), out["y_hist"][b, :n]), b
        assert np.array_equal(one["flags"][0].cpu().numpy(), out["flags"][b, :n])
        assert not out["y_hist"][b, n:].any()
    with pytest.raises(engine.LipmError, match="F16X3"):
        engine.get_plan(638, 0.75, 9.81, 1.5 / 638, 1.0, 1e-6, precision=engine.PRECISION_F16X3)
    # the unconstrained mode of a reduced-precision plan is the same FP64 gain-row kernel
    a = plan150_f16.rollout(engine.MODE_UNCONSTRAINED, zero, zero, z_min[:70], z_max[:70])
    assert np.isfinite(a["y_hist"].cpu().numpy()).all()


# ----------------------------------------------------------------------------- parameter sweep (BASELINE config 5)
def test_param_sweep_matches_individual_walks(engine, default_config):
    """Each scenario of a (h, m, step_length, foot_spread, F_ext) sweep equals the same walk run alone
    through the mirrored reference API with that scenario's own MPCConfig, whatever the number of ranks
    it is dealt to; the corners of BASELINE config 5's box (h 0.6 / 0.9, m 30 / 60) against the oracle."""
    from dataclasses import replace

    import torch

    from mpc_bipedal import parallel, sweeps
    from mpc_bipedal.controllers import ZMPController
    from mpc_bipedal.generators import CoPGenerator

    cfg = replace(default_config, distance=0.9, standing_duration=0.4)
    table = sweeps.make_sweep(96, seed=5, h_values=(0.6, 0.9), step_lengths=(0.25, 0.35), foot_spreads=(0.09,),
                              mass_range=(30.0, 60.0))
    table.mass[[0, 40, 95]] = [30.0, 60.0, 30.0]                                  # pin the corners of the mass range
    whole = sweeps.run_shard(cfg, table, 0, 1)
    full = parallel.gather_by_index(whole["record"], whole["index"], len(table), 1)
    parts = [sweeps.run_shard(cfg, table, rank, 3) for rank in range(3)]          # three ranks' shares on one GPU
    dealt = torch.empty_like(full)
    for p in parts:
        dealt[p["index"]] = p["record"]
    assert torch.equal(dealt, full)                                               # the partition does not change a bit
    assert sum(p["solves"] for p in parts) == whole["solves"] and sum(p["launched"] for p in parts) == 96
    out = parallel.unpack_outcomes(full.cpu())
    assert (out["status"] == 0).all() and out["max_violation"].max().item() < 1e-9
    _, _, lengths, first, last = sweeps.device_bounds(cfg, table.variants)
    assert whole["solves"] == int((2 * (lengths[table.variant].astype(np.int64) - 1)).sum())
    for k in (0, 40, 95):
        step_length, foot_spread = table.variants[table.variant[k]]
        h = float(table.h_values[table.h_index[k]])
        one = replace(cfg, h=h, m=float(table.mass[k]), F_ext=float(table.F_ext[k]),
                      step_length=step_length, foot_spread=foot_spread, add_force=True, strict=True)
        z_max, z_min, _ = CoPGenerator(one).generate_cop_trajectory(save_footsteps=False)
        assert len(z_min) == lengths[table.variant[k]]
        tick = sweeps.push_ticks(table, np.array([k]), lengths, first, last)
        res = ZMPController(one).generate_com_trajectory_batched(np.zeros((1, 3)), np.zeros((1, 3)), z_max, z_min,
                                                                 F_ext=[one.F_ext], m=[one.m], force_tick=tick)
        assert int(res["iters"].sum()) == int(out["iters"][k])                    # same solves, iteration for iteration
        assert bool(res["fell"][0]) == bool(out["fell"][k])
        assert float(res["lean_max"][0]) == float(out["lean_max"][k])
        # ... and the walk itself against the oracle (its own Px / Pu for this CoM height): up to the push plus one second
        p = orc.Problem.make(one.horizon, one.h, one.g, one.Q, one.R)
        ticks = min(int(tick[0]) + 100, len(z_min) - 1)
        want = orc.rollout_wieber(p, np.zeros(3), np.zeros(3), z_max, z_min, strict=True, impulse=one.dt * one.F_ext / one.m,
                                  force_tick=int(tick[0]), solver="admm", n_ticks=ticks)
        for key in ("x_hist", "y_hist"):
            got = res[key][0][: ticks + 1]
            assert np.max(np.abs(got - want[key])) <= 1e-9 * max(1.0, np.abs(want[key]).max()), (k, key)


@pytest.mark.parametrize("N", [152, 318, 477, 637])
def test_reduced_long_horizons_against_the_oracle(engine, N):
    """The largest horizon of each reduced-precision kernel shape (N = 152: resident operands; 318 / 477 / 637: 2 / 3 / 4
    streamed slabs) against the ORACLE's FP64 ADMM (own Px / Pu / Hessian for that horizon), not just against the
    library's FP64 path: ZMP plan and next state within 1e-4 m, and the FP64 path within 1e-9 of the same oracle."""
    rng = np.random.default_rng(300 + N)
    p = orc.Problem.make(N)
    cases = 3
    x = rng.normal(size=(cases, 3)) * np.array([0.05, 0.1, 0.5])
    centre = rng.uniform(-0.05, 0.05, size=(cases, 1)) + np.cumsum(
        (rng.random((cases, N)) < 0.04) * rng.uniform(-0.1, 0.1, size=(cases, N)), axis=1)
    lo, hi = centre - 0.055, centre + 0.055
    rho = float(np.sqrt(p.Q * (p.Q + np.linalg.eigvalsh(p._cache()["Hz"] - p.Q * np.eye(N)).max())))   # fast for cond(H) ~ 100
    f16 = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, precision=engine.PRECISION_F16X3)
    fp64 = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6)
    got16 = f16.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    got64 = fp64.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    assert (got16["status"].cpu().numpy() == 0).all() and (got64["status"].cpu().numpy() == 0).all()
    active = 0
    for k in range(cases):
        nxt, z, w, it = orc.strict_axis_admm(p, x[k], hi[k], lo[k], rho=rho, max_iter=20000)
        assert it < 20000
        active += int(((z == lo[k]) | (z == hi[k])).sum())
        assert np.max(np.abs(got64["zmp_plan"][k].cpu().numpy() - z)) < 1e-9
        assert np.max(np.abs(got64["next_state"][k].cpu().numpy() - nxt)) < 1e-9 * max(1.0, np.abs(nxt).max()) * 10
        assert np.max(np.abs(got16["zmp_plan"][k].cpu().numpy() - z)) < TOL_REDUCED_M
        # position and velocity of the next state; its acceleration carries the first jerk (z0 - Px[0] x) / Pu[0,0] with
        # |dt / Pu[0,0]| = g / h = 13 per metre of ZMP error
        assert (np.abs(got16["next_state"][k].cpu().numpy() - nxt) < TOL_REDUCED_M * np.array([1.0, 1.0, 20.0])).all()
    assert active > 0


@pytest.mark.parametrize("N", [24, 100, 128, 152, 153, 200, 300, 318, 477, 600, 637])
def test_reduced_other_horizons_against_fp64_path(engine, N):
    """Every horizon the reduced-precision kernels serve (N <= 152: operands resident in shared memory, one 160-row
    tile; beyond: operands streamed, 1-4 slabs, BASELINE config 3's N = 300 / 600): solution within 1e-4 m of the FP64
    path of the same library, feasible after the FP64 clamp, far fewer iterations, bit-identical when repeated."""
    rng = np.random.default_rng(21 + N)
    fp64 = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, solver="admm")      # iteration counts are compared below
    f16 = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, precision=engine.PRECISION_F16X3)
    cases = 300                      # more than two 128-column units: partial last unit
    x = rng.normal(size=(cases, 3)) * np.array([0.05, 0.1, 0.5])
    centre = rng.uniform(-0.05, 0.05, size=(cases, 1)) + np.cumsum(
        (rng.random((cases, N)) < 0.04) * rng.uniform(-0.1, 0.1, size=(cases, N)), axis=1)
    lo, hi = centre - 0.055, centre + 0.055
    want = fp64.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    got = f16.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    assert (got["status"].cpu().numpy() == 0).all()
    z, zw = got["zmp_plan"].cpu().numpy(), want["zmp_plan"].cpu().numpy()
    assert np.max(np.abs(z - zw)) < TOL_REDUCED_M
    assert (z >= lo).all() and (z <= hi).all()
    assert np.max(np.abs(got["next_state"].cpu().numpy()[:, 0] - want["next_state"].cpu().numpy()[:, 0])) < TOL_REDUCED_M
    assert got["iters"].sum().item() < want["iters"].sum().item()
    again = f16.solve_tick(engine.MODE_STRICT, x, lo, hi, want_plan=True)
    assert np.array_equal(again["zmp_plan"].cpu().numpy(), z)


def test_reduced_streamed_walks(engine):
    """Whole walks on the streamed reduced-precision kernel (N = 200, 2 s preview: two slabs; the warm start is shifted
    through the per-CTA workspace, the axis state rides in contraction slots 200..202): within 1e-4 m of the FP64
    path, ZMP inside the box, ragged lengths, and each walk bit-identical when it is run alone."""
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(7, seed=9)
    n, N = 130, 200
    z_max, z_min = z_max[:n], z_min[:n]
    fp64 = engine.get_plan(N, 0.75, 9.81, 0.01, 1.0, 1e-6, solver="admm")      # iteration counts are compared below
    f16 = engine.get_plan(N, 0.75, 9.81, 0.01, 1.0, 1e-6, precision=engine.PRECISION_F16X3)
    n_steps = np.array([130, 130, 60, 130, 2, 101, 130], np.int32)
    tick = np.array([105, 110, 30, 120, 0, 50, 115], np.int32)
    zero = np.zeros((7, 3))
    kw = dict(n_steps=n_steps, impulse=impulse, force_tick=tick)
    want = {k: v.cpu().numpy() for k, v in fp64.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, **kw).items() if not k.startswith("_")}
    got = {k: v.cpu().numpy() for k, v in f16.rollout(engine.MODE_STRICT, zero, zero, z_min, z_max, **kw).items() if not k.startswith("_")}
    assert not want["status"].any() and not got["status"].any()
    for key in ("x_hist", "y_hist"):
        assert np.max(np.abs(got[key][:, :, 0] - want[key][:, :, 0])) < TOL_REDUCED_M, key
    assert got["max_violation"].max() < 1e-9
    assert got["iters"].sum() < want["iters"].sum()
    for b in (0, 2, 4, 5):
        nb = int(n_steps[b])
        one = f16.rollout(engine.MODE_STRICT, zero[:1], zero[:1], z_min[:nb], z_max[:nb], impulse=impulse[b:b + 1],
                          force_tick=tick[b:b + 1])
        assert np.array_equal(one["y_hist"][0].cpu().numpy(), got["y_hist"][b, :nb]), b
        assert np.array_equal(one["x_hist"][0].cpu().numpy(), got["x_hist"][b, :nb]), b
        assert np.array_equal(one["flags"][0].cpu().numpy(), got["flags"][b, :nb]), b
        assert not got["y_hist"][b, nb:].any()


# ----------------------------------------------------------------------------- device-side generators
def test_device_bound_generation_is_bit_identical(engine, default_config):
    """lipm_generate_bounds against the host generators (which are pinned against the reference itself):
    same samples, same state timeline, same bits -- including the irregular 24/25-sample single supports."""
    from dataclasses import replace

    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.generators import CoPGenerator
    from mpc_bipedal.herdt import state_codes

    rng = np.random.default_rng(4)
    for base in (default_config, MPCConfig(), replace(default_config, horizon=100, dt=1.5 / 100, ssp_duration=0.31, dsp_duration=0.05)):
        V = 40
        distance = np.concatenate([[base.distance, 2.0, 0.29, 0.3, 0.61], rng.uniform(0.2, 3.0, V - 5)])
        step = np.concatenate([[base.step_length, 0.33, 0.3, 0.3, 0.3], rng.uniform(0.15, 0.45, V - 5)])
        spread = np.concatenate([[base.foot_spread, 0.1, 0.1, 0.1, 0.1], rng.uniform(0.07, 0.13, V - 5)])
        out = engine.generate_bounds(base, distance, step, spread, n_max=900)
        n = out["n"].cpu().numpy()
        assert (n > 0).all()
        z_max, z_min, states = (out[k].cpu().numpy() for k in ("z_max", "z_min", "states"))
        for v in range(V):
            cfg = replace(base, distance=float(distance[v]), step_length=float(step[v]), foot_spread=float(spread[v]))
            want_max, want_min, want_states = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
            assert n[v] == len(want_max), (v, n[v], len(want_max))
            assert np.array_equal(z_max[v, : n[v]], want_max) and np.array_equal(z_min[v, : n[v]], want_min), v
            assert np.array_equal(states[v, : n[v]], state_codes(want_states)), v
            assert not z_max[v, n[v]:].any()
    short = engine.generate_bounds(default_config, [2.1], [0.3], [0.1], n_max=100)
    assert short["n"].item() == -420                    # does not fit: reports the length it needs
    # the generated table drives a sweep directly
    plan = engine.get_plan(150, 0.75, 9.81, 0.01, 1.0, 1e-6)
    table = engine.generate_bounds(default_config, [2.1, 1.2], [0.3, 0.25], [0.1, 0.09], n_max=450)
    idx = np.array([0, 1, 1, 0], np.int32)
    zero = np.zeros((4, 3))
    res = plan.rollout(engine.MODE_STRICT, zero, zero, table["z_min"], table["z_max"], n_steps=table["n"][idx].cpu().numpy(),
                       bounds_index=idx)
    z_max0, z_min0, _ = CoPGenerator(default_config).generate_cop_trajectory(save_footsteps=False)
    alone = plan.rollout(engine.MODE_STRICT, zero[:1], zero[:1], z_min0, z_max0)
    assert np.array_equal(res["y_hist"][0, :420].cpu().numpy(), alone["y_hist"][0].cpu().numpy())
    assert np.array_equal(res["y_hist"][3].cpu().numpy(), res["y_hist"][0].cpu().numpy())


# ----------------------------------------------------------------------------- edge cases
def test_edge_cases(engine, plan150, plan150_f16):
    import torch

    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(4)
    dev = plan150.device
    # empty batch: a no-op that succeeds
    empty = plan150.rollout(engine.MODE_STRICT, np.zeros((0, 3)), np.zeros((0, 3)), z_min[:30], z_max[:30])
    assert empty["x_hist"].shape == (0, 30, 3)
    # walks of 0, 1 and 2 samples inside a batch: nothing / nothing / one tick; the others are unaffected
    n_steps = np.array([0, 1, 2, 30], np.int32)
    zero = np.zeros((4, 3))
    for plan in (plan150, plan150_f16):
        out = plan.rollout(engine.MODE_STRICT, zero, zero, z_min[:30], z_max[:30], n_steps=n_steps)
        y = out["y_hist"].cpu().numpy()
        assert not y[0].any() and not y[1].any() and not y[2, 2:].any()
        alone = plan.rollout(engine.MODE_STRICT, zero[:1], zero[:1], z_min[:30], z_max[:30])
        assert np.array_equal(alone["y_hist"][0].cpu().numpy(), y[3])
        two = plan.rollout(engine.MODE_STRICT, zero[:1], zero[:1], z_min[:2], z_max[:2])
        assert np.array_equal(two["y_hist"][0].cpu().numpy(), y[2, :2])
    # degenerate box (z_min == z_max): the ZMP plan is the bound itself, both active bits are set
    x = np.array([[0.01, 0.02, 0.0]])
    line = np.full((1, 150), 0.02)
    for plan, tol in ((plan150, 1e-12), (plan150_f16, 1e-12)):
        res = plan.solve_tick(engine.MODE_STRICT, x, line, line, want_plan=True)
        assert res["status"].item() == 0
        assert np.max(np.abs(res["zmp_plan"].cpu().numpy() - 0.02)) <= tol
    walk = plan150.rollout(engine.MODE_STRICT, zero[:1], zero[:1], np.zeros((20, 2)), np.zeros((20, 2)))
    assert (walk["flags"][0, 1:].cpu().numpy() & 3 == 3).all()
    # the smallest horizons
    for N in (2, 3, 8):
        plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6)
        p = orc.Problem.make(N)
        rng = np.random.default_rng(N)
        xs = rng.normal(size=(9, 3)) * [0.02, 0.05, 0.1]
        lo = rng.uniform(-0.06, -0.02, size=(9, N)); hi = lo + 0.08
        res = plan.solve_tick(engine.MODE_STRICT, xs, lo, hi, want_plan=True)
        for k in range(9):
            _, want, _ = orc.strict_axis_exact(p, xs[k], hi[k], lo[k])
            assert np.max(np.abs(res["zmp_plan"][k].cpu().numpy() - want)) < 1e-10, (N, k)
        free = plan.solve_tick(engine.MODE_UNCONSTRAINED, xs, lo, hi, want_plan=True)
        for k in range(9):
            assert np.max(np.abs(free["next_state"][k].cpu().numpy() - orc.unconstrained_axis(p, xs[k], hi[k], lo[k]))) < 1e-9
    # a QP that cannot converge within max_iter stops its walk there, like the reference raises (:193-194)
    tight = engine.get_plan(150, 0.75, 9.81, 0.01, 1.0, 1e-6, max_iter=3, solver="admm")
    out = tight.rollout(engine.MODE_STRICT, zero[:2], zero[:2], z_min[:60], z_max[:60], impulse=[0.0, 0.2], force_tick=[5, 5])
    status = out["status"].cpu().numpy()
    flags = out["flags"].cpu().numpy()
    assert status.max() == 1                                       # at most one failed tick per axis: the walk ends there
    first_bad = np.argmax(flags[1, :, 1] & engine.FLAG_NOT_CONVERGED != 0)
    assert first_bad > 0 and not out["y_hist"][1, first_bad + 1:].cpu().numpy().any()
    from mpc_bipedal.controllers import ZMPController
    ctrl = ZMPController(cfg)
    ctrl.solver_options = {"max_iter": 3, "solver": "admm"}
    with pytest.raises(RuntimeError, match="QP solver did not find a solution"):
        ctrl.generate_com_trajectory(np.zeros((3, 1)), np.zeros((3, 1)), z_max, z_min)
    # the active-set solver reports a QP whose set is still changing after max_sweeps corrections the same way
    one_sweep = engine.LipmPlan(150, 0.75, 9.81, 0.01, 1.0, 1e-6)
    one_sweep.set_max_sweeps(1)
    out = one_sweep.rollout(engine.MODE_STRICT, zero[:2], zero[:2], z_min, z_max, impulse=[0.0, 0.2], force_tick=[150, 150])
    status, flags = out["status"].cpu().numpy(), out["flags"].cpu().numpy()
    assert status.max() == 1 and status[1, 1] == 1                 # the first step transition that enters the preview needs a correction
    first_bad = np.argmax(flags[1, :, 1] & engine.FLAG_NOT_CONVERGED != 0)
    assert first_bad > 0 and not out["y_hist"][1, first_bad + 1:].cpu().numpy().any()
    one_sweep.close()
    # ... and the batched entry point of the mirrored API then solves those scenarios on the ADMM kernel instead
    ctrl = ZMPController(cfg)
    shared = ctrl._plan()
    shared.set_max_sweeps(1)
    try:
        res = ctrl.generate_com_trajectory_batched(zero[:2], zero[:2], z_max, z_min, F_ext=[0.0, 400.0], force_tick=[150, 150])
    finally:
        shared.set_max_sweeps(0)
    assert (res["status"] == 0).all()
    ctrl.solver_options = {"solver": "admm"}
    want = ctrl.generate_com_trajectory_batched(zero[:2], zero[:2], z_max, z_min, F_ext=[0.0, 400.0], force_tick=[150, 150])
    assert np.array_equal(res["y_hist"], want["y_hist"]) and np.array_equal(res["flags"], want["flags"])
    # bad arguments are rejected with a message, never a crash
    with pytest.raises(ValueError, match="bounds batch"):
        plan150.rollout(engine.MODE_STRICT, zero, zero, np.zeros((3, 30, 2)), np.zeros((3, 30, 2)))
    with pytest.raises(ValueError):
        plan150.solve_tick(engine.MODE_STRICT, zero, np.zeros((4, 149)), np.zeros((4, 149)))
    with pytest.raises(engine.LipmError, match="mode"):
        plan150.rollout(7, zero, zero, z_min[:30], z_max[:30])


# ----------------------------------------------------------------------------- SpeedTrajectoryGenerator "wieber" mode
def test_speed_generator_wieber_mode_matches_reference(engine, golden, default_config):
    """`SpeedTrajectoryGenerator` in "wieber" mode is a second caller of the hot path (speed_generation.py:53-67): a whole
    push-free walk on the engine, velocities read off the histories.  Against the REFERENCE's own output (strict=False
    runs there without cvxpy; tests/golden/ref_speed_wieber.npz), single and batched."""
    from dataclasses import replace

    from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator
    from mpc_bipedal.herdt import state_codes

    g = golden("ref_speed_wieber")
    cases = {"default_json": replace(default_config, strict=False, speed_generation="wieber"),
             "short": replace(default_config, strict=False, speed_generation="wieber", distance=0.9, step_length=0.25,
                              standing_duration=0.4, horizon=100, dt=1.5 / 100)}
    for tag, cfg in cases.items():
        vx, vy, states = SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
        assert np.array_equal(state_codes(states), g[f"{tag}/states"])
        assert rel_err(vx, g[f"{tag}/vx"]) < REL and rel_err(vy, g[f"{tag}/vy"]) < REL, tag
    # batched over walk variants: variant 0 is the default walk, the others are each equal to their own single run
    cfg = cases["default_json"]
    gen = SpeedTrajectoryGenerator(cfg)
    dist, step, spread = np.array([cfg.distance, 0.9, 1.4]), np.array([cfg.step_length, 0.25, 0.35]), np.array([cfg.foot_spread, 0.09, 0.12])
    out = gen.generate_speed_and_state_batched(dist, step, spread)
    n0 = int(out["n"][0])
    assert n0 == len(g["default_json/vx"])
    assert rel_err(out["v"][0, :n0, 0], g["default_json/vx"]) < REL and rel_err(out["v"][0, :n0, 1], g["default_json/vy"]) < REL
    for v in (1, 2):
        one = replace(cfg, distance=float(dist[v]), step_length=float(step[v]), foot_spread=float(spread[v]))
        vx, vy, states = SpeedTrajectoryGenerator(one).generate_speed_and_state(save_footsteps=False)
        nv = int(out["n"][v])
        assert nv == len(vx) and np.array_equal(out["states"][v, :nv], state_codes(states))
        assert np.array_equal(out["v"][v, :nv, 0], vx) and np.array_equal(out["v"][v, :nv, 1], vy)      # same kernel, same bits
        assert not out["v"][v, nv:].any()
    classic = SpeedTrajectoryGenerator(replace(cfg, speed_generation="classic")).generate_speed_and_state_batched(dist, step, spread)
    vx, _, _ = SpeedTrajectoryGenerator(replace(cfg, speed_generation="classic")).generate_speed_and_state(save_footsteps=False)
    assert np.array_equal(classic["v"][0, :n0, 0], vx) and not classic["v"][..., 1].any()


# ----------------------------------------------------------------------------- CTA-pair kernel (tcgen05.mma cta_group::2)
def test_reduced_pair_kernel_equals_single_cta_kernel(engine):
    """The reduced-precision walks run on CTA pairs (cta_group::2, 64 columns per CTA, "2 x 2" tensor-memory layout; 20 warps
    per CTA for small batches, two 8-warp CTAs per SM for large ones).  Same arithmetic as the single-CTA tcgen05 kernel
    (LIPM_OPT_SINGLE_CTA_F16): every output is bit-identical -- ragged walks, pushes, partial units, both variants,
    single ticks with plans."""
    cfg, z_max, z_min, impulse, tick = push_sweep_inputs(9700, seed=13)         # 152 units: the two-CTAs-per-SM variant
    n = 60
    z_max, z_min = z_max[100:100 + n], z_min[100:100 + n]
    rng = np.random.default_rng(13)
    n_steps = rng.integers(2, n + 1, 9700).astype(np.int32)
    n_steps[:5] = [n, 0, 1, 2, n]
    tick = (n_steps // 2).astype(np.int32)
    x0 = rng.normal(size=(9700, 3)) * [0.01, 0.02, 0.05]
    plan = engine.get_plan(150, 0.75, 9.81, 0.01, 1.0, 1e-6, precision=engine.PRECISION_F16X3)
    for count in (9700, 300, 64, 5):                                             # 300 -> 5 units: the 20-warp variant
        results = []
        for single in (True, False):
            plan.set_single_cta_f16(single)
            out = plan.rollout(engine.MODE_STRICT, x0[:count], x0[:count], z_min, z_max, n_steps=n_steps[:count],
                               impulse=impulse[:count], force_tick=tick[:count], want_readout=True)
            results.append({k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")})
        for key in results[0]:
            assert np.array_equal(results[0][key], results[1][key]), (count, key)
    # single ticks (the predict_wieber_axis path), other horizons
    for N in (16, 100, 152):
        p = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, precision=engine.PRECISION_F16X3)
        x = rng.normal(size=(200, 3)) * [0.05, 0.1, 0.5]
        centre = rng.uniform(-0.05, 0.05, size=(200, 1)) + 0.0 * np.zeros((200, N))
        got = []
        for single in (True, False):
            p.set_single_cta_f16(single)
            res = p.solve_tick(engine.MODE_STRICT, x, centre - 0.03, centre + 0.03, want_plan=True)
            got.append({k: res[k].cpu().numpy() for k in ("next_state", "zmp_plan", "iters", "status")})
        for key in got[0]:
            assert np.array_equal(got[0][key], got[1][key]), (N, key)
        p.set_single_cta_f16(False)
    plan.set_single_cta_f16(False)
