"""The oracle against the reference's own outputs (committed golden vectors) and against itself.

CPU only.  ``ref_*`` fixtures were produced by importing the reference in the
build container (tests/golden/make_golden.py); ``oracle_*`` fixtures by the
oracle's exact solver and are used here for regression and for the
two-independent-solvers agreement that defines the strict oracle.
"""

import numpy as np
import pytest

from oracle import lipm_oracle as orc


def test_preview_matrices_are_markov_parameters():
    # Px[i] = C A^(i+1), Pu[i,j] = C A^(i-j) B  
    p = orc.Problem.make(40, h=0.8)
    Apow = np.eye(3)
    for i in range(p.N):
        markov = p.C @ Apow @ p.B
        assert np.allclose(np.diag(p.Pu, -i), markov, rtol=1e-12, atol=1e-18)
        Apow = Apow @ p.A
        assert np.allclose(p.Px[i], p.C @ Apow, rtol=1e-13)
    assert np.all(np.triu(p.Pu, 1) == 0)


def test_nonstrict_ticks_match_reference(golden):
    g = golden("ref_nonstrict_ticks")
    for N in (10, 50, 100, 150):
        for h in (0.75, 0.62):
            key = f"N{N}_h{h}"
            p = orc.Problem.make(N, h=h)
            for x, lo, hi, want in zip(g[f"{key}/x"], g[f"{key}/z_min"], g[f"{key}/z_max"], g[f"{key}/next"]):
                got = orc.unconstrained_axis(p, x, hi, lo)
                assert np.allclose(got, want, rtol=1e-9, atol=1e-12), key


def test_nonstrict_default_walk_matches_reference(golden):
    g = golden("ref_nonstrict_walk")
    p = orc.Problem.make(150)
    res = orc.rollout_wieber(p, np.zeros(3), np.zeros(3), g["z_max"], g["z_min"], strict=False,
                             impulse=float(g["impulse"]), force_tick=int(g["force_tick"]))
    com, y_hist = orc.reference_outputs(res)
    assert com.shape == g["com"].shape and y_hist.shape == g["y_hist"].shape
    scale = np.max(np.abs(g["y_hist"]))
    assert np.max(np.abs(com - g["com"])) <= 1e-9 * np.max(np.abs(g["com"]))
    assert np.max(np.abs(y_hist - g["y_hist"])) <= 1e-9 * scale
    # generate_state_trajectory_wieber: same loop, never pushes
    res = orc.rollout_wieber(p, np.zeros(3), np.zeros(3), g["z_max"], g["z_min"], strict=False)
    assert np.max(np.abs(res["x_hist"] - g["nopush_x_hist"][:, :, 0])) <= 1e-9 * np.max(np.abs(g["nopush_x_hist"]))
    assert np.max(np.abs(res["y_hist"] - g["nopush_y_hist"][:, :, 0])) <= 1e-9 * np.max(np.abs(g["nopush_y_hist"]))
    # the un-constrained controller lets the ZMP leave the sole by 0.038 m after the push
    assert 0.03 < res_max_violation(p, g) < 0.045


def res_max_violation(p, g):
    res = orc.rollout_wieber(p, np.zeros(3), np.zeros(3), g["z_max"], g["z_min"], strict=False,
                             impulse=float(g["impulse"]), force_tick=int(g["force_tick"]))
    return res["max_violation"]


@pytest.mark.parametrize("N", [16, 50, 150])
def test_strict_two_solvers_agree(golden, N):
    g = golden("oracle_strict_ticks")
    p = orc.Problem.make(N)
    M = orc.admm_matrix(p, 2.0)
    for k in range(0, len(g[f"N{N}/x"]), 3 if N == 150 else 1):
        x, lo, hi = g[f"N{N}/x"][k], g[f"N{N}/z_min"][k], g[f"N{N}/z_max"][k]
        nxt_a, z_a, _, iters = orc.strict_axis_admm(p, x, hi, lo, M=M)
        assert iters < 5000
        assert np.max(np.abs(z_a - g[f"N{N}/z"][k])) <= 1e-11
        assert np.allclose(nxt_a, g[f"N{N}/next"][k], rtol=0, atol=1e-10)
        # ADMM clips exactly, so its active set is readable by equality; ties aside it equals BVLS's
        act = np.where(z_a == lo, -1, 0) + np.where(z_a == hi, 1, 0)
        free_margin = np.minimum(g[f"N{N}/z"][k] - lo, hi - g[f"N{N}/z"][k])
        decided = (g[f"N{N}/active"][k] != 0) | (free_margin > 1e-9)
        assert np.array_equal(act[decided], g[f"N{N}/active"][k][decided])


def test_strict_exact_regression(golden):
    g = golden("oracle_strict_ticks")
    p = orc.Problem.make(50)
    for k in range(6):
        nxt, z, act = orc.strict_axis_exact(p, g["N50/x"][k], g["N50/z_max"][k], g["N50/z_min"][k])
        assert np.allclose(z, g["N50/z"][k], rtol=0, atol=1e-13)
        assert np.array_equal(act, g["N50/active"][k])


def test_strict_equals_unconstrained_when_inactive():
    # property: wide boxes -> strict optimum is the closed form
    p = orc.Problem.make(60)
    rng = np.random.default_rng(3)
    x = rng.normal(size=3) * np.array([0.02, 0.05, 0.1])
    lo, hi = np.full(60, -5.0), np.full(60, 5.0)
    want = orc.unconstrained_axis(p, x, hi, lo)
    got, z, act = orc.strict_axis_exact(p, x, hi, lo)
    assert not act.any()
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    got2, *_ = orc.strict_axis_admm(p, x, hi, lo)
    assert np.allclose(got2, want, rtol=1e-9, atol=1e-12)


def test_strict_walk_properties(golden):
    g = golden("oracle_strict_walk")
    # ZMP never leaves the box in strict mode and the final y state is the surveyed one
    assert g["margin"].min() > -1e-12
    assert not (g["flags"] & (orc.FLAG_VIOL_LO | orc.FLAG_VIOL_HI)).any()
    assert np.allclose(g["y_hist"][-1], [-0.0142, 0.0317, -0.0638], atol=1e-4)
    # the FIRST preview row (the one that is applied) sits on a bound on ~100 y ticks and few x ticks
    active = (g["flags"] & 3) != 0
    assert 80 < active[:, 1].sum() < 140 and active[:, 0].sum() < 30


def test_strict_walk_admm_rollout_matches_exact_prefix(golden):
    g = golden("oracle_strict_walk")
    p = orc.Problem.make(150)
    ticks = 40
    start = int(g["force_tick"]) - 5          # straddle the push: the hardest ticks
    x0, y0 = g["x_hist"][start], g["y_hist"][start]
    res = orc.rollout_wieber(p, x0, y0, g["z_max"][start:], g["z_min"][start:], strict=True,
                             impulse=float(g["impulse"]), force_tick=5, solver="admm", n_ticks=ticks)
    assert np.max(np.abs(res["x_hist"] - g["x_hist"][start:start + ticks + 1])) < 1e-10
    assert np.max(np.abs(res["y_hist"] - g["y_hist"][start:start + ticks + 1])) < 1e-10


# ----------------------------------------------------------------------------- C restatement
def test_c_oracle_matrices_and_walks(golden):
    from oracle.c_oracle import COracle

    c = COracle(150)
    p = orc.Problem.make(150)
    m = c.export()
    assert np.array_equal(m["Px"], p.Px) and np.array_equal(m["Pu"], p.Pu)     # bit-identical to the reference loop
    assert np.max(np.abs(m["M"] - orc.admm_matrix(p, 2.0))) < 1e-12
    K = (p._cache()["Hu_inv"] @ p.Pu.T)[0]
    assert np.max(np.abs(m["gain"][:150] - K)) / np.max(np.abs(K)) < 1e-9

    # non-strict walk against the reference's own output
    g = golden("ref_nonstrict_walk")
    zero = np.zeros((1, 3))
    res = c.rollout(False, zero, zero, g["z_min"], g["z_max"], impulse=[float(g["impulse"])],
                    force_tick=[int(g["force_tick"])])
    assert np.max(np.abs(res["y_hist"][0] - g["y_hist"][:, :, 0])) <= 1e-9 * np.max(np.abs(g["y_hist"]))
    assert np.max(np.abs(res["x_hist"][0, :, 0] - g["com"][:, 0])) <= 1e-9 * np.max(np.abs(g["com"]))

    # strict walk against the exact (BVLS) oracle, flags included
    g = golden("oracle_strict_walk")
    res = c.rollout(True, zero, zero, g["z_min"], g["z_max"], impulse=[float(g["impulse"])],
                    force_tick=[int(g["force_tick"])])
    assert (res["status"] == 0).all()
    assert np.max(np.abs(res["x_hist"][0] - g["x_hist"])) <= 1e-9 * np.max(np.abs(g["x_hist"]))
    assert np.max(np.abs(res["y_hist"][0] - g["y_hist"])) <= 1e-9 * np.max(np.abs(g["y_hist"]))
    zmp = np.stack([res["x_hist"][0] @ p.C, res["y_hist"][0] @ p.C], 1)
    margin = np.minimum(zmp - g["z_min"], g["z_max"] - zmp)
    diff = res["flags"][0] != g["flags"]
    assert not (diff & (np.abs(margin) > 1e-9)).any()          # flags bit-exact away from ties
    assert diff.sum() <= 8
    assert 10 < res["iters"].sum() / 838 < 60


def test_c_oracle_active_set_solver(golden):
    """The third strict solver of the oracle -- primal-dual active set with exact O(N) Riccati sweeps over the 3-state
    model (lipm_oracle.c:riccati_pdas_solve, the CPU restatement of csrc/lipm_rollout_riccati.cu) -- against the exact
    BVLS oracle and against the C ADMM: it shares no arithmetic with either (no condensed N x N matrix at all)."""
    from oracle.c_oracle import COracle

    c = COracle(150)
    p = orc.Problem.make(150)
    g = golden("oracle_strict_walk")
    zero = np.zeros((1, 3))
    res = c.rollout("riccati", zero, zero, g["z_min"], g["z_max"], impulse=[float(g["impulse"])], force_tick=[int(g["force_tick"])])
    assert (res["status"] == 0).all()
    assert np.max(np.abs(res["x_hist"][0] - g["x_hist"])) <= 1e-9 * np.max(np.abs(g["x_hist"]))
    assert np.max(np.abs(res["y_hist"][0] - g["y_hist"])) <= 1e-9 * np.max(np.abs(g["y_hist"]))
    zmp = np.stack([res["x_hist"][0] @ p.C, res["y_hist"][0] @ p.C], 1)
    margin = np.minimum(zmp - g["z_min"], g["z_max"] - zmp)
    assert margin.min() > -1e-12
    diff = res["flags"][0] != g["flags"]
    assert not (diff & (np.abs(margin) > 1e-9)).any() and diff.sum() <= 8
    assert 838 <= res["iters"].sum() < 1.5 * 838           # sweep pairs: the predicted set is right on most ticks
    # pushed scenarios (one of them falls), ragged lengths: same walks as the ADMM solver, same flags
    n = len(g["z_min"])
    B = 6
    imp = np.array([0.0, 0.05, 0.1, 0.16, 0.2, 0.25])
    tick = np.array([100, 150, 210, 120, 319, 101], np.int32)
    n_steps = np.array([n, n, 300, n, n, 250], np.int32)
    zero = np.zeros((B, 3))
    a = c.rollout(True, zero, zero, g["z_min"], g["z_max"], n_steps=n_steps, impulse=imp, force_tick=tick)
    r = c.rollout("riccati", zero, zero, g["z_min"], g["z_max"], n_steps=n_steps, impulse=imp, force_tick=tick)
    assert not a["status"].any() and not r["status"].any()
    for key in ("x_hist", "y_hist"):
        scale = np.maximum(1.0, np.abs(a[key]).max(axis=(1, 2)))
        assert (np.abs(a[key] - r[key]).max(axis=(1, 2)) <= 1e-9 * scale).all(), key
    assert np.abs(a["y_hist"][:, :, 0]).max() > 1.0            # at least one scenario diverges (all rows pinned)
    assert (a["flags"] != r["flags"]).sum() <= 4
    assert r["max_violation"].max() < 1e-12


def test_c_oracle_ragged_and_batched():
    from oracle.c_oracle import COracle

    c = COracle(40)
    rng = np.random.default_rng(5)
    n = 55
    centre = np.cumsum((rng.random((n, 2)) < 0.1) * rng.uniform(-0.05, 0.05, (n, 2)), 0)
    z_min, z_max = centre - 0.04, centre + 0.04
    B = 6
    x0 = rng.normal(size=(B, 3)) * [0.01, 0.02, 0.05]
    y0 = rng.normal(size=(B, 3)) * [0.01, 0.02, 0.05]
    n_steps = np.array([55, 2, 30, 55, 17, 41], np.int32)
    imp = rng.uniform(0, 0.05, B)
    tick = (n_steps // 2).astype(np.int32)
    res = c.rollout(True, x0, y0, z_min, z_max, n_steps=n_steps, impulse=imp, force_tick=tick, threads=3)
    p = orc.Problem.make(40)
    for b in (0, 2, 4):
        nb = n_steps[b]
        want = orc.rollout_wieber(p, x0[b], y0[b], z_max[:nb], z_min[:nb], strict=True, impulse=imp[b],
                                  force_tick=int(tick[b]), solver="exact")
        assert np.max(np.abs(res["y_hist"][b, :nb] - want["y_hist"])) < 1e-9
        assert np.max(np.abs(res["x_hist"][b, :nb] - want["x_hist"])) < 1e-9
        assert not res["y_hist"][b, nb:].any()


# ----------------------------------------------------------------------------- the oracle IS the optimum of the reference's QP
KKT_TOL = 1e-12     # multipliers relative to max |multiplier|; primal feasibility absolute in metres


@pytest.mark.parametrize("N", [16, 50, 150])
def test_strict_golden_ticks_are_kkt_points_of_the_reference_qp(golden, N):
    """Every strict golden tick, mapped back to the jerk coordinates and the literal matrices of the reference's QP
    (`zmp_controller.py:174-189`: H = R I, A_ineq = [Pu; -Pu]), satisfies its KKT conditions: the golden vectors are
    the unique optimum of the problem the reference poses, not just the point two z-coordinate solvers agree on."""
    g = golden("oracle_strict_ticks")
    p = orc.Problem.make(N)
    active = 0
    for x, lo, hi, z, nxt in zip(g[f"N{N}/x"], g[f"N{N}/z_min"], g[f"N{N}/z_max"], g[f"N{N}/z"], g[f"N{N}/next"]):
        cert = orc.kkt_certificate_jerk(p, x, hi, lo, z)
        assert cert["primal"] <= 1e-12 and cert["free_multiplier"] <= KKT_TOL and cert["wrong_sign"] <= KKT_TOL, cert
        # the applied jerk is the first component of J = Pu^-1 (z - Px x): A x + B J0 is the golden next state
        assert np.allclose(p.A @ x + p.B * cert["jerk0"], nxt, rtol=1e-9, atol=1e-12)
        active += cert["active_rows"]
    assert active > 0                                      # the set exercises active constraints


def test_strict_golden_walk_ticks_are_kkt_points(golden):
    """Sampled ticks of the golden strict walk (before, at and after the push, where up to 43 rows are active):
    re-solving from the recorded state gives a certified KKT point whose first jerk reproduces the next recorded state."""
    g = golden("oracle_strict_walk")
    p = orc.Problem.make(150)
    N, n = 150, len(g["z_min"])
    hi = np.vstack([g["z_max"], np.tile(g["z_max"][-1:], (N, 1))])
    lo = np.vstack([g["z_min"], np.tile(g["z_min"][-1:], (N, 1))])
    force_tick, impulse = int(g["force_tick"]), float(g["impulse"])
    most_active = 0
    for i in (0, 99, 150, 209, 210, 211, 215, 230, 260, 300, 417):
        for ax, hist in ((0, g["x_hist"]), (1, g["y_hist"])):
            hw, lw = hi[i + 1:i + 1 + N, ax], lo[i + 1:i + 1 + N, ax]
            nxt, z, _ = orc.strict_axis_exact(p, hist[i], hw, lw)
            cert = orc.kkt_certificate_jerk(p, hist[i], hw, lw, z)
            assert cert["primal"] <= 1e-12 and cert["free_multiplier"] <= KKT_TOL and cert["wrong_sign"] <= KKT_TOL, (i, ax, cert)
            most_active = max(most_active, cert["active_rows"])
            want = hist[i + 1].copy()
            if ax == 1 and i == force_tick:
                want[1] += impulse                          # the recorded sample already carries the push
            assert np.max(np.abs(nxt - want)) <= 1e-9 * max(1.0, np.abs(want).max()), (i, ax)
    assert most_active >= 30


# ----------------------------------------------------------------------------- what the reference's solver would return
def test_osqp_restatement_known_answers():
    """The only solver output recorded anywhere in the reference: `src/mpc_bipedal/test.ipynb` cell 3 projects the point
    (0.6, -0.25) on the hull of the left footstep polytope with cvxpy -> OSQP and prints `Status: optimal / Optimal u:
    [ 0.6 -0.25]`.  Plus the solver's two other verdicts on toy problems."""
    from scipy.spatial import ConvexHull

    from oracle import osqp_restatement as osq

    pts = np.array([(-0.1, -0.3), (-0.1, -0.4), (0.0, -0.4), (0.0, -0.2), (0.1, -0.17), (0.2, -0.13), (0.3, -0.1), (0.7, -0.05),
                    (0.8, -0.05), (0.8, -0.3), (0.4, -0.35)])
    eq = ConvexHull(pts).equations
    A_h, b_h = eq[:, :2], -eq[:, 2]
    target = np.array([0.6, -0.25])
    # cvxpy: sum_squares(u - target) -> auxiliary t = u - target, minimise t't: variables [u; t]
    P = np.zeros((4, 4)); P[2:, 2:] = 2 * np.eye(2)
    A = np.zeros((2 + len(b_h), 4)); A[:2, :2], A[:2, 2:], A[2:, :2] = np.eye(2), -np.eye(2), A_h
    res = osq.solve(P, np.zeros(4), A, np.concatenate([target, np.full(len(b_h), -osq.INF)]), np.concatenate([target, b_h]))
    assert res["status"] == "solved"
    assert np.array_equal(np.round(res["x"][:2], 4), target)          # what the notebook printed
    # a target outside the hull lands on its boundary (projection), to the solver's tolerance
    out = np.array([0.9, 0.1])
    res = osq.solve(P, np.zeros(4), A, np.concatenate([out, np.full(len(b_h), -osq.INF)]), np.concatenate([out, b_h]))
    assert res["status"] == "solved" and abs(np.max(A_h @ res["x"][:2] - b_h)) < 1e-4
    # contradictory rows -> primal infeasible (cvxpy: u.value is None -> the reference's fallback / RuntimeError)
    res = osq.solve(np.eye(1), np.zeros(1), np.array([[1.0], [1.0]]), np.array([1.0, -osq.INF]), np.array([osq.INF, 0.0]))
    assert res["status"] == "primal_infeasible"
    # unbounded below along a free direction -> dual infeasible
    res = osq.solve(np.zeros((1, 1)), np.array([1.0]), np.array([[1.0]]), np.array([-osq.INF]), np.array([0.0]))
    assert res["status"] == "dual_infeasible"


def test_reference_solver_output_sits_within_its_tolerance_of_the_oracle(golden):
    """OSQP as cvxpy calls it (non-condensed form, eps_abs = eps_rel = 1e-5, restated) on golden strict ticks: it
    reports "solved", and what the reference would have returned differs from the exact optimum by about the solver's
    tolerance -- which is why parity is stated against the optimum, not against the reference's own ~1e-5 output."""
    from oracle import osqp_restatement as osq

    g = golden("oracle_strict_ticks")
    N = 150
    p = orc.Problem.make(N)
    worst_z = worst_state = 0.0
    for k in (0, 3, 7, 12):
        x, lo, hi, z = g[f"N{N}/x"][k], g[f"N{N}/z_min"][k], g[f"N{N}/z_max"][k], g[f"N{N}/z"][k]
        for interval in (25, 100):
            nxt, info = osq.wieber_strict_axis(p, x, hi, lo, osq.Settings(adaptive_rho_interval=interval))
            assert info["status"] == "solved" and info["iters"] <= 500
            plan = p.Px @ x + p.Pu @ info["x"][:N]
            worst_z = max(worst_z, float(np.abs(plan - z).max()))
            worst_state = max(worst_state, float(np.abs(nxt - g[f"N{N}/next"][k]).max()))
            assert max((plan - hi).max(), (lo - plan).max()) < 1e-4           # feasible to the solver's tolerance only
    assert worst_z < 1e-3 and worst_state < 1e-3
    assert worst_z > 1e-9                        # ... and NOT to 1e-9: the reference's own output is not the optimum
