"""Herdt free-footstep MPC: host layout on CPU, GPU joint-QP ticks and walks against the certified oracle.

Parity status: UNPINNED by the reference (cvxpy/OSQP not installed, no golden vectors); the oracle's
solutions carry a KKT certificate of optimality for the reference's own QP matrices.  Timeline, nb_steps
and polytope facets come from the reference's outputs recorded in tests/golden/ref_herdt_helpers.npz.
"""

import os

import numpy as np
import pytest

from oracle import herdt_oracle as ho

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def herdt_case(golden):
    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.models.lipm_model import discrete_model

    g = golden("ref_herdt_helpers")
    cfg = MPCConfig()                       # class defaults, as scripts/run_herdt_method.py:53
    cfg.method = "herdt"
    A, B, _ = discrete_model(cfg.dt, cfg.h, cfg.g)
    states = g["class_defaults/states"]
    v_ref = np.stack([g["class_defaults/vx"], g["class_defaults/vy"]], 1)
    polys = {s: (g[f"polytope/{s}/A"], g[f"polytope/{s}/b"]) for s in ("left", "right")}
    return dict(cfg=cfg, A=A, B=B, states=states, v_ref=v_ref, nb=g["class_defaults/nb_steps"], polys=polys)


# ----------------------------------------------------------------------------- CPU
def test_timeline_layout_matches_reference_selection_matrices(herdt_case):
    from mpc_bipedal import herdt

    N = herdt_case["cfg"].horizon
    lay = herdt.timeline_layout(herdt_case["states"], N)
    assert lay.phase_of_row.shape == (301, N) and lay.m.max() == 7
    assert lay.foot_side[0] == 0 and lay.commit.sum() == 8            # eight steps are committed
    for i in (0, 49, 50, 60, 100, 200, 251, 280, 300):
        runs = ho.support_phases(int(lay.current[i]), lay.codes[i + 1: i + 1 + N])
        assert sum(runs) == N + 1                                     # the current tick is counted
        U, U_c = ho.selection_matrices(N, runs)
        assert lay.m[i] == U.shape[1]
        assert np.array_equal(U_c[:, 0] == 1, lay.phase_of_row[i] == -1)
        for k in range(U.shape[1]):
            assert np.array_equal(U[:, k] == 1, lay.phase_of_row[i] == k)
        window = lay.codes[i + 1: i + 1 + N]
        assert np.array_equal(lay.row_kind[i] == 1, window != ho.STANDING)
    # side toggles exactly when a SINGLE_SUPPORT phase ends
    leaving = [(lay.current[i] == ho.SINGLE_SUPPORT) and (lay.codes[i + 1] != lay.current[i]) for i in range(300)]
    assert np.array_equal(lay.commit[:300], np.array(leaving))
    verts = herdt.polygon_vertices(herdt_case["cfg"].left_foot_polytope)
    A_poly, b_poly = herdt_case["polys"]["left"]
    assert len(verts) == len(b_poly) == 9 and np.max(verts @ A_poly.T - b_poly) < 1e-12
    e1, e2 = np.roll(verts, -1, 0) - verts, np.roll(verts, -2, 0) - np.roll(verts, -1, 0)
    assert (e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0] > 0).all()                                          # counter-clockwise, strictly convex


def test_oracle_qp_is_certified(herdt_case):
    c = herdt_case
    N = c["cfg"].horizon
    ext = np.concatenate([c["states"], np.repeat(c["states"][-1:], N)])
    v_ext = np.vstack([c["v_ref"], np.repeat(c["v_ref"][-1:], N, 0)])
    qp = ho.build_qp(c["cfg"], [0.01, 0.2, 0.0], [0.03, 0.0, -0.1], v_ext[61:61 + N], 0.0, 0.1, int(ext[60]),
                     ext[61:61 + N], "left", c["polys"]["left"])
    assert qp["m"] == 6 and qp["A"].shape == (2 * 2 * N + 9, 2 * (N + 6))       # no STANDING rows in this window
    u, lam, info = ho.solve_qp(qp)
    assert info["kkt"]["ok"], info
    assert (qp["A"] @ u <= qp["b"] + 1e-9).all()


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_herdt_ticks_match_oracle(herdt_case):
    import torch

    from mpc_bipedal import engine, herdt

    c = herdt_case
    cfg, N = c["cfg"], c["cfg"].horizon
    lay = herdt.timeline_layout(c["states"], N)
    plan = engine.HerdtPlan(N, cfg.h, cfg.g, cfg.dt, cfg.alpha, cfg.beta, cfg.gamma, cfg.foot_length, cfg.foot_width,
                            cfg.foot_spread)
    # W = alpha T'T + beta V'V against the reference's prediction matrices
    Pzx, Pzu, Pvs, Pvu = mats = ho.prediction_matrices(N, cfg.dt, cfg.h, cfg.g)
    T = np.linalg.inv(Pzu)
    W = cfg.alpha * T.T @ T + cfg.beta * (Pvu @ T).T @ (Pvu @ T)
    assert np.max(np.abs(plan.export_w() - W)) / np.max(np.abs(W)) < 1e-9
    ticks = [0, 60, 100, 251, 260]
    timeline = plan.timeline(lay.phase_of_row[ticks], lay.row_kind[ticks], lay.m[ticks])
    polygons = (herdt.polygon_vertices(cfg.left_foot_polytope), herdt.polygon_vertices(cfg.right_foot_polytope))
    v_ext = np.vstack([c["v_ref"], np.repeat(c["v_ref"][-1:], N, 0)])
    rng = np.random.default_rng(0)
    dev = plan.device
    for slot, i in enumerate(ticks):
        side = int(lay.foot_side[i])
        fc = np.array([[0.3 * (i // 50), 0.1 if side == 0 else -0.1]] * 3)
        x = rng.normal(size=(3, 3)) * [0.02, 0.05, 0.1] + [fc[0, 0], 0.2 if 50 < i < 250 else 0.0, 0.0]
        y = rng.normal(size=(3, 3)) * [0.01, 0.03, 0.1] + [fc[0, 1] * 0.3, 0.0, 0.0]
        # the tick API reads vref samples tick+1+j: slot s of this 5-tick timeline -> give it its own window
        vref = torch.as_tensor(np.vstack([np.zeros((slot + 1, 2)), v_ext[i + 1: i + 1 + N]])).to(dev)
        out = timeline.tick(slot, torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev), torch.as_tensor(fc).to(dev),
                            vref, side, polygons[side] if lay.m[i] > 0 else None)
        assert (out["status"].cpu().numpy() == 0).all()
        for b in range(3):
            xn, yn, fx1, fy1, u, lam, info = ho.predict_joint(
                cfg, c["A"], c["B"], x[b], y[b], v_ext[i + 1: i + 1 + N], fc[b, 0], fc[b, 1], int(lay.current[i]),
                lay.codes[i + 1: i + 1 + N], "left" if side == 0 else "right", c["polys"]["left" if side == 0 else "right"], mats)
            assert info["kkt"]["ok"]
            # next states (what the rollout consumes): position / velocity to 1e-9; the acceleration carries
            # the first jerk = (z0 - Pzx[0] x) / Pzu[0,0] with Pzu[0,0] = -7.6e-4, i.e. dt/7.6e-4 = 13 x the
            # error of the ZMP sample (ADMM stop: 1e-12 relative to |q|_inf, which is O(100) in these coordinates)
            tol = np.array([1e-9, 1e-9, 2e-7])
            assert (np.abs(out["next_x"][b].cpu().numpy() - xn) < tol).all(), (i, b)
            assert (np.abs(out["next_y"][b].cpu().numpy() - yn) < tol).all(), (i, b)
            if lay.m[i] > 0:
                assert abs(out["f1"][b, 0].item() - fx1) < 1e-7 and abs(out["f1"][b, 1].item() - fy1) < 1e-7, (i, b)
            else:
                assert torch.isnan(out["f1"][b]).all()
    plan.close()


@pytest.mark.gpu
def test_herdt_walk_matches_oracle(herdt_case, golden):
    from mpc_bipedal.controllers import ZMPController

    c = herdt_case
    g = golden("oracle_herdt_walk")
    cfg = c["cfg"]
    cfg.add_force = True
    ctrl = ZMPController(cfg)
    from mpc_bipedal.generators.cop_generator import State
    state_ref = np.array([State.STANDING, State.DOUBLE_SUPPORT, State.SINGLE_SUPPORT], dtype=object)[c["states"]]
    com, y_hist, foot_hist = ctrl.generate_com_trajectory(np.zeros((3, 1)), np.zeros((3, 1)), v_ref=c["v_ref"],
                                                          state_ref=state_ref)
    assert com.shape == (302, 2) and y_hist.shape == (302, 3, 1) and foot_hist.shape == (302, 2)
    # the oracle certifies the walk up to the final stance (tests/golden/make_golden.py explains why the
    # reference's formulation stops having a converging QP there); that prefix is the golden vector
    k = int(g["certified_ticks"]) + 1
    assert k > 250
    assert np.max(np.abs(foot_hist[:k] - g["foot_hist"])) < 1e-6
    assert np.max(np.abs(y_hist[:k, :2, 0] - g["y_hist"][:, :2])) < 1e-7          # position, velocity
    assert np.max(np.abs(y_hist[:k, 2, 0] - g["y_hist"][:, 2])) < 1e-5            # acceleration carries the jerk
    assert np.max(np.abs(com[:k, 0] - g["x_hist"][:, 0])) < 1e-7
    # batched: scenarios sharing the timeline, each equal to the same walk run alone
    B = 24
    rng = np.random.default_rng(2)
    x0 = rng.normal(size=(B, 3)) * [0.01, 0.02, 0.05]
    y0 = rng.normal(size=(B, 3)) * [0.01, 0.02, 0.05]
    F = rng.uniform(0, 300, B)
    v = c["v_ref"][None] * rng.uniform(0.8, 1.2, (B, 1, 1))
    res = ctrl.generate_com_trajectory_herdt_batched(x0, y0, v, state_ref, F_ext=F)
    assert np.isfinite(res["com"]).all()
    for b in (0, 23):
        one = ctrl.generate_com_trajectory_herdt_batched(x0[b:b + 1], y0[b:b + 1], v[b], state_ref, F_ext=F[b:b + 1])
        assert np.array_equal(one["y_hist"][0], res["y_hist"][b]) and np.array_equal(one["foot_hist"][0], res["foot_hist"][b])
    # every committed footstep lies in the reachable polygon of the previous one (reference check :814-823)
    lay_commit = np.where(np.any(np.diff(res["foot_hist"][0], axis=0) != 0, axis=1))[0]
    assert len(lay_commit) == 8
    # the device-side walk (herdt_rollout: one call) is the host-driven tick loop (herdt_tick per tick), bit for bit,
    # and walks of different timelines enqueued on different streams do not disturb each other
    import torch
    batch = ctrl._herdt_cache[1]
    spread = rng.uniform(0.08, 0.12, B)
    n = len(state_ref)
    whole = batch.rollout(x0, y0, v, F_ext=F, foot_spread=spread)
    ticked = batch.rollout(x0, y0, v, F_ext=F, foot_spread=spread, n_ticks=n - 1)
    for key in ("x_hist", "y_hist", "foot_hist", "status", "iters"):
        assert torch.equal(whole[key], ticked[key]), key
    short_ref = state_ref[:200]
    other = ZMPController(cfg)
    alone = other.generate_com_trajectory_herdt_batched(x0, y0, v[:, :200], short_ref, F_ext=F)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        a = ctrl.generate_com_trajectory_herdt_batched(x0, y0, v, state_ref, F_ext=F, to_numpy=False)
    with torch.cuda.stream(s2):
        b2 = other.generate_com_trajectory_herdt_batched(x0, y0, v[:, :200], short_ref, F_ext=F, to_numpy=False)
    torch.cuda.synchronize()
    assert np.array_equal(a["y_hist"].cpu().numpy(), res["y_hist"])
    assert np.array_equal(b2["y_hist"].cpu().numpy(), alone["y_hist"]) and np.array_equal(b2["foot_hist"].cpu().numpy(), alone["foot_hist"])


@pytest.mark.gpu
def test_predict_herdt_joint_reference_signature(herdt_case):
    """The reference's 13-argument call (zmp_controller.py:533-547) on the mirrored controller."""
    from mpc_bipedal.controllers import ZMPController
    from mpc_bipedal.generators.cop_generator import State

    c = herdt_case
    cfg, N = c["cfg"], c["cfg"].horizon
    ctrl = ZMPController(cfg)
    names = np.array([State.STANDING, State.DOUBLE_SUPPORT, State.SINGLE_SUPPORT], dtype=object)
    ext = np.concatenate([c["states"], np.repeat(c["states"][-1:], N)])
    v_ext = np.vstack([c["v_ref"], np.repeat(c["v_ref"][-1:], N, 0)])
    i = 60
    x, y = np.array([[0.01], [0.2], [0.0]]), np.array([[0.03], [0.0], [-0.1]])
    out = ctrl.predict_herdt_joint(x, y, v_ext[i + 1: i + 1 + N], np.array([[0.0]]), np.array([[0.1]]), names[ext[i]],
                                   names[ext[i + 1: i + 1 + N]], N, tuple(c["nb"][i]), np.array([[0.0]]),
                                   np.array([[0.1]]), "left", i)
    xn, yn, fx1, fy1, u, lam, info = ho.predict_joint(cfg, c["A"], c["B"], x, y, v_ext[i + 1: i + 1 + N], 0.0, 0.1,
                                                      int(ext[i]), ext[i + 1: i + 1 + N], "left", c["polys"]["left"])
    assert info["kkt"]["ok"]
    assert out[0].shape == (3, 1) and out[1].shape == (3, 1)
    assert (np.abs(out[0][:, 0] - xn) < [1e-9, 1e-9, 2e-7]).all() and (np.abs(out[1][:, 0] - yn) < [1e-9, 1e-9, 2e-7]).all()
    assert abs(out[2] - fx1) < 1e-7 and abs(out[3] - fy1) < 1e-7
    with pytest.raises(ValueError):
        ctrl.generate_com_trajectory(x, y)                  # herdt needs v_ref and state_ref (reference :53-54)


@pytest.mark.gpu
def test_herdt_randomized_plans_match_oracle(golden):
    """BASELINE configs[3]: three plans of the randomized set (seed 2) with three DIFFERENT timelines (452 / 377 / 302
    samples), own step length / foot spread / distance / velocity scale / push, against the KKT-certified oracle walks
    (tests/golden/oracle_herdt_plans.npz: timeline, reference speeds and nb_steps from the REFERENCE's own generators).
    Every tick up to the final stance carries a certificate; compared over that prefix."""
    from dataclasses import replace

    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.controllers import ZMPController
    from mpc_bipedal.generators.cop_generator import State
    from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator
    from mpc_bipedal.herdt import randomized_plans, state_codes

    g = golden("oracle_herdt_plans")
    names = np.array([State.STANDING, State.DOUBLE_SUPPORT, State.SINGLE_SUPPORT], dtype=object)
    lengths = set()
    plans = randomized_plans(MPCConfig(), 8192, seed=2)
    for k, b in enumerate(g["plan_ids"]):
        step_length, spread, distance, speed, F_ext = g[f"p{k}/params"]
        # the draw is the one the drivers and the bench use
        assert plans["step_length"][b] == step_length and plans["foot_spread"][b] == spread and plans["speed"][b] == speed
        cfg = replace(MPCConfig(), step_length=float(step_length), foot_spread=float(spread), distance=float(distance),
                      speed_generation="classic", add_force=True, F_ext=float(F_ext))
        cfg.method = "herdt"
        vx, vy, states = SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
        codes = g[f"p{k}/states"]
        assert np.array_equal(state_codes(states), codes)                   # host timeline == the reference's
        v_ref = np.stack([vx, vy], 1) * speed
        assert np.array_equal(v_ref, g[f"p{k}/v_ref"])
        ctrl = ZMPController(cfg)
        padded = np.concatenate([names[codes], np.repeat(names[codes][-1:], cfg.horizon)])
        assert np.array_equal(np.array(ctrl.find_nb_steps(padded), np.int32), g[f"p{k}/nb_steps"])
        com, y_hist, foot_hist = ctrl.generate_com_trajectory(np.zeros((3, 1)), np.zeros((3, 1)), v_ref=v_ref, state_ref=names[codes])
        m = int(g[f"p{k}/certified_ticks"]) + 1
        lengths.add(len(codes))
        assert m > 0.8 * len(codes)
        assert np.max(np.abs(foot_hist[:m] - g[f"p{k}/foot_hist"])) < 1e-6, k
        assert np.max(np.abs(y_hist[:m, :2, 0] - g[f"p{k}/y_hist"][:, :2])) < 1e-7, k          # position, velocity
        assert np.max(np.abs(y_hist[:m, 2, 0] - g[f"p{k}/y_hist"][:, 2])) < 1e-5, k            # acceleration carries the jerk
        assert np.max(np.abs(com[:m, 0] - g[f"p{k}/x_hist"][:, 0])) < 1e-7, k
    assert len(lengths) == 3
