#!/usr/bin/env python3
"""Regenerate the golden fixtures in this directory.

Runs ONLY in the build container, where ``/root/reference`` is mounted: it
imports the reference package itself (with stub ``cvxpy`` / ``matplotlib``
modules, because neither is installed and the non-strict path and the
generators do not need them) and records its outputs.  The GPU box has no
``/root/reference``; tests there read the committed ``.npz`` files.

Fixtures written
----------------
ref_generators.npz     reference generators: footsteps, z_max/z_min/states for several walks
ref_nonstrict_walk.npz reference ``generate_com_trajectory_wieber`` (strict=False, push) and
                       ``generate_state_trajectory_wieber`` on configs/default.json
ref_nonstrict_ticks.npz reference ``predict_wieber_axis`` (strict=False) on seeded random inputs
ref_herdt_helpers.npz  reference ``find_nb_steps`` and ``_polytope_halfspace`` outputs
ref_speed_wieber.npz   reference ``SpeedTrajectoryGenerator`` in "wieber" mode (strict=False) on two configs
oracle_herdt_walk.npz  ORACLE Herdt walk of the class-default config (--herdt)
oracle_herdt_plans.npz ORACLE Herdt walks of three randomized plans with different timelines (--herdt-plans)
oracle_strict_walk.npz ORACLE (not reference) strict default walk: exact BVLS solver
oracle_strict_ticks.npz ORACLE strict single ticks on seeded random inputs (N = 16, 50, 150)

Usage:  python tests/golden/make_golden.py [--only NAME ...]
"""

import argparse
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "cvxpy"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["cvxpy"].OSQP = "OSQP"
    sys.path.insert(0, REFERENCE)
    from src.mpc_bipedal.config import MPCConfig
    from src.mpc_bipedal.controllers.zmp_controller import ZMPController
    from src.mpc_bipedal.generators.cop_generator import CoPGenerator, State
    from src.mpc_bipedal.generators.footstep_generator import generate_footsteps
    return MPCConfig, ZMPController, CoPGenerator, State, generate_footsteps


STATE_CODE = {"STANDING": 0, "DOUBLE_SUPPORT": 1, "SINGLE_SUPPORT": 2}

# walks used for generator parity: (tag, overrides on top of configs/default.json or class defaults)
WALKS = [
    ("default_json", "json", {}),
    ("class_defaults", "class", {}),
    ("short_steps", "json", {"step_length": 0.2, "distance": 1.5}),
    ("long_steps", "json", {"step_length": 0.4, "foot_spread": 0.12, "distance": 2.5}),
    ("narrow", "class", {"foot_spread": 0.08, "step_length": 0.25, "distance": 1.9}),
    ("h100", "json", {"horizon": 100}),
    ("h300", "class", {"horizon": 300, "step_length": 0.35}),
    ("odd", "json", {"step_length": 0.33, "distance": 2.0, "ssp_duration": 0.31, "dsp_duration": 0.05}),
]


def default_json_section():
    with open(os.path.join(REFERENCE, "configs", "default.json")) as fh:
        return json.load(fh)["mpc"]


def make_config(MPCConfig, base, overrides):
    values = dict(default_json_section()) if base == "json" else {}
    values.update(overrides)
    values.pop("dt", None)
    return MPCConfig(**values)


def gen_generators(out):
    MPCConfig, _, CoPGenerator, _, generate_footsteps = import_reference()
    data = {}
    for tag, base, over in WALKS:
        cfg = make_config(MPCConfig, base, over)
        steps = generate_footsteps(cfg.distance, cfg.step_length, cfg.foot_spread)
        data[f"{tag}/footsteps"] = np.array([[c.x, c.y, *c.z_max, *c.z_min] for c in steps])
        z_max, z_min, states = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
        data[f"{tag}/z_max"] = z_max
        data[f"{tag}/z_min"] = z_min
        data[f"{tag}/states"] = np.array([STATE_CODE[s.value] for s in states], np.int8)
        data[f"{tag}/config"] = np.array(json.dumps({"base": base, "overrides": over}))
    np.savez_compressed(out, **data)


def gen_nonstrict_walk(out):
    MPCConfig, ZMPController, CoPGenerator, _, _ = import_reference()
    cfg = make_config(MPCConfig, "json", {"strict": False})
    z_max, z_min, _ = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
    ctrl = ZMPController(cfg)
    zero = np.zeros((3, 1))
    com, y_hist = ctrl.generate_com_trajectory(zero.copy(), zero.copy(), z_max, z_min)
    x_state, y_state = ctrl.generate_state_trajectory_wieber(zero.copy(), zero.copy(), z_max, z_min)
    np.savez_compressed(out, z_max=z_max, z_min=z_min, com=com, y_hist=y_hist,
                        nopush_x_hist=x_state, nopush_y_hist=y_state,
                        impulse=np.array(cfg.dt * cfg.F_ext / cfg.m), force_tick=np.array(len(z_min) // 2))


def gen_nonstrict_ticks(out):
    MPCConfig, ZMPController, _, _, _ = import_reference()
    rng = np.random.default_rng(20261018)
    data = {}
    for N in (10, 50, 100, 150):
        for h in (0.75, 0.62):
            cfg = MPCConfig(horizon=N, h=h, strict=False)
            ctrl = ZMPController(cfg)
            cases = 6
            x = rng.normal(size=(cases, 3)) * np.array([0.05, 0.1, 0.5])
            centre = rng.uniform(-0.05, 0.05, size=(cases, 1)) + np.cumsum(
                (rng.random((cases, N)) < 0.04) * rng.uniform(-0.1, 0.1, size=(cases, N)), axis=1)
            z_min, z_max = centre - 0.055, centre + 0.055
            nxt = np.stack([
                ctrl.predict_wieber_axis(x[k].reshape(3, 1), N, z_max[k].reshape(N, 1), z_min[k].reshape(N, 1))[:, 0]
                for k in range(cases)])
            key = f"N{N}_h{h}"
            data[f"{key}/x"], data[f"{key}/z_min"], data[f"{key}/z_max"], data[f"{key}/next"] = x, z_min, z_max, nxt
    np.savez_compressed(out, **data)


def gen_herdt_helpers(out):
    MPCConfig, ZMPController, _, State, _ = import_reference()
    from src.mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator
    data = {}
    code_to_state = {0: State.STANDING, 1: State.DOUBLE_SUPPORT, 2: State.SINGLE_SUPPORT}
    for tag, base, over in WALKS[:5]:
        cfg = make_config(MPCConfig, base, over)
        vx, vy, states = SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
        codes = np.array([STATE_CODE[s.value] for s in states], np.int8)
        padded = np.concatenate([codes, np.repeat(codes[-1:], cfg.horizon)])
        ctrl = ZMPController(cfg)
        nb = ctrl.find_nb_steps(np.array([code_to_state[int(c)] for c in padded], dtype=object))
        data[f"{tag}/vx"], data[f"{tag}/vy"], data[f"{tag}/states"] = np.asarray(vx), np.asarray(vy), codes
        data[f"{tag}/nb_steps"] = np.array(nb, np.int32)
    cfg = MPCConfig()
    ctrl = ZMPController(cfg)
    for side in ("left", "right"):
        A, b = ctrl._polytope_halfspace(np.array(getattr(cfg, f"{side}_foot_polytope")))
        data[f"polytope/{side}/A"], data[f"polytope/{side}/b"] = A, b
    np.savez_compressed(out, **data)


def _oracle():
    sys.path.insert(0, REPO)
    from oracle import lipm_oracle
    return lipm_oracle


def gen_oracle_strict_walk(out):
    MPCConfig, _, CoPGenerator, _, _ = import_reference()
    orc = _oracle()
    cfg = make_config(MPCConfig, "json", {})
    z_max, z_min, _ = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
    p = orc.Problem.make(cfg.horizon, cfg.h, cfg.g, cfg.Q, cfg.R)
    impulse = cfg.dt * cfg.F_ext / cfg.m
    res = orc.rollout_wieber(p, np.zeros(3), np.zeros(3), z_max, z_min, strict=True,
                             impulse=impulse, force_tick=len(z_min) // 2, solver="exact")
    np.savez_compressed(out, z_max=z_max, z_min=z_min, x_hist=res["x_hist"], y_hist=res["y_hist"],
                        flags=res["flags"], margin=res["margin"], impulse=np.array(impulse),
                        force_tick=np.array(len(z_min) // 2))


def gen_oracle_strict_ticks(out):
    orc = _oracle()
    rng = np.random.default_rng(7)
    data = {}
    for N in (16, 50, 150):
        p = orc.Problem.make(N)
        cases = 24
        x = rng.normal(size=(cases, 3)) * np.array([0.05, 0.1, 0.5])
        centre = rng.uniform(-0.05, 0.05, size=(cases, 1)) + np.cumsum(
            (rng.random((cases, N)) < 0.04) * rng.uniform(-0.1, 0.1, size=(cases, N)), axis=1)
        z_min, z_max = centre - 0.055, centre + 0.055
        nxt, zs, act = [], [], []
        for k in range(cases):
            s, z, a = orc.strict_axis_exact(p, x[k], z_max[k], z_min[k])
            nxt.append(s), zs.append(z), act.append(a)
        key = f"N{N}"
        data[f"{key}/x"], data[f"{key}/z_min"], data[f"{key}/z_max"] = x, z_min, z_max
        data[f"{key}/next"], data[f"{key}/z"], data[f"{key}/active"] = np.array(nxt), np.array(zs), np.array(act)
    np.savez_compressed(out, **data)


def gen_speed_wieber(out):
    """REFERENCE ``SpeedTrajectoryGenerator`` in "wieber" mode (speed_generation.py:53-67): a whole push-free Wieber walk
    through the reference's own ``generate_state_trajectory_wieber`` (non-strict: runs without cvxpy), velocities read
    off the state histories.  Two configs: default.json and a short walk with another step length."""
    MPCConfig, _, _, _, _ = import_reference()
    from src.mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator
    data = {}
    for tag, over in (("default_json", {"strict": False, "speed_generation": "wieber"}),
                      ("short", {"strict": False, "speed_generation": "wieber", "distance": 0.9, "step_length": 0.25,
                                 "standing_duration": 0.4, "horizon": 100})):
        cfg = make_config(MPCConfig, "json", over)
        vx, vy, states = SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
        data[f"{tag}/vx"], data[f"{tag}/vy"] = np.asarray(vx), np.asarray(vy)
        data[f"{tag}/states"] = np.array([STATE_CODE[s.value] for s in states], np.int8)
    np.savez_compressed(out, **data)


TARGETS = {
    "ref_speed_wieber": gen_speed_wieber,
    "ref_generators": gen_generators,
    "ref_nonstrict_walk": gen_nonstrict_walk,
    "ref_nonstrict_ticks": gen_nonstrict_ticks,
    "ref_herdt_helpers": gen_herdt_helpers,
    "oracle_strict_walk": gen_oracle_strict_walk,
    "oracle_strict_ticks": gen_oracle_strict_ticks,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=list(TARGETS))
    args = ap.parse_args()
    if not os.path.isdir(REFERENCE):
        sys.exit("make_golden.py needs /root/reference (build container only)")
    for name in args.only:
        path = os.path.join(HERE, name + ".npz")
        print("writing", path, flush=True)
        TARGETS[name](path)


# ----------------------------------------------------------------------------- Herdt (oracle, not reference)
def make_oracle_herdt_walk():
    """ORACLE Herdt walk (KKT-certified QP solves) on the class-default config the reference's
    run_herdt_method.py uses (:53), timeline / nb_steps / polytope facets taken from the REFERENCE outputs
    recorded in ref_herdt_helpers.npz.  ~5 minutes of CPU."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"))
    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.models.lipm_model import discrete_model
    from oracle import herdt_oracle as ho

    g = np.load(os.path.join(HERE, "ref_herdt_helpers.npz"))
    cfg = MPCConfig()
    cfg.add_force = True
    A, B, _ = discrete_model(cfg.dt, cfg.h, cfg.g)
    states = g["class_defaults/states"]
    v_ref = np.stack([g["class_defaults/vx"], g["class_defaults/vy"]], 1)
    polys = {s: (g[f"polytope/{s}/A"], g[f"polytope/{s}/b"]) for s in ("left", "right")}
    res = ho.rollout_herdt(cfg, A, B, np.zeros(3), np.zeros(3), v_ref, states, g["class_defaults/nb_steps"], polys)
    # The walk is certified tick by tick up to the final stance.  There the reference's formulation commits the
    # "footstep" of the final STANDING phase (forced one step AHEAD by the reachable polytope, :773-790) and then
    # asks the ZMP to stay in a hull that lies in front of the decelerating CoM (:715-771): every row of the QP is
    # active, the CoM is pushed backwards and the solves no longer converge (the reference would print "Joint QP
    # solver failed" and fall back, :796-804).  Only the certified prefix is a golden vector.
    ok = np.array([c["ok"] for c in res["kkt"]])
    last = int(np.where(~ok)[0][0]) if (~ok).any() else len(ok)
    np.savez_compressed(os.path.join(HERE, "oracle_herdt_walk.npz"), x_hist=res["x_hist"][: last + 1],
                        y_hist=res["y_hist"][: last + 1], foot_hist=res["foot_hist"][: last + 1], certified_ticks=last,
                        final_com_uncertified=np.array([res["x_hist"][-1, 0], res["y_hist"][-1, 0]]))
    print("oracle_herdt_walk.npz written; certified ticks:", last)


def make_oracle_herdt_plans():
    """ORACLE Herdt walks of three RANDOMIZED footstep plans with three different timelines (BASELINE configs[3]:
    step_length ~ U(0.2, 0.4), foot_spread ~ U(0.08, 0.12), distance ~ U(1.5, 2.5), velocity scale ~ U(0.8, 1.2), seed 2
    -- the draw of mpc_bipedal.herdt.randomized_plans).  Timeline, reference velocities and nb_steps come from the
    REFERENCE's own SpeedTrajectoryGenerator / find_nb_steps (imported here); each walk is kept up to its last
    KKT-certified tick.  ~20 minutes of CPU."""
    MPCConfig_ref, ZMPController_ref, _, State, _ = import_reference()
    from src.mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator as RefSpeed
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200"))
    from dataclasses import replace

    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.herdt import randomized_plans
    from mpc_bipedal.models.lipm_model import discrete_model
    from oracle import herdt_oracle as ho

    g = np.load(os.path.join(HERE, "ref_herdt_helpers.npz"))
    polys = {s: (g[f"polytope/{s}/A"], g[f"polytope/{s}/b"]) for s in ("left", "right")}
    base = MPCConfig()
    plans = randomized_plans(base, 8192, seed=2)
    groups = sorted(plans["groups"], key=lambda e: len(e["ids"]))
    picks = [int(groups[0]["ids"][0]), int(groups[len(groups) // 2]["ids"][3]), int(groups[-1]["ids"][7])]
    code_to_state = {0: State.STANDING, 1: State.DOUBLE_SUPPORT, 2: State.SINGLE_SUPPORT}
    data = {"plan_ids": np.array(picks)}
    for k, b in enumerate(picks):
        over = dict(step_length=float(plans["step_length"][b]), foot_spread=float(plans["foot_spread"][b]),
                    distance=float(plans["distance"][b]))
        ref_cfg = MPCConfig_ref(**over)
        ref_cfg.speed_generation = "classic"
        vx, vy, states = RefSpeed(ref_cfg).generate_speed_and_state(save_footsteps=False)
        codes = np.array([STATE_CODE[s.value] for s in states], np.int8)
        padded = np.concatenate([codes, np.repeat(codes[-1:], ref_cfg.horizon)])
        nb = np.array(ZMPController_ref(ref_cfg).find_nb_steps(np.array([code_to_state[int(c)] for c in padded], dtype=object)), np.int32)
        cfg = replace(base, **over)
        cfg.add_force = True
        cfg.F_ext = 60.0 + 40.0 * k
        A, B, _ = discrete_model(cfg.dt, cfg.h, cfg.g)
        v_ref = np.stack([np.asarray(vx), np.asarray(vy)], 1) * float(plans["speed"][b])
        res = ho.rollout_herdt(cfg, A, B, np.zeros(3), np.zeros(3), v_ref, codes, nb, polys, stop_when_uncertified=True)
        ok = np.array([c["ok"] for c in res["kkt"]])
        last = int(np.where(~ok)[0][0]) if (~ok).any() else len(ok)
        print(f"plan {b}: n = {len(codes)}, certified ticks = {last}", flush=True)
        data[f"p{k}/states"], data[f"p{k}/v_ref"], data[f"p{k}/nb_steps"] = codes, v_ref, nb
        data[f"p{k}/params"] = np.array([over["step_length"], over["foot_spread"], over["distance"], float(plans["speed"][b]), cfg.F_ext])
        data[f"p{k}/x_hist"], data[f"p{k}/y_hist"] = res["x_hist"][: last + 1], res["y_hist"][: last + 1]
        data[f"p{k}/foot_hist"], data[f"p{k}/certified_ticks"] = res["foot_hist"][: last + 1], last
    np.savez_compressed(os.path.join(HERE, "oracle_herdt_plans.npz"), **data)
    print("oracle_herdt_plans.npz written")


if __name__ == "__main__":
    if "--herdt-plans" in sys.argv:
        make_oracle_herdt_plans()
    elif "--herdt" in sys.argv:
        make_oracle_herdt_walk()
    else:
        main()
