"""Herdt free-footstep MPC on the batched engine (reference ``zmp_controller.py:435-826``).

Host side of the method: the support-phase layout of every tick's preview window (what the reference
recomputes inside ``predict_herdt_joint``, ``:581-603`` and ``:682-698``) and the rollout bookkeeping of
``generate_com_trajectory_herdt`` (``:450-531``: swing-foot interpolation, side toggle and footstep commit
on leaving SINGLE_SUPPORT, push, current-state tracking).  The joint QP of every tick is solved on the GPU
for all scenarios that share the timeline (``engine.HerdtTimeline.tick`` -> ``herdt_tick``).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .generators.cop_generator import State

STANDING, DOUBLE_SUPPORT, SINGLE_SUPPORT = 0, 1, 2


def state_codes(state_ref) -> np.ndarray:
    """``State`` members (or ints) -> int8 codes 0 STANDING / 1 DOUBLE_SUPPORT / 2 SINGLE_SUPPORT."""
    table = {State.STANDING: STANDING, State.DOUBLE_SUPPORT: DOUBLE_SUPPORT, State.SINGLE_SUPPORT: SINGLE_SUPPORT}
    return np.array([table.get(s, s) for s in state_ref], dtype=np.int8)


def support_runs(current_state: int, window: Sequence[int]) -> List[int]:
    """Run lengths of the support phases seen from the current tick (reference ``:581-594``): a run ends when
    the state changes, except DOUBLE_SUPPORT -> SINGLE_SUPPORT which is one footstep.  The current tick is
    counted in the first run, so the lengths sum to N + 1."""
    runs, count, s = [], 1, current_state
    for state in window:
        if state == s or (s == DOUBLE_SUPPORT and state == SINGLE_SUPPORT):
            count += 1
        else:
            runs.append(count)
            count = 1
        s = state
    runs.append(count)
    return runs


@dataclass
class TimelineLayout:
    """Everything about a walk that depends on the state timeline only (shared by a batch)."""

    codes: np.ndarray            # (n + N,) padded state codes (reference :468-470)
    current: np.ndarray          # (n-1,) state the controller is in at each tick
    foot_side: np.ndarray        # (n-1,) 0 left / 1 right at each tick
    commit: np.ndarray           # (n-1,) bool: the tick leaves SINGLE_SUPPORT (footstep is committed, side toggles)
    new_run: np.ndarray          # (n-1,) bool: the current support run started at this tick
    phase_of_row: np.ndarray     # (n-1, N) int8
    row_kind: np.ndarray         # (n-1, N) int8
    m: np.ndarray                # (n-1,) int32


def timeline_layout(state_ref, horizon: int) -> TimelineLayout:
    codes = state_codes(state_ref)
    n, N = len(codes), horizon
    ext = np.concatenate([codes, np.repeat(codes[-1:], N)])
    ticks = n - 1
    current = np.zeros(ticks, np.int8)
    side = np.zeros(ticks, np.int8)
    commit = np.zeros(ticks, bool)
    phase = np.zeros((ticks, N), np.int8)
    kind = np.zeros((ticks, N), np.int8)
    m = np.zeros(ticks, np.int32)
    new_run = np.zeros(ticks, bool)
    cur, foot = int(ext[0]), 0                       # reference :459, :457 ("left")
    for i in range(ticks):
        window = ext[i + 1: i + 1 + N]
        runs = support_runs(cur, window)
        current[i], side[i], m[i] = cur, foot, len(runs) - 1
        row = np.full(N, -2, np.int8)
        row[: runs[0]] = -1                          # U_c, :597-598
        start = runs[0]
        for k, length in enumerate(runs[1:]):        # U, :599-603 (clipped at N like the numpy slices)
            row[start: start + length] = k
            start += length
        phase[i] = row
        walking = window != STANDING
        kind[i][walking] = 1                         # ZMP box around the supporting foot, rows kept by :682-698
        if (cur == STANDING or not walking.any()) and (~walking).any():
            kind[i][~walking] = 2                    # standing hull, :715-771
        nxt = int(ext[i + 1])
        if nxt != cur and cur == SINGLE_SUPPORT:     # :501-503
            commit[i] = True
            foot = 1 - foot
        if nxt != cur:                               # :524-525
            if i + 1 < ticks and not (cur == DOUBLE_SUPPORT and nxt == SINGLE_SUPPORT):
                new_run[i + 1] = True                # the first future phase becomes the current one
            cur = nxt
    return TimelineLayout(ext, current, side, commit, new_run, phase, kind, m)


def polygon_vertices(polytope) -> np.ndarray:
    """Counter-clockwise hull vertices of a footstep polytope (what ``_polytope_halfspace`` turns into facets)."""
    from scipy.spatial import ConvexHull

    verts = np.asarray(polytope, dtype=float)
    return np.ascontiguousarray(verts[ConvexHull(verts).vertices])


class HerdtBatch:
    """Batched ``generate_com_trajectory_herdt`` for scenarios sharing one state timeline."""

    def __init__(self, config, state_ref, nb_steps, *, solver_options: Optional[dict] = None, warm_start: bool = True):
        from . import engine

        self.config = config
        self.layout = timeline_layout(state_ref, config.horizon)
        self.nb_steps = nb_steps          # find_nb_steps of the PADDED timeline (reference :471)
        self.warm_start = warm_start      # the reference cold-starts every tick; the optimum is the same
        opts = solver_options or {}
        self.plan = engine.HerdtPlan(config.horizon, config.h, config.g, config.dt, config.alpha, config.beta,
                                     config.gamma, config.foot_length, config.foot_width, config.foot_spread, **opts)
        self.timeline = self.plan.timeline(self.layout.phase_of_row, self.layout.row_kind, self.layout.m)
        self.polygons = (polygon_vertices(config.left_foot_polytope), polygon_vertices(config.right_foot_polytope))

    def rollout(self, x_init, y_init, v_ref, *, F_ext=None, mass=None, foot_spread=None,
                add_force: Optional[bool] = None, n_ticks: Optional[int] = None) -> Dict[str, "torch.Tensor"]:
        """x_init, y_init (B,3); v_ref (n,2) shared or (B,n,2).  Returns device tensors x_hist, y_hist (B,n,3),
        foot_hist (B,n,2), status (B,) = ticks whose QP failed (they fall back to zero jerk and the interpolated
        footstep like the reference, :796-804), accepted (B,) = ticks whose QP stopped short of eps but within the
        reference solver's own tolerance (applied, like the reference applies what OSQP calls "solved"), iters (B,)."""
        import torch

        cfg, lay = self.config, self.layout
        dev = self.plan.device
        N = cfg.horizon
        x = torch.as_tensor(np.asarray(x_init, float).reshape(-1, 3)).to(dev).contiguous()
        y = torch.as_tensor(np.asarray(y_init, float).reshape(-1, 3)).to(dev).contiguous()
        B = x.shape[0]
        v = np.asarray(v_ref, float)
        n = v.shape[-2]
        if n != len(lay.codes) - N:
            raise ValueError("v_ref and state_ref must have the same length")
        pad = np.repeat(v[..., -1:, :], N, axis=-2)                              # reference :465-467
        vref = torch.as_tensor(np.concatenate([v, pad], axis=-2)).to(dev).contiguous()
        fc = torch.zeros((B, 2), dtype=torch.float64, device=dev)
        spread = None
        if foot_spread is not None:                                               # a plan parameter per scenario
            spread = torch.as_tensor(np.broadcast_to(np.asarray(foot_spread, float), (B,)).copy()).to(dev)
        fc[:, 1] = cfg.foot_spread if spread is None else spread                  # :455-456
        air = fc.clone()
        push = cfg.add_force if add_force is None else add_force
        impulse = None
        if push:
            F = np.broadcast_to(np.asarray(cfg.F_ext if F_ext is None else F_ext, float), (B,))
            m_ = np.broadcast_to(np.asarray(cfg.m if mass is None else mass, float), (B,))
            impulse = torch.as_tensor(cfg.dt * F / m_).to(dev)
        if n_ticks is None:
            # the whole walk in one call: tick loop and foot bookkeeping run on the device (herdt_rollout)
            steps_left = np.array([self.nb_steps[i][0] for i in range(n - 1)], np.int32)
            out = self.timeline.rollout(x, y, fc, vref, n=n, foot_side=lay.foot_side, steps_left=steps_left, commit=lay.commit,
                                        new_run=lay.new_run, polygons=self.polygons, spread=spread, warm_start=self.warm_start,
                                        impulse=impulse, force_tick=n // 2 if push else -1)
            res = {k: out[k] for k in ("x_hist", "y_hist", "foot_hist", "iters", "_keepalive")}
            res["status"], res["accepted"] = out["status"] & 0xFFFF, out["status"] >> 16
            return res
        # a prefix of the walk, tick by tick from the host (the same kernels; used by tests and probes)
        ticks = n_ticks
        x_hist = torch.zeros((B, n, 3), dtype=torch.float64, device=dev)
        y_hist = torch.zeros((B, n, 3), dtype=torch.float64, device=dev)
        foot_hist = torch.zeros((B, n, 2), dtype=torch.float64, device=dev)
        x_hist[:, 0], y_hist[:, 0], foot_hist[:, 0] = x, y, fc
        status = torch.zeros((B,), dtype=torch.int32, device=dev)
        iters = torch.zeros((B,), dtype=torch.int64, device=dev)
        for i in range(ticks):
            side = int(lay.foot_side[i])
            has_step = lay.m[i] > 0
            out = self.timeline.tick(i, x, y, fc, vref, side, self.polygons[side] if has_step else None,
                                     air=air.contiguous(), spread=spread,
                                     warm=0 if (i == 0 or not self.warm_start) else (2 if lay.new_run[i] else 1))
            x, y = out["next_x"], out["next_y"]
            if has_step:
                air = air + (1.0 / self.nb_steps[i][0]) * (out["f1"] - air)      # :495-498
            if lay.commit[i]:                                                     # :501-517
                fc = out["f1"].clone() if has_step else air.clone()
                air = fc.clone()
            if push and i == n // 2:                                              # :522-523
                y = y.clone()
                y[:, 1] -= impulse
            x_hist[:, i + 1], y_hist[:, i + 1], foot_hist[:, i + 1] = x, y, fc
            st = out["status"].view(B, 2)
            failed = (st == 1).any(dim=1)
            status += torch.where(failed, 1, torch.where((st == 2).any(dim=1), 1 << 16, 0)).to(torch.int32)
            iters += out["iters"].view(B, 2).amax(dim=1)
        return {"x_hist": x_hist, "y_hist": y_hist, "foot_hist": foot_hist, "status": status & 0xFFFF, "accepted": status >> 16,
                "iters": iters}


# ====================================================================================== randomized plans (BASELINE config 4)
def randomized_plans(config, count: int, seed: int = 2, on_device: bool = False) -> Dict[str, object]:
    """BASELINE configs[3] / SURVEY 8(d) config 4: ``count`` footstep plans with step_length ~ U(0.2, 0.4),
    foot_spread ~ U(0.08, 0.12), distance ~ U(1.5, 2.5) (and a velocity scale ~ U(0.8, 1.2)), grouped by the state
    timeline of their gait (``CoPGenerator``'s state machine, `cop_generator.py:73-115`) -- scenarios of a group share
    the per-tick Hessians.  The reference velocity is ``speed_generation = "classic"`` (`speed_generation.py:48-51`:
    0.3 m/s forward while not STANDING).  ``on_device``: the timelines come from ``lipm_generate_bounds`` (bit-identical
    to the host generator, tested) instead of ``count`` runs of the Python state machine.
    Returns dict(step_length, foot_spread, distance, speed (count,), groups = list of dict(ids, states, v))."""
    from dataclasses import replace

    from .generators.speed_generation import CLASSIC_FORWARD_SPEED, SpeedTrajectoryGenerator

    rng = np.random.default_rng(seed)
    step_length = rng.uniform(0.2, 0.4, count)
    spread = rng.uniform(0.08, 0.12, count)
    distance = rng.uniform(1.5, 2.5, count)
    speed = rng.uniform(0.8, 1.2, count)
    groups: Dict[bytes, dict] = {}
    if on_device:
        from . import engine
        from .sweeps import series_capacity

        cap = series_capacity(replace(config, distance=2.5), [(0.2, 0.1)])
        table = engine.generate_bounds(config, distance, step_length, spread, n_max=cap)
        n = table["n"].cpu().numpy()
        if (n <= 0).any():
            raise RuntimeError("timeline capacity too small")
        codes = table["states"].cpu().numpy()
        for b in range(count):
            row = codes[b, : n[b]]
            key = row.tobytes()
            entry = groups.get(key)
            if entry is None:
                v_x = np.where(row == STANDING, 0.0, CLASSIC_FORWARD_SPEED)
                states = np.array([State.STANDING, State.DOUBLE_SUPPORT, State.SINGLE_SUPPORT], dtype=object)[row]
                entry = groups[key] = {"ids": [], "states": states, "v": np.stack([v_x, np.zeros_like(v_x)], axis=1)}
            entry["ids"].append(b)
    else:
        for b in range(count):
            cfg_b = replace(config, step_length=float(step_length[b]), foot_spread=float(spread[b]), distance=float(distance[b]),
                            speed_generation="classic")
            v_x, v_y, states = SpeedTrajectoryGenerator(cfg_b).generate_speed_and_state(save_footsteps=False)
            key = bytes(state_codes(states))
            entry = groups.get(key)
            if entry is None:
                entry = groups[key] = {"ids": [], "states": np.array(states), "v": np.stack([v_x, v_y], axis=1)}
            entry["ids"].append(b)
    for entry in groups.values():
        entry["ids"] = np.array(entry["ids"])
    return {"step_length": step_length, "foot_spread": spread, "distance": distance, "speed": speed,
            "groups": list(groups.values())}


def run_plan_groups(config, plans: Dict[str, object], controllers=None, *, solver_options: Optional[dict] = None):
    """Every timeline group of ``randomized_plans`` as one batch on its own stream (a group's whole walk is ONE enqueue,
    ``herdt_rollout``, so the groups run concurrently and the small ones fill the SMs the large ones leave idle).
    Returns (results per group as device-tensor dicts incl. ``com``, the controllers -- reusable: they hold the built
    timelines -- and the streams to synchronise on)."""
    import torch

    from .controllers import ZMPController

    groups = plans["groups"]
    if controllers is None:
        controllers = []
        for entry in groups:                 # per-timeline Hessians first: building them synchronises the device
            ctrl = ZMPController(config)
            if solver_options:
                ctrl.herdt_options = dict(solver_options)
            ctrl.prepare_herdt(entry["states"])
            controllers.append(ctrl)
    results, streams = [], []
    for entry, ctrl in zip(groups, controllers):
        ids = entry["ids"]
        v = entry["v"][None] * plans["speed"][ids, None, None]
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            res = ctrl.generate_com_trajectory_herdt_batched(np.zeros((len(ids), 3)), np.zeros((len(ids), 3)), v,
                                                             entry["states"], foot_spread=plans["foot_spread"][ids],
                                                             to_numpy=False)
        results.append(res)
        streams.append(stream)
    return results, controllers, streams
