"""Reference CoM speeds aligned with the gait timeline (input of the Herdt method).

Restates `generators/speed_generation.py:11-69` of the reference.  The "wieber"
mode is a second caller of the hot path: it runs a whole push-free Wieber walk
(`ZMPController.generate_state_trajectory_wieber`) on the GPU engine and reads
the velocities off the state histories.
"""

from __future__ import annotations

from typing import List, Tuple

import numpy as np

from ..config import MPCConfig
from .cop_generator import CoPGenerator, State

CLASSIC_FORWARD_SPEED = 0.3   # m/s while not standing (reference :49)


class SpeedTrajectoryGenerator:
    """Generates reference CoM speed trajectories aligned with CoP/state timeline."""

    def __init__(self, config: MPCConfig):
        from ..controllers.zmp_controller import ZMPController   # late: avoids an import cycle

        self.config = config
        self._cop_generator = CoPGenerator(config)
        self._zmp_controller = ZMPController(config)

    def generate_speed_and_state(self, save_footsteps: bool = True,
                                 output_dir: str = "results") -> Tuple[np.ndarray, np.ndarray, List[State]]:
        """Return ``(v_x (n,), v_y (n,), states)`` according to ``config.speed_generation``."""
        z_max, z_min, states = self._cop_generator.generate_cop_trajectory(
            save_footsteps=save_footsteps, output_dir=output_dir)
        mode = (self.config.speed_generation or "wieber").lower()

        if mode == "classic":
            v_x = np.array([0.0 if s == State.STANDING else CLASSIC_FORWARD_SPEED for s in states])
            return v_x, np.array([0.0 for _ in states]), states

        if mode == "wieber":
            rest = np.zeros((3, 1))
            x_hist, y_hist = self._zmp_controller.generate_state_trajectory_wieber(
                x_init=rest, y_init=rest.copy(), z_max=z_max, z_min=z_min)
            return x_hist[:, 1, 0], y_hist[:, 1, 0], states

        raise ValueError(f"Unknown speed_generation mode: {self.config.speed_generation}")

    def generate_speed_and_state_batched(self, distance, step_length, foot_spread, to_numpy: bool = True):
        """``generate_speed_and_state`` for V walk variants at once (additive API): per-variant ``distance``,
        ``step_length``, ``foot_spread`` (arrays of length V, or scalars), everything else from ``config``.

        The bound series and state timelines are generated on the device (``lipm_generate_bounds``, bit-identical to
        ``CoPGenerator``); in "wieber" mode the V push-free walks of ``generate_state_trajectory_wieber`` (reference
        `speed_generation.py:53-67`) run as ONE batched rollout that reads each variant's series through the bounds
        index.  Returns dict(v (V, n_max, 2), states (V, n_max) int8 codes, n (V,) samples per variant); rows past a
        variant's ``n`` are zero."""
        import torch

        from .. import engine
        from ..sweeps import series_capacity

        cfg = self.config
        d = np.atleast_1d(np.asarray(distance, dtype=np.float64))
        V = len(d)
        s = np.broadcast_to(np.asarray(step_length, dtype=np.float64), (V,))
        f = np.broadcast_to(np.asarray(foot_spread, dtype=np.float64), (V,))
        from dataclasses import replace
        cap = series_capacity(replace(cfg, distance=float(d.max())), [(float(s.min()), 0.1)])
        table = engine.generate_bounds(cfg, d, s, f, n_max=cap)
        n = table["n"]
        if int(n.min().item()) <= 0:
            raise RuntimeError("bound series capacity too small")
        mode = (cfg.speed_generation or "wieber").lower()
        states = table["states"]
        if mode == "classic":
            valid = torch.arange(cap, device=n.device)[None, :] < n[:, None]
            v = torch.zeros((V, cap, 2), dtype=torch.float64, device=n.device)
            v[..., 0] = ((states != 0) & valid).to(torch.float64) * CLASSIC_FORWARD_SPEED      # float64: 0.3 exactly as the host value
        elif mode == "wieber":
            plan = self._zmp_controller._plan()
            zero = torch.zeros((V, 3), dtype=torch.float64, device=n.device)
            out = plan.rollout(1 if cfg.strict else 0, zero, zero, table["z_min"], table["z_max"], n_steps=n,
                               bounds_index=torch.arange(V, dtype=torch.int32, device=n.device), want_flags=False)
            v = torch.stack([out["x_hist"][..., 1], out["y_hist"][..., 1]], dim=-1)
        else:
            raise ValueError(f"Unknown speed_generation mode: {cfg.speed_generation}")
        res = {"v": v, "states": states, "n": n}
        return engine.to_host(res) if to_numpy else res
