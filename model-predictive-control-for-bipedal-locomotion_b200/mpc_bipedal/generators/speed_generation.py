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
