"""Per-sample ZMP bounds (support-polygon boxes) and the gait-phase timeline.

Restates `generators/cop_generator.py:11-115` of the reference.  The timeline is
driven by a floating-point clock that accumulates ``t += dt`` and compares it to
an accumulated switching time; the resulting phase lengths are irregular (e.g.
24- and 25-sample single supports for configs/default.json) and are part of the
behaviour the controller sees, so the same accumulation order is kept here.
The footstep picture is only drawn on request and matplotlib is imported lazily.
"""

from __future__ import annotations

from enum import Enum
from typing import List, Tuple

import numpy as np

from ..config import MPCConfig
from .footstep_generator import Contact, generate_footsteps


class State(Enum):
    """Gait phase of one bound sample."""
    STANDING = 'STANDING'
    DOUBLE_SUPPORT = 'DOUBLE_SUPPORT'
    SINGLE_SUPPORT = 'SINGLE_SUPPORT'


def _draw_footsteps(footsteps: List[Contact], output_dir: str) -> None:
    import os
    import matplotlib.pyplot as plt  # optional dependency, only for the picture

    os.makedirs(output_dir, exist_ok=True)
    fig, ax = plt.subplots()
    for c in footsteps:
        w, h = c.shape
        ax.add_patch(plt.Rectangle((c.x - w / 2, c.y - h / 2), w, h, edgecolor='b', facecolor='none'))
    ax.scatter([c.x for c in footsteps], [c.y for c in footsteps], color='r', s=0.2)
    ax.set_xlabel("X")
    ax.set_ylabel("Y")
    ax.set_title("Footsteps (rectangles centered on contacts)")
    ax.set_aspect('equal')
    plt.savefig(f'{output_dir}/footsteps.png')
    plt.close(fig)


class CoPGenerator:
    """Turns a footstep plan into ZMP box bounds sampled every ``dt``."""

    def __init__(self, config: MPCConfig):
        if config.dt is None:
            raise ValueError("dt must be set in MPCConfig (it is shared with the controller)")
        self.ssp_duration = config.ssp_duration
        self.dsp_duration = config.dsp_duration
        self.standing_duration = config.standing_duration
        self.dt = config.dt
        self.distance = config.distance
        self.step_length = config.step_length
        self.foot_spread = config.foot_spread

    # Phase graph: STANDING -> DS -> SS -> DS -> SS ... -> DS -> STANDING -> end.
    # Returns (new_state, added_duration, footsteps_consumed).
    def _advance(self, state: State, on_last_footstep: bool):
        if state is State.STANDING:
            if on_last_footstep:
                return state, 0.0, 1          # final stance elapsed: terminate
            return State.DOUBLE_SUPPORT, self.dsp_duration, 0
        if state is State.SINGLE_SUPPORT:
            return State.DOUBLE_SUPPORT, self.dsp_duration, 1
        if state is State.DOUBLE_SUPPORT:
            if on_last_footstep:
                return State.STANDING, self.standing_duration, 0
            return State.SINGLE_SUPPORT, self.ssp_duration, 0
        raise ValueError(f"Invalid state: {state}")

    def generate_cop_trajectory(
        self, save_footsteps: bool = True, output_dir: str = 'results'
    ) -> Tuple[np.ndarray, np.ndarray, List[State]]:
        """Return ``(z_max (n,2), z_min (n,2), states)`` for the configured walk."""
        footsteps = generate_footsteps(
            distance=self.distance, step_length=self.step_length, foot_spread=self.foot_spread
        )
        if save_footsteps:
            _draw_footsteps(footsteps, output_dir)

        last = len(footsteps) - 1
        current = 1
        state = State.STANDING
        clock = 0.0
        switch_at = self.standing_duration
        upper, lower, states = [], [], []

        while current < len(footsteps):
            if clock > switch_at:
                terminal_stance = state is State.STANDING and current == last
                state, extra, consumed = self._advance(state, current == last)
                if not terminal_stance:
                    switch_at += extra
                current += consumed

            if current < len(footsteps):
                lead = footsteps[current]
                if state is State.SINGLE_SUPPORT:
                    upper.append([lead.z_max[0], lead.z_max[1]])
                    lower.append([lead.z_min[0], lead.z_min[1]])
                else:
                    # both feet on the ground: bounding box of the two soles
                    trail = footsteps[current - 1]
                    upper.append([max(trail.z_max[0], lead.z_max[0]), max(trail.z_max[1], lead.z_max[1])])
                    lower.append([min(trail.z_min[0], lead.z_min[0]), min(trail.z_min[1], lead.z_min[1])])
                states.append(state)

            clock += self.dt

        return np.array(upper), np.array(lower), states
