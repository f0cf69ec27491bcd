"""Alternating footstep plan for a straight walk.

Behavioural restatement of the reference's `generators/footstep_generator.py:7-49`
(plotting demo dropped; no matplotlib import at module load).  The plan is a
handful of scalars per walk, so it stays on the host.
"""

from __future__ import annotations

from typing import List, Tuple

# The reference hard-codes the sole size and friction of every contact (`:34`).
SOLE_SHAPE: Tuple[float, float] = (0.11, 0.05)
SOLE_FRICTION: float = 0.7


class Contact:
    """A foothold: centre, rectangular sole, and the ZMP box it supports."""

    def __init__(self, x: float, y: float, shape: Tuple[float, float], friction: float):
        self.x = x
        self.y = y
        self.shape = shape
        self.friction = friction
        half_len, half_wid = shape[0] / 2, shape[1] / 2
        self.z_max = [x + half_len, y + half_wid]
        self.z_min = [x - half_len, y - half_wid]

    def __repr__(self) -> str:  # debugging aid only
        return f"Contact(x={self.x!r}, y={self.y!r})"


def generate_footsteps(distance: float, step_length: float, foot_spread: float) -> List[Contact]:
    """Footholds from a two-feet stance at x=0 to a two-feet stance at ``distance``.

    Feet alternate sides every step.  Once the remaining distance is at most one
    step, the stride is shortened to ``min(remaining, step_length / 2)``; the walk
    ends with the trailing foot brought next to the leading one.
    """
    plan = [
        Contact(0.0, -foot_spread, SOLE_SHAPE, SOLE_FRICTION),
        Contact(0.0, +foot_spread, SOLE_SHAPE, SOLE_FRICTION),
    ]
    x, side = 0.0, foot_spread
    while x < distance:
        remaining = distance - x
        stride = min(remaining, 0.5 * step_length) if remaining <= step_length else step_length
        x += stride
        side = -side
        plan.append(Contact(x, side, SOLE_SHAPE, SOLE_FRICTION))
    plan.append(Contact(x, -side, SOLE_SHAPE, SOLE_FRICTION))
    return plan
