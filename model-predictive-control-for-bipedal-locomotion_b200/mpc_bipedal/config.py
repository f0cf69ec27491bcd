"""Parameter record shared by the CoP generator and the MPC controller.

This mirrors the reference's single configuration object
(`src/mpc_bipedal/config.py:13-87`): the field names, defaults and the rule
``dt = 1.5 / horizon`` when ``dt`` is not given are the JSON schema of
``configs/default.json`` (``{"mpc": {...}}``) and must not drift.  Batched
sweeps never extend this record; per-scenario quantities (F_ext, m, push tick,
bounds) are passed to the batched entry points as arrays instead.
"""

from __future__ import annotations

import json
from dataclasses import asdict, dataclass, fields
from typing import Any, Dict, Optional, Tuple

Polytope = Tuple[Tuple[float, float], ...]

# Relative vertices [dx, dy] that the next footstep must stay inside when the
# left / right foot swings (reference: config.py:57-82).  The right polytope is
# the mirror image of the left one about the x axis.
_LEFT_SWING_POLYTOPE: Polytope = (
    (-0.1, -0.3), (-0.1, -0.4), (0.0, -0.4), (0.0, -0.2), (0.1, -0.17), (0.2, -0.13),
    (0.3, -0.1), (0.7, -0.05), (0.8, -0.05), (0.8, -0.3), (0.4, -0.35),
)
_RIGHT_SWING_POLYTOPE: Polytope = tuple((dx, -dy) for dx, dy in _LEFT_SWING_POLYTOPE)


@dataclass
class MPCConfig:
    """All knobs of the walking-pattern generator (flat, JSON-serialisable)."""

    # gait timeline (seconds) and footstep plan (metres) -- CoP generator side
    ssp_duration: float = 24 * 0.01
    dsp_duration: float = 1 * 0.01
    standing_duration: float = 50 * 0.01
    distance: float = 2.1
    step_length: float = 0.3
    foot_spread: float = 0.1

    # sample time; None means "derive from the horizon" (see __post_init__)
    dt: Optional[float] = None

    # preview controller
    horizon: int = 150
    Q: float = 1.0
    R: float = 1e-6
    S: float = 1.0
    h: float = 0.75
    g: float = 9.81
    m: float = 40.0
    F_ext: float = 400.0
    strict: bool = True
    add_force: bool = True

    # walking-pattern method ("wieber" | "herdt") and Herdt weights / geometry
    method: str = "wieber"
    alpha: float = 1e-6
    beta: float = 1.0
    gamma: float = 1.0
    vx_ref: float = 0.0
    vy_ref: float = 0.0
    foot_length: float = 0.11
    foot_width: float = 0.05
    v_max_x: float = 0.9
    v_max_y: float = 0.5
    # "classic": constant forward speed; "wieber": speeds taken from a Wieber rollout
    speed_generation: str = "classic"
    left_foot_polytope: tuple = _LEFT_SWING_POLYTOPE
    right_foot_polytope: tuple = _RIGHT_SWING_POLYTOPE

    def __post_init__(self) -> None:
        # The preview window always spans 1.5 s (reference config.py:84-87).
        if self.dt is None:
            self.dt = 1.5 / self.horizon

    # ------------------------------------------------------------------ helpers
    @classmethod
    def field_names(cls) -> Tuple[str, ...]:
        return tuple(f.name for f in fields(cls))

    @classmethod
    def from_dict(cls, section: Dict[str, Any]) -> "MPCConfig":
        """Build from the ``"mpc"`` section of a config file.

        ``dt`` is dropped so that the horizon rule wins, exactly like the
        drivers' ``load_config_from_json`` (reference scripts/run_mpc.py:23-39).
        Unknown keys raise ``TypeError`` like the dataclass constructor does.
        """
        values = dict(section)
        values.pop("dt", None)
        for key in ("left_foot_polytope", "right_foot_polytope"):
            if key in values:
                values[key] = tuple(tuple(v) for v in values[key])
        return cls(**values)

    @classmethod
    def from_json(cls, path: str) -> "MPCConfig":
        with open(path, "r") as handle:
            document = json.load(handle)
        return cls.from_dict(document.get("mpc", {}))

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)
