"""Linear Inverted Pendulum Model with jerk input, plus its preview matrices.

The reference keeps the model inside ``ZMPController.__init__``
(`controllers/zmp_controller.py:15-21`) and rebuilds the preview matrices with a
Python double loop on every solve (`:162-171`, `:566-579`).  Its documentation
still advertises ``models/lipm_model.py`` / ``LIPMModel`` (README.md:119-121);
the surface recovered from the stale ``__pycache__/model.cpython-312.pyc`` is
``LIPMModel(config)`` with ``g, h, dt, A, B, C`` and ``step / get_zmp /
get_state_dimension``.  This module provides that surface and closed-form,
vectorised builders for the preview matrices (host side; the device builds its
own copy inside the plan, see csrc/lipm_plan.cu).
"""

from __future__ import annotations

import numpy as np

from ..config import MPCConfig


def discrete_model(dt: float, h: float, g: float):
    """(A, B, C) of the sampled triple integrator; state = [pos, vel, acc]."""
    T = dt
    A = np.array([[1.0, T, T ** 2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]])
    B = np.array([T ** 3 / 6.0, T ** 2 / 2.0, T]).reshape((3, 1))
    C = np.array([1.0, 0.0, -h / g])
    return A, B, C


def zmp_preview_matrices(horizon: int, dt: float, h: float, g: float):
    """Px (N,3) and Pu (N,N): ZMP over the window = Px @ x + Pu @ jerks.

    Row i is time k+i+1.  ``Px[i] = C A^(i+1)`` and ``Pu[i,j] = C A^(i-j) B``;
    the literal formulas follow `zmp_controller.py:166-171` term by term so that
    the entries are bit-identical to the reference's loop.
    """
    T = dt
    i = np.arange(horizon)
    Px = np.empty((horizon, 3))
    Px[:, 0] = 1
    Px[:, 1] = T * (i + 1)
    Px[:, 2] = (T ** 2) / 2 * (i + 1) ** 2 - h / g
    lag = i[:, None] - i[None, :]
    toeplitz = (T ** 3) / 6 * (1 + 3 * lag + 3 * lag ** 2) - T * h / g
    Pu = np.where(lag >= 0, toeplitz, 0.0)
    return Px, Pu


def velocity_preview_matrices(horizon: int, dt: float):
    """Pvs (N,3) and Pvu (N,N): CoM velocity over the window (Herdt cost, `:571-579`)."""
    T = dt
    i = np.arange(horizon)
    Pvs = np.zeros((horizon, 3))
    Pvs[:, 1] = 1.0
    Pvs[:, 2] = (i + 1) * T
    lag = i[:, None] - i[None, :]
    Pvu = np.where(lag >= 0, (T ** 2) / 2.0 * (2 * lag + 1), 0.0)
    return Pvs, Pvu


class LIPMModel:
    """CoM dynamics of one horizontal axis at constant height ``h``."""

    def __init__(self, config: MPCConfig):
        self.g = config.g
        self.h = config.h
        self.dt = config.dt
        self.A, self.B, c_row = discrete_model(self.dt, self.h, self.g)
        self.C = c_row.reshape((1, 3))

    def step(self, x: np.ndarray, u: np.ndarray) -> np.ndarray:
        return self.A @ x + self.B @ u

    def get_zmp(self, x: np.ndarray) -> np.ndarray:
        return self.C @ x

    def get_state_dimension(self) -> int:
        return self.A.shape[1]

    def preview(self, horizon: int):
        return zmp_preview_matrices(horizon, self.dt, self.h, self.g)
