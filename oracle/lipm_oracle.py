"""CPU oracle for the receding-horizon ZMP preview controller (TEST INFRASTRUCTURE).

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The shipped path is the CUDA library
(``include/lipm_mpc.h``) and must fail loudly when that library is missing.

It restates, in plain NumPy float64, the algorithm of the reference
``/root/reference/src/mpc_bipedal/controllers/zmp_controller.py``:

* model ``A, B, C``                                   -- ``:15-21``
* preview matrices ``Px, Pu``                         -- ``:162-171``
* non-strict closed form                              -- ``:196-199``
* strict QP (objective, box rows)                     -- ``:173-195``
* receding-horizon rollout, bound padding, push       -- ``:59-108`` (``:110-147`` without push)

Pinning status
--------------
* Non-strict path and the generators: PINNED against the reference itself,
  imported in the build container with stub ``cvxpy``/``matplotlib`` modules;
  the resulting vectors are committed under ``tests/golden/`` together with
  ``tests/golden/make_golden.py``.
* Strict path: **parity unpinned**.  The reference delegates the arithmetic to
  cvxpy -> OSQP (``requirements.txt:2``: ``cvxpy>=1.2.0``, un-pinned; OSQP
  transitive), neither vendored nor installed, and ships no test or golden
  vector.  OSQP stops at eps 1e-5, so its output only approximates the QP's
  minimiser.  The QP is strictly convex (H >= R*I > 0) and always feasible
  (Pu is square, lower-triangular with non-zero diagonal), hence the minimiser
  is unique; the oracle is that unique minimiser, computed by two independent
  FP64 solvers that must agree (``strict_axis_exact``: SciPy BVLS in ZMP
  coordinates; ``strict_axis_admm``: ADMM run to 1e-13).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------- model


def model_matrices(dt: float, h: float, g: float):
    """A (3,3), B (3,), C (3,) -- zmp_controller.py:15-21."""
    T = dt
    A = np.array([[1.0, T, T ** 2 / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]])
    B = np.array([T ** 3 / 6.0, T ** 2 / 2.0, T])
    C = np.array([1.0, 0.0, -h / g])
    return A, B, C


def preview_matrices(N: int, dt: float, h: float, g: float):
    """Px (N,3), Pu (N,N) by the reference's literal loop -- zmp_controller.py:162-171."""
    Px = np.zeros((N, 3))
    Pu = np.zeros((N, N))
    T = dt
    for i in range(N):
        Px[i, 0] = 1
        Px[i, 1] = T * (i + 1)
        Px[i, 2] = (T ** 2) / 2 * (i + 1) ** 2 - h / g
        j = np.arange(i + 1)
        Pu[i, : i + 1] = (T ** 3) / 6 * (1 + 3 * (i - j) + 3 * (i - j) ** 2) - T * h / g
    return Px, Pu


@dataclass
class Problem:
    """Everything that is fixed for one (N, h, g, dt, Q, R)."""

    N: int
    dt: float
    h: float
    g: float
    Q: float
    R: float
    A: np.ndarray
    B: np.ndarray
    C: np.ndarray
    Px: np.ndarray
    Pu: np.ndarray

    @classmethod
    def make(cls, N: int, h: float = 0.75, g: float = 9.81, Q: float = 1.0, R: float = 1e-6,
             dt: Optional[float] = None) -> "Problem":
        dt = 1.5 / N if dt is None else dt
        A, B, C = model_matrices(dt, h, g)
        Px, Pu = preview_matrices(N, dt, h, g)
        return cls(N, dt, h, g, Q, R, A, B, C, Px, Pu)

    # cached derived matrices ------------------------------------------------
    def _cache(self):
        if not hasattr(self, "_c"):
            N = self.N
            Pinv = np.linalg.solve(self.Pu, np.eye(N))      # jerks = Pinv @ (zmp - Px x)
            G = self.R * Pinv.T @ Pinv
            Hz = self.Q * np.eye(N) + G                      # Hessian in ZMP coordinates
            Hu = self.Pu.T @ self.Pu + self.R / self.Q * np.eye(N)
            self._c = dict(Pinv=Pinv, G=G, Hz=Hz, Hu_inv=np.linalg.inv(Hu), GPx=G @ self.Px)
        return self._c


# ------------------------------------------------------------------ per-tick solves


def unconstrained_axis(p: Problem, x: np.ndarray, zmax: np.ndarray, zmin: np.ndarray) -> np.ndarray:
    """Next axis state, non-strict mode -- zmp_controller.py:196-199 (matrices cached)."""
    zref = (zmax + zmin) / 2
    jerks = -p._cache()["Hu_inv"] @ p.Pu.T @ (p.Px @ x - zref)
    return p.A @ x + p.B * jerks[0]


def strict_axis_exact(p: Problem, x: np.ndarray, zmax: np.ndarray, zmin: np.ndarray):
    """Unique minimiser of the strict QP (zmp_controller.py:173-191), via BVLS.

    With z = Px x + Pu J (the previewed ZMP itself) the QP reads
    ``min Q/2 |z - zref|^2 + R/2 |Pu^-1 (z - Px x)|^2  s.t. zmin <= z <= zmax``:
    a bounded-variable least-squares problem that SciPy solves by an active-set
    method to machine precision.  Returns (next_state, z, active) where
    active[j] = -1 / 0 / +1 for lower-active / free / upper-active rows.
    """
    from scipy.optimize import lsq_linear

    c = p._cache()
    N = p.N
    zref = (zmax + zmin) / 2
    sq, sr = np.sqrt(p.Q), np.sqrt(p.R)
    lhs = np.vstack([sq * np.eye(N), sr * c["Pinv"]])
    rhs = np.concatenate([sq * zref, sr * (c["Pinv"] @ (p.Px @ x))])
    sol = lsq_linear(lhs, rhs, bounds=(zmin, zmax), method="bvls", tol=1e-15, max_iter=10 * N)
    z = sol.x
    jerk0 = (z[0] - p.Px[0] @ x) / p.Pu[0, 0]
    return p.A @ x + p.B * jerk0, z, sol.active_mask.astype(np.int8)


def kkt_certificate_jerk(p: Problem, x: np.ndarray, zmax: np.ndarray, zmin: np.ndarray, z: np.ndarray,
                         bound_tol: float = 1e-12) -> dict:
    """Optimality certificate of a previewed-ZMP plan ``z`` in the reference's OWN coordinates and matrices.

    The reference's strict QP (zmp_controller.py:174-189) is over the jerks J:
        min 1/2 Q |Px x + Pu J - zref|^2 + 1/2 J'(R I)J     s.t.  [Pu; -Pu] J <= [zmax - Px x; -zmin + Px x].
    Both oracle solvers and every GPU kernel work in z = Px x + Pu J instead; a common-mode error in that change of
    variables would pass their mutual agreement.  Here J = Pu^-1 (z - Px x) is mapped back and the KKT conditions are
    evaluated literally: stationarity  Q Pu'(Pu J + Px x - zref) + R J + Pu'(lam_hi - lam_lo) = 0  fixes
    mu = lam_hi - lam_lo = -Pu^-T grad; optimality <=> mu vanishes on rows strictly inside the box, mu >= 0 only on
    rows at zmax, mu <= 0 only on rows at zmin, and z is inside the box.  Returns the residuals, relative to the larger of
    max|mu| and the magnitude of the gradient's own terms Q |z|_inf, Q |zref|_inf (the usual relative KKT residual)."""
    x = np.asarray(x, float).reshape(3)
    Pxx = p.Px @ x
    J = np.linalg.solve(p.Pu, z - Pxx)
    zref = (zmax + zmin) / 2
    z_back = Pxx + p.Pu @ J
    grad = p.Q * (p.Pu.T @ (z_back - zref)) + p.R * J                      # H = R I, :174
    mu = -np.linalg.solve(p.Pu.T, grad)
    at_hi = z >= zmax - bound_tol
    at_lo = z <= zmin + bound_tol
    free = ~(at_hi | at_lo)
    scale = max(float(np.abs(mu).max()), p.Q * float(np.abs(z_back).max()), p.Q * float(np.abs(zref).max()), 1e-300)
    wrong = 0.0
    if at_hi.any():
        wrong = max(wrong, float(np.max(-mu[at_hi & ~at_lo], initial=0.0)))
    if at_lo.any():
        wrong = max(wrong, float(np.max(mu[at_lo & ~at_hi], initial=0.0)))
    return dict(primal=float(max(0.0, np.max(z - zmax), np.max(zmin - z))),
                free_multiplier=float(np.abs(mu[free]).max(initial=0.0)) / scale,
                wrong_sign=wrong / scale, active_rows=int((at_hi | at_lo).sum()), scale=scale, jerk0=float(J[0]))


def admm_matrix(p: Problem, rho: float) -> np.ndarray:
    """(Hz + rho I)^-1 -- the shared iteration matrix."""
    return np.linalg.inv(p._cache()["Hz"] + rho * np.eye(p.N))


def strict_axis_admm(p: Problem, x: np.ndarray, zmax: np.ndarray, zmin: np.ndarray,
                     rho: float = 2.0, alpha: float = 1.6, eps: float = 1e-13,
                     max_iter: int = 5000, z0: Optional[np.ndarray] = None,
                     w0: Optional[np.ndarray] = None, M: Optional[np.ndarray] = None):
    """Same minimiser by fixed-rho ADMM with over-relaxation (independent second solver).

    Splitting: f(v) = 1/2 v'Hz v + q'v, g(z) = indicator(zmin <= z <= zmax), v = z;
    ``w`` is the scaled dual.  Returns (next_state, z, w, iterations).
    """
    c = p._cache()
    M = admm_matrix(p, rho) if M is None else M
    q = -p.Q * (zmax + zmin) / 2 - c["GPx"] @ x
    z = np.clip(np.zeros(p.N) if z0 is None else z0, zmin, zmax)
    w = np.zeros(p.N) if w0 is None else w0.copy()
    it = 0
    # Stop test on d = |v - z_old|: primal residual <= (1+alpha) d, dual residual <= rho alpha d; the
    # threshold is relative to the linear term (diverging "falls" have a huge q and an absolute
    # test would sit below their FP64 rounding floor).
    thr = eps * max(1.0, np.max(np.abs(q))) / max(1.0 + alpha, rho * alpha)
    for it in range(1, max_iter + 1):
        v = M @ (rho * (z - w) - q)
        d = np.max(np.abs(v - z))
        t = alpha * v + (1.0 - alpha) * z + w
        z = np.minimum(np.maximum(t, zmin), zmax)
        w = t - z
        if d <= thr:
            break
    jerk0 = (z[0] - p.Px[0] @ x) / p.Pu[0, 0]
    return p.A @ x + p.B * jerk0, z, w, it


# ------------------------------------------------------------------------- rollout

FLAG_ACTIVE_LO = 1      # first preview row sits on its lower bound (strict mode)
FLAG_ACTIVE_HI = 2      # first preview row sits on its upper bound
FLAG_VIOL_LO = 4        # realised ZMP below zmin - viol_tol
FLAG_VIOL_HI = 8        # realised ZMP above zmax + viol_tol
FLAG_NOT_CONVERGED = 16  # the QP of the tick that produced this sample hit max_iter


def rollout_wieber(p: Problem, x_init, y_init, z_max: np.ndarray, z_min: np.ndarray, *,
                   strict: bool, impulse: float = 0.0, force_tick: Optional[int] = None,
                   solver: str = "exact", viol_tol: float = 1e-9, n_ticks: Optional[int] = None):
    """Receding-horizon walk of one scenario -- zmp_controller.py:59-108.

    ``impulse`` = dt*F_ext/m is subtracted from the y velocity after the update
    of tick ``force_tick`` (the reference hard-wires ``n_steps // 2``, ``:90,105``;
    ``None`` = no push = ``generate_state_trajectory_wieber``, ``:110-147``).
    Returns dict(x_hist (n,3), y_hist (n,3), flags (n,2) uint8, max_violation, margin (n,2))
    where ``margin`` is the distance of the realised ZMP to the nearer bound
    (negative = outside), used by the tests' tie tolerance.
    """
    n = len(z_min)
    N = p.N
    hi = np.vstack([z_max, np.tile(z_max[-1:, :], (N, 1))])
    lo = np.vstack([z_min, np.tile(z_min[-1:, :], (N, 1))])
    state = [np.asarray(x_init, float).reshape(3).copy(), np.asarray(y_init, float).reshape(3).copy()]
    hist = [[state[0].copy()], [state[1].copy()]]
    flags = np.zeros((n, 2), np.uint8)
    warm = [None, None]
    ticks = n - 1 if n_ticks is None else n_ticks
    for i in range(ticks):
        for ax in (0, 1):
            hi_w, lo_w = hi[i + 1:i + 1 + N, ax], lo[i + 1:i + 1 + N, ax]
            if not strict:
                nxt = unconstrained_axis(p, state[ax], hi_w, lo_w)
            elif solver == "exact":
                nxt, z, act = strict_axis_exact(p, state[ax], hi_w, lo_w)
                flags[i + 1, ax] |= (FLAG_ACTIVE_LO if act[0] < 0 else 0) | (FLAG_ACTIVE_HI if act[0] > 0 else 0)
            else:
                z0 = w0 = None
                if warm[ax] is not None:
                    z0 = np.concatenate([warm[ax][0][1:], warm[ax][0][-1:]])
                    w0 = np.concatenate([warm[ax][1][1:], [0.0]])
                nxt, z, w, _ = strict_axis_admm(p, state[ax], hi_w, lo_w, z0=z0, w0=w0)
                warm[ax] = (z, w)
                flags[i + 1, ax] |= (FLAG_ACTIVE_LO if z[0] == lo_w[0] else 0) | (FLAG_ACTIVE_HI if z[0] == hi_w[0] else 0)
            state[ax] = nxt
            hist[ax].append(nxt.copy())
        if force_tick is not None and i == force_tick:
            state[1] = state[1] - np.array([0.0, impulse, 0.0])
            hist[1][-1] = state[1].copy()
    x_hist, y_hist = np.array(hist[0]), np.array(hist[1])
    m = len(x_hist)
    zmp = np.stack([x_hist @ p.C, y_hist @ p.C], axis=1)
    below = z_min[:m] - zmp
    above = zmp - z_max[:m]
    flags[:m][below > viol_tol] |= FLAG_VIOL_LO
    flags[:m][above > viol_tol] |= FLAG_VIOL_HI
    margin = -np.maximum(below, above)
    return dict(x_hist=x_hist, y_hist=y_hist, flags=flags[:m], zmp=zmp,
                max_violation=float(max(0.0, np.max(np.maximum(below, above)))), margin=margin)


def reference_outputs(res) -> Tuple[np.ndarray, np.ndarray]:
    """Shape the rollout like the reference's return value: com (n,2), y_hist (n,3,1)."""
    com = np.stack([res["x_hist"][:, 0], res["y_hist"][:, 0]], axis=1)
    return com, res["y_hist"][:, :, None]
