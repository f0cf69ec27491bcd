"""OSQP as the reference calls it, restated in NumPy (TEST INFRASTRUCTURE; parity unpinned).

Checker only: ``tests/``, ``tools/`` probes and ``__graft_entry__.smoke()`` may import it; the product never does.

The reference hands its strict QPs to ``cvxpy`` -> ``OSQP`` (``zmp_controller.py:181-191`` Wieber, ``:792-794``
Herdt).  Neither package is vendored, pinned (``requirements.txt:2``: ``cvxpy>=1.2.0``) nor installed here, so
what the reference ACTUALLY returns -- a point that satisfies OSQP's stopping test at cvxpy's default
``eps_abs = eps_rel = 1e-5``, or nothing at all when OSQP reports infeasibility -- cannot be observed.  This
module restates the published algorithm (Stellato, Banjac, Goulart, Bemporad, Boyd: "OSQP: an operator
splitting solver for quadratic programs", Math. Prog. Comp. 2020; solver release 0.6.x defaults) with dense
linear algebra, and the way cvxpy's OSQP interface canonicalises the two problems, in order to

* bound how far the reference's own output can sit from the exact minimiser the oracle computes
  (``wieber_strict_problem`` + ``solve``: same stopping test, same tolerances), and
* decide what the reference does on the ticks where the Herdt QP has no (numerically reachable) solution:
  OSQP "solved" / "maximum iterations reached" -> cvxpy fills ``u.value`` (USER_LIMIT counts as "solution
  present") and the reference applies it; "primal infeasible" -> ``u.value is None`` and the reference falls back
  to zero jerk and the interpolated footstep (``:796-804``).

What is NOT reproducible: OSQP picks the interval of its adaptive-rho updates from wall-clock timings
(``adaptive_rho_fraction`` of the setup time, rounded to a multiple of ``check_termination``), so the iterate
at which it stops depends on the machine.  ``adaptive_rho_interval`` is therefore a parameter here (default 50;
the callers sweep 25 / 50 / 100 to show the verdicts do not depend on it).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
from scipy.linalg import lu_factor, lu_solve

INF = 1e20          # OSQP_INFTY (bounds beyond it are treated as infinite)


@dataclass
class Settings:
    """Solver 0.6.x defaults, with the four values cvxpy's OSQP interface overrides (eps 1e-5, max_iter 10000)."""
    rho: float = 0.1
    sigma: float = 1e-6
    alpha: float = 1.6
    eps_abs: float = 1e-5
    eps_rel: float = 1e-5
    eps_prim_inf: float = 1e-4
    eps_dual_inf: float = 1e-4
    max_iter: int = 10000
    scaling: int = 10
    check_termination: int = 25
    adaptive_rho: bool = True
    adaptive_rho_interval: int = 50
    adaptive_rho_tolerance: float = 5.0
    polish: bool = False


def _limit_scaling(v: np.ndarray) -> np.ndarray:
    v = np.where(v < 1e-4, 1.0, v)
    return np.minimum(v, 1e4)


def _ruiz(P, q, A, iters):
    """Modified Ruiz equilibration of the KKT matrix with cost scaling (paper section 5.1)."""
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, q, A = P.copy(), q.copy(), A.copy()
    for _ in range(iters):
        col_top = np.maximum(np.abs(P).max(axis=0) if n else np.zeros(0), np.abs(A).max(axis=0) if m else np.zeros(n))
        d = 1.0 / np.sqrt(_limit_scaling(col_top))
        e = 1.0 / np.sqrt(_limit_scaling(np.abs(A).max(axis=1))) if m else np.zeros(0)
        P = d[:, None] * P * d[None, :]
        A = e[:, None] * A * d[None, :]
        q = d * q
        D, E = D * d, E * e
        cost = max(float(np.mean(np.abs(P).max(axis=0))), float(np.abs(q).max()))
        cost = 1.0 / float(_limit_scaling(np.array([cost]))[0])
        P, q, c = P * cost, q * cost, c * cost
    return P, q, A, D, E, c


def solve(P: np.ndarray, q: np.ndarray, A: np.ndarray, l: np.ndarray, u: np.ndarray,
          settings: Optional[Settings] = None) -> Dict[str, object]:
    """``min 1/2 x'Px + q'x  s.t.  l <= Ax <= u``.  Returns dict(status, x, y, iters, prim_res, dual_res, rho_updates);
    status in {"solved", "primal_infeasible", "dual_infeasible", "max_iter"}; x / y are the unscaled iterate at exit
    (OSQP returns NaN for infeasible problems; callers only use ``x`` when status is "solved" or "max_iter")."""
    st = settings or Settings()
    n, m = len(q), len(l)
    l = np.maximum(np.asarray(l, float), -INF)
    u = np.minimum(np.asarray(u, float), INF)
    Ps, qs, As, D, E, c = _ruiz(np.asarray(P, float), np.asarray(q, float), np.asarray(A, float), st.scaling)
    ls, us = E * l, E * u
    ls[l <= -INF], us[u >= INF] = -INF, INF
    eq = np.abs(u - l) < 1e-4                       # RHO_TOL
    loose = (l <= -INF) & (u >= INF)

    def rho_vector(rho):
        r = np.full(m, rho)
        r[eq] = 1e3 * rho
        r[loose] = 1e-6
        return r

    def factor(rho_vec):
        K = np.block([[Ps + st.sigma * np.eye(n), As.T], [As, -np.diag(1.0 / rho_vec)]])
        return lu_factor(K)

    rho = st.rho
    rho_vec = rho_vector(rho)
    lu = factor(rho_vec)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    status, updates = "max_iter", 0
    prim = dual = np.inf
    it = 0
    for it in range(1, st.max_iter + 1):
        rhs = np.concatenate([st.sigma * x - qs, z - y / rho_vec])
        sol = lu_solve(lu, rhs)
        xt, nu = sol[:n], sol[n:]
        zt = z + (nu - y) / rho_vec
        x_new = st.alpha * xt + (1 - st.alpha) * x
        zr = st.alpha * zt + (1 - st.alpha) * z
        z_new = np.minimum(np.maximum(zr + y / rho_vec, ls), us)
        y_new = y + rho_vec * (zr - z_new)
        dx, dy = x_new - x, y_new - y
        x, z, y = x_new, z_new, y_new
        check = (it % st.check_termination == 0) or it == st.max_iter
        adapt = st.adaptive_rho and st.adaptive_rho_interval > 0 and it % st.adaptive_rho_interval == 0
        if not (check or adapt):
            continue
        # residuals in the ORIGINAL problem's units (scaled_termination = 0)
        Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
        prim = float(np.abs((Ax - z) / E).max()) if m else 0.0
        dual = float(np.abs((Px + qs + Aty) / D).max()) / c
        prim_norm = max(float(np.abs(Ax / E).max()), float(np.abs(z / E).max())) if m else 0.0
        dual_norm = max(float(np.abs(Px / D).max()), float(np.abs(Aty / D).max()), float(np.abs(qs / D).max())) / c
        if check:
            if prim <= st.eps_abs + st.eps_rel * prim_norm and dual <= st.eps_abs + st.eps_rel * dual_norm:
                status = "solved"
                break
            # primal infeasibility certificate: delta_y with A' dy = 0 and u' dy+ + l' dy- < 0 (paper section 3.4)
            norm_dy = float(np.abs(E * dy).max()) if m else 0.0
            if norm_dy > 1e-30:
                dy_pos, dy_neg = np.maximum(dy, 0.0), np.minimum(dy, 0.0)
                finite_u, finite_l = us < INF, ls > -INF
                # an infinite bound paired with a multiplier step in its direction cannot certify anything
                blocked = bool((dy_pos[~finite_u] > 1e-4 * norm_dy).any() or (dy_neg[~finite_l] < -1e-4 * norm_dy).any())
                support = float(us[finite_u] @ dy_pos[finite_u] + ls[finite_l] @ dy_neg[finite_l])
                if not blocked and support < -st.eps_prim_inf * norm_dy:
                    if float(np.abs((As.T @ dy) / D).max()) < st.eps_prim_inf * norm_dy:
                        status = "primal_infeasible"
                        break
            norm_dx = float(np.abs(D * dx).max())
            if norm_dx > 1e-30 and float(qs @ dx) / c < -st.eps_dual_inf * norm_dx:
                if float(np.abs((Ps @ dx) / D).max()) / c < st.eps_dual_inf * norm_dx:
                    Adx = (As @ dx) / E
                    tol = st.eps_dual_inf * norm_dx
                    ok_u = np.where(us < INF, Adx < tol, True)
                    ok_l = np.where(ls > -INF, Adx > -tol, True)
                    if bool((ok_u & ok_l).all()):
                        status = "dual_infeasible"
                        break
        if adapt and it < st.max_iter:
            num = prim / max(prim_norm, 1e-10)
            den = dual / max(dual_norm, 1e-10)
            rho_new = float(np.clip(rho * np.sqrt(num / max(den, 1e-10)), 1e-6, 1e6))
            if rho_new > rho * st.adaptive_rho_tolerance or rho_new < rho / st.adaptive_rho_tolerance:
                rho = rho_new
                rho_vec = rho_vector(rho)
                lu = factor(rho_vec)
                updates += 1
    return {"status": status, "x": D * x, "y": (E * y) / c, "iters": it, "prim_res": prim, "dual_res": dual,
            "rho": rho, "rho_updates": updates}


# ------------------------------------------------------------------------------------------------ canonical forms
def wieber_strict_problem(p, x_state: np.ndarray, zmax: np.ndarray, zmin: np.ndarray):
    """The strict Wieber QP as cvxpy hands it to OSQP (``zmp_controller.py:174-190``).  ``sum_squares`` of an affine
    expression is canonicalised with an auxiliary variable ``t = Px x + Pu J - zref`` (N equality rows), the
    ``quad_form(J, R I)`` term stays on J: variables ``[J; t]``, ``P = blkdiag(R I, Q I)``, rows
    ``[Pu, -I] = zref - Px x`` (equalities) and ``[Pu; -Pu] J <= b_ineq`` (``:175-180,189``)."""
    N = p.N
    Pxx = p.Px @ np.asarray(x_state, float).reshape(3)
    zref = (zmax + zmin) / 2
    P = np.zeros((2 * N, 2 * N))
    P[:N, :N] = p.R * np.eye(N)
    P[N:, N:] = p.Q * np.eye(N)
    q = np.zeros(2 * N)
    A = np.zeros((3 * N, 2 * N))
    A[:N, :N], A[:N, N:] = p.Pu, -np.eye(N)
    A[N:2 * N, :N], A[2 * N:, :N] = p.Pu, -p.Pu
    rhs = zref - Pxx
    l = np.concatenate([rhs, np.full(2 * N, -INF)])
    u = np.concatenate([rhs, zmax - Pxx, -zmin + Pxx])
    return P, q, A, l, u


def wieber_strict_axis(p, x_state, zmax, zmin, settings: Optional[Settings] = None):
    """What the reference's strict ``predict_wieber_axis`` would return under this restatement:
    (next_state | None, info).  ``None`` <-> ``J.value is None`` <-> the reference raises (``:193-194``)."""
    P, q, A, l, u = wieber_strict_problem(p, x_state, zmax, zmin)
    res = solve(P, q, A, l, u, settings)
    if res["status"] not in ("solved", "max_iter"):
        return None, res
    J = res["x"][: p.N]
    return p.A @ np.asarray(x_state, float).reshape(3) + p.B * J[0], res


def herdt_problem(qp):
    """The joint Herdt QP of ``oracle.herdt_oracle.build_qp`` in OSQP form: ``quad_form(u, psd_wrap(Q)) / 2 + p'u``
    needs no auxiliary variable; every constraint is a one-sided ``A u <= b`` row (``:656-790``)."""
    return qp["Q"], qp["p"], qp["A"], np.full(len(qp["b"]), -INF), qp["b"]
