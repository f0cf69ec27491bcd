"""CPU oracle for the Herdt free-footstep MPC (TEST INFRASTRUCTURE; parity unpinned).

Checker only: ``tests/`` and ``__graft_entry__.smoke()`` may import it; the product never does.

Restates ``/root/reference/src/mpc_bipedal/controllers/zmp_controller.py``:

* ``predict_herdt_joint``          -- ``:533-826``: prediction matrices ``:566-579``, support-phase run
  lengths and selection matrices ``:581-603``, cost blocks ``:614-653``, ZMP-around-foot rows ``:661-713``,
  standing convex-hull rows ``:715-771``, footstep polytope rows ``:773-790``, state update ``:811-812``;
* ``generate_com_trajectory_herdt`` -- ``:435-531``: reference padding, swing-foot interpolation, side
  toggle and footstep commit, push, state bookkeeping.

``find_nb_steps`` (``:203-433``) and ``_polytope_halfspace`` (``:828-865``) are NOT restated here: their
outputs are arguments (the tests pass the reference's own outputs from ``tests/golden/ref_herdt_helpers.npz``
or the host mirror's, which is pinned against that file).

**Parity unpinned.**  The reference solves this QP with cvxpy -> OSQP (``:792-794``, un-pinned, not
installed here; OSQP stops at eps 1e-5 and no golden vector exists).  The QP is strictly convex
(``alpha > 0``), so its minimiser is unique whenever it is feasible; ``solve_qp`` computes it with
an ADMM pass in well-conditioned coordinates, polishes on the identified active set and CERTIFIES the
result through the KKT conditions of the reference's own matrices (stationarity, primal and dual
feasibility, complementarity) -- a proof of optimality independent of how the point was found.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

STANDING, DOUBLE_SUPPORT, SINGLE_SUPPORT = 0, 1, 2      # codes of the reference's State enum members


def prediction_matrices(N: int, T: float, h: float, g: float):
    """Pzx, Pzu, Pvs, Pvu by the reference's literal loops -- zmp_controller.py:566-579."""
    Pzx, Pzu = np.zeros((N, 3)), np.zeros((N, N))
    Pvs, Pvu = np.zeros((N, 3)), np.zeros((N, N))
    for i in range(N):
        Pzx[i, 0] = 1.0
        Pzx[i, 1] = T * (i + 1)
        Pzx[i, 2] = (T ** 2) / 2.0 * (i + 1) ** 2 - h / g
        Pvs[i, 1] = 1.0
        Pvs[i, 2] = (i + 1) * T
        j = np.arange(i + 1)
        Pzu[i, : i + 1] = (T ** 3) / 6.0 * (1 + 3 * (i - j) + 3 * (i - j) ** 2) - T * h / g
        Pvu[i, : i + 1] = (T ** 2) / 2.0 * (2 * (i - j) + 1)
    return Pzx, Pzu, Pvs, Pvu


def support_phases(current_state: int, state_ref: Sequence[int]) -> List[int]:
    """Run lengths ``l`` of the support phases in the window -- zmp_controller.py:581-594 (sum = N + 1)."""
    s, runs, c = current_state, [], 1
    for state in state_ref:
        if state == s or (s == DOUBLE_SUPPORT and state == SINGLE_SUPPORT):
            c += 1
        else:
            runs.append(c)
            c = 1
        s = state
    runs.append(c)
    return runs


def selection_matrices(N: int, runs: Sequence[int]):
    """U (N, m), U_c (N, 1) -- zmp_controller.py:596-603 (slices past N are clipped, as numpy does)."""
    m = len(runs) - 1
    U, U_c = np.zeros((N, m)), np.zeros((N, 1))
    U_c[: runs[0], 0] = 1
    n_c = runs[0]
    for i, n_f in enumerate(runs[1:]):
        U[n_c: n_c + n_f, i] = 1
        n_c += n_f
    return U, U_c


def build_qp(cfg, x, y, v_ref, x_fc: float, y_fc: float, current_state: int, state_ref, foot_side: str,
             polytope: Tuple[np.ndarray, np.ndarray], mats=None):
    """The reference's joint QP: returns dict(Q, p, A, b, N, m) with ``min 1/2 u'Qu + p'u  s.t. A u <= b``,
    ``u = [Jx (N), Fx (m), Jy (N), Fy (m)]`` -- zmp_controller.py:562-790."""
    N = len(state_ref)
    T = cfg.dt
    Pzx, Pzu, Pvs, Pvu = mats if mats is not None else prediction_matrices(N, T, cfg.h, cfg.g)
    state_ref = np.asarray(state_ref)
    runs = support_phases(current_state, state_ref)
    U, U_c = selection_matrices(N, runs)
    m = U.shape[1]
    x = np.asarray(x, float).reshape(3, 1)
    y = np.asarray(y, float).reshape(3, 1)

    # cost, :605-653
    Qxx = cfg.alpha * np.eye(N) + cfg.beta * (Pvu.T @ Pvu) + cfg.gamma * (Pzu.T @ Pzu)
    Qxf = -cfg.gamma * (Pzu.T @ U)
    Qff = cfg.gamma * (U.T @ U)
    Qx = np.block([[Qxx, Qxf], [Qxf.T, Qff]])
    Qx = 0.5 * (Qx + Qx.T)
    n1 = N + m
    Q = np.zeros((2 * n1, 2 * n1))
    Q[:n1, :n1] = Qx
    Q[n1:, n1:] = Qx

    def linear(state, vcol, fc):
        e_v = Pvs @ state - v_ref[:, vcol:vcol + 1]
        e_z = Pzx @ state - U_c * fc
        return np.vstack([cfg.beta * (Pvu.T @ e_v) + cfg.gamma * (Pzu.T @ e_z), -cfg.gamma * (U.T @ e_z)]).flatten()

    p = np.concatenate([linear(x, 0, x_fc), linear(y, 1, y_fc)])

    rows_A, rows_b = [], []

    def embed(block, axis):
        full = np.zeros((block.shape[0], 2 * n1))
        full[:, axis * n1: axis * n1 + n1] = block
        return full

    # ZMP around the supporting foot, :661-713 (rows of STANDING preview samples are removed)
    keep = state_ref != STANDING
    A_d = np.hstack([Pzu, -U])
    n_kept = 0
    for axis, (state, fc, size) in enumerate(((x, x_fc, cfg.foot_length), (y, y_fc, cfg.foot_width))):
        nominal = Pzx @ state - U_c * fc
        upper = (0.5 * size - nominal).flatten()
        lower = (0.5 * size + nominal).flatten()
        block = np.vstack([A_d[keep], -A_d[keep]])
        n_kept = block.shape[0]
        if n_kept:
            rows_A.append(embed(block, axis))
            rows_b.append(np.concatenate([upper[keep], lower[keep]]))

    # standing rows: ZMP inside the hull of both feet, :715-771
    standing = np.where(~keep)[0]
    if (current_state == STANDING or n_kept == 0) and len(standing) > 0:
        if foot_side == "left":
            y_left, y_right = y_fc, y_fc - 2 * cfg.foot_spread
        else:
            y_right, y_left = y_fc, y_fc + 2 * cfg.foot_spread
        box = ((x_fc - 0.5 * cfg.foot_length, x_fc + 0.5 * cfg.foot_length),
               (min(y_left, y_right) - 0.5 * cfg.foot_width, max(y_left, y_right) + 0.5 * cfg.foot_width))
        for axis, state in enumerate((x, y)):
            nominal = (Pzx[standing] @ state).flatten()
            block = np.zeros((2 * len(standing), n1))
            block[: len(standing), :N] = Pzu[standing]
            block[len(standing):, :N] = -Pzu[standing]
            rows_A.append(embed(block, axis))
            rows_b.append(np.concatenate([box[axis][1] - nominal, -box[axis][0] + nominal]))

    # first free footstep inside the reachable polytope of the swing foot, :773-790
    if m > 0:
        A_poly, b_poly = polytope
        full = np.zeros((A_poly.shape[0], 2 * n1))
        full[:, N] = A_poly[:, 0]
        full[:, n1 + N] = A_poly[:, 1]
        rows_A.append(full)
        rows_b.append(b_poly + A_poly @ np.array([x_fc, y_fc]))

    # A support phase that starts exactly at the end of the window has an all-zero column in U (the run
    # lengths count the current tick, so sum(l) = N + 1 and the last run can fall off the window): that
    # footstep variable appears nowhere in the cost -- the reference's QP is only positive SEMI-definite
    # there and OSQP returns an arbitrary value for it.  It is never consumed unless it is the FIRST
    # footstep (which needs a phase longer than the horizon); the oracle pins it to zero.
    for k in np.where(~U.any(axis=0))[0]:
        Q[N + k, N + k] += 1.0
        Q[n1 + N + k, n1 + N + k] += 1.0

    A = np.vstack(rows_A) if rows_A else np.zeros((0, 2 * n1))
    b = np.concatenate(rows_b) if rows_b else np.zeros(0)
    return dict(Q=0.5 * (Q + Q.T), p=p, A=A, b=b, N=N, m=m, runs=runs, U=U, U_c=U_c, Pzx=Pzx, Pzu=Pzu)


def kkt_certificate(Q, p, A, b, u, lam, tol=1e-8):
    """Residuals of the KKT conditions of ``min 1/2 u'Qu + p'u s.t. Au <= b`` (all must be ~0)."""
    scale = max(1.0, np.abs(p).max())
    r = dict(stationarity=float(np.abs(Q @ u + p + A.T @ lam).max()) / scale if len(u) else 0.0,
             primal=float(max(0.0, (A @ u - b).max())) if len(b) else 0.0,
             dual=float(max(0.0, (-lam).max())) if len(b) else 0.0,
             complementarity=float(np.abs(lam * (A @ u - b)).max()) / scale if len(b) else 0.0)
    r["ok"] = all(v <= tol for v in r.values())
    return r


def solve_qp(qp, warm=None, rho=1.0, alpha=1.6, sigma=1e-6, eps=1e-7, max_iter=20000):
    """Unique minimiser of the QP + multipliers, certified.  Returns (u, lam, info)."""
    Q, p, A, b = qp["Q"], qp["p"], qp["A"], qp["b"]
    n, nc = len(p), len(b)
    if nc == 0:
        u = np.linalg.solve(Q, -p)
        return u, np.zeros(0), dict(iters=0, kkt=kkt_certificate(Q, p, A, b, u, np.zeros(0)))
    # change of variables u = L xi (per axis: J = Pzu^-1 zeta) -- Hessian and rows become O(1)-conditioned
    N, m = qp["N"], qp["m"]
    n1 = N + m
    Linv_block = np.eye(n1)
    Linv_block[:N, :N] = qp["Pzu"]
    L1 = np.linalg.inv(Linv_block)
    L = np.zeros((n, n))
    L[:n1, :n1] = L1
    L[n1:, n1:] = L1
    Qt, pt, At = L.T @ Q @ L, L.T @ p, A @ L
    K = np.linalg.inv(Qt + sigma * np.eye(n) + rho * (At.T @ At))
    xi = np.zeros(n) if warm is None else np.linalg.solve(L, warm[0])
    z = np.minimum(At @ xi, b)
    yv = np.zeros(nc) if warm is None else warm[1].copy()
    iters = 0
    for iters in range(1, max_iter + 1):
        xt = K @ (sigma * xi - pt + At.T @ (rho * z - yv))
        zt = At @ xt
        xi_new = alpha * xt + (1 - alpha) * xi
        zr = alpha * zt + (1 - alpha) * z
        z_new = np.minimum(zr + yv / rho, b)
        yv = yv + rho * (zr - z_new)
        r_prim = np.abs(At @ xi_new - z_new).max()
        r_dual = np.abs(Qt @ xi_new + pt + At.T @ yv).max()
        xi, z = xi_new, z_new
        if r_prim <= eps and r_dual <= eps * max(1.0, np.abs(pt).max()):
            break
    u = L @ xi
    # polish: equality-constrained solve on the active set read from the multipliers
    lam = np.maximum(yv, 0.0)
    active = lam > 1e-7 * max(1.0, lam.max())
    for _ in range(8):
        Aw = A[active]
        k = Aw.shape[0]
        kkt = np.block([[Q, Aw.T], [Aw, -1e-12 * np.eye(k)]])
        sol = np.linalg.solve(kkt, np.concatenate([-p, b[active]]))
        u_p, lam_p = sol[:n], np.zeros(nc)
        lam_p[active] = sol[n:]
        viol = A @ u_p - b
        drop = active & (lam_p < -1e-10)
        add = (~active) & (viol > 1e-11)
        if not drop.any() and not add.any():
            u, lam = u_p, lam_p
            break
        active = (active & ~drop) | add
    cert = kkt_certificate(Q, p, A, b, u, lam)
    return u, lam, dict(iters=iters, kkt=cert, active=int(active.sum()))


def predict_joint(cfg, A_model, B_model, x, y, v_ref, x_fc, y_fc, current_state, state_ref, foot_side, polytope,
                  mats=None, warm=None):
    """One tick -- zmp_controller.py:533-826.  Returns (x_next, y_next, fx1|None, fy1|None, u, lam, info)."""
    qp = build_qp(cfg, x, y, v_ref, x_fc, y_fc, current_state, state_ref, foot_side, polytope, mats)
    u, lam, info = solve_qp(qp, warm)
    N, m = qp["N"], qp["m"]
    B_vec = np.asarray(B_model, float).reshape(3)
    x_next = A_model @ np.asarray(x, float).reshape(3) + B_vec * u[0]
    y_next = A_model @ np.asarray(y, float).reshape(3) + B_vec * u[N + m]
    fx1 = u[N] if m > 0 else None
    fy1 = u[N + m + N] if m > 0 else None
    info["m"] = m
    return x_next, y_next, fx1, fy1, u, lam, info


def rollout_herdt(cfg, A_model, B_model, x_init, y_init, v_ref, state_ref, nb_steps, polytopes,
                  n_ticks: Optional[int] = None, stop_when_uncertified: bool = False):
    """Whole walk -- zmp_controller.py:435-531.  ``state_ref``: int codes (n,); ``nb_steps``: the reference's
    ``find_nb_steps`` output for the PADDED timeline; ``polytopes``: {"left": (A, b), "right": (A, b)}."""
    N = cfg.horizon
    n = len(v_ref)
    force_time = n // 2
    x, y = np.asarray(x_init, float).reshape(3).copy(), np.asarray(y_init, float).reshape(3).copy()
    x_hist, y_hist = [x.copy()], [y.copy()]
    x_fc, y_fc = 0.0, float(cfg.foot_spread)
    x_air, y_air = x_fc, y_fc
    foot_hist = [(x_fc, y_fc)]
    foot_side = "left"
    v_ext = np.vstack([v_ref, np.repeat(v_ref[-1:], N, axis=0)])
    s_ext = np.concatenate([np.asarray(state_ref), np.repeat(np.asarray(state_ref)[-1:], N)])
    current = int(s_ext[0])
    mats = prediction_matrices(N, cfg.dt, cfg.h, cfg.g)
    certs = []
    ticks = n - 1 if n_ticks is None else n_ticks
    for i in range(ticks):
        x_next, y_next, fx1, fy1, _, _, info = predict_joint(
            cfg, A_model, B_model, x, y, v_ext[i + 1: i + 1 + N], x_fc, y_fc, current, s_ext[i + 1: i + 1 + N],
            foot_side, polytopes[foot_side], mats)
        certs.append(info["kkt"])
        if stop_when_uncertified and not info["kkt"]["ok"]:
            break                                    # the caller only keeps the certified prefix
        x, y = x_next, y_next
        if fx1 is not None:
            x_air += (1 / nb_steps[i][0]) * (fx1 - x_air)
            y_air += (1 / nb_steps[i][0]) * (fy1 - y_air)
        if s_ext[i + 1] != current and current == SINGLE_SUPPORT:
            foot_side = "left" if foot_side == "right" else "right"
            if fx1 is not None:
                x_fc, y_fc = float(fx1), float(fy1)
            else:
                x_fc, y_fc = x_air, y_air
            x_air, y_air = x_fc, y_fc
        foot_hist.append((x_fc, y_fc))
        if cfg.add_force and i == force_time:
            y = y - np.array([0.0, cfg.dt * cfg.F_ext / cfg.m, 0.0])
        x_hist.append(x.copy())
        y_hist.append(y.copy())
        if s_ext[i + 1] != current:
            current = int(s_ext[i + 1])
    return dict(x_hist=np.array(x_hist), y_hist=np.array(y_hist), foot_hist=np.array(foot_hist), kkt=certs)
