#!/usr/bin/env python3
"""Design probe (CPU, not product, not a test): the Herdt joint QP in the coordinates the GPU kernel uses.

Per axis xi = (e, F):  e = z - U F - U_c fc  (previewed ZMP relative to its supporting foot),  F = future
footsteps.  The ZMP-around-foot rows become a plain box on e, the footstep polytope a 2-D projection of the
pair (F1x, F1y), the Hessian P = S'(alpha T'T + beta V'V)S + gamma diag(I,0) (S = [I U], T = Pzu^-1,
V = Pvu T) is well conditioned and IDENTICAL for both axes and for every scenario that shares the timeline.
Solves one tick by the fixed-rho ADMM of the strict Wieber kernel (t-form, projection = box / polygon) and
compares with the certified oracle.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    sys.path.insert(0, _p)
from oracle import herdt_oracle as ho  # noqa: E402
from mpc_bipedal.config import MPCConfig  # noqa: E402


def project_polygon(p, verts):
    """Euclidean projection of a 2-D point onto a convex polygon given by counter-clockwise vertices."""
    inside = True
    best, best_d = None, np.inf
    k = len(verts)
    for i in range(k):
        a, b = verts[i], verts[(i + 1) % k]
        ab, ap = b - a, p - a
        if ab[0] * ap[1] - ab[1] * ap[0] < 0:
            inside = False
        s = np.clip(ap @ ab / (ab @ ab), 0.0, 1.0)
        c = a + s * ab
        d = (p - c) @ (p - c)
        if d < best_d:
            best, best_d = c, d
    return p.copy() if inside else best


def main():
    from scipy.spatial import ConvexHull
    g = np.load(os.path.join(REPO, "tests", "golden", "ref_herdt_helpers.npz"))
    cfg = MPCConfig()
    N = cfg.horizon
    states = g["class_defaults/states"]
    v_ref = np.stack([g["class_defaults/vx"], g["class_defaults/vy"]], 1)
    s_ext = np.concatenate([states, np.repeat(states[-1:], N)])
    v_ext = np.vstack([v_ref, np.repeat(v_ref[-1:], N, 0)])
    Pzx, Pzu, Pvs, Pvu = mats = ho.prediction_matrices(N, cfg.dt, cfg.h, cfg.g)
    T = np.linalg.inv(Pzu)
    V = Pvu @ T
    W = cfg.alpha * T.T @ T + cfg.beta * V.T @ V
    poly = {"left": (g["polytope/left/A"], g["polytope/left/b"]), "right": (g["polytope/right/A"], g["polytope/right/b"])}
    verts_all = {"left": np.array(cfg.left_foot_polytope), "right": np.array(cfg.right_foot_polytope)}
    rng = np.random.default_rng(0)
    for i, side, fc in ((0, "left", (0.0, 0.1)), (60, "left", (0.0, 0.1)), (100, "right", (0.3, -0.1)), (251, "right", (1.9, -0.1)),
                        (260, "left", (1.9, 0.1))):
        x = rng.normal(size=3) * [0.02, 0.05, 0.1] + [fc[0], 0.2 if 50 < i < 250 else 0, 0]
        y = rng.normal(size=3) * [0.01, 0.03, 0.1] + [fc[1] * 0.3, 0, 0]
        cur, win = int(s_ext[i]), s_ext[i + 1: i + 1 + N]
        qp = ho.build_qp(cfg, x, y, v_ext[i + 1: i + 1 + N], fc[0], fc[1], cur, win, side, poly[side], mats)
        u, lam, info = ho.solve_qp(qp)
        assert info["kkt"]["ok"], info
        m, U, U_c = qp["m"], qp["U"], qp["U_c"][:, 0]
        n = N + m
        walking_rows = win != ho.STANDING
        hull = (cur == ho.STANDING or not walking_rows.any()) and (~walking_rows).any()
        absolute = (~walking_rows) & U.any(axis=1) if hull else np.zeros(N, bool)   # hull rows tied to a future footstep
        S = np.hstack([np.eye(N), U * (~absolute)[:, None]])
        R = np.hstack([np.eye(N), -U * absolute[:, None]])
        P = S.T @ W @ S + cfg.gamma * R.T @ R
        for k in np.where(~U.any(axis=0))[0]:
            P[N + k, N + k] += 1.0
        rho, alpha = float(os.environ.get("RHO", 2.0)), 1.6
        # Jacobi scaling xi = D xi' (same D for both axes, so the polygon is scaled uniformly)
        D = 1.0 / np.sqrt(np.diag(P)) if os.environ.get("SCALE") else np.ones(n)
        if os.environ.get("SCALE") == "F":
            D[:N] = 1.0
        P = D[:, None] * P * D[None, :]
        M = np.linalg.inv(P + rho * np.eye(n))
        hull = ConvexHull(verts_all[side])
        verts = verts_all[side][hull.vertices]                 # counter-clockwise
        sol = []
        for ax, (st, half) in enumerate(((x, 0.5 * cfg.foot_length), (y, 0.5 * cfg.foot_width))):
            d1 = U_c * fc[ax] - Pzx @ st
            d2 = Pvs @ st - v_ext[i + 1: i + 1 + N, ax]
            qz = W @ d1 + cfg.beta * V.T @ d2
            q = S.T @ qz
            lo, hi = np.full(n, -1e30), np.full(n, 1e30)
            walking = win != ho.STANDING
            lo[:N][walking], hi[:N][walking] = -half, half
            standing = np.where(~walking)[0]
            if (cur == ho.STANDING or not walking.any()) and len(standing):
                if ax == 0:
                    blo, bhi = fc[0] - 0.5 * cfg.foot_length, fc[0] + 0.5 * cfg.foot_length
                else:
                    other = fc[1] - 2 * cfg.foot_spread if side == "left" else fc[1] + 2 * cfg.foot_spread
                    blo, bhi = min(fc[1], other) - 0.5 * cfg.foot_width, max(fc[1], other) + 0.5 * cfg.foot_width
                lo[standing], hi[standing] = blo - U_c[standing] * fc[ax], bhi - U_c[standing] * fc[ax]
            sol.append(dict(q=D * q, lo=lo / D, hi=hi / D))
        # joint ADMM (the two axes only meet in the polygon projection of the pair (F1x, F1y))
        t = [np.clip(np.concatenate([np.zeros(N), np.full(m, fc[ax])]) / D, s_["lo"], s_["hi"]) for ax, s_ in enumerate(sol)]

        def project(tx, ty):
            zx, zy = np.clip(tx, sol[0]["lo"], sol[0]["hi"]), np.clip(ty, sol[1]["lo"], sol[1]["hi"])
            if m > 0:
                pt = project_polygon(np.array([tx[N] * D[N] - fc[0], ty[N] * D[N] - fc[1]]), verts)
                zx[N], zy[N] = (pt[0] + fc[0]) / D[N], (pt[1] + fc[1]) / D[N]
            return zx, zy

        for it in range(1, 5000):
            zo = project(t[0], t[1])
            v = [M @ (rho * (2 * zo[ax] - t[ax]) - sol[ax]["q"]) for ax in (0, 1)]
            d = max(np.abs(v[0] - zo[0]).max(), np.abs(v[1] - zo[1]).max())
            t = [t[ax] + alpha * (v[ax] - zo[ax]) for ax in (0, 1)]
            if d < 1e-11:
                break
        z = project(t[0], t[1])
        z = (z[0] * D, z[1] * D)
        # back to the reference's variables
        Jx = T @ (S @ z[0] + U_c * fc[0] - Pzx @ x)
        Jy = T @ (S @ z[1] + U_c * fc[1] - Pzx @ y)
        got = np.concatenate([Jx, z[0][N:], Jy, z[1][N:]])
        scale = np.abs(u).max()
        print(f"tick {i:3d} m={m} iters={it:4d}  max|u - oracle|/max|u| = {np.abs(got - u).max() / scale:.2e}  "
              f"J0 err = {abs(got[0] - u[0]):.2e} {abs(got[n] - u[n]):.2e}  F1 err = "
              + (f"{abs(got[N] - u[N]):.2e} {abs(got[n + N] - u[n + N]):.2e}" if m else "-") + f"  cond(P)={np.linalg.cond(P):.1f}")


if __name__ == "__main__":
    main()
