import os, sys, numpy as np
REPO='/root/repo'
sys.path.insert(0,REPO); sys.path.insert(0,os.path.join(REPO,'model-predictive-control-for-bipedal-locomotion_b200')); sys.path.insert(0,os.path.join(REPO,'tools'))
from oracle import herdt_oracle as ho
from oracle import osqp_restatement as osq
from mpc_bipedal.config import MPCConfig
from herdt_probe import project_polygon
from scipy.spatial import ConvexHull
g = np.load(os.path.join(REPO, "tests", "golden", "ref_herdt_helpers.npz"))
w = np.load(os.path.join(REPO, "tests", "golden", "oracle_herdt_walk.npz"))
cfg = MPCConfig(); cfg.add_force=True
N = cfg.horizon
states = g["class_defaults/states"]
v_ref = np.stack([g["class_defaults/vx"], g["class_defaults/vy"]], 1)
s_ext = np.concatenate([states, np.repeat(states[-1:], N)])
v_ext = np.vstack([v_ref, np.repeat(v_ref[-1:], N, 0)])
Pzx, Pzu, Pvs, Pvu = mats = ho.prediction_matrices(N, cfg.dt, cfg.h, cfg.g)
T = np.linalg.inv(Pzu); V = Pvu @ T
W = cfg.alpha * T.T @ T + cfg.beta * V.T @ V
poly = {"left": (g["polytope/left/A"], g["polytope/left/b"]), "right": (g["polytope/right/A"], g["polytope/right/b"])}
verts_all = {"left": np.array(cfg.left_foot_polytope), "right": np.array(cfg.right_foot_polytope)}
i = int(sys.argv[1]) if len(sys.argv)>1 else 251
k = int(w["certified_ticks"]); print('certified', k, w['x_hist'].shape)
x, y = w["x_hist"][i], w["y_hist"][i]
fc = w["foot_hist"][i]
print('state x', x, 'y', y, 'fc', fc, 'cur', s_ext[i], 'window states', np.unique(s_ext[i+1:i+1+N], return_counts=True))
# which side? count commits
from mpc_bipedal import herdt as H
lay = H.timeline_layout(states, N)
side = "left" if lay.foot_side[i]==0 else "right"
cur, win = int(s_ext[i]), s_ext[i + 1: i + 1 + N]
qp = ho.build_qp(cfg, x, y, v_ext[i + 1: i + 1 + N], fc[0], fc[1], cur, win, side, poly[side], mats)
print('m', qp['m'], 'runs', qp['runs'], 'rows', qp['A'].shape)
u, lam, info = ho.solve_qp(qp)
print('oracle', info)
res = osq.solve(*osq.herdt_problem(qp)); print('osqp', res['status'], res['iters'], res['prim_res'], res['dual_res'])
# feasibility check via LP-ish: min violation using OSQP with zero cost
Z = np.zeros_like(qp['Q']); 
viol = (qp['A'] @ res['x'] - qp['b']).max(); print('osqp sol max viol', viol)
print('oracle sol viol', (qp['A']@u - qp['b']).max(), 'cost diff', 0.5*u@qp['Q']@u+qp['p']@u - (0.5*res['x']@qp['Q']@res['x']+qp['p']@res['x']))

# ---- ADMM in the kernel's coordinates for several rho
m, U, U_c = qp["m"], qp["U"], qp["U_c"][:, 0]
n = N + m
walking_rows = win != ho.STANDING
hullflag = (cur == ho.STANDING or not walking_rows.any()) and (~walking_rows).any()
absolute = (~walking_rows) & U.any(axis=1) if hullflag else np.zeros(N, bool)
S = np.hstack([np.eye(N), U * (~absolute)[:, None]])
R = np.hstack([np.eye(N), -U * absolute[:, None]])
P = S.T @ W @ S + cfg.gamma * R.T @ R
for k in np.where(~U.any(axis=0))[0]:
    P[N + k, N + k] += 1.0
hull = ConvexHull(verts_all[side]); verts = verts_all[side][hull.vertices]
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
    sol.append(dict(q=q, lo=lo, hi=hi))
print('|q|', [np.abs(s_['q']).max() for s_ in sol], 'cond P', np.linalg.cond(P))
def project(tx, ty):
    zx, zy = np.clip(tx, sol[0]["lo"], sol[0]["hi"]), np.clip(ty, sol[1]["lo"], sol[1]["hi"])
    if m > 0:
        pt = project_polygon(np.array([tx[N] - fc[0], ty[N] - fc[1]]), verts)
        zx[N], zy[N] = pt[0] + fc[0], pt[1] + fc[1]
    return zx, zy
for rho in (6.0, 60.0, 600.0, 6000.0, 60000.0):
    alpha = 1.8
    M = np.linalg.inv(P + rho * np.eye(n))
    t = [np.clip(np.concatenate([np.zeros(N), np.full(m, fc[ax])]), s_["lo"], s_["hi"]) for ax, s_ in enumerate(sol)]
    hist=[]
    for it in range(1, 20001):
        zo = project(t[0], t[1])
        v = [M @ (rho * (2 * zo[ax] - t[ax]) - sol[ax]["q"]) for ax in (0, 1)]
        d = max(np.abs(v[0] - zo[0]).max(), np.abs(v[1] - zo[1]).max())
        t = [t[ax] + alpha * (v[ax] - zo[ax]) for ax in (0, 1)]
        if it in (10,100,1000,5000,20000): hist.append((it,d))
        if d < 1e-12*max(1,np.abs(sol[0]['q']).max())/max(1+alpha,rho*alpha):
            break
    z = project(t[0], t[1])
    Jx = T @ (S @ z[0] + U_c * fc[0] - Pzx @ x); Jy = T @ (S @ z[1] + U_c * fc[1] - Pzx @ y)
    got = np.concatenate([Jx, z[0][N:], Jy, z[1][N:]])
    print('rho',rho,'iters',it,'d',d, hist, 'J0 err vs oracle', abs(got[0]-u[0]), abs(got[n]-u[n]), 'cost diff', 0.5*got@qp['Q']@got+qp['p']@got-(0.5*u@qp['Q']@u+qp['p']@u), 'viol', (qp['A']@got-qp['b']).max())
