#!/usr/bin/env python3
"""CPU emulation of the reduced-precision strict path (design probe, not product, not a test).

Runs the default.json strict walk (400 N push) with the ADMM contraction v = M r evaluated the way
the tensor-core path would: operands rounded to bf16 / fp16 / tf32, optionally split into hi + lo
parts (3 products: hi*hi + hi*lo + lo*hi), FP32 accumulation, FP32 epilogue, FP64 state update,
tick-relative coordinates.  Prints the trajectory error against the committed exact-oracle walk.
"""

import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from oracle import lipm_oracle as orc  # noqa: E402


def round_bits(x, keep):
    """Round float32 to `keep` explicit mantissa bits (RNE): bf16 = 7, tf32/fp16 = 10."""
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - keep
    bias = ((u >> drop) & 1) + (1 << (drop - 1)) - 1
    u = ((u + bias) >> drop) << drop
    return u.astype(np.uint32).view(np.float32)


def split(x, keep, parts):
    hi = round_bits(x, keep)
    if parts == 1:
        return [hi]
    lo = round_bits(np.float32(x) - hi, keep)
    return [hi, lo]


def walk(keep, parts, eps, rho=2.0, alpha=1.6, max_iter=500, relative=True):
    g = np.load(os.path.join(REPO, "tests", "golden", "oracle_strict_walk.npz"))
    p = orc.Problem.make(150)
    c = p._cache()
    N = p.N
    M = orc.admm_matrix(p, rho)
    Ms = split(M, keep, parts)
    gpx = c["GPx"]
    z_min, z_max = g["z_min"], g["z_max"]
    n = len(z_min)
    lo_all = np.vstack([z_min, np.tile(z_min[-1:], (N, 1))])
    hi_all = np.vstack([z_max, np.tile(z_max[-1:], (N, 1))])
    state = [np.zeros(3), np.zeros(3)]
    hist = [[state[0].copy()], [state[1].copy()]]
    tv = [None, None]
    iters = 0
    f32 = np.float32
    for i in range(n - 1):
        for ax in (0, 1):
            x = state[ax]
            off = x[0] if relative else 0.0
            lo = (lo_all[i + 1:i + 1 + N, ax] - off).astype(f32)
            hi = (hi_all[i + 1:i + 1 + N, ax] - off).astype(f32)
            xr = np.array([x[0] - off, x[1], x[2]])
            # with Px[:,0] = 1 the G Px x term of the offset cancels against the shift of the box
            q = (-(p.Q / 2) * (lo.astype(np.float64) + hi) - gpx @ xr + (0.0 if relative else 0.0)).astype(f32)
            if tv[ax] is None:
                t = (0.5 * (lo + hi)).astype(f32)
            else:
                prev, prev_off = tv[ax]
                t = np.concatenate([prev[1:], prev[-1:]]) + f32(prev_off - off)
                t[-1] = np.clip(t[-1], lo[-1], hi[-1])
            thr = f32(eps * max(1.0, float(np.max(np.abs(q)))) / max(1 + alpha, rho * alpha))
            zc = np.clip(t, lo, hi)
            r = (f32(rho) * (f32(2) * zc - t) - q).astype(f32)
            for k in range(max_iter):
                rs = split(r, keep, parts)
                v = (Ms[0] @ rs[0]).astype(f32)
                if parts == 2:
                    v = (v + Ms[0] @ rs[1] + Ms[1] @ rs[0]).astype(f32)
                zo = np.clip(t, lo, hi)
                d = (v - zo).astype(f32)
                t = (t + f32(alpha) * d).astype(f32)
                zn = np.clip(t, lo, hi)
                r = (f32(2 * rho) * zn - f32(rho) * t - q).astype(f32)
                iters += 1
                if np.max(np.abs(d)) <= thr:
                    break
            tv[ax] = (t, off)
            z0 = float(np.clip(t[0], lo[0], hi[0])) + off
            jerk = (z0 - p.Px[0] @ x) / p.Pu[0, 0]
            nxt = p.A @ x + p.B * jerk
            if ax == 1 and i == int(g["force_tick"]):
                nxt = nxt - np.array([0.0, float(g["impulse"]), 0.0])
            state[ax] = nxt
            hist[ax].append(nxt.copy())
    xh, yh = np.array(hist[0]), np.array(hist[1])
    zmp = np.stack([xh @ p.C, yh @ p.C], 1)
    viol = np.maximum(z_min - zmp, zmp - z_max).max()
    return (np.abs(xh[:, 0] - g["x_hist"][:, 0]).max(), np.abs(yh[:, 0] - g["y_hist"][:, 0]).max(),
            np.abs(zmp - np.stack([g["x_hist"] @ p.C, g["y_hist"] @ p.C], 1)).max(), viol, iters / (2 * (n - 1)))


if __name__ == "__main__":
    cases = [("bf16x3", 7, 2), ("fp16x3", 10, 2), ("tf32x1", 10, 1), ("bf16x1", 7, 1)]
    for name, keep, parts in cases:
        for eps in (1e-5, 1e-6):
            ex, ey, ez, viol, it = walk(keep, parts, eps)
            print(f"{name:8s} eps={eps:g}: |dx|={ex:.2e} |dy|={ey:.2e} |dzmp|={ez:.2e} max box violation={viol:.2e} "
                  f"iters/solve={it:.1f}", flush=True)
