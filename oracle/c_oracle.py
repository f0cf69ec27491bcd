"""ctypes wrapper of oracle/lipm_oracle.c (TEST INFRASTRUCTURE: checker and CPU baseline only)."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build(native: bool = False, force: bool = False) -> str:
    name = "liblipm_oracle_native.so" if native else "liblipm_oracle.so"
    out = os.path.join(BUILD, name)
    src = os.path.join(HERE, "lipm_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(BUILD, exist_ok=True)
        arch = "-march=native" if native else "-march=x86-64-v3"
        subprocess.run(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", arch, "-o", out, src, "-lm"], check=True)
    return out


def _load(path):
    lib = ctypes.CDLL(path)
    vp, i32, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.lipm_oracle_plan_create.argtypes = [i32, dbl, dbl, dbl, dbl, dbl, dbl, dbl, dbl, i32, dbl]
    lib.lipm_oracle_plan_create.restype = vp
    lib.lipm_oracle_plan_free.argtypes = [vp]
    lib.lipm_oracle_plan_export.argtypes = [vp] * 6
    lib.lipm_oracle_rollout.argtypes = [vp, i32, i32, i32] + [vp] * 5 + [i32] + [vp] * 8 + [i32]
    lib.lipm_oracle_rollout.restype = i32
    lib.lipm_oracle_threads.restype = i32
    return lib


class COracle:
    def __init__(self, N, h=0.75, g=9.81, dt=None, Q=1.0, R=1e-6, rho=2.0, alpha=1.75, eps=1e-12, max_iter=2000,
                 viol_tol=1e-9, native=False):
        self.lib = _load(build(native=native))
        self.N = N
        dt = 1.5 / N if dt is None else dt
        self.plan = self.lib.lipm_oracle_plan_create(N, h, g, dt, Q, R, rho, alpha, eps, max_iter, viol_tol)
        if not self.plan:
            raise RuntimeError("lipm_oracle_plan_create failed")

    def __del__(self):
        try:
            self.lib.lipm_oracle_plan_free(self.plan)
        except Exception:
            pass

    def max_threads(self):
        return int(self.lib.lipm_oracle_threads())

    def export(self):
        N = self.N
        Px, Pu, M, Hinv, gain = np.empty((N, 3)), np.empty((N, N)), np.empty((N, N)), np.empty((N, N)), np.empty(N + 3)
        self.lib.lipm_oracle_plan_export(self.plan, *(a.ctypes.data for a in (Px, Pu, M, Hinv, gain)))
        return dict(Px=Px, Pu=Pu, M=M, Hinv=Hinv, gain=gain)

    def rollout(self, strict, x0, y0, z_min, z_max, n_steps=None, impulse=None, force_tick=None, threads=0):
        x0 = np.ascontiguousarray(x0, np.float64).reshape(-1, 3)
        y0 = np.ascontiguousarray(y0, np.float64).reshape(-1, 3)
        B = len(x0)
        z_min = np.ascontiguousarray(z_min, np.float64)
        z_max = np.ascontiguousarray(z_max, np.float64)
        bounds_batch = 1 if z_min.ndim == 2 else z_min.shape[0]
        n_max = z_min.shape[-2]
        opt = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
        n_steps, impulse, force_tick = opt(n_steps, np.int32), opt(impulse, np.float64), opt(force_tick, np.int32)
        x_hist, y_hist = np.zeros((B, n_max, 3)), np.zeros((B, n_max, 3))
        flags = np.zeros((B, n_max, 2), np.uint8)
        viol, status, iters = np.zeros((B, 2)), np.zeros((B, 2), np.int32), np.zeros((B, 2), np.int32)
        ptr = lambda a: None if a is None else a.ctypes.data
        mode = 2 if strict == "riccati" else (1 if strict else 0)     # "riccati": active-set + Riccati solver
        self.lib.lipm_oracle_rollout(self.plan, mode, B, n_max, ptr(n_steps), ptr(x0), ptr(y0),
                                     ptr(z_min), ptr(z_max), bounds_batch, ptr(impulse), ptr(force_tick),
                                     ptr(x_hist), ptr(y_hist), ptr(flags), ptr(viol), ptr(status), ptr(iters),
                                     int(threads))
        return dict(x_hist=x_hist, y_hist=y_hist, flags=flags, max_violation=viol, status=status, iters=iters)
