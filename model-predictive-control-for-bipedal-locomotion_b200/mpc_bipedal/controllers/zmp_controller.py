"""ZMP preview controller: the reference's Python surface on top of the batched CUDA engine.

Mirrors ``src/mpc_bipedal/controllers/zmp_controller.py`` of the reference: same
class, method names, argument meaning, return shapes (NumPy float64) and error
behaviour.  The per-tick QP, the state update and the whole receding-horizon
loop run on the GPU (``mpc_bipedal.engine`` -> ``liblipm_mpc.so``); the
single-walk methods are the B=1 case of ``generate_com_trajectory_batched``.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from ..config import MPCConfig
from ..generators.cop_generator import State
from ..models.lipm_model import discrete_model


class ZMPController:
    """ZMP controller using model predictive control (Wieber 2006) for a walking biped."""

    def __init__(self, config: MPCConfig):
        self.config = config
        self.A, self.B, self.C = discrete_model(config.dt, config.h, config.g)   # reference :15-21
        self.external_force = True
        # ADMM settings of the strict mode (library defaults when left at 0 / -1)
        self.solver_options: Dict[str, float] = {}
        self.herdt_options: Dict[str, float] = {}     # rho / relax / eps / max_iter of the Herdt ADMM

    # ------------------------------------------------------------------ engine access
    def _plan(self, horizon: Optional[int] = None):
        from .. import engine   # imported lazily: needs the built library and a CUDA device
        c = self.config
        return engine.get_plan(c.horizon if horizon is None else horizon, c.h, c.g, c.dt, c.Q, c.R,
                               **self.solver_options)

    def _mode(self) -> int:
        return 1 if self.config.strict else 0

    # ------------------------------------------------------------------ dispatcher (reference :23-57)
    def generate_com_trajectory(self, x_init: np.ndarray, y_init: np.ndarray, z_max: np.ndarray = None,
                                z_min: np.ndarray = None, v_ref: np.ndarray = None,
                                state_ref: np.ndarray = None):
        method = self.config.method.lower()
        if method == "wieber":
            if z_max is None or z_min is None:
                raise ValueError("z_max and z_min are required for wieber method")
            return self.generate_com_trajectory_wieber(x_init, y_init, z_max, z_min)
        elif method == "herdt":
            if v_ref is None or state_ref is None:
                raise ValueError("v_ref and state_ref are required for herdt method")
            return self.generate_com_trajectory_herdt(x_init, y_init, v_ref, state_ref)
        else:
            raise ValueError(f"Unknown method: {method}. Must be 'wieber' or 'herdt'")

    # ------------------------------------------------------------------ batched entry (new, additive)
    def generate_com_trajectory_batched(self, x_init, y_init, z_max, z_min, *, F_ext=None, m=None,
                                        force_tick=None, n_steps=None, add_force: Optional[bool] = None,
                                        strict: Optional[bool] = None, to_numpy=True,
                                        fall_angle_deg: float = 30.0):
        """Walks of B independent scenarios in one device launch.

        x_init, y_init: (B,3).  z_max, z_min: (n,2) shared by all scenarios or (B,n,2).
        F_ext, m: (B,) or scalars (default: config values); the push ``dt*F_ext/m`` hits the y
        velocity after tick ``force_tick`` (default ``n_steps // 2`` as the reference, :90).
        n_steps: (B,) int, bound samples per scenario (default: all n).
        Returns dict(com (B,n,2), x_hist, y_hist (B,n,3), zmp (B,n,2), flags (B,n,2) uint8,
        max_violation (B,), lean_max (B,), fell (B,) bool, status (B,), iters (B,2)), every one of them written by the
        walk kernel itself (no host or library post-processing).  ``to_numpy=True``: NumPy views of page-locked
        staging buffers that are recycled every second call (fastest; copy what must live longer);
        ``to_numpy="copy"``: NumPy arrays the caller owns; ``to_numpy=False``: device tensors.

        ``fell`` is a definition of this build (the reference computes no outcome): the walk counts
        as a fall when the CoM gets further from the ZMP than ``h * tan(fall_angle_deg)``, i.e. the
        pendulum leans more than that angle -- in strict mode the ZMP cannot leave the sole, so an
        unrecoverable push shows up as the CoM running away from it.
        """
        import torch
        from .. import engine

        cfg = self.config
        plan = self._plan()
        x_init = np.asarray(x_init, dtype=np.float64).reshape(-1, 3) if not torch.is_tensor(x_init) else x_init
        y_init = np.asarray(y_init, dtype=np.float64).reshape(-1, 3) if not torch.is_tensor(y_init) else y_init
        B = x_init.shape[0]
        n = z_min.shape[-2]
        push = cfg.add_force if add_force is None else add_force
        impulse = tick = None
        if push:
            F = np.broadcast_to(np.asarray(cfg.F_ext if F_ext is None else F_ext, dtype=np.float64), (B,))
            mass = np.broadcast_to(np.asarray(cfg.m if m is None else m, dtype=np.float64), (B,))
            impulse = cfg.dt * F / mass                                      # reference :106
            if force_tick is None:
                steps = np.full((B,), n) if n_steps is None else np.asarray(n_steps)
                tick = (steps // 2).astype(np.int32)                         # reference :90
            else:
                tick = np.broadcast_to(np.asarray(force_tick, dtype=np.int32), (B,)).copy()
        mode = self._mode() if strict is None else (1 if strict else 0)
        # device output buffers are kept between calls of the same shape (cleared: ragged walks leave the
        # samples past their end untouched)
        cache = getattr(self, "_rollout_buffers", None) if to_numpy else None    # device results are the caller's
        if cache is not None:
            for name in ("x_hist", "y_hist", "flags", "com", "zmp"):
                if name in cache:
                    cache[name].zero_()
        out = plan.rollout(mode, x_init, y_init, z_min, z_max, n_steps=n_steps, impulse=impulse, force_tick=tick,
                           want_readout=True, fall_lean=cfg.h * float(np.tan(np.radians(fall_angle_deg))), out=cache)
        if to_numpy:
            self._rollout_buffers = out
        res = {
            "com": out["com"],                                                # reference :108
            "x_hist": out["x_hist"], "y_hist": out["y_hist"],
            "zmp": out["zmp"],                                                # reference run_mpc.py:294
            "flags": out["flags"],
            "iters": out["iters"],
            "outcome": out["outcome"],
        }
        if not to_numpy:
            res.update({k: v for k, v in engine.unpack_outcome(out["outcome"]).items() if k != "iters"})
            return res
        # page-locked staging buffers are recycled: the arrays of a call stay valid until the call after next
        if getattr(self, "_host_pool", None) is None:
            self._host_pool = engine.HostPool()
        res = engine.to_host(res, self._host_pool)
        res.update({k: v for k, v in engine.unpack_outcome_host(res.pop("outcome")).items() if k != "iters"})
        if mode == 1 and plan.precision == engine.PRECISION_FP64 and plan.solver == engine.SOLVER_ACTIVE_SET:
            # Safety net: an active-set iteration whose set keeps changing (never seen on the BASELINE sweeps) is
            # reported like an ADMM solve that hit max_iter; those scenarios run again on the ADMM kernel, which
            # converges for every feasible QP.
            bad = np.flatnonzero(np.asarray(res["status"]) != 0)
            if bad.size:
                admm = dict(self.solver_options, solver="admm")
                c = self.config
                other = engine.get_plan(c.horizon, c.h, c.g, c.dt, c.Q, c.R, **admm)
                pick = lambda a: None if a is None else np.asarray(a)[bad]
                zmin_b = z_min if np.ndim(z_min) == 2 else np.asarray(z_min)[bad]
                zmax_b = z_max if np.ndim(z_max) == 2 else np.asarray(z_max)[bad]
                redo = other.rollout(mode, np.asarray(x_init)[bad], np.asarray(y_init)[bad], zmin_b, zmax_b, n_steps=pick(n_steps),
                                     impulse=pick(impulse), force_tick=pick(tick), want_readout=True,
                                     fall_lean=cfg.h * float(np.tan(np.radians(fall_angle_deg))))
                res = {k: np.array(v) for k, v in res.items()}
                for name in ("com", "x_hist", "y_hist", "zmp", "flags", "iters"):
                    res[name][bad] = redo[name].cpu().numpy()
                for name, val in engine.unpack_outcome_host(redo["outcome"].cpu().numpy()).items():
                    if name != "iters":
                        res[name][bad] = val
                return res
        if to_numpy == "copy":
            res = {k: np.array(v) for k, v in res.items()}
        return res

    # ------------------------------------------------------------------ single walk (reference :59-147)
    def _single_walk(self, x_init, y_init, z_max, z_min, push: bool):
        z_max = np.ascontiguousarray(z_max, dtype=np.float64)
        z_min = np.ascontiguousarray(z_min, dtype=np.float64)
        res = self.generate_com_trajectory_batched(
            np.asarray(x_init, dtype=np.float64).reshape(1, 3), np.asarray(y_init, dtype=np.float64).reshape(1, 3),
            z_max, z_min, add_force=push)
        if self.config.strict and int(res["status"][0]) != 0:
            raise RuntimeError("QP solver did not find a solution (infeasible or other).")   # reference :193-194
        return res

    def generate_com_trajectory_wieber(self, x_init: np.ndarray, y_init: np.ndarray, z_max: np.ndarray,
                                       z_min: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """CoM trajectory of one walk: ``(com (n,2), y_hist (n,3,1))``."""
        n_steps = len(z_min)
        force_time = n_steps // 2
        print(f"Time of the external force: {(force_time * self.config.dt):.2f}s")
        res = self._single_walk(x_init, y_init, z_max, z_min, push=bool(self.config.add_force))
        return res["com"][0].copy(), res["y_hist"][0][:, :, None].copy()      # own memory, like the reference

    def generate_state_trajectory_wieber(self, x_init: np.ndarray, y_init: np.ndarray, z_max: np.ndarray,
                                         z_min: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Full state histories ``(x_hist, y_hist)``, each (n,3,1); never pushes (reference :110-147)."""
        res = self._single_walk(x_init, y_init, z_max, z_min, push=False)
        return res["x_hist"][0][:, :, None].copy(), res["y_hist"][0][:, :, None].copy()

    # ------------------------------------------------------------------ one tick (reference :149-201)
    def predict_wieber_axis(self, x_init: np.ndarray, nb_steps: int, z_max: np.ndarray, z_min: np.ndarray) -> np.ndarray:
        """Next axis state (3,1) from state (3,1) and the bound window (nb_steps,1)."""
        plan = self._plan(int(nb_steps))
        lo = np.ascontiguousarray(z_min, dtype=np.float64).reshape(1, nb_steps)
        hi = np.ascontiguousarray(z_max, dtype=np.float64).reshape(1, nb_steps)
        res = plan.solve_tick(self._mode(), np.asarray(x_init, dtype=np.float64).reshape(1, 3), lo, hi)
        if self.config.strict and int(res["status"][0].item()) != 0:
            raise RuntimeError("QP solver did not find a solution (infeasible or other).")
        return res["next_state"].cpu().numpy().reshape(3, 1)

    def predict_wieber_axis_batched(self, states, z_max, z_min, want_plan: bool = False):
        """B independent ticks at once: states (B,3), z_max/z_min (B,N) or (N,)."""
        plan = self._plan()
        return plan.solve_tick(self._mode(), states, z_min, z_max, want_plan=want_plan)

    # ------------------------------------------------------------------ Herdt helpers (host side)
    def find_nb_steps(self, state_ref) -> list:
        """Per sample ``(steps_to_next_footstep_change, samples_of_current_footstep)``.

        Same table as the reference's O(n^2) scan (`:203-433`), computed in O(n) from four
        index maps: next DOUBLE_SUPPORT after i, next SINGLE_SUPPORT after i, last DOUBLE_SUPPORT
        before i and the start of the DOUBLE_SUPPORT run that contains an index.
        """
        n = len(state_ref)
        is_ds = [s == State.DOUBLE_SUPPORT for s in state_ref]
        is_ss = [s == State.SINGLE_SUPPORT for s in state_ref]
        next_ds, next_ss = [None] * n, [None] * n
        seen_ds = seen_ss = None
        for i in range(n - 1, -1, -1):
            next_ds[i], next_ss[i] = seen_ds, seen_ss
            if is_ds[i]:
                seen_ds = i
            if is_ss[i]:
                seen_ss = i
        prev_ds, run_start = [None] * n, [None] * n
        last_ds = None
        for i in range(n):
            prev_ds[i] = last_ds
            if is_ds[i]:
                run_start[i] = run_start[i - 1] if i > 0 and is_ds[i - 1] else i
                last_ds = i

        to_next = [0] * n
        for i in range(n):
            remaining = n - i
            s = state_ref[i]
            if s == State.STANDING:
                ds = next_ds[i]
                ss = next_ss[ds] if ds is not None else None
                to_next[i] = remaining if ss is None else ss - i - 1
            elif s in (State.DOUBLE_SUPPORT, State.SINGLE_SUPPORT):
                to_next[i] = remaining if next_ds[i] is None else next_ds[i] - i
            else:
                to_next[i] = remaining

        result = [(0, 0)] * n
        for i in range(n):
            remaining = n - i
            s = state_ref[i]
            end = n if next_ds[i] is None else next_ds[i]
            if s == State.STANDING:
                if i == 0:
                    total = to_next[0]
                elif prev_ds[i] is not None:
                    total = to_next[prev_ds[i]]
                else:
                    total = result[0][1]
            elif s == State.DOUBLE_SUPPORT:
                total = end - run_start[i]
            elif s == State.SINGLE_SUPPORT:
                total = end - run_start[prev_ds[i]] if prev_ds[i] is not None else remaining
            else:
                total = remaining
            result[i] = (to_next[i], total)
        return result

    def _polytope_halfspace(self, vertices: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Convex polygon vertices -> ``A x <= b`` via Qhull facets (reference :828-865)."""
        from scipy.spatial import ConvexHull

        verts = np.asarray(vertices, dtype=float)
        if verts.ndim != 2 or verts.shape[1] != 2 or len(verts) < 3:
            raise ValueError("Polytope must be array-like of shape (k, 2), k>=3")
        eq = ConvexHull(verts).equations            # rows [a, b, c], a x + b y + c <= 0 inside
        A, b = eq[:, :2], -eq[:, 2]
        worst = np.max(verts @ A.T - b)
        if worst > 1e-10:
            raise ValueError(f"Polytope half-space conversion failed: max violation = {worst}")
        return A, b

    # ------------------------------------------------------------------ Herdt (reference :435-826)
    def generate_com_trajectory_herdt_batched(self, x_init, y_init, v_ref, state_ref, *, F_ext=None, m=None,
                                              foot_spread=None, add_force: Optional[bool] = None,
                                              to_numpy: bool = True):
        """Herdt walks of B scenarios that share ``state_ref`` (one launch per tick for the whole batch).

        x_init, y_init (B,3); v_ref (n,2) shared or (B,n,2); state_ref (n,) of ``State``.  Returns dict(com
        (B,n,2), x_hist, y_hist (B,n,3), foot_hist (B,n,2), status (B,), iters (B,))."""
        import torch
        from .. import engine

        batch = self.prepare_herdt(state_ref)
        out = batch.rollout(x_init, y_init, v_ref, F_ext=F_ext, mass=m, foot_spread=foot_spread, add_force=add_force)
        if not to_numpy:
            out["com"] = torch.stack([out["x_hist"][..., 0], out["y_hist"][..., 0]], dim=-1)
            return out
        host = engine.to_host(out)
        host["com"] = np.stack([host["x_hist"][..., 0], host["y_hist"][..., 0]], axis=-1)    # reference :526-528
        return host

    def prepare_herdt(self, state_ref):
        """Builds (once per timeline) what the Herdt walks of ``state_ref`` share: the support layout, the plan and the
        per-tick inverted Hessians on the device.  Called implicitly by the walk methods; call it up front to keep
        the build (which synchronises the device) out of a region where several timelines run on their own streams."""
        from ..herdt import HerdtBatch, state_codes

        codes = state_codes(state_ref)
        padded = np.concatenate([np.asarray(state_ref, dtype=object),
                                 np.repeat(np.asarray(state_ref, dtype=object)[-1:], self.config.horizon)])
        # every input the plan bakes in (HerdtBatch / HerdtPlan / timeline): a change of any of them rebuilds it
        c = self.config
        key = (codes.tobytes(), c.horizon, c.h, c.g, c.dt, c.alpha, c.beta, c.gamma, c.foot_length, c.foot_width,
               c.foot_spread, np.asarray(c.left_foot_polytope, float).tobytes(), np.asarray(c.right_foot_polytope, float).tobytes(),
               tuple(sorted(self.herdt_options.items())))
        cache = getattr(self, "_herdt_cache", None)
        if cache is None or cache[0] != key:
            if cache is not None:
                cache[1].plan.close()                      # the evicted plan's device memory goes now, not at GC time
            batch = HerdtBatch(self.config, state_ref, self.find_nb_steps(padded), solver_options=self.herdt_options)
            self._herdt_cache = (key, batch)
        return self._herdt_cache[1]

    def generate_com_trajectory_herdt(self, x_init, y_init, v_ref, state_ref):
        """One Herdt walk: ``(com (n,2), y_hist (n,3,1), foot_hist (n,2))`` -- reference :435-531."""
        res = self.generate_com_trajectory_herdt_batched(
            np.asarray(x_init, dtype=np.float64).reshape(1, 3), np.asarray(y_init, dtype=np.float64).reshape(1, 3),
            np.asarray(v_ref, dtype=np.float64), state_ref)
        if int(res["status"][0]) != 0:
            print("Joint QP solver failed on at least one tick")               # reference :797 prints and falls back
        return res["com"][0], res["y_hist"][0][:, :, None], res["foot_hist"][0]

    def predict_herdt_joint(self, x_init, y_init, v_ref, x_fc, y_fc, current_state, state_ref, nb_steps,
                            nb_steps_to_next_state, x_airc, y_airc, foot_side, idx=0):
        """One tick of the joint QP -- reference :533-826 (13 arguments, same order).  Returns
        ``(x_next (3,1), y_next (3,1), first_x_footstep | None, first_y_footstep | None)``."""
        import torch
        from .. import engine
        from ..herdt import STANDING, polygon_vertices, state_codes, support_runs

        cfg = self.config
        N = int(nb_steps)
        window = state_codes(state_ref)[:N]
        cur = int(state_codes([current_state])[0])
        runs = support_runs(cur, window)
        phase = np.full((1, N), -2, np.int8)
        phase[0, : runs[0]] = -1
        start = runs[0]
        for k, length in enumerate(runs[1:]):
            phase[0, start: start + length] = k
            start += length
        kind = np.zeros((1, N), np.int8)
        walking = window != STANDING
        kind[0][walking] = 1
        if (cur == STANDING or not walking.any()) and (~walking).any():
            kind[0][~walking] = 2
        plan = engine.HerdtPlan(N, cfg.h, cfg.g, cfg.dt, cfg.alpha, cfg.beta, cfg.gamma, cfg.foot_length,
                                cfg.foot_width, cfg.foot_spread, **self.herdt_options)
        timeline = plan.timeline(phase, kind, np.array([len(runs) - 1], np.int32))
        dev = plan.device
        to_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        v = np.asarray(v_ref, dtype=np.float64).reshape(-1, 2)[:N]
        vref = to_dev(np.vstack([v[:1], v]))                                      # the tick reads samples 1 .. N
        side = 0 if foot_side == "left" else 1
        poly = polygon_vertices(cfg.left_foot_polytope if side == 0 else cfg.right_foot_polytope)
        fc = to_dev([[float(np.asarray(x_fc).reshape(-1)[0]), float(np.asarray(y_fc).reshape(-1)[0])]])
        out = timeline.tick(0, to_dev(np.asarray(x_init).reshape(1, 3)), to_dev(np.asarray(y_init).reshape(1, 3)), fc,
                            vref, side, poly if len(runs) > 1 else None)
        res = engine.to_host({k: out[k] for k in ("next_x", "next_y", "f1", "status")})
        plan.close()
        has_step = len(runs) > 1
        return (res["next_x"].reshape(3, 1), res["next_y"].reshape(3, 1),
                float(res["f1"][0, 0]) if has_step else None, float(res["f1"][0, 1]) if has_step else None)
