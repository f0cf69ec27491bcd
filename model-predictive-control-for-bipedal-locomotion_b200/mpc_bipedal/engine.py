"""ctypes binding of the CUDA library (C ABI: ``include/lipm_mpc.h``).

PyTorch is used for what it is good at here -- device memory, streams, NCCL --
and nothing else: every tensor is handed to the library as a raw device pointer
and the current stream handle.  There is no CPU path; importing this module on
a machine without the built library or calling it without a CUDA device raises.
"""

from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

MODE_UNCONSTRAINED = 0
MODE_STRICT = 1

PRECISION_FP64 = 0     # FP64 tensor pipe (DMMA): <= 1e-9 relative
PRECISION_F16X3 = 1    # tcgen05, FP16 hi/lo split operands, FP32 accumulate: <= 1e-4 m, horizon <= 637

FLAG_ACTIVE_LO = 1
FLAG_ACTIVE_HI = 2
FLAG_VIOL_LO = 4
FLAG_VIOL_HI = 8
FLAG_NOT_CONVERGED = 16

EXPORT_PX, EXPORT_PU, EXPORT_ABC, EXPORT_ADMM_M, EXPORT_HINV, EXPORT_GAIN, EXPORT_GPX = range(7)

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBRARY_PATH = os.path.join(_PKG_ROOT, "_lib", "liblipm_mpc.so")

EXPORTED_SYMBOLS = (
    "lipm_plan_create", "lipm_plan_destroy", "lipm_plan_export", "lipm_rollout", "lipm_rollout_indexed", "lipm_rollout_ex",
    "lipm_set_option", "lipm_solve_tick", "lipm_plan_stats", "lipm_last_error", "lipm_version", "lipm_max_strict_horizon",
    "lipm_max_reduced_horizon", "lipm_generate_bounds",
    "herdt_plan_create", "herdt_plan_destroy", "herdt_plan_export_w", "herdt_timeline_create",
    "herdt_timeline_destroy", "herdt_timeline_size", "herdt_tick", "herdt_rollout",
)


class LipmError(RuntimeError):
    """A call into liblipm_mpc failed; the message comes from ``lipm_last_error``."""

    def __init__(self, code: int, message: str):
        super().__init__(f"liblipm_mpc error {code}: {message}")
        self.code = code


class _PlanDesc(ctypes.Structure):
    _fields_ = [
        ("horizon", ctypes.c_int32),
        ("h", ctypes.c_double), ("g", ctypes.c_double), ("dt", ctypes.c_double),
        ("Q", ctypes.c_double), ("R", ctypes.c_double),
        ("rho", ctypes.c_double), ("alpha", ctypes.c_double), ("eps", ctypes.c_double),
        ("max_iter", ctypes.c_int32),
        ("viol_tol", ctypes.c_double),
        ("device", ctypes.c_int32),
        ("precision", ctypes.c_int32),
    ]


class _RolloutIO(ctypes.Structure):
    """``lipm_rollout_io_t`` of include/lipm_mpc.h."""
    _fields_ = [
        ("B", ctypes.c_int32), ("n_max", ctypes.c_int32),
        ("n_steps", ctypes.c_void_p), ("x0", ctypes.c_void_p), ("y0", ctypes.c_void_p),
        ("z_min", ctypes.c_void_p), ("z_max", ctypes.c_void_p),
        ("bounds_batch", ctypes.c_int32), ("bounds_index", ctypes.c_void_p),
        ("impulse", ctypes.c_void_p), ("force_tick", ctypes.c_void_p),
        ("x_hist", ctypes.c_void_p), ("y_hist", ctypes.c_void_p), ("com", ctypes.c_void_p), ("zmp", ctypes.c_void_p),
        ("flags", ctypes.c_void_p), ("max_violation", ctypes.c_void_p), ("lean_max", ctypes.c_void_p),
        ("status", ctypes.c_void_p), ("iters", ctypes.c_void_p), ("outcome", ctypes.c_void_p),
        ("fall_lean", ctypes.c_double),
    ]


OPT_KEEP_POOL_MEMORY = 1
OPT_TICK_RANGES = 2
OPT_SINGLE_CTA_F16 = 3
OPT_STRICT_SOLVER = 4
OPT_RICCATI_LANES = 5
OPT_MAX_SWEEPS = 6
OPT_RICCATI_SCRATCH = 7
OPT_FEW_PIN_ROWS = 8
OPT_ROW_PARALLEL = 9
OPT_WALK_CHUNKS = 10
# solver of the FP64 strict mode: both return the unique minimiser of the reference's QP
SOLVER_ACTIVE_SET = 0   # default: primal-dual active set + O(N) Riccati sweeps (csrc/lipm_rollout_riccati.cu)
SOLVER_ADMM = 1         # ADMM on the condensed QP, FP64 tensor pipe (csrc/lipm_rollout_strict.cu)
_SOLVERS = {"active_set": SOLVER_ACTIVE_SET, "riccati": SOLVER_ACTIVE_SET, "admm": SOLVER_ADMM,
            SOLVER_ACTIVE_SET: SOLVER_ACTIVE_SET, SOLVER_ADMM: SOLVER_ADMM}
OUTCOME_WORDS = 4      # lipm_outcome_t = 4 x 8 bytes: max_violation f64, lean_max f64, (status i32, fell i32), iters i64

_lib = None


def load_library() -> ctypes.CDLL:
    """Load the in-tree shared library; build it first with ``__graft_entry__.build()``."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBRARY_PATH):
        raise ImportError(
            f"{LIBRARY_PATH} not found: the CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIBRARY_PATH)
    vp, i32, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.lipm_plan_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_PlanDesc)]
    lib.lipm_plan_create.restype = i32
    lib.lipm_plan_destroy.argtypes = [vp]
    lib.lipm_plan_destroy.restype = None
    lib.lipm_plan_export.argtypes = [vp, i32, vp, ctypes.c_size_t]
    lib.lipm_plan_export.restype = i32
    lib.lipm_rollout.argtypes = [vp, i32, i32, i32] + [vp] * 5 + [i32] + [vp] * 9
    lib.lipm_rollout.restype = i32
    lib.lipm_rollout_indexed.argtypes = [vp, i32, i32, i32] + [vp] * 5 + [i32] + [vp] * 10
    lib.lipm_rollout_indexed.restype = i32
    lib.lipm_rollout_ex.argtypes = [vp, i32, ctypes.POINTER(_RolloutIO), vp]
    lib.lipm_rollout_ex.restype = i32
    lib.lipm_set_option.argtypes = [vp, i32, ctypes.c_longlong]
    lib.lipm_set_option.restype = i32
    lib.lipm_solve_tick.argtypes = [vp, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.lipm_solve_tick.restype = i32
    lib.lipm_plan_stats.argtypes = [vp, vp]
    lib.lipm_plan_stats.restype = i32
    lib.lipm_last_error.restype = ctypes.c_char_p
    lib.lipm_version.restype = ctypes.c_char_p
    lib.lipm_max_strict_horizon.restype = i32
    lib.lipm_max_reduced_horizon.restype = i32
    lib.lipm_generate_bounds.argtypes = [i32, vp, vp, vp, dbl, dbl, dbl, dbl, i32, vp, vp, vp, vp, vp]
    lib.lipm_generate_bounds.restype = i32
    _lib = lib
    return lib


def _check(code: int) -> None:
    if code != 0:
        raise LipmError(code, load_library().lipm_last_error().decode("utf-8", "replace"))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dev_tensor(value, dtype, device, shape=None) -> torch.Tensor:
    t = torch.as_tensor(value, dtype=dtype)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


@dataclass(frozen=True)
class PlanKey:
    horizon: int
    h: float
    g: float
    dt: float
    Q: float
    R: float
    rho: float = 0.0
    alpha: float = 0.0
    eps: float = 0.0
    max_iter: int = 0
    viol_tol: float = -1.0
    device: int = 0
    precision: int = 0
    solver: int = 0


class LipmPlan:
    """Device-resident matrices for one ``(N, h, g, dt, Q, R)`` -- see ``lipm_plan_create``."""

    def __init__(self, horizon: int, h: float, g: float, dt: Optional[float], Q: float, R: float, *,
                 rho: float = 0.0, alpha: float = 0.0, eps: float = 0.0, max_iter: int = 0,
                 viol_tol: float = -1.0, device: Optional[int] = None, precision: int = PRECISION_FP64,
                 solver=SOLVER_ACTIVE_SET):
        if solver not in _SOLVERS:
            raise ValueError(f"unknown solver {solver!r}: 'active_set' or 'admm'")
        if not torch.cuda.is_available():
            raise RuntimeError("liblipm_mpc needs a CUDA device (B200 / sm_100a); there is no CPU path")
        self._lib = load_library()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.init()
        with torch.cuda.device(self.device_index):
            torch.zeros(1, device=self.device)     # make sure the primary context exists
        self.horizon = int(horizon)
        self.dt = 1.5 / horizon if dt is None else float(dt)
        self.precision = int(precision)
        desc = _PlanDesc(self.horizon, h, g, self.dt, Q, R, rho, alpha, eps, max_iter, viol_tol, self.device_index,
                         self.precision)
        handle = ctypes.c_void_p()
        _check(self._lib.lipm_plan_create(ctypes.byref(handle), ctypes.byref(desc)))
        self._handle = handle
        self.h, self.g, self.Q, self.R = h, g, Q, R
        self.solver = _SOLVERS[solver]
        if self.solver != SOLVER_ACTIVE_SET:
            _check(self._lib.lipm_set_option(self._handle, OPT_STRICT_SOLVER, self.solver))

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self._lib.lipm_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ read-back
    def export(self, which: int) -> np.ndarray:
        N = self.horizon
        shape = {EXPORT_PX: (N, 3), EXPORT_PU: (N, N), EXPORT_ABC: (15,), EXPORT_ADMM_M: (N, N),
                 EXPORT_HINV: (N, N), EXPORT_GAIN: (N + 3,), EXPORT_GPX: (N, 3)}[which]
        out = np.empty(shape, dtype=np.float64)
        _check(self._lib.lipm_plan_export(self._handle, which, out.ctypes.data, out.size))
        return out

    def stats(self) -> Dict[str, int]:
        """Work accounting of the last strict launch (synchronises)."""
        out = np.zeros(4, dtype=np.uint64)
        _check(self._lib.lipm_plan_stats(self._handle, out.ctypes.data))
        # ADMM kernels: [0] iterations of the solver groups, [1] units.  Active-set solver: [0] sweep pairs, [1] columns,
        # [2] columns the row-parallel kernel parked on the Riccati kernel (Riccati kernel alone: stages of the general
        # recursion), [3] columns handed back (alone: table stages).  The old key names stay for the tools.
        return {"group_iterations": int(out[0]), "units": int(out[1]),
                "debug_cycles_contraction": int(out[2]), "debug_cycles_iteration": int(out[3]),
                "sweep_pairs": int(out[0]), "columns": int(out[1]), "columns_parked": int(out[2]), "columns_handed_back": int(out[3])}

    # ------------------------------------------------------------------ walks
    def rollout(self, mode: int, x0, y0, z_min, z_max, *, n_steps=None, impulse=None, force_tick=None,
                bounds_index=None, want_flags: bool = True, want_iters: bool = True, want_hist: bool = True,
                want_readout: bool = False, want_outcome: bool = True, fall_lean: float = 0.0,
                out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """Batched walks (``lipm_rollout``).  Inputs may be host or device arrays.

        x0, y0: (B,3); z_min, z_max: (n,2) shared or (B,n,2); n_steps: (B,) int32 or None;
        impulse: (B,) = dt*F_ext/m or None; force_tick: (B,) int32 or None.
        bounds_index: (B,) int32 -> z_min, z_max are a table (V,n,2) of series and scenario b uses
        series bounds_index[b] (``lipm_rollout_indexed``).
        Returns device tensors x_hist, y_hist (B,n,3), flags (B,n,2) uint8,
        max_violation (B,2), lean_max (B,2), status (B,2) int32, iters (B,2) int32; with ``want_readout`` also com and
        zmp (B,n,2), written by the walk kernel itself; with ``want_outcome`` the packed per-scenario record
        ``outcome`` (B,4) int64 = ``lipm_outcome_t`` (decode with ``unpack_outcome``).
        """
        dev = self.device
        x0 = _dev_tensor(x0, torch.float64, dev)
        y0 = _dev_tensor(y0, torch.float64, dev)
        if x0.ndim != 2 or x0.shape[1] != 3 or x0.shape != y0.shape:
            raise ValueError("x0 and y0 must both have shape (B, 3)")
        B = x0.shape[0]
        z_min = _dev_tensor(z_min, torch.float64, dev)
        z_max = _dev_tensor(z_max, torch.float64, dev)
        if z_min.shape != z_max.shape or z_min.shape[-1] != 2 or z_min.ndim not in (2, 3):
            raise ValueError("z_min and z_max must both have shape (n, 2) or (B, n, 2)")
        bounds_batch = 1 if z_min.ndim == 2 else z_min.shape[0]
        index_t = None
        if bounds_index is not None:
            if z_min.ndim != 3:
                raise ValueError("bounds_index needs a (V, n, 2) table of bound series")
            index_t = _dev_tensor(bounds_index, torch.int32, dev, (B,))
        elif bounds_batch not in (1, B):
            raise ValueError("bounds batch must be 1 or B")
        n_max = z_min.shape[-2]
        n_steps_t = None if n_steps is None else _dev_tensor(n_steps, torch.int32, dev, (B,))
        impulse_t = None if impulse is None else _dev_tensor(impulse, torch.float64, dev, (B,))
        force_t = None if force_tick is None else _dev_tensor(force_tick, torch.int32, dev, (B,))
        if out is None:
            out = {}
        def buf(name, shape, dtype):
            t = out.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
                t = torch.zeros(shape, dtype=dtype, device=dev)
                out[name] = t
            return t
        x_hist = buf("x_hist", (B, n_max, 3), torch.float64) if want_hist else None
        y_hist = buf("y_hist", (B, n_max, 3), torch.float64) if want_hist else None
        flags = buf("flags", (B, n_max, 2), torch.uint8) if want_flags else None
        viol = buf("max_violation", (B, 2), torch.float64)
        status = buf("status", (B, 2), torch.int32)
        iters = buf("iters", (B, 2), torch.int32) if want_iters else None
        lean = buf("lean_max", (B, 2), torch.float64)
        com = buf("com", (B, n_max, 2), torch.float64) if want_readout else None
        zmp = buf("zmp", (B, n_max, 2), torch.float64) if want_readout else None
        outcome = buf("outcome", (B, OUTCOME_WORDS), torch.int64) if want_outcome else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        io = _RolloutIO(B, n_max, _ptr(n_steps_t), _ptr(x0), _ptr(y0), _ptr(z_min), _ptr(z_max), bounds_batch,
                        _ptr(index_t), _ptr(impulse_t), _ptr(force_t), _ptr(x_hist), _ptr(y_hist), _ptr(com), _ptr(zmp),
                        _ptr(flags), _ptr(viol), _ptr(lean), _ptr(status), _ptr(iters), _ptr(outcome), float(fall_lean))
        _check(self._lib.lipm_rollout_ex(self._handle, mode, ctypes.byref(io), stream))
        # keep inputs alive until the stream has consumed them
        out["_keepalive"] = (x0, y0, z_min, z_max, n_steps_t, impulse_t, force_t, index_t)
        return out

    def set_tick_ranges(self, ranges: int) -> None:
        """Force the number of tick ranges of the FP64 strict kernel (0 = automatic); results do not depend on it."""
        _check(self._lib.lipm_set_option(self._handle, OPT_TICK_RANGES, int(ranges)))

    def set_riccati_lanes(self, lanes: int) -> None:
        """Columns per warp of the active-set kernel (0 = automatic); results do not depend on it."""
        _check(self._lib.lipm_set_option(self._handle, OPT_RICCATI_LANES, int(lanes)))

    def set_riccati_scratch(self, where: int) -> None:
        """Feed-forward terms of the active-set kernel's table stages: 0 automatic, 1 shared memory, 2 global scratch."""
        _check(self._lib.lipm_set_option(self._handle, OPT_RICCATI_SCRATCH, int(where)))

    def set_few_pin_rows(self, rows: int) -> None:
        """Active-set kernel: pinned rows up to which a sweep uses table solve + dense correction (0 = never, the default; up to 16)."""
        _check(self._lib.lipm_set_option(self._handle, OPT_FEW_PIN_ROWS, int(rows)))

    def set_row_parallel(self, on) -> None:
        """FP64 active-set walks on shared / indexed bound series start on the row-parallel kernel: True (default) for
        batches of up to 16384 columns, False never, 2 whatever the batch size."""
        _check(self._lib.lipm_set_option(self._handle, OPT_ROW_PARALLEL, 2 if on == 2 else (1 if on else 0)))

    def set_walk_chunks(self, chunks: int) -> None:
        """``LIPM_OPT_WALK_CHUNKS``: 0 automatic, 1 no side streams, 2..4 chunks of the batch overlapped on side streams."""
        _check(self._lib.lipm_set_option(self._handle, OPT_WALK_CHUNKS, int(chunks)))

    def set_max_sweeps(self, sweeps: int) -> None:
        """Active-set corrections per QP before the solve is reported as not converged (0 = default 64)."""
        _check(self._lib.lipm_set_option(self._handle, OPT_MAX_SWEEPS, int(sweeps)))

    def set_single_cta_f16(self, on: bool) -> None:
        """Reduced-precision walks (N <= 152) on the single-CTA tcgen05 kernel instead of the CTA-pair one."""
        _check(self._lib.lipm_set_option(self._handle, OPT_SINGLE_CTA_F16, 1 if on else 0))

    # ------------------------------------------------------------------ single ticks
    def solve_tick(self, mode: int, state, z_min, z_max, *, want_plan: bool = False) -> Dict[str, torch.Tensor]:
        """One control tick for B independent axis problems (``lipm_solve_tick``).

        state: (B,3); z_min, z_max: (N,) shared or (B,N).  Returns next_state (B,3), status (B,),
        iters (B,) and, with ``want_plan``, zmp_plan and jerk_plan (B,N).
        """
        dev = self.device
        N = self.horizon
        state = _dev_tensor(state, torch.float64, dev)
        if state.ndim != 2 or state.shape[1] != 3:
            raise ValueError("state must have shape (B, 3)")
        B = state.shape[0]
        z_min = _dev_tensor(z_min, torch.float64, dev)
        z_max = _dev_tensor(z_max, torch.float64, dev)
        if z_min.shape != z_max.shape or z_min.shape[-1] != N or z_min.ndim not in (1, 2):
            raise ValueError(f"z_min and z_max must both have shape ({N},) or (B, {N})")
        bounds_batch = 1 if z_min.ndim == 1 else z_min.shape[0]
        if bounds_batch not in (1, B):
            raise ValueError("bounds batch must be 1 or B")
        nxt = torch.empty((B, 3), dtype=torch.float64, device=dev)
        status = torch.zeros((B,), dtype=torch.int32, device=dev)
        iters = torch.zeros((B,), dtype=torch.int32, device=dev)
        zmp_plan = torch.empty((B, N), dtype=torch.float64, device=dev) if want_plan else None
        jerk_plan = torch.empty((B, N), dtype=torch.float64, device=dev) if want_plan else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(self._lib.lipm_solve_tick(
            self._handle, mode, B, _ptr(state), _ptr(z_min), _ptr(z_max), bounds_batch, _ptr(nxt),
            _ptr(zmp_plan), _ptr(jerk_plan), _ptr(status), _ptr(iters), stream))
        res = {"next_state": nxt, "status": status, "iters": iters, "_keepalive": (state, z_min, z_max)}
        if want_plan:
            res["zmp_plan"], res["jerk_plan"] = zmp_plan, jerk_plan
        return res


OUTCOME_DTYPE = np.dtype([("max_violation", "<f8"), ("lean_max", "<f8"), ("status", "<i4"), ("fell", "<i4"), ("iters", "<i8")])


def unpack_outcome_host(record: np.ndarray) -> Dict[str, np.ndarray]:
    """(B,4) int64 ``lipm_outcome_t`` records on the host -> named columns (views, no arithmetic)."""
    rec = np.ascontiguousarray(record).view(OUTCOME_DTYPE).reshape(-1)
    return {"max_violation": rec["max_violation"], "lean_max": rec["lean_max"], "status": rec["status"],
            "fell": rec["fell"] != 0, "iters": rec["iters"]}


def unpack_outcome(record: torch.Tensor) -> Dict[str, torch.Tensor]:
    """(B,4) int64 ``lipm_outcome_t`` records (device or host tensor) -> named columns."""
    from .parallel import unpack_outcomes
    out = unpack_outcomes(record)
    out["status"] = out["status"].to(torch.int32)
    return out


def keep_pool_memory(device: Optional[int] = None) -> None:
    """Opt in to ``LIPM_OPT_KEEP_POOL_MEMORY`` for a device: the rollouts' stream-ordered scratch buffers stay cached
    between calls instead of being returned to the driver at every synchronisation (process-wide effect on that
    device's default memory pool, which is why the library does not do it on its own)."""
    dev = torch.cuda.current_device() if device is None else int(device)
    _check(load_library().lipm_set_option(None, OPT_KEEP_POOL_MEMORY, dev))


_plan_cache: Dict[PlanKey, LipmPlan] = {}


def get_plan(horizon: int, h: float, g: float, dt: float, Q: float, R: float, *, device: Optional[int] = None,
             **solver) -> LipmPlan:
    """Plans are immutable and cheap to keep: cache one per parameter set and device."""
    dev = torch.cuda.current_device() if device is None else int(device)
    if "solver" in solver:
        if solver["solver"] not in _SOLVERS:
            raise ValueError(f"unknown solver {solver['solver']!r}: 'active_set' or 'admm'")
        solver = dict(solver, solver=_SOLVERS[solver["solver"]])
    key = PlanKey(int(horizon), float(h), float(g), float(dt), float(Q), float(R), device=dev, **solver)
    plan = _plan_cache.get(key)
    if plan is None:
        plan = LipmPlan(horizon, h, g, dt, Q, R, device=dev, **solver)
        _plan_cache[key] = plan
    return plan


def generate_bounds(config, distance, step_length, foot_spread, n_max: int, *, device: Optional[int] = None,
                    want_states: bool = True) -> Dict[str, torch.Tensor]:
    """Bound series of V walk variants generated on the device (``lipm_generate_bounds``): the batched, bit-exact
    counterpart of ``CoPGenerator(config).generate_cop_trajectory()`` for per-variant (distance, step_length,
    foot_spread); durations and dt come from ``config``.  Returns z_max, z_min (V, n_max, 2), states (V, n_max)
    int8, n (V,) int32 -- ready for ``LipmPlan.rollout(..., bounds_index=..., n_steps=n[bounds_index])``."""
    if not torch.cuda.is_available():
        raise RuntimeError("liblipm_mpc needs a CUDA device (B200 / sm_100a); there is no CPU path")
    lib = load_library()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    d = _dev_tensor(np.asarray(distance, dtype=np.float64).reshape(-1), torch.float64, dev)
    V = d.shape[0]
    s = _dev_tensor(np.broadcast_to(np.asarray(step_length, dtype=np.float64), (V,)).copy(), torch.float64, dev)
    f = _dev_tensor(np.broadcast_to(np.asarray(foot_spread, dtype=np.float64), (V,)).copy(), torch.float64, dev)
    z_min = torch.zeros((V, n_max, 2), dtype=torch.float64, device=dev)
    z_max = torch.zeros((V, n_max, 2), dtype=torch.float64, device=dev)
    states = torch.zeros((V, n_max), dtype=torch.int8, device=dev) if want_states else None
    n = torch.zeros((V,), dtype=torch.int32, device=dev)
    _check(lib.lipm_generate_bounds(V, d.data_ptr(), s.data_ptr(), f.data_ptr(), float(config.ssp_duration),
                                    float(config.dsp_duration), float(config.standing_duration), float(config.dt),
                                    int(n_max), z_min.data_ptr(), z_max.data_ptr(), _ptr(states), n.data_ptr(),
                                    torch.cuda.current_stream(dev).cuda_stream))
    return {"z_max": z_max, "z_min": z_min, "states": states, "n": n}


class HostPool:
    """Page-locked staging buffers for results, reused across calls (allocating 141 MB of pinned memory per
    call costs more than the walk kernel itself and made the end-to-end time erratic).  Two generations
    alternate, so the arrays handed out by a call stay valid until the call after the next one."""

    def __init__(self, generations: int = 2):
        self._slots = [dict() for _ in range(generations)]
        self._turn = 0

    def next_generation(self) -> dict:
        slot = self._slots[self._turn]
        self._turn = (self._turn + 1) % len(self._slots)
        return slot


def to_host(tensors: Dict[str, torch.Tensor], pool: Optional[HostPool] = None) -> Dict[str, np.ndarray]:
    """Device -> host for a dict of result tensors: asynchronous copies into page-locked buffers, one
    synchronisation for the whole set.  With a ``pool`` the buffers are recycled (see ``HostPool``: the
    returned arrays are views that stay valid for two calls); without, they are allocated per call."""
    staged = {}
    stream = None
    slot = pool.next_generation() if pool is not None else None
    for key, t in tensors.items():
        if key.startswith("_"):
            continue
        t = t.contiguous()
        host = slot.get(key) if slot is not None else None
        if host is None or host.shape != t.shape or host.dtype != t.dtype:
            host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            if slot is not None:
                slot[key] = host
        host.copy_(t, non_blocking=True)
        staged[key] = host
        stream = torch.cuda.current_stream(t.device)
    if stream is not None:
        # sleep until the copies are done instead of spinning: on a shared host a spinning waiter gets
        # descheduled for hundreds of milliseconds now and then, a sleeping one is woken by the interrupt
        done = torch.cuda.Event(blocking=True)
        done.record(stream)
        done.synchronize()
    return {k: v.numpy() for k, v in staged.items()}


def library_version() -> str:
    return load_library().lipm_version().decode()


# ====================================================================================== Herdt (free footsteps)
class _HerdtDesc(ctypes.Structure):
    _fields_ = [
        ("horizon", ctypes.c_int32),
        ("h", ctypes.c_double), ("g", ctypes.c_double), ("dt", ctypes.c_double),
        ("alpha", ctypes.c_double), ("beta", ctypes.c_double), ("gamma", ctypes.c_double),
        ("foot_length", ctypes.c_double), ("foot_width", ctypes.c_double), ("foot_spread", ctypes.c_double),
        ("rho", ctypes.c_double), ("relax", ctypes.c_double), ("eps", ctypes.c_double),
        ("max_iter", ctypes.c_int32), ("device", ctypes.c_int32),
        ("stall_window", ctypes.c_int32), ("accept_eps", ctypes.c_double),
    ]


HERDT_SYMBOLS = ("herdt_plan_create", "herdt_plan_destroy", "herdt_plan_export_w", "herdt_timeline_create",
                 "herdt_timeline_destroy", "herdt_timeline_size", "herdt_tick", "herdt_rollout")


def _herdt_lib() -> ctypes.CDLL:
    lib = load_library()
    if not getattr(lib, "_herdt_ready", False):
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        lib.herdt_plan_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_HerdtDesc)]
        lib.herdt_plan_create.restype = i32
        lib.herdt_plan_destroy.argtypes = [vp]
        lib.herdt_plan_destroy.restype = None
        lib.herdt_plan_export_w.argtypes = [vp, vp, ctypes.c_size_t]
        lib.herdt_plan_export_w.restype = i32
        lib.herdt_timeline_create.argtypes = [vp, i32, i32, vp, vp, vp, ctypes.POINTER(vp)]
        lib.herdt_timeline_create.restype = i32
        lib.herdt_timeline_destroy.argtypes = [vp]
        lib.herdt_timeline_destroy.restype = None
        lib.herdt_timeline_size.argtypes = [vp]
        lib.herdt_timeline_size.restype = i32
        lib.herdt_tick.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        lib.herdt_tick.restype = i32
        lib.herdt_rollout.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp, i32, i32, vp, i32,
                                      vp, vp, vp, vp, vp, vp]
        lib.herdt_rollout.restype = i32
        lib._herdt_ready = True
    return lib


class HerdtPlan:
    """Device-resident matrices of the Herdt QP for one ``(N, h, g, dt, alpha, beta, gamma, foot)``."""

    def __init__(self, horizon, h, g, dt, alpha, beta, gamma, foot_length, foot_width, foot_spread, *,
                 rho=0.0, relax=0.0, eps=0.0, max_iter=0, stall_window=0, accept_eps=0.0, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("liblipm_mpc needs a CUDA device (B200 / sm_100a); there is no CPU path")
        self._lib = _herdt_lib()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        with torch.cuda.device(self.device_index):
            torch.zeros(1, device=self.device)
        self.horizon = int(horizon)
        desc = _HerdtDesc(self.horizon, h, g, dt, alpha, beta, gamma, foot_length, foot_width, foot_spread,
                          rho, relax, eps, max_iter, self.device_index, stall_window, accept_eps)
        handle = ctypes.c_void_p()
        _check(self._lib.herdt_plan_create(ctypes.byref(handle), ctypes.byref(desc)))
        self._handle = handle
        self._timelines = []

    def export_w(self) -> np.ndarray:
        N = self.horizon
        out = np.empty((N, N))
        _check(self._lib.herdt_plan_export_w(self._handle, out.ctypes.data, out.size))
        return out

    def timeline(self, phase_of_row: np.ndarray, row_kind: np.ndarray, m_per_tick: np.ndarray) -> "HerdtTimeline":
        return HerdtTimeline(self, phase_of_row, row_kind, m_per_tick)

    def close(self) -> None:
        for t in self._timelines:
            t.close()
        self._timelines = []
        if getattr(self, "_handle", None):
            self._lib.herdt_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class HerdtTimeline:
    """Per-tick Hessians (inverted, in DMMA fragment order) of one support timeline."""

    def __init__(self, plan: HerdtPlan, phase_of_row, row_kind, m_per_tick):
        self.plan = plan
        self._lib = plan._lib
        phase = np.ascontiguousarray(phase_of_row, dtype=np.int8)
        kind = np.ascontiguousarray(row_kind, dtype=np.int8)
        m = np.ascontiguousarray(m_per_tick, dtype=np.int32)
        if phase.shape != kind.shape or phase.ndim != 2 or phase.shape[1] != plan.horizon or len(m) != len(phase):
            raise ValueError("phase_of_row / row_kind must be (n_ticks, N) and m_per_tick (n_ticks,)")
        self.n_ticks = len(m)
        self.m = m
        handle = ctypes.c_void_p()
        _check(self._lib.herdt_timeline_create(plan._handle, self.n_ticks, int(m.max()), phase.ctypes.data,
                                               kind.ctypes.data, m.ctypes.data, ctypes.byref(handle)))
        self._handle = handle
        self.size = int(self._lib.herdt_timeline_size(handle))
        plan._timelines.append(self)

    def tick(self, i: int, x: torch.Tensor, y: torch.Tensor, fc: torch.Tensor, vref: torch.Tensor, foot_side: int,
             poly_verts: Optional[np.ndarray], *, air: Optional[torch.Tensor] = None,
             spread: Optional[torch.Tensor] = None, warm: int = 0,
             want_solution: bool = False) -> Dict[str, torch.Tensor]:
        """``predict_herdt_joint`` for B scenarios (device tensors in, device tensors out)."""
        dev = self.plan.device
        B = x.shape[0]
        vref_batch = 1 if vref.ndim == 2 else vref.shape[0]
        n_ext = vref.shape[-2]
        nx = torch.empty((B, 3), dtype=torch.float64, device=dev)
        ny = torch.empty((B, 3), dtype=torch.float64, device=dev)
        f1 = torch.empty((B, 2), dtype=torch.float64, device=dev)
        status = torch.zeros((2 * B,), dtype=torch.int32, device=dev)
        iters = torch.zeros((2 * B,), dtype=torch.int32, device=dev)
        sol = torch.empty((2 * B, self.size), dtype=torch.float64, device=dev) if want_solution else None
        verts = None if poly_verts is None else np.ascontiguousarray(poly_verts, dtype=np.float64)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(self._lib.herdt_tick(self._handle, int(i), B, x.data_ptr(), y.data_ptr(), fc.data_ptr(), _ptr(air),
                                    _ptr(spread), vref.data_ptr(), vref_batch, n_ext, int(foot_side), None if verts is None else verts.ctypes.data,
                                    0 if verts is None else len(verts), int(warm), nx.data_ptr(), ny.data_ptr(), f1.data_ptr(),
                                    _ptr(sol), status.data_ptr(), iters.data_ptr(), stream))
        out = {"next_x": nx, "next_y": ny, "f1": f1, "status": status, "iters": iters}
        if want_solution:
            out["solution"] = sol
        return out

    def rollout(self, x0: torch.Tensor, y0: torch.Tensor, fc0: torch.Tensor, vref: torch.Tensor, *, n: int, foot_side,
                steps_left, commit, new_run, polygons, spread: Optional[torch.Tensor] = None, warm_start: bool = True,
                impulse: Optional[torch.Tensor] = None, force_tick: int = -1) -> Dict[str, torch.Tensor]:
        """``generate_com_trajectory_herdt`` for B scenarios in ONE call (``herdt_rollout``): the tick loop, the swing-foot
        interpolation, the footstep commits, the push and the histories are enqueued on the current stream without
        coming back to the host between ticks.  Device tensors in, device tensors out."""
        dev = self.plan.device
        B = x0.shape[0]
        ticks = n - 1
        side = np.ascontiguousarray(foot_side[:ticks], dtype=np.int8)
        left = np.ascontiguousarray(steps_left[:ticks], dtype=np.int32)
        com = np.ascontiguousarray(commit[:ticks], dtype=np.uint8)
        new = np.ascontiguousarray(new_run[:ticks], dtype=np.uint8)
        pl, pr = (np.ascontiguousarray(p, dtype=np.float64) for p in polygons)
        vref_batch = 1 if vref.ndim == 2 else vref.shape[0]
        x_hist = torch.empty((B, n, 3), dtype=torch.float64, device=dev)
        y_hist = torch.empty((B, n, 3), dtype=torch.float64, device=dev)
        foot_hist = torch.empty((B, n, 2), dtype=torch.float64, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        iters = torch.empty((B,), dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(self._lib.herdt_rollout(self._handle, B, int(n), x0.data_ptr(), y0.data_ptr(), fc0.data_ptr(), _ptr(spread),
                                       vref.data_ptr(), vref_batch, vref.shape[-2], side.ctypes.data, left.ctypes.data,
                                       com.ctypes.data, new.ctypes.data, pl.ctypes.data, len(pl), pr.ctypes.data, len(pr),
                                       1 if warm_start else 0, _ptr(impulse), int(force_tick), x_hist.data_ptr(),
                                       y_hist.data_ptr(), foot_hist.data_ptr(), status.data_ptr(), iters.data_ptr(), stream))
        return {"x_hist": x_hist, "y_hist": y_hist, "foot_hist": foot_hist, "status": status, "iters": iters,
                "_keepalive": (x0, y0, fc0, vref, spread, impulse)}

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self._lib.herdt_timeline_destroy(self._handle)
            self._handle = None
