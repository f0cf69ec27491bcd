#!/usr/bin/env python3
"""Headline benchmark: QP solves/sec of the batched N=150 strict LIPM-MPC push sweep.

Workload (BASELINE.json configs[1]): 4096 scenarios per GPU, each a full configs/default.json
walk (420 bound samples -> 419 ticks x 2 axes = 838 strict QP solves), with a push
F_ext ~ U(0,1000) N at a tick ~ U{100..319}, m = 40 kg, seed 0, bounds shared by all scenarios.
One "step" = one sweep = B x 838 solves.  Scenarios shard across GPUs with no data-path
collective; the only exchange is one NCCL all-gather of the per-scenario outcomes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (see the keys below).  `--impl reference` times the CPU restatement of the
reference path (oracle port; the reference's own strict path needs cvxpy/OSQP, which are not
installed, and its Python sources do not travel to the GPU box) on all host cores.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
PACKAGE_DIR = os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")
for _p in (REPO, PACKAGE_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "QP solves/sec (N=150 LIPM MPC, strict, batched push sweep)"
UNIT = "solves/s"
SCENARIOS_PER_GPU = 4096
RICCATI_FLOPS_PER_STAGE = 184      # one stage of a sweep pair: backward 75 + forward 17 FP64 multiply-adds
WORKLOAD = "push-sweep-4096 (BASELINE configs[1]): default.json walk n=420, N=150, strict, F_ext~U(0,1000)N, tick~U{100..319}"


def bench_config(B, n, N, mode):
    """The `config` of the JSON line: static facts of the workload only, identical for both arms."""
    return {"workload": WORKLOAD, "scenarios_per_gpu": B, "horizon": N, "samples_per_walk": n, "mode": mode,
            "solves_per_step_per_gpu": B * 2 * (n - 1),
            "l2": "flushed between timed steps (256 MiB write, untimed)",
            "qp": "box-constrained preview QP of zmp_controller.py:173-191, both axes, solved to its unique minimiser in FP64"}


def sweep_inputs(B, seed=0):
    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.generators import CoPGenerator

    cfg = MPCConfig.from_json(os.path.join(REPO, "configs", "default.json"))
    z_max, z_min, _ = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
    rng = np.random.default_rng(seed)
    F = rng.uniform(0.0, 1000.0, B)
    tick = rng.integers(100, 320, B).astype(np.int32)
    return cfg, z_max, z_min, F, tick


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU baseline
def host_threads():
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would silently turn the
    # baseline into a single-thread run)
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(steps=1, warmup=1, budget_s=25.0, batch=SCENARIOS_PER_GPU):
    """Oracle port (oracle/lipm_oracle.c: same ADMM, cached matrices, warm start) on all host cores.

    A bounded sample of the SAME workload: the first `walks` scenarios of the 4096-scenario push sweep (seed 0, the
    set the GPU arm runs), full default.json strict walks of 838 solves each, OpenMP over scenarios; `walks` is
    sized from a calibration pass so that warmup + steps passes take about `budget_s` seconds.  The reference itself
    would also rebuild Px/Pu in Python and go through cvxpy/OSQP on every solve (not runnable here), so this port is
    an optimistic stand-in for it."""
    from oracle.c_oracle import COracle

    try:
        oracle = COracle(150, native=True)           # -march=native, built on this host
    except Exception:
        oracle = COracle(150)                        # portable prebuilt (x86-64-v3)
    threads = host_threads()
    cfg, z_max, z_min, F, tick = sweep_inputs(batch)
    impulse = cfg.dt * F / cfg.m
    zero = np.zeros((batch, 3))

    def run(count):
        t0 = time.perf_counter()
        res = oracle.rollout(True, zero[:count], zero[:count], z_min, z_max, impulse=impulse[:count],
                             force_tick=tick[:count], threads=threads)
        return time.perf_counter() - t0, res

    run(threads)                                      # page in the library, spin up the OpenMP team
    calib_s, _ = run(2 * threads)                     # two walks per thread
    per_walk = calib_s / (2 * threads)
    walks = int(budget_s / max(steps + warmup, 1) / per_walk)
    walks = max(threads, min(batch, (walks // threads) * threads))
    for _ in range(warmup):
        run(walks)
    times = []
    for _ in range(steps):
        dt, res = run(walks)
        times.append(dt)
    assert (res["status"] == 0).all()
    solves = walks * 2 * (len(z_min) - 1)
    # the GPU arm's own algorithm (active set + Riccati sweeps) restated in C on the same cores, same sample: separates
    # what the algorithm buys from what the hardware buys
    t0 = time.perf_counter()
    res_as = oracle.rollout("riccati", zero[:walks], zero[:walks], z_min, z_max, impulse=impulse[:walks],
                            force_tick=tick[:walks], threads=threads)
    t_as = time.perf_counter() - t0
    same_algorithm = {"value": solves / t_as, "unit": UNIT, "cores": threads, "seconds": t_as,
                      "mean_sweep_pairs_per_solve": float(res_as["iters"].sum()) / solves,
                      "not_converged": int((res_as["status"] != 0).sum()),
                      "note": "oracle/lipm_oracle.c mode 2: primal-dual active set with O(N) Riccati sweeps, the CPU "
                              "restatement of csrc/lipm_rollout_riccati.cu (not the reference's algorithm: the reference "
                              "calls OSQP, an ADMM method, cold-started on every solve)"}
    return {"value": solves / float(np.mean(times)), "unit": UNIT, "cores": threads, "kind": "port",
            "active_set_port": same_algorithm,
            "sample": f"first {walks} of the {batch} scenarios of the push sweep (seed 0), full default.json strict walks "
                      f"({solves} solves) per step; C restatement (oracle/lipm_oracle.c), FP64 ADMM rho=2 alpha=1.75 "
                      f"eps=1e-12, matrices cached, warm start, OpenMP over scenarios on {threads} threads",
            "mean_iters_per_solve": float(res["iters"].sum()) / solves,
            "seconds_per_step": float(np.mean(times))}, times, (cfg, len(z_min))


def numpy_single_core(budget_s=6.0):
    """BASELINE.md section 4, items 1 and 3, on ONE core with NumPy (BLAS pinned to one thread): the reference's
    per-solve cost structure restated -- (a) "as the reference": Px / Pu rebuilt by the Python double loop and the
    N x N inverse recomputed on every solve (zmp_controller.py:162-171,196-198; oracle.Problem.make is that loop),
    (b) "fair NumPy": matrices cached, the solve alone, non-strict and strict (FP64 ADMM, warm start along a walk).
    The reference's own sources are not on the GPU box, so these are restatements, labelled as such."""
    from oracle import lipm_oracle as orc

    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    cfg, z_max, z_min, F, tick = sweep_inputs(1)
    N, n = cfg.horizon, len(z_min)
    hi = np.vstack([z_max, np.tile(z_max[-1:], (N, 1))])
    lo = np.vstack([z_min, np.tile(z_min[-1:], (N, 1))])
    x = np.array([0.01, 0.05, 0.1])
    out = {"cores": 1, "horizon": N, "kind": "NumPy restatement (oracle/lipm_oracle.py), not the reference's own files"}
    # (a) rebuild + invert per solve, non-strict
    t0, k = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s / 3 or k < 2:
        p = orc.Problem.make(N, cfg.h, cfg.g, cfg.Q, cfg.R)
        zref = (hi[1:1 + N, 1] + lo[1:1 + N, 1]) / 2
        X = -np.linalg.inv(p.Pu.T @ p.Pu + cfg.R / cfg.Q * np.eye(N)) @ p.Pu.T @ (p.Px @ x - zref)
        _ = p.A @ x + p.B * X[0]
        k += 1
    out["as_reference_nonstrict_solves_per_s"] = k / (time.perf_counter() - t0)
    # (b) cached matrices
    p = orc.Problem.make(N, cfg.h, cfg.g, cfg.Q, cfg.R)
    p._cache()
    t0, k = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s / 6:
        orc.unconstrained_axis(p, x, hi[1:1 + N, 1], lo[1:1 + N, 1])
        k += 1
    out["cached_nonstrict_solves_per_s"] = k / (time.perf_counter() - t0)
    ticks = 0
    t0 = time.perf_counter()
    res = None
    for ticks in (12, 40):                               # a slice of the default strict walk around the push
        t0 = time.perf_counter()
        res = orc.rollout_wieber(p, np.zeros(3), np.zeros(3), z_max[190:], z_min[190:], strict=True,
                                 impulse=cfg.dt * cfg.F_ext / cfg.m, force_tick=20, solver="admm", n_ticks=ticks)
        if time.perf_counter() - t0 > budget_s / 4:
            break
    out["cached_strict_admm_solves_per_s"] = 2 * ticks / (time.perf_counter() - t0)
    out["strict_sample"] = f"{ticks} ticks x 2 axes of the default walk from sample 190 (push at its tick 20), eps 1e-13, warm start"
    if limiter is not None:
        limiter.restore_original_limits() if hasattr(limiter, "restore_original_limits") else None
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    base, times, (cfg, n) = cpu_baseline(steps=steps, warmup=warmup, budget_s=120.0, batch=args.batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.batch, n, cfg.horizon, args.mode),
        "reference_note": "CPU restatement of the reference path on all host cores; the reference's strict branch needs "
                          "cvxpy+OSQP (absent) and its sources are not on this box; each step is a bounded sample of the "
                          "workload (cpu_baseline.sample)",
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def ncu_traffic(kernel, batch):
    """dram__bytes_read + dram__bytes_write of the dominant kernel, per launch, from the committed
    `ncu --set full` summary of this workload (profiles/*.json written by tools/ncu_summary.py)."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(REPO, "profiles", "r*_ncu_*.json"))):
        try:
            with open(path) as fh:
                for rec in json.load(fh):
                    if kernel in rec.get("kernel", "") and rec.get("scenarios_per_gpu") == batch:
                        total = 0.0
                        for key, val in rec.items():
                            if key.startswith("dram__bytes_read.sum") or key.startswith("dram__bytes_write.sum"):
                                scale = {"[byte]": 1.0, "[Kbyte]": 1e3, "[Mbyte]": 1e6, "[Gbyte]": 1e9}[key.split()[-1]]
                                total += val * scale
                        best = {"bytes_per_launch": total, "source": os.path.relpath(path, REPO)}
        except Exception:
            continue
    return best


# ------------------------------------------------------------------------------ GPU arm
def measure_fp64_peak(torch, device, n=4096, reps=5):
    """cuBLAS DGEMM throughput measured like MEASURED_PEAKS.json measures bf16 (best of reps)."""
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def other_configs(torch, engine, device):
    """BASELINE configs[2] (horizon scaling, one strict tick of 16384 problems) and configs[3] (Herdt, 2048 walks of
    the class-default timeline) in one device-timed number each.  Not the headline: context for it."""
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import horizon_scaling
    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.controllers import ZMPController
    from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator

    out = {}
    B = 16384
    for N in (300, 600):
        for precision, tag, solver in ((engine.PRECISION_FP64, "fp64", "active_set"), (engine.PRECISION_FP64, "fp64_admm", "admm"),
                                       (engine.PRECISION_F16X3, "f16x3", "admm")):
            plan = engine.get_plan(N, 0.75, 9.81, 1.5 / N, 1.0, 1e-6, precision=precision, solver=solver)
            x, lo, hi = horizon_scaling.tick_inputs(N, B)
            xd, lod, hid = (torch.as_tensor(a).to(device) for a in (x, lo, hi))
            plan.solve_tick(engine.MODE_STRICT, xd, lod, hid)
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                res = plan.solve_tick(engine.MODE_STRICT, xd, lod, hid)
            e1.record()
            torch.cuda.synchronize(device)
            ms = e0.elapsed_time(e1) / 3
            iters = int(res["iters"].sum().item())
            rec = {"batch": B, "ms_per_tick": ms, "solves_per_s": B / (ms * 1e-3), "start": "cold",
                   "not_converged": int((res["status"] != 0).sum().item())}
            if tag == "fp64":
                rec.update({"solver": "active set + Riccati sweeps", "mean_sweep_pairs_per_solve": iters / B})
            else:
                rec.update({"solver": "ADMM", "algorithmic_tflops": 2.0 * N * N * iters / (ms * 1e-3) / 1e12,
                            "mean_iters_per_solve": iters / B})
            out[f"strict_tick_N{N}_{tag}"] = rec
    out["herdt_randomized_plans_fp64"] = herdt_plans_leg(torch, engine, device)
    return out


def herdt_plans_leg(torch, engine, device, plans_count=8192):
    """BASELINE configs[3]: 8192 randomized footstep plans (step_length, foot_spread, distance, speed; seed 2) through the
    Herdt free-footstep MPC, strict constraints -- one batch per timeline group, each on its own stream, timed on the
    device from the first enqueue to the last group's completion (the per-timeline Hessians are built before)."""
    from mpc_bipedal.config import MPCConfig
    from mpc_bipedal.herdt import randomized_plans, run_plan_groups

    cfg = MPCConfig()                      # class defaults, as the reference's driver (run_herdt_method.py:53)
    cfg.method = "herdt"
    t0 = time.perf_counter()
    plans = randomized_plans(cfg, plans_count, seed=2, on_device=True)
    t_plans = time.perf_counter() - t0
    t0 = time.perf_counter()
    warm = dict(plans)
    warm["groups"] = [dict(g, ids=g["ids"][:2]) for g in plans["groups"]]
    _, controllers, streams = run_plan_groups(cfg, warm)             # builds every timeline, loads the kernels
    torch.cuda.synchronize(device)
    t_build = time.perf_counter() - t0
    start = torch.cuda.Event(enable_timing=True)
    start.record()
    w0 = time.perf_counter()
    results, _, streams = run_plan_groups(cfg, plans, controllers)
    ends = []
    for st in streams:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(st)
        ends.append(ev)
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - w0
    seconds = max(start.elapsed_time(ev) for ev in ends) * 1e-3
    solves = iters = flops = 0
    failed_plans = failed_ticks = 0
    for g, res, ctrl in zip(plans["groups"], results, controllers):
        ticks = len(g["states"]) - 1
        it = int(res["iters"].sum().item())
        n_var = cfg.horizon + float(ctrl.prepare_herdt(g["states"]).layout.m.mean())
        solves += len(g["ids"]) * ticks
        iters += it
        flops += 2.0 * 2.0 * n_var * n_var * it
        failed_plans += int((res["status"] != 0).sum().item())
        failed_ticks += int(res["status"].sum().item())
    peak = measure_fp64_peak(torch, device)
    return {"plans": plans_count, "timeline_groups": len(plans["groups"]), "joint_qp_solves": solves, "seconds": seconds,
            "wall_seconds": wall, "joint_qp_solves_per_s": solves / seconds, "mean_iters_per_joint_solve": iters / max(solves, 1),
            "plans_with_fallback_ticks": failed_plans, "fallback_ticks": failed_ticks,
            "fallback_tick_fraction": failed_ticks / max(solves, 1),
            "seconds_generating_plans_on_device": t_plans, "seconds_building_timelines_and_warmup": t_build,
            "roofline": {"bound": "tensor", "achieved": flops / seconds / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / seconds / 1e12 / peak, "traffic": None, "kernel": "k_qp_tick (FP64 DMMA ADMM, joint x/y)",
                         "note": "algorithmic flops = ADMM iterations x 2 axes x 2 (N + m)^2 with m = mean free footsteps "
                                 "of the timeline; peak = cuBLAS DGEMM measured in this run"}}


def param_sweep_leg(torch, dist, engine, device, rank, world, per_gpu=131072):
    """BASELINE configs[4]: the (h, m, step_length, foot_spread, F_ext) sweep, `per_gpu` scenarios per GPU (8 GPUs =
    the 1,048,576 scenarios the config names), bound series generated on the device, scenarios dealt round-robin inside
    every CoM-height bucket, ONE all-gather of the packed outcome records.  Wall clock between barriers, max over
    ranks; FP64 and reduced precision."""
    from mpc_bipedal import parallel, sweeps
    from mpc_bipedal.config import MPCConfig

    cfg = MPCConfig.from_json(os.path.join(REPO, "configs", "default.json"))
    total = per_gpu * world
    table = sweeps.make_sweep(total, seed=3)
    warm = sweeps.make_sweep(256 * world, seed=3)
    result = {"scenarios": total, "scenarios_per_gpu": per_gpu, "gpus": world, "scaling": "weak",
              "h_values": len(table.h_values), "bound_variants": len(table.variants)}
    for tag, precision in (("f16x3", engine.PRECISION_F16X3), ("fp64", engine.PRECISION_FP64)):
        w = sweeps.run_shard(cfg, warm, rank, world, precision=precision)          # plans of all 8 CoM heights, kernels
        parallel.gather_by_index(w["record"], w["index"], len(warm), world)
        torch.cuda.synchronize(device)
        # one untimed pass at full size: the stream-ordered pool grows to the sweep's scratch size on first use (page
        # mapping: the first pass varied between 0.65 and 0.96 s from box to box), like the warm-up steps of the headline
        t_first = time.perf_counter()
        sweeps.run_shard(cfg, table, rank, world, precision=precision)
        torch.cuda.synchronize(device)
        t_first = time.perf_counter() - t_first
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        local = sweeps.run_shard(cfg, table, rank, world, precision=precision)
        torch.cuda.synchronize(device)
        t_walks = time.perf_counter() - t0
        gathered = parallel.gather_by_index(local["record"], local["index"], total, world)
        torch.cuda.synchronize(device)
        seconds = time.perf_counter() - t0
        agg = torch.tensor([seconds, t_walks, float(local["solves"])], dtype=torch.float64, device=device)
        if world > 1:
            mx = agg.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            seconds, t_walks, solves = mx[0].item(), mx[1].item(), agg[2].item()
        else:
            solves = agg[2].item()
        rec = parallel.unpack_outcomes(gathered)
        result[tag] = {"seconds": seconds, "seconds_first_untimed_pass_this_rank": t_first, "seconds_walks": t_walks, "seconds_all_gather_and_scatter": seconds - t_walks,
                       "qp_solves": int(solves), "qp_solves_per_s": solves / seconds,
                       "not_converged": int((rec["status"] != 0).sum().item()), "fallen": int(rec["fell"].sum().item()),
                       "max_zmp_violation_m": rec["max_violation"].max().item(),
                       "mean_iters_per_solve": rec["iters"][rec["status"] == 0].sum().item() / max(solves, 1)}
    return result


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from mpc_bipedal import engine
    from mpc_bipedal.controllers import ZMPController

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # the version banner goes to stdout, next to the JSON line
        dist.init_process_group("nccl", device_id=device)

    B = args.batch
    cfg, z_max, z_min, F, tick = sweep_inputs(B, seed=0 if args.same_sweep else rank)
    n = len(z_min)
    solves_per_step = B * 2 * (n - 1)
    precision = engine.PRECISION_F16X3 if args.precision == "f16x3" else engine.PRECISION_FP64
    plan = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, precision=precision, solver=args.solver)
    mode = engine.MODE_STRICT if args.mode == "strict" else engine.MODE_UNCONSTRAINED
    active_set = precision == engine.PRECISION_FP64 and args.solver == "active_set"

    # device-resident inputs for `value`
    d = lambda a, dt: torch.as_tensor(a, dtype=dt).to(device)
    x0 = torch.zeros((B, 3), dtype=torch.float64, device=device)
    y0 = torch.zeros((B, 3), dtype=torch.float64, device=device)
    zmin_d, zmax_d = d(z_min, torch.float64), d(z_max, torch.float64)
    imp_d, tick_d = d(cfg.dt * F / cfg.m, torch.float64), d(tick, torch.int32)
    out = {}
    engine.keep_pool_memory(local_rank)      # opt-in: the rollouts' scratch buffers stay cached between calls
    gathered = [torch.empty((B, engine.OUTCOME_WORDS), dtype=torch.int64, device=device) for _ in range(world)] if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2

    def step():
        plan.rollout(mode, x0, y0, zmin_d, zmax_d, impulse=imp_d, force_tick=tick_d, out=out)
        if world > 1:   # the path's only exchange: the packed per-scenario outcome records (lipm_outcome_t, 32 B each)
            dist.all_gather(gathered, out["outcome"])

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)

    events = []
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            flush.fill_(1)                      # evict L2 between timed iterations (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record()
            events.append((e0, e1))
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
    step_ms = [a.elapsed_time(b) for a, b in events]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = total_ms.item() * 1e-3
    value = world * solves_per_step * args.steps / total_s

    slots = 128 if precision == engine.PRECISION_F16X3 else 8      # columns a solver unit iterates together
    iters = out["iters"].sum().item() if mode == engine.MODE_STRICT else 0
    # latency floor of the active-set path, measured live: the scenario with the most sweep pairs run ALONE (its two
    # columns have the machine to themselves); no launch shape can bring the step below the walk of its slowest column
    floor_ms = None
    if rank == 0 and active_set and mode == engine.MODE_STRICT:
        worst = int(out["iters"].sum(dim=1).argmax().item())
        one = {}
        run_one = lambda: plan.rollout(mode, x0[:1], y0[:1], zmin_d, zmax_d, impulse=imp_d[worst:worst + 1].contiguous(),
                                       force_tick=tick_d[worst:worst + 1].contiguous(), out=one)
        run_one()
        torch.cuda.synchronize(device)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(); run_one(); f1.record()
        torch.cuda.synchronize(device)
        floor_ms = f0.elapsed_time(f1)
    stats = plan.stats() if mode == engine.MODE_STRICT else {"group_iterations": 0, "units": 0}
    status_bad = int((out["status"] != 0).sum().item())
    max_viol = out["max_violation"].max().item()

    # ---- the same sweep on the ADMM / FP64 tensor-pipe kernel (the headline solver of round 1): its own roofline and
    #      the agreement of the two solvers, which share no arithmetic ---------------------------------------------
    admm_leg = None
    if mode == engine.MODE_STRICT and active_set and not args.no_reduced:
        plan_admm = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, solver="admm")
        out_a = {}
        for _ in range(2):
            plan_admm.rollout(mode, x0, y0, zmin_d, zmax_d, impulse=imp_d, force_tick=tick_d, out=out_a)
        torch.cuda.synchronize(device)
        ev_a = []
        for _ in range(min(args.steps, 3)):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan_admm.rollout(mode, x0, y0, zmin_d, zmax_d, impulse=imp_d, force_tick=tick_d, out=out_a); e1.record()
            ev_a.append((e0, e1))
        torch.cuda.synchronize(device)
        ms_a = float(np.mean([a.elapsed_time(b) for a, b in ev_a]))
        it_a = out_a["iters"].sum().item()
        st_a = plan_admm.stats()
        peak_a = measure_fp64_peak(torch, device)
        tf_a = 2.0 * cfg.horizon ** 2 * it_a / (ms_a * 1e-3) / 1e12
        scale = torch.clamp(out_a["y_hist"].abs().amax(dim=(1, 2)), min=1.0)
        dxy = torch.maximum((out["x_hist"] - out_a["x_hist"]).abs().amax(dim=(1, 2)), (out["y_hist"] - out_a["y_hist"]).abs().amax(dim=(1, 2)))
        admm_leg = {
            "value_per_gpu": solves_per_step / (ms_a * 1e-3), "unit": UNIT, "ms_per_step": ms_a,
            "kernel": "k_strict_rollout (FP64 DMMA.8x8x4 ADMM, rho=2 alpha=1.75 eps=1e-12)",
            "mean_iters_per_solve": it_a / max(solves_per_step, 1),
            "useful_column_slot_fraction": it_a / max(8 * st_a["group_iterations"], 1),
            "roofline": {"bound": "tensor", "achieved": tf_a, "peak": peak_a, "unit": "TFLOP/s", "frac": tf_a / peak_a,
                         "traffic": ncu_traffic("k_strict_rollout<", B),
                         "note": "algorithmic K_exec * 2 N^2 flops per solve (SURVEY 8(d)); peak = cuBLAS DGEMM measured in this run"},
            "agreement_with_headline_solver": {
                "max_rel_diff_trajectories": float((dxy / scale).max().item()),
                "median_abs_diff_m": float(dxy.median().item()),
                "flag_bytes_that_differ": int((out["flags"] != out_a["flags"]).sum().item()),
                "fall_verdict_mismatches": int((engine.unpack_outcome(out["outcome"])["fell"] != engine.unpack_outcome(out_a["outcome"])["fell"]).sum().item())},
        }

    # ---- the same sweep on the reduced-precision (tcgen05) strict path -------------------------------
    reduced = None
    if mode == engine.MODE_STRICT and precision == engine.PRECISION_FP64 and not args.no_reduced:
        plan16 = engine.get_plan(cfg.horizon, cfg.h, cfg.g, cfg.dt, cfg.Q, cfg.R, precision=engine.PRECISION_F16X3)
        out16 = {}
        for _ in range(args.warmup):
            plan16.rollout(mode, x0, y0, zmin_d, zmax_d, impulse=imp_d, force_tick=tick_d, out=out16)
        torch.cuda.synchronize(device)
        ev16 = []
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan16.rollout(mode, x0, y0, zmin_d, zmax_d, impulse=imp_d, force_tick=tick_d, out=out16); e1.record()
            ev16.append((e0, e1))
        torch.cuda.synchronize(device)
        ms16 = [a.elapsed_time(b) for a, b in ev16]
        tot16 = torch.tensor([sum(ms16)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tot16, op=dist.ReduceOp.MAX)
        it16 = out16["iters"].sum().item()
        st16 = plan16.stats()
        # agreement with the FP64 path: over ALL scenarios, and over those that absorb their push with margin
        # (a saturated ZMP leaves the open-loop unstable pendulum, which amplifies any
        # solver difference by e^(t / 0.28 s))
        lean = out["lean_max"][:, 1]                                    # max |CoM - ZMP| on the pushed axis (FP64 walk), from the kernel
        calm = lean < 0.12
        dy = (out16["y_hist"][..., 0] - out["y_hist"][..., 0]).abs().amax(dim=1)
        dx = (out16["x_hist"][..., 0] - out["x_hist"][..., 0]).abs().amax(dim=1)
        diff = torch.maximum(dx, dy)
        fell64 = engine.unpack_outcome(out["outcome"])["fell"]
        fell16 = engine.unpack_outcome(out16["outcome"])["fell"]
        quant = torch.quantile(diff, torch.tensor([0.5, 0.9, 0.99], dtype=diff.dtype, device=device)).tolist()
        flops16 = 2.0 * cfg.horizon ** 2 * it16
        s16 = tot16.item() * 1e-3 / args.steps
        peak16 = None
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
                peak16 = float(json.load(fh)["bf16_tflops"])
        except Exception:
            pass
        reduced = {
            "value": world * solves_per_step / s16, "unit": UNIT, "ms_per_step": float(np.mean(ms16)),
            "dtype": "f16 hi/lo x3 operands, f32 accumulate + ADMM state, f64 walk state",
            "kernel": "k_strict_rollout_f16p (tcgen05.mma.cta_group::2 kind::f16 on CTA pairs, TMEM accumulators)",
            "eps": 1e-5, "mean_iters_per_solve": it16 / max(solves_per_step, 1),
            "useful_column_slot_fraction": it16 / max(128 * st16["group_iterations"], 1),
            "not_converged_columns": int((out16["status"] != 0).sum().item()),
            "max_zmp_violation_m": out16["max_violation"].max().item(),
            "tolerance_m": 1e-4,
            "vs_fp64_path": {
                "max_abs_diff_m_all_scenarios": diff.max().item(),
                "fraction_within_tolerance": float((diff <= 1e-4).double().mean().item()),
                "abs_diff_m_p50_p90_p99": quant,
                "max_abs_diff_m_calm_scenarios": diff[calm].max().item() if bool(calm.any()) else None,
                "calm_scenarios": int(calm.sum().item()),
                "fell_fp64": int(fell64.sum().item()), "fell_reduced": int(fell16.sum().item()),
                "fall_verdict_mismatches": int((fell64 != fell16).sum().item()),
                "max_abs_diff_m_among_scenarios_that_do_not_fall": diff[~fell64].max().item() if bool((~fell64).any()) else None,
                "fraction_within_tolerance_among_scenarios_that_do_not_fall":
                    float((diff[~fell64] <= 1e-4).double().mean().item()) if bool((~fell64).any()) else None,
                "note": "CoM position, both axes, whole walk; 'calm' = FP64 walk with max |CoM_y - ZMP_y| < 0.12 m; a scenario "
                        "'falls' when the pendulum leans more than 30 degrees (lipm_outcome_t.fell)"},
            "roofline": {"bound": "tensor", "achieved": flops16 / s16 / 1e12, "peak": peak16, "unit": "TFLOP/s",
                         "frac": (flops16 / s16 / 1e12 / peak16) if peak16 else None,
                         "executed_tflops": 3.0 * (160.0 / cfg.horizon) ** 2 * 128.0 * st16["group_iterations"] * 2.0 * cfg.horizon ** 2 / s16 / 1e12,
                         "note": "achieved = algorithmic K_exec*2N^2 per solve; the kernel executes 3 FP16 passes on a "
                                 "160x160 tile for every column slot (executed_tflops); peak = MEASURED_PEAKS bf16 burst"},
        }

    # ---- end to end through the public API with host buffers ---------------------------------------
    ctrl = ZMPController(cfg)
    ctrl.config.strict = (args.mode == "strict")
    ctrl.solver_options = {"precision": precision, "solver": args.solver}
    hx0 = np.zeros((B, 3)); hy0 = np.zeros((B, 3))
    e2e_steps = max(1, args.steps)
    for _ in range(3):   # warm-up: the page-locked staging buffers are allocated on the first calls
        res = ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
    torch.cuda.synchronize(device)
    e2e_calls = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        c0 = time.perf_counter()
        res = ctrl.generate_com_trajectory_batched(hx0, hy0, z_max, z_min, F_ext=F, force_tick=tick)
        e2e_calls.append(time.perf_counter() - c0)      # the call returns host arrays: it has synchronised
    torch.cuda.synchronize(device)
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = hx0.nbytes + hy0.nbytes + z_max.nbytes + z_min.nbytes + F.nbytes + tick.nbytes
    d2h = int(sum(v.nbytes for v in res.values()))
    e2e_value = world * solves_per_step / e2e_s.item()

    # ---- p50 latency of ONE control tick for the whole batch (BASELINE metric, second half): x-solve + y-solve of
    #      all scenarios = 2 B strict QPs in one lipm_solve_tick call, cold start, each call timed by its own
    #      pair of CUDA events, 100 calls ---------------------------------------------------------------------
    tick_latency = None
    if rank == 0 and mode == engine.MODE_STRICT:
        sys.path.insert(0, os.path.join(REPO, "tools"))
        import horizon_scaling
        xt, lot, hit = horizon_scaling.tick_inputs(cfg.horizon, 2 * B, seed=7)
        xt, lot, hit = (torch.as_tensor(a).to(device) for a in (xt, lot, hit))
        for _ in range(3):
            plan.solve_tick(mode, xt, lot, hit)
        torch.cuda.synchronize(device)
        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
        for a, b in events:
            a.record()
            plan.solve_tick(mode, xt, lot, hit)
            b.record()
        torch.cuda.synchronize(device)
        tick_ms = np.array([a.elapsed_time(b) for a, b in events])
        tick_latency = {"p50_ms": float(np.median(tick_ms)), "p99_ms": float(np.percentile(tick_ms, 99)), "ticks": 100,
                        "problems_per_tick": 2 * B, "start": "cold (no previous iterate), random states and bounds"}

    # ---- the other BASELINE configs in one number each (single GPU run only; seconds, not minutes) ----------
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = other_configs(torch, engine, device)
    # BASELINE configs[4] (parameter sweep, the config that is multi-GPU by definition): every rank, every world size
    sweep = None
    if not args.no_extras and mode == engine.MODE_STRICT:
        sweep = param_sweep_leg(torch, dist, engine, device, rank, world)

    if rank == 0:
        N = cfg.horizon
        kernel_s = np.mean(step_ms) * 1e-3
        peak = measure_fp64_peak(torch, device)
        if active_set:
            # one sweep pair of the Riccati recursion = N stages x (backward 75 + forward 17 FP64 multiply-adds)
            # ~ 184 N flops; the row-parallel stage does less arithmetic per sweep on the ticks it takes, so this is
            # an upper estimate of the work actually executed (the time is that of ALL kernels of the step)
            flops = RICCATI_FLOPS_PER_STAGE * N * iters
            traffic = ncu_traffic("k_riccati_rollout", B)
            if traffic is not None:
                traffic["launch"] = ("the longest of the step's Riccati launches (first chunk of the processing order, parked columns); "
                                     "the step has two such launches per chunk")
            st = stats                       # of the last timed step
            roofline = {"bound": "fp64 pipe (not tensor, not hbm: scalar FP64 recursions with one dependent chain per column; "
                                 "latency bound: the step ends with the walks of the scenarios that fall)",
                        "kernel": "k_riccati_rollout (active set + Riccati sweeps, one thread per column: the columns the "
                                  "row-parallel kernel k_rowpar_rollout parks after a strong push; ~80 % of the step)",
                        "columns_parked_on_the_riccati_kernel": int(st.get("columns_parked", 0)),
                        "columns_handed_back": int(st.get("columns_handed_back", 0)),
                        "latency_floor": {"heaviest_scenario_alone_ms": floor_ms, "step_ms": float(np.mean(step_ms)),
                                          "frac": (floor_ms / float(np.mean(step_ms))) if floor_ms else None,
                                          "note": "the scenario with the most sweep pairs of this sweep run ALONE, measured in this "
                                                  "run (two columns, one warp each: one dependent chain of ~650 sweep pairs on its "
                                                  "pushed axis); frac = that walk / the whole 4096-scenario step"},
                        "note": "achieved = sweep pairs x N x %d flops (the active-set algorithm's own work; the ADMM formulation "
                                "needs ~150x more flops per solve, see admm_tensor_path) / time of the whole step; peak = cuBLAS "
                                "DGEMM measured in this run (B200's FP64 FMA pipe and DMMA have the same nominal rate)" % RICCATI_FLOPS_PER_STAGE}
        else:
            flops = 2.0 * N * N * iters                      # K_exec * 2 N^2 per solve
            traffic = ncu_traffic("k_strict_rollout_f16" if precision == engine.PRECISION_F16X3 else "k_strict_rollout<", B)
            roofline = {"bound": "tensor", "kernel": "k_strict_rollout (FP64 DMMA.8x8x4 ADMM)"}
        achieved = flops / kernel_s / 1e12
        roofline.update({"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "cuBLAS DGEMM 4096^3 via torch.matmul, best of 5, measured in this run "
                                        "(MEASURED_PEAKS.json has no FP64 entry)",
                         "algorithmic_flops_per_step": flops})
        if active_set:
            solver_block = {"name": "active set (LIPM_SOLVER_ACTIVE_SET): row-parallel kernel with dense corrections, Riccati sweeps for large sets",
                            "mean_sweep_pairs_per_solve": iters / max(solves_per_step, 1),
                            "mean_sweep_pairs_per_solve_x_axis": out["iters"][:, 0].sum().item() / max(B * (n - 1), 1),
                            "mean_sweep_pairs_per_solve_y_axis": out["iters"][:, 1].sum().item() / max(B * (n - 1), 1),
                            "max_sweep_pairs_of_one_column": int(out["iters"].max().item())}
        else:
            solver_block = {"name": "ADMM (rho=2, alpha=1.75, eps=1e-12)" if precision == engine.PRECISION_FP64 else "ADMM f16x3 (eps=1e-5)",
                            "mean_iters_per_solve": iters / max(solves_per_step, 1),
                            "useful_column_slot_fraction": iters / max(slots * stats["group_iterations"], 1),
                            "mean_iters_per_solve_x_axis": out["iters"][:, 0].sum().item() / max(B * (n - 1), 1),
                            "mean_iters_per_solve_y_axis": out["iters"][:, 1].sum().item() / max(B * (n - 1), 1)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(B, n, N, args.mode),
            "solver": solver_block,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                    "api": "ZMPController.generate_com_trajectory_batched (NumPy in, NumPy out)",
                    "ms_per_call": [round(1e3 * c, 2) for c in e2e_calls],
                    "median_ms_per_call": round(1e3 * float(np.median(e2e_calls)), 2),
                    "host_loadavg": list(os.getloadavg()),
                    "note": "value = mean over the calls (wall clock, NumPy in / NumPy out, all histories copied back)"},
            # per timed step: k_tick_span + k_order_keys (+ k_unit_work, ADMM) + the rollout kernels + k_pack_outcomes (the
            # CUB sort passes between them are library kernels and not counted).  Active set: k_rowpar_pack, k_preview_u,
            # then per chunk of the batch (four from 4096 scenarios on, LIPM_OPT_WALK_CHUNKS) 2 x k_rowpar_rollout and
            # 2 x (k_pack_bounds + k_riccati_rollout).  Unconstrained: preview + rollout + pack
            "gpu_launches": ((5 + 6 * (4 if B >= 4096 else 1) if active_set else 5) if mode == engine.MODE_STRICT else 3) * args.steps,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "checks": {"not_converged_columns": status_bad, "max_zmp_violation_m": max_viol},
            "latency": {"p50_step_ms": float(np.median(step_ms)),
                        "p50_tick_ms_amortised": float(np.median(step_ms)) / (n - 1),
                        "single_tick": tick_latency},
        }
        if admm_leg is not None:
            line["admm_tensor_path"] = admm_leg
        if reduced is not None:
            line["reduced_precision"] = reduced
        if extras is not None or sweep is not None:
            line["other_configs"] = dict(extras or {})
            if sweep is not None:
                line["other_configs"]["param_sweep"] = sweep
        if world == 1 and not args.no_cpu_baseline:
            base, _, _ = cpu_baseline(steps=1, warmup=0, budget_s=20.0)
            base["numpy_single_core"] = numpy_single_core()
            line["cpu_baseline"] = base
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner with
    NCCL_DEBUG=VERSION or WARN), so file descriptor 1 is pointed at stderr for the whole run and the result line
    goes to a private copy of the original stdout."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=SCENARIOS_PER_GPU, help="scenarios per GPU")
    ap.add_argument("--mode", choices=["strict", "unconstrained"], default="strict")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reduced", action="store_true", help="skip the reduced-precision leg")
    ap.add_argument("--same-sweep", action="store_true", help="diagnostic: every rank runs the sweep of rank 0 (default: each rank draws its own)")
    ap.add_argument("--no-extras", action="store_true", help="skip the one-number summaries of the other BASELINE configs")
    ap.add_argument("--precision", choices=["fp64", "f16x3"], default="fp64")
    ap.add_argument("--solver", choices=["active_set", "admm"], default="active_set",
                    help="solver of the FP64 strict path (library default: active_set)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
