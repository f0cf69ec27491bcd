"""Parameter sweeps over many walking scenarios (BASELINE config 5).

A scenario is one walk of the reference's Wieber controller (`zmp_controller.py:59-108`) with its own
CoM height h, mass m, step length, foot spread and push (F_ext, tick).  What the reference would run
as a Python loop over `MPCConfig`s is organised here for the batched engine:

* `h` changes Pu (`zmp_controller.py:169,171`) and therefore the plan: scenarios are bucketed by `h`;
* (step_length, foot_spread) only change the bound series (and its length n): one host-generated
  series per distinct pair (`CoPGenerator`, bit-identical to the reference), shared by its scenarios;
* m and F_ext enter only through the impulse dt * F_ext / m (`:106`).

Scenarios are sorted by (h, bounds variant), split into contiguous balanced shards (one per GPU), and
each shard is run bucket by bucket -- one `lipm_rollout_indexed` per CoM height, the bound series picked
per scenario from the table of variants.
The only thing that leaves a rank is the packed per-scenario outcome record.
"""

from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .config import MPCConfig
from .generators.cop_generator import CoPGenerator
from .parallel import shard_range


@dataclass
class SweepTable:
    """Column-wise scenario table, already sorted by (h_index, variant)."""

    h_values: np.ndarray          # [H] distinct CoM heights
    variants: List[Tuple[float, float]]   # distinct (step_length, foot_spread)
    h_index: np.ndarray           # [B] int32
    variant: np.ndarray           # [B] int32
    mass: np.ndarray              # [B]
    F_ext: np.ndarray             # [B]
    tick_fraction: np.ndarray     # [B] push time as a fraction of the walking phase, in [0, 1)
    scenario_id: np.ndarray       # [B] position in the caller's (unsorted) order

    def __len__(self) -> int:
        return len(self.h_index)


def make_sweep(batch: int, seed: int = 3, h_values: Sequence[float] = tuple(np.linspace(0.6, 0.9, 8)),
               step_lengths: Sequence[float] = (0.2, 0.25, 0.3, 0.35, 0.4),
               foot_spreads: Sequence[float] = (0.08, 0.1, 0.12), mass_range=(30.0, 60.0),
               force_range=(0.0, 1000.0)) -> SweepTable:
    """h on a grid, m ~ U, step_length / foot_spread on grids, F_ext ~ U."""
    rng = np.random.default_rng(seed)
    variants = [(float(s), float(f)) for s in step_lengths for f in foot_spreads]
    h_index = rng.integers(0, len(h_values), batch).astype(np.int32)
    variant = rng.integers(0, len(variants), batch).astype(np.int32)
    mass = rng.uniform(*mass_range, batch)
    force = rng.uniform(*force_range, batch)
    frac = rng.uniform(0.0, 1.0, batch)
    order = np.lexsort((variant, h_index))
    return SweepTable(np.asarray(h_values, float), variants, h_index[order], variant[order], mass[order], force[order],
                      frac[order], order.astype(np.int64))


def variant_bounds(config: MPCConfig, variants: Sequence[Tuple[float, float]]):
    """Host-side bound series per (step_length, foot_spread): list of (z_max, z_min, first_walking_tick, n)."""
    out = []
    for step_length, foot_spread in variants:
        cfg = replace(config, step_length=step_length, foot_spread=foot_spread)
        z_max, z_min, states = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
        walking = [k for k, s in enumerate(states) if getattr(s, "name", str(s)) != "STANDING"]
        first = walking[0] if walking else 0
        last = walking[-1] if walking else len(states) - 1
        out.append((np.ascontiguousarray(z_max), np.ascontiguousarray(z_min), first, last, len(z_min)))
    return out


def groups_of_shard(table: SweepTable, rank: int, world: int) -> List[Tuple[int, int, int]]:
    """(h_index, lo, hi) runs of identical h inside this rank's contiguous shard: one plan and one launch
    each (the bound series of a run's scenarios are selected per scenario from the table of variants)."""
    lo, hi = shard_range(len(table), rank, world)
    runs = []
    start = lo
    while start < hi:
        key = int(table.h_index[start])
        stop = start + 1
        while stop < hi and int(table.h_index[stop]) == key:
            stop += 1
        runs.append((key, start, stop))
        start = stop
    return runs


def push_ticks(table: SweepTable, bounds, lo: int, hi: int) -> np.ndarray:
    """Push tick inside the walking phase of each scenario's own timeline."""
    first = np.array([bd[2] for bd in bounds])[table.variant[lo:hi]]
    last = np.array([bd[3] for bd in bounds])[table.variant[lo:hi]]
    n = np.array([bd[4] for bd in bounds])[table.variant[lo:hi]]
    span = np.maximum(last - first, 1)
    tick = first + np.floor(table.tick_fraction[lo:hi] * span).astype(np.int64)
    return np.clip(tick, 0, n - 2).astype(np.int32)


def bounds_table(bounds) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Stack the variants' series into (V, n_max, 2) tables (rows past a series' end are never read:
    the engine repeats the LAST VALID row of each series, `zmp_controller.py:81-88`) + lengths (V,)."""
    n_max = max(bd[4] for bd in bounds)
    z_max = np.zeros((len(bounds), n_max, 2))
    z_min = np.zeros((len(bounds), n_max, 2))
    for v, (hi_series, lo_series, _, _, n) in enumerate(bounds):
        z_max[v, :n], z_min[v, :n] = hi_series, lo_series
    return z_max, z_min, np.array([bd[4] for bd in bounds], np.int32)


def run_shard(config: MPCConfig, table: SweepTable, rank: int = 0, world: int = 1, *, strict: bool = True,
              precision: int = 0, fall_angle_deg: float = 30.0) -> Dict[str, np.ndarray]:
    """Run this rank's shard on the current CUDA device; returns the local outcome record (device tensor,
    `parallel.RECORD_WIDTH` float64 columns) and the work done."""
    import torch

    from . import engine
    from .parallel import pack_outcomes

    bounds = variant_bounds(config, table.variants)
    z_max_t, z_min_t, lengths = bounds_table(bounds)
    lo_all, hi_all = shard_range(len(table), rank, world)
    device = torch.device("cuda", torch.cuda.current_device())
    z_max_d, z_min_d = torch.as_tensor(z_max_t).to(device), torch.as_tensor(z_min_t).to(device)
    record = torch.zeros((hi_all - lo_all, 4), dtype=torch.float64, device=device)
    solves = 0
    mode = engine.MODE_STRICT if strict else engine.MODE_UNCONSTRAINED
    for h_idx, lo, hi in groups_of_shard(table, rank, world):
        h = float(table.h_values[h_idx])
        plan = engine.get_plan(config.horizon, h, config.g, config.dt, config.Q, config.R, precision=precision)
        count = hi - lo
        zero = torch.zeros((count, 3), dtype=torch.float64, device=device)
        impulse = config.dt * table.F_ext[lo:hi] / table.mass[lo:hi]
        n_steps = lengths[table.variant[lo:hi]]
        out = plan.rollout(mode, zero, zero, z_min_d, z_max_d, n_steps=n_steps, bounds_index=table.variant[lo:hi],
                           impulse=impulse, force_tick=push_ticks(table, bounds, lo, hi), want_flags=False)
        lean = (h / config.g) * out["y_hist"][..., 2].abs().amax(dim=1)     # |CoM - ZMP| = (h/g) |acceleration|
        lean = torch.maximum(lean, (h / config.g) * out["x_hist"][..., 2].abs().amax(dim=1))
        fell = lean > h * float(np.tan(np.radians(fall_angle_deg)))
        record[lo - lo_all:hi - lo_all] = pack_outcomes(out["max_violation"], out["status"], out["iters"], fell)
        solves += int((2 * (n_steps.astype(np.int64) - 1)).sum())
    return {"record": record, "solves": solves, "range": (lo_all, hi_all)}
