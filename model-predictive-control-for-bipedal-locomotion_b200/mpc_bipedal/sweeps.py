"""Parameter sweeps over many walking scenarios (BASELINE config 5).

A scenario is one walk of the reference's Wieber controller (`zmp_controller.py:59-108`) with its own
CoM height h, mass m, step length, foot spread and push (F_ext, tick).  What the reference would run
as a Python loop over `MPCConfig`s is organised here for the batched engine:

* `h` changes Pu (`zmp_controller.py:169,171`) and therefore the plan: scenarios are bucketed by `h`;
* (step_length, foot_spread) only change the bound series (and its length n): one series per distinct
  pair, generated ON THE DEVICE by `lipm_generate_bounds` (bit-identical to the reference's
  `CoPGenerator`, tested) and shared by its scenarios through `lipm_rollout_ex`'s bounds index;
* m and F_ext enter only through the impulse dt * F_ext / m (`:106`).

Scenarios are sorted by (h, bounds variant).  Inside every h bucket the scenarios are dealt ROUND-ROBIN
to the ranks, so every GPU runs the same number of launches on the same mix of CoM heights, footstep
plans and pushes (a contiguous split gives each rank different heights, whose iteration counts differ:
measured 0.69 weak-scaling efficiency at 8 GPUs in round 1).  The only thing that leaves a rank is the
packed per-scenario outcome record the library writes (`lipm_outcome_t`).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .config import MPCConfig


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


def series_capacity(config: MPCConfig, variants: Sequence[Tuple[float, float]]) -> int:
    """Upper bound on the samples of any variant's bound series (`cop_generator.py:73-115`: two standing phases,
    one double + one single support per footstep of `footstep_generator.py:19-49`, a few samples of clock slack)."""
    steps = max(int(np.ceil(config.distance / s)) for s, _ in variants) + 4
    seconds = 2.0 * config.standing_duration + steps * (config.ssp_duration + config.dsp_duration)
    return int(np.ceil(seconds / config.dt)) + 2 * steps + 16


def device_bounds(config: MPCConfig, variants: Sequence[Tuple[float, float]]):
    """Bound series of every variant on the current device + their lengths and walking-phase spans (host)."""
    from . import engine

    step = np.array([v[0] for v in variants])
    spread = np.array([v[1] for v in variants])
    table = engine.generate_bounds(config, np.full(len(variants), config.distance), step, spread,
                                   n_max=series_capacity(config, variants))
    n = table["n"].cpu().numpy()
    if (n <= 0).any():
        raise RuntimeError(f"bound series capacity too small: {n.tolist()}")
    states = table["states"].cpu().numpy()
    first = np.zeros(len(variants), np.int64)
    last = np.zeros(len(variants), np.int64)
    for v in range(len(variants)):
        walking = np.flatnonzero(states[v, : n[v]] != 0)          # 0 = STANDING
        first[v] = walking[0] if len(walking) else 0
        last[v] = walking[-1] if len(walking) else n[v] - 1
    return table["z_max"], table["z_min"], n.astype(np.int32), first, last


def bucket_ranges(table: SweepTable) -> List[Tuple[int, int, int]]:
    """(h_index, lo, hi) runs of identical CoM height in the sorted table."""
    edges = np.flatnonzero(np.diff(table.h_index)) + 1
    starts = np.concatenate([[0], edges]).astype(np.int64)
    stops = np.concatenate([edges, [len(table)]]).astype(np.int64)
    return [(int(table.h_index[a]), int(a), int(b)) for a, b in zip(starts, stops) if b > a]


def rank_indices(table: SweepTable, rank: int, world: int) -> List[Tuple[int, np.ndarray]]:
    """Per h bucket, the (sorted-table) indices this rank runs: every `world`-th scenario of the bucket."""
    return [(h, np.arange(lo + rank, hi, world, dtype=np.int64)) for h, lo, hi in bucket_ranges(table)]


def push_ticks(table: SweepTable, idx: np.ndarray, n: np.ndarray, first: np.ndarray, last: np.ndarray) -> np.ndarray:
    """Push tick inside the walking phase of each scenario's own timeline."""
    v = table.variant[idx]
    span = np.maximum(last[v] - first[v], 1)
    tick = first[v] + np.floor(table.tick_fraction[idx] * span).astype(np.int64)
    return np.clip(tick, 0, n[v] - 2).astype(np.int32)


def run_shard(config: MPCConfig, table: SweepTable, rank: int = 0, world: int = 1, *, strict: bool = True,
              precision: int = 0, fall_angle_deg: float = 30.0) -> Dict[str, object]:
    """Run this rank's share on the current CUDA device.  Returns `record` (count, 4) int64 device tensor of
    `lipm_outcome_t`, `index` (count,) int64 device tensor with the sorted-table index of each record, `solves`
    (QP solves of the walks that ran to their end; a walk that stopped at an unconverged QP counts for nothing) and
    `launched` (scenarios run)."""
    import torch

    from . import engine

    device = torch.device("cuda", torch.cuda.current_device())
    z_max_d, z_min_d, lengths, first, last = device_bounds(config, table.variants)
    mode = engine.MODE_STRICT if strict else engine.MODE_UNCONSTRAINED
    shares = rank_indices(table, rank, world)
    total = sum(len(idx) for _, idx in shares)
    record = torch.empty((total, 4), dtype=torch.int64, device=device)
    index = torch.empty((total,), dtype=torch.int64, device=device)
    n_all = np.zeros(total, np.int64)
    pos = 0
    # Reduced precision: one stream per CoM-height bucket -- a launch is as long as its slowest unit, and the other buckets'
    # launches fill that tail (0.83 -> 0.65 s for 131,072 scenarios).  The FP64 active-set kernel keeps per-stage feedback
    # data of every resident column in scratch memory; eight buckets in flight push it out of the L2 (measured: 0.90 ->
    # 1.20 s), so its buckets run one after the other.
    concurrent = precision == engine.PRECISION_F16X3
    main = torch.cuda.current_stream(device)
    streams = []
    for h_idx, idx in shares:
        count = len(idx)
        if count == 0:
            continue
        h = float(table.h_values[h_idx])
        plan = engine.get_plan(config.horizon, h, config.g, config.dt, config.Q, config.R, precision=precision)
        variant = table.variant[idx]
        n_steps = lengths[variant]
        side = torch.cuda.Stream(device) if concurrent else main
        if concurrent:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            zero = torch.zeros((count, 3), dtype=torch.float64, device=device)
            out = plan.rollout(mode, zero, zero, z_min_d, z_max_d, n_steps=n_steps, bounds_index=variant,
                               impulse=config.dt * table.F_ext[idx] / table.mass[idx],
                               force_tick=push_ticks(table, idx, lengths, first, last), want_flags=False, want_hist=False,
                               fall_lean=h * float(np.tan(np.radians(fall_angle_deg))))
            record[pos:pos + count] = out["outcome"]
            index[pos:pos + count] = torch.as_tensor(idx).to(device, non_blocking=True)
        streams.append((side, out, zero))
        n_all[pos:pos + count] = n_steps
        pos += count
    for side, _, _ in streams:
        if concurrent:
            main.wait_stream(side)
    status = (record[:, 2] & 0xFFFFFFFF).cpu().numpy()
    solves = int((2 * (n_all - 1))[status == 0].sum())
    return {"record": record, "index": index, "solves": solves, "launched": total}
