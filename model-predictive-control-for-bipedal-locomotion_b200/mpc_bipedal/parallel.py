"""Scenario sharding across GPUs and the one collective of the path.

Scenarios are independent for a whole walk (reference loop `zmp_controller.py:93-106` has no
cross-scenario term), so a sweep shards by contiguous scenario ranges, one process per GPU, with no
data-path collective.  The only exchange is an all-gather of a packed per-scenario outcome record
(max ZMP violation, fall flag, solver status, iterations) once the walks are done; histories stay
on the rank that produced them.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""

from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

RECORD_WIDTH = 4   # float64 columns: max_violation, fell, status, iters


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most 1) range of scenarios owned by ``rank``."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_outcomes(max_violation: torch.Tensor, status: torch.Tensor, iters: torch.Tensor,
                  fell: torch.Tensor) -> torch.Tensor:
    """(B,2) per-axis kernel outputs + (B,) fall flags -> (B,4) float64 record (exact for these ranges)."""
    return torch.stack([
        max_violation.amax(dim=1).to(torch.float64),
        fell.to(torch.float64),
        status.sum(dim=1).to(torch.float64),
        iters.sum(dim=1).to(torch.float64),
    ], dim=1).contiguous()


def unpack_outcomes(record: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {
        "max_violation": record[:, 0],
        "fell": record[:, 1] != 0,
        "status": record[:, 2].to(torch.int64),
        "iters": record[:, 3].to(torch.int64),
    }


def gather_outcomes(local: torch.Tensor, batch: int, rank: int, world: int) -> torch.Tensor:
    """All-gather the shards' records into the (batch, 4) table, in global scenario order."""
    if world == 1:
        return local
    sizes = [hi - lo for lo, hi in (shard_range(batch, r, world) for r in range(world))]
    width = max(sizes)
    padded = torch.zeros((width, RECORD_WIDTH), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
