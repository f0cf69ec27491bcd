"""Scenario sharding across GPUs and the one collective of the path.

Scenarios are independent for a whole walk (reference loop `zmp_controller.py:93-106` has no
cross-scenario term), so a sweep shards by scenario, one process per GPU, with no data-path
collective.  The only exchange is an all-gather of the packed per-scenario outcome record
(`lipm_outcome_t` of include/lipm_mpc.h: max ZMP violation, max lean, solver status, fall flag,
iterations -- 32 bytes, written by the library's `k_pack_outcomes` straight into the send buffer)
once the walks are done; histories stay on the rank that produced them.  Backend: NCCL on GPUs,
gloo in the CPU tests.
"""

from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

RECORD_WIDTH = 4   # int64 words of lipm_outcome_t: max_violation (f64 bits), lean_max (f64 bits), status | fell << 32, iters


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most 1) range of scenarios owned by ``rank``."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_outcomes(max_violation: torch.Tensor, status: torch.Tensor, iters: torch.Tensor,
                  fell: torch.Tensor, lean_max: torch.Tensor = None) -> torch.Tensor:
    """Host-side twin of the library's ``k_pack_outcomes`` (used by the CPU tests and by callers that hold per-axis
    columns): (B,2) per-axis results + (B,) fall flags -> (B,4) int64 ``lipm_outcome_t`` records, bit for bit."""
    viol = max_violation.amax(dim=1).to(torch.float64).contiguous()
    lean = torch.zeros_like(viol) if lean_max is None else lean_max.amax(dim=1).to(torch.float64).contiguous()
    word = status.sum(dim=1).to(torch.int64) | (fell.to(torch.int64) << 32)
    return torch.stack([viol.view(torch.int64), lean.view(torch.int64), word, iters.sum(dim=1).to(torch.int64)],
                       dim=1).contiguous()


def unpack_outcomes(record: torch.Tensor) -> Dict[str, torch.Tensor]:
    """(B,4) int64 records -> named columns."""
    word = record[:, 2]
    return {
        "max_violation": record[:, 0].contiguous().view(torch.float64),
        "lean_max": record[:, 1].contiguous().view(torch.float64),
        "status": word & 0xFFFFFFFF,
        "fell": (word >> 32) != 0,
        "iters": record[:, 3],
    }


def gather_outcomes(local: torch.Tensor, batch: int, rank: int, world: int) -> torch.Tensor:
    """All-gather the contiguous shards' records into the (batch, 4) table, in global scenario order."""
    if world == 1:
        return local
    sizes = [hi - lo for lo, hi in (shard_range(batch, r, world) for r in range(world))]
    width = max(sizes)
    padded = torch.zeros((width, RECORD_WIDTH), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def gather_by_index(local: torch.Tensor, index: torch.Tensor, batch: int, world: int) -> torch.Tensor:
    """All-gather for a NON-contiguous partition (e.g. round-robin inside parameter buckets): every rank contributes its
    records and the global scenario index of each; the result is the (batch, 4) table in global order."""
    if world == 1:
        table = torch.empty((batch, RECORD_WIDTH), dtype=local.dtype, device=local.device)
        table[index] = local
        return table
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.empty_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    width = int(max(c.item() for c in counts))
    send = torch.zeros((width, RECORD_WIDTH + 1), dtype=torch.int64, device=local.device)
    send[: local.shape[0], :RECORD_WIDTH] = local
    send[: local.shape[0], RECORD_WIDTH] = index.to(torch.int64)
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send)
    table = torch.empty((batch, RECORD_WIDTH), dtype=local.dtype, device=local.device)
    for part, c in zip(parts, counts):
        k = int(c.item())
        table[part[:k, RECORD_WIDTH]] = part[:k, :RECORD_WIDTH]
    return table
