"""Multi-rank host logic on CPU (gloo, world_size 2): scenario sharding and the outcome all-gather.

The data path has no collective: ranks own contiguous scenario ranges and only exchange the packed
per-scenario outcomes at the end.  The GPU kernels are not involved here."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_bipedal.parallel import gather_by_index, gather_outcomes, pack_outcomes, shard_range, unpack_outcomes


def test_shard_ranges_cover_batch_exactly():
    for B in (0, 1, 7, 4096, 1_048_576 + 3):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(B, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_outcome_record_round_trip():
    rng = np.random.default_rng(0)
    viol = torch.tensor(rng.random((9, 2)))
    status = torch.tensor(rng.integers(0, 3, (9, 2)), dtype=torch.int32)
    iters = torch.tensor(rng.integers(0, 50000, (9, 2)), dtype=torch.int32)
    fell = torch.tensor(rng.random(9) < 0.3)
    lean = torch.tensor(rng.random((9, 2)))
    rec = pack_outcomes(viol, status, iters, fell, lean)
    assert rec.shape == (9, 4) and rec.dtype == torch.int64              # lipm_outcome_t: 32 bytes per scenario
    out = unpack_outcomes(rec)
    assert torch.equal(out["max_violation"], viol.amax(dim=1))
    assert torch.equal(out["lean_max"], lean.amax(dim=1))
    # the record is the C struct of include/lipm_mpc.h, bit for bit
    from mpc_bipedal.engine import OUTCOME_DTYPE, unpack_outcome_host
    assert OUTCOME_DTYPE.itemsize == 32
    host = unpack_outcome_host(rec.numpy())
    assert np.array_equal(host["max_violation"], viol.amax(dim=1).numpy()) and np.array_equal(host["fell"], fell.numpy())
    assert np.array_equal(host["status"], status.sum(dim=1).numpy()) and np.array_equal(host["iters"], iters.sum(dim=1).numpy())
    assert torch.equal(out["status"], status.sum(dim=1).to(torch.int64))
    assert torch.equal(out["iters"], iters.sum(dim=1).to(torch.int64))
    assert torch.equal(out["fell"], fell)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B, rank, world)
    # each rank "solves" its scenarios: outcome = f(global scenario index)
    idx = torch.arange(lo, hi, dtype=torch.float64)
    viol = torch.stack([idx * 1e-3, idx * 2e-3], dim=1)
    status = torch.stack([(idx % 3 == 0).to(torch.int32), torch.zeros_like(idx, dtype=torch.int32)], dim=1)
    iters = torch.stack([idx.to(torch.int32), 2 * idx.to(torch.int32)], dim=1)
    fell = idx % 5 == 0
    local = pack_outcomes(viol, status, iters, fell)
    everything = gather_outcomes(local, B, rank, world)
    torch.save(everything, os.path.join(result_dir, f"rank{rank}.pt"))
    # the same outcomes, dealt round-robin instead (parameter sweeps): rank r owns scenarios r, r + world, ...
    mine = torch.arange(rank, B, world)
    f = mine.to(torch.float64)
    rr = pack_outcomes(torch.stack([f * 1e-3, f * 2e-3], dim=1),
                       torch.stack([(f % 3 == 0).to(torch.int32), torch.zeros_like(f, dtype=torch.int32)], dim=1),
                       torch.stack([f.to(torch.int32), 2 * f.to(torch.int32)], dim=1), f % 5 == 0)
    torch.save(gather_by_index(rr, mine, B, world), os.path.join(result_dir, f"rr{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [10, 11])
def test_all_gather_of_outcomes_world2(tmp_path, B):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    assert torch.equal(got[0], got[1])                       # every rank ends with the same table
    out = unpack_outcomes(got[0])
    idx = torch.arange(B, dtype=torch.float64)
    assert torch.equal(out["max_violation"], idx * 2e-3)     # scenario order preserved across ranks
    assert torch.equal(out["iters"], (3 * idx).to(torch.int64))
    assert torch.equal(out["fell"], idx % 5 == 0)
    assert torch.equal(out["status"], (idx % 3 == 0).to(torch.int64))
    rr = [torch.load(os.path.join(tmp_path, f"rr{r}.pt")) for r in range(world)]
    assert torch.equal(rr[0], got[0]) and torch.equal(rr[1], got[0])     # round-robin partition: same table
