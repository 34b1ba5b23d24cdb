"""Multi-GPU plumbing for proof-level data parallelism (one process per GPU, torch.distributed).

The unit of work the reference exposes is one proof per `vm::prove` call (vm/src/lib.rs:13-29); proofs are
independent objects, so N ranks prove N disjoint sets of traces with NO data-path collective.  The only
communication is control-plane: a barrier around the timed region, a MAX over ranks of the device time and a
gather of (unit id, proof digest) pairs to rank 0.  Works with the `nccl` backend on GPUs and with `gloo` on CPU
(the latter is what the world_size-2 test uses).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Callable, Dict, List, Sequence, Tuple


def assign_units(num_units: int, world_size: int, rank: int) -> List[int]:
    """Round-robin assignment of proof units to ranks: unit u belongs to rank u % world_size."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("invalid rank / world_size")
    return list(range(rank, num_units, world_size))


def unit_seed(base_seed: int, log_n: int, unit: int) -> int:
    """Seed of the synthetic program of a unit (distinct traces on every rank; rank 0 / unit 0 is the N=1 workload)."""
    return base_seed + log_n + 1000 * unit


@dataclass
class ShardReport:
    units: Dict[int, str]          # unit id -> sha256 of the proof bytes (all ranks, on rank 0; own units elsewhere)
    seconds: float                 # max over ranks of the local time
    total_units: int               # units proved by all ranks
    throughput: float              # total_units / seconds


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def max_over_ranks(value: float, device=None) -> float:
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def prove_units(units: Sequence[int], prove_one: Callable[[int], bytes], timer: Callable[[], float], device=None) -> ShardReport:
    """Runs `prove_one(unit)` for this rank's units between two barriers and aggregates.

    `timer()` returns a monotonically increasing time in seconds (device timer on GPUs, perf_counter on CPU)."""
    dist = _dist()
    if dist is not None:
        dist.barrier()
    t0 = timer()
    mine: List[Tuple[int, str]] = []
    for u in units:
        mine.append((u, hashlib.sha256(prove_one(u)).hexdigest()))
    local = timer() - t0
    if dist is not None:
        dist.barrier()
    seconds = max_over_ranks(local, device)
    total = int(sum_over_ranks(float(len(mine)), device))
    merged: Dict[int, str] = dict(mine)
    if dist is not None and dist.get_world_size() > 1:
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, mine)
        if dist.get_rank() == 0:
            merged = {}
            for part in gathered:
                for u, h in part:
                    if u in merged:
                        raise RuntimeError(f"unit {u} was proved by two ranks")
                    merged[u] = h
    return ShardReport(merged, seconds, total, total / seconds if seconds > 0 else 0.0)
