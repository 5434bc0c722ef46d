"""Host-side multi-GPU logic: reads shard, the index is replicated, results are gathered on the host.

There is no collective on the data path.  One process per GPU (torchrun) uses
``torch.distributed`` only for plumbing: a barrier around the timed region, a max-over-ranks of the step
time, and -- when a caller wants the whole batch's intervals in one place -- a gather of the per-rank CSR
results to rank 0 with the offsets fixed up exactly as smem_gpu_collect does inside one multi-GPU handle
(csrc/smem_gpu.cu: shard(), do_fetch()).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_reads: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank ``rank``: ceil(n/world) reads each, same rule as the C library."""
    per = (n_reads + world - 1) // world
    return min(n_reads, rank * per), min(n_reads, (rank + 1) * per)


def shard_batch(seq: np.ndarray, offs: np.ndarray, rank: int, world: int):
    lo, hi = shard_range(len(offs) - 1, rank, world)
    o = offs[lo:hi + 1]
    return seq[int(o[0]):int(o[-1])], o - o[0], (lo, hi)


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: int, device=None) -> int:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(value)
    t = torch.tensor([value], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def gather_rows(values, device=None) -> "list[list[float]]":
    """One row of floats per rank, on every rank (per-rank stage times of the end-to-end leg in bench.py)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [[float(v) for v in values]]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [[float(v) for v in o.tolist()] for o in out]


def gather_results(local: dict, dst: int = 0):
    """Gather per-rank CSR results (intv [t,4] uint64, read_off [n+1], optional step) on ``dst`` in rank
    (= read) order.  Returns the merged dict on ``dst`` and None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [None] * world if rank == dst else None
    dist.gather_object({k: (np.asarray(v) if v is not None else None) for k, v in local.items()}, parts, dst=dst)
    if rank != dst:
        return None
    intv = np.concatenate([p["intv"] for p in parts])
    offs = [np.zeros(1, np.int64)]
    base = 0
    for p in parts:
        offs.append(p["read_off"][1:] + base)
        base += int(p["read_off"][-1])
    out = dict(intv=intv, read_off=np.concatenate(offs))
    if all(p.get("step") is not None for p in parts):
        out["step"] = np.concatenate([p["step"] for p in parts])
    return out
