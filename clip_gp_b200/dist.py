"""Host-side sharding plumbing (torch.distributed; NCCL on the B200 box, gloo in the CPU tests) — SURVEY.md 8e.

The data path has NO collective inside the kernels: eval images and training MC samples are independent, so every
rank runs the single-GPU kernels on its shard and only tiny results cross NVLink:
  * training: ONE all-reduce (sum) of the flat gradient buffer (+ loss in its last slot) per step;
  * evaluation: ONE all-reduce of the integer calibration counters (exact, order independent) and, for AECE only,
    an all-gather of the per-image (confidence, hit) pairs because equal-count bins need global ranks.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as td


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items: ceil(n / world) per rank, the tail ranks may be short or empty."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def sample_split(S: int, rank: int, world: int) -> Tuple[int, int]:
    """(offset, count) of this rank's MC samples: even split, the first S % world ranks take one more
    (S=10 over 8 ranks: 2,2,1,1,1,1,1,1).  Offsets index the shared Philox stream, so sharding never changes the draws."""
    base, extra = divmod(S, world)
    return rank * base + min(rank, extra), base + (1 if rank < extra else 0)


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def gather_variable(t: torch.Tensor, counts: List[int]) -> torch.Tensor:
    """All-gather 1-D shards of known (possibly unequal) lengths into the global order."""
    if not (td.is_available() and td.is_initialized()) or td.get_world_size() == 1:
        return t
    world = td.get_world_size()
    mx = max(counts)
    if min(counts) == mx:                                   # equal shards (the usual case): ONE collective into one tensor, no padding
        out = torch.empty(world * mx, dtype=t.dtype, device=t.device)
        td.all_gather_into_tensor(out, t.contiguous())
        return out
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    bufs = [torch.empty(mx, dtype=t.dtype, device=t.device) for _ in range(world)]
    td.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def global_calibration(hist: torch.Tensor, conf: torch.Tensor, correct: torch.Tensor, n_total: int, world: int):
    """Combine per-rank eval results: returns (hist summed over ranks, conf and correct gathered in image order)."""
    if world == 1:
        return hist, conf, correct
    h = hist.clone()
    allreduce_sum_(h)
    counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    return h, gather_variable(conf, counts), gather_variable(correct, counts)
