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


# ---------------------------------------------------------------------------------------------------- NVLink peer memory
class _CudaArray:
    """Minimal __cuda_array_interface__ carrier: lets torch view memory the library allocated with cudaMalloc."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerBlock:
    """One cudaMalloc block per rank, mapped into every other rank's address space with CUDA IPC (one process per GPU on ONE node;
    torch.distributed carries the 64-byte handles once).  `views(q)` are rank q's regions as seen from this process: raw device
    pointers that kernels on THIS GPU may load / store through NVLink (csrc/peer.cu).  `local(name)` is a torch view of this rank's
    own region, so the rest of the engine (kernel argument blocks, optimiser state) uses it like any other tensor.

    layout: {name: (offset_bytes, nbytes)}; identical on every rank."""

    def __init__(self, layout, device: torch.device, rank: int, world: int, group=None):
        """Collective over the group.  Raises RuntimeError ON EVERY RANK if any rank could not allocate, export or map a block (CUDA IPC
        unavailable in the container, peers on another node, ...), so that the caller can fall back to NCCL consistently."""
        from . import _lib
        import ctypes as C
        if world > _lib.PEER_MAX:
            raise ValueError(f"PeerBlock: world {world} > {_lib.PEER_MAX}")
        self.lib, self.rank, self.world, self.layout, self.device = _lib.load(), rank, world, dict(layout), device
        self.nbytes = max(o + b for o, b in layout.values())
        self.base, self._mapped, self._bytes, self.bases = 0, [], None, [0] * world
        err, handle = None, C.create_string_buffer(64)
        try:
            import os
            if os.environ.get("CLIPGP_PEER_DISABLE_IPC"):             # test hook: exercise the collective fall-back
                raise RuntimeError("CUDA IPC disabled by CLIPGP_PEER_DISABLE_IPC")
            base = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(self.lib.clipgp_peer_alloc(self.nbytes, C.byref(base)), "peer_alloc")
                self.base = int(base.value)
                _lib.check(self.lib.clipgp_ipc_export(base, handle), "ipc_export")
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        got = [None] * world
        torch.distributed.all_gather_object(got, (err, bytes(handle.raw)), group=group)
        errs = [g[0] for g in got if g[0]]
        if not errs:
            try:
                with torch.cuda.device(device):
                    for q in range(world):
                        if q == rank:
                            self.bases[q] = self.base
                            continue
                        ptr = C.c_void_p()
                        _lib.check(self.lib.clipgp_ipc_open(got[q][1], C.byref(ptr)), f"ipc_open(rank {q})")
                        self.bases[q] = int(ptr.value)
                        self._mapped.append(int(ptr.value))
            except Exception as e:  # noqa: BLE001
                err = f"rank {rank}: {e}"
            got2 = [None] * world
            torch.distributed.all_gather_object(got2, err, group=group)     # also the barrier: nobody touches a block before all are mapped
            errs = [g for g in got2 if g]
        if errs:
            self.close()
            raise RuntimeError("PeerBlock: CUDA IPC peer mapping failed (" + "; ".join(errs) + ")")
        self._carrier = _CudaArray(self.base, self.nbytes)
        self._bytes = torch.as_tensor(self._carrier, device=device)

    def local(self, name: str, dtype: torch.dtype) -> torch.Tensor:
        o, b = self.layout[name]
        return self._bytes[o:o + b].view(dtype)

    def ptr(self, q: int, name: str) -> int:
        return self.bases[q] + self.layout[name][0]

    def close(self):
        """Unmap the peers' blocks and free this rank's (call on every rank, after a barrier: peers may still be reading)."""
        import ctypes as C
        if self.lib is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._mapped:
                self.lib.clipgp_ipc_close(C.c_void_p(p))
            self._mapped = []
            self._bytes = None
            if self.base:
                self.lib.clipgp_peer_free(C.c_void_p(self.base))
            self.base = 0
        self.lib = None
