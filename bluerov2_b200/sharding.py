"""Multi-GPU plumbing: instances are independent, so the batch shards into contiguous blocks (one process per GPU,
no data-path collective inside the solve) and a tick ends with ONE all-gather of the thrust vectors
(SURVEY 8e).  torch.distributed only -- NCCL over NVLink on the GPUs, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist

NTHRUST = 6


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous block [lo, hi) of `total` instances owned by `rank`; the remainder goes to the first ranks"""
    if not (0 <= rank < world) or total < 0:
        raise ValueError((total, world, rank))
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ThrustGather:
    """Owns the [total, 6] fp64 buffer every rank ends a tick with.  ``slot`` is this rank's block: hand it to the
    solver as the thrust output so the IPM epilogue writes straight into the collective's send region."""

    def __init__(self, total: int, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total = total
        self.bounds = [shard_bounds(total, self.world, r) for r in range(self.world)]
        self.buf = torch.zeros((total, NTHRUST), dtype=torch.float64, device=device)
        lo, hi = self.bounds[self.rank]
        self.slot = self.buf[lo:hi]
        self.equal = len({hi - lo for lo, hi in self.bounds}) == 1

    def all_gather(self):
        """in place; enqueued on the current stream for NCCL"""
        if self.world == 1:
            return self.buf
        if self.equal:
            dist.all_gather_into_tensor(self.buf, self.slot, group=self.group)
        else:
            # ragged shards: gather fixed-size padded blocks (collectives want equal sizes), then unpack
            m = max(hi - lo for lo, hi in self.bounds)
            send = torch.zeros((m, NTHRUST), dtype=self.buf.dtype, device=self.buf.device)
            send[: self.slot.shape[0]] = self.slot
            recv = torch.empty((self.world, m, NTHRUST), dtype=self.buf.dtype, device=self.buf.device)
            dist.all_gather_into_tensor(recv.view(self.world * m, NTHRUST), send, group=self.group)
            for r, (lo, hi) in enumerate(self.bounds):
                if r != self.rank:
                    self.buf[lo:hi] = recv[r, : hi - lo]
        return self.buf
