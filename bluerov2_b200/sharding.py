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


class PipelinedThrustGather:
    """Double-buffered ThrustGather: the all-gather of tick t runs on the collective's own stream while tick t + 1 is
    already linearising (the thrust vectors are an OUTPUT of the tick -- nothing in the next solve reads them), so the
    collective's launch + NVLink latency leaves the critical path.  Still exactly one all-gather per tick.

        thr = g.slot(t)          # this rank's block of buffer t % depth; waits (stream-side) for the gather that last read it
        solver.solve_*(..., out=(u0, thr, status))
        g.all_gather_async(t)    # enqueued behind the solve, returns at once
        ...
        full = g.result(t)       # [total, 6] of tick t, valid on the current stream from here on
    """

    def __init__(self, total: int, device, depth: int = 2, group=None):
        if depth < 1:
            raise ValueError(depth)
        self.gathers = [ThrustGather(total, device, group) for _ in range(depth)]
        self.works = [None] * depth
        self.depth = depth
        self.world = self.gathers[0].world
        self.bounds = self.gathers[0].bounds

    def _wait(self, i):
        if self.works[i] is not None:
            self.works[i].wait()          # NCCL: the current stream waits for the collective; gloo: the host does
            self.works[i] = None

    def slot(self, t: int):
        i = t % self.depth
        self._wait(i)
        return self.gathers[i].slot

    def all_gather_async(self, t: int):
        i = t % self.depth
        g = self.gathers[i]
        if g.world == 1:
            return
        if not g.equal:
            g.all_gather()                # ragged shards: the padded blocking path
            return
        self.works[i] = dist.all_gather_into_tensor(g.buf, g.slot, group=g.group, async_op=True)

    def result(self, t: int):
        i = t % self.depth
        self._wait(i)
        return self.gathers[i].buf

    def wait_all(self):
        for i in range(self.depth):
            self._wait(i)


class PeerExchangeUnavailable(RuntimeError):
    """raised on EVERY rank when any rank could not set the peer-to-peer exchange up (the caller may then run the NCCL all-gather)"""


class PeerThrustExchange:
    """The per-tick exchange of the thrust vectors without a collective (include/bluerov2_b200.h "Sharding"): every
    rank's solver owns a gather buffer [2][world * B][6]; the QP epilogue stores each instance's thrusts into the same row of
    every rank's buffer over NVLink (CUDA IPC mappings), the last warp of a tick publishes the tick index in every rank's flags.
    This class is the rendezvous: it exchanges the 64-byte IPC handles through torch.distributed (any backend: the handles are
    host bytes) and connects the peers.  Afterwards the ticks deliver by themselves; ``wait()`` enqueues the consumer-side wait,
    ``result(tick)`` views the local buffer of that tick's parity as a tensor.

    ``solver`` needs ``shard_init / shard_handle / shard_connect / shard_wait / shard_gathered_ptr`` (BatchSolver has them; the
    CPU tests pass a recording stand-in)."""

    def __init__(self, solver, batch_per_rank: int, group=None):
        self.solver, self.group = solver, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.B = int(batch_per_rank)
        self.bounds = [(r * self.B, (r + 1) * self.B) for r in range(self.world)]
        # Every step that can fail on ONE rank (no IPC in this container, no peer access between two devices) is followed by an
        # exchange of the outcome, so that all ranks raise PeerExchangeUnavailable together instead of one raising while the
        # others wait in a collective.  A solver left half connected never ships anything (engine.cu shard_pending).
        mine, err = None, None
        try:
            solver.shard_init(self.rank, self.world)
            mine = solver.shard_handle()
        except Exception as e:                          # noqa: BLE001 -- reported to every rank below
            err = f"rank {self.rank}: {type(e).__name__}: {e}"
        got = self._all_gather((mine, err))
        self._raise_if_any([g[1] for g in got], "shard_init / shard_handle")
        self.handles = [g[0] for g in got]
        try:
            for r, h in enumerate(self.handles):
                if r != self.rank:
                    solver.shard_connect(r, h)
        except Exception as e:                          # noqa: BLE001
            err = f"rank {self.rank}: {type(e).__name__}: {e}"
        self._raise_if_any(self._all_gather(err), "shard_connect")     # also the barrier: nobody ticks before everybody is mapped

    def _all_gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    @staticmethod
    def _raise_if_any(errors, step):
        bad = [e for e in errors if e]
        if bad:
            raise PeerExchangeUnavailable(f"peer-to-peer thrust exchange unavailable ({step}): " + "; ".join(bad))

    def wait(self, stream=None):
        self.solver.shard_wait(stream)

    def result(self, tick_index: int, device=None):
        """[world * B, 6] tensor view of the local gather buffer holding tick `tick_index` (1-based count of ticks solved)"""
        ptr = self.solver.shard_gathered_ptr(tick_index & 1)

        class _View:
            __cuda_array_interface__ = {"shape": (self.world * self.B, NTHRUST), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_View(), device=device if device is not None else torch.device("cuda", torch.cuda.current_device()))
