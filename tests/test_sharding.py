"""world_size-2 gloo test of the multi-GPU host logic (shard bounds + the single all-gather of the thrust vectors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bluerov2_b200.sharding import PipelinedThrustGather, ThrustGather, shard_bounds

RC = 0.026546960744430276


def _alloc(u):
    return np.stack([-u[:, 0] + u[:, 1] + u[:, 3], -u[:, 0] - u[:, 1] - u[:, 3], u[:, 0] + u[:, 1] - u[:, 3],
                     u[:, 0] - u[:, 1] + u[:, 3], -u[:, 2], -u[:, 2]], axis=1) / RC


def test_shard_bounds_cover_the_batch():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        u = np.random.default_rng(0).uniform(-50, 50, (total, 4))       # every rank can regenerate the global inputs
        g = ThrustGather(total, "cpu")
        lo, hi = g.bounds[rank]
        g.slot.copy_(torch.from_numpy(_alloc(u[lo:hi])))               # stands in for the solver epilogue writing its block
        full = g.all_gather().numpy()
        q.put((rank, bool(np.array_equal(full, _alloc(u))), (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_two_rank_all_gather_gloo(total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert sorted(b for _, _, b in res) == [shard_bounds(total, 2, 0), shard_bounds(total, 2, 1)]


def _worker_pipelined(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = PipelinedThrustGather(total, "cpu", depth=2)
        lo, hi = g.bounds[rank]
        T, ok = 5, True
        us = [np.random.default_rng(t).uniform(-50, 50, (total, 4)) for t in range(T)]
        for t in range(T):
            g.slot(t).copy_(torch.from_numpy(_alloc(us[t][lo:hi])))    # tick t's epilogue writes its block of buffer t % 2
            g.all_gather_async(t)
            if t >= 1:                                                  # tick t-1 is read while tick t is in flight
                ok &= bool(np.array_equal(g.result(t - 1).numpy(), _alloc(us[t - 1])))
        ok &= bool(np.array_equal(g.result(T - 1).numpy(), _alloc(us[T - 1])))
        g.wait_all()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_two_rank_pipelined_all_gather_gloo(total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_pipelined, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res



# ---- PeerThrustExchange: the rendezvous of the peer-to-peer exchange (handles through torch.distributed, then connect) ----
class _RecordingSolver:
    """stands in for BatchSolver on a box without GPUs: records what the rendezvous does"""

    def __init__(self, rank):
        self.rank, self.calls = rank, []

    def shard_init(self, rank, world):
        self.calls.append(("init", rank, world))

    def shard_handle(self):
        return bytes([self.rank]) * 64

    def shard_connect(self, peer, handle):
        assert len(handle) == 64
        self.calls.append(("connect", peer, handle[0]))


def _peer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bluerov2_b200.sharding import PeerThrustExchange
        s = _RecordingSolver(rank)
        ex = PeerThrustExchange(s, batch_per_rank=5)
        q.put((rank, s.calls, ex.bounds, [h[0] for h in ex.handles]))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_rendezvous_gloo():
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, calls, bounds, handles in res:
        assert calls[0] == ("init", rank, world)
        # every OTHER rank is connected exactly once, with that rank's own handle
        assert sorted(calls[1:]) == [("connect", r, r) for r in range(world) if r != rank]
        assert bounds == [(0, 5), (5, 10)] and handles == [0, 1]


class _FailingSolver(_RecordingSolver):
    """rank 1 cannot map its peer (what a box without IPC / peer access looks like)"""

    def shard_connect(self, peer, handle):
        if self.rank == 1:
            raise RuntimeError("cudaIpcOpenMemHandle: invalid device context")
        super().shard_connect(peer, handle)


def _peer_fail_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bluerov2_b200.sharding import PeerThrustExchange, PeerExchangeUnavailable
        try:
            PeerThrustExchange(_FailingSolver(rank), batch_per_rank=5)
            q.put((rank, None))
        except PeerExchangeUnavailable as e:
            q.put((rank, str(e)))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_failure_on_one_rank_raises_on_all_gloo():
    """one rank failing to connect must not leave the others waiting in a collective: every rank gets the same exception"""
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_fail_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r for r, _ in res] == [0, 1]
    for _, msg in res:
        assert msg is not None and "shard_connect" in msg and "rank 1" in msg and "cudaIpcOpenMemHandle" in msg


def test_peer_exchange_single_rank_needs_no_process_group():
    from bluerov2_b200.sharding import PeerThrustExchange
    s = _RecordingSolver(0)
    ex = PeerThrustExchange(s, batch_per_rank=3)
    assert s.calls == [("init", 0, 1)] and ex.world == 1 and ex.bounds == [(0, 3)]
