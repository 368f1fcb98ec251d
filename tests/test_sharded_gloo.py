"""world_size-2 gloo test of the sharded-run host logic (mcpar_b200/sharded.py): the
in-place pool all-gather with the engine's slice rule and the burn-in tuning all-reduce,
driven through ShardedRunner with a CPU stand-in for the engine."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcpar_b200.sharded import DistGroup, Shard, ShardedRunner, check_even_pool, pool_slots


class _Buf:
    def __init__(self, t):
        self.t = t; self.ptr = t.data_ptr(); self.nbytes = t.numel() * 8


class FakeEngine:
    """Mimics the C-ABI engine's sharded protocol on CPU tensors: same tuning boundaries
    (isamp = 51, 101, ...), same pool slice rule, two pool buffers swapped at exchange_end."""

    def __init__(self, shard, pool_m, d, sync):
        self.sh, self.d, self.sync = shard, d, sync
        self.s0, self.s1, self.M, self.stride = pool_slots(shard.chain0, shard.chains_per_rank, shard.total, pool_m)
        self.pool = [torch.zeros(self.M * d * 2, dtype=torch.float64) for _ in range(2)]
        self.cur = 0
        self.cnt = torch.zeros(2, dtype=torch.int64)
        self.burn_done, self.irate, self.pending_tune = 0, 50, False
        self.t, self.pending_x = 0, False
        self.tuned = []

    def tuning_counters(self):
        return _Buf(self.cnt)

    def burnin_some(self, nmax):
        assert not self.pending_tune
        n = min(nmax, self.irate + 2 - self.burn_done)
        self.burn_done += n
        self.cnt += torch.tensor([n * (self.sh.rank + 1), n * self.sh.chains_per_rank])
        self.pending_tune = self.burn_done == self.irate + 2
        return n, self.pending_tune

    def tune(self):
        self.tuned.append(self.cnt.clone()); self.cnt.zero_(); self.irate += 50; self.pending_tune = False

    def sample_begin(self, nsamp):
        self.t = 0

    def sample(self, n):
        assert not self.pending_x and (self.t % self.sync) + n <= self.sync
        self.t += n
        nxt = self.pool[self.cur ^ 1]
        for s in range(self.s0, self.s1):                         # publish own slots
            nxt[s * self.d * 2:(s + 1) * self.d * 2] = float(1000 * self.t + s)
        self.pending_x = self.t % self.sync == 0

    def exchange_begin(self):
        return _Buf(self.pool[self.cur ^ 1]), self.s0 * self.d * 16, (self.s1 - self.s0) * self.d * 16

    def exchange_end(self):
        self.cur ^= 1; self.pending_x = False


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = Shard(rank, world, 64)
        pool_m, d, sync = 16, 2, 10
        check_even_pool(sh, pool_m)
        e = FakeEngine(sh, pool_m, d, sync)
        views = {}

        def as_tensor(buf):                                       # alias, not copy
            return buf.t
        r = ShardedRunner(e, DistGroup(dist), as_tensor)
        r.burnin(130)
        r.sample(35, sync)
        q.put((rank, [t.tolist() for t in e.tuned], e.pool[e.cur].tolist(), e.t, e.burn_done))
    finally:
        dist.destroy_process_group()


def test_sharded_runner_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, tuned, pool, t, burn in res:
        assert t == 35 and burn == 130
        # two tuning boundaries inside 130 steps (after 52 and 102 steps); counters are GLOBAL sums
        assert tuned == [[52 * 1 + 52 * 2, 52 * 64 * 2], [50 * 1 + 50 * 2, 50 * 64 * 2]]
        # the current pool is the one all-gathered at t = 30: every slot from its owner
        exp = np.repeat([1000.0 * 30 + s for s in range(16)], 4)
        assert pool == exp.tolist()


def test_pool_slot_rule():
    assert pool_slots(0, 1 << 20, 1 << 23, 256) == (0, 32, 256, 1 << 15)
    assert pool_slots(3 << 20, 1 << 20, 1 << 23, 256) == (96, 128, 256, 1 << 15)
    assert pool_slots(0, 64, 64, 0) == (0, 64, 64, 1)
    with pytest.raises(ValueError):
        check_even_pool(Shard(0, 3, 64), 16)


class FakeP2PEngine(FakeEngine):
    """The same stand-in with the peer-to-peer protocol of mcgpu_p2p_*: handles are exported and
    attached in rank order, after which sample() may cross window boundaries and the runner must
    not call the exchange entry points."""
    p2p = False

    def p2p_export(self):
        return bytes([self.sh.rank]) * 64

    def p2p_attach(self, world, rank, handles):
        assert world == self.sh.world and rank == self.sh.rank and len(handles) == world
        assert [h[0] for h in handles] == list(range(world)) and all(len(h) == 64 for h in handles)
        self.p2p = True

    def sample(self, n):
        assert self.p2p
        self.t += n

    def exchange_begin(self):
        raise AssertionError("the peer-to-peer exchange needs no host-side exchange call")


def _worker_p2p(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e = FakeP2PEngine(Shard(rank, world, 64), 16, 2, 10)
        r = ShardedRunner(e, DistGroup(dist), lambda buf: buf.t)

        def gather_bytes(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        r.enable_p2p(rank, world, gather_bytes)
        r.burnin(60)
        r.sample(35, 10)
        q.put((rank, e.p2p, e.t, [t.tolist() for t in e.tuned]))
    finally:
        dist.destroy_process_group()


def test_sharded_runner_p2p_wiring_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker_p2p, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, p2p, t, tuned in res:
        assert p2p and t == 35
        assert tuned == [[52 * 1 + 52 * 2, 52 * 64 * 2]]          # burn-in tuning still sums the counters over ranks
