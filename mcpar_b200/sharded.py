"""Chains sharded over several engines (one engine per GPU / per process).

Replaces what the reference does with MPI between ranks (src/mcpar.cc:127-140 in-place
MPI_Allgather of the (mu, sigma^2) table; per-rank tuning counters) by collectives on
torch tensors that ALIAS the engines' device buffers.  The collectives are injected so
the same code runs over NCCL (one process per GPU, torch.distributed), over gloo on CPU
tensors (tests), or between several engines inside one process (LocalGroup).

Partition rule: contiguous blocks of global chain id, g = rank*C + j -- the reference's
musigall slot rule (mcpar.cc:206, :342).
"""
from dataclasses import dataclass


@dataclass
class Shard:
    rank: int
    world: int
    chains_per_rank: int

    @property
    def chain0(self):
        return self.rank * self.chains_per_rank

    @property
    def total(self):
        return self.world * self.chains_per_rank


def pool_slots(chain0, nchain, ntotal, pool_m):
    """Pool slots [s0, s1) owned by chains [chain0, chain0+nchain): slot s is global chain
    s*stride, stride = ntotal // M (same rule as the kernel's publication)."""
    M = pool_m if 0 < pool_m < ntotal else ntotal
    stride = ntotal // M
    s0 = (chain0 + stride - 1) // stride
    s1 = min(M, (chain0 + nchain + stride - 1) // stride)
    return s0, max(s0, s1), M, stride


def check_even_pool(shard, pool_m):
    """An in-place all-gather needs equal slices: every rank must own M/world slots."""
    s = [pool_slots(r * shard.chains_per_rank, shard.chains_per_rank, shard.total, pool_m) for r in range(shard.world)]
    sizes = {b - a for a, b, _, _ in s}
    if len(sizes) != 1 or s[0][0] != 0 or any(s[i][1] != s[i + 1][0] for i in range(len(s) - 1)):
        raise ValueError("pool_m=%d does not split evenly over %d ranks of %d chains" %
                         (pool_m, shard.world, shard.chains_per_rank))
    return s[shard.rank]


class DistGroup:
    """torch.distributed collectives (nccl on GPU tensors, gloo on CPU tensors)."""

    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group

    def all_gather_inplace(self, full, start, count):
        # send buffer is the rank's own slice of the receive buffer: NCCL's in-place form,
        # and exactly MPI_IN_PLACE as the reference uses it
        self.dist.all_gather_into_tensor(full, full[start:start + count], group=self.group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, group=self.group)


class ShardedRunner:
    """Drives ONE engine of a sharded run; every rank executes the same sequence."""

    def __init__(self, engine, group, as_tensor, cuda_stream=None):
        """engine: mcpar_b200.engine.Engine created with nchain < nchain_total.
        group: DistGroup-like.  as_tensor: DevicePtr -> tensor aliasing that memory.
        The collectives below run on torch's CURRENT stream and touch buffers the engine's kernels write
        (the next pool, the tuning counters), so the engine is put on that stream here (`cuda_stream`: a
        cudaStream_t as int, default torch.cuda.current_stream()): launches and collectives are then ordered."""
        self.e, self.g, self.as_tensor = engine, group, as_tensor
        self._views = {}
        self._cnt = None
        if hasattr(engine, "set_stream"):
            if cuda_stream is None:
                try:
                    import torch
                    if torch.cuda.is_available():
                        cuda_stream = torch.cuda.current_stream().cuda_stream
                except ImportError:
                    pass
            if cuda_stream is not None:
                engine.set_stream(cuda_stream)

    def burnin(self, nburn):
        """Burn-in with GLOBAL acceptance-rate tuning: the window counters are summed over
        all ranks before each tuning decision, so the proposal scale -- and therefore every
        chain's trajectory -- does not depend on how chains are sharded."""
        if self._cnt is None:
            self._cnt = self.as_tensor(self.e.tuning_counters())
        left = nburn
        while left > 0:
            done, pending = self.e.burnin_some(left)
            left -= done
            if pending:
                self.g.all_reduce_sum(self._cnt)
                self.e.tune()

    def enable_p2p(self, rank, world, all_gather_bytes):
        """Switch the exchange from an all-gather call per window to in-kernel peer-to-peer stores
        (mcgpu_p2p_*): export this engine's exchange-region handle, gather all handles in rank
        order with `all_gather_bytes(bytes) -> [bytes]*world` (any host-side means), attach."""
        handles = all_gather_bytes(self.e.p2p_export())
        self.e.p2p_attach(world, rank, handles)

    def exchange(self):
        if getattr(self.e, "p2p", False):   # the window kernels exchanged over NVLink themselves
            return
        buf, off, own = self.e.exchange_begin()
        if buf.ptr not in self._views:
            self._views[buf.ptr] = self.as_tensor(buf)
        self.g.all_gather_inplace(self._views[buf.ptr], off // 8, own // 8)
        self.e.exchange_end()

    def sample(self, nsamp, sync):
        """Main loop in exchange windows of `sync` steps (mcpar.cc:113-210)."""
        self.e.sample_begin(nsamp)
        t = 0
        while t < nsamp:
            n = min(sync, nsamp - t)
            self.e.sample(n)
            t += n
            if t % sync == 0:
                self.exchange()

    def window(self, sync):
        self.e.sample(sync)
        self.exchange()
