"""Multi-process check (run under torchrun on >= 2 GPUs; tests/test_gpu_dist.py launches it):
chains sharded one engine per rank, tuning counters all-reduced over NCCL, the pool exchanged
either by an in-place NCCL all-gather per window or by the window kernels themselves with
peer-to-peer stores over NVLink (mcgpu_p2p_*), must reproduce the single-engine run bit for bit."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import tiled_pinit                        # noqa: E402
from mcpar_b200 import engine                           # noqa: E402
from mcpar_b200.sharded import DistGroup, Shard, ShardedRunner, check_even_pool   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    def gather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    cases = [("dualgaussian", [5.0], 2, 4096, 16, 0), ("rosenbrock1", None, 2, 2048, 8, 32), ("rosenbrock1", None, 16, 512, 8, 0)]
    runs = [(c, x, 0, 0) for x in ("nccl", "p2p") for c in cases]
    # remote mode 1 (sum-mixture proposal) with the pool read one exchange late (4 pool buffers), both exchanges
    runs += [(c, "p2p", 1, 1) for c in cases] + [(cases[0], "nccl", 1, 1)]
    for (lik, par, d, Cg, M, cg), xchg, rmode, lag in runs:
        N, nburn, nsamp, sync, pl = Cg * world, 130, 60, 10, 0.7
        check_even_pool(Shard(rank, world, Cg), M)
        pin = tiled_pinit(N, d)
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        e = engine.Engine(d, Cg, mode="normal", nchain_total=N, chain0=rank * Cg, pool_m=M, pl=pl, sync=sync,
                          coin_group=cg, history_steps=nsamp, device=local, remote_mode=rmode, pool_lag=lag)
        e.set_stream(stream.cuda_stream)
        e.set_likelihood(lik, par); e.set_covariance(None); e.set_state(pin[rank * Cg:(rank + 1) * Cg])
        r = ShardedRunner(e, DistGroup(dist), lambda ptr: torch.as_tensor(ptr, device=dev))
        if xchg == "p2p":
            r.enable_p2p(rank, world, gather_bytes)
        r.burnin(nburn)
        r.sample(nsamp, sync)
        e.synchronize(); torch.cuda.synchronize()
        mine = e.state()["p"]; hist = e.history(); fac = e.factor()
        dist.barrier()                                  # peers map this engine's exchange region until here
        e.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, hist, fac))
        if rank == 0:
            one = engine.Engine(d, N, mode="normal", pool_m=M, pl=pl, sync=sync, coin_group=cg, history_steps=nsamp, device=local,
                                remote_mode=rmode, pool_lag=lag)
            one.run(nsamp, nburn, pin, lik, par)
            same = (np.array_equal(np.concatenate([g[0] for g in gathered]), one.state()["p"])
                    and np.array_equal(np.concatenate([g[1] for g in gathered], axis=1), one.history())
                    and all(np.array_equal(g[2], one.factor()) for g in gathered))
            print("dist_check", xchg, lik, "d=%d" % d, "world=%d" % world, "remote_mode=%d lag=%d" % (rmode, lag),
                  "OK" if same else "MISMATCH", flush=True)
            ok = ok and same
            one.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
