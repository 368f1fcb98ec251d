"""Peer-to-peer exchange (mcgpu_p2p_*): sharded engines whose window kernels store the published
(mu, sigma^2) pool slots into each other's memory and wait on arrival counters, instead of an
all-gather call per window (the reference's MPI_Allgather(MPI_IN_PLACE), src/mcpar.cc:127-140).

Runs on ONE GPU as well: the engines of a group may share a device (the peers' exchange regions
are then ordinary device pointers); with several GPUs the engines are spread over them and the
stores travel over NVLink.  Bar: bit-identical to the single-engine run, as for the NCCL exchange.
"""
import numpy as np
import pytest

from conftest import tiled_pinit

pytestmark = pytest.mark.gpu


def _group_run(eng, lik, par, d, Cg, world, M, cg, nburn, nsamp, sync, pl, thin=1, incov=None, chunks=None, mode=0, lag=0):
    N = Cg * world
    ndev = eng.device_count()
    pin = tiled_pinit(N, d)
    es = [eng.Engine(d, Cg, mode="normal", nchain_total=N, chain0=r * Cg, pool_m=M, pl=pl, sync=sync, thin=thin,
                     coin_group=cg, history_steps=(nsamp + thin - 1) // thin, device=r % ndev, remote_mode=mode, pool_lag=lag)
          for r in range(world)]
    for r, e in enumerate(es):
        e.set_likelihood(lik, par); e.set_covariance(incov); e.set_state(pin[r * Cg:(r + 1) * Cg])
    eng.p2p_attach_local(es)
    eng.burnin_group(es, nburn)
    for e in es:
        e.sample_begin(nsamp)
    # engines driven by one host thread are fed window by window, in turn (mcgpu_sample_group): a kernel that
    # waits for a publication is never queued in front of the kernel that makes it; `chunks` exercises
    # calls that stop inside a window and calls that cross several boundaries
    t, k = 0, 0
    while t < nsamp:
        n = min(chunks[k % len(chunks)] if chunks else sync, nsamp - t)
        eng.sample_group(es, n)
        t += n; k += 1
    for e in es:
        e.synchronize()
    out = dict(p=np.concatenate([e.state()["p"] for e in es]), hist=np.concatenate([e.history() for e in es], axis=1),
               fac=[e.factor() for e in es], pool=[e.musig() for e in es],
               acc=sum(e.stats()["accepted"] for e in es), rem=sum(e.stats()["remote_steps"] for e in es))
    for e in es:
        e.close()
    one = eng.Engine(d, N, mode="normal", pool_m=M, pl=pl, sync=sync, thin=thin, coin_group=cg,
                     history_steps=(nsamp + thin - 1) // thin, remote_mode=mode, pool_lag=lag)
    one.run(nsamp, nburn, pin, lik, par, incov)
    ref = dict(p=one.state()["p"], hist=one.history(), fac=one.factor(), pool=one.musig(), acc=one.stats()["accepted"])
    one.close()
    return out, ref


@pytest.mark.parametrize("lik,par,d,Cg,world,M,cg,chunks,mode,lag", [
    ("dualgaussian", [5.0], 2, 2048, 2, 16, 0, None, 0, 0),
    ("dualgaussian", [5.0], 2, 1024, 4, 16, 0, [3, 7, 10, 4], 0, 0),       # partial windows
    ("rosenbrock1", None, 2, 1024, 2, 8, 32, None, 0, 0),                  # per-group coins: one mixed kernel per window
    ("rosenbrock1", None, 16, 256, 2, 8, 0, None, 0, 0),                   # wide kernel, waits in pool_prep
    ("rosenbrock1", None, 2, 512, 2, 0, 0, None, 0, 0),                    # every chain in the pool
    ("dualgaussian", [5.0], 2, 1024, 4, 16, 0, [3, 7, 25, 4], 0, 1),       # pool read one exchange late (4 pool buffers)
    ("dualgaussian", [5.0], 2, 2048, 2, 64, 0, None, 1, 0),                # sum-mixture remote mode
    ("rosenbrock1", None, 2, 1024, 2, 16, 32, [13, 10, 7], 1, 1),          # sum-mixture, per-group coins, lagged pool
    ("rosenbrock1", None, 16, 256, 2, 8, 0, None, 1, 1),                   # wide kernel, sum-mixture, lagged pool
])
def test_p2p_group_equals_single_engine(lik, par, d, Cg, world, M, cg, chunks, mode, lag):
    from mcpar_b200 import engine as eng
    out, ref = _group_run(eng, lik, par, d, Cg, world, M, cg, nburn=130, nsamp=70, sync=10, pl=0.7, chunks=chunks, mode=mode, lag=lag)
    assert ref["acc"] > 0 and out["rem"] > 0
    assert np.array_equal(out["p"], ref["p"])
    assert np.array_equal(out["hist"], ref["hist"])
    assert all(np.array_equal(f, ref["fac"]) for f in out["fac"])
    assert all(np.array_equal(m, ref["pool"]) for m in out["pool"])      # every GPU holds the whole pool
    assert out["acc"] == ref["acc"]


@pytest.mark.parametrize("lag", [0, 1])
def test_p2p_multi_device(lag):
    """One engine per GPU: the stores travel over NVLink; the host enqueues 40 windows per engine in turn."""
    from mcpar_b200 import engine as eng
    if eng.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(4, eng.device_count())
    out, ref = _group_run(eng, "dualgaussian", [5.0], 2, 1 << 16, world, 16, 0, nburn=120, nsamp=400, sync=10, pl=0.8,
                          thin=10, chunks=[400], lag=lag)
    assert np.array_equal(out["p"], ref["p"]) and np.array_equal(out["hist"], ref["hist"])


def test_p2p_local_engines_advance_window_by_window():
    """An engine attached with attach_local refuses a sample call that crosses an exchange boundary: queued
    behind its own later windows, the peers' launches could never run (spin-wait vs launch order)."""
    from mcpar_b200 import engine as eng
    es = [eng.Engine(2, 64, nchain_total=128, chain0=64 * r, pool_m=4, coin_group=0, history_steps=30) for r in range(2)]
    for r, e in enumerate(es):
        e.set_likelihood("rosenbrock1"); e.set_state(tiled_pinit(128, 2)[64 * r:64 * (r + 1)])
    eng.p2p_attach_local(es)
    eng.burnin_group(es, 20)
    for e in es:
        e.sample_begin(30)
    with pytest.raises(eng.McgpuError, match="ESTATE"):
        es[0].sample(11)
    eng.sample_group(es, 30)
    for e in es:
        e.synchronize()
    assert es[0].stats()["main_steps"] == 30 and es[1].stats()["main_steps"] == 30
    for e in es:
        e.close()


def test_p2p_attach_rules():
    from mcpar_b200 import engine as eng
    a = eng.Engine(2, 64, nchain_total=128, chain0=0, pool_m=4)
    b = eng.Engine(2, 64, nchain_total=128, chain0=64, pool_m=4)
    with pytest.raises(eng.McgpuError):
        eng.p2p_attach_local([b, a])                    # rank order = chain order
    eng.p2p_attach_local([a, b])
    with pytest.raises(eng.McgpuError):
        eng.p2p_attach_local([a, b])                    # once
    v = eng.Engine(2, 8, mode="verify", chains_per_rank=4)
    with pytest.raises(eng.McgpuError):
        v.p2p_export()
    h = a.p2p_export()
    assert len(h) == eng.P2P_HANDLE_BYTES
    for e in (a, b, v):
        e.close()
