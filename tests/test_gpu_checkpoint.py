"""Checkpoint / restart (mcgpu_checkpoint_*): an engine rebuilt from a checkpoint continues the run
bit for bit -- mid burn-in (tuning state), mid main loop inside an exchange window, thread-per-chain
and wide kernels.  (The reference has no resume)"""
import numpy as np
import pytest

from conftest import tiled_pinit

pytestmark = pytest.mark.gpu


def _mk(eng, d, N, M, cg, thin, nsamp):
    return eng.Engine(d, N, mode="normal", pool_m=M, pl=0.7, coin_group=cg, thin=thin, history_steps=(nsamp + thin - 1) // thin)


@pytest.mark.parametrize("lik,par,d,N,M,cg,thin", [
    ("dualgaussian", [5.0], 2, 4096, 16, 0, 1),
    ("rosenbrock1", None, 2, 1000, 8, 8, 3),          # ragged chain count, per-group coins, thinned history
    ("rosenbrock1", None, 16, 128, 8, 0, 1),          # wide kernel
])
def test_restart_continues_bit_for_bit(lik, par, d, N, M, cg, thin):
    from mcpar_b200 import engine as eng
    nburn, nsamp, cut_b, cut_s = 130, 75, 70, 33       # cuts: inside a tuning window / inside an exchange window
    pin = tiled_pinit(N, d)
    a = _mk(eng, d, N, M, cg, thin, nsamp)
    a.run(nsamp, nburn, pin, lik, par)
    ref = dict(p=a.state(), hist=a.history(), fac=a.factor(), pool=a.musig(), st=a.stats())
    a.close()

    b = _mk(eng, d, N, M, cg, thin, nsamp)
    b.set_likelihood(lik, par); b.set_covariance(None); b.set_state(pin)
    b.burnin(cut_b)
    blob1 = b.checkpoint()
    b.close()
    c = _mk(eng, d, N, M, cg, thin, nsamp)
    c.set_likelihood(lik, par)
    c.restore(blob1)
    c.burnin(nburn - cut_b)
    c.sample_begin(nsamp); c.sample(cut_s)
    head = c.history()                                 # rows kept before the second checkpoint
    blob2 = c.checkpoint()
    c.close()
    e = _mk(eng, d, N, M, cg, thin, nsamp)
    e.set_likelihood(lik, par)
    e.restore(blob2)
    e.sample(nsamp - cut_s); e.synchronize()
    st = e.state()
    for k in ("p", "ly", "mu", "psum2"):
        assert np.array_equal(st[k], ref["p"][k]), k
    assert np.array_equal(e.factor(), ref["fac"]) and np.array_equal(e.musig(), ref["pool"])
    nh = head.shape[0]
    assert np.array_equal(head, ref["hist"][:nh])
    # the history is not part of the blob: after a load the device serves only the rows produced since
    assert np.array_equal(e.history(first=nh), ref["hist"][nh:])
    with pytest.raises(eng.McgpuError, match="no longer on the device"):
        e.history(first=0, count=1)
    s = e.stats()
    assert (s["accepted"], s["tried"], s["remote_steps"], s["remote_iterations"]) == \
        (ref["st"]["accepted"], ref["st"]["tried"], ref["st"]["remote_steps"], ref["st"]["remote_iterations"])
    e.close()


def test_checkpoint_rejects_a_different_engine():
    from mcpar_b200 import engine as eng
    a = eng.Engine(2, 256, pool_m=8)
    a.set_likelihood("rosenbrock1"); a.set_state(tiled_pinit(256, 2)); a.burnin(10)
    blob = a.checkpoint()
    b = eng.Engine(2, 512, pool_m=8)
    with pytest.raises(eng.McgpuError):
        b.restore(blob)                                # likelihood first
    b.set_likelihood("rosenbrock1")
    with pytest.raises(eng.McgpuError, match="truncated|different shape"):
        b.restore(blob)                                # an engine of another size
    d = eng.Engine(2, 256, pool_m=8, seed=7)
    d.set_likelihood("rosenbrock1")
    with pytest.raises(eng.McgpuError, match="different shape"):
        d.restore(blob)                                # same size, another seed
    d.close()
    c = eng.Engine(2, 256, pool_m=8)
    c.set_likelihood("rosenbrock1")
    with pytest.raises(eng.McgpuError):
        c.restore(blob[:100])
    c.restore(blob)
    for e in (a, b, c):
        e.close()
