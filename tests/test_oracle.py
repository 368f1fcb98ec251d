"""CPU tests pinning the oracle: the plain-C restatement (oracle/mh_oracle.c) against
(1) golden fixtures generated from the reference's own sources (tests/golden/make_golden.py),
(2) the live reference build oracle/_ref when it is present, (3) analytic known answers
(SURVEY.md section 4), (4) Philox4x32-10 known-answer vectors (Random123)."""
import importlib.util
import os
import numpy as np
import pytest

from conftest import make_streams, tiled_pinit
from oracle import mh
from oracle import ref as refmod

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = make_golden.CASES


def _oracle_case(case, trace=True):
    lik, d, C, R, nsamp, nburn, par, incov, pl, sync, seed = case
    Z, U, I = make_streams(R, C, d, nsamp + nburn, seed)
    return mh.run_replay(lik, d, C, R, nsamp, nburn, tiled_pinit(C, d), Z, U, I, incov=incov, par=par,
                         pl=pl, sync=sync, trace=trace)


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_matches_golden_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    o = _oracle_case(CASES[name])
    for k in ("rows", "p", "ly", "mu", "sig", "psum2", "cov", "musig", "used", "maxl"):
        assert np.array_equal(o[k], g[k], equal_nan=True), k       # bit-identical to the reference
    acc = np.unpackbits(g["accept"])[:int(np.prod(g["accept_shape"]))].reshape(g["accept_shape"]).astype(bool)
    assert np.array_equal(o["accept"], acc)
    assert np.array_equal(o["trace"]["trial_ly"], g["trial_ly"], equal_nan=True)
    assert np.array_equal(o["trace"]["cfac"], g["cfac"])
    assert np.array_equal(o["trace"]["cursors"], g["cursors"])
    assert o["used"][:, 3].sum() == 0                                # no stream overrun


@pytest.mark.skipif(not refmod.available(64), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["rosen1_c1", "dgauss", "rosen2_d4", "rosen1_sync3"])
def test_restatement_matches_live_reference_build(name):
    r = refmod.Ref(64)
    a = make_golden.run_case(r, CASES[name])
    b = _oracle_case(CASES[name])
    for k in ("rows", "p", "ly", "mu", "sig", "psum2", "cov", "musig", "used", "maxl"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    for k in ("pre_p", "pre_ly", "trial_p", "trial_ly", "cfac", "cov", "cursors"):
        assert np.array_equal(a["trace"][k], b["trace"][k], equal_nan=True), k
    assert np.array_equal(a["accept"], b["accept"])


@pytest.mark.skipif(not refmod.available(32), reason="oracle/_ref not built")
def test_fp64_build_tracks_native_fp32_reference():
    """The prelude (float->double) build is the same algorithm as the reference's native
    float build: same accept/reject decisions while trajectories stay within fp32 rounding."""
    lik, d, C, R, nsamp, nburn = "rosenbrock1", 2, 4, 1, 0, 40
    Z, U, I = make_streams(R, C, d, 60, 5)
    a = refmod.Ref(64).run(lik, d, C, R, nsamp, nburn, tiled_pinit(C, d), Z=Z, U=U, I=I, trace=True)
    b = refmod.Ref(32).run(lik, d, C, R, nsamp, nburn, tiled_pinit(C, d), Z=Z, U=U, I=I, trace=True)
    assert (a["accept"] == b["accept"]).mean() > 0.97
    assert np.allclose(a["trace"]["trial_ly"][:, :5], b["trace"]["trial_ly"][:, :5], rtol=2e-4, atol=1e-3)


def test_text_output_row_format():
    """MCout::output format (mcout.cc:37-47): `value  value  value  \\n`, 6 significant digits."""
    g = np.load(os.path.join(GOLD, "rosen1_c1.npz"))
    first = str(g["text_head"]).splitlines()[0]
    toks = first.split("  ")
    assert toks[-1] == "" and len(toks) == 4
    assert np.allclose([float(t) for t in toks[:3]], g["rows"][0, 0], rtol=1e-5, atol=1e-6)
    log = str(g["log"]).splitlines()
    assert log[0] == "Starting burn-in.  Samples = 160"
    assert log[1] == "Starting main sample loop:  nsamp = 120" and log[2] == "Output after each 12 steps."


def test_likelihood_known_answers():
    kat = np.load(os.path.join(GOLD, "likelihood_kat.npz"))
    for lik, d, par in [("rosenbrock1", 2, None), ("rosenbrock1", 16, None), ("rosenbrock2", 4, None),
                        ("gaussian", 2, [1.0, -1.0, 0.5, 2.0]), ("dualgaussian", 2, [5.0])]:
        y = mh.loglik(lik, d, kat["x_%s_%d" % (lik, d)], par)
        assert np.array_equal(y, kat["y_%s_%d" % (lik, d)]), lik
    assert np.array_equal(mh.loglik("rosenbrock1", 2, [[1, 1], [0, 0], [2, 2]]), [0.0, -1.0, -401.0])
    assert np.isclose(mh.loglik("dualgaussian", 2, [[0, 0]], [5.0])[0], np.log(5 + np.exp(-25)), rtol=1e-15)
    assert mh.loglik("dualgaussian", 2, [[100, 100]], [5.0])[0] == -np.inf
    assert np.array_equal(mh.covar_setup(2, kat["chol_in"]), kat["chol_out"])
    L = np.tril(kat["chol_out"])
    assert np.allclose(L @ L.T, kat["chol_in"])
    assert np.array_equal(mh.qriguess(2, 5, 3, [0, -1, 2.0], [1, 1, 4.0]), kat["qri"])
    with pytest.raises(ValueError):
        mh.loglik("rosenbrock1", 3, np.zeros((1, 3)))              # rosenbrock.hh:13-16


def test_rosenbrock2_quirks():
    """rosenbrock.cc:25-41: the flat loop credits chain j with a term that pairs its LAST
    parameter with chain j+1's FIRST; the last chain of a batch has d-1 terms; sign is
    t1^2 - 100 t2^2."""
    x = np.array([[1.0, 2.0], [3.0, 4.0]])
    t = lambda a, b: (1 - a) ** 2 - 100 * (b - a * a) ** 2
    y = mh.loglik("rosenbrock2", 2, x)
    assert np.isclose(y[0], -(t(1, 2) + t(2, 3))) and np.isclose(y[1], -t(3, 4))


def test_gaussmix_reduces_to_dualgaussian():
    rng = np.random.default_rng(1)
    x = rng.normal(2, 3, size=(200, 2))
    par = mh.gaussmix_params(2, 2, [[0, 0], [5, 5]], np.ones((2, 2)), [5.0, 1.0])
    assert np.allclose(mh.loglik("gaussmix", 2, x, par), mh.loglik("dualgaussian", 2, x, [5.0]), rtol=1e-12, atol=1e-12)


def test_philox_known_answers():
    """Philox4x32-10 KAT vectors from Random123 (kat_vectors)."""
    assert mh.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert mh.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert mh.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sobol_first_points():
    """Dimension 1 is the van der Corput sequence in gray-code order; all points in [0,1)."""
    q = mh.qriguess(0, 8, 1, [0.0], [1.0]).ravel()
    assert np.allclose(q, [0, 0.5, 0.75, 0.25, 0.375, 0.875, 0.625, 0.125])
    q3 = mh.qriguess(0, 64, 3, [0, 0, 0], [1, 1, 1])
    assert q3.min() >= 0 and q3.max() < 1 and np.allclose(q3.mean(0), 0.5, atol=0.02)
    # rank skip-ahead (mcutil.cc:22-23): rank r continues where rank r-1's npset*nparam scalars end
    a = mh.qriguess(0, 10, 3, [0, 0, 0], [1, 1, 1]); b = mh.qriguess(1, 5, 3, [0, 0, 0], [1, 1, 1])
    assert np.array_equal(b, a[5:])


def test_counter_mode_local_moments_rosenbrock():
    """Normal-mode semantics, local proposals only: analytic Rosenbrock moments
    (x ~ N(1,1/2), E y = 1.5, Var y = 2.505, Cov = 1)."""
    N = 256                                           # the banana mixes slowly: 20000 steps, drop the first 5000
    o = mh.run_counter("rosenbrock1", 2, N, 20000, 500, tiled_pinit(N, 2), pl=1.0, thin=10)
    r = o["rows"][500:].reshape(-1, 3)
    assert abs(r[:, 0].mean() - 1.0) < 0.02 and abs(r[:, 1].mean() - 1.5) < 0.05
    assert abs(r[:, 0].var() - 0.5) < 0.03 and abs(r[:, 1].var() - 2.505) < 0.3
    assert abs(np.cov(r[:, 0], r[:, 1])[0, 1] - 1.0) < 0.08


def test_counter_mode_pool_equal_all_chains_when_m_zero():
    """pool_m = 0 keeps every chain in the remote mixture (the reference's choice)."""
    N = 64
    o = mh.run_counter("rosenbrock1", 2, N, 40, 60, tiled_pinit(N, 2), pool_m=0, trace=True)
    assert o["pool"].shape == (N, 2, 2) and o["remote"][60 + 10:].any() and not o["remote"][:70].any()


# ---------------------------------------------------------------- counter mode: remote mode 1, pool lag, Sobol
def test_sobol_points_match_scipy():
    """orc_sobol_points (Joe-Kuo direction numbers, gray-code order, 32-bit integers) against scipy's
    unscrambled Sobol generator -- an independent implementation over the same public table."""
    import ctypes as C
    from scipy.stats import qmc
    lib = mh.lib()
    for d in (1, 2, 16, 17, 64):
        n = 2048
        out = np.empty(n * d)
        assert lib.orc_sobol_points(d, C.c_uint64(0), C.c_size_t(n * d), out.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(out.reshape(n, d), qmc.Sobol(d, scramble=False, bits=32).random(n))
    skip = np.empty(3 * 5)
    lib.orc_sobol_points(5, C.c_uint64(7 * 5), C.c_size_t(15), skip.ctypes.data_as(C.c_void_p))
    assert np.array_equal(skip.reshape(3, 5), qmc.Sobol(5, scramble=False, bits=32).random(10)[7:])   # rank skip-ahead, mcutil.cc:22-23


def test_sum_mixture_mode_leaves_the_target_invariant():
    """Remote mode 1 (x' ~ uniform mixture of the pool's Gaussians, Hastings factor q(x)/q(x') with normalised
    components) is an exact independence sampler: chains started from the DualGaussian target stay there --
    mean and small-mode mass within 4.5 standard errors -- while the reference's max-mixture correction with
    unnormalised Q_i (mode 0, mcpar.cc:367-439) visibly drifts at the same size."""
    from scipy import stats
    N, nsamp = 16384, 300
    rng = np.random.default_rng(11)
    comp = rng.random(N) < 1.0 / 6.0
    pin = rng.standard_normal((N, 2)) + 5.0 * comp[:, None]
    pright = (5.0 * stats.norm.sf(2.5) + stats.norm.cdf(2.5)) / 6.0
    var = 1.0 + 25.0 * 5.0 / 36.0
    se_m, se_p = np.sqrt(var / N), np.sqrt(pright * (1.0 - pright) / N)
    for lag in (0, 1):
        o = mh.run_counter("dualgaussian", 2, N, nsamp, 100, pin, par=[5.0], pool_m=16, pl=0.9, coin_group=0, thin=100,
                           remote_mode=1, pool_lag=lag, trace=True)
        h = o["rows"][-1]
        assert abs(h[:, 0].mean() - 5.0 / 6.0) < 4.5 * se_m and abs((h[:, 0] > 2.5).mean() - pright) < 4.5 * se_p
        rem = o["remote"][100:]
        assert not rem[:10 * (1 + lag)].any() and rem[10 * (1 + lag):].any()     # remote steps start at t = sync (1 + lag)
        assert o["remote_iters"][0] == rem.sum()                                   # one candidate per remote step
        assert o["accept"][100:][rem].mean() > 0.5                                # and most of them are accepted
    o0 = mh.run_counter("dualgaussian", 2, N, nsamp, 100, pin, par=[5.0], pool_m=16, pl=0.9, coin_group=0, thin=100)
    assert abs((o0["rows"][-1][:, 0] > 2.5).mean() - pright) > 3.0 * se_p         # the reference algorithm's bias


def test_pool_lag_changes_nothing_without_remote_steps():
    pin = tiled_pinit(64, 2)
    a = mh.run_counter("rosenbrock1", 2, 64, 40, 60, pin, pool_m=8, pl=1.0, coin_group=0)
    b = mh.run_counter("rosenbrock1", 2, 64, 40, 60, pin, pool_m=8, pl=1.0, coin_group=0, pool_lag=1, remote_mode=1)
    assert np.array_equal(a["rows"], b["rows"]) and np.array_equal(a["pool"], b["pool"])
