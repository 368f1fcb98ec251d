"""GPU tests of the C++ host mirror (mcpar_b200/host: MCPar / MCout / VLFunc classes above
the C ABI) through the driver mains: the reference's command-line, stdout, log-file and
per-rank-file contract (src/mcpar-rosen1.cc, src/mcpar-dgauss.cc, src/mcpar.cc:23-28,
:110-119; row order relied on by src/anly/mcpar-analysis.R:80-120)."""
import os
import subprocess
import numpy as np
import pytest

from conftest import tiled_pinit

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "mcpar_b200", "bin")


def _engine_rows(lik, par, d, ranks, nsamp, nburn, incov=None):
    """Same run through the Python binding, re-ordered the way MCPar::run feeds MCout:
    batches of outstep steps -> rank-major -> step-major -> chain."""
    from mcpar_b200 import engine
    N = 4 * ranks
    e = engine.Engine(d, N, mode="normal", coin_group=4, pool_m=0, history_steps=nsamp)
    e.run(nsamp, nburn, np.tile(tiled_pinit(4, d), (ranks, 1)), lik, par, incov)
    h = e.history().reshape(nsamp, ranks, 4, d + 1)
    e.close()
    outstep = nsamp // 10 if nsamp > 50 else 5
    out = []
    for s0 in range(0, nsamp, outstep):
        blk = h[s0:s0 + outstep]                         # [step][rank][chain]
        out.append(blk.transpose(1, 0, 2, 3).reshape(-1, d + 1))
    return np.concatenate(out)


def test_mcpar_rosen1_contract(tmp_path):
    nsamp, ranks = 60, 3
    r = subprocess.run([os.path.join(BIN, "mcpar-rosen1"), str(nsamp), "--ranks=%d" % ranks], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "nsamp = %d" % nsamp                              # mcpar-rosen1.cc:35-36
    rows = lines[1:]
    assert len(rows) == nsamp * 4 * ranks                                 # nsamp x nc x mpisiz rows
    assert all(l.endswith("  ") and len(l.split()) == 3 for l in rows)   # "v  v  v  "
    got = np.array([[float(t) for t in l.split()] for l in rows])
    exp = _engine_rows("rosenbrock1", None, 2, ranks, nsamp, 500)
    assert np.allclose(got, exp, rtol=2e-5, atol=1e-6)                    # 6 significant digits in the text
    log = open(tmp_path / "mcpar-log.000.txt").read().splitlines()        # mcpar.cc:56,:111-112,:116-118
    assert log[0] == "Starting burn-in.  Samples = 500"
    assert log[1] == "Starting main sample loop:  nsamp = 60" and log[2] == "Output after each 6 steps."
    assert log[3] == "Beginning output at step 6" and log[4] == "Output finished"


def test_mcpar_rosen1_default_positional_only(tmp_path):
    r = subprocess.run([os.path.join(BIN, "mcpar-rosen1"), "12"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.splitlines()[0] == "nsamp = 12"
    assert len(r.stdout.splitlines()) == 1 + 12 * 4


def test_mcpar_dgauss_contract(tmp_path):
    ranks = 2
    r = subprocess.run([os.path.join(BIN, "mcpar-dgauss"), "--ranks=%d" % ranks], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert len(lines) == 8 * 4 * ranks + 2                                # run(8, 500): rows + 2 max-likelihood lines
    assert lines[-2].startswith("max likelihood value: ")
    exp = _engine_rows("dualgaussian", [5.0], 2, ranks, 8, 500)
    got = np.array([[float(t) for t in l.split()] for l in lines[:-2]])
    assert np.allclose(got, exp, rtol=2e-5, atol=1e-6)
    best = exp[np.argmax(exp[:, 2])]
    assert np.isclose(float(lines[-2].split(":")[1]), best[2], rtol=2e-5)
    assert np.allclose([float(t) for t in lines[-1].split()], best[:2], rtol=2e-5, atol=1e-6)
    for rk in range(ranks):                                               # mcpar-dgauss.RRR.txt, parameters only, tabs
        f = open(tmp_path / ("mcpar-dgauss.%03d.txt" % rk)).read().splitlines()
        assert len(f) == 8 * 4 and all(l.endswith("\t") and len(l.split()) == 2 for l in f)


def test_mcpar_rosen2_d16(tmp_path):
    r = subprocess.run([os.path.join(BIN, "mcpar-rosen2"), "20", "--ranks=2"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "nsamp = 20" and len(lines) == 1 + 20 * 8
    got = np.array([[float(t) for t in l.split()] for l in lines[1:]])
    assert got.shape[1] == 17
    blk = (2.38 ** 2 / 16) * np.array([[0.5, 1.0], [1.0, 2.505]])
    exp = _engine_rows("rosenbrock1", None, 16, 2, 20, 500, np.kron(np.eye(8), blk))
    assert np.allclose(got, exp, rtol=2e-5, atol=1e-6)


def test_mcpar_rosen1_ngpu_equals_single_engine(tmp_path):
    """--ngpu=2: the ranks are sharded over two engines (two GPUs when the box has them, one shared
    device otherwise) that exchange their (mu, sigma^2) slots peer to peer inside the window
    kernels; stdout must be the single-engine run's, row for row."""
    nsamp, ranks = 40, 16                                                 # 8 ranks x 4 chains = 32 chains per engine
    outs = []
    for extra in ([], ["--ngpu=2"]):
        r = subprocess.run([os.path.join(BIN, "mcpar-rosen1"), str(nsamp), "--ranks=%d" % ranks] + extra, cwd=tmp_path,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert len(outs[0].splitlines()) == 1 + nsamp * 4 * ranks
    assert outs[0] == outs[1]
    r = subprocess.run([os.path.join(BIN, "mcpar-rosen1"), "10", "--ranks=3", "--ngpu=2"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 2 and "ngpu" in r.stderr                        # ranks must split evenly, 32-chain blocks


def test_binary_output_and_iteration_bookkeeping(tmp_path):
    """--binary=FILE: the same rows, same order, raw fp64; and the reference analysis script's iteration
    bookkeeping (mcparam.itercount, src/anly/mcpar-analysis.R:80-120, ported in mcpar_b200/mcout_io.py)
    names the right iteration for every row of the driver's output."""
    from mcpar_b200 import mcout_io, engine
    nsamp, ranks = 100, 3
    r = subprocess.run([os.path.join(BIN, "mcpar-rosen1"), str(nsamp), "--ranks=%d" % ranks], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    text = mcout_io.read_text(r.stdout.splitlines())
    r = subprocess.run([os.path.join(BIN, "mcpar-rosen1"), str(nsamp), "--ranks=%d" % ranks, "--binary=out.bin"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout == "nsamp = %d\n" % nsamp
    rows = mcout_io.read_binary(str(tmp_path / "out.bin"))
    assert rows.shape == (nsamp * 4 * ranks, 3) and np.allclose(text, rows, rtol=2e-5, atol=1e-6)
    # iteration of every row from the engine's own history: row (t, chain) of the history is unique
    N = 4 * ranks
    e = engine.Engine(2, N, mode="normal", coin_group=4, pool_m=0, history_steps=nsamp)
    e.run(nsamp, 500, np.tile(tiled_pinit(4, 2), (ranks, 1)), "rosenbrock1")
    h = e.history(); e.close()
    it = mcout_io.itercount(nsamp, ranks, 4)
    assert np.array_equal(it, mcout_io.itercount_exact(nsamp, ranks, 4))      # niter = 100: the R rule is exact
    chain = np.tile(np.tile(np.arange(4), mcout_io.outstep_of(nsamp)), ranks * (nsamp // mcout_io.outstep_of(nsamp)))
    rank = np.tile(np.repeat(np.arange(ranks), 4 * mcout_io.outstep_of(nsamp)), nsamp // mcout_io.outstep_of(nsamp))
    assert np.array_equal(rows, h[it - 1, rank * 4 + chain])


@pytest.mark.parametrize("remote_mode", [0, 1])
def test_user_written_host_vlfunc_runs_through_mcpar(tmp_path, remote_mode):
    """a VLFunc without a device functor is called on the host once per step (and per rank
    batch), between the engine's propose and accept kernels; the run equals the fused run with the same likelihood
    on the device."""
    host = os.path.join(ROOT, "mcpar_b200", "host")
    exe = str(tmp_path / "host_plugin_check")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-O2", "-I", host, "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "host_plugin_check.cc"), "-L", os.path.join(ROOT, "mcpar_b200"),
                           "-lmcpar", "-lmcgpu", "-Wl,-rpath," + os.path.join(ROOT, "mcpar_b200")])
    r = subprocess.run([exe, "16", "60", str(remote_mode)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr


def test_host_likelihood_steps_match_counter_oracle():
    """mcgpu_step_propose / mcgpu_step_accept with the likelihood evaluated by the CALLER (here: the oracle's
    orc_loglik on the host) against the oracle's counter-mode run: the split path uses plain fp64 arithmetic and
    CUDA libm, so the two agree to rounding -- both remote modes, lagged pool, d = 2 and a d = 6 case that no fused
    kernel is instantiated for."""
    from mcpar_b200 import engine
    from oracle import mh
    for lik, par, d, N, M, cg, rmode, lag in [("dualgaussian", [5.0], 2, 128, 8, 0, 0, 0), ("dualgaussian", [5.0], 2, 128, 16, 4, 1, 1),
                                              ("rosenbrock1", None, 6, 64, 8, 32, 1, 0)]:
        nburn, nsamp = 110, 45
        pin = tiled_pinit(N, d)
        o = mh.run_counter(lik, d, N, nsamp, nburn, pin, par=par, pool_m=M, pl=0.7, coin_group=cg, remote_mode=rmode, pool_lag=lag, trace=True)
        e = engine.Engine(d, N, mode="normal", pool_m=M, pl=0.7, coin_group=cg, history_steps=nsamp, remote_mode=rmode, pool_lag=lag)
        calls = []
        e.run_host(nsamp, nburn, pin, lambda x: (calls.append(len(x)), mh.loglik(lik, d, x, par))[1])
        h = e.history()
        same = np.all(np.isclose(h[:, :, :d], o["rows"][:, :, :d], rtol=1e-9, atol=1e-11), axis=-1)
        assert same.mean() > 0.999, (lik, d, rmode, same.mean())
        assert np.array_equal(e.factor(), o["cov"]), "burn-in tuning differs"
        s = e.stats()
        assert s["remote_steps"] == int(o["remote"][nburn:].sum()) and s["tried"] == nsamp * N
        assert abs(s["remote_iterations"] - int(o["remote_iters"][0])) <= max(2, int(o["remote_iters"][0]) // 200)
        assert len(calls) == nburn + nsamp + 1
        with pytest.raises(engine.McgpuError, match="ESTATE"):
            e.sample(1)
        e.close()


@pytest.mark.parametrize("remote_mode", [0, 1])
def test_mcpar_gmix_contract(tmp_path, remote_mode):
    """mcpar-gmix (BASELINE config 4's likelihood: d = 64, K = 64 mixture; no counterpart main in the reference, it follows
    the shape of src/mcpar-dgauss.cc): banner, rows of 64 parameters + logL for every kept step of every chain, the
    maximum-likelihood line, the log file.  Job-wide coin, so the steps run on the wide / cooperative kernels."""
    nsamp, ranks, thin = 40, 4, 10
    r = subprocess.run([os.path.join(BIN, "mcpar-gmix"), str(nsamp), "--ranks=%d" % ranks, "--thin=%d" % thin,
                        "--remote-mode=%d" % remote_mode], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "nsamp = %d" % nsamp
    rows = np.array([[float(t) for t in l.split()] for l in lines[1:]])
    assert rows.shape == ((nsamp // thin) * 4 * ranks, 65)
    assert np.isfinite(rows).all() and (rows[:, 64] < 0).all() and (np.abs(rows[:, :64]) < 12).all()   # logL of a mixture in [-5, 5]^64
    assert "max likelihood value:" in r.stderr
    assert float(r.stderr.split("max likelihood value:")[1].split()[0]) >= rows[:, 64].max() - 1e-3 * abs(rows[:, 64].max())
    assert open(tmp_path / "mcpar-log.000.txt").read().startswith("Starting burn-in.  Samples = 200")


def test_mcpar_bench_lines(tmp_path):
    """mcpar-bench (config 5's sweep through the C++ mirror): one header, one line per chain count."""
    r = subprocess.run([os.path.join(BIN, "mcpar-bench"), "--nsamp=100", "--remote-mode=1", "--lag=1", "1024", "4096"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0].startswith("# chains") and "nsamp 100" in lines[0] and "remote mode 1" in lines[0]
    data = [l.split() for l in lines[1:]]
    assert [int(d[0]) for d in data] == [1024, 4096]
    for d in data:
        assert float(d[1]) > 1e6 and 0.05 < float(d[2]) < 0.95 and float(d[3]) >= 0.0
