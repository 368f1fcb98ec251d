"""ctypes binding for oracle/libmh_oracle.so, the plain-C restatement (TEST INFRASTRUCTURE).

Same call shapes and result dictionaries as oracle.ref.Ref.run so that tests can
swap the two.  Built by `make -C oracle oracle` (also by __graft_entry__.build()).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIK = {"rosenbrock1": 0, "rosenbrock2": 1, "gaussian": 2, "dualgaussian": 3, "gaussmix": 4}
SLOT_ACCEPT = 0x10000000
SLOT_REMOTE = 0x40000000


class Cfg(C.Structure):
    _fields_ = [("nparam", C.c_int), ("nchain", C.c_int), ("nranks", C.c_int),
                ("nsamp", C.c_int), ("nburn", C.c_int), ("lik", C.c_int), ("n_lik_par", C.c_int),
                ("pl", C.c_double), ("armin", C.c_double), ("armax", C.c_double),
                ("dfac", C.c_double), ("ifac", C.c_double), ("sync", C.c_int),
                ("pinit_per_rank", C.c_int), ("trace_steps", C.c_int), ("trace_musig", C.c_int),
                ("seed", C.c_uint64), ("coin_group", C.c_int), ("pool_m", C.c_int), ("thin", C.c_int),
                ("remote_mode", C.c_int), ("pool_lag", C.c_int)]


def lib_path():
    return os.path.join(_HERE, "libmh_oracle.so")


def build(force=False):
    if force or not os.path.exists(lib_path()) or \
            os.path.getmtime(lib_path()) < os.path.getmtime(os.path.join(_HERE, "mh_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return lib_path()


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_u53.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def loglik(lik, nparam, x, par=None):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, nparam)
    y = np.empty(x.shape[0], dtype=np.float64)
    par_a = None if par is None else np.ascontiguousarray(par, dtype=np.float64)
    rc = lib().orc_loglik(LIK[lik], nparam, _p(par_a), x.shape[0], _p(x), _p(y))
    if rc != 0:
        raise ValueError("oracle likelihood rejected the arguments rc=%d" % rc)
    return y


def covar_setup(nparam, incov=None):
    inc = None if incov is None else np.ascontiguousarray(incov, dtype=np.float64).ravel()
    out = np.empty(nparam * nparam, dtype=np.float64)
    lib().orc_covar_setup(nparam, _p(inc), _p(out))
    return out.reshape(nparam, nparam)


def qriguess(rank, npset, nparam, plo, phi):
    plo = np.ascontiguousarray(plo, dtype=np.float64); phi = np.ascontiguousarray(phi, dtype=np.float64)
    out = np.empty(npset * nparam, dtype=np.float64)
    lib().orc_qriguess(rank, npset, nparam, _p(plo), _p(phi), _p(out))
    return out.reshape(npset, nparam)


def gaussmix_params(K, d, mu, s2, w):
    return np.concatenate([[float(K)], np.asarray(mu, float).ravel(), np.asarray(s2, float).ravel(),
                           np.asarray(w, float).ravel()])


def _cfg(lik, d, Cn, R, nsamp, nburn, par, pl, armin, armax, dfac, ifac, sync):
    cfg = Cfg()
    cfg.nparam, cfg.nchain, cfg.nranks, cfg.nsamp, cfg.nburn = d, Cn, R, nsamp, nburn
    cfg.lik = LIK[lik]
    cfg.n_lik_par = 0 if par is None else len(par)
    cfg.pl, cfg.armin, cfg.armax, cfg.dfac, cfg.ifac, cfg.sync = pl, armin, armax, dfac, ifac, sync
    return cfg


def run_replay(lik, nparam, nchain, nranks, nsamp, nburn, pinit, Z, U, I=None, incov=None, par=None,
               pl=0.9, armin=0.2, armax=0.5, dfac=0.2, ifac=1.5, sync=10, trace=False,
               trace_musig=False):
    """The reference engine restated, on nranks ranks, consuming replay streams
    Z,U,I of shape [nranks][n] with the reference's consumption protocol."""
    R, Cn, d = nranks, nchain, nparam
    pinit = np.ascontiguousarray(pinit, dtype=np.float64)
    assert pinit.size in (Cn * d, R * Cn * d)
    cfg = _cfg(lik, d, Cn, R, nsamp, nburn, par, pl, armin, armax, dfac, ifac, sync)
    cfg.pinit_per_rank = int(pinit.size == R * Cn * d and R > 1)
    T = (nburn + nsamp) if trace else 0
    cfg.trace_steps, cfg.trace_musig = T, int(trace_musig)
    inc = None if incov is None else np.ascontiguousarray(incov, dtype=np.float64).ravel()
    par_a = None if par is None else np.ascontiguousarray(par, dtype=np.float64)
    Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(R, -1)
    U = np.ascontiguousarray(U, dtype=np.float64).reshape(R, -1)
    if I is None:
        I = np.zeros((R, 1), dtype=np.int32)
    I = np.ascontiguousarray(I, dtype=np.int32).reshape(R, -1)
    out = {"rows": np.zeros((R, nsamp * Cn, d + 1)), "p": np.zeros((R, Cn, d)), "ly": np.zeros((R, Cn)),
           "mu": np.zeros((R, Cn, d)), "sig": np.zeros((R, Cn, d)), "psum2": np.zeros((R, Cn, d)),
           "cov": np.zeros((R, d, d)), "musig": np.zeros((R, R * Cn, d, 2)),
           "used": np.zeros((R, 4), dtype=np.int64), "maxl": np.zeros(d + 1)}
    tr = {}
    if T:
        tr = {"pre_p": np.zeros((R, T, Cn, d)), "pre_ly": np.zeros((R, T, Cn)),
              "trial_p": np.zeros((R, T, Cn, d)), "trial_ly": np.zeros((R, T, Cn)),
              "cfac": np.zeros((R, T, Cn)), "cov": np.zeros((R, T, d, d)),
              "musig": np.zeros((R, T, R * Cn, d, 2)) if trace_musig else None,
              "cursors": np.zeros((R, T, 3), dtype=np.int64),
              "accept": np.zeros((R, T, Cn), dtype=np.int32), "remote": np.zeros((R, T), dtype=np.int32),
              "iters": np.zeros((R, T), dtype=np.int32)}
    g = lambda k: _p(tr.get(k)) if T else None
    rc = lib().orc_run_replay(
        C.byref(cfg), _p(pinit), _p(inc), _p(par_a),
        _p(Z), C.c_size_t(Z.shape[1]), _p(U), C.c_size_t(U.shape[1]), _p(I), C.c_size_t(I.shape[1]),
        _p(out["rows"]), _p(out["p"]), _p(out["ly"]), _p(out["mu"]), _p(out["sig"]), _p(out["psum2"]),
        _p(out["cov"]), _p(out["musig"]), _p(out["used"]), _p(out["maxl"]),
        g("pre_p"), g("pre_ly"), g("trial_p"), g("trial_ly"), g("cfac"), g("cov"), g("musig"),
        g("cursors"), g("accept"), g("remote"), g("iters"))
    if rc != 0:
        raise RuntimeError("orc_run_replay rc=%d" % rc)
    if T:
        out["trace"] = tr
        out["accept"] = tr["accept"].astype(bool)
    return out


def run_counter(lik, nparam, nchain, nsamp, nburn, pinit, incov=None, par=None, seed=8675309,
                coin_group=32, pool_m=0, thin=1, pl=0.9, armin=0.2, armax=0.5, dfac=0.2, ifac=1.5,
                sync=10, trace=False, want_rows=True, remote_mode=0, pool_lag=0):
    """Normal-mode semantics (counter-based Philox per global chain) on N=nchain chains.
    remote_mode 0: the reference's max-mixture rejection loop; 1: normalised sum-mixture proposal."""
    N, d = nchain, nparam
    pinit = np.ascontiguousarray(pinit, dtype=np.float64)
    assert pinit.size == N * d
    cfg = _cfg(lik, d, N, 1, nsamp, nburn, par, pl, armin, armax, dfac, ifac, sync)
    cfg.seed, cfg.coin_group, cfg.pool_m, cfg.thin = seed, coin_group, pool_m, thin
    cfg.remote_mode, cfg.pool_lag = remote_mode, pool_lag
    M = pool_m if 0 < pool_m < N else N
    nkeep = (nsamp + thin - 1) // thin
    inc = None if incov is None else np.ascontiguousarray(incov, dtype=np.float64).ravel()
    par_a = None if par is None else np.ascontiguousarray(par, dtype=np.float64)
    out = {"rows": np.zeros((nkeep, N, d + 1)) if want_rows else None, "p": np.zeros((N, d)),
           "ly": np.zeros(N), "mu": np.zeros((N, d)), "psum2": np.zeros((N, d)),
           "pool": np.zeros((M, d, 2)), "counts": np.zeros(2, dtype=np.int64), "cov": np.zeros((d, d)),
           "flags": np.zeros((nburn + nsamp, N), dtype=np.uint8) if trace else None,
           "remote_iters": np.zeros(1, dtype=np.int64)}
    rc = lib().orc_run_counter(C.byref(cfg), _p(pinit), _p(inc), _p(par_a), _p(out["rows"]), _p(out["p"]),
                               _p(out["ly"]), _p(out["mu"]), _p(out["psum2"]), _p(out["pool"]),
                               _p(out["counts"]), _p(out["cov"]), _p(out["flags"]), _p(out["remote_iters"]))
    if rc != 0:
        raise RuntimeError("orc_run_counter rc=%d" % rc)
    if trace:
        out["accept"] = (out["flags"] & 1).astype(bool)
        out["remote"] = (out["flags"] & 2).astype(bool)
    return out
