"""ctypes binding for oracle/_ref/libmcpar_ref{64,32}.so (TEST INFRASTRUCTURE).

The libraries are the reference's unmodified sources + shim + ref_harness.cc,
built by oracle/Makefile.  They exist wherever `make -C oracle ref` ran with
/root/reference present (the dev container) and travel to the GPU box as built
files; /root/reference itself is never read at run time.
"""
import ctypes as C
import os
import tempfile
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIK = {"rosenbrock1": 0, "rosenbrock2": 1, "gaussian": 2, "dualgaussian": 3}


class _Cfg(C.Structure):
    pass


def _cfg_fields(real):
    return [("nparam", C.c_int), ("nchain", C.c_int), ("nranks", C.c_int),
            ("nsamp", C.c_int), ("nburn", C.c_int), ("lik", C.c_int),
            ("pl", real), ("armin", real), ("armax", real), ("dfac", real), ("ifac", real),
            ("sync", C.c_int), ("rng_mode", C.c_int), ("seed", C.c_ulonglong),
            ("text_sink", C.c_int), ("trace_steps", C.c_int), ("trace_musig", C.c_int),
            ("pinit_per_rank", C.c_int)]


def lib_path(bits=64):
    return os.path.join(_HERE, "_ref", "libmcpar_ref%d.so" % bits)


def available(bits=64):
    return os.path.exists(lib_path(bits))


class Ref:
    """One loaded reference build (bits = 64: prelude fp64 build; 32: native float)."""

    def __init__(self, bits=64):
        self.bits = bits
        self.lib = C.CDLL(lib_path(bits))
        self.dtype = np.float64 if bits == 64 else np.float32
        self.creal = C.c_double if bits == 64 else C.c_float
        assert self.lib.ref_real_bytes() == bits // 8

        class Cfg(C.Structure):
            _fields_ = _cfg_fields(self.creal)
        self.Cfg = Cfg
        self.lib.ref_run.restype = C.c_double
        self.lib.ref_last_text.restype = C.c_size_t

    def _p(self, a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def loglik(self, lik, nparam, x, par=None):
        x = np.ascontiguousarray(x, dtype=self.dtype).reshape(-1, nparam)
        y = np.empty(x.shape[0], dtype=self.dtype)
        par_a = None if par is None else np.ascontiguousarray(par, dtype=self.dtype)
        rc = self.lib.ref_loglik(LIK[lik], nparam, self._p(par_a), x.shape[0], self._p(x), self._p(y))
        if rc != 0:
            raise ValueError("reference likelihood constructor/eval failed rc=%d" % rc)
        return y

    def covar_setup(self, nparam, incov=None):
        inc = None if incov is None else np.ascontiguousarray(incov, dtype=self.dtype).ravel()
        out = np.empty(nparam * nparam, dtype=self.dtype)
        self.lib.ref_covar_setup(nparam, self._p(inc), self._p(out))
        return out.reshape(nparam, nparam)

    def qriguess(self, rank, npset, nparam, plo, phi):
        plo = np.ascontiguousarray(plo, dtype=self.dtype)
        phi = np.ascontiguousarray(phi, dtype=self.dtype)
        out = np.empty(npset * nparam, dtype=self.dtype)
        self.lib.ref_qriguess(rank, npset, nparam, self._p(plo), self._p(phi), self._p(out))
        return out.reshape(npset, nparam)

    def run(self, lik, nparam, nchain, nranks, nsamp, nburn, pinit, incov=None, par=None,
            Z=None, U=None, I=None, seed=8675309, pl=0.9, armin=0.2, armax=0.5, dfac=0.2,
            ifac=1.5, sync=10, trace=False, trace_musig=False, text=False, want_rows=True,
            want_maxl=True):
        """Run MCPar::run on nranks thread-ranks.  Replay mode when Z/U(/I) given
        (arrays [nranks][n]); otherwise the shim's host Philox."""
        R, Cn, d = nranks, nchain, nparam
        dt = self.dtype
        pinit = np.ascontiguousarray(pinit, dtype=dt)
        per_rank = int(pinit.size == R * Cn * d and R > 1)
        assert pinit.size in (Cn * d, R * Cn * d)
        cfg = self.Cfg()
        cfg.nparam, cfg.nchain, cfg.nranks, cfg.nsamp, cfg.nburn = d, Cn, R, nsamp, nburn
        cfg.lik = LIK[lik]
        cfg.pl, cfg.armin, cfg.armax, cfg.dfac, cfg.ifac, cfg.sync = pl, armin, armax, dfac, ifac, sync
        replay = Z is not None
        cfg.rng_mode = 0 if replay else 1
        cfg.seed = seed
        cfg.text_sink = int(text)
        T = (nburn + nsamp) if trace else 0
        cfg.trace_steps, cfg.trace_musig, cfg.pinit_per_rank = T, int(trace_musig), per_rank
        inc = None if incov is None else np.ascontiguousarray(incov, dtype=dt).ravel()
        par_a = None if par is None else np.ascontiguousarray(par, dtype=dt)
        nz = nu = ni = 0
        if replay:
            Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(R, -1); nz = Z.shape[1]
            U = np.ascontiguousarray(U, dtype=np.float64).reshape(R, -1); nu = U.shape[1]
            if I is None:
                I = np.zeros((R, 1), dtype=np.int32)
            I = np.ascontiguousarray(I, dtype=np.int32).reshape(R, -1); ni = I.shape[1]
        nt, nc, nm = Cn * d, d * d, 2 * R * Cn * d
        out = {}
        out["rows"] = np.zeros((R, nsamp * Cn, d + 1), dtype=dt) if want_rows else None
        out["p"] = np.zeros((R, Cn, d), dtype=dt); out["ly"] = np.zeros((R, Cn), dtype=dt)
        out["mu"] = np.zeros((R, Cn, d), dtype=dt); out["sig"] = np.zeros((R, Cn, d), dtype=dt)
        out["psum2"] = np.zeros((R, Cn, d), dtype=dt); out["cov"] = np.zeros((R, d, d), dtype=dt)
        out["musig"] = np.zeros((R, R * Cn, d, 2), dtype=dt)
        out["used"] = np.zeros((R, 4), dtype=np.int64)
        out["maxl"] = np.zeros(d + 1, dtype=dt) if want_maxl else None
        tr = {}
        if T:
            tr["pre_p"] = np.zeros((R, T, Cn, d), dtype=dt); tr["pre_ly"] = np.zeros((R, T, Cn), dtype=dt)
            tr["trial_p"] = np.zeros((R, T, Cn, d), dtype=dt); tr["trial_ly"] = np.zeros((R, T, Cn), dtype=dt)
            tr["cfac"] = np.zeros((R, T, Cn), dtype=dt); tr["cov"] = np.zeros((R, T, d, d), dtype=dt)
            tr["musig"] = np.zeros((R, T, R * Cn, d, 2), dtype=dt) if trace_musig else None
            tr["cursors"] = np.zeros((R, T, 3), dtype=np.int64)
        nsteps = np.zeros(R, dtype=np.int32)
        g = lambda k: self._p(tr.get(k)) if T else None
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)      # the reference drops mcpar-log.000.txt in cwd (mcpar.cc:23-28)
            try:
                secs = self.lib.ref_run(
                    C.byref(cfg), self._p(pinit), self._p(inc), self._p(par_a),
                    self._p(Z) if replay else None, C.c_size_t(nz),
                    self._p(U) if replay else None, C.c_size_t(nu),
                    self._p(I) if replay else None, C.c_size_t(ni),
                    self._p(out["rows"]), self._p(out["p"]), self._p(out["ly"]), self._p(out["mu"]),
                    self._p(out["sig"]), self._p(out["psum2"]), self._p(out["cov"]), self._p(out["musig"]),
                    self._p(out["used"]), self._p(out["maxl"]),
                    g("pre_p"), g("pre_ly"), g("trial_p"), g("trial_ly"), g("cfac"), g("cov"),
                    g("musig"), g("cursors"), self._p(nsteps))
                log = open("mcpar-log.000.txt").read() if os.path.exists("mcpar-log.000.txt") else ""
            finally:
                os.chdir(cwd)
        if secs < 0:
            raise RuntimeError("ref_run failed")
        out["seconds"] = secs
        out["log"] = log
        if text:
            n = self.lib.ref_last_text(None, C.c_size_t(0))
            buf = C.create_string_buffer(n + 1)
            self.lib.ref_last_text(buf, C.c_size_t(n))
            out["text"] = buf.raw[:n].decode()
        if T:
            out["trace"] = tr
            out["trace_nsteps"] = nsteps
            out["accept"] = derive_accept(tr, out["p"])
        return out


def derive_accept(tr, final_p):
    """accept[r,k,j] = chain j of rank r took its trial at traced step k (state after
    the step equals the trial point)."""
    pre = tr["pre_p"]; trial = tr["trial_p"]
    R, T = pre.shape[:2]
    post = np.concatenate([pre[:, 1:], final_p[:, None]], axis=1)
    return np.all(post == trial, axis=-1)
