#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref = the reference's own
unmodified sources built against the shim MPI/MKL headers by oracle/Makefile).

Run in the dev container, where /root/reference exists:   python tests/golden/make_golden.py
The fixtures let boxes without oracle/_ref still pin oracle/mh_oracle.c (and through it the
CUDA engine) to the reference's behaviour.  Inputs are regenerated from seeds
(tests/conftest.py: make_streams / tiled_pinit), so only outputs are stored.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_streams, tiled_pinit          # noqa: E402
from oracle.ref import Ref                               # noqa: E402

SPD4 = np.diag([0.5, 2, 0.5, 2.0]) + 0.1
# name: (lik, d, C, R, nsamp, nburn, par, incov, pl, sync, seed)
CASES = {
    "rosen1_c1": ("rosenbrock1", 2, 4, 3, 120, 160, None, None, 0.9, 10, 101),     # BASELINE configs[0] shape, small
    "rosen1_single": ("rosenbrock1", 2, 4, 1, 150, 120, None, None, 0.9, 10, 102),
    "dgauss": ("dualgaussian", 2, 4, 2, 100, 120, [5.0], None, 0.9, 10, 103),
    "gauss": ("gaussian", 2, 8, 2, 60, 110, [1.0, -1.0, 0.5, 2.0], None, 0.9, 10, 104),
    "rosen2_d4": ("rosenbrock2", 4, 4, 2, 60, 110, None, None, 0.9, 10, 105),
    "rosen1_d4_cov": ("rosenbrock1", 4, 5, 2, 60, 320, None, SPD4, 0.9, 10, 106),
    "rosen1_sync3": ("rosenbrock1", 2, 4, 2, 55, 0, None, None, 0.5, 3, 107),
    "rosen1_local": ("rosenbrock1", 2, 24, 1, 60, 170, None, None, 1.0, 10, 108),
}


def run_case(ref, case):
    lik, d, C, R, nsamp, nburn, par, incov, pl, sync, seed = case
    Z, U, I = make_streams(R, C, d, nsamp + nburn, seed)
    return ref.run(lik, d, C, R, nsamp, nburn, tiled_pinit(C, d), incov=incov, par=par, Z=Z, U=U, I=I,
                   pl=pl, sync=sync, trace=True, text=True)


def main():
    ref = Ref(64)
    for name, case in CASES.items():
        o = run_case(ref, case)
        tr = o["trace"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            rows=o["rows"], p=o["p"], ly=o["ly"], mu=o["mu"], sig=o["sig"], psum2=o["psum2"],
                            cov=o["cov"], musig=o["musig"], used=o["used"], maxl=o["maxl"],
                            accept=np.packbits(o["accept"]), accept_shape=np.array(o["accept"].shape),
                            trial_ly=tr["trial_ly"], cfac=tr["cfac"], cursors=tr["cursors"],
                            text_head=np.array(o["text"][:2000]), log=np.array(o["log"]))
        print(name, "rows", o["rows"].shape, "accept rate %.3f" % o["accept"].mean())
    # likelihood known-answer vectors through the reference's own VLFunc classes
    rng = np.random.default_rng(99)
    kat = {}
    for lik, d, par in [("rosenbrock1", 2, None), ("rosenbrock1", 16, None), ("rosenbrock2", 4, None),
                        ("gaussian", 2, [1.0, -1.0, 0.5, 2.0]), ("dualgaussian", 2, [5.0])]:
        x = rng.normal(1.0, 2.0, size=(64, d))
        kat["x_%s_%d" % (lik, d)] = x
        kat["y_%s_%d" % (lik, d)] = ref.loglik(lik, d, x, par)
    kat["chol_in"] = np.array([[0.5, 1.0], [1.0, 2.505]])
    kat["chol_out"] = ref.covar_setup(2, kat["chol_in"])
    kat["qri"] = ref.qriguess(2, 5, 3, [0, -1, 2.0], [1, 1, 4.0])
    np.savez_compressed(os.path.join(HERE, "likelihood_kat.npz"), **kat)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
