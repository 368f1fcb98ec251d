#!/usr/bin/env python
"""Diagnostic for test_full_size_stationarity's sub-sample KS: p-values of EVERY stride-64 offset, both
coordinates, both kept steps, for several configurations; the p-values of a correct sampler are uniform."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcpar_b200 import engine
from scipy import stats
N = 1 << 20
cdf = lambda v: (5.0 * stats.norm.cdf(v) + stats.norm.cdf(v - 5.0)) / 6.0
for seed, rmode, M, pl, kseed in [(11, 1, 16, 0.9, 0), (11, 1, 16, 0.9, 1), (12, 1, 16, 0.9, 0), (11, 0, 16, 1.0, 0), (11, 1, 256, 0.9, 0)]:
    rng = np.random.default_rng(seed)
    comp = rng.random(N) < 1.0 / 6.0
    pin = rng.standard_normal((N, 2)) + 5.0 * comp[:, None]
    e = engine.Engine(2, N, mode="normal", pl=pl, pool_m=M, coin_group=0, thin=100, history_steps=3, remote_mode=rmode, seed=8675309 + kseed)
    e.run(300, 100, pin, "dualgaussian", [5.0])
    h = e.history()
    e.close()
    out = {"data_seed": seed, "key_seed": kseed, "mode": rmode, "M": M, "pl": pl}
    P = np.zeros((2, 2, 64))
    for ki, k in enumerate((1, 2)):
        for i in (0, 1):
            for j in range(64):
                P[ki, i, j] = stats.kstest(h[k][j::64, i], cdf).pvalue
    out["min_p"] = float(P.min()); out["argmin(k,i,off)"] = [int(v) for v in np.unravel_index(P.argmin(), P.shape)]
    out["n_below_0.01"] = int((P < 0.01).sum()); out["n_tests"] = int(P.size)
    out["uniformity_of_p"] = float(stats.kstest(P.ravel(), "uniform").pvalue)
    out["p_offset0"] = [[round(float(P[ki, i, 0]), 5) for i in (0, 1)] for ki in (0, 1)]
    out["start_p_offset0"] = [round(float(stats.kstest(pin[0::64, i], cdf).pvalue), 5) for i in (0, 1)]
    out["ks_all"] = [[round(float(stats.kstest(h[k][:, i], cdf).pvalue), 4) for i in (0, 1)] for k in (1, 2)]
    # per-lane means of x1 at kept step 1 (chain index mod 128): any lane-dependent defect shows as an outlier
    x = h[1][:, 1]
    lm = x.reshape(-1, 128).mean(axis=0) - 5.0 / 6.0
    se = np.sqrt((1.0 + 25.0 * 5.0 / 36.0) / (N / 128))
    out["lane_mean_max_sigma"] = float(np.abs(lm / se).max()); out["lane_of_max"] = int(np.abs(lm).argmax())
    print(json.dumps(out), flush=True)
