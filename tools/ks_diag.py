#!/usr/bin/env python
"""Diagnostic: KS / moment checks of the sum-mixture mode at full size over several data seeds and sub-samples
(tests/test_gpu_parity.py::test_full_size_stationarity saw one KS p-value of 1e-4 on a stride-64 sub-sample)."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcpar_b200 import engine
from scipy import stats
N = 1 << 20
cdf = lambda v: (5.0 * stats.norm.cdf(v) + stats.norm.cdf(v - 5.0)) / 6.0
for seed, rmode, M, pl in [(11, 1, 16, 0.9), (12, 1, 16, 0.9), (13, 1, 16, 0.9), (11, 1, 256, 0.9), (11, 0, 16, 1.0), (14, 1, 16, 0.9)]:
    rng = np.random.default_rng(seed)
    comp = rng.random(N) < 1.0 / 6.0
    pin = rng.standard_normal((N, 2)) + 5.0 * comp[:, None]
    e = engine.Engine(2, N, mode="normal", pl=pl, pool_m=M, coin_group=0, thin=100, history_steps=3, remote_mode=rmode, seed=8675309 + seed - 11)
    e.run(300, 100, pin, "dualgaussian", [5.0])
    h = e.history()
    e.close()
    out = {"seed": seed, "mode": rmode, "M": M, "pl": pl}
    for k in (1, 2):
        x = h[k][:, 0]
        out["k%d" % k] = {"ks_all": float(stats.kstest(x, cdf).pvalue), "ks_stride64": [round(float(stats.kstest(x[j::64], cdf).pvalue), 4) for j in (0, 1, 7, 31, 32, 63)],
                          "ks_rand16k": float(stats.kstest(np.random.default_rng(5).choice(x, 16384, replace=False), cdf).pvalue),
                          "dmean": float(x.mean() - 5 / 6), "ks_start": float(stats.kstest(pin[:, 0], cdf).pvalue) if k == 1 else None}
    print(json.dumps(out), flush=True)
