"""bench.py places its timed region by evaluating the job-wide local/remote coins on the host (remote_plan,
place_timed_region): the plan must be the one the engine follows, i.e. the oracle's counter-mode coins
(oracle/mh_oracle.c orc_run_counter; src/mcpar.cc:142-152 for the rule)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                      # noqa: E402
from conftest import tiled_pinit  # noqa: E402
from oracle import mh             # noqa: E402


def test_remote_plan_matches_the_counter_oracle():
    for d, lag, pl in [(2, 0, 0.7), (2, 1, 0.9), (4, 1, 0.6)]:
        nburn, nsamp, sync = 30, 200, 10
        o = mh.run_counter("rosenbrock1", d, 8, nsamp, nburn, tiled_pinit(8, d), pool_m=4, pl=pl, coin_group=0, trace=True, pool_lag=lag)
        got = np.asarray(o["remote"])[nburn:, 0].astype(int)
        plan = bench.remote_plan(bench.SEED, d, nburn, nsamp, pl, sync * (1 + lag))
        assert got.tolist() == plan and 0 < sum(plan) < nsamp


def test_placement_picks_a_representative_region():
    K, W, sync, pl = 20, 3, 10, 0.9
    adv, share = bench.place_timed_region(bench.SEED, 2, 500, sync, 1, pl, 3000, W, K)
    assert 3000 <= adv <= 3200
    plan = bench.remote_plan(bench.SEED, 2, 500, (adv + W + K) * sync, pl, sync * 2)
    lo = (adv + W) * sync
    assert abs(sum(plan[lo:lo + K * sync]) / float(K * sync) - share) < 1e-12
    assert abs(share - (1.0 - pl)) <= 0.005
    assert bench.place_timed_region(bench.SEED, 2, 500, sync, 1, 1.0, 300, W, K) == (300, 0.0)      # no remote steps: nothing to place
