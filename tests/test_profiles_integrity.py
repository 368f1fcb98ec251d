"""The roofline inputs bench.py reads (profiles/fp64_work.json: fp64 flops per chain-step, pipe and issue utilisation,
DRAM bytes per launch) must be what tools/fp64_work.py derives from the committed ncu metric lists (profiles/r02_ops_*.csv)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fp64_work   # noqa: E402


def test_fp64_work_json_follows_from_the_committed_metric_lists():
    committed = json.load(open(os.path.join(ROOT, "profiles", "fp64_work.json")))
    per_launch = {"dgauss": 10485760.0, "rosen16": 10485760.0, "gmix64": 1048576.0}     # chains x steps one launch advances
    n = 0
    for key, entry in committed.items():
        if key.startswith("_"):
            continue
        path = os.path.join(ROOT, entry["source"])
        assert os.path.exists(path), "%s names %s, which is not committed" % (key, entry["source"])
        again = fp64_work.one(path, per_launch[key.split("/")[0]], entry["source"])
        for k in ("fp64_flops_per_chain_step", "dadd", "dmul", "dfma", "fp64_pipe_pct", "issue_active_pct", "dram_bytes_per_launch", "launches"):
            assert again[k] == entry[k], (key, k, again[k], entry[k])
        n += 1
    assert n >= 8
    # the default bench line's key exists and carries a plausible count (SURVEY 8d guessed 280 for DualGaussian)
    assert 100 < committed["dgauss/reference/M16"]["fp64_flops_per_chain_step"] < 280
