#!/bin/bash
# round 2, GPU call F (1 GPU): the cooperative d = 64 kernel -- parity tests, then A/B against the wide kernel (MCGPU_NO_COOP=1)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/f_tests.log 2>&1
echo "pytest rc=$?" >> $O/f_tests.log
grep -E 'passed|failed' $O/f_tests.log | tail -2; grep -E '^FAILED|^ERROR' $O/f_tests.log | head -20
S="--workload gmix64 --steps 100 --advance 100 --no-cpu --no-e2e --no-modes"
run() { tag=$1; shift; timeout 300 python bench.py $S "$@" > $O/f_$tag.json 2>> $O/f_err.log || echo "FAILED $tag" >> $O/f_err.log; }
run coop_local --pl 1.0; run coop_sum256 --remote-mode summix --pool 256; run coop_sum16 --remote-mode summix --pool 16; run coop_ref16
export MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_coop1.so
run coop1_local --pl 1.0; run coop1_sum256 --remote-mode summix --pool 256
unset MCGPU_LIB
export MCGPU_NO_COOP=1
run wide_local --pl 1.0; run wide_sum256 --remote-mode summix --pool 256; run wide_sum16 --remote-mode summix --pool 16
unset MCGPU_NO_COOP
S="--steps 100 --advance 200 --no-cpu --no-e2e --no-modes"
run dg_local --pl 1.0; run dg_ref16; run dg_sum16 --remote-mode summix; run dg_sum256 --remote-mode summix --pool 256; run r2d_local --workload rosen2d --pl 1.0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/f_*.json")):
    try:
        d = json.load(open(f)); print("%-28s %.4g  %.4f ms  fallback %.2e  accept %.3f  mean %s" % (f.split("/")[-1][2:-5], d["value"], d["ms_per_step"], d["exact_fallback_rate"], d["accept_rate"], ["%.4f" % x for x in d["posterior_mean"][:3]]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -20 $O/f_err.log
