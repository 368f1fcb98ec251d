#!/bin/bash
# round 2, GPU call D (1 GPU): KS diagnostic over every stride-64 offset, issue-model micro-benchmark, default bench line,
# wide launch-bound A/B (d = 16), full GPU suite on the retuned build
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 ) > $O/d_tests.log 2>&1
echo "pytest rc=$?" >> $O/d_tests.log
timeout 900 python tools/ks_diag2.py > $O/d_ks2.log 2>&1
timeout 120 tools/micro/issue_model > $O/d_issue_model.txt 2>&1
( time timeout 900 python bench.py ) > $O/d_bench.json 2> $O/d_bench.err
echo "bench rc=$?" >> $O/d_bench.err
S="--steps 100 --advance 200 --no-cpu --no-e2e --no-modes"
run() { tag=$1; shift; timeout 300 python bench.py $S "$@" > $O/d_$tag.json 2>> $O/d_err.log || echo "FAILED $tag" >> $O/d_err.log; }
run base_r16_local --workload rosen16 --pl 1.0; run base_r16_sum16 --workload rosen16 --remote-mode summix; run base_r16_ref16 --workload rosen16
for v in wsm7 wsm8; do
  export MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_$v.so
  run ${v}_r16_local --workload rosen16 --pl 1.0; run ${v}_r16_sum16 --workload rosen16 --remote-mode summix; run ${v}_r16_ref16 --workload rosen16
done
unset MCGPU_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/d_*_r16_*.json")):
    try:
        d = json.load(open(f)); print("%-28s %.4g  %.4f ms  fallback %.2e" % (f.split("/")[-1][2:-5], d["value"], d["ms_per_step"], d["exact_fallback_rate"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
grep -E "passed|failed" $O/d_tests.log | tail -2; grep -E "^FAILED|^ERROR" $O/d_tests.log | head
cat $O/d_issue_model.txt
cat $O/d_ks2.log | cut -c1-700
