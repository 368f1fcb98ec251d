#!/bin/bash
# round 2, GPU call C (1 GPU): full GPU suite (host-likelihood path, wide FPEPS fast path), KS diagnostic, launch-bound A/B
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 ) > $O/c_tests.log 2>&1
echo "pytest rc=$?" >> $O/c_tests.log
timeout 600 python tools/ks_diag.py > $O/c_ks.log 2>&1
S="--steps 100 --advance 200 --no-cpu --no-e2e --no-modes"
run() { tag=$1; shift; timeout 300 python bench.py $S "$@" > $O/c_$tag.json 2>> $O/c_err.log || echo "FAILED $tag" >> $O/c_err.log; }
run base_ref16; run base_sum16 --remote-mode summix; run base_sum256 --remote-mode summix --pool 256; run base_local --pl 1.0
for v in mx5 mx7 mx8; do
  export MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_$v.so
  run ${v}_ref16; run ${v}_sum16 --remote-mode summix; run ${v}_sum256 --remote-mode summix --pool 256
done
unset MCGPU_LIB
run base_r16_local --workload rosen16 --pl 1.0; run base_r16_sum16 --workload rosen16 --remote-mode summix; run base_r16_ref16 --workload rosen16
for v in wsm5 wsm6; do
  export MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_$v.so
  run ${v}_r16_local --workload rosen16 --pl 1.0; run ${v}_r16_sum16 --workload rosen16 --remote-mode summix
done
unset MCGPU_LIB
run base_g64_local --workload gmix64 --pl 1.0; run base_g64_sum256 --workload gmix64 --remote-mode summix --pool 256; run base_g64_ref16 --workload gmix64
for v in w64_2 w64_4; do
  export MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_$v.so
  run ${v}_g64_local --workload gmix64 --pl 1.0; run ${v}_g64_sum256 --workload gmix64 --remote-mode summix --pool 256
done
unset MCGPU_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c_*.json")):
    try:
        d = json.load(open(f)); print("%-28s %.4g  %.4f ms  fallback %.2e" % (f.split("/")[-1][2:-5], d["value"], d["ms_per_step"], d["exact_fallback_rate"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
grep -E "passed|failed" $O/c_tests.log | tail -2; grep -E "^FAILED|^ERROR" $O/c_tests.log | head
cat $O/c_ks.log | cut -c1-400
