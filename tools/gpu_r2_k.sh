#!/bin/bash
# round 2, GPU call K (1 GPU): remote cooperative kernel with the pool read through L1 (two chains per owner warp) vs staged in
# shared memory; ncu --set full of the window kernels (fixed kernel-name filters); placement of the timed region; printed drift
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "gaussmix or gmix" ) > $O/k_tests_gmix.log 2>&1
echo "pytest rc=$?" >> $O/k_tests_gmix.log
grep -E 'passed|failed' $O/k_tests_gmix.log | tail -2; grep -E '^FAILED|^ERROR' $O/k_tests_gmix.log | head
S="--workload gmix64 --steps 200 --no-cpu --no-e2e --no-modes"
run() { tag=$1; shift; timeout 300 python bench.py "$@" > $O/k_$tag.json 2>> $O/k_err.log || echo "FAILED $tag" >> $O/k_err.log; }
run g64_l1pool_sum256 $S --remote-mode summix --pool 256; run g64_l1pool_sum16 $S --remote-mode summix --pool 16
export MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_poolsm.so
run g64_smpool_sum256 $S --remote-mode summix --pool 256; run g64_smpool_sum16 $S --remote-mode summix --pool 16
unset MCGPU_LIB
D="--no-cpu --no-e2e --no-modes"
run dg_20 $D --steps 20; run dg_50 $D --steps 50 --warmup 5; run dg_1000 $D --steps 1000
run dg_sum_20 $D --steps 20 --remote-mode summix; run dg_sum_1000 $D --steps 1000 --remote-mode summix
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/k_*.json")):
    try:
        d = json.load(open(f)); print("%-28s %.4g  %.4f ms  fallback %.2e  remote %.4f  iters %.1f" % (f.split("/")[-1][2:-5], d["value"], d["ms_per_step"], d["exact_fallback_rate"], d["remote_fraction"], d["remote_iterations_mean"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -20 $O/k_err.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "stationarity" > $O/k_stationarity.log 2>&1
grep -E "^pl=|remote mode|passed|failed" $O/k_stationarity.log
B="python bench.py --no-cpu --no-e2e --no-modes --advance 300 --steps 60 --no-place"
G="python bench.py --workload gmix64 --no-cpu --no-e2e --no-modes --advance 60 --steps 40 --no-place"
full() { tag=$1; kn=$2; skip=$3; div=$4; shift 4
  "$@" > $O/k_plain_$tag.json 2>> $O/k_err.log &&
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$kn" -s $skip -c 3 -o /tmp/prof_$tag -f "$@" > $O/k_ncu_full_$tag.log 2>&1
  python tools/summarize_profile.py full /tmp/prof_$tag.ncu-rep $div $div $div > $O/r02_full_$tag.txt 2>> $O/k_err.log
  python tools/profile_lines.py /tmp/prof_$tag.ncu-rep 0 $div 70 > $O/r02_lines_$tag.txt 2>> $O/k_err.log; }
full ref16 "mh_steps_kernel<[^>]*1>" 200 327680 $B --remote-mode reference --pool 16
full sum16 "mh_steps_kernel<[^>]*4>" 200 327680 $B --remote-mode summix --pool 16
full g64local "mh_coop_kernel<[^>]*4, [^>]*2, " 30 32768 $G --pl 1.0
full g64sum256 "mh_coop_kernel<[^>]*5, " 3 65536 $G --remote-mode summix --pool 256
ls -la $O/r02_full_*.txt | awk '{print $5, $9}'
head -24 $O/r02_full_g64sum256.txt
