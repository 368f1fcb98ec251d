#!/bin/bash
# round 2, GPU call H (1 GPU): warp-specialised cooperative d = 64 kernel -- parity tests first (under a short timeout: named barriers),
# then A/B of chains per owner warp against the wide kernel, then ncu of both roles
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "gaussmix or gmix" ) > $O/h_tests_gmix.log 2>&1
echo "pytest rc=$?" >> $O/h_tests_gmix.log
grep -E 'passed|failed' $O/h_tests_gmix.log | tail -2; grep -E '^FAILED|^ERROR' $O/h_tests_gmix.log | head
if grep -q "failed\|rc=124" $O/h_tests_gmix.log; then tail -40 $O/h_tests_gmix.log; fi
S="--workload gmix64 --steps 100 --advance 100 --no-cpu --no-e2e --no-modes"
run() { tag=$1; shift; timeout 300 python bench.py $S "$@" > $O/h_$tag.json 2>> $O/h_err.log || echo "FAILED $tag" >> $O/h_err.log; }
run coop2_local --pl 1.0; run coop2_sum256 --remote-mode summix --pool 256; run coop2_sum16 --remote-mode summix --pool 16; run coop2_ref16
export MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_coop1.so
run coop1_local --pl 1.0; run coop1_sum256 --remote-mode summix --pool 256
export MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_coop4.so
run coop4_local --pl 1.0; run coop4_sum256 --remote-mode summix --pool 256
unset MCGPU_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/h_*.json")):
    try:
        d = json.load(open(f)); print("%-28s %.4g  %.4f ms  fallback %.2e  accept %.3f  mean %s" % (f.split("/")[-1][2:-5], d["value"], d["ms_per_step"], d["exact_fallback_rate"], d["accept_rate"], ["%.4f" % x for x in d["posterior_mean"][:3]]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -20 $O/h_err.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/h_tests.log 2>&1
echo "pytest rc=$?" >> $O/h_tests.log
grep -E 'passed|failed' $O/h_tests.log | tail -2; grep -E '^FAILED|^ERROR' $O/h_tests.log | head -20
B="python bench.py --workload gmix64 --no-cpu --no-e2e --no-modes --advance 20 --steps 20"
for cfg in "g64local:--pl 1.0:30" "g64sum256:--remote-mode summix --pool 256 --pl 0.0:4"; do
  tag=${cfg%%:*}; rest=${cfg#*:}; fl=${rest%%:*}; skip=${rest#*:}
  $B $fl > $O/h_plain_$tag.json 2>> $O/h_err.log &&
  ncu --set full --clock-control none --import-source on -k regex:mh_coop -s $skip -c 2 -o /tmp/prof_$tag -f $B $fl > $O/h_ncu_$tag.log 2>&1
  python tools/summarize_profile.py full /tmp/prof_$tag.ncu-rep 65536 65536 > $O/r02_full_$tag.txt 2>> $O/h_err.log
  python tools/profile_lines.py /tmp/prof_$tag.ncu-rep 0 65536 120 > $O/r02_lines_$tag.txt 2>> $O/h_err.log
  ncu -i /tmp/prof_$tag.ncu-rep --page details --csv --launch-count 1 > $O/r02_details_$tag.csv 2>> $O/h_err.log
done
head -24 $O/r02_full_g64local.txt; head -24 $O/r02_full_g64sum256.txt
